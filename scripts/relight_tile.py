"""One shared relighting tile (32 000 rays of the frame centre, both env maps) inside an NVTX range, for
  ncu --nvtx --nvtx-include "tile/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python scripts/relight_tile.py
(the launch list committed under profiles/)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from rise_sdf_b200 import synthetic as syn
from rise_sdf_b200.relight import EnvSet, render_frame_shard, synthetic_envs
from rise_sdf_b200.split_mixed_occ import SplitMixedOCCModel, split_mixed_occ_config
dev = torch.device('cuda')
torch.manual_seed(42)
model = SplitMixedOCCModel(split_mixed_occ_config()).to(dev)
with torch.no_grad():
    model.geometry.network.layers[0].weight_v[:, 3:].normal_(0.0, 0.05); model.variance.variance.fill_(0.5)
model.train(); model.update_step(0, 80000)
model.occupancy_grid._update(0, model.occ_eval_fn, occ_thre=0.001, jitter=torch.rand(128**3, 3, generator=torch.Generator().manual_seed(7)))
model.eval(); model.background_color = torch.ones(3, device=dev)
envs = EnvSet(model, synthetic_envs())
tile = syn.frame_rays(3).to(dev)[320000 - 16000:320000 + 16000].contiguous()
render_frame_shard(model, tile, envs, tile=32000)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("tile")
out, _ = render_frame_shard(model, tile, envs, tile=32000)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("mean rgb", float(out[0]["comp_rgb_phys_full"].mean()), float(out[1]["comp_rgb_phys_full"].mean()))
