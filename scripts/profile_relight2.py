"""Wall / CUDA time of the two passes of a shared-tile relighting render (first env map: everything; second env
map: emitter lookups + compositing on the tile cache), for an average and for the densest tile."""
import sys, time, torch
sys.path.insert(0, '/root/repo')
from rise_sdf_b200 import synthetic as syn
from rise_sdf_b200.relight import EnvSet, synthetic_envs
from rise_sdf_b200.split_mixed_occ import SplitMixedOCCModel, split_mixed_occ_config
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda')
torch.manual_seed(42)
model = SplitMixedOCCModel(split_mixed_occ_config()).to(dev)
with torch.no_grad():
    model.geometry.network.layers[0].weight_v[:, 3:].normal_(0.0, 0.05); model.variance.variance.fill_(0.5)
model.train(); model.update_step(0, 80000)
model.occupancy_grid._update(0, model.occ_eval_fn, occ_thre=0.001, jitter=torch.rand(128**3, 3, generator=torch.Generator().manual_seed(7)))
model.eval(); model.background_color = torch.ones(3, device=dev)
envs = EnvSet(model, synthetic_envs())
rays = syn.frame_rays(3).to(dev)


def two_passes(tile, prof=False):
    res = []
    model._tile_cache = {}
    for e in range(2):
        envs.use(e)
        torch.cuda.synchronize(); t = time.time()
        if prof:
            with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as p:
                model.forward_(tile, relighting=True); torch.cuda.synchronize()
            res.append(p)
        else:
            model.forward_(tile, relighting=True); torch.cuda.synchronize()
            res.append(time.time() - t)
    model._tile_cache = None
    return res


with torch.no_grad():
    for name, a in (("centre tile", 320000 - 16000), ("upper tile", 96000)):
        tile = rays[a:a + 32000].contiguous()
        two_passes(tile); two_passes(tile)
        w = two_passes(tile)
        print(f"{name}: pass 1 {w[0]*1e3:.1f} ms, pass 2 {w[1]*1e3:.1f} ms")
    p1, p2 = two_passes(rays[320000 - 16000:320000 + 16000].contiguous(), prof=True)
print("==== pass 2 by CUDA time")
print(p2.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
print("==== pass 2 by CPU time")
print(p2.key_averages().table(sort_by="self_cpu_time_total", row_limit=22, max_name_column_width=70))
