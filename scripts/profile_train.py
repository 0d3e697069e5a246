import sys, torch
sys.path.insert(0, '/root/repo')
from rise_sdf_b200 import synthetic as syn
from rise_sdf_b200.neus import NeuSModel, neus_blender_config
from rise_sdf_b200.train import NeusTrainer
dev = torch.device('cuda'); torch.manual_seed(42)
model = NeuSModel(neus_blender_config()).to(dev).train()
with torch.no_grad(): model.geometry.network.layers[0].weight_v[:, 3:].normal_(0.0, 0.05)
tr = NeusTrainer(model); model.cos_anneal_ratio = 0.0
model.occupancy_grid._update(0, model.occ_eval_fn, occ_thre=0.001, jitter=torch.rand(128**3, 3, generator=torch.Generator().manual_seed(7)))
rays, rgb, fg, bg = [t.to(dev) for t in syn.training_rays(8192, seed=42)]
for _ in range(3): tr.step(rays, rgb, fg, bg)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.step(rays, rgb, fg, bg); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=int(sys.argv[1]) if len(sys.argv) > 1 else 40, max_name_column_width=70))
