import sys, time, torch
sys.path.insert(0, '/root/repo')
from rise_sdf_b200 import synthetic as syn
from rise_sdf_b200.split_mixed_occ import SplitMixedOCCModel, split_mixed_occ_config
from rise_sdf_b200.train import SplitTrainer
dev = torch.device('cuda'); torch.manual_seed(42)
model = SplitMixedOCCModel(split_mixed_occ_config()).to(dev)
with torch.no_grad(): model.geometry.network.layers[0].weight_v[:, 3:].normal_(0.0, 0.05)
model.train(); model.update_step(0, 20000)
model.occupancy_grid._update(0, model.occ_eval_fn, occ_thre=0.001, jitter=torch.rand(128**3, 3, generator=torch.Generator().manual_seed(7)))
tr = SplitTrainer(model)
b = tuple(t.to(dev) for t in syn.training_rays(4096, seed=7))
for _ in range(3): loss, out = tr.step(*b)
torch.cuda.synchronize(); t = time.time()
for _ in range(3): loss, out = tr.step(*b)
torch.cuda.synchronize(); print('ms/step', (time.time() - t) / 3 * 1e3, 'loss', float(loss), 'samples', int(out['num_samples'].sum()))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.step(*b); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by=sys.argv[2] if len(sys.argv) > 2 else "cuda_time_total",
                                row_limit=int(sys.argv[1]) if len(sys.argv) > 1 else 40, max_name_column_width=70))
