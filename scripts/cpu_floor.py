"""Host-side cost of one training step: the same step on a tiny ray batch (GPU work negligible)."""
import sys, time, torch
sys.path.insert(0, '/root/repo')
from rise_sdf_b200 import synthetic as syn
from rise_sdf_b200.neus import NeuSModel, neus_blender_config
from rise_sdf_b200.train import NeusTrainer
dev = torch.device('cuda'); torch.manual_seed(42)
model = NeuSModel(neus_blender_config()).to(dev).train()
tr = NeusTrainer(model); model.cos_anneal_ratio = 0.0
model.occupancy_grid._update(0, model.occ_eval_fn, occ_thre=0.001, jitter=torch.rand(128**3, 3, generator=torch.Generator().manual_seed(7)))
for n in (64, 8192):
    b = [t.to(dev) for t in syn.training_rays(n, seed=42)]
    for _ in range(5): tr.step(*b)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(20): tr.step(*b)
    torch.cuda.synchronize(); print(n, 'rays: ms/step', (time.perf_counter() - t) / 20 * 1e3)
import cProfile, pstats
b = [t.to(dev) for t in syn.training_rays(64, seed=42)]
pr = cProfile.Profile(); pr.enable()
for _ in range(10): tr.step(*b)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
