"""Split-sum training step: device time vs host enqueue time, with and without a per-step host sync."""
import time
import torch
from rise_sdf_b200 import synthetic as syn
from rise_sdf_b200.split_mixed_occ import SplitMixedOCCModel, split_mixed_occ_config
from rise_sdf_b200.train import SplitTrainer

dev = torch.device("cuda:0")
torch.manual_seed(42)
model = SplitMixedOCCModel(split_mixed_occ_config()).to(dev)
with torch.no_grad():
    model.geometry.network.layers[0].weight_v[:, 3:].normal_(0.0, 0.05)
model.train()
trainer = SplitTrainer(model)
trainer.global_step = 20001
model.update_step(0, 20000)
gj = torch.Generator().manual_seed(7)
model.occupancy_grid._update(0, model.occ_eval_fn, occ_thre=0.001, jitter=torch.rand(128 ** 3, 3, generator=gj))
batch = tuple(t.to(dev) for t in syn.training_rays(4096, seed=7))
for _ in range(4):
    trainer.step(*batch, update=False)
torch.cuda.synchronize()
for mode in ("async", "sync-each-step", "async"):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 6
    cpu = 0.0
    for _ in range(n):
        c0 = time.perf_counter()
        loss, _ = trainer.step(*batch, update=False)
        cpu += time.perf_counter() - c0
        if mode != "async":
            float(loss.item())
    torch.cuda.synchronize()
    print(f"{mode:16s} wall {1e3 * (time.perf_counter() - t0) / n:7.2f} ms/step   host time inside step() {1e3 * cpu / n:7.2f} ms")
