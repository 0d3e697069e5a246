"""Per-step wall time of the split-train e2e loop (find the outlier steps)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rise_sdf_b200 import synthetic as syn
from rise_sdf_b200.split_mixed_occ import SplitMixedOCCModel, split_mixed_occ_config
from rise_sdf_b200.train import SplitTrainer
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
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
poses, dirs = syn.camera_poses(), syn.ray_directions()
host = [tuple(t.pin_memory() for t in syn.training_rays(4096, seed=7 + 1000 * b, rank=0, poses=poses, directions=dirs)) for b in range(2)]
batches = [tuple(t.to(dev) for t in h) for h in host]
rs = model.render_step_size
model.render_step_size = rs / 1.1
trainer.step(*batches[0], update=False)
model.render_step_size = rs
for i in range(5):
    trainer.step(*batches[i % 2])
torch.cuda.synchronize()
for rep in range(4):
    sync = rep % 2 == 0
    ts, evs = [], [torch.cuda.Event(enable_timing=True) for _ in range(17)]
    torch.cuda.synchronize()
    evs[0].record()
    for i in range(16):
        t0 = time.perf_counter()
        if sync:
            b = tuple(t.to(dev, non_blocking=True) for t in host[i % 2])
        else:
            b = batches[i % 2]
        loss, out = trainer.step(*b)
        if sync:
            float(loss.item())
        evs[i + 1].record()
        ts.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
    dev_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(16)]
    print("sync " if sync else "async", "host:", " ".join(f"{t:5.0f}" for t in ts))
    print("      ", "dev: ", " ".join(f"{t:5.0f}" for t in dev_ms), " step", trainer.global_step, " reserved GB",
          round(torch.cuda.memory_reserved() / 2**30, 1), "segments", torch.cuda.memory_stats()["num_device_alloc"],
          "samples", int(out["num_samples"].sum()))
