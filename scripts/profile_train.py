"""torch.profiler view of one neus-blender training step: device kernels by time, host ops by self time, syncs."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from rise_sdf_b200 import synthetic as syn
from rise_sdf_b200.neus import NeuSModel, neus_blender_config
from rise_sdf_b200.train import NeusTrainer
dev = torch.device('cuda'); torch.manual_seed(42)
model = NeuSModel(neus_blender_config()).to(dev).train()
with torch.no_grad(): model.geometry.network.layers[0].weight_v[:, 3:].normal_(0.0, 0.05)
tr = NeusTrainer(model); tr.global_step = 5001
model.update_step(0, 5000)
model.occupancy_grid._update(0, model.occ_eval_fn, occ_thre=0.001, jitter=torch.rand(128**3, 3, generator=torch.Generator().manual_seed(7)))
rays, rgb, fg, bg = [t.to(dev) for t in syn.training_rays(8192, seed=42)]
for _ in range(4): tr.step(rays, rgb, fg, bg, update=False)
torch.cuda.synchronize()
for mode in ("async", "sync"):
    t0 = time.perf_counter(); cpu = 0.0
    for _ in range(6):
        c0 = time.perf_counter()
        loss, _ = tr.step(rays, rgb, fg, bg, update=False)
        cpu += time.perf_counter() - c0
        if mode == "sync": float(loss.item())
    torch.cuda.synchronize()
    print(f"{mode}: wall {(time.perf_counter() - t0) / 6 * 1e3:.2f} ms/step, host inside step {cpu / 6 * 1e3:.2f} ms")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.step(rays, rgb, fg, bg, update=False); torch.cuda.synchronize()
rows = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        r = rows.setdefault(e.name[:100], [0, 0.0]); r[0] += 1; r[1] += e.device_time
print(f"device {sum(v[1] for v in rows.values()) / 1e3:.2f} ms over {sum(v[0] for v in rows.values())} launches")
for k, (c, t) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:32]:
    print(f"{t / 1e3:8.3f} ms {c:4d} x {k}")
print("== host")
for e in sorted(prof.key_averages(), key=lambda e: -e.self_cpu_time_total)[:40]:
    print(f"{e.self_cpu_time_total / 1e3:8.2f} ms  {e.count:5d} x  {e.key[:90]}")
