"""Host-side op counts and device kernels of ONE relighting tile (32 000 rays, both env maps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from rise_sdf_b200 import synthetic as syn
from rise_sdf_b200.relight import EnvSet, render_frame_shard, synthetic_envs
from rise_sdf_b200.split_mixed_occ import SplitMixedOCCModel, split_mixed_occ_config
dev = torch.device("cuda:0")
torch.manual_seed(42)
model = SplitMixedOCCModel(split_mixed_occ_config()).to(dev)
with torch.no_grad():
    model.geometry.network.layers[0].weight_v[:, 3:].normal_(0.0, 0.05)
    model.variance.variance.fill_(0.5)
model.train()
model.update_step(0, 80000)
gj = torch.Generator().manual_seed(7)
model.occupancy_grid._update(0, model.occ_eval_fn, occ_thre=0.001, jitter=torch.rand(128 ** 3, 3, generator=gj))
model.eval()
model.background_color = torch.ones(3, device=dev)
envs = EnvSet(model, synthetic_envs())
poses, dirs = syn.camera_poses(), syn.ray_directions()
rays = syn.frame_rays(3, poses, dirs).to(dev)
tile = rays[304000:336000].contiguous()          # a centre tile
for _ in range(2):
    render_frame_shard(model, tile, envs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); render_frame_shard(model, tile, envs); e1.record(); torch.cuda.synchronize()
print(f"centre tile, 2 maps: {e0.elapsed_time(e1):.2f} ms")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    render_frame_shard(model, tile, envs)
    torch.cuda.synchronize()
rows = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        r = rows.setdefault(e.name[:100], [0, 0.0]); r[0] += 1; r[1] += e.device_time
print(f"device {sum(v[1] for v in rows.values()) / 1e3:.2f} ms over {sum(v[0] for v in rows.values())} launches")
for k, (c, t) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{t / 1e3:8.3f} ms {c:4d} x {k}")
print("== device kernels by launch count")
for k, (c, t) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{c:4d} x {t / 1e3:8.3f} ms {k}")
print("== host")
for e in sorted(prof.key_averages(), key=lambda e: -e.self_cpu_time_total)[:45]:
    print(f"{e.self_cpu_time_total / 1e3:8.2f} ms  {e.count:5d} x  {e.key[:90]}")
