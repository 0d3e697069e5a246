"""Hash-grid kernel microbench on ray-ordered sample points (the access pattern of a training step).
python scripts/bench_hashgrid.py [n_rays]"""
import sys, time
import torch
sys.path.insert(0, '/root/repo')
from rise_sdf_b200 import _lib as L, synthetic as syn
from rise_sdf_b200.tinycudann import Encoding

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device('cuda')
enc = Encoding(3, dict(otype="HashGrid", n_levels=16, n_features_per_level=2, log2_hashmap_size=19,
                       base_resolution=16, per_level_scale=1.447269237440378)).to(dev)
rays = syn.training_rays(R, seed=42)[0].to(dev)
n_per = 408
t = torch.linspace(2.6, 5.4, n_per, device=dev)
pts = rays[:, None, :3] + rays[:, None, 3:6] * t[None, :, None]
x = ((pts.reshape(-1, 3) + 1.5) / 3.0).clamp(0, 1).contiguous()
S = x.shape[0]
meta, table = enc.meta, enc.params.detach()
y = torch.empty(S, 32, device=dev); dy = torch.empty(S, 32, 3, device=dev)
gy = torch.randn(S, 32, device=dev); v = torch.randn(S, 3, device=dev)
gt = torch.zeros_like(table); ggy = torch.empty_like(gy); gx = torch.empty(S, 3, device=dev)
def timeit(name, fn, bytes_per_sample, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:28s} {ms:7.3f} ms  {S / ms / 1e6:7.2f} Gsamples/s  {bytes_per_sample * S / ms / 1e6:8.1f} GB/s algorithmic "
          f"({bytes_per_sample * S / ms / 1e6 / 6553:.3f} of HBM peak)")
st = L.stream()
print("samples", S)
timeit("fwd+dydx", lambda: L.call("rsdf_hashgrid_fwd", L.ptr(x), L.ptr(table), meta.ref, S, L.ptr(y), L.ptr(dy), st), 1548)
timeit("fwd", lambda: L.call("rsdf_hashgrid_fwd", L.ptr(x), L.ptr(table), meta.ref, S, L.ptr(y), None, st), 1164)
timeit("bwd_table", lambda: L.call("rsdf_hashgrid_bwd_table", L.ptr(x), L.ptr(gy), meta.ref, S, L.ptr(gt), st), 2188)
timeit("bwd_input", lambda: L.call("rsdf_hashgrid_bwd_input", L.ptr(dy), L.ptr(gy), S, 32, L.ptr(gx), st), 524)
timeit("bwd_bwd(table,dLdy)", lambda: L.call("rsdf_hashgrid_bwd_bwd", L.ptr(x), L.ptr(table), L.ptr(v), L.ptr(gy), meta.ref, S,
                                             L.ptr(gt), L.ptr(ggy), None, st), 2188)
