"""Developer tool: per-phase clock64 breakdown of the fused SDF kernels (thread 0 of CTA 0).
Build: nvcc ... -DRSDF_PROFILE_PHASES -shared -o scripts/_dbg/libsdfprof.so csrc/sdf_train.cu csrc/mlp_tc.cu"""
import ctypes, sys
import torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from rise_sdf_b200 import _lib as L, sdf_field
from test_gpu_sdf_field import make_mlp
lib = ctypes.CDLL('/root/repo/scripts/_dbg/libsdfprof.so')
S = 3340000
m = make_mlp()
x01 = torch.rand(S, 3, device='cuda'); enc = torch.randn(S, 32, device='cuda') * 0.1
(W1, b1), (W2, b2), (W3, b3) = [(w.detach(), b.detach()) for w, b in m.effective_weights()]
net, keep = sdf_field._net_struct(W1, b1, W2, b2, W3, b3)
out = torch.empty(S, 48, device='cuda'); g0a = torch.empty(S, 3, device='cuda'); g0b = torch.empty(S, 32, device='cuda')
go = torch.randn(S, 48, device='cuda'); gg = torch.randn(S, 35, device='cuda'); gga = gg[:, :3].contiguous(); ggb = gg[:, 3:].contiguous(); gin0 = torch.empty(S, 3, device='cuda'); gin1 = torch.empty(S, 32, device='cuda')
gs = [torch.zeros_like(t) for t in (W1, b1, W2, b2, W3, b3)]
amax = torch.empty(1, device='cuda', dtype=torch.int32)
L.call("rsdf_absmax2", L.ptr(go), go.numel(), L.ptr(gg), gg.numel(), L.ptr(amax), 0, L.stream())
c_p, c_i, c_f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
lib.rsdf_sdf_mlp_fwd.argtypes = [c_p, c_p, c_i, c_f, c_f, c_p, c_i, c_i, c_p, c_p, c_p, c_p, c_p]
lib.rsdf_sdf_mlp_bwd.argtypes = [c_p, c_p, c_i, c_f, c_f, c_p, c_i, c_i] + [c_p] * 14
buf = (ctypes.c_ulonglong * 8)()
def show(name):
    torch.cuda.synchronize(); lib.rsdf_debug_read_prof(buf)
    v = list(buf)[:4]; tot = sum(v); tiles = (S + 63) // 64 // 148 + 1
    print(f"{name}: total {tot / 1.965e3 / tiles:.2f} us/tile  " + "  ".join(f"{n} {x / 1.965e3 / tiles:.2f}" for n, x in zip(("epilogue", "fence+sync", "issue", "mma_wait"), v)))
for _ in range(2):
    lib.rsdf_sdf_mlp_fwd(ctypes.byref(net), x01.data_ptr(), 3, 2.0, -1.0, enc.data_ptr(), 32, S, out.data_ptr(), None, g0a.data_ptr(), g0b.data_ptr(), None)
show("fwd+g0")
for _ in range(2):
    lib.rsdf_sdf_mlp_bwd(ctypes.byref(net), x01.data_ptr(), 3, 2.0, -1.0, enc.data_ptr(), 32, S, go.data_ptr(), None, gga.data_ptr(), ggb.data_ptr(), amax.data_ptr(), gin0.data_ptr(), gin1.data_ptr(),
                         *[g.data_ptr() for g in gs], None)
show("bwd")
