import sys, time, torch
sys.path.insert(0, '/root/repo')
from rise_sdf_b200.network_utils import VanillaMLP
from rise_sdf_b200.fused_mlp import PackedMLP
torch.manual_seed(0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
for name, (di, do, hid, sphere) in {"geo": (35, 48, 2, True), "tex": (67, 3, 4, False)}.items():
    m = VanillaMLP(di, do, {"n_neurons": 128, "n_hidden_layers": hid, "sphere_init": sphere, "weight_norm": sphere, "output_activation": "none"}).cuda()
    x = torch.rand(S, di, device="cuda")
    f = PackedMLP(m)
    for _ in range(3): y = f(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): y = f(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    flops = 2 * S * sum(a * b for a, b in zip([di] + [128] * hid, [128] * hid + [do]))
    print(f"{name}: S={S} {ms:.3f} ms  {S/ms/1e3:.1f} Mrows/s  {flops/ms/1e9:.1f} TFLOP/s (fp32-equivalent), x3 MMA = {3*flops/ms/1e9:.1f}")
    with torch.no_grad():
        e0.record(); 
        for _ in range(5): y2 = m.layers(x)
        e1.record(); torch.cuda.synchronize()
    print(f"   torch fp32 (cuBLAS): {e0.elapsed_time(e1)/5:.3f} ms")
