import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from rise_sdf_b200 import relu_mlp
from test_gpu_relu_mlp import make
S = int(sys.argv[1]) if len(sys.argv) > 1 else 3340000
m = make(67, 3, 4)
x = torch.randn(S, 67, device='cuda', requires_grad=True)
def step():
    out = relu_mlp.relu_mlp(m, x)
    out.backward(torch.ones_like(out) / S)
for _ in range(2): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
