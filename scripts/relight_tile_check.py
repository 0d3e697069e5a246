"""Relit frame at two tile sizes (variance given on the command line): identical pixels? largest evaluation batch?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rise_sdf_b200 import synthetic as syn, nerfacc, tinycudann
from rise_sdf_b200.relight import EnvSet, render_frame_shard, synthetic_envs
from rise_sdf_b200.split_mixed_occ import SplitMixedOCCModel, split_mixed_occ_config
var = float(sys.argv[1]) if len(sys.argv) > 1 else 0.3
dev = torch.device("cuda:0")
torch.manual_seed(42)
model = SplitMixedOCCModel(split_mixed_occ_config()).to(dev)
with torch.no_grad():
    model.geometry.network.layers[0].weight_v[:, 3:].normal_(0.0, 0.05)
    model.variance.variance.fill_(var)
model.train()
model.update_step(0, 80000)
gj = torch.Generator().manual_seed(7)
model.occupancy_grid._update(0, model.occ_eval_fn, occ_thre=0.001, jitter=torch.rand(128 ** 3, 3, generator=gj))
model.eval()
model.background_color = torch.ones(3, device=dev)
envs = EnvSet(model, synthetic_envs())
poses, dirs = syn.camera_poses(), syn.ray_directions()
rays = syn.frame_rays(3, poses, dirs).to(dev)
biggest = [0]
orig = tinycudann.hashgrid_fd6
def spy(inner, points, eps, radius):
    biggest[0] = max(biggest[0], points.shape[0])
    return orig(inner, points, eps, radius)
tinycudann.hashgrid_fd6 = spy
res = {}
for tile in (65536, 131072):
    biggest[0] = 0
    out, _ = render_frame_shard(model, rays, envs, tile=tile, keys=("comp_rgb_phys_full", "depth"))
    torch.cuda.synchronize()
    res[tile] = out
    print(f"tile {tile}: largest FD batch {biggest[0]} samples = {biggest[0] * 6 * 32 / 2**31:.2f} x 2^31 feature elements; "
          f"reserved {torch.cuda.memory_reserved() / 2**30:.1f} GB")
for e in (0, 1):
    for k in ("comp_rgb_phys_full", "depth"):
        d = (res[65536][e][k] - res[131072][e][k]).abs()
        print("env", e, k, "pixels differing", int((d.reshape(d.shape[0], -1).max(-1).values > 0).sum()), "max", float(d.max()))
