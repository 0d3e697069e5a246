"""torch.profiler breakdown of the split-sum training step and one relighting frame: every CUDA kernel (ours and
torch's glue) by total device time and launch count.  Run on the GPU box:
    PYTHONPATH=. python scripts/profile_split.py [train|relight] > gpurun_out/profile_split.txt"""
import sys

import torch
from torch.profiler import ProfilerActivity, profile

from rise_sdf_b200 import synthetic as syn
from rise_sdf_b200.split_mixed_occ import SplitMixedOCCModel, split_mixed_occ_config


def table(prof, n_iter, top=45):
    rows = {}
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            r = rows.setdefault(e.name[:110], [0, 0.0])
            r[0] += 1
            r[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    tot = sum(v[1] for v in rows.values())
    print(f"device time {tot / 1e3 / n_iter:.2f} ms/iter over {sum(v[0] for v in rows.values()) / n_iter:.0f} launches/iter")
    for k, (c, t) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{t / 1e3 / n_iter:9.3f} ms {c / n_iter:7.1f} x  {k}")


def train():
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
    n = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        trainer.step(*batch, update=False)
    e1.record()
    torch.cuda.synchronize()
    print(f"split train step (no occupancy update) {e0.elapsed_time(e1) / n:.2f} ms wall")
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            trainer.step(*batch, update=False)
        torch.cuda.synchronize()
    table(prof, n)


def relight(variance=0.5):
    from rise_sdf_b200.relight import EnvSet, render_frame_shard, synthetic_envs
    dev = torch.device("cuda:0")
    torch.manual_seed(42)
    model = SplitMixedOCCModel(split_mixed_occ_config()).to(dev)
    with torch.no_grad():
        model.geometry.network.layers[0].weight_v[:, 3:].normal_(0.0, 0.05)
        model.variance.variance.fill_(variance)
    model.train()
    model.update_step(0, 80000)
    gj = torch.Generator().manual_seed(7)
    model.occupancy_grid._update(0, model.occ_eval_fn, occ_thre=0.001, jitter=torch.rand(128 ** 3, 3, generator=gj))
    model.eval()
    model.background_color = torch.ones(3, device=dev)
    envs = EnvSet(model, synthetic_envs())
    poses, dirs = syn.camera_poses(), syn.ray_directions()
    rays = syn.frame_rays(3, poses, dirs).to(dev)
    render_frame_shard(model, rays, envs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    render_frame_shard(model, rays, envs)
    e1.record()
    torch.cuda.synchronize()
    print(f"relit frame x{len(envs.maps)} maps: {e0.elapsed_time(e1):.1f} ms wall")
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        render_frame_shard(model, rays, envs)
        torch.cuda.synchronize()
    table(prof, 1)


if __name__ == "__main__":
    what = sys.argv[1:] or ["train", "relight"]
    if "train" in what:
        train()
    if "relight" in what:
        relight()
