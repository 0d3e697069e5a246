#!/usr/bin/env python
"""bench.py -- headline benchmark of the ray-marched neural-SDF hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload train|relight]

Workload at N=1 (BASELINE.json configs[1]): one neus-blender TRAINING STEP -- 8192 synthetic
rays, occupancy-grid march, hash-grid SDF with analytic normals, eikonal loss through the
second-order hash-grid gradient, radiance MLP, fused alpha/scan/accumulate, all four losses,
backward, (NCCL all-reduce of one flat gradient bucket when N>1) and the Adam update.
`value` = whole-job train rays/s with the batch already resident in HBM; `e2e` = the same step
through the public API with HOST (pinned) rays/targets copied in and the loss read back inside
the timed region.  Prints ONE JSON line (rank 0).

--impl reference: the reference has no CPU path and its CUDA dependencies (tiny-cuda-nn,
nerfacc 0.5.3) cannot be installed offline, so the reference arm times the CPU restatement of
the same step (oracle/, kind "port") on all host cores, on a bounded ray sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# dynamic sample counts: see rise_sdf_b200/__init__.py (must be set before torch's first CUDA allocation)
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")

N_RAYS = 8192
METRIC = "train_rays_per_s"
WORKLOAD = ("neus-blender training step, 8192 rays/GPU/step, occupancy grid from random-init SDF refreshed every 16 steps, "
            "analytic normals + eikonal (2nd-order hash grid), update_step+fwd+loss+bwd+Adam (configs[1])")
# ALGORITHMIC work per sample and per launch of every hand-written kernel on the step.
#   hbm kernels: bytes/sample;  tensor kernels: fp32-equivalent flops/sample (2*in*out per product; the
#   3-product fp16 split that delivers fp32-class accuracy is NOT counted three times).
_GEO = (35, 128, 128, 48)
_F_FWD = 2 * (35 * 128 + 128 * 128 + 128 * 48)
_F_CHAIN = 2 * (128 * 128 + 128 * 35)                      # g0 = W1^T (s1 . (W2^T (s2 . w3)))
_F_BWD = (2 * (35 * 128 + 128 * 128) + _F_CHAIN            # recomputed forward + chain
          + 2 * (35 * 128 + 128 * 128 + 48 * 128 + 128 * 128 + 128 * 35)     # data gradients
          + 2 * (2 * 35 * 128 + 2 * 128 * 128 + 48 * 128))                   # weight gradients
_F_RAD = 2 * (67 * 128 + 3 * 128 * 128 + 128 * 3)             # radiance net forward
KERNEL_WORK = {
    # hash-grid forward with fused dy/dx: 12 (x) + 16*8*2*4 (corner reads) + 16*2*4 (y) + 3*16*2*4 (dy_dx)
    "rsdf_hashgrid_fwd": ("hbm", 12 + 1024 + 128 + 384, "hashgrid_fwd_kernel<true>"),
    "rsdf_hashgrid_bwd_table": ("hbm", 12 + 128 + 2048, "hashgrid_bwd_table_kernel"),
    "rsdf_hashgrid_bwd_bwd": ("hbm", 12 + 12 + 128 + 2048 + 1024 + 128, "hashgrid_bwd_bwd_kernel<table,dLdy>"),
    "rsdf_hashgrid_bwd_input": ("hbm", 384 + 128 + 12, "hashgrid_bwd_input_kernel"),
    # first + second order table scatter in one pass: x, v, dL_dy, g2, one atomic RMW per corner
    "rsdf_hashgrid_bwd_table2": ("hbm", 12 + 12 + 128 + 128 + 2048, "hashgrid_bwd_table2_kernel"),
    "rsdf_hashgrid_jvp": ("hbm", 384 + 12 + 128, "hashgrid_jvp_kernel"),
    "rsdf_sdf_mlp_fwd": ("tensor", _F_FWD + _F_CHAIN, "sdf_fwd_kernel<true, true>"),
    "rsdf_sdf_mlp_bwd": ("tensor", _F_BWD, "sdf_bwd_kernel<true, true>"),
    # radiance MLP 67 -> 128 x4 -> 3 (a TENSOR kernel, 116 224 flop/sample forward, 2x that backward),
    # five launches each way: per-launch average.  4th entry = algorithmic HBM bytes/sample of the whole net (inputs
    # in, colours out / cotangents in, input gradients out) for the traffic-over-algorithmic ratio.
    "rsdf_relu_layer_fwd": ("tensor", _F_RAD / 5.0, "relu_layer_fwd_kernel", (268 + 12) / 5.0),
    "rsdf_relu_layer_bwd": ("tensor", 2 * _F_RAD / 5.0, "relu_layer_bwd_kernel", (12 + 268 + 268) / 5.0),
    "rsdf_neus_render_fwd": ("hbm", 64, "sdf_render_fwd_kernel<3, true, false>"),
    "rsdf_neus_render_bwd": ("hbm", 64 + 32, "sdf_render_bwd_kernel<3, true, false>"),
}
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of THIS round's
# kernels (profiles/kernels_r02_full.txt: one training step of `bench.py --steps 2`, 3 374 096 samples), expressed per
# sample; per-launch averages for the multi-launch entry points
NCU_TRAFFIC_SOURCE = "profiles/kernels_r02_full.txt (ncu --set full, one capture; not re-measured by this run)"
NCU_TRAFFIC_PER_SAMPLE = {
    "rsdf_hashgrid_fwd": 686, "rsdf_hashgrid_bwd_input": 526, "rsdf_hashgrid_jvp": 520, "rsdf_hashgrid_bwd_table2": 322,
    "rsdf_sdf_mlp_fwd": 461, "rsdf_sdf_mlp_bwd": 604,
    "rsdf_relu_layer_fwd": (1087 + 3 * 1013 + 527) / 5.0, "rsdf_relu_layer_bwd": (1026 + 3 * 1532 + 1034) / 5.0,
    "rsdf_neus_render_fwd": 43.4, "rsdf_neus_render_bwd": 74.5,
}


def peaks():
    """(hbm GB/s, dense bf16 TFLOP/s sustained, source).  The step is long, so the sustained tensor figure applies."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1500.0))), "measured"
    return 6650.0, 1500.0, "fallback"


def roofline_of(name, calls, total_ms, n_samples, hbm_peak, tc_peak, peak_src):
    bound, work, kernel = KERNEL_WORK[name][:3]
    alg_bytes = KERNEL_WORK[name][3] if len(KERNEL_WORK[name]) > 3 else (work if bound == "hbm" else None)
    avg_ms = total_ms / max(calls, 1)
    if avg_ms <= 0:
        return None
    if bound == "hbm":
        achieved, peak, unit = work * n_samples / (avg_ms / 1e3) / 1e9, hbm_peak, "GB/s"
    else:
        achieved, peak, unit = work * n_samples / (avg_ms / 1e3) / 1e12, tc_peak, "TFLOP/s"
    tr = NCU_TRAFFIC_PER_SAMPLE.get(name)
    return {"bound": bound, "kernel": kernel, "achieved": achieved, "peak": peak, "peak_source": peak_src,
            "unit": unit, "frac": achieved / peak, "traffic": tr * n_samples if tr else None,
            "traffic_source": NCU_TRAFFIC_SOURCE if tr else None,
            "traffic_over_algorithmic_bytes": (tr / alg_bytes) if (tr and alg_bytes) else None,
            "work_per_sample": work, "avg_launch_ms": avg_ms, "launches_per_step": None}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------
class CpuTrainer:
    """The CPU arm: the oracle port of the SAME loop the GPU arm times (rise_sdf_b200.train.NeusTrainer.step):
    update_step (cos-anneal + occupancy refresh every 16 steps, lib/nerfacc/grid.py:196-239), render, losses
    (systems/neus.py:98-135), backward incl. the eikonal double-backward, torch.optim.Adam with the config's groups.
    Pure PyTorch fp32 on the host cores; only bench.py's baseline legs use it."""

    def __init__(self, threads, seed=42, global_step=5000):
        import torch
        from oracle import neus as oneus
        from rise_sdf_b200 import synthetic as syn
        torch.set_num_threads(threads)
        self.torch, self.oneus, self.syn = torch, oneus, syn
        self.P = P = oneus.make_params(seed=seed)
        with torch.no_grad():
            P.geo_mlp[0]["weight_v"][:, 3:].normal_(0.0, 0.05, generator=torch.Generator().manual_seed(1))
        for t in P.tensors():
            t.requires_grad_(True)
        geo = [P.table] + [v for l in P.geo_mlp for v in l.values()]
        tex = [v for l in P.tex_mlp for v in l.values()]
        self.opt = torch.optim.Adam([{"params": geo, "lr": 0.01}, {"params": tex, "lr": 0.01},
                                     {"params": [P.variance], "lr": 0.001}], lr=0.01, betas=(0.9, 0.99), eps=1e-15)
        self.step_size = 1.732 * 2 * 1.5 / 1024
        self.global_step = global_step
        self.gen = torch.Generator().manual_seed(seed)
        self.binary = syn.analytic_grid("ball")
        self.occs = self.binary.flatten().float() * 0.01
        self.poses, self.dirs = syn.camera_poses(), syn.ray_directions()

    def step(self, n_rays, seed):
        torch, oneus = self.torch, self.oneus
        t0 = time.perf_counter()
        if self.global_step % 16 == 0:
            with torch.no_grad():
                self.occs, self.binary = oneus.grid_update(self.occs, self.global_step,
                                                           oneus.occ_eval_fn(self.P, self.step_size), [-1.5] * 3 + [1.5] * 3,
                                                           occ_thre=0.001, gen=self.gen)
        rays, rgb, fg, bg = self.syn.training_rays(n_rays, seed=seed, poses=self.poses, directions=self.dirs)
        self.opt.zero_grad(set_to_none=True)
        out = oneus.forward(self.P, rays, self.binary.numpy(), self.step_size, min(1.0, self.global_step / 20000),
                            background=bg, jitter=torch.rand(n_rays, generator=self.gen), training=True, create_graph=True)
        loss, _ = oneus.loss(out, rgb, fg)
        loss.backward()
        self.opt.step()
        self.global_step += 1
        return time.perf_counter() - t0, int(out["num_samples"])


CPU_RAYS = 2048          # rays per CPU step: a bounded sample of the 8192-ray step (fixed: like-for-like per-ray cost)


def run_reference(args):
    """Reference arm: the reference has no CPU path and its CUDA dependencies (tiny-cuda-nn, nerfacc 0.5.3) cannot be
    installed offline, so this arm times the CPU port (oracle/, kind "port") of the same training loop on all host
    cores, CPU_RAYS rays per step (of the 8192 the GPU arm steps)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    tr = CpuTrainer(cores, global_step=5001)
    tr.step(64, seed=1)                                   # import / allocator warm-up (untimed, tiny)
    for i in range(max(args.warmup, 1) - 1):
        tr.step(CPU_RAYS, seed=100 + i)
    runs = [tr.step(CPU_RAYS, seed=200 + i) for i in range(args.steps)]
    ms = 1e3 * sum(r[0] for r in runs) / len(runs)
    v = CPU_RAYS / (ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD + "; CPU port of the reference path",
                   "rays_per_step_sample": CPU_RAYS, "rays_per_step_full": N_RAYS,
                   "samples_per_step_sample": int(sum(r[1] for r in runs) / len(runs))},
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": f"{CPU_RAYS} of {N_RAYS} rays per step (ball occupancy grid refreshed every 16 steps, "
                                   f"update_step + fwd + loss + bwd incl. eikonal double-backward + Adam), "
                                   f"{args.steps} steps after {args.warmup} warm-up"},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------
def _max_over_ranks(vals, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor(vals, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def run_relight(args, dev, world, rank, variance, n_frames=2):
    """BASELINE configs[3]: relit 800x800 frames under 2 synthetic env maps, pixels sharded over the
    ranks (no collective).  Returns a dict for the JSON line (whole-job frames/s, max over ranks).
    `value`: rays resident on the device, shards left on the device.  `e2e`: what the reference's test loop delivers
    (models/utils.py:37-41 `.cpu()`s every chunk): pose rays start in pinned host memory, the rank's shard of every relit
    frame is written into its rows of a full 800x800x3 pinned host frame (`frame[rank::world]`), copies inside the
    timed region."""
    import torch
    import torch.distributed as dist
    from rise_sdf_b200 import synthetic as syn
    from rise_sdf_b200.relight import EnvSet, balanced_tile, my_pixels, render_frame_shard, synthetic_envs
    from rise_sdf_b200.split_mixed_occ import SplitMixedOCCModel, split_mixed_occ_config

    torch.manual_seed(42)
    model = SplitMixedOCCModel(split_mixed_occ_config()).to(dev)
    with torch.no_grad():
        model.geometry.network.layers[0].weight_v[:, 3:].normal_(0.0, 0.05)
        model.variance.variance.fill_(variance)     # 0.3 = the config's init_val (inv_s = e^3); 0.5 = a trained sharpness
    model.train()
    model.update_step(0, 80000)                      # trainer.max_steps: all levels on, stage 1 (systems/base.py:123-126)
    gj = torch.Generator().manual_seed(7)
    model.occupancy_grid._update(0, model.occ_eval_fn, occ_thre=0.001, jitter=torch.rand(128 ** 3, 3, generator=gj))
    model.eval()
    model.background_color = torch.ones(3, device=dev)
    envs = EnvSet(model, synthetic_envs())
    poses, dirs = syn.camera_poses(), syn.ray_directions()
    host_rays = [syn.frame_rays(7 * k + 3, poses, dirs).pin_memory() for k in range(n_frames)]
    frames = [r.to(dev) for r in host_rays]
    n_env = len(envs.maps)
    host_frames = [[torch.zeros(640000, 3).pin_memory() for _ in range(n_env)] for _ in range(n_frames)]
    # warm-up: one full frame of another pose at a 25 % finer march, so that the caching allocator already holds
    # blocks for every tile size of the timed frames (a first-time size is a cudaMalloc in the timed region)
    rs = model.render_step_size
    model.render_step_size = rs / 1.25
    render_frame_shard(model, syn.frame_rays(50, poses, dirs).to(dev), envs, rank, world)
    model.render_step_size = rs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        barrier()
        return e0.elapsed_time(e1), r

    def resident():
        for f in frames:
            out, tiles = render_frame_shard(model, f, envs, rank, world)
        return out

    def e2e():
        for k, hr in enumerate(host_rays):
            f = hr.to(dev, non_blocking=True)
            out, _ = render_frame_shard(model, f, envs, rank, world)
            for e in range(n_env):
                host_frames[k][e][my_pixels(640000, rank, world)].copy_(out[e]["comp_rgb_phys_full"], non_blocking=True)
        torch.cuda.current_stream().synchronize()        # the frames are on the host when the clock stops
        return None

    resident()                                       # untimed: with the finer-march frame above, 3 warm-up frame renders
    ms, out = timed(resident)
    ms_e2e, _ = timed(e2e)
    # the reference's loop order for comparison: every env map re-renders the frame from scratch
    ms_indep, _ = timed(lambda: render_frame_shard(model, frames[0], envs, rank, world, share_across_envs=False))
    ms, ms_e2e, ms_indep = _max_over_ranks([ms, ms_e2e, ms_indep], dev, world)
    n_relit = n_frames * n_env
    shard = 640000 // world
    return {"metric": "relit_800x800_frames_per_s", "value": n_relit / (ms / 1e3), "unit": "frames/s",
            "variance": variance, "ms_per_frame": ms / n_relit, "frames": n_relit, "env_maps": n_env, "n_gpus": world,
            "scaling": "strong", "occupied_fraction": round(float(model.occupancy_grid.binaries.float().mean()), 4),
            "e2e": {"value": n_relit / (ms_e2e / 1e3), "unit": "frames/s",
                    "h2d_bytes_per_frame": 640000 * 6 * 4 // n_env, "d2h_bytes_per_frame": shard * 3 * 4,
                    "what": "pose rays from pinned host memory; every relit frame's shard copied into its rows of a full "
                            "800x800x3 pinned host frame inside the timed region"},
            "sharding": f"pixels interleaved over the ranks (rank r renders pixels r, r+{world}, ...), "
                        f"{balanced_tile(640000, world)}-ray tiles inside a shard, no collective",
            "env_sharing": "each tile is rendered under both env maps back to back; sampling, field evaluations, "
                           "material networks and the secondary bounce run once per tile, emitter lookups + "
                           "compositing per map (frames bit-identical to independent renders: "
                           "tests/test_gpu_splitsum.py::test_relighting_reuse_is_bit_identical)",
            "frames_per_s_independent_renders": n_env / (ms_indep / 1e3),
            "mean_rgb": float(out[0]["comp_rgb_phys_full"].mean()) if out[0]["comp_rgb_phys_full"].numel() else None}


def run_split_train(args, dev, world, rank, n_rays=4096):
    """BASELINE configs[2]: split-mixed-occ training step -- update_step (occupancy refresh every 16 steps), split-sum
    PBR shading at stage 1, env-light mip pyramid rebuilt every step, finite-difference normals + curvature probe,
    reflection bounce, all losses of systems/split_occ.py, backward, (all-reduce), Adam + schedule.  4096 rays/GPU/step."""
    import torch
    import torch.distributed as dist
    from rise_sdf_b200 import _lib as L
    from rise_sdf_b200 import synthetic as syn
    from rise_sdf_b200.split_mixed_occ import SplitMixedOCCModel, split_mixed_occ_config
    from rise_sdf_b200.train import SplitTrainer

    torch.manual_seed(42)
    model = SplitMixedOCCModel(split_mixed_occ_config()).to(dev)
    with torch.no_grad():
        model.geometry.network.layers[0].weight_v[:, 3:].normal_(0.0, 0.05)
    model.train()
    trainer = SplitTrainer(model)
    trainer.global_step = 20001                      # mid-training: all hash levels on, stage 1 (split-sum shading active)
    model.update_step(0, 20000)
    gj = torch.Generator().manual_seed(7)
    model.occupancy_grid._update(0, model.occ_eval_fn, occ_thre=0.001, jitter=torch.rand(128 ** 3, 3, generator=gj))
    poses, dirs = syn.camera_poses(), syn.ray_directions()
    host = [tuple(t.pin_memory() for t in syn.training_rays(n_rays, seed=7 + 1000 * b, rank=rank, poses=poses, directions=dirs))
            for b in range(2)]
    batches = [tuple(t.to(dev) for t in h) for h in host]
    torch.cuda.manual_seed(4321 + rank)
    rs = model.render_step_size                      # allocator priming step, as in run_ours -- with more headroom: the
    model.render_step_size = rs / 1.3                # sample count of this config drifts upwards as the step count grows
    trainer.step(*batches[0], update=False)
    model.render_step_size = rs
    trainer.global_step = 19984                      # one untimed occupancy-refresh step (see run_ours)
    trainer.step(*batches[1])
    trainer.global_step = 20001
    # more than two occupancy-refresh periods of warm-up: this config's sample count keeps drifting (the model is learning), and
    # every first-time size is an allocator growth event (measured: one 150-200 ms step in an otherwise 67 ms loop)
    for i in range(max(args.warmup, 40)):
        trainer.step(*batches[i % 2])
    steps = max(3, args.steps // 2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            r = fn(i)
        e1.record()
        barrier()
        return e0.elapsed_time(e1), r

    L.stats_reset(True, ())
    ms, (_, out) = timed(lambda i: trainer.step(*batches[i % 2]), steps)
    launches = L.STATS["launches"]
    L.stats_reset(False)

    def step_e2e(i):
        b = tuple(t.to(dev, non_blocking=True) for t in host[i % 2])
        loss, _ = trainer.step(*b)
        return float(loss.item())
    ms_e2e, _ = timed(step_e2e, steps)
    L.stats_reset(True, SPLIT_TIMED)
    n_prof = 3
    timed(lambda i: trainer.step(*batches[i % 2]), n_prof)
    ktimes = L.stats_times_ms()
    L.stats_reset(False)
    ms, ms_e2e = _max_over_ranks([ms / steps, ms_e2e / steps], dev, world)
    return {"metric": "split_train_rays_per_s", "value": world * n_rays / (ms / 1e3), "unit": "rays/s",
            "ms_per_step": ms, "steps": steps, "rays_per_gpu": n_rays, "n_gpus": world, "scaling": "weak",
            "e2e": {"value": world * n_rays / (ms_e2e / 1e3), "unit": "rays/s",
                    "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host[0]), "d2h_bytes_per_step": 4},
            "gpu_launches_per_step": launches / steps,
            "kernel_ms_per_step": {k: round(v[1] / n_prof, 4) for k, v in sorted(ktimes.items())},
            "primary_samples_per_step": int(out["num_samples"].sum()),
            "workload": "split-mixed-occ-tensoir training step (configs[2]): update_step (occupancy refresh every 16 "
                        "steps), stage 1 split-sum shading, build_mips per step, FD normals + curvature, reflection "
                        "bounce, fwd+loss+bwd+Adam"}


SPLIT_TIMED = ["rsdf_hashgrid_fwd", "rsdf_hashgrid_bwd_table", "rsdf_hashgrid_bwd_input", "rsdf_hashgrid_bwd_bwd",
               "rsdf_hashgrid_fd6", "rsdf_sdf_mlp_fwd", "rsdf_sdf_mlp_bwd", "rsdf_sdf_mlp_eval", "rsdf_relu_layer_fwd",
               "rsdf_relu_layer_bwd", "rsdf_mlp_fwd", "rsdf_mm_stream", "rsdf_mm_tn", "rsdf_specular_cubemap", "rsdf_specular_apply",
               "rsdf_diffuse_cubemap", "rsdf_cube_sample_fwd", "rsdf_cube_sample_bwd", "rsdf_tex2d_fwd", "rsdf_tex2d_bwd",
               "rsdf_weight_from_alpha_fwd", "rsdf_weight_from_alpha_bwd", "rsdf_accumulate_fwd", "rsdf_accumulate_bwd",
               "rsdf_march_count_keep", "rsdf_march_compact", "rsdf_adam_step", "rsdf_sh_fwd", "rsdf_sh_bwd",
               "rsdf_split_shade_fwd", "rsdf_split_shade_bwd", "rsdf_split_render_fwd", "rsdf_split_render_bwd",
               "rsdf_freq_encode_fwd", "rsdf_freq_encode_bwd", "rsdf_normalize3_fwd", "rsdf_normalize3_bwd"]


# ------------------------------------------------------------------------------------------
def gpu_reference(dev, model, rays, n_rep=20):
    """what of the reference's GPU build can be timed on this box, next to its replacement, on
    the same inputs (CUDA events, 3 warm-ups, median of n_rep):
      * kernel vs kernel: the reference's in-tree kernels compiled unmodified (oracle/_ref/*.so) -- ray marching
        (lib/nerfacc/cuda/csrc/ray_marching.cu:81-289), weight_from_alpha fwd+bwd (render_weight.cu:86-154), the GGX /
        diffuse cube-map prefilter fwd+bwd at the six pyramid levels (lib/renderutils/c_src/cubemap.cu);
      * a PyTorch-CUDA STAND-IN training step for the parts whose reference implementation (tiny-cuda-nn, nerfacc
        0.5.3) cannot be installed: the oracle's torch code moved to the device (gather hash grid, cuBLAS nn.Linear,
        autograd double backward, index_add accumulate) behind the reference's compiled march."""
    import statistics

    import torch
    from oracle import ref as oref
    from rise_sdf_b200 import _lib as L
    from rise_sdf_b200 import nerfacc as rn
    from rise_sdf_b200 import renderutils as ru

    def med(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(n_rep):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts)

    out = {"note": "reference = the reference's own sources compiled unmodified for sm_100a (oracle/_ref); same inputs"}
    C, P = oref.nerfacc_cuda(), oref.renderutils_plugin()
    ro, rd = rays[:, :3].contiguous(), rays[:, 3:].contiguous()
    grid = model.occupancy_grid
    step = model.render_step_size
    roi = grid.aabbs[0].contiguous()
    if C is not None:
        def ref_march():
            tmin, tmax = C.ray_aabb_intersect(ro, rd, roi)
            return C.ray_marching(ro, rd, tmin, tmax, roi, grid.binaries[0], C.ContractionType.AABB, step, 0.0)
        def our_march():
            return grid.sampling(ro, rd, render_step_size=step, _return_packed=True)
        pk, ri, ts, te = ref_march()
        S = int(ri.shape[0])
        out["march"] = {"rays": int(ro.shape[0]), "samples": S, "reference_ms": med(ref_march), "ours_ms": med(our_march),
                        "what": "ray_aabb_intersect + two-pass ray_marching (incl. its cumsum and .item() sync) vs "
                                "sampling() (one warp-per-ray march + device scan + copy, one read-back)"}
        a = torch.rand(S, 1, device=dev) * 0.2
        gw = torch.randn(S, 1, device=dev)
        packed = pk.contiguous()
        def ref_w():
            w = C.weight_from_alpha_forward_naive(packed, a)
            return C.weight_from_alpha_backward_naive(w, gw, packed, a)
        a1, gw1 = a[:, 0].contiguous(), gw[:, 0].contiguous()
        def our_w():
            w, T = rn._WeightFromAlpha.forward(_Ctx(), packed, a1)
            ga = torch.zeros_like(a1)
            L.call("rsdf_weight_from_alpha_bwd", L.ptr(packed), L.ptr(a1), L.ptr(w), L.ptr(T), L.ptr(gw1), None,
                   packed.shape[0], L.ptr(ga), L.stream())
            return ga
        out["weight_from_alpha_fwd_bwd"] = {"samples": S, "reference_ms": med(ref_w), "ours_ms": med(our_w),
                                            "what": "weight_from_alpha_{forward,backward}_naive (1 thread/ray) vs the "
                                                    "warp-per-ray shuffle scan"}
        out["march"]["ratio"] = out["march"]["reference_ms"] / out["march"]["ours_ms"]
        out["weight_from_alpha_fwd_bwd"]["ratio"] = (out["weight_from_alpha_fwd_bwd"]["reference_ms"]
                                                     / out["weight_from_alpha_fwd_bwd"]["ours_ms"])
    if P is not None:
        base = torch.rand(6, 512, 512, 3, device=dev) * 0.5 + 0.25
        levels = [base]
        while levels[-1].shape[1] > 16:
            x = levels[-1].permute(0, 3, 1, 2)
            levels.append(torch.nn.functional.avg_pool2d(x, (2, 2)).permute(0, 2, 3, 1).contiguous())
        n = len(levels)
        rough = [(i / (n - 2)) * 0.42 + 0.08 for i in range(n - 1)] + [1.0]
        cuts = [ru.ndf_cutoff(r, 0.99) for r in rough]
        rb = [P.specular_bounds(l.shape[1], c) for l, c in zip(levels, cuts)]
        go = [torch.randn(6, l.shape[1], l.shape[1], 4, device=dev) for l in levels]
        gd = torch.randn(6, 16, 16, 3, device=dev)
        def ref_pf():
            for l, b, r, c, g in zip(levels, rb, rough, cuts, go):
                P.specular_cubemap_fwd(l, b, r, c)
                P.specular_cubemap_bwd(l, b, g, r, c)
            P.diffuse_cubemap_fwd(levels[-1]); P.diffuse_cubemap_bwd(levels[-1], gd)
        lv = [l.clone().requires_grad_(True) for l in levels]
        def our_pf():
            for l, r, g in zip(lv, rough, go):
                o = ru.specular_cubemap(l, r, 0.99)
                o.backward(g[..., :3])
            d = ru.diffuse_cubemap(lv[-1]); d.backward(gd)
        out["cubemap_prefilter_fwd_bwd"] = {"levels": [int(l.shape[1]) for l in levels], "reference_ms": med(ref_pf),
                                            "ours_ms": med(our_pf),
                                            "what": "specular_cubemap fwd+bwd at 512..16 + diffuse_cubemap fwd+bwd "
                                                    "(bounds cached on both sides); ours incl. autograd node overhead"}
        out["cubemap_prefilter_fwd_bwd"]["ratio"] = (out["cubemap_prefilter_fwd_bwd"]["reference_ms"]
                                                     / out["cubemap_prefilter_fwd_bwd"]["ours_ms"])
    out["standin_train_step"] = standin_step(dev, model, rays, C)
    return out


class _Ctx:
    """minimal autograd-ctx stand-in for calling a Function's forward directly (bench only)"""
    def save_for_backward(self, *a):
        pass


def standin_step(dev, model, rays, C, n_rep=3):
    """PyTorch-CUDA stand-in for the reference's tcnn + nerfacc-0.5.3 training step: the oracle's
    torch restatement (oracle/fields.py: gather-based hash grid, F.linear = cuBLAS SGEMM, autograd.grad(create_graph)
    for the analytic normal, padded-cumprod weights, index_add accumulation) run on the device on the SAME rays, grid
    and weights, behind the reference's compiled march when it is available.  fwd + loss + bwd (no optimizer).
    NOT the reference's kernels -- tiny-cuda-nn's fused fp16 hash grid is far faster than this -- labelled as such."""
    import torch
    import torch.nn.functional as F
    from oracle import fields as of
    from oracle import neus as oneus
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_params_from_model
    P = oracle_params_from_model(model)
    P = oneus.NeusParams(P.table.to(dev), [{k: v.to(dev) for k, v in l.items()} for l in P.geo_mlp],
                         [{k: v.to(dev) for k, v in l.items()} for l in P.tex_mlp], P.variance.to(dev), P.meta,
                         P.radius, P.sh_degree)
    for t in P.tensors():
        t.requires_grad_(True)
    ro, rd = rays[:, :3].contiguous(), rays[:, 3:].contiguous()
    n_rays = ro.shape[0]
    grid, step = model.occupancy_grid, model.render_step_size
    roi = grid.aabbs[0].contiguous()
    rgb_t = torch.rand(n_rays, 3, device=dev)
    fg = torch.ones(n_rays, device=dev)
    bg = torch.ones(3, device=dev)

    def one():
        with torch.no_grad():
            if C is not None:
                tmin, tmax = C.ray_aabb_intersect(ro, rd, roi)
                pk, ri, ts, te = C.ray_marching(ro, rd, tmin, tmax, roi, grid.binaries[0], C.ContractionType.AABB, step, 0.0)
                ts, te = ts[:, 0], te[:, 0]
            else:
                ri, ts, te = grid.sampling(ro, rd, render_step_size=step)
        t_o, t_d = ro[ri], rd[ri]
        mid = (ts + te)[:, None] / 2.0
        pos = t_o + t_d * mid
        sdf, grad, feat = of.sdf_field(pos, P.table, P.meta, P.geo_mlp, P.radius, with_grad=True, create_graph=True)
        normal = F.normalize(grad, p=2, dim=-1)
        alpha = of.get_alpha(sdf, normal, t_d, te - ts, P.inv_s.view(1, 1), model.cos_anneal_ratio)
        rgb = of.radiance(feat, t_d, normal, P.tex_mlp, P.sh_degree)
        # nerfacc-style scan via the compiled reference kernel is not differentiable from here: exclusive cumprod by
        # segment in torch (sort-free: samples are packed by ray)
        counts = torch.bincount(ri, minlength=n_rays)
        starts = torch.cumsum(counts, 0) - counts
        logt = torch.log1p(-alpha.clamp(max=1 - 1e-7))
        cs = torch.cumsum(logt, 0)
        excl = cs - logt
        T = torch.exp(excl - excl[starts][ri])
        w = T * alpha
        z = lambda d: torch.zeros(n_rays, d, device=dev)
        op = z(1).index_add(0, ri, w[:, None])
        comp = z(3).index_add(0, ri, w[:, None] * rgb)
        out = {"comp_rgb_full": comp + bg * (1 - op), "opacity": op, "rays_valid": op > 0, "sdf_grad_samples": grad,
               "sdf_samples": sdf}
        loss, _ = oneus.loss(out, rgb_t, fg)
        for t in P.tensors():
            t.grad = None
        loss.backward()
        return int(ri.shape[0])

    S = one()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n_rep):
        one()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n_rep
    return {"ms_per_step": ms, "rays_per_s": n_rays / (ms / 1e3), "rays": n_rays, "samples": S,
            "kind": "PyTorch-CUDA stand-in for tcnn/nerfacc-0.5.3 (oracle torch code on the device, cuBLAS fp32 nn.Linear, "
                    "gather hash grid, autograd double backward); fwd+loss+bwd, no optimizer, no occupancy update",
            "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}


# ------------------------------------------------------------------------------------------
TRAIN_TIMED = ["rsdf_hashgrid_fwd", "rsdf_hashgrid_bwd_table", "rsdf_hashgrid_bwd_input", "rsdf_hashgrid_bwd_bwd",
               "rsdf_march_count", "rsdf_march_fill", "rsdf_march_count_keep", "rsdf_march_compact", "rsdf_adam_step",
               "rsdf_neus_render_fwd", "rsdf_neus_render_bwd", "rsdf_sdf_mlp_fwd", "rsdf_sdf_mlp_bwd", "rsdf_sdf_mlp_eval",
               "rsdf_relu_layer_fwd", "rsdf_relu_layer_bwd", "rsdf_absmax2", "rsdf_hashgrid_bwd_table2", "rsdf_hashgrid_jvp",
               "rsdf_sh_fwd", "rsdf_sh_bwd", "rsdf_sdf_reg_fwd", "rsdf_sdf_reg_bwd", "rsdf_sample_setup",
               "rsdf_normalize3_fwd", "rsdf_normalize3_bwd", "rsdf_grid_pack_bits"]


def run_ours(args):
    import torch
    import torch.distributed as dist
    from rise_sdf_b200 import _lib as L
    from rise_sdf_b200 import synthetic as syn
    from rise_sdf_b200.neus import NeuSModel, neus_blender_config
    from rise_sdf_b200.train import NeusTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the hot path has no CPU fallback")
    if os.environ.get("RSDF_NCU_RANGE"):
        # profiling aid only: NVTX push/pop ranges are per thread, so `ncu --nvtx --nvtx-include "timed/"` sees the
        # backward kernels only when autograd runs them on the calling thread
        torch.autograd.set_multithreading_enabled(False)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.lib()

    torch.manual_seed(42)                      # identical weights on every rank (DDP broadcast)
    model = NeuSModel(neus_blender_config()).to(dev).train()
    with torch.no_grad():                      # "mid-training" state: hash features are live
        model.geometry.network.layers[0].weight_v[:, 3:].normal_(0.0, 0.05)
    trainer = NeusTrainer(model)
    # occupancy grid from the random-init SDF (warm-up branch of update_every_n_steps, untimed); the timed loop then
    # runs the reference's per-batch sequence from global step 5001 on: cos-anneal ratio 0.25, occupancy refresh
    # (quarter of the cells + the occupied ones) whenever step % 16 == 0
    gj = torch.Generator().manual_seed(7)
    model.occupancy_grid._update(0, model.occ_eval_fn, occ_thre=0.001, jitter=torch.rand(128 ** 3, 3, generator=gj))
    trainer.global_step = 5001
    occ_frac = float(model.occupancy_grid.binaries.float().mean())

    poses, dirs = syn.camera_poses(), syn.ray_directions()
    n_batches = 4
    host = []
    for b in range(n_batches):
        rays, rgb, fg, bg = syn.training_rays(N_RAYS, seed=42 + 1000 * b, rank=rank, poses=poses, directions=dirs)
        host.append(tuple(t.pin_memory() for t in (rays, rgb, fg, bg)))
    devb = [tuple(t.to(dev) for t in h) for h in host]
    # stratified jitter draws come from the CUDA generator inside sampling(): seed per rank
    torch.cuda.manual_seed(1234 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(i):
        return trainer.step(*devb[i % n_batches])

    def step_e2e(i):
        rays, rgb, fg, bg = (t.to(dev, non_blocking=True) for t in host[i % n_batches])
        loss, out = trainer.step(rays, rgb, fg, bg)
        return float(loss.item()), out         # D2H read of the step's result

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            r = fn(i)
        e1.record()
        barrier()
        return e0.elapsed_time(e1), r

    # Prime the caching allocator: stratified jitter makes the sample count differ from step to step, and a size the
    # allocator has not seen yet costs a cudaMalloc inside the timed region.  One untimed step at a 10 % finer march
    # leaves cached blocks that cover every size the timed steps ask for.
    rs = model.render_step_size
    model.render_step_size = rs / 1.1
    trainer.step(*devb[0], update=False)
    model.render_step_size = rs
    # ... and one untimed step ON an occupancy-refresh step (step % 16 == 0), so that the refresh inside the timed region
    # is not the first one the allocator sees (measured at N=2, 12 steps: 65 ms of cudaMalloc in that one step)
    trainer.global_step = 4992
    trainer.step(*devb[1])
    trainer.global_step = 5001
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                        # nvidia-smi is forked before, not inside, the timed region
    for i in range(args.warmup):
        step_resident(i)
    # ---- timed: resident inputs, NO per-kernel instrumentation (launch counting only) ------------
    L.stats_reset(True, ())
    sampler.rows.clear()                       # keep only the samples taken during the timed region
    torch.cuda.nvtx.range_push("timed")
    g0 = trainer.global_step
    ms_total, (_, out) = timed(step_resident, args.steps)
    torch.cuda.nvtx.range_pop()
    clocks = sampler.stop() if rank == 0 else None
    launches = L.STATS["launches"]
    L.stats_reset(False)
    n_updates = sum(1 for s in range(g0, g0 + args.steps) if s % 16 == 0)
    # ---- timed: end to end (host buffers) ---------------------------------------------------
    for i in range(2):
        step_e2e(i)
    ms_e2e, _ = timed(step_e2e, args.steps)
    # ---- instrumented pass: per-kernel CUDA events on the launching stream (perturbs the step: not the headline) ----
    n_prof = min(args.steps, 8)
    L.stats_reset(True, TRAIN_TIMED)
    ms_prof, (_, out) = timed(step_resident, n_prof)
    ktimes = L.stats_times_ms()
    L.stats_reset(False)
    n_samples = int(out["num_samples"].item())

    # ---- the reduced-precision MLP variant (single fp16 plane, one product per GEMM; tests/test_gpu_mlp_fp16.py) ----
    variant = None
    if not args.no_variant:
        from rise_sdf_b200.network_utils import VanillaMLP
        VanillaMLP.mlp_precision = "fp16"
        try:
            for i in range(3):
                step_resident(i)
            ms_v, _ = timed(step_resident, args.steps)
            L.stats_reset(True, TRAIN_TIMED)
            timed(step_resident, n_prof)
            kv = L.stats_times_ms()
            L.stats_reset(False)
        finally:
            VanillaMLP.mlp_precision = "fp32"
        (ms_v,) = _max_over_ranks([ms_v], dev, world)
        variant = {"mlp_precision": "fp16 single plane, one tcgen05 product per GEMM, fp32 TMEM accumulation",
                   "value": world * N_RAYS / (ms_v / args.steps / 1e3), "unit": "rays/s", "ms_per_step": ms_v / args.steps,
                   "stated_tolerance": "rendered rgb/opacity/depth 5e-3, loss 1e-3, parameter gradients 3e-2 rel-L2 vs "
                                       "float64 (tests/test_gpu_mlp_fp16.py); NOT the parity path, not the headline",
                   "kernel_ms_per_step": {k: round(v[1] / n_prof, 4) for k, v in sorted(kv.items())
                                          if k in ("rsdf_sdf_mlp_fwd", "rsdf_sdf_mlp_bwd", "rsdf_relu_layer_fwd",
                                                   "rsdf_relu_layer_bwd")}}
    # ---- the reference's GPU build, as far as it can be had on this box (rank 0 of a 1-GPU run) ----
    gref = None
    if world == 1 and not args.no_gpu_reference:
        gref = gpu_reference(dev, model, devb[0][0])
    # ---- second headlines: split-sum training step, relit frames/s (all ranks take part; no collective on the data path)
    relight = split_train = relight_03 = None
    if not args.no_relight:
        del trainer, devb
        torch.cuda.empty_cache()
        split_train = run_split_train(args, dev, world, rank)
        torch.cuda.empty_cache()
        relight = run_relight(args, dev, world, rank, variance=0.5)
        torch.cuda.empty_cache()
        relight_03 = run_relight(args, dev, world, rank, variance=0.3)

    ms_total, ms_e2e = _max_over_ranks([ms_total, ms_e2e], dev, world)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_step = ms_total / args.steps
    value = world * N_RAYS / (ms_step / 1e3)
    e2e_v = world * N_RAYS / (ms_e2e / args.steps / 1e3)
    hbm_peak, tc_peak, peak_src = peaks()
    rooflines = {}
    for name, (calls, tot) in ktimes.items():
        if name in KERNEL_WORK and calls:
            r = roofline_of(name, calls, tot, n_samples, hbm_peak, tc_peak, peak_src)
            if r:
                r["launches_per_step"] = calls / n_prof
                r["ms_per_step"] = tot / n_prof
                rooflines[name] = r
    dominant = max(rooflines, key=lambda k: rooflines[k]["ms_per_step"]) if rooflines else None
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        tr = CpuTrainer(cores, global_step=5001)
        tr.step(64, seed=1)
        k_cpu = 8                              # ~15-25 s of host work on the 16-core B200 box
        runs = [tr.step(CPU_RAYS, seed=42 + j) for j in range(k_cpu)]
        tc, s_cpu = sum(r[0] for r in runs), sum(r[1] for r in runs)
        cpu = {"value": CPU_RAYS * k_cpu / tc, "unit": "rays/s", "cores": cores, "kind": "port",
               "sample": f"{k_cpu} training steps on {CPU_RAYS} of {N_RAYS} rays each ({s_cpu} samples in total), ball "
                         f"occupancy grid refreshed at step % 16 == 0 (once in this sample), oracle port (pure PyTorch "
                         f"fp32): update_step + fwd + loss + bwd incl. eikonal double-backward + Adam; {tc:.1f} s"}
    ksum = sum(v[1] for v in ktimes.values()) / n_prof
    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "rays_per_gpu": N_RAYS, "samples_per_step": n_samples, "occupied_fraction": round(occ_frac, 4),
                   "occupancy_updates_in_timed_region": n_updates, "first_global_step": g0,
                   "cache": "4 rotating ray batches; per-step working set (hash table 50 MB + ~2 GB activations) "
                            "exceeds the 126 MB L2, no explicit flush",
                   "parallelism": f"dp{world}",
                   "mlp": "tcgen05 kernels, fp16 hi/lo 3-product split with fp32 TMEM accumulation (fp32-class): "
                          "fused SDF field fwd/bwd incl. 2nd order (csrc/sdf_train.cu), per-layer ReLU MLP with "
                          "operand-image streams (csrc/relu_mlp.cu)"},
        "e2e": {"value": e2e_v, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        # the kernel with the largest share of the step; every other hand-written kernel in `rooflines`
        "roofline": rooflines.get(dominant),
        "rooflines": rooflines,
        "instrumented_pass": {"steps": n_prof, "ms_per_step": ms_prof / n_prof, "kernel_sum_ms_per_step": ksum,
                              "note": "per-kernel CUDA events bracket every launch in this pass only; `value` / `e2e` "
                                      "are timed without them"},
        "kernel_ms_per_step": {k: round(v[1] / n_prof, 4) for k, v in sorted(ktimes.items())},
        "clocks": clocks,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    if variant:
        line["mlp_variant_fp16"] = variant
    if gref:
        line["gpu_reference"] = gref
    if split_train:
        line["split_train"] = split_train
    if relight:
        line["relight"] = relight
        line["relight_variance_0.3"] = relight_03
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """the one JSON line, on the real stdout"""
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variant", action="store_true", help="skip the fp16 single-plane MLP variant timing")
    ap.add_argument("--no-gpu-reference", action="store_true",
                    help="skip the reference-GPU-kernel comparison and the PyTorch-CUDA stand-in step (1-GPU runs only)")
    ap.add_argument("--no-relight", action="store_true",
                    help="skip the split-train rays/s and relit-frames/s sections (configs[2], configs[3])")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # stdout carries exactly ONE JSON line: anything a library prints on fd 1 meanwhile (NCCL's version banner under
    # torchrun, for one) goes to stderr instead
    sys.stdout.flush()
    global _REAL_STDOUT
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
