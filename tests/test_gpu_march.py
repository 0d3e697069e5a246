"""K1 parity: CUDA march / AABB / pack vs the C oracle (bit-exact), the reference's own compiled
kernel (oracle/_ref, when present) and the committed golden vectors it produced."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import march as omarch
from oracle import ref as oref
from rise_sdf_b200 import nerfacc as rn
from rise_sdf_b200 import _lib as L
from rise_sdf_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
ROI = [-1.5] * 3 + [1.5] * 3
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def special_rays():
    o = torch.tensor([[0, 0, -4], [0.2, 0.1, -4], [0, 0, 0], [0.3, -0.2, 0.1], [5, 5, 5], [0, 0, -4],
                      [1.5, 0, -4], [-4, 1.49999, 0.3], [0, 4, 0]], dtype=torch.float32)
    d = torch.tensor([[0, 0, 1], [0, 0, 1], [0.6, 0.8, 0], [0, 1, 0], [0, 0, 1], [1, 0, 0],
                      [0, 0, 1], [1, 0, 0], [0.0, -1.0, 0.0]], dtype=torch.float32)
    return torch.cat([o, d], -1)


def rays_for(R, seed=11):
    rays, _, _, _ = syn.training_rays(R, seed=seed)
    if R >= 64:
        sp = special_rays()
        rays[: sp.shape[0]] = sp
    return rays


def bits(a):
    """bit pattern with NaNs canonicalised (0/0 on the slab planes: CPU and GPU emit different
    NaN payloads; the value class is what the reference's comparisons see)."""
    a = np.ascontiguousarray(a, dtype=np.float32).copy()
    a[np.isnan(a)] = np.float32(np.nan)
    return a.view(np.uint32)


SPAN = 3.0 * 3 ** 0.5          # AABB diagonal: the host-side bound `OccGridEstimator.sampling` passes to `_march`
PATHS = ["warp", "thread"]     # the hot path (march_warp_kernel: rsdf_march_count_keep + rsdf_march_compact) and the
                               # reference-ordered two-round march (march_kernel: rsdf_march_count + rsdf_march_fill)


def run_ours(rays, grid, step, use_bits=True, path="warp", cone=0.0, calls=None):
    """`path="warp"` passes the span, so the kernel compared is the one `sampling()` launches on the render path."""
    dev = "cuda"
    o, d = rays[:, :3].contiguous().to(dev), rays[:, 3:].contiguous().to(dev)
    g = grid.to(dev).contiguous()
    tmin, tmax = rn.ray_aabb_intersect(o, d, torch.tensor(ROI))
    gb = rn.pack_bits(g) if use_bits else None
    real = L.call
    seen = []
    L.call = lambda name, *a: (seen.append(name), real(name, *a))[1]
    try:
        packed, ri, ts, te = rn._march(o, d, tmin, tmax, L.host6(ROI), g, gb, step, cone,
                                       SPAN if path == "warp" else None)
    finally:
        L.call = real
    if path == "warp":          # the warp kernel really ran, and (no overflow at this span) the copy round, not a re-march
        assert "rsdf_march_count_keep" in seen and "rsdf_march_fill" not in seen, seen
    else:
        assert "rsdf_march_count" in seen and "rsdf_march_count_keep" not in seen, seen
    if calls is not None:
        calls.extend(seen)
    return [t.cpu().numpy() for t in (tmin, tmax, packed, ri, ts, te)]


@pytest.mark.parametrize("kind", ["ball", "shell", "ones", "zeros", "voxel", "random"])
@pytest.mark.parametrize("R", [1, 4096])
@pytest.mark.parametrize("use_bits", [True, False])
@pytest.mark.parametrize("path", PATHS)
def test_march_bit_exact_vs_oracle(kind, R, use_bits, path):
    rays = rays_for(R)
    grid = syn.analytic_grid(kind)
    step = 1.732 * 2 * 1.5 / (128 if kind == "ones" else 1024)
    tmin, tmax, packed, ri, ts, te = run_ours(rays, grid, step, use_bits, path)
    o, d = rays[:, :3].numpy(), rays[:, 3:].numpy()
    otmin, otmax = omarch.ray_aabb_intersect(o, d, np.array(ROI, np.float32))
    assert np.array_equal(bits(tmin), bits(otmin)) and np.array_equal(bits(tmax), bits(otmax))
    pk, ori, ots, ote = omarch.ray_marching_raw(o, d, otmin, otmax, np.array(ROI, np.float32), grid.numpy(), step)
    assert np.array_equal(packed, pk)
    assert np.array_equal(ri, ori)
    assert np.array_equal(bits(ts), bits(ots)) and np.array_equal(bits(te), bits(ote))


@pytest.mark.parametrize("kind", ["ball", "ones", "random", "zeros"])
def test_one_march_path_equals_two_round_path(kind, monkeypatch):
    """The count round that keeps its intervals + the copy (rsdf_march_count_keep / rsdf_march_compact) against
    the reference's count-then-march-again order (already pinned to the oracle above): same bits.  Also the two
    fallbacks: a ray overflowing its slots (span too small) and a scratch over budget."""
    rays = rays_for(4096)
    grid = syn.analytic_grid(kind)
    step = 1.732 * 2 * 1.5 / (128 if kind == "ones" else 1024)
    o, d = rays[:, :3].contiguous().cuda(), rays[:, 3:].contiguous().cuda()
    g = grid.cuda().contiguous()
    tmin, tmax = rn.ray_aabb_intersect(o, d, torch.tensor(ROI))
    gb = rn.pack_bits(g)
    calls = []
    real = L.call
    monkeypatch.setattr(L, "call", lambda name, *a: (calls.append(name), real(name, *a))[1])
    base = rn._march(o, d, tmin, tmax, L.host6(ROI), g, gb, step, 0.0)
    assert "rsdf_march_count" in calls and "rsdf_march_count_keep" not in calls
    for span, budget, want_compact in ((3.0 * 3 ** 0.5, 1 << 30, True), (0.05, 1 << 30, kind == "zeros"),
                                       (3.0 * 3 ** 0.5, 1 << 20, False)):
        calls.clear()
        monkeypatch.setattr(rn, "KEEP_BUDGET_BYTES", budget)
        got = rn._march(o, d, tmin, tmax, L.host6(ROI), g, gb, step, 0.0, span)
        if int(base[0][:, 1].sum()) > 0:
            assert ("rsdf_march_compact" in calls) == want_compact, (span, budget, calls)
        for a, b in zip(base, got):
            assert a.dtype == b.dtype and torch.equal(a.view(torch.int32) if a.dtype == torch.float32 else a,
                                                      b.view(torch.int32) if b.dtype == torch.float32 else b)
    # through the public entry point (span from the estimator's box): the keep path is the one that runs
    calls.clear()
    monkeypatch.setattr(rn, "KEEP_BUDGET_BYTES", 1 << 30)
    est = rn.OccGridEstimator(torch.tensor(ROI), resolution=128).cuda()
    est.binaries = g[None]
    ri, ts, te, packed = est.sampling(o, d, render_step_size=step, _return_packed=True)
    assert "rsdf_march_count_keep" in calls
    assert torch.equal(ri, base[1]) and torch.equal(ts.view(torch.int32), base[2].view(torch.int32)) \
        and torch.equal(te.view(torch.int32), base[3].view(torch.int32)) and torch.equal(packed, base[0])


@pytest.mark.parametrize("chunks", [(4,), (16, 32), (64, 128, 256)])
def test_front_to_back_visibility_equals_one_shot(chunks, monkeypatch):
    """`ray_marching(alpha_fn=...)`: evaluating the candidates in front-to-back rounds and dropping rays whose
    transmittance is already under `early_stop_eps` must give the same survivors, with the same alphas, as
    lib/nerfacc/ray_marching.py:198-218 (evaluate everything, then mask), while calling alpha_fn on fewer samples."""
    rays = rays_for(2048, seed=3)
    o, d = rays[:, :3].contiguous().cuda(), rays[:, 3:].contiguous().cuda()
    est = rn.OccGridEstimator(torch.tensor(ROI), resolution=128).cuda()
    est.binaries = syn.analytic_grid("ball").cuda()[None]
    step = 1.732 * 2 * 1.5 / 512
    seen = []

    def alpha_fn(ts, te, ri):           # a dense ball of radius 1 inside thin haze: per-sample, batch-independent
        x = o[ri] + d[ri] * ((ts + te) * 0.5)[:, None]
        r = x.norm(dim=-1)
        a = torch.where(r < 1.0, 0.5 + 0.1 * torch.sin(40.0 * r), 0.002 + 0.001 * torch.sin(40.0 * r))
        seen.append((a, ts.shape[0]))
        return a

    monkeypatch.setattr(rn, "VISIBILITY_CHUNKS", ())
    ri0, ts0, te0, rows0 = est.sampling(o, d, alpha_fn=alpha_fn, render_step_size=step, _return_mask=True)
    a0, n0 = seen[0][0][rows0], seen[0][1]
    assert len(seen) == 1 and 0 < ri0.shape[0] < n0
    seen.clear()
    monkeypatch.setattr(rn, "VISIBILITY_CHUNKS", chunks)
    ri1, ts1, te1, rows1 = est.sampling(o, d, alpha_fn=alpha_fn, render_step_size=step, _return_mask=True)
    assert torch.equal(ri0, ri1) and torch.equal(ts0, ts1) and torch.equal(te0, te1)
    assert int(rows1.min()) >= 0                                   # every survivor was evaluated
    assert torch.equal(torch.cat([a for a, _ in seen])[rows1], a0)
    n1 = sum(n for _, n in seen)
    assert len(seen) <= len(chunks) + 1 and n1 <= n0, (len(seen), n1, n0)
    if chunks[0] >= 16:                 # rays go opaque ~15 samples into the ball: later rounds only see the misses
        assert n1 < 0.5 * n0, (n1, n0)
    # early_stop_eps = 0: nothing can be dropped -> the one-shot pass is used
    seen.clear()
    est.sampling(o, d, alpha_fn=alpha_fn, render_step_size=step, early_stop_eps=0.0)
    assert len(seen) == 1 and seen[0][1] == n0


@pytest.mark.parametrize("kind", ["ball", "shell", "random"])
@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("cone", [0.0, 0.004])
def test_march_bit_exact_vs_reference_kernel(kind, path, cone):
    """Both of our kernels against the reference's compiled `ray_marching` (lib/nerfacc/cuda/csrc/ray_marching.cu:81-192),
    with cone_angle = 0 (both configs) and cone_angle > 0 (dt grows with t: march_warp_kernel<false>)."""
    C = oref.nerfacc_cuda()
    if C is None:
        pytest.skip("oracle/_ref/nerfacc_cuda.so not built")
    rays = rays_for(8192, seed=5)
    grid = syn.analytic_grid(kind)
    step = 1.732 * 2 * 1.5 / 1024
    tmin, tmax, packed, ri, ts, te = run_ours(rays, grid, step, path=path, cone=cone)
    o, d = rays[:, :3].contiguous().cuda(), rays[:, 3:].contiguous().cuda()
    roi = torch.tensor(ROI, device="cuda")
    rtmin, rtmax = C.ray_aabb_intersect(o, d, roi)
    assert np.array_equal(bits(tmin), bits(rtmin.cpu().numpy()))
    assert np.array_equal(bits(tmax), bits(rtmax.cpu().numpy()))
    rp, rri, rts, rte = C.ray_marching(o, d, rtmin, rtmax, roi, grid.cuda(), C.ContractionType.AABB, step, cone)
    assert np.array_equal(packed, rp.cpu().numpy())
    assert np.array_equal(ri, rri.cpu().numpy())
    assert np.array_equal(bits(ts), bits(rts[:, 0].cpu().numpy()))
    assert np.array_equal(bits(te), bits(rte[:, 0].cpu().numpy()))


@pytest.mark.parametrize("kind", ["ball", "shell", "random", "ones"])
@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("cone", [0.003, 0.02])
def test_march_cone_angle_vs_oracle(kind, path, cone):
    """cone_angle > 0 (`dt = clamp(t * cone, step, 1e10)`, ray_marching.cu:121-125,165-176): never taken by the two
    configs, but part of the `ray_marching` / `sampling` signature -> both kernels against the C oracle."""
    rays = rays_for(2048, seed=17)
    grid = syn.analytic_grid(kind)
    step = 1.732 * 2 * 1.5 / (128 if kind == "ones" else 512)
    tmin, tmax, packed, ri, ts, te = run_ours(rays, grid, step, path=path, cone=cone)
    o, d = rays[:, :3].numpy(), rays[:, 3:].numpy()
    pk, ori, ots, ote = omarch.ray_marching_raw(o, d, tmin, tmax, np.array(ROI, np.float32), grid.numpy(), step, cone)
    assert np.array_equal(packed, pk) and np.array_equal(ri, ori)
    assert np.array_equal(bits(ts), bits(ots)) and np.array_equal(bits(te), bits(ote))
    if len(ts) and cone >= 0.02:
        assert float((te - ts).max()) > step * 1.5      # the cone really widened the far intervals


@pytest.mark.parametrize("gold", sorted(glob.glob(os.path.join(GOLD, "march_*.npz"))))
@pytest.mark.parametrize("path", PATHS)
def test_march_vs_golden(gold, path):
    z = np.load(gold)
    rays = torch.from_numpy(np.concatenate([z["rays_o"], z["rays_d"]], 1))
    cone = float(z["cone"]) if "cone" in z else 0.0
    tmin, tmax, packed, ri, ts, te = run_ours(rays, torch.from_numpy(z["grid"]), float(z["step"]), path=path, cone=cone)
    assert np.array_equal(packed, z["packed_info"]) and np.array_equal(ri, z["ray_indices"])
    assert np.array_equal(bits(ts), bits(z["t_starts"])) and np.array_equal(bits(te), bits(z["t_ends"]))


def test_frame_sized_march_properties():
    """640 000 rays (one 800x800 frame): size-independent properties instead of the serial oracle."""
    rays = syn.frame_rays(3)
    grid = syn.analytic_grid("ball")
    step = 1.732 * 2 * 1.5 / 1024
    tmin, tmax, packed, ri, ts, te = run_ours(rays, grid, step, path="thread")
    warp = run_ours(rays[::8], grid, step, path="warp")      # (the keep scratch of a whole frame exceeds its budget)
    sel = np.arange(0, len(rays), 8)
    assert np.array_equal(warp[2][:, 1], packed[sel, 1])
    assert packed[:, 1].sum() == len(ri) and packed[-1, 0] + packed[-1, 1] == len(ri)
    assert np.all(np.diff(ri) >= 0)
    assert np.array_equal(packed[:, 0], np.concatenate([[0], np.cumsum(packed[:-1, 1])]))
    assert np.array_equal(np.bincount(ri, minlength=len(rays)), packed[:, 1])
    assert np.all(te > ts)
    mid = (ts + te) * 0.5
    pts = rays[ri, :3].numpy() + rays[ri, 3:].numpy() * mid[:, None]
    assert np.all(omarch.grid_query(pts[::97], np.array(ROI, np.float32), grid.numpy()))
    # a subset of rays re-marched by the oracle must agree bit for bit
    sub = np.arange(0, len(rays), 1601)
    o, d = rays[sub, :3].numpy(), rays[sub, 3:].numpy()
    pk, ori, ots, ote = omarch.ray_marching_raw(o, d, tmin[sub], tmax[sub], np.array(ROI, np.float32), grid.numpy(), step)
    assert np.array_equal(pk[:, 1], packed[sub, 1])
    for k, r in enumerate(sub[:200]):
        b, n = packed[r]
        assert np.array_equal(bits(ts[b:b + n]), bits(ots[pk[k, 0]:pk[k, 0] + n]))


def test_sampling_api_shapes_and_empty():
    est = rn.OccGridEstimator(torch.tensor(ROI), 128).cuda()
    rays = rays_for(256).cuda()
    ri, ts, te = est.sampling(rays[:, :3].contiguous(), rays[:, 3:].contiguous(), render_step_size=0.01)
    assert ri.numel() == 0 and ts.shape == (0,) and ri.dtype == torch.int64      # empty grid -> no samples
    est.binaries = syn.analytic_grid("ball")[None].cuda()
    ri, ts, te = est.sampling(rays[:, :3].contiguous(), rays[:, 3:].contiguous(), render_step_size=0.01,
                              near_plane=3.0, far_plane=4.0)
    assert ts.dim() == 1 and ts.min() >= 3.0 and te.max() <= 4.0 + 0.02
    pk = rn.pack_info(ri, 256).cpu().numpy()
    assert np.array_equal(pk, omarch.pack_info(ri.cpu().numpy(), 256))
    # alpha_fn visibility filter == oracle's
    def alpha_fn(t0, t1, idx):
        return torch.full_like(t0, 0.3)
    ri2, ts2, te2 = est.sampling(rays[:, :3].contiguous(), rays[:, 3:].contiguous(), render_step_size=0.01,
                                 alpha_fn=alpha_fn)
    ori, ots, ote = omarch.ray_marching(rays[:, :3].cpu().numpy(), rays[:, 3:].cpu().numpy(),
                                        scene_aabb=np.array(ROI, np.float32), grid_roi=np.array(ROI, np.float32),
                                        grid_binary=syn.analytic_grid("ball").numpy(), render_step_size=0.01,
                                        alpha_fn=lambda a, b, c: np.full_like(a, 0.3), near_plane=0.0, far_plane=1e10)
    assert np.array_equal(ri2.cpu().numpy(), ori) and np.array_equal(bits(ts2.cpu().numpy()), bits(ots))
