"""Oracle self-checks for the march / scan restatement (CPU)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import march
from rise_sdf_b200 import synthetic as syn

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROI = np.array([-1.5] * 3 + [1.5] * 3, np.float32)


def test_docstring_known_answers():
    # lib/nerfacc/vol_rendering.py:430-434 and :493-500
    a = np.array([0.4, 0.8, 0.1, 0.8, 0.1, 0.0, 0.9], np.float32)
    pk = march.pack_info([0, 0, 0, 1, 1, 2, 2], 3)
    w, T = march.weight_from_alpha(pk, a)
    np.testing.assert_allclose(w, [0.4, 0.48, 0.012, 0.8, 0.02, 0.0, 0.9], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(T, [1.0, 0.6, 0.12, 1.0, 0.2, 1.0, 1.0], rtol=1e-6)
    vis = (T >= 0.3) & (a >= 0.2)
    assert vis.tolist() == [True, True, False, True, False, False, True]


def test_pack_info_shape_example():
    pk = march.pack_info([0, 0, 2, 2, 2], 4)
    assert pk.tolist() == [[0, 2], [2, 0], [2, 3], [5, 0]]


def test_march_all_ones_is_uniform():
    rays, _, _, _ = syn.training_rays(256, seed=3)
    o, d = rays[:, :3].numpy(), rays[:, 3:].numpy()
    tmin, tmax = march.ray_aabb_intersect(o, d, ROI)
    step = 1.732 * 2 * 1.5 / 128
    pk, ri, ts, te = march.ray_marching_raw(o, d, tmin, tmax, ROI, np.ones((128, 128, 128), bool), step)
    assert pk[:, 1].sum() == len(ri) and np.all(np.diff(ri) >= 0)
    hit = tmax < 1e9
    assert hit.sum() > 50 and np.all(pk[~hit, 1] == 0)
    # expected count: midpoints t_min + (k+.5)step < t_max
    exp = np.ceil((tmax[hit] - tmin[hit]) / step - 0.5)
    assert np.all(np.abs(pk[hit, 1] - exp) <= 1)
    # contiguous intervals along each ray
    for r in np.nonzero(hit)[0][:20]:
        b, n = pk[r]
        assert np.all(ts[b + 1:b + n] == te[b:b + n - 1])
        assert ts[b] == tmin[r]


def test_march_zero_grid_and_voxel():
    rays, _, _, _ = syn.training_rays(512, seed=4)
    o, d = rays[:, :3].numpy(), rays[:, 3:].numpy()
    tmin, tmax = march.ray_aabb_intersect(o, d, ROI)
    pk, ri, ts, te = march.ray_marching_raw(o, d, tmin, tmax, ROI, np.zeros((128,) * 3, bool), 0.005)
    assert len(ri) == 0 and pk[:, 1].sum() == 0
    g = syn.analytic_grid("ball").numpy()
    pk, ri, ts, te = march.ray_marching_raw(o, d, tmin, tmax, ROI, g, 0.005)
    mid = (ts + te) * 0.5
    pts = o[ri] + d[ri] * mid[:, None]
    assert np.all(march.grid_query(pts, ROI, g))            # every sample centre sits in an occupied cell
    assert np.all(np.linalg.norm(pts, axis=1) < 1.1 + 0.05)


def test_aabb_edge_cases():
    o = np.array([[0, 0, 0], [5, 5, 5], [0, 0, -4], [0, 0, -4]], np.float32)
    d = np.array([[0, 0, 1], [0, 0, 1], [0, 0, 1], [1, 0, 0]], np.float32)
    tmin, tmax = march.ray_aabb_intersect(o, d, ROI)
    assert tmin[0] == 0.0 and tmax[0] == 1.5        # origin inside: near clamped to 0
    assert tmin[1] == 1e10 and tmax[1] == 1e10      # miss
    assert tmin[2] == 2.5 and tmax[2] == 5.5        # axis-parallel hit (inv_dir = inf on x, y)
    assert tmin[3] == 1e10                          # axis-parallel miss


def test_scan_backward_matches_autograd():
    from oracle import fields
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 40, size=50)
    ri = np.repeat(np.arange(50), counts)
    a = rng.uniform(0, 0.9, size=len(ri)).astype(np.float32)
    pk = march.pack_info(ri, 50)
    w, T = march.weight_from_alpha(pk, a)
    gw = rng.normal(size=len(ri)).astype(np.float32)
    ga = march.weight_from_alpha_backward(pk, a, w, gw)
    at = torch.tensor(a, dtype=torch.float64, requires_grad=True)
    wt, _ = fields.render_weight_from_alpha(at, torch.tensor(ri), 50)
    (wt * torch.tensor(gw, dtype=torch.float64)).sum().backward()
    np.testing.assert_allclose(ga, at.grad.numpy(), rtol=2e-3, atol=2e-4)
    np.testing.assert_allclose(w, wt.detach().numpy(), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "march_*.npz"))))
def test_oracle_vs_reference_golden(path):
    """tests/golden/march_*.npz = outputs of the reference's own compiled kernel (oracle/_ref,
    built from lib/nerfacc/cuda/csrc) on a B200 (tests/golden/make_golden.py)."""
    z = np.load(path)
    tmin, tmax = march.ray_aabb_intersect(z["rays_o"], z["rays_d"], z["roi"])

    def canon(a):   # 0/0 on a slab plane: x86 and the GPU emit different NaN payloads
        a = a.copy(); a[np.isnan(a)] = np.float32(np.nan); return a.view(np.uint32)
    assert np.array_equal(canon(tmin), canon(z["t_min"]))
    assert np.array_equal(canon(tmax), canon(z["t_max"]))
    # scan known-answers from the reference's naive kernels (render_weight.cu:86-154)
    w, T = march.weight_from_alpha(z["packed_info"], z["alphas"])
    assert np.array_equal(w.view(np.uint32), z["weights"].view(np.uint32))
    assert np.array_equal(T.view(np.uint32), z["trans"].view(np.uint32))
    ga = march.weight_from_alpha_backward(z["packed_info"], z["alphas"], w, z["grad_weights"])
    np.testing.assert_allclose(ga, z["grad_alphas"], rtol=1e-5, atol=1e-6)
    pk, ri, ts, te = march.ray_marching_raw(z["rays_o"], z["rays_d"], z["t_min"], z["t_max"], z["roi"],
                                            z["grid"], float(z["step"]), float(z["cone"]) if "cone" in z else 0.0)
    assert np.array_equal(pk, z["packed_info"])
    assert np.array_equal(ri, z["ray_indices"])
    assert np.array_equal(ts.view(np.uint32), z["t_starts"].view(np.uint32))
    assert np.array_equal(te.view(np.uint32), z["t_ends"].view(np.uint32))
