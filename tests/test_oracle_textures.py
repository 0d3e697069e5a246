"""Pins oracle/textures.py against the reference's own compiled cubemap kernels
(tests/golden/cubemap_prefilter.npz, produced on a B200 by tests/golden/make_golden.py from
oracle/_ref/renderutils_plugin.so) and checks the lookup restatement's invariants."""
import os

import numpy as np
import torch

from oracle import textures as ot

GOLD = os.path.join(os.path.dirname(__file__), "golden", "cubemap_prefilter.npz")


def test_diffuse_matches_reference_kernel():
    z = np.load(GOLD)
    cube, go = torch.from_numpy(z["cube16"]).double(), torch.from_numpy(z["go16"]).double()
    out = ot.diffuse_cubemap(cube)
    assert float((out - torch.from_numpy(z["diffuse_fwd"])).abs().max()) <= 1e-5 * float(out.abs().max())
    W = torch.from_numpy(ot.diffuse_weights(16))
    g = (W.T @ go.reshape(-1, 3)).reshape(6, 16, 16, 3)
    assert float((g - torch.from_numpy(z["diffuse_bwd"])).abs().max()) <= 1e-4 * float(g.abs().max())


def test_specular_bounds_and_filter_match_reference_kernel():
    z = np.load(GOLD)
    for N in (16, 32):
        rough, cut = float(z[f"spec{N}_rough"]), float(z[f"spec{N}_cut"])
        assert ot.ndf_cutoff(rough, 0.99) == cut
        b = ot.specular_bounds(N, cut)
        assert np.array_equal(b.astype(np.int16), z[f"spec{N}_bounds"])
        W = ot.specular_weights(N, rough, np.float32(cut), b)
        cube = z[f"spec{N}_cube"].reshape(-1, 3).astype(np.float64)
        raw = z[f"spec{N}_raw"].reshape(-1, 4)
        # what the caller sees (ops.py:458): the NORMALISED filter, pinned to 5e-5.  The raw
        # weight sum is only loosely comparable: at roughness 0.08 the reference evaluates
        # d = 1 - c^2 (1 - alpha^4) in fp32 next to c ~ 1 (cubemap.cu:174-179), a ~1e-3 cancellation
        # error that is common to numerator and denominator and divides out.
        np.testing.assert_allclose((W @ cube) / W.sum(1)[:, None], raw[:, :3] / raw[:, 3:], rtol=5e-5, atol=1e-7)
        np.testing.assert_allclose(W.sum(1), raw[:, 3], rtol=2e-2)


def test_cube_lookup_invariants():
    # constant map -> constant everywhere (weights renormalised at corners), incl. edges/corners
    tex = torch.full((6, 8, 8, 3), 0.7, dtype=torch.float64)
    d = torch.nn.functional.normalize(torch.randn(2000, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(0)), dim=-1)
    d[:3] = torch.nn.functional.normalize(torch.tensor([[1.0, 1, 1], [1, -1, 0.999], [-1, 1, 0]], dtype=torch.float64), dim=-1)
    assert float((ot.cube_linear(tex, d) - 0.7).abs().max()) <= 1e-12
    # texel centres reproduce the texel; face convention == cube_to_dir (light_utils.py:85-92)
    N = 8
    tex = torch.arange(6 * N * N, dtype=torch.float64).view(6, N, N, 1)
    dirs = torch.from_numpy(ot.texel_dirs(N)).reshape(-1, 3).double()
    assert float((ot.cube_linear(tex, dirs)[:, 0] - torch.arange(6 * N * N)).abs().max()) <= 1e-4
    # continuity across a face edge
    a = torch.tensor([[1.0, 0.3, 0.999999]], dtype=torch.float64)
    b = torch.tensor([[0.999999, 0.3, 1.0]], dtype=torch.float64)
    t = torch.rand(6, N, N, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    assert float((ot.cube_linear(t, a) - ot.cube_linear(t, b)).abs().max()) <= 1e-4
    # mip blend: bias exactly on a level returns that level; clamped outside
    lv = [torch.rand(6, r, r, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(r)) for r in (16, 8, 4)]
    for k, bias in [(0, -3.0), (0, 0.0), (1, 1.0), (2, 2.0), (2, 9.0)]:
        o = ot.cube_sample(lv, d, torch.full((2000,), bias, dtype=torch.float64))
        assert float((o - ot.cube_linear(lv[k], d)).abs().max()) <= 1e-12


def test_tex2d_clamp_and_centres():
    tex = torch.arange(12, dtype=torch.float64).view(3, 4, 1)
    uv = torch.tensor([[(x + 0.5) / 4, (y + 0.5) / 3] for y in range(3) for x in range(4)], dtype=torch.float64)
    assert torch.allclose(ot.tex2d(tex, uv)[:, 0], torch.arange(12, dtype=torch.float64))
    assert float(ot.tex2d(tex, torch.tensor([[0.0, 0.0], [1.0, 1.0]], dtype=torch.float64))[0, 0]) == 0.0
