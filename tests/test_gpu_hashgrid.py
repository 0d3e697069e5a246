"""K2 parity: hash-grid forward / backward / second-order and SH vs the torch oracle
(tiny-cuda-nn itself is absent -> parity unpinned, oracle-defined)."""
import numpy as np
import pytest
import torch

from oracle import fields as of
from rise_sdf_b200 import tinycudann as tcnn

pytestmark = pytest.mark.gpu

CFG = {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 19,
       "base_resolution": 16, "per_level_scale": 1.447269237440378}


def make(cfg, table_scale=1.0, seed=0):
    enc = tcnn.Encoding(3, cfg).cuda()
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        enc.params.copy_(((torch.rand(enc.params.shape, generator=g) * 2 - 1) * table_scale).cuda())
    meta = of.HashGridMeta(cfg["n_levels"], 2, cfg["log2_hashmap_size"], cfg["base_resolution"], cfg["per_level_scale"])
    assert meta.n_params == enc.params.numel() and meta.res == enc.meta.res and meta.offset == enc.meta.offset
    return enc, meta


def points(S, seed=1):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(S, 3, generator=g)
    x[0] = 0.0; x[1] = 1.0; x[2] = torch.tensor([0.5, 0.25, 0.75])   # cell boundaries / box corners
    x[3] = torch.tensor([1.0, 0.0, 0.5])
    return x


@pytest.mark.parametrize("base,table_scale", [(16, 1.0), (32, 1.0), (16, 1e-4)])
def test_forward_matches_oracle(base, table_scale):
    cfg = dict(CFG, base_resolution=base)
    enc, meta = make(cfg, table_scale)
    assert meta.n_params == {16: 12599920, 32: 14533536}[base]
    x = points(4096)
    y = enc(x.cuda()).cpu()
    y0 = of.hash_encode(x, enc.params.detach().cpu(), meta)
    assert y.shape == (4096, 32)
    assert float((y - y0).abs().max()) <= 2e-6 * table_scale


def test_first_order_grads():
    cfg = dict(CFG, n_levels=8, log2_hashmap_size=12)
    enc, meta = make(cfg)
    x = points(2048)
    gy = torch.randn(2048, meta.n_output_dims, generator=torch.Generator().manual_seed(2))
    xc = x.cuda().requires_grad_(True)
    y = enc(xc)
    (y * gy.cuda()).sum().backward()
    x0 = x.clone().requires_grad_(True)
    t0 = enc.params.detach().cpu().clone().requires_grad_(True)
    (of.hash_encode(x0, t0, meta) * gy).sum().backward()
    assert float((enc.params.grad.cpu() - t0.grad).abs().max()) <= 1e-4 * float(t0.grad.abs().max())
    # d/dx is piecewise constant with 1/ulp-sized jumps at cell faces: compare away from faces
    sc = torch.tensor(meta.scale)
    pos = x[:, None, :] * sc[None, :, None] + 0.5
    frac = pos - pos.floor()
    interior = ((frac > 1e-3) & (frac < 1 - 1e-3)).all(-1).all(-1)
    err = (xc.grad.cpu() - x0.grad).abs()[interior].max()
    assert float(err) <= 1e-4 * float(x0.grad.abs().max())


def test_second_order_grads_match_autograd_oracle():
    """grad-of-grad: L = <c, d/dx sum(y*gy)> backpropagated to table, gy and x
    (what eikonal / normal-dependent losses need: models/geometry.py:224-228)."""
    cfg = dict(CFG, n_levels=6, log2_hashmap_size=10, base_resolution=4)
    enc, meta = make(cfg)
    S = 1024
    x = points(S, seed=5) * 0.98 + 0.01
    gen = torch.Generator().manual_seed(3)
    gy = torch.randn(S, meta.n_output_dims, generator=gen)
    c = torch.randn(S, 3, generator=gen)

    def run(xi, table, gyi, fn):
        y = fn(xi, table)
        (gx,) = torch.autograd.grad((y * gyi).sum(), xi, create_graph=True)
        loss = (gx * c.to(gx.device, gx.dtype)).sum()
        return torch.autograd.grad(loss, [table, gyi, xi], allow_unused=True)

    xc = x.cuda().requires_grad_(True)
    gyc = gy.cuda().requires_grad_(True)
    gt, ggy, gx2 = run(xc, enc.params, gyc, lambda a, b: enc(a))
    x0 = x.double().requires_grad_(True)
    t0 = enc.params.detach().cpu().double().requires_grad_(True)
    gy0 = gy.double().requires_grad_(True)
    rt, rgy, rx = run(x0, t0, gy0, lambda a, b: of.hash_encode(a, b, meta))
    assert float((gt.cpu() - rt).abs().max()) <= 2e-4 * float(rt.abs().max())
    assert float((ggy.cpu() - rgy).abs().max()) <= 2e-4 * float(rgy.abs().max())
    assert float((gx2.cpu() - rx).abs().max()) <= 5e-4 * float(rx.abs().max())


def test_table_grad_is_conservative():
    """size-independent property at full size: sum of the table gradient == sum of dL_dy
    (trilinear weights sum to one per level), 1M samples."""
    enc, meta = make(CFG)
    S = 1 << 20
    x = torch.rand(S, 3, device="cuda")
    gy = torch.rand(S, 32, device="cuda")
    y = enc(x)
    (y * gy).sum().backward()
    g = enc.params.grad.view(-1, 2)
    for l in (0, 4, 5, 15):
        a, b = meta.offset[l], meta.offset[l + 1]
        got = g[a:b].sum(0).double().cpu()
        want = gy[:, 2 * l:2 * l + 2].double().sum(0).cpu()
        assert float((got - want).abs().max()) <= 1e-3 * float(want.abs().max())


@pytest.mark.parametrize("degree", [4, 5])
def test_sh_fwd_bwd(degree):
    enc = tcnn.Encoding(3, {"otype": "SphericalHarmonics", "degree": degree}).cuda()
    g = torch.Generator().manual_seed(0)
    d = torch.nn.functional.normalize(torch.randn(4096, 3, generator=g), dim=-1)
    u = (d + 1) / 2
    uc = u.cuda().requires_grad_(True)
    out = enc(uc)
    ref_in = u.double().requires_grad_(True)
    ref = of.sh_encode(ref_in, degree)
    assert out.shape == (4096, degree * degree)
    assert float((out.cpu() - ref).abs().max()) <= 2e-6
    go = torch.randn(4096, degree * degree, generator=g)
    (out * go.cuda()).sum().backward()
    (ref * go.double()).sum().backward()
    assert float((uc.grad.cpu() - ref_in.grad).abs().max()) <= 1e-5 * float(ref_in.grad.abs().max())


def test_fd6_neighbours_bit_identical_to_six_forward_passes():
    """rsdf_hashgrid_fd6 == the reference's construction of the finite-difference neighbours
    (models/geometry.py:229-237: add offset, clamp to +-radius, scale to the unit cube) followed by the plain
    forward kernel: identical unit-cube coordinates and identical encodings, bit for bit."""
    from rise_sdf_b200 import tinycudann as tcnn
    from rise_sdf_b200.geometry import scale_anything
    enc = tcnn.Encoding(3, dict(otype="HashGrid", n_levels=16, n_features_per_level=2, log2_hashmap_size=19,
                                base_resolution=32, per_level_scale=1.447269237440378)).cuda()
    with torch.no_grad():
        enc.params.uniform_(-1.0, 1.0)
    g = torch.Generator().manual_seed(11)
    r = 1.5
    for S, eps in ((1000, 2 * r / 8180.0), (33, 0.01), (5000, 1e-3)):
        p = ((torch.rand(S, 3, generator=g) * 2 - 1) * r).cuda()
        p[:7] = torch.tensor([[r, -r, 0.3], [r, r, r], [-r, -r, -r], [r - 1e-4, 0, 0], [0, -r + 1e-5, 0], [0, 0, r], [0, 0, 0]])
        x01, y = tcnn.hashgrid_fd6(enc, p, eps, r)
        offsets = torch.as_tensor([[eps, 0.0, 0.0], [-eps, 0.0, 0.0], [0.0, eps, 0.0],
                                   [0.0, -eps, 0.0], [0.0, 0.0, eps], [0.0, 0.0, -eps]]).to(p)
        pd = scale_anything((p[:, None, :] + offsets).clamp(-r, r), (-r, r), (0, 1)).view(-1, 3)
        with torch.no_grad():
            y_ref = enc(pd.contiguous())
        assert torch.equal(x01, pd)
        assert torch.equal(y, y_ref)
