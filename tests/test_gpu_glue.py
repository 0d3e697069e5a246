"""csrc/glue.cu: the elementwise stages around the big kernels, each against the torch expression of the reference it
replaces (values and gradients)."""
import pytest
import torch
import torch.nn.functional as F

from rise_sdf_b200 import glue
from rise_sdf_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _ref_freq(x, n, scale, offset, mask):
    """models/network_utils.py:27-33"""
    out = []
    x = x * scale + offset
    for k in range(n):
        for func in (torch.sin, torch.cos):
            out += [func((2.0 ** k) * x) * mask[k]]
    return torch.cat(out, -1)


@pytest.mark.parametrize("masked", [False, True])
def test_frequency_encoding(masked):
    from rise_sdf_b200.network_utils import VanillaFrequency
    g = torch.Generator().manual_seed(0)
    x = ((torch.rand(5001, 3, generator=g) * 2 - 1) * 1.5).cuda()
    enc = VanillaFrequency(3, {"n_frequencies": 6, "n_masking_step": 100 if masked else 0, "x_scale": 1.3 if masked else 1.0,
                               "x_offset": -0.2 if masked else 0.0})
    enc.update_step(0, 37)
    xc = x.clone().requires_grad_(True)
    out = enc(xc)
    xr = x.clone().requires_grad_(True)
    ref = _ref_freq(xr, 6, enc.x_scale, enc.x_offset, enc.mask.cuda())
    assert out.shape == (5001, 36) and torch.equal(out, ref)                # same sinf / cosf: bit-identical
    go = torch.randn(5001, 36, generator=g).cuda()
    (out * go).sum().backward(); (ref * go).sum().backward()
    assert float((xc.grad - xr.grad).abs().max()) <= 1e-5 * float(xr.grad.abs().max())
    if masked:
        assert 0 < float(enc.mask.min()) < 1 or float(enc.mask.min()) == 0


@pytest.mark.parametrize("srgb", [False, True])
def test_composite_epilogue(srgb):
    from rise_sdf_b200.light import rgb_to_srgb
    g = torch.Generator().manual_seed(1)
    rgb = (torch.rand(4097, 3, generator=g) * 1.4 - 0.1).cuda()             # some values clamp on both sides
    rgb[:50] *= 0.003                                                        # the linear toe of the sRGB curve
    op = torch.rand(4097, 1, generator=g).cuda()
    bg = torch.rand(3, generator=g).cuda()
    a, o = rgb.clone().requires_grad_(True), op.clone().requires_grad_(True)
    out = glue.composite(a, o, bg, srgb=srgb)
    b, p = rgb.clone().requires_grad_(True), op.clone().requires_grad_(True)
    ref = b + bg[None, :] * (1.0 - p)
    if srgb:
        ref = rgb_to_srgb(ref).clamp(0, 1)
    assert float((out - ref).abs().max()) <= (2e-7 if srgb else 0.0)
    go = torch.randn(4097, 3, generator=g).cuda()
    (out * go).sum().backward(); (ref * go).sum().backward()
    assert float((a.grad - b.grad).abs().max()) <= 2e-5 * float(b.grad.abs().max())
    assert float((o.grad - p.grad).abs().max()) <= 2e-5 * float(p.grad.abs().max())


def test_ray_loss_terms():
    """systems/neus.py:103 F.mse_loss(full[valid], rgb[valid]) and :123-125 BCE on the clamped opacity."""
    from rise_sdf_b200.train import binary_cross_entropy
    g = torch.Generator().manual_seed(2)
    n = 8192
    full = torch.rand(n, 3, generator=g).cuda()
    op = torch.rand(n, 1, generator=g)
    op[::7] = 0.0
    op[1::11] = 1.0
    op[2::13] = 5e-4
    op = op.cuda()
    tgt, fg = torch.rand(n, 3, generator=g).cuda(), (torch.rand(n, generator=g) < 0.5).float().cuda()
    a, o = full.clone().requires_grad_(True), op.clone().requires_grad_(True)
    mse, bce = glue.ray_loss_terms(a, o, tgt, fg)
    b, p = full.clone().requires_grad_(True), op.clone().requires_grad_(True)
    valid = p[:, 0] > 0
    rmse = F.mse_loss(b[valid], tgt[valid])
    rbce = binary_cross_entropy(torch.clamp(p.squeeze(-1), 1e-3, 1 - 1e-3), fg)
    assert abs(float(mse) - float(rmse)) <= 2e-6 * float(rmse) and abs(float(bce) - float(rbce)) <= 2e-6 * float(rbce)
    (3.0 * mse + 0.7 * bce).backward(); (3.0 * rmse + 0.7 * rbce).backward()
    assert float((a.grad - b.grad).abs().max()) <= 1e-5 * float(b.grad.abs().max())
    assert float((o.grad - p.grad).abs().max()) <= 1e-5 * float(p.grad.abs().max())
    m2, b2 = glue.ray_loss_terms(full, op, tgt, fg)
    assert float(m2) == float(mse) and float(b2) == float(bce)              # fixed-order reduction: bit-reproducible


def test_get_rays_and_ray_sampler():
    """models/ray_utils.py:32-56 + F.normalize (systems/split_occ.py:103) on per-ray (image, x, y) draws."""
    poses, dirs = syn.camera_poses(), syn.ray_directions()
    g = torch.Generator().manual_seed(3)
    n = 6000
    index = torch.randint(0, 100, (n,), generator=g)
    x, y = torch.randint(0, 800, (n,), generator=g), torch.randint(0, 800, (n,), generator=g)
    ro, rd = syn.get_rays(dirs[y, x], poses[index])
    ref = torch.cat([ro, F.normalize(rd, p=2, dim=-1)], -1)
    got = glue.get_rays(dirs.cuda(), poses.cuda(), index.cuda(), x.cuda(), y.cuda())
    assert got.shape == (n, 6) and float((got.cpu() - ref).abs().max()) <= 5e-7
    images = torch.rand(4, 800, 800, 3, generator=g).cuda()
    masks = (torch.rand(4, 800, 800, generator=g) < 0.5).cuda()
    s = glue.RaySampler(dirs.cuda(), poses[:4].cuda(), images, masks, train_num_rays=256, max_train_num_rays=4096,
                        num_samples_per_ray=1024, dynamic=True, generator=torch.Generator("cuda").manual_seed(5))
    rays, rgb, fg, bg = s.sample()
    assert rays.shape == (256, 6) and rgb.shape == (256, 3) and fg.shape == (256,) and bg.shape == (3,)
    assert float((rays[:, 3:].norm(dim=-1) - 1).abs().max()) <= 1e-6
    # dynamic ray count (systems/split_occ.py:159-161): 256 rays x 1024 target samples, the step produced 100 per ray
    assert s.update_ray_count(256 * 100) == min(int(256 * 0.9 + int(256 * (256 * 1024 / 25600)) * 0.1), 4096)
    for _ in range(60):
        s.update_ray_count(s.train_num_rays * 100)
    assert s.train_num_rays <= 4096 and s.train_num_rays > 2000


def test_occupancy_update_kernels_equal_the_torch_form():
    """lib/nerfacc/grid.py:196-239 with duplicate indices: the result is the maximum over the duplicates."""
    from rise_sdf_b200 import nerfacc as rn
    est = rn.OccGridEstimator(torch.tensor([-1.5] * 3 + [1.5] * 3), 128).cuda().train()
    g = torch.Generator().manual_seed(4)
    fn = lambda p: 0.02 * torch.exp(-4.0 * (p.norm(dim=-1, keepdim=True) - 0.8).abs())
    est._update(0, fn, occ_thre=0.001, jitter=torch.rand(128 ** 3, 3, generator=g))
    occ0 = est.occs.clone()
    assert 0.05 < float(est.binaries.float().mean()) < 0.9
    assert torch.equal(rn.pack_bits(est.binaries), est.bits)                 # packed grid written by the same kernel
    # the cells `_update` will draw (same generator, same calls: lib/nerfacc/grid.py:205-215)
    torch.cuda.manual_seed(7)
    n = 128 ** 3 // 4
    uniform = torch.randint(128 ** 3, (n,), device="cuda")
    occupied = torch.nonzero(est.binaries.flatten())[:, 0]
    if n < len(occupied):
        occupied = occupied[torch.randint(len(occupied), (n,), device="cuda")]
    flat = torch.cat([uniform, occupied])
    assert int(flat.numel() - flat.unique().numel()) > 1000                   # duplicates are the point of this test
    torch.cuda.manual_seed(7)
    seen = {}
    def fn2(p):
        seen["x"] = p
        return fn(p) * 1.5
    est._update(256, fn2, occ_thre=0.001)                                     # quarter of the cells + the occupied ones
    x = seen["x"]
    assert x.shape[0] == flat.shape[0]
    upd = torch.zeros_like(occ0).scatter_reduce(0, flat, fn2(x).reshape(-1), "amax", include_self=True)
    touched = torch.zeros_like(occ0, dtype=torch.bool)
    touched[flat] = True
    want = torch.where(touched, torch.maximum(occ0 * 0.95, upd), occ0)
    assert float((est.occs - want).abs().max()) == 0.0
    thr = torch.clamp(est.occs.mean(), max=0.001)
    assert int((est.binaries.flatten() != (est.occs > thr)).sum()) <= 4       # cells AT the threshold (sum order of the mean)
    assert torch.equal(rn.pack_bits(est.binaries), est.bits)


@pytest.mark.parametrize("ratio", [0.0, 0.37, 1.0])
def test_neus_alpha_kernel_vs_torch_chain(ratio):
    """get_alpha (models/neus.py:128-150) as one launch, same op order as the torch chain.  alpha = (p + 1e-5) / (c + 1e-5)
    with p = sigmoid(a) - sigmoid(b) a cancellation, so a last-bit difference in one sigmoid shows up as a few ulps of
    alpha: bound 1e-6 absolute (measured 3.3e-7), and most samples bit-identical."""
    from rise_sdf_b200.neus import NeuSModel, neus_alpha, neus_blender_config
    model = NeuSModel(neus_blender_config()).cuda()
    model.cos_anneal_ratio = ratio
    g = torch.Generator().manual_seed(3)
    n = 200000
    sdf = (torch.randn(n, generator=g) * 0.02).cuda()
    normal = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).cuda()
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).cuda()
    dists = (torch.rand(n, generator=g) * 0.01 + 0.001).cuda()
    with torch.no_grad():
        got = neus_alpha(sdf, normal, dirs, dists, model.variance.inv_s, ratio)
    with torch.enable_grad():                       # the torch chain (get_alpha takes the kernel only under no_grad)
        ref = model.get_alpha(sdf, normal, dirs, dists).detach()
    assert got.shape == ref.shape == (n,)
    diff = (got - ref).abs()
    print("bit-identical fraction", float((diff == 0).float().mean()), "max", float(diff.max()))
    assert float(diff.max()) <= 1e-6
    assert float((diff == 0).float().mean()) >= 0.5


def test_visibility_round_kernels_vs_torch():
    """One front-to-back round: lens / candidate list / scatter against the index arithmetic they replace."""
    from rise_sdf_b200 import _lib as L
    g = torch.Generator().manual_seed(5)
    n_rays = 3000
    count = torch.randint(0, 300, (n_rays,), generator=g)
    count[::9] = 0
    base = torch.cumsum(count, 0) - count
    S0 = int(count.sum())
    packed = torch.stack([base, count], 1).int().cuda()
    ts = torch.rand(S0, generator=g).cuda()
    te = ts + 0.01
    T = torch.rand(S0, generator=g).cuda() * 3e-4           # about a third below the threshold
    for done, chunk in ((0, 64), (64, 128), (192, 1 << 20)):
        lens = torch.empty(n_rays, dtype=torch.int64, device="cuda")
        L.call("rsdf_vis_round_lens", L.ptr(packed), L.ptr(T if done else None), done, min(chunk, 2 ** 30), 1e-4, n_rays, S0,
               L.ptr(lens), L.stream())
        c, b = count.cuda(), base.cuda()
        active = c > done
        if done:
            active &= T[(b + done).clamp(max=S0 - 1)] >= 1e-4
        want = torch.where(active, (c - done).clamp(max=chunk), torch.zeros_like(c))
        assert torch.equal(lens, want)
        csum = torch.cumsum(lens, 0)
        total = int(csum[-1])
        idx = torch.empty(total, dtype=torch.int64, device="cuda")
        ts_sel, te_sel = torch.empty(total, device="cuda"), torch.empty(total, device="cuda")
        ri_sel = torch.empty(total, dtype=torch.int64, device="cuda")
        L.call("rsdf_vis_round_fill", L.ptr(packed), L.ptr(lens), L.ptr(csum), done, n_rays, L.ptr(ts), L.ptr(te), L.ptr(idx),
               L.ptr(ts_sel), L.ptr(te_sel), L.ptr(ri_sel), L.stream())
        ray_of = torch.repeat_interleave(torch.arange(n_rays, device="cuda"), want)
        first = csum - want
        want_idx = (b + done)[ray_of] + (torch.arange(total, device="cuda") - first[ray_of])
        assert torch.equal(idx, want_idx) and torch.equal(ri_sel, ray_of)
        assert torch.equal(ts_sel, ts[want_idx]) and torch.equal(te_sel, te[want_idx])
        a = torch.rand(total, device="cuda")
        alphas, rows = torch.zeros(S0, device="cuda"), torch.full((S0,), -1, dtype=torch.int64, device="cuda")
        L.call("rsdf_vis_round_scatter", L.ptr(idx), L.ptr(a), 1000, total, L.ptr(alphas), L.ptr(rows), L.stream())
        wa, wr = torch.zeros(S0, device="cuda"), torch.full((S0,), -1, dtype=torch.int64, device="cuda")
        wa[want_idx] = a
        wr[want_idx] = torch.arange(1000, 1000 + total, device="cuda")
        assert torch.equal(alphas, wa) and torch.equal(rows, wr)
