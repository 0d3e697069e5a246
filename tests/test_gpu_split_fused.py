"""The two fused split-sum stages against the op-by-op path they replace (the same host code the reference runs:
models/texture.py:330-377, models/split_mixed_occ.py:151-177,384-394, models/volrend.py:851-885), which in turn is pinned
to the reference's Python by tests/test_gpu_reference_host.py and to oracle/split.py by tests/test_gpu_splitsum.py /
tests/test_gpu_split_grads.py (those run with the fused kernels ON, the default).

Both sides are fp32 on the same device and differ only in the association order of a handful of adds/FMAs per sample,
so the bounds are a few ulps of the output scale; gradients summed over all samples (texels, network weights) get 1e-4
relative L2.
"""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
from helpers import rel_l2  # noqa: E402

pytestmark = pytest.mark.gpu


def _texture_and_light(base_res=64, seed=0):
    from rise_sdf_b200.light import EnvironmentLightMipCube
    from rise_sdf_b200.split_mixed_occ import split_mixed_occ_config
    from rise_sdf_b200.texture import VolumeMixedMipSplitOcc
    torch.manual_seed(seed)
    cfg = split_mixed_occ_config()
    cfg["light"]["envlight_config"]["base_res"] = base_res
    tex = VolumeMixedMipSplitOcc(cfg.texture).cuda()
    light = EnvironmentLightMipCube(cfg.light).cuda()
    return tex, light


def _shade_inputs(n, seed=1):
    g = torch.Generator().manual_seed(seed)
    feats = (torch.randn(n, 48, generator=g) * 0.5).cuda().requires_grad_(True)
    normals = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).cuda().requires_grad_(True)
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).cuda()
    pos = (torch.rand(n, 3, generator=g) * 2 - 1).cuda()
    return feats, dirs, normals, pos


@pytest.mark.parametrize("stage", [0, 1])
def test_fused_shade_vs_op_by_op(stage):
    tex, light = _texture_and_light()
    n = 20000
    feats, dirs, normals, pos = _shade_inputs(n)
    cot = torch.randn(n, 24 if stage else 7, generator=torch.Generator().manual_seed(5)).cuda()
    res = []
    for fused in (True, False):
        tex.fused_shade = fused
        for p in list(tex.parameters()) + [light.base, feats, normals]:
            p.grad = None
        light.build_mips()
        out = tex(feats, dirs, normals, pos, light, stage)
        (out * cot).sum().backward()
        grads = {"features": feats.grad.clone(), "normals": normals.grad.clone()}
        if stage:
            grads["emitter.base"] = light.base.grad.clone()
        for k, p in tex.named_parameters():
            if p.grad is not None and p.numel():
                grads[k] = p.grad.clone()
        res.append((out.detach(), grads))
    (a, ga), (b, gb) = res
    assert a.shape == b.shape == (n, 24 if stage else 7)
    assert float((a - b).abs().max()) <= 5e-6 * max(float(b.abs().max()), 1.0)   # measured 2.4e-6 (mip lerp order)
    # stage 0 never reads the roughness network: torch leaves its grads None, the fused node reports exact zeros
    assert set(gb) <= set(ga) and all(k.startswith("roughness_network") and stage == 0 for k in set(ga) - set(gb))
    for k in ga:
        if k not in gb:
            assert float(ga[k].abs().max()) == 0.0, k
            continue
        e = rel_l2(ga[k].cpu().numpy(), gb[k].cpu().numpy())
        assert e <= 1e-4, (k, e)
    if stage:
        assert float(gb["normals"].abs().max()) > 0 and float(gb["emitter.base"].abs().max()) > 0


def test_fused_shade_roughness_branches_and_clamps():
    """get_mip has two branches (roughness below / above MAX_ROUGHNESS) and the LUT clamps NoV: drive the raw outputs so
    that all of them are hit, straight through the kernel wrapper against torch autograd on the same formulas."""
    from rise_sdf_b200.split_shade import split_shade
    tex, light = _texture_and_light()
    n = 4096
    g = torch.Generator().manual_seed(3)
    mk = lambda c, s: (torch.randn(n, c, generator=g) * s).cuda().requires_grad_(True)
    raw6, rawr, raw2, raw3 = mk(6, 1.0), mk(1, 2.5), mk(2, 1.0), mk(3, 1.0)     # sigmoid(2.5 randn) covers (0.01, 0.99)
    normals = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).cuda().requires_grad_(True)
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).cuda()   # half of them face away: NoV < 0
    cot = torch.randn(n, 24, generator=g).cuda()
    res = []
    for fused in (True, False):
        for t in (raw6, rawr, raw2, raw3, normals, light.base):
            t.grad = None
        light.build_mips()
        if fused:
            out = split_shade(raw6, rawr, raw2, raw3, normals, dirs, light, tex.FG_LUT, 1)
        else:
            s = torch.sigmoid
            wi = -dirs
            wo = torch.sum(wi * normals, -1, keepdim=True) * normals * 2 - wi
            NoV = torch.sum(normals * wi, -1, keepdim=True)
            D, A, R, B, M, E = s(raw6[:, :3]), s(raw6[:, 3:]), s(rawr), s(raw2[:, :1]), s(raw2[:, 1:]), s(raw3)
            fg = tex._fg(NoV, R)
            sr = (0.04 * (1 - M) + M * A) * fg[:, 0:1] + fg[:, 1:2]
            SL = light.eval_mip(wo, specular=True, roughness=R)
            out = torch.cat([(1 - B) * D, B * E, B, (1 - M) * A * light.eval_mip(normals), sr * SL, sr, SL, A, M, R], -1)
        (out * cot).sum().backward()
        res.append((out.detach(), [t.grad.clone() for t in (raw6, rawr, raw2, raw3, normals, light.base)]))
    (a, ga), (b, gb) = res
    R = torch.sigmoid(rawr.detach())
    assert float((R < 0.08).float().mean()) > 0.02 and float((R > 0.5).float().mean()) > 0.2
    assert float((a - b).abs().max()) <= 5e-6 * max(float(b.abs().max()), 1.0)   # measured 2.4e-6 (mip lerp order)
    for name, x, y in zip(("albedo", "roughness", "metallic", "env", "normals", "base"), ga, gb):
        assert rel_l2(x.cpu().numpy(), y.cpu().numpy()) <= 1e-4, name


def _render_inputs(n_rays=3000, seed=0, cd=24):
    g = torch.Generator().manual_seed(seed)
    counts = torch.randint(0, 90, (n_rays,), generator=g)
    counts[::7] = 0                                            # rays without samples
    counts[5] = 300                                            # longer than several warp chunks
    ray_indices = torch.repeat_interleave(torch.arange(n_rays), counts).cuda()
    S = int(counts.sum())
    t0 = torch.rand(S, generator=g).cuda() * 3
    t1 = t0 + 0.02
    sdf = (torch.randn(S, generator=g) * 0.02).cuda().requires_grad_(True)
    normals = torch.nn.functional.normalize(torch.randn(S, 3, generator=g), dim=-1).cuda().requires_grad_(True)
    colors = torch.rand(S, cd, generator=g).cuda().requires_grad_(True)
    rays_d = torch.nn.functional.normalize(torch.randn(n_rays, 3, generator=g), dim=-1).cuda()
    return ray_indices, t0, t1, sdf, normals, colors, rays_d


@pytest.mark.parametrize("cd", [7, 24])
@pytest.mark.parametrize("ratio", [0.37, 1.0])
def test_fused_split_render_vs_op_by_op(cd, ratio):
    from rise_sdf_b200.nerfacc import accumulate_along_rays, pack_info, render_weight_from_alpha
    from rise_sdf_b200.split_mixed_occ import SplitMixedOCCModel, split_mixed_occ_config
    from rise_sdf_b200.split_shade import split_render
    cfg = split_mixed_occ_config()
    cfg["light"]["envlight_config"]["base_res"] = 32
    model = SplitMixedOCCModel(cfg).cuda()
    model.cos_anneal_ratio = ratio
    with torch.no_grad():
        model.variance.variance.fill_(0.5)
    ri, t0, t1, sdf, normals, colors, rays_d = _render_inputs(cd=cd)
    n_rays = rays_d.shape[0]
    g = torch.Generator().manual_seed(9)
    cot, cot_w = torch.randn(n_rays, cd + 6, generator=g).cuda(), torch.randn(ri.shape[0], generator=g).cuda()
    res = []
    for fused in (True, False):
        for t in (sdf, normals, colors, model.variance.variance):
            t.grad = None
        if fused:
            out, w, alpha = split_render(pack_info(ri, n_rays), rays_d, t0, t1, sdf, normals, colors, model.variance.inv_s,
                                         ratio)
        else:
            alpha = model.get_alpha(sdf, normals, rays_d[ri], (t1 - t0)[:, None])
            w, _ = render_weight_from_alpha(alpha, ray_indices=ri, n_rays=n_rays)
            kw = dict(ray_indices=ri, n_rays=n_rays)
            orient = torch.sum(rays_d[ri] * normals, dim=-1, keepdim=True).clamp(min=0)
            out = torch.cat([accumulate_along_rays(w, values=colors, **kw), accumulate_along_rays(w, values=normals, **kw),
                             accumulate_along_rays(w, values=None, **kw),
                             accumulate_along_rays(w, values=(t0 + t1)[:, None] / 2.0, **kw),
                             accumulate_along_rays(w, values=orient, **kw)], -1)
        ((out * cot).sum() + (w.view(-1) * cot_w).sum()).backward()
        res.append((out.detach(), w.detach().view(-1), alpha.detach().view(-1),
                    [t.grad.clone() for t in (sdf, normals, colors, model.variance.variance)]))
    (a, wa, aa, ga), (b, wb, ab, gb) = res
    assert float((aa - ab).abs().max()) <= 2e-6
    assert float((wa - wb).abs().max()) <= 2e-6
    assert float((a - b).abs().max()) <= 1e-5 * max(float(b.abs().max()), 1.0)
    for name, x, y in zip(("sdf", "normals", "colors", "variance"), ga, gb):
        assert rel_l2(x.cpu().numpy(), y.cpu().numpy()) <= 1e-4, name


def test_split_model_fused_stages_equal_op_by_op_training_step():
    """Whole model, training mode, stage 1: outputs and every parameter gradient with the fused stages on vs off."""
    from test_gpu_split_grads import split_model
    from rise_sdf_b200 import synthetic as syn
    from rise_sdf_b200.train import split_loss
    m = split_model(base_res=128)
    rays, rgb, fg, bg = (t.cuda() for t in syn.training_rays(192, seed=4))
    res = []
    for fused in (True, False):
        m.fused_render = m.texture.fused_shade = fused
        m.train()
        m.randomized = False
        m.background_color = bg
        m.zero_grad(set_to_none=True)
        m.emitter.base.grad = None
        torch.manual_seed(11)
        torch.cuda.manual_seed(11)
        m.emitter.build_mips()
        out = m(rays)
        loss, parts = split_loss(m, out, rgb, fg)
        loss.backward()
        grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None and p.numel()}
        grads["emitter.base"] = m.emitter.base.grad.clone()
        res.append(({k: v.detach().clone() for k, v in out.items() if torch.is_tensor(v)}, float(loss), grads))
    (a, la, ga), (b, lb, gb) = res
    assert torch.equal(a["ray_indices"], b["ray_indices"])
    for k in ("comp_rgb", "comp_rgb_phys", "comp_normal", "opacity", "depth", "comp_albedo", "comp_roughness",
              "comp_metallic", "normals_orientation_loss_map", "weights", "comp_rgb_full", "comp_rgb_phys_full"):
        e = float((a[k] - b[k]).abs().max())
        # measured 8e-5 .. 2e-4: ~1e-6 per alpha over a ray's scan; the reflection bounce then re-marches from
        # origin + depth * d, and its DISCRETE sample set turns a 1e-6 shift of depth into a visible change of `tr` on a few rays
        assert e <= 1e-3 * max(float(b[k].abs().max()), 1.0), (k, e)
    assert abs(la - lb) <= 1e-5 * abs(lb)
    assert set(ga) == set(gb)
    for k in gb:
        # both sides run the same fp32 FD-normal field; what differs is the summation order in shade/render -- and, through
        # the 1e-6 shift of depth it causes, the DISCRETE sample set of the reflection bounce on a few rays (see above),
        # which a scalar gradient like `variance` feels at the 3e-3 level (measured); a logic error is O(1)
        assert rel_l2(ga[k].cpu().numpy(), gb[k].cpu().numpy()) <= 1e-2, k


def test_fused_stages_edge_cases():
    """Empty sample sets, rays without samples, a single sample: the fused stages keep the shapes and zeros the op-by-op
    path produces (the reference's own guards: models/volrend.py:826-834, models/texture.py:331-332)."""
    from rise_sdf_b200.nerfacc import pack_info
    from rise_sdf_b200.split_shade import split_render, split_shade
    tex, light = _texture_and_light(base_res=64)          # 64 -> 32 -> 16: the smallest pyramid get_mip is defined for
    light.build_mips()
    z = lambda *s: torch.zeros(*s, device="cuda")
    out = split_shade(z(0, 6), z(0, 1), z(0, 2), z(0, 3), z(0, 3), z(0, 3), light, tex.FG_LUT, 1)
    assert out.shape == (0, 24)
    out0 = split_shade(z(0, 6), z(0, 1), z(0, 2), z(0, 3), z(0, 3), z(0, 3), light, tex.FG_LUT, 0)
    assert out0.shape == (0, 7)
    # three rays, only the middle one has (one) sample
    ri = torch.tensor([1], device="cuda")
    packed = pack_info(ri, 3)
    rays_d = torch.nn.functional.normalize(torch.randn(3, 3), dim=-1).cuda()
    sdf = torch.tensor([0.001], device="cuda", requires_grad=True)
    n = (-rays_d[1:2]).clone().requires_grad_(True)                  # facing the ray
    col = torch.rand(1, 24, device="cuda", requires_grad=True)
    inv_s = torch.tensor(148.4, device="cuda", requires_grad=True)
    acc, w, alpha = split_render(packed, rays_d, torch.tensor([1.0], device="cuda"), torch.tensor([1.02], device="cuda"), sdf, n,
                                 col, inv_s, 1.0)
    assert acc.shape == (3, 30) and float(acc[0].abs().sum()) == 0 and float(acc[2].abs().sum()) == 0
    assert torch.allclose(acc[1, :24], (w * col)[0]) and torch.allclose(acc[1, 27], w[0]) and float(acc[1, 29]) == 0.0
    assert 0.0 < float(alpha) <= 1.0 and torch.allclose(w, alpha)      # first sample of its ray: T = 1
    acc.sum().backward()
    for t in (sdf, n, col, inv_s):
        assert t.grad is not None and torch.isfinite(t.grad).all()
