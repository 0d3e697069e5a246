"""CPU restatement of the split-sum (`split-mixed-occ`) render.  TEST INFRASTRUCTURE ONLY.

Follows models/split_mixed_occ.py:179-443 (forward_, compute_indirect_radiance),
models/texture.py:292-427 (VolumeMixedMipSplitOcc.forward / secondary_shading /
secondary_shading_pbr), lib/pbr/light.py:169-206 (build_mips / get_mip / eval_mip),
lib/pbr/utils/light_utils.py:71-74 (avg-pool mip), models/volrend.py:18-127,739-895,
lib/pbr/utils/nvdiffrecmc_util.py:95-103 (rgb_to_srgb), models/geometry.py:229-244,304-318
(finite-difference normals with the progressive eps).  Third-party pieces (tcnn, nerfacc 0.5.3,
nvdiffrast) are restated as in oracle/fields.py / oracle/textures.py (PARITY UNPINNED there).
`forward` = the no-grad eval / relighting render; `forward_train` = the differentiable training render
(curvature probe models/geometry.py:246-282, normal-orientation map models/split_mixed_occ.py:384-394, the
reflection bounce with gradients into the secondary origin / direction) and `loss` = systems/split_occ.py:150-237.
Both accept `dtype=torch.float64`: everything downstream of the fp32 sample positions / cell lookups is then
evaluated in double, the rounding-free yardstick the GPU gradient tests measure against.
tests/test_reference_host_cpu.py pins this file to the reference's own Python (run on the CPU over oracle shims).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import fields, march
from . import textures as tx

MIN_ROUGHNESS, MAX_ROUGHNESS, LIGHT_MIN_RES = 0.08, 0.5, 16


class SplitParams:
    def __init__(self, table, meta, geo_mlp, nets, variance, fg_lut, base, radius=1.5, level_mask=None,
                 fd_eps=None):
        self.table, self.meta, self.geo_mlp, self.nets = table, meta, geo_mlp, nets   # nets: dict name -> layers
        self.variance, self.fg_lut, self.base, self.radius = variance, fg_lut, base, radius
        self.level_mask, self.fd_eps = level_mask, fd_eps
        self.specular, self.diffuse = None, None

    @property
    def inv_s(self):
        return torch.exp(self.variance * 10.0)

    def named_tensors(self):
        """name (as in the model's state dict) -> learnable tensor"""
        out = {"geometry.encoding.encoding.encoding.params": self.table, "variance.variance": self.variance,
               "emitter.base": self.base}
        for i, layer in enumerate(self.geo_mlp):
            for k, v in layer.items():
                out[f"geometry.network.layers.{2 * i}.{k}"] = v
        for n, net in self.nets.items():
            for i, layer in enumerate(net):
                for k, v in layer.items():
                    out[f"texture.{n}_network.layers.{2 * i}.{k}"] = v
        return out

    def to(self, dtype):
        """Copy with every learnable tensor cast to `dtype` (fresh leaves)."""
        c = lambda t: t.detach().to(dtype).clone()
        P = SplitParams(c(self.table), self.meta, [{k: c(v) for k, v in l.items()} for l in self.geo_mlp],
                        {n: [{k: c(v) for k, v in l.items()} for l in net] for n, net in self.nets.items()},
                        c(self.variance), c(self.fg_lut), c(self.base), self.radius,
                        None if self.level_mask is None else self.level_mask.to(dtype), self.fd_eps)
        if self.specular is not None:
            P.specular, P.diffuse = [c(t) for t in self.specular], c(self.diffuse)
        return P


class _CubemapMip(torch.autograd.Function):
    """lib/pbr/utils/light_utils.py:94-109: 2x2 average pool whose backward is the reference's own -- a cube-filtered
    upsample of 0.25 * dout at the fine level's texel directions -- NOT the adjoint of the pooling."""

    @staticmethod
    def forward(ctx, cubemap):
        x = cubemap.permute(0, 3, 1, 2)
        return F.avg_pool2d(x, (2, 2)).permute(0, 2, 3, 1).contiguous()

    @staticmethod
    def backward(ctx, dout):
        res = dout.shape[1] * 2
        dirs = torch.from_numpy(tx.texel_dirs(res)).reshape(-1, 3).to(dout.dtype)
        return tx.cube_sample([dout * 0.25], dirs).reshape(6, res, res, dout.shape[-1])


def cubemap_mip(x):
    return _CubemapMip.apply(x)


def build_mips(P, cutoff=0.99):
    spec = [P.base]
    while spec[-1].shape[1] > LIGHT_MIN_RES:
        spec.append(cubemap_mip(spec[-1]))
    P.diffuse = tx.diffuse_cubemap(spec[-1])
    n = len(spec)
    for idx in range(n - 1):
        rough = (idx / (n - 2)) * (MAX_ROUGHNESS - MIN_ROUGHNESS) + MIN_ROUGHNESS
        spec[idx] = tx.specular_cubemap(spec[idx], rough, cutoff)
    spec[-1] = tx.specular_cubemap(spec[-1], 1.0, cutoff)
    P.specular = spec


def get_mip(P, roughness):
    n = len(P.specular)
    return torch.where(roughness < MAX_ROUGHNESS,
                       (torch.clamp(roughness, MIN_ROUGHNESS, MAX_ROUGHNESS) - MIN_ROUGHNESS)
                       / (MAX_ROUGHNESS - MIN_ROUGHNESS) * (n - 2),
                       (torch.clamp(roughness, MAX_ROUGHNESS, 1.0) - MAX_ROUGHNESS) / (1.0 - MAX_ROUGHNESS) + n - 2)


def eval_mip(P, directions, specular=False, roughness=None):
    if specular:
        return tx.cube_sample(P.specular, directions, get_mip(P, roughness)[:, 0])
    return tx.cube_sample([P.diffuse], directions)


def _field(P, points, dtype):
    """network(encoding(contract(points))) (models/geometry.py:214-217); the cell lookup stays fp32-faithful for
    fp32 points, everything after it runs in `dtype`."""
    x01 = fields.scale_to_unit(points, P.radius)
    enc = fields.hash_encode(x01, P.table, P.meta, dtype=dtype)
    if P.level_mask is not None:
        enc = enc * P.level_mask.to(dtype)
    return fields.vanilla_mlp(torch.cat([(x01 * 2.0 - 1.0).to(dtype), enc], -1), P.geo_mlp, "softplus100")


def geometry(P, points, with_grad=True, dtype=None, laplace_dirs=None):
    """VolumeSDF.forward with grad_type='finite_difference' (models/geometry.py:206-292).  `laplace_dirs` [S,3]
    (the `torch.rand_like(points)` draw of :253) adds the curvature probe and returns a fourth value."""
    dtype = dtype or points.dtype
    out = _field(P, points, dtype)
    if not with_grad:
        return out[:, 0], None, out
    eps = P.fd_eps
    offs = torch.tensor([[eps, 0, 0], [-eps, 0, 0], [0, eps, 0], [0, -eps, 0], [0, 0, eps], [0, 0, -eps]],
                        dtype=points.dtype)
    pd = (points[:, None, :] + offs).clamp(-P.radius, P.radius)
    sd = _field(P, pd.view(-1, 3), dtype)[:, 0].view(-1, 6)
    grad = fields.cuda_scalar_div(0.5 * (sd[:, 0::2] - sd[:, 1::2]), eps)
    if laplace_dirs is None:
        return out[:, 0], grad, out
    eps_c = 1e-4
    rand_directions = F.normalize(laplace_dirs.to(dtype), dim=-1, eps=1e-6)
    normal = F.normalize(grad, dim=-1, eps=1e-6)
    tangent = torch.cross(normal, rand_directions, dim=-1)
    pt = points + eps_c * tangent
    if not pt.requires_grad:
        pt = pt.requires_grad_(True)
    sdf_t = _field(P, pt, dtype)[:, 0]
    (grad_t,) = torch.autograd.grad(sdf_t, pt, torch.ones_like(sdf_t), create_graph=True, retain_graph=True)
    dot = torch.sum(F.normalize(grad, dim=-1, eps=1e-6) * F.normalize(grad_t, dim=-1, eps=1e-6), dim=-1)
    laplace = torch.acos(torch.clamp(dot, -1.0 + 1e-6, 1.0 - 1e-6)) / np.pi
    return out[:, 0], grad, out, laplace


def _fg(P, NoV, roughness):
    uv = torch.cat([NoV.clamp(0.0, 1.0), roughness.clamp(0.0, 1.0)], -1)
    return tx.tex2d(P.fg_lut[0], uv)


def texture_forward(P, features, dirs, normals, positions, stage):
    dt = features.dtype
    wi = -dirs
    wo = torch.sum(wi * normals, -1, keepdim=True) * normals * 2 - wi
    NoV = torch.sum(normals * wi, -1, keepdim=True)
    inp = torch.cat([features, fields.frequency_encode(positions, 6).to(dt)], -1)
    sig = torch.sigmoid
    albedo6 = fields.vanilla_mlp(inp, P.nets["albedo"], "relu")
    diff_rgb, albedo = sig(albedo6[:, :3]), sig(albedo6[:, 3:])
    roughness = sig(fields.vanilla_mlp(inp, P.nets["roughness"], "relu"))
    met2 = fields.vanilla_mlp(inp, P.nets["metallic"], "relu")
    blend, metallic = sig(met2[:, :1]), sig(met2[:, 1:])
    wo_enc = fields.sh_encode((wo + 1.0) / 2.0, 5)
    spec_rgb = sig(fields.vanilla_mlp(torch.cat([features, wo_enc], -1), P.nets["env"], "relu"))
    spec_rgb = blend * spec_rgb
    diff_rgb = (1 - blend) * diff_rgb
    if stage == 0:
        return torch.cat([diff_rgb, spec_rgb, blend], -1)
    diff_pbr = (1 - metallic) * albedo * eval_mip(P, normals)
    spec_albedo = 0.04 * (1 - metallic) + metallic * albedo
    spec_light = eval_mip(P, wo, specular=True, roughness=roughness)
    fg = _fg(P, NoV, roughness)
    spec_ref = spec_albedo * fg[:, 0:1] + fg[:, 1:2]
    return torch.cat([diff_rgb, spec_rgb, blend, diff_pbr, spec_ref * spec_light, spec_ref, spec_light, albedo,
                      metallic, roughness], -1)


def secondary_shading(P, features, rays_d, normal):
    emb = fields.sh_encode((rays_d + 1.0) / 2.0, 5)
    return torch.sigmoid(fields.vanilla_mlp(torch.cat([features, emb, normal], -1), P.nets["secondary"], "relu"))


def secondary_shading_pbr(P, features, dirs, normals, positions):
    dt = features.dtype
    wi = -dirs
    NoV = torch.sum(normals * wi, -1, keepdim=True)
    inp = torch.cat([features, fields.frequency_encode(positions, 6).to(dt)], -1)
    sig = torch.sigmoid
    albedo = sig(fields.vanilla_mlp(inp, P.nets["albedo"], "relu")[:, 3:])
    roughness = sig(fields.vanilla_mlp(inp, P.nets["roughness"], "relu"))
    metallic = sig(fields.vanilla_mlp(inp, P.nets["metallic"], "relu")[:, 1:])
    diff_pbr = (1 - metallic) * albedo * eval_mip(P, normals)
    spec_albedo = 0.04 * (1 - metallic) + metallic * albedo
    spec_light = eval_mip(P, dirs, specular=True, roughness=roughness)
    fg = _fg(P, NoV, roughness)
    return diff_pbr + (spec_albedo * fg[:, 0:1] + fg[:, 1:2]) * spec_light


def _alpha_fn(P, rays_o, rays_d, ratio, dtype=torch.float32, keep=None):
    """alpha_fn of models/split_mixed_occ.py:182-195,228-240 (always called under no_grad by `sampling`).  With
    `keep` (a dict) the candidates' alphas and packed positions are recorded for the sample-set diagnostics."""
    def fn(ts, te, ri):
        ts, te, ri = torch.as_tensor(ts), torch.as_tensor(te), torch.as_tensor(ri)
        if len(ri) == 0:
            return np.zeros(0, np.float32)
        with torch.no_grad():
            t_o, t_d = rays_o[ri], rays_d[ri]
            pos = t_o + t_d * (ts + te)[:, None] / 2.0
            sdf, grad, _ = geometry(P, pos, dtype=dtype)
            normal = F.normalize(grad, p=2, dim=-1, eps=1e-6)
            a = fields.get_alpha(sdf, normal, t_d.to(dtype), (te - ts)[:, None].to(dtype), P.inv_s.view(1, 1), ratio)
        if keep is not None:
            keep.update(ts=ts, te=te, ri=ri, alphas=a)
        return a.float().numpy()
    return fn


def _sample(P, rays_o, rays_d, grid, step, ratio, near=0.0, far=1e10, dtype=torch.float32, keep=None, jitter=None):
    roi = np.array([-P.radius] * 3 + [P.radius] * 3, np.float32)
    ri, ts, te = march.ray_marching(rays_o.numpy(), rays_d.numpy(), scene_aabb=roi, grid_roi=roi, grid_binary=grid,
                                    alpha_fn=_alpha_fn(P, rays_o, rays_d, ratio, dtype, keep), render_step_size=step,
                                    near_plane=near, far_plane=far, jitter=jitter)
    return torch.from_numpy(ri), torch.from_numpy(ts), torch.from_numpy(te)


def rgb_to_srgb(f):
    return torch.where(f <= 0.0031308, f * 12.92, torch.pow(torch.clamp(f, 0.0031308), 1.0 / 2.4) * 1.055 - 0.055)


def _render(P, rays, grid, render_step_size, stage, relighting, cos_anneal_ratio, background, relighting_threshold,
            secondary, training, dtype, laplace_dirs, samples, keep, jitter):
    """SplitMixedOCCModel.forward_ (models/split_mixed_occ.py:224-443)."""
    n_rays = rays.shape[0]
    rays_o, rays_d = rays[:, :3].contiguous(), rays[:, 3:6].contiguous()
    if samples is None:
        ri, ts, te = _sample(P, rays_o, rays_d, grid, render_step_size, cos_anneal_ratio, dtype=dtype, keep=keep,
                             jitter=jitter)
    else:                                # a sample set fixed by the caller (gradient tests: same set on both sides)
        ri, ts, te = (torch.as_tensor(t) for t in samples)
    t_o, t_d = rays_o[ri], rays_d[ri]
    pos = t_o + t_d * (ts + te)[:, None] / 2.0          # fp32, as the product computes them
    cdim = 7 if stage == 0 else 24
    laplace = None
    if len(ri):
        if training:
            sdf, grad, feat, laplace = geometry(P, pos, dtype=dtype, laplace_dirs=laplace_dirs)
        else:
            sdf, grad, feat = geometry(P, pos, dtype=dtype)
        normal = F.normalize(grad, p=2, dim=-1, eps=1e-6)
        alpha = fields.get_alpha(sdf, normal, t_d.to(dtype), (te - ts)[:, None].to(dtype), P.inv_s.view(1, 1),
                                 cos_anneal_ratio)
        colors = texture_forward(P, feat, t_d.to(dtype), normal, pos, stage)
    else:
        z = lambda *sh: torch.zeros(*sh, dtype=dtype)
        sdf, grad, normal, alpha, colors = z(0), z(0, 3), z(0, 3), z(0), z(0, cdim)
    w, _ = fields.render_weight_from_alpha(alpha, ri, n_rays)
    rgb_map = fields.accumulate_along_rays(w, colors, ri, n_rays)
    normal_map = fields.accumulate_along_rays(w, normal, ri, n_rays)
    acc = fields.accumulate_along_rays(w, None, ri, n_rays)
    depth = fields.accumulate_along_rays(w, ((ts + te)[:, None] / 2.0).to(dtype), ri, n_rays)
    valid = torch.nonzero(acc > 0.5)[:, 0]
    # the reference writes through slices of rgb_map in place (:296-318); out-of-place here, same values / gradients
    diff, spec, blend = rgb_map[:, :3], rgb_map[:, 3:6], rgb_map[:, 6:7]
    if stage:
        diff_pbr, spec_pbr = rgb_map[:, 7:10], rgb_map[:, 10:13]
        spec_ref, spec_light = rgb_map[:, 13:16], rgb_map[:, 16:19]
        albedo_map, metallic_map, rough_map = rgb_map[:, 19:22], rgb_map[:, 22:23], rgb_map[:, 23:]
    if len(valid):
        so = rays_o[valid].to(dtype) + depth[valid] * rays_d[valid].to(dtype)
        wo = -rays_d[valid].to(dtype)
        nm = normal_map[valid]
        sd = 2 * torch.sum(wo * nm, -1, keepdim=True) * nm - wo
        near, far, ns = secondary
        sstep = (far - near) / (ns - 1)
        with torch.no_grad():            # compute_indirect_radiance (:179-222)
            so_d, sd_d = so.detach().contiguous(), sd.detach().contiguous()
            ri2, ts2, te2 = _sample(P, so_d, sd_d, grid, sstep, cos_anneal_ratio, near, far, dtype=dtype)
            keep2 = {}
            _alpha_fn(P, so_d, sd_d, cos_anneal_ratio, dtype, keep2)(ts2, te2, ri2)
            a2 = keep2["alphas"] if len(ri2) else torch.zeros(0, dtype=dtype)
            w2, _ = fields.render_weight_from_alpha(a2, ri2, len(valid))
            acc2 = fields.accumulate_along_rays(w2, None, ri2, len(valid))
            depth2 = fields.accumulate_along_rays(w2, ((ts2 + te2)[:, None] / 2.0).to(dtype), ri2, len(valid))
            tr = (1.0 - acc2).clamp(0, 1)
        _, _, sfeat = geometry(P, so if training else so.detach(), with_grad=False, dtype=dtype)
        srgb = secondary_shading(P, sfeat, sd, nm)
        scatter = lambda base, rows: base.index_put((valid,), rows)
        spec = scatter(spec, tr * spec[valid] + (1 - tr) * srgb)
        if stage:
            if not relighting:
                spec_pbr = scatter(spec_pbr, tr * spec_pbr[valid] + (1 - tr) * srgb)
            else:
                mask = (rough_map[valid] <= relighting_threshold)[:, 0]
                to = so[mask] + depth2[mask] * sd[mask]
                if mask.any():
                    _, tgrad, tfeat = geometry(P, to if training else to.detach(), dtype=dtype)
                    tn = F.normalize(tgrad, p=2, dim=-1, eps=1e-6)
                    trgb = secondary_shading_pbr(P, tfeat, sd[mask], tn, to)
                    slv = spec_light[valid]
                    slv = slv.index_put((torch.nonzero(mask)[:, 0],), tr[mask] * slv[mask] + (1 - tr[mask]) * trgb)
                    spec_light = scatter(spec_light, slv)
                spec_pbr = spec_ref * spec_light
    out = {"comp_rgb": diff + spec, "comp_diffuse_rgb": diff, "comp_spec_rgb": spec, "comp_blend": blend,
           "comp_normal": normal_map, "opacity": acc, "depth": depth, "num_samples": len(ri),
           "ray_indices": ri, "t_starts": ts, "t_ends": te, "valid_indices": valid, "rays_valid": acc > 0}
    if stage:
        out.update({"comp_rgb_phys": diff_pbr + spec_pbr, "comp_albedo": albedo_map, "comp_metallic": metallic_map,
                    "comp_roughness": rough_map, "comp_spec_rgb_phys": spec_pbr, "comp_diffuse_rgb_phys": diff_pbr})
    if training:
        out.update({"sdf_samples": sdf, "sdf_grad_samples": grad, "weights": w, "sdf_laplace_samples": laplace})
        if len(ri):
            orient = torch.sum(rays_d[ri].to(dtype) * normal, -1, keepdim=True).clamp(min=0)
            out["normals_orientation_loss_map"] = fields.accumulate_along_rays(w, orient, ri, n_rays)
        else:
            out["normals_orientation_loss_map"] = torch.zeros(n_rays, 1, dtype=dtype)
    if background is not None:
        bgc = background[None, :].to(dtype)
        comp = lambda x: rgb_to_srgb(x + bgc * (1.0 - acc)).clamp(0, 1)
        out["comp_rgb_full"] = comp(out["comp_rgb"])
        out["rays_valid_full"] = out["rays_valid"]
        if stage:
            out["comp_rgb_phys_full"] = comp(out["comp_rgb_phys"])
            out["comp_spec_rgb_full"] = comp(out["comp_spec_rgb"])
            out["comp_spec_rgb_phys_full"] = comp(out["comp_spec_rgb_phys"])
    return out


@torch.no_grad()
def forward(P, rays, grid, render_step_size, stage=1, relighting=False, cos_anneal_ratio=1.0, background=None,
            relighting_threshold=0.3, secondary=(0.05, 1.5, 96), dtype=torch.float32, samples=None, keep=None):
    """eval-mode render (no grad; `relighting=True` adds the third bounce of :323-332)."""
    return _render(P, rays, grid, render_step_size, stage, relighting, cos_anneal_ratio, background,
                   relighting_threshold, secondary, False, dtype, None, samples, keep, None)


def forward_train(P, rays, grid, render_step_size, laplace_dirs, stage=1, cos_anneal_ratio=1.0, background=None,
                  secondary=(0.05, 1.5, 96), dtype=torch.float32, samples=None, jitter=None):
    """training-mode render: differentiable w.r.t. every tensor of `P` (incl. P.specular / P.diffuse or, when they
    were built from it with build_mips under grad, P.base).  `laplace_dirs` [S,3]: the uniform draws of the curvature
    probe (models/geometry.py:253), S = the number of samples the march keeps."""
    return _render(P, rays, grid, render_step_size, stage, False, cos_anneal_ratio, background, 0.3, secondary, True,
                   dtype, laplace_dirs, samples, None, jitter)


SPLIT_LAMBDAS = dict(lambda_rgb_mse=10.0, lambda_rgb_l1=0.0, lambda_rgb_phys_mse=10.0, lambda_rgb_phys_l1=0.0,
                     lambda_mask=0.1, lambda_eikonal=0.05, lambda_sparsity=0.01, lambda_curvature=1.0,
                     lambda_opaque=0.0, lambda_normal_orientation=0.05, lambda_emitter_distillation=0.0,
                     sparsity_scale=1.0)      # configs/split-mixed-occ-tensoir.yaml:139-152


def loss(out, rgb, fg_mask, stage=1, has_mask=True, **overrides):
    """systems/split_occ.py:163-225 (the distortion terms have lambda 0 in the config and need a library that is
    not part of the path)."""
    lam = dict(SPLIT_LAMBDAS, **overrides)
    valid = out["rays_valid_full"][:, 0]
    dt = out["comp_rgb_full"].dtype
    rgb = rgb.to(dt)
    parts = {"rgb_mse": F.mse_loss(out["comp_rgb_full"][valid], rgb[valid])}
    total = parts["rgb_mse"] * lam["lambda_rgb_mse"]
    total = total + F.l1_loss(out["comp_rgb_full"][valid], rgb[valid]) * lam["lambda_rgb_l1"]
    if stage != 0:
        parts["rgb_phys_mse"] = F.mse_loss(out["comp_rgb_phys_full"][valid], rgb[valid])
        total = total + parts["rgb_phys_mse"] * lam["lambda_rgb_phys_mse"]
        total = total + F.l1_loss(out["comp_rgb_phys_full"][valid], rgb[valid]) * lam["lambda_rgb_phys_l1"]
    parts["eikonal"] = ((torch.linalg.norm(out["sdf_grad_samples"], ord=2, dim=-1) - 1.0) ** 2).mean()
    total = total + parts["eikonal"] * lam["lambda_eikonal"]
    opacity = torch.clamp(out["opacity"].squeeze(-1), 1e-3, 1 - 1e-3)
    bce = lambda x, t: -(t * torch.log(x) + (1 - t) * torch.log(1 - x)).mean()
    parts["mask"] = bce(opacity, fg_mask.to(dt))
    total = total + parts["mask"] * (lam["lambda_mask"] if has_mask else 0.0)
    total = total + bce(opacity, opacity) * lam["lambda_opaque"]
    parts["sparsity"] = torch.exp(-lam["sparsity_scale"] * out["sdf_samples"].abs()).mean()
    total = total + parts["sparsity"] * lam["lambda_sparsity"]
    if lam["lambda_curvature"] > 0:
        parts["curvature"] = out["sdf_laplace_samples"].abs().mean()
        total = total + parts["curvature"] * lam["lambda_curvature"]
    if lam["lambda_emitter_distillation"] > 0 and stage != 0:
        total = total + F.mse_loss(out["comp_spec_rgb_full"][valid], out["comp_spec_rgb_phys_full"][valid]) \
            * lam["lambda_emitter_distillation"]
    parts["normal_orientation"] = out["normals_orientation_loss_map"].mean()       # models/geometry.py:321-326
    total = total + parts["normal_orientation"] * lam["lambda_normal_orientation"]
    return total, parts
