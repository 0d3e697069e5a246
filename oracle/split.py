"""CPU restatement of the split-sum (`split-mixed-occ`) render.  TEST INFRASTRUCTURE ONLY.

Follows models/split_mixed_occ.py:179-443 (forward_, compute_indirect_radiance),
models/texture.py:292-427 (VolumeMixedMipSplitOcc.forward / secondary_shading /
secondary_shading_pbr), lib/pbr/light.py:169-206 (build_mips / get_mip / eval_mip),
lib/pbr/utils/light_utils.py:71-74 (avg-pool mip), models/volrend.py:18-127,739-895,
lib/pbr/utils/nvdiffrecmc_util.py:95-103 (rgb_to_srgb), models/geometry.py:229-244,304-318
(finite-difference normals with the progressive eps).  Third-party pieces (tcnn, nerfacc 0.5.3,
nvdiffrast) are restated as in oracle/fields.py / oracle/textures.py (PARITY UNPINNED there).
Forward only (the relighting path).
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


def build_mips(P, cutoff=0.99):
    spec = [P.base]
    while spec[-1].shape[1] > LIGHT_MIN_RES:
        x = spec[-1].permute(0, 3, 1, 2)
        spec.append(F.avg_pool2d(x, (2, 2)).permute(0, 2, 3, 1).contiguous())
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


def geometry(P, points, with_grad=True):
    """VolumeSDF.forward with grad_type='finite_difference' (eval mode: no laplace)."""
    if with_grad:
        sdf, grad, feat = fields.sdf_field_fd(points, P.table, P.meta, P.geo_mlp, P.radius, P.fd_eps,
                                              level_mask=P.level_mask)
        return sdf, grad, feat
    x01 = fields.scale_to_unit(points, P.radius)
    enc = fields.hash_encode(x01, P.table, P.meta)
    if P.level_mask is not None:
        enc = enc * P.level_mask
    out = fields.vanilla_mlp(torch.cat([x01 * 2.0 - 1.0, enc], -1), P.geo_mlp, "softplus100")
    return out[:, 0], None, out


def _fg(P, NoV, roughness):
    uv = torch.cat([NoV.clamp(0.0, 1.0), roughness.clamp(0.0, 1.0)], -1)
    return tx.tex2d(P.fg_lut[0], uv)


def texture_forward(P, features, dirs, normals, positions, stage):
    wi = -dirs
    wo = torch.sum(wi * normals, -1, keepdim=True) * normals * 2 - wi
    NoV = torch.sum(normals * wi, -1, keepdim=True)
    inp = torch.cat([features, fields.frequency_encode(positions, 6)], -1)
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
    wi = -dirs
    NoV = torch.sum(normals * wi, -1, keepdim=True)
    inp = torch.cat([features, fields.frequency_encode(positions, 6)], -1)
    sig = torch.sigmoid
    albedo = sig(fields.vanilla_mlp(inp, P.nets["albedo"], "relu")[:, 3:])
    roughness = sig(fields.vanilla_mlp(inp, P.nets["roughness"], "relu"))
    metallic = sig(fields.vanilla_mlp(inp, P.nets["metallic"], "relu")[:, 1:])
    diff_pbr = (1 - metallic) * albedo * eval_mip(P, normals)
    spec_albedo = 0.04 * (1 - metallic) + metallic * albedo
    spec_light = eval_mip(P, dirs, specular=True, roughness=roughness)
    fg = _fg(P, NoV, roughness)
    return diff_pbr + (spec_albedo * fg[:, 0:1] + fg[:, 1:2]) * spec_light


def _alpha_fn(P, rays_o, rays_d, ratio):
    def fn(ts, te, ri):
        ts, te, ri = torch.as_tensor(ts), torch.as_tensor(te), torch.as_tensor(ri)
        if len(ri) == 0:
            return np.zeros(0, np.float32)
        t_o, t_d = rays_o[ri], rays_d[ri]
        pos = t_o + t_d * (ts + te)[:, None] / 2.0
        sdf, grad, _ = geometry(P, pos)
        normal = F.normalize(grad, p=2, dim=-1, eps=1e-6)
        return fields.get_alpha(sdf, normal, t_d, (te - ts)[:, None], P.inv_s.view(1, 1), ratio).numpy()
    return fn


def _sample(P, rays_o, rays_d, grid, step, ratio, near=0.0, far=1e10):
    roi = np.array([-P.radius] * 3 + [P.radius] * 3, np.float32)
    ri, ts, te = march.ray_marching(rays_o.numpy(), rays_d.numpy(), scene_aabb=roi, grid_roi=roi, grid_binary=grid,
                                    alpha_fn=_alpha_fn(P, rays_o, rays_d, ratio), render_step_size=step,
                                    near_plane=near, far_plane=far)
    return torch.from_numpy(ri), torch.from_numpy(ts), torch.from_numpy(te)


def rgb_to_srgb(f):
    return torch.where(f <= 0.0031308, f * 12.92, torch.pow(torch.clamp(f, 0.0031308), 1.0 / 2.4) * 1.055 - 0.055)


@torch.no_grad()
def forward(P, rays, grid, render_step_size, stage=1, relighting=False, cos_anneal_ratio=1.0, background=None,
            relighting_threshold=0.3, secondary=(0.05, 1.5, 96)):
    n_rays = rays.shape[0]
    rays_o, rays_d = rays[:, :3].contiguous(), rays[:, 3:6].contiguous()
    ri, ts, te = _sample(P, rays_o, rays_d, grid, render_step_size, cos_anneal_ratio)
    t_o, t_d = rays_o[ri], rays_d[ri]
    pos = t_o + t_d * (ts + te)[:, None] / 2.0
    cdim = 7 if stage == 0 else 24
    if len(ri):
        sdf, grad, feat = geometry(P, pos)
        normal = F.normalize(grad, p=2, dim=-1, eps=1e-6)
        alpha = fields.get_alpha(sdf, normal, t_d, (te - ts)[:, None], P.inv_s.view(1, 1), cos_anneal_ratio)
        colors = texture_forward(P, feat, t_d, normal, pos, stage)
    else:
        normal, alpha, colors = torch.zeros(0, 3), torch.zeros(0), torch.zeros(0, cdim)
    w, _ = fields.render_weight_from_alpha(alpha, ri, n_rays)
    rgb_map = fields.accumulate_along_rays(w, colors, ri, n_rays)
    normal_map = fields.accumulate_along_rays(w, normal, ri, n_rays)
    acc = fields.accumulate_along_rays(w, None, ri, n_rays)
    depth = fields.accumulate_along_rays(w, (ts + te)[:, None] / 2.0, ri, n_rays)
    valid = torch.nonzero(acc > 0.5)[:, 0]
    diff, spec, blend = rgb_map[:, :3].clone(), rgb_map[:, 3:6].clone(), rgb_map[:, 6:7]
    if stage:
        diff_pbr, spec_pbr = rgb_map[:, 7:10].clone(), rgb_map[:, 10:13].clone()
        spec_ref, spec_light = rgb_map[:, 13:16], rgb_map[:, 16:19].clone()
        albedo_map, metallic_map, rough_map = rgb_map[:, 19:22], rgb_map[:, 22:23], rgb_map[:, 23:]
    if len(valid):
        so = rays_o[valid] + depth[valid] * rays_d[valid]
        wo = -rays_d[valid]
        nm = normal_map[valid]
        sd = 2 * torch.sum(wo * nm, -1, keepdim=True) * nm - wo
        near, far, ns = secondary
        sstep = (far - near) / (ns - 1)
        ri2, ts2, te2 = _sample(P, so, sd, grid, sstep, cos_anneal_ratio, near, far)
        a2 = torch.from_numpy(_alpha_fn(P, so, sd, cos_anneal_ratio)(ts2, te2, ri2))
        w2, _ = fields.render_weight_from_alpha(a2, ri2, len(valid))
        acc2 = fields.accumulate_along_rays(w2, None, ri2, len(valid))
        depth2 = fields.accumulate_along_rays(w2, (ts2 + te2)[:, None] / 2.0, ri2, len(valid))
        tr = (1.0 - acc2).clamp(0, 1)
        _, _, sfeat = geometry(P, so, with_grad=False)
        srgb = secondary_shading(P, sfeat, sd, nm)
        spec[valid] = tr * spec[valid] + (1 - tr) * srgb
        if stage:
            if not relighting:
                spec_pbr[valid] = tr * spec_pbr[valid] + (1 - tr) * srgb
            else:
                mask = (rough_map[valid] <= relighting_threshold)[:, 0]
                to = so[mask] + depth2[mask] * sd[mask]
                if mask.any():
                    _, tgrad, tfeat = geometry(P, to)
                    tn = F.normalize(tgrad, p=2, dim=-1, eps=1e-6)
                    trgb = secondary_shading_pbr(P, tfeat, sd[mask], tn, to)
                    slv = spec_light[valid]
                    slv[mask] = tr[mask] * slv[mask] + (1 - tr[mask]) * trgb
                    spec_light[valid] = slv
                spec_pbr = spec_ref * spec_light
    out = {"comp_rgb": diff + spec, "comp_diffuse_rgb": diff, "comp_spec_rgb": spec, "comp_blend": blend,
           "comp_normal": normal_map, "opacity": acc, "depth": depth, "num_samples": len(ri),
           "ray_indices": ri, "valid_indices": valid}
    if stage:
        out.update({"comp_rgb_phys": diff_pbr + spec_pbr, "comp_albedo": albedo_map, "comp_metallic": metallic_map,
                    "comp_roughness": rough_map, "comp_spec_rgb_phys": spec_pbr})
    if background is not None:
        comp = lambda x: rgb_to_srgb(x + background[None, :] * (1.0 - acc)).clamp(0, 1)
        out["comp_rgb_full"] = comp(out["comp_rgb"])
        if stage:
            out["comp_rgb_phys_full"] = comp(out["comp_rgb_phys"])
    return out
