"""CPU restatement of the `neus` render + training loss.  TEST INFRASTRUCTURE ONLY.

Follows models/neus.py:227-317 (forward_), :90-122 (update_step / occ_eval_fn),
lib/nerfacc/grid.py:196-239 (occupancy EMA update), systems/neus.py:98-135 (losses),
systems/criterions.py (binary_cross_entropy).  OccGridEstimator.sampling (nerfacc 0.5.3,
third-party, absent) is DEFINED as the in-tree 0.3.5 ray_marching with scene_aabb = roi
(SURVEY.md Appendix A.5); PARITY UNPINNED for that one definitional choice.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import fields, march


class NeusParams:
    """Plain container: everything the neus model learns."""

    def __init__(self, table, geo_mlp, tex_mlp, variance, meta, radius=1.5, sh_degree=4):
        self.table, self.geo_mlp, self.tex_mlp, self.variance = table, geo_mlp, tex_mlp, variance
        self.meta, self.radius, self.sh_degree = meta, radius, sh_degree

    def tensors(self):
        out = [self.table, self.variance]
        for net in (self.geo_mlp, self.tex_mlp):
            for layer in net:
                out += list(layer.values())
        return out

    @property
    def inv_s(self):
        return torch.exp(self.variance * 10.0)

    def to(self, dtype):
        """Copy with every learnable tensor cast to `dtype` (leaf tensors)."""
        cast = lambda t: t.detach().to(dtype).clone()
        return NeusParams(cast(self.table), [{k: cast(v) for k, v in l.items()} for l in self.geo_mlp],
                          [{k: cast(v) for k, v in l.items()} for l in self.tex_mlp], cast(self.variance),
                          self.meta, self.radius, self.sh_degree)


def make_params(seed=42, table_scale=1e-4, base_resolution=16, n_neurons=128, variance=0.3):
    gen = torch.Generator().manual_seed(seed)
    meta = fields.HashGridMeta(base_resolution=base_resolution)
    table = (torch.rand(meta.n_params, generator=gen) * 2 - 1) * table_scale
    geo = fields.init_mlp(3 + meta.n_output_dims, 48, n_neurons, 2, True, True, gen)
    tex = fields.init_mlp(48 + 16 + 3, 3, n_neurons, 4, False, False, gen)
    return NeusParams(table, geo, tex, torch.tensor(variance), meta)


def grid_update(occs, step, occ_eval_fn, roi, resolution=128, occ_thre=0.01, ema_decay=0.95,
                warmup_steps=256, jitter=None, gen=None):
    """lib/nerfacc/grid.py:196-239.  occs: float32 [res^3] (updated copy returned)."""
    n_cells = resolution ** 3
    if step < warmup_steps:
        indices = torch.arange(n_cells)
    else:
        n = n_cells // 4
        uniform = torch.randint(n_cells, (n,), generator=gen)
        occupied = torch.nonzero(occs > torch.clamp(occs.mean(), max=occ_thre))[:, 0]
        if n < len(occupied):
            occupied = occupied[torch.randint(len(occupied), (n,), generator=gen)]
        indices = torch.cat([uniform, occupied])
    r = resolution
    coords = torch.stack([indices // (r * r), (indices // r) % r, indices % r], -1).float()
    if jitter is None:
        jitter = torch.rand(coords.shape, generator=gen)
    x = (coords + jitter) / r
    roi = torch.as_tensor(roi, dtype=torch.float32)
    x = x * (roi[3:] - roi[:3]) + roi[:3]
    occ = occ_eval_fn(x).squeeze(-1)
    occs = occs.clone()
    occs[indices] = torch.maximum(occs[indices] * ema_decay, occ)
    binary = (occs > torch.clamp(occs.mean(), max=occ_thre)).view(r, r, r)
    return occs, binary


def occ_eval_fn(P, render_step_size):
    def fn(x):
        with torch.no_grad():
            out = []
            for c in torch.split(x, 1 << 18):
                sdf, _, _ = fields.sdf_field(c, P.table, P.meta, P.geo_mlp, P.radius, with_grad=False)
                out.append(fields.occ_alpha(sdf, P.inv_s.detach().view(1, 1), render_step_size))
            return torch.cat(out)
    return fn


def forward(P, rays, grid_binary, render_step_size, cos_anneal_ratio=1.0, background=None,
            jitter=None, training=False, create_graph=False, dtype=torch.float32):
    """NeuSModel.forward_ on CPU.  rays [R,6] float32 torch tensor."""
    n_rays = rays.shape[0]
    rays_o, rays_d = rays[:, :3], rays[:, 3:6]
    roi = np.array([-P.radius] * 3 + [P.radius] * 3, np.float32)
    ri, ts, te = march.ray_marching(
        rays_o.numpy(), rays_d.numpy(), scene_aabb=roi, grid_roi=roi,
        grid_binary=grid_binary, render_step_size=render_step_size, jitter=jitter,
        near_plane=0.0, far_plane=1e10)
    ray_indices = torch.from_numpy(ri)
    t_starts, t_ends = torch.from_numpy(ts), torch.from_numpy(te)
    t_o, t_d = rays_o[ray_indices], rays_d[ray_indices]
    midpoints = (t_starts + t_ends)[:, None] / 2.0
    positions = t_o + t_d * midpoints          # fp32, as the product computes them
    dists = (t_ends - t_starts).to(dtype)
    t_d, midpoints = t_d.to(dtype), midpoints.to(dtype)
    if len(ri) == 0:
        z = torch.zeros
        out = {"comp_rgb": z(n_rays, 3), "comp_normal": z(n_rays, 3), "opacity": z(n_rays, 1),
               "depth": z(n_rays, 1), "num_samples": 0}
    else:
        sdf, sdf_grad, feature = fields.sdf_field(positions, P.table, P.meta, P.geo_mlp, P.radius,
                                                  with_grad=True, create_graph=create_graph, dtype=dtype)
        sdf_grad = sdf_grad.to(dtype)
        normal = F.normalize(sdf_grad, p=2, dim=-1)
        alpha = fields.get_alpha(sdf, normal, t_d, dists, P.inv_s.view(1, 1), cos_anneal_ratio)
        rgb = fields.radiance(feature, t_d, normal, P.tex_mlp, P.sh_degree)
        weights, _ = fields.render_weight_from_alpha(alpha, ray_indices, n_rays)
        opacity = fields.accumulate_along_rays(weights, None, ray_indices, n_rays)
        depth = fields.accumulate_along_rays(weights, midpoints, ray_indices, n_rays)
        comp_rgb = fields.accumulate_along_rays(weights, rgb, ray_indices, n_rays)
        comp_normal = fields.accumulate_along_rays(weights, normal, ray_indices, n_rays)
        comp_normal = F.normalize(comp_normal, p=2, dim=-1)
        out = {"comp_rgb": comp_rgb, "comp_normal": comp_normal, "opacity": opacity, "depth": depth,
               "num_samples": len(ri)}
        if training:
            out.update({"sdf_samples": sdf, "sdf_grad_samples": sdf_grad, "weights": weights,
                        "alpha": alpha, "rgb": rgb, "normal": normal})
    out["rays_valid"] = out["opacity"] > 0
    out["ray_indices"], out["t_starts"], out["t_ends"] = ray_indices, t_starts, t_ends
    if background is not None:
        out["comp_rgb_full"] = out["comp_rgb"] + background[None, :].to(dtype) * (1.0 - out["opacity"])
    return out


def binary_cross_entropy(inp, target):
    """systems/criterions.py: -(t log x + (1-t) log(1-x)).mean()"""
    return -(target * torch.log(inp) + (1 - target) * torch.log(1 - inp)).mean()


def loss(out, target_rgb, fg_mask, lambda_rgb_mse=10.0, lambda_mask=0.1, lambda_eikonal=0.1,
         lambda_sparsity=0.01, sparsity_scale=1.0):
    """systems/neus.py:98-121 with configs/neus-blender.yaml:83-91 weights."""
    valid = out["rays_valid"][:, 0]
    l_rgb = F.mse_loss(out["comp_rgb_full"][valid], target_rgb[valid])
    l_eik = ((torch.linalg.norm(out["sdf_grad_samples"], ord=2, dim=-1) - 1.0) ** 2).mean()
    opacity = torch.clamp(out["opacity"].squeeze(-1), 1e-3, 1 - 1e-3)
    l_mask = binary_cross_entropy(opacity, fg_mask.float())
    l_sparse = torch.exp(-sparsity_scale * out["sdf_samples"].abs()).mean()
    total = l_rgb * lambda_rgb_mse + l_eik * lambda_eikonal + l_mask * lambda_mask + l_sparse * lambda_sparsity
    return total, {"rgb_mse": l_rgb, "eikonal": l_eik, "mask": l_mask, "sparsity": l_sparse}
