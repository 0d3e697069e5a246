"""Host-side mirror of the two models/volrend.py renderers the split-sum model calls:
    secondary_rendering .......... models/volrend.py:18-127
    rendering_with_normals_sdf ... models/volrend.py:739-895
Scan + accumulation run on librsdf_b200.so (deterministic segmented reductions)."""
from typing import Callable, Dict, Optional, Tuple

import torch
from torch import Tensor

from .nerfacc import accumulate_along_rays, pack_info, render_weight_from_alpha
from .neus import chunk_batch


def secondary_rendering(t_starts: Tensor, t_ends: Tensor, ray_indices: Optional[Tensor] = None,
                        n_rays: Optional[int] = None, alpha_fn: Optional[Callable] = None,
                        chunk_size: Optional[int] = None):
    if ray_indices is not None:
        assert t_starts.shape == t_ends.shape == ray_indices.shape, \
            "Since nerfacc 0.5.0, t_starts, t_ends and ray_indices must have the same shape (N,). "
    if t_starts.shape[0] != 0:
        if chunk_size is None:
            alphas = alpha_fn(t_starts, t_ends, ray_indices)
        else:
            alphas = chunk_batch(alpha_fn, chunk_size, False, t_starts, t_ends, ray_indices)
    else:
        alphas = torch.empty((0,), device=t_starts.device)
    assert alphas.shape == t_starts.shape, "alphas must have shape of (N,)! Got {}".format(alphas.shape)
    packed = pack_info(ray_indices, n_rays)
    weights, trans = render_weight_from_alpha(alphas, ray_indices=ray_indices, n_rays=n_rays, packed_info=packed)
    extras = {"weights": weights, "trans": trans, "alphas": alphas}
    kw = dict(ray_indices=ray_indices, n_rays=n_rays, packed_info=packed)
    opacities = accumulate_along_rays(weights, values=None, **kw)
    depths = accumulate_along_rays(weights, values=(t_starts + t_ends)[..., None] / 2.0, **kw)
    return opacities, depths, extras


def rendering_with_normals_sdf(t_starts: Tensor, t_ends: Tensor, ray_indices: Optional[Tensor] = None,
                               n_rays: Optional[int] = None, rgb_sigma_fn: Optional[Callable] = None,
                               rgb_alpha_fn: Optional[Callable] = None, render_bkgd: Optional[Tensor] = None,
                               has_laplace: bool = False, color_dim=3, normal_dim=3
                               ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Dict]:
    if ray_indices is not None:
        assert t_starts.shape == t_ends.shape == ray_indices.shape, \
            "Since nerfacc 0.5.0, t_starts, t_ends and ray_indices must have the same shape (N,). "
    if rgb_sigma_fn is None and rgb_alpha_fn is None:
        raise ValueError("At least one of `rgb_sigma_fn` and `rgb_alpha_fn` should be specified.")
    if rgb_sigma_fn is not None:
        raise NotImplementedError("rgb_sigma_fn is not implemented yet.")
    dev = t_starts.device
    sdf_laplace = None
    if t_starts.shape[0] != 0:
        if has_laplace:
            rgbs, normals, alphas, sdf, sdf_grad, sdf_laplace = rgb_alpha_fn(t_starts, t_ends, ray_indices)
        else:
            rgbs, normals, alphas, sdf, sdf_grad = rgb_alpha_fn(t_starts, t_ends, ray_indices)
    else:
        rgbs = torch.empty((0, color_dim), device=dev)
        normals = torch.empty((0, normal_dim), device=dev)
        alphas = torch.empty((0,), device=dev)
        sdf = torch.empty((0,), device=dev)
        sdf_grad = torch.empty((0, 3), device=dev)
        sdf_laplace = torch.empty((0, 3), device=dev)
    assert rgbs.shape[-1] == color_dim, "rgbs must have 3 channels, got {}".format(rgbs.shape)
    assert normals.shape[-1] == normal_dim, "normals must have 3 channels, got {}".format(normals.shape)
    assert alphas.shape == t_starts.shape, "alphas must have shape of (N,)! Got {}".format(alphas.shape)
    assert sdf.shape == t_starts.shape, "sdf must have shape of (N,)! Got {}".format(sdf.shape)
    assert sdf_grad.shape[-1] == 3, "sdf_grad must have 3 channels, got {}".format(sdf_grad.shape)
    packed = pack_info(ray_indices, n_rays)
    weights, trans = render_weight_from_alpha(alphas, ray_indices=ray_indices, n_rays=n_rays, packed_info=packed)
    extras = {"weights": weights, "trans": trans, "rgbs": rgbs, "alphas": alphas, "normals": normals,
              "sdf": sdf, "sdf_grad": sdf_grad}
    if has_laplace:
        extras["sdf_laplace"] = sdf_laplace
    kw = dict(ray_indices=ray_indices, n_rays=n_rays, packed_info=packed)
    colors = accumulate_along_rays(weights, values=rgbs, **kw)
    normals = accumulate_along_rays(weights, values=normals, **kw)
    opacities = accumulate_along_rays(weights, values=None, **kw)
    depths = accumulate_along_rays(weights, values=(t_starts + t_ends)[..., None] / 2.0, **kw)
    if render_bkgd is not None:
        colors = colors + render_bkgd * (1.0 - opacities)
        normals = normals + render_bkgd * (1 - opacities) * torch.tensor([0.0, 0.0, 1.0], device=normals.device)
    return colors, normals, opacities, depths, extras
