"""Autograd wrappers of the two fused split-sum stages (csrc/split_shade.cu, csrc/render.cu):

    split_shade ...... models/texture.py:330-377 after the material networks: activations, mixes, FG LUT + emitter
                       lookups (lib/pbr/light.py:168-206), 7/24-channel packing (models/texture.py:345)
    split_render ..... models/split_mixed_occ.py:151-177 get_alpha + models/volrend.py:851-885 weights and the four
                       accumulations + the orientation map of models/split_mixed_occ.py:384-394
"""
import ctypes

import torch

from . import _lib as L


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])


def _args(raw_albedo, raw_rough, raw_metal, raw_env, normals, dirs, lut, diffuse, levels, lo, hi, stage, keep):
    a = L.SplitShadeC()
    a.raw_albedo, a.raw_roughness, a.raw_metallic, a.raw_env = (L.ptr(t) for t in (raw_albedo, raw_rough, raw_metal, raw_env))
    a.normals, a.dirs = L.ptr(normals), L.ptr(dirs)
    a.stage, a.n = int(stage), int(normals.shape[0])
    a.min_roughness, a.max_roughness = float(lo), float(hi)
    if stage != 0:
        a.fg_lut, a.lut_h, a.lut_w = L.ptr(lut), int(lut.shape[-3]), int(lut.shape[-2])
        a.diffuse, a.diffuse_res = L.ptr(diffuse), int(diffuse.shape[1])
        lv = _ptr_array(levels)
        res = (ctypes.c_int * len(levels))(*[int(t.shape[1]) for t in levels])
        keep += [lv, res]
        a.specular_levels = ctypes.cast(lv, ctypes.c_void_p)
        a.specular_res = ctypes.cast(res, ctypes.c_void_p)
        a.n_levels = len(levels)
    return a


class _SplitShade(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw_albedo, raw_rough, raw_metal, raw_env, normals, dirs, stage, lo, hi, lut, diffuse, *levels):
        f = lambda t: t.contiguous().float()
        raw_albedo, raw_rough, raw_metal, raw_env, normals, dirs = map(f, (raw_albedo, raw_rough, raw_metal, raw_env, normals, dirs))
        if stage != 0:
            lut, diffuse, levels = f(lut), f(diffuse), [f(t) for t in levels]
        n = normals.shape[0]
        out = torch.empty(n, 24 if stage != 0 else 7, device=normals.device, dtype=torch.float32)
        keep = []
        a = _args(raw_albedo, raw_rough, raw_metal, raw_env, normals, dirs, lut, diffuse, levels, lo, hi, stage, keep)
        L.call("rsdf_split_shade_fwd", ctypes.byref(a), L.ptr(out), L.stream())
        ctx.save_for_backward(raw_albedo, raw_rough, raw_metal, raw_env, normals, dirs,
                              *(([lut, diffuse] + list(levels)) if stage != 0 else []))
        ctx.cfg = (int(stage), float(lo), float(hi))
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, go):
        raw_albedo, raw_rough, raw_metal, raw_env, normals, dirs, *tex = ctx.saved_tensors
        stage, lo, hi = ctx.cfg
        lut, diffuse, levels = (tex[0], tex[1], tex[2:]) if stage != 0 else (None, None, [])
        keep = []
        a = _args(raw_albedo, raw_rough, raw_metal, raw_env, normals, dirs, lut, diffuse, levels, lo, hi, stage, keep)
        g = [torch.empty_like(t) for t in (raw_albedo, raw_rough, raw_metal, raw_env)]
        g_n = torch.empty_like(normals) if ctx.needs_input_grad[4] else None
        need = ctx.needs_input_grad
        g_lut = torch.zeros_like(lut) if (stage != 0 and need[9]) else None
        g_dif = torch.zeros_like(diffuse) if (stage != 0 and need[10]) else None
        g_lv = [torch.zeros_like(t) if need[11 + i] else None for i, t in enumerate(levels)]
        L.call("rsdf_split_shade_bwd", ctypes.byref(a), L.ptr(go.contiguous().float()), *(L.ptr(t) for t in g), L.ptr(g_n),
               L.ptr(g_lut), L.ptr(g_dif), _ptr_array(g_lv) if g_lv else None, L.stream())
        return (*g, g_n, None, None, None, None, g_lut, g_dif, *g_lv)


def split_shade(raw_albedo, raw_rough, raw_metal, raw_env, normals, dirs, emitter, fg_lut, stage):
    """-> colors [S, 7 | 24] in the channel order of models/texture.py:345 from the PRE-activation outputs of the albedo
    (6), roughness (1), metallic (2) and env (3) networks.  color_activation must be sigmoid (the only one the configs
    use; VolumeMixedMipSplitOcc falls back to the op-by-op path otherwise)."""
    L.require_cuda(raw_albedo, normals, dirs)
    if stage == 0:
        return _SplitShade.apply(raw_albedo, raw_rough, raw_metal, raw_env, normals, dirs, 0, 0.0, 0.0, None, None)
    return _SplitShade.apply(raw_albedo, raw_rough, raw_metal, raw_env, normals, dirs, int(stage), emitter.MIN_ROUGHNESS,
                             emitter.MAX_ROUGHNESS, fg_lut, emitter.diffuse, *emitter.specular)


class _SplitRender(torch.autograd.Function):
    """out[R, CD + 6] = (colours CD, sum w n (3), opacity, depth, sum w relu(d . n)); also returns weights, alphas."""

    @staticmethod
    def forward(ctx, packed, rays_d, t_starts, t_ends, sdf, normals, colors, inv_s, ratio):
        R, S, dev = packed.shape[0], sdf.shape[0], sdf.device
        sdf, normals, colors = sdf.contiguous().float(), normals.contiguous().float(), colors.contiguous().float()
        cd = colors.shape[1]
        ctx.inv_shape = inv_s.shape
        inv_s = inv_s.reshape(1).contiguous().float()
        alpha, w, T = (torch.empty(S, device=dev) for _ in range(3))
        out = torch.empty(R, cd + 6, device=dev)
        L.call("rsdf_split_render_fwd", L.ptr(packed), L.ptr(rays_d), L.ptr(t_starts), L.ptr(t_ends), L.ptr(sdf),
               L.ptr(normals), L.ptr(colors), cd, L.ptr(inv_s), float(ratio), R, L.ptr(alpha), L.ptr(w), L.ptr(T),
               L.ptr(out), L.stream())
        ctx.save_for_backward(packed, rays_d, t_starts, t_ends, sdf, normals, colors, inv_s, alpha, w, T)
        ctx.ratio = float(ratio)
        ctx.mark_non_differentiable(alpha)
        return out, w, alpha

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_out, g_w, _g_alpha):
        packed, rays_d, t_starts, t_ends, sdf, normals, colors, inv_s, alpha, w, T = ctx.saved_tensors
        R, S, dev, cd = packed.shape[0], sdf.shape[0], sdf.device, colors.shape[1]
        g_out = g_out.contiguous() if g_out is not None else torch.zeros(R, cd + 6, device=dev)
        g_w = g_w.contiguous() if g_w is not None else None
        g_sdf = torch.empty(S, device=dev)
        g_n = torch.empty(S, 3, device=dev)
        g_c = torch.empty(S, cd, device=dev)
        g_inv = torch.zeros(R, device=dev)
        L.call("rsdf_split_render_bwd", L.ptr(packed), L.ptr(rays_d), L.ptr(t_starts), L.ptr(t_ends), L.ptr(sdf),
               L.ptr(normals), L.ptr(colors), cd, L.ptr(alpha), L.ptr(w), L.ptr(T), L.ptr(g_out), L.ptr(g_w), L.ptr(inv_s),
               ctx.ratio, R, L.ptr(g_sdf), L.ptr(g_n), L.ptr(g_c), L.ptr(g_inv), L.stream())
        return None, None, None, None, g_sdf, g_n, g_c, g_inv.sum().reshape(ctx.inv_shape), None


def split_render(packed, rays_d, t_starts, t_ends, sdf, normals, colors, inv_s, cos_anneal_ratio):
    return _SplitRender.apply(packed, rays_d.contiguous(), t_starts.contiguous(), t_ends.contiguous(), sdf, normals, colors,
                              inv_s, cos_anneal_ratio)
