"""Drop-in for `nvdiffrast.torch.texture` in the three modes RISE-SDF uses
(lib/pbr/light.py:194-206, models/texture.py:340-341, lib/pbr/utils/light_utils.py:108,124,138):

    tex[1,H,W,C],     uv[1,S,1,2] | [1,H',W',2], filter 'linear', boundary 'clamp' | 'wrap' (the default; taps wrap around both axes)
    tex[1,6,N,N,3],   uv[1,S,1,3] | [1,H',W',3], filter 'linear', boundary 'cube'
    tex[1,6,N,N,3] + mip=[...] + mip_level_bias[1,S,1], filter 'linear-mipmap-linear', boundary 'cube'

Differentiable wrt tex, every mip level, uv / directions and mip_level_bias.  nvdiffrast itself is
third-party and absent from the reference tree: semantics restated in SURVEY.md Appendix A.3
(PARITY UNPINNED); the cube face convention is pinned by lib/renderutils/c_src/cubemap.cu:32-60.
"""
import ctypes

import torch

from . import _lib as L


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def _int_array(vals):
    return (ctypes.c_int * len(vals))(*vals)


class _Tex2D(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tex, uv, wrap=False):
        H, W, C = tex.shape
        n = uv.shape[0]
        out = torch.empty(n, C, device=uv.device, dtype=torch.float32)
        L.call("rsdf_tex2d_fwd", L.ptr(tex), H, W, C, int(wrap), L.ptr(uv), n, L.ptr(out), L.stream())
        ctx.save_for_backward(tex, uv)
        ctx.wrap = int(wrap)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, go):
        tex, uv = ctx.saved_tensors
        H, W, C = tex.shape
        n = uv.shape[0]
        g_tex = torch.zeros_like(tex) if ctx.needs_input_grad[0] else None
        g_uv = torch.empty_like(uv) if ctx.needs_input_grad[1] else None
        if g_tex is not None or g_uv is not None:
            L.call("rsdf_tex2d_bwd", L.ptr(tex), H, W, C, ctx.wrap, L.ptr(uv), L.ptr(go.contiguous()), n,
                   L.ptr(g_tex), L.ptr(g_uv), L.stream())
        return g_tex, g_uv, None


class _CubeSample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dirs, bias, *levels):
        n = dirs.shape[0]
        res = [int(t.shape[1]) for t in levels]
        out = torch.empty(n, 3, device=dirs.device, dtype=torch.float32)
        L.call("rsdf_cube_sample_fwd", _ptr_array(levels), _int_array(res), len(levels), L.ptr(dirs), L.ptr(bias), n,
               L.ptr(out), L.stream())
        ctx.save_for_backward(dirs, bias, *levels)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, go):
        dirs, bias, *levels = ctx.saved_tensors
        n = dirs.shape[0]
        res = [int(t.shape[1]) for t in levels]
        g_levels = [torch.zeros_like(t) if ctx.needs_input_grad[2 + i] else None for i, t in enumerate(levels)]
        g_dirs = torch.empty_like(dirs) if ctx.needs_input_grad[0] else None
        g_bias = torch.empty_like(bias) if (bias is not None and ctx.needs_input_grad[1]) else None
        L.call("rsdf_cube_sample_bwd", _ptr_array(levels), _ptr_array(g_levels), _int_array(res), len(levels),
               L.ptr(dirs), L.ptr(bias), L.ptr(go.contiguous()), n, L.ptr(g_bias), L.ptr(g_dirs), L.stream())
        return (g_dirs, g_bias, *g_levels)


def texture(tex, uv, uv_da=None, mip_level_bias=None, mip=None, filter_mode="auto", boundary_mode="wrap",
            max_mip_level=None):
    """nvdiffrast.torch.texture (subset).  Shapes as in the reference call sites; returns
    [1, ..., C] with uv's leading dims."""
    L.require_cuda(tex, uv)
    if uv_da is not None:
        raise NotImplementedError("uv_da (screen-space derivatives) is not used by RISE-SDF")
    if filter_mode == "auto":
        filter_mode = "linear-mipmap-linear" if mip is not None else "linear"
    lead = uv.shape[:-1]
    if boundary_mode == "cube":
        assert tex.dim() == 5 and tex.shape[0] == 1 and tex.shape[1] == 6 and tex.shape[-1] == 3, tex.shape
        dirs = uv.reshape(-1, 3).contiguous().float()
        levels = [tex[0].contiguous().float()]
        bias = None
        if filter_mode == "linear-mipmap-linear":
            assert mip is not None and mip_level_bias is not None, \
                "cube linear-mipmap-linear needs the explicit mip stack and mip_level_bias"
            levels += [m[0].contiguous().float() for m in mip]
            bias = mip_level_bias.reshape(-1).contiguous().float()
        elif filter_mode != "linear":
            raise NotImplementedError(filter_mode)
        out = _CubeSample.apply(dirs, bias, *levels)
        return out.view(*lead, 3)
    if boundary_mode not in ("clamp", "wrap"):
        raise NotImplementedError(boundary_mode)
    if filter_mode != "linear":
        raise NotImplementedError(filter_mode)
    assert tex.dim() == 4 and tex.shape[0] == 1, tex.shape
    uv2 = uv.reshape(-1, 2).contiguous().float()
    out = _Tex2D.apply(tex[0].contiguous().float(), uv2, boundary_mode == "wrap")
    return out.view(*lead, tex.shape[-1])
