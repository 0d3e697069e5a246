"""Host side of the fused SDF-field MLP kernels (csrc/sdf_train.cu): the `VanillaMLP` of
`VolumeSDF` (models/geometry.py:206-228; 35 -> 128 -> 128 -> 48, Softplus(beta=100), weight-norm folded
on the host) evaluated TOGETHER with g0 = d sdf / d h0 -- what the reference obtains through
`torch.autograd.grad(sdf, points, create_graph=True)` -- as ONE autograd node whose backward is ONE
kernel that also carries the second-order terms (eikonal loss, normal-dependent shading).

    out, g0 = fused_sdf(mlp, in0, scale0, shift0, in1)        h0 = cat(in0 * scale0 + shift0, in1)

`out`/`g0` are differentiable w.r.t. in0, in1 and every MLP parameter (first order in `out`, and through
`g0` second order in the network); the node itself is once-differentiable, which is all the render
path needs (no third-order terms exist in the reference's losses).
"""
import ctypes

import torch

from . import _lib as L
from .fused_mlp import pack_weight

HID, KP = 128, 48


def supports(mlp, dim_in=None):
    """True when `mlp` is the SDF-shaped VanillaMLP the fused kernels are built for."""
    return (getattr(mlp, "sphere_init", False) and mlp.n_hidden_layers == 2 and mlp.n_neurons == HID
            and (dim_in or mlp.dim_in) <= KP and mlp.dim_out <= KP)


def _net_struct(W1, b1, W2, b2, W3, b3):
    keep = [pack_weight(W1, HID, KP), pack_weight(W2, HID, HID), pack_weight(W3, KP, HID),
            b1.detach().contiguous().float(), b2.detach().contiguous().float(), b3.detach().contiguous().float(),
            W3.detach()[0].contiguous().float()]
    c = L.SdfMlpC()
    c.w1_blob, c.w2_blob, c.w3_blob = (k.data_ptr() for k in keep[:3])
    c.b1, c.b2, c.b3, c.w3_row0 = (k.data_ptr() for k in keep[3:])
    c.n_in, c.n_out = W1.shape[1], W3.shape[0]
    return c, keep


class _FusedSDF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, in0, in1, W1, b1, W2, b2, W3, b3, scale0, shift0, want_g0):
        L.require_cuda(in0, in1, W1)
        in0 = in0.contiguous().float()
        in1 = None if in1 is None else in1.contiguous().float()
        S, w0 = in0.shape
        w1 = 0 if in1 is None else in1.shape[1]
        net, keep = _net_struct(W1, b1, W2, b2, W3, b3)
        out = torch.empty(S, W3.shape[0], device=in0.device, dtype=torch.float32)
        g0 = torch.empty(S, w0 + w1, device=in0.device, dtype=torch.float32) if want_g0 else None
        if S:
            L.call("rsdf_sdf_mlp_fwd", ctypes.byref(net), L.ptr(in0), w0, float(scale0), float(shift0), L.ptr(in1), w1,
                   S, L.ptr(out), L.ptr(g0), L.stream())
        ctx.save_for_backward(in0, in1, W1, b1, W2, b2, W3, b3)
        ctx.net, ctx.keep, ctx.scale0, ctx.shift0 = net, keep, float(scale0), float(shift0)
        if not want_g0:
            g0 = in0.new_zeros(0)
            ctx.mark_non_differentiable(g0)
        return out, g0

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_out, g_g0):
        in0, in1, W1, b1, W2, b2, W3, b3 = ctx.saved_tensors
        S, w0 = in0.shape
        w1 = 0 if in1 is None else in1.shape[1]
        dev = in0.device
        g_out = g_out.contiguous().float()
        g_g0 = None if (g_g0 is None or g_g0.numel() == 0) else g_g0.contiguous().float()
        need_in = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        g_in = torch.empty(S, w0 + w1, device=dev, dtype=torch.float32) if need_in else None
        gW1, gb1, gW2, gb2 = torch.zeros_like(W1), torch.zeros_like(b1), torch.zeros_like(W2), torch.zeros_like(b2)
        gW3, gb3 = torch.zeros_like(W3), torch.zeros_like(b3)
        if S:
            amax = torch.empty(1, device=dev, dtype=torch.int32)
            L.call("rsdf_absmax2", L.ptr(g_out), g_out.numel(), L.ptr(g_g0), 0 if g_g0 is None else g_g0.numel(),
                   L.ptr(amax), L.stream())
            L.call("rsdf_sdf_mlp_bwd", ctypes.byref(ctx.net), L.ptr(in0), w0, ctx.scale0, ctx.shift0, L.ptr(in1), w1, S,
                   L.ptr(g_out), L.ptr(g_g0), L.ptr(amax), L.ptr(g_in), L.ptr(gW1), L.ptr(gb1), L.ptr(gW2),
                   L.ptr(gb2), L.ptr(gW3), L.ptr(gb3), L.stream())
        elif need_in:
            g_in.zero_()
        g_in0 = g_in1 = None
        if ctx.needs_input_grad[0]:
            g_in0 = g_in[:, :w0] * ctx.scale0
        if ctx.needs_input_grad[1] and in1 is not None:
            g_in1 = g_in[:, w0:]
        return g_in0, g_in1, gW1, gb1, gW2, gb2, gW3, gb3, None, None, None


def fused_sdf(mlp, in0, scale0=1.0, shift0=0.0, in1=None, want_g0=True):
    """-> (out [S, dim_out], g0 [S, dim_in] or None).  `mlp`: a VanillaMLP for which supports() holds."""
    (W1, b1), (W2, b2), (W3, b3) = mlp.effective_weights()
    out, g0 = _FusedSDF.apply(in0, in1, W1.float(), b1.float(), W2.float(), b2.float(), W3.float(), b3.float(),
                              scale0, shift0, want_g0)
    return out, (g0 if want_g0 else None)
