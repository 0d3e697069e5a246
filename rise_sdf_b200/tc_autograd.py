"""Training-path matmuls on tcgen05 with full autograd support (first AND second order).

Three products, closed under differentiation, all executed by csrc/gemm_stream.cu:

    mm_nt(x[S,K], W[N,K]) = x W^T        d/dx -> mm_nn(g, W)      d/dW -> mm_tn(g, x)
    mm_nn(g[S,N], W[N,K]) = g W          d/dg -> mm_nt(gy, W)     d/dW -> mm_tn(g, gy)
    mm_tn(a[S,Fa], b[S,Fb]) = a^T b      d/da -> mm_nt(b, gG)     d/db -> mm_nn(a, gG)

Because every backward is expressed with the same differentiable primitives, torch autograd can
recurse through them: `torch.autograd.grad(sdf, points, create_graph=True)` followed by a backward
through that result (models/geometry.py:224-228; the eikonal loss) never leaves the tensor cores.
`linear(x, W, b)` is the drop-in for F.linear inside VanillaMLP's training path.
S is the (large) sample dimension; the small dimensions must be <= 128.
"""
import torch

from . import _lib as L
from .fused_mlp import pack_weight, pad16


def _stream(x, W, transposed):
    x = x.contiguous().float()
    W = W.contiguous().float()
    S, K = x.shape
    rows, cols = W.shape
    N = cols if transposed else rows
    assert (rows if transposed else cols) == K, (x.shape, W.shape, transposed)
    y = torch.empty(S, N, device=x.device, dtype=torch.float32)
    if S == 0:
        return y
    blob = pack_weight(W)
    L.call("rsdf_mm_stream", L.ptr(x), L.ptr(blob), None, L.ptr(y), S, K, N, pad16(rows), pad16(cols),
           1 if transposed else 0, 0, L.stream())
    return y


def _tn(a, b):
    a = a.contiguous().float()
    b = b.contiguous().float()
    S, Fa = a.shape
    Fb = b.shape[1]
    G = torch.zeros(Fa, Fb, device=a.device, dtype=torch.float32)
    if S:
        L.call("rsdf_mm_tn", L.ptr(a), L.ptr(b), L.ptr(G), S, Fa, Fb, L.stream())
    return G


class _MMNT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W):
        ctx.save_for_backward(x, W)
        return _stream(x, W, False)

    @staticmethod
    def backward(ctx, gy):
        x, W = ctx.saved_tensors
        gx = mm_nn(gy, W) if ctx.needs_input_grad[0] else None
        gW = mm_tn(gy, x) if ctx.needs_input_grad[1] else None
        return gx, gW


class _MMNN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g, W):
        ctx.save_for_backward(g, W)
        return _stream(g, W, True)

    @staticmethod
    def backward(ctx, gy):
        g, W = ctx.saved_tensors
        gg = mm_nt(gy, W) if ctx.needs_input_grad[0] else None
        gW = mm_tn(g, gy) if ctx.needs_input_grad[1] else None
        return gg, gW


class _MMTN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return _tn(a, b)

    @staticmethod
    def backward(ctx, gG):
        a, b = ctx.saved_tensors
        ga = mm_nt(b, gG) if ctx.needs_input_grad[0] else None
        gb = mm_nn(a, gG) if ctx.needs_input_grad[1] else None
        return ga, gb


def _ok(*ts):
    return all(t.is_cuda and t.dim() == 2 for t in ts)


def mm_nt(x, W):
    assert _ok(x, W) and W.shape[0] <= 128 and W.shape[1] <= 128
    return _MMNT.apply(x, W)


def mm_nn(g, W):
    assert _ok(g, W) and W.shape[0] <= 128 and W.shape[1] <= 128
    return _MMNN.apply(g, W)


def mm_tn(a, b):
    assert _ok(a, b) and a.shape[1] <= 128 and b.shape[1] <= 128
    return _MMTN.apply(a, b)


def linear(x, W, b=None):
    """F.linear(x, W, b) for x [S,K], W [N,K] with K, N <= 128."""
    y = mm_nt(x, W)
    return y if b is None else y + b
