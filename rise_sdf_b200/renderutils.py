"""Drop-in for the two `lib.renderutils` entry points on the render path
(lib/pbr/light.py:174,179,180): `diffuse_cubemap` and `specular_cubemap`
(lib/renderutils/ops.py:391-458), backed by csrc/cubemap.cu.  Backward passes are deterministic
gathers (the reference scatters with fp32 atomics)."""
import numpy as np
import torch

from . import _lib as L

_tables = {}
_bounds = {}


def texel_table(res, device):
    """float4[6*res*res]: unit direction + solid angle of every texel (built once per resolution)."""
    key = (res, str(device))
    if key not in _tables:
        t = torch.empty(6 * res * res, 4, device=device, dtype=torch.float32)
        L.call("rsdf_cubemap_texel_table", res, L.ptr(t), L.stream())
        _tables[key] = t
    return _tables[key]


class _DiffuseCubemap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cubemap):
        cubemap = cubemap.contiguous()
        res = cubemap.shape[1]
        out = torch.empty_like(cubemap)
        L.call("rsdf_diffuse_cubemap", L.ptr(texel_table(res, cubemap.device)), L.ptr(cubemap), res, 0, L.ptr(out),
               L.stream())
        ctx.res = res
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        dout = dout.contiguous()
        g = torch.empty_like(dout)
        L.call("rsdf_diffuse_cubemap", L.ptr(texel_table(ctx.res, dout.device)), L.ptr(dout), ctx.res, 1, L.ptr(g),
               L.stream())
        return g


def diffuse_cubemap(cubemap, use_python=False):
    assert not use_python
    L.require_cuda(cubemap)
    assert cubemap.shape[0] == 6 and cubemap.shape[1] == cubemap.shape[2] and cubemap.shape[3] == 3
    return _DiffuseCubemap.apply(cubemap.float())


def ndf_cutoff(roughness, cutoff):
    """cos(theta) that retains `cutoff` of the GGX NDF energy: host-side numpy, bit-identical to
    lib/renderutils/ops.py:428-443 (1 000 000-point cumsum)."""
    def ndfGGX(alphaSqr, costheta):
        costheta = np.clip(costheta, 0.0, 1.0)
        d = (costheta * alphaSqr - costheta) * costheta + 1.0
        return alphaSqr / (d * d * np.pi)
    nSamples = 1000000
    costheta = np.cos(np.linspace(0, np.pi / 2.0, nSamples))
    D = np.cumsum(ndfGGX(roughness ** 4, costheta))
    idx = np.argmax(D >= D[..., -1] * cutoff)
    return float(costheta[idx])


def specular_bounds(res, costheta_cutoff, device):
    """-> float32 [6,res,res,24] (xmin,xmax,ymin,ymax per source face), ops.py:444 / cubemap.cu:181-244."""
    table = texel_table(res, device)
    nc = (res + 15) // 16 + 1
    scratch = torch.empty(6 * nc * nc, 4, device=device, dtype=torch.float32)
    bounds = torch.empty(6, res, res, 24, device=device, dtype=torch.float32)
    L.call("rsdf_specular_bounds", L.ptr(table), res, float(costheta_cutoff), L.ptr(scratch), L.ptr(bounds), L.stream())
    return bounds


def _ndf_bounds(res, roughness, cutoff, device):
    key = (res, roughness, cutoff, str(device))
    if key not in _bounds:
        c = ndf_cutoff(roughness, cutoff)
        _bounds[key] = (c, specular_bounds(res, c, device))
    return _bounds[key]


# ---- the prefilter as a cached sparse operator (csrc/cubemap.cu: specular_build / specular_apply) ------------------
# Training rebuilds the pyramid every step (systems/split_occ.py:151-152) from a map that is the only thing that
# changes: the 1.2 G pair weights of the six levels (4.9 GB per direction) are evaluated once per (res, roughness, cutoff) and
# streamed from HBM afterwards.  Used when autograd is recording (training); a one-off no-grad build (relighting's
# EnvSet) runs the direct kernel and allocates nothing.
OPERATOR_CACHE = True
OPERATOR_CACHE_MIN_RES = 32
_operators = {}


class _Operator:
    def __init__(self, res, roughness, cutoff, device):
        self.c, self.bounds = _ndf_bounds(res, roughness, cutoff, device)
        b = self.bounds.view(-1, 6, 4)
        area = ((b[..., 1] - b[..., 0] + 1).clamp(min=0) * (b[..., 3] - b[..., 2] + 1).clamp(min=0)).sum(-1)
        self.offset = torch.zeros(area.numel() + 1, device=device, dtype=torch.int64)
        torch.cumsum(area.to(torch.int64), 0, out=self.offset[1:])
        self.n_pairs = int(self.offset[-1])
        self.table = texel_table(res, device)
        # one array per direction: a pair's weight is evaluated from the forward OUTPUT texel's side either way
        self.weights = torch.empty(max(self.n_pairs, 1), device=device, dtype=torch.float32)
        self.weights_t = torch.empty(max(self.n_pairs, 1), device=device, dtype=torch.float32)
        self.wsum = torch.empty(6, res, res, 1, device=device, dtype=torch.float32)
        for tr, w in ((0, self.weights), (1, self.weights_t)):
            L.call("rsdf_specular_build", L.ptr(self.table), L.ptr(self.bounds), L.ptr(self.offset), res, float(roughness),
                   float(self.c), tr, L.ptr(w), L.ptr(self.wsum), L.stream())
        self.area4 = (self.table[:, 3] / 4.0).view(6, res, res, 1)
        self.res = res

    def apply(self, src, transposed):
        """src [6,res,res,3] -> [6,res,res,3]; the kernel gathers 16-byte texels, so the rgb is padded here"""
        src = torch.nn.functional.pad(src, (0, 1)).contiguous()
        dst = torch.empty(6, self.res, self.res, 3, device=src.device, dtype=torch.float32)
        L.call("rsdf_specular_apply", L.ptr(self.table), L.ptr(self.bounds), L.ptr(self.offset),
               L.ptr(self.weights_t if transposed else self.weights), L.ptr(src), self.res, int(transposed), L.ptr(dst),
               L.stream())
        return dst


def _operator(res, roughness, cutoff, device):
    key = (res, roughness, cutoff, str(device))
    if key not in _operators:
        _operators[key] = _Operator(res, roughness, cutoff, device)
    return _operators[key]


class _SpecularOperator(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cubemap, op):
        ctx.op = op
        rgb = op.apply(cubemap * op.area4, False)
        return torch.cat([rgb, op.wsum], -1)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        return ctx.op.apply(dout[..., 0:3], True), None


class _SpecularCubemap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cubemap, roughness, costheta_cutoff, bounds):
        cubemap = cubemap.contiguous()
        res = cubemap.shape[1]
        out = torch.empty(6, res, res, 4, device=cubemap.device, dtype=torch.float32)
        L.call("rsdf_specular_cubemap", L.ptr(texel_table(res, cubemap.device)), L.ptr(bounds), L.ptr(cubemap), res,
               float(roughness), float(costheta_cutoff), 0, L.ptr(out), L.stream())
        ctx.save_for_backward(bounds)
        ctx.args = (res, float(roughness), float(costheta_cutoff))
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        (bounds,) = ctx.saved_tensors
        res, roughness, cutoff = ctx.args
        g_rgb = dout[..., 0:3].contiguous()     # the weight-sum channel does not depend on the cubemap
        g = torch.empty(6, res, res, 3, device=dout.device, dtype=torch.float32)
        L.call("rsdf_specular_cubemap", L.ptr(texel_table(res, dout.device)), L.ptr(bounds), L.ptr(g_rgb), res,
               roughness, cutoff, 1, L.ptr(g), L.stream())
        return g, None, None, None


def specular_cubemap(cubemap, roughness, cutoff=0.99, use_python=False):
    assert not use_python
    L.require_cuda(cubemap)
    assert cubemap.shape[0] == 6 and cubemap.shape[1] == cubemap.shape[2], \
        "Bad shape for cubemap tensor: %s" % str(cubemap.shape)
    res = cubemap.shape[1]
    if OPERATOR_CACHE and res >= OPERATOR_CACHE_MIN_RES and torch.is_grad_enabled() and cubemap.requires_grad:
        out = _SpecularOperator.apply(cubemap.float().contiguous(), _operator(res, roughness, cutoff, cubemap.device))
    else:
        c, bounds = _ndf_bounds(res, roughness, cutoff, cubemap.device)
        out = _SpecularCubemap.apply(cubemap.float(), roughness, c, bounds)
    return out[..., 0:3] / out[..., 3:]
