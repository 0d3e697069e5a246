"""Pure-PyTorch CPU restatement of the field / encoding / shading arithmetic.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Every function is written with plain differentiable torch ops so that first- and
second-order gradients come from torch.autograd (that is how the reference obtains them:
models/geometry.py:224-228 `create_graph=True`).

Follows:
  hash grid ........ tiny-cuda-nn `HashGrid` encoding, version unpinned in the reference
                     (README.md:56, call sites models/network_utils.py:50,99); third-party,
                     absent from /root/reference -> restated from its published algorithm
                     (SURVEY.md Appendix A.1).  PARITY UNPINNED.
  SH ............... tiny-cuda-nn `SphericalHarmonics` (Appendix A.2).  PARITY UNPINNED.
  VanillaFrequency . models/network_utils.py:14-40
  CompositeEncoding  models/network_utils.py:71-79
  VanillaMLP ....... models/network_utils.py:109-157 (weight-norm: torch.nn.utils.weight_norm, dim=0)
  VolumeSDF ........ models/geometry.py:206-292, contract_to_unisphere :17-19, scale_anything models/utils.py:109-114
  VolumeRadiance ... models/texture.py:28-35
  get_alpha ........ models/neus.py:128-150
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

PRIMES = (1, 2654435761, 805459861)


# ----------------------------------------------------------------------------------------
# hash grid
# ----------------------------------------------------------------------------------------
class HashGridMeta:
    """Per-level constants, computed once on the host in float32 like tcnn's
    grid_scale()/grid_resolution() (scale = exp2f(l*log2f(pls))*base - 1, res = ceil(scale)+1,
    n_l = min(next_multiple(res^3, 8), 2^log2_T))."""

    def __init__(self, n_levels=16, n_features=2, log2_hashmap_size=19, base_resolution=16,
                 per_level_scale=1.447269237440378):
        self.n_levels, self.n_features = n_levels, n_features
        self.base_resolution, self.per_level_scale = base_resolution, per_level_scale
        log2_pls = np.float32(np.log2(per_level_scale))
        self.scale, self.res, self.size, self.offset = [], [], [], [0]
        for l in range(n_levels):
            s = np.float32(np.exp2(np.float32(np.float32(l) * log2_pls))) * np.float32(base_resolution) - np.float32(1.0)
            s = np.float32(s)
            r = int(np.ceil(s)) + 1
            dense = r ** 3
            n = min(dense, (2 ** 32 - 1) // 2)
            n = (n + 7) // 8 * 8
            n = min(n, 1 << log2_hashmap_size)
            self.scale.append(float(s))
            self.res.append(r)
            self.size.append(n)
            self.offset.append(self.offset[-1] + n)
        self.n_entries = self.offset[-1]
        self.n_params = self.n_entries * n_features
        self.n_output_dims = n_levels * n_features


def _grid_index(cx, cy, cz, res, size):
    """tcnn grid_index(): dense stride walk with early exit, coherent-prime hash otherwise.
    Integer tensors are int64 holding uint32 values."""
    M = 0xFFFFFFFF
    stride, dims = 1, 0
    idx = torch.zeros_like(cx)
    for c in (cx, cy, cz):
        if stride > size:
            break
        idx = (idx + c * stride) & M
        stride *= res
        dims += 1
    if size < stride:
        idx = ((cx * PRIMES[0]) & M) ^ ((cy * PRIMES[1]) & M) ^ ((cz * PRIMES[2]) & M)
    return idx % size


def hash_encode(x, table, meta, dtype=None):
    """x [S,3] in [0,1]; table flat [n_params] (level-major, features interleaved).
    Returns [S, L*F], level-major.  pos = fmaf(scale, x, 0.5) is emulated by a float64
    multiply-add rounded once to float32 (the product of two floats is exact in double)."""
    dtype = dtype or x.dtype
    Fdim = meta.n_features
    tab = table.view(-1, Fdim)
    outs = []
    for l in range(meta.n_levels):
        scale, res, size, off = meta.scale[l], meta.res[l], meta.size[l], meta.offset[l]
        if x.dtype == torch.float32:
            pos = (x.double() * scale + 0.5).float()
        else:
            pos = x * scale + 0.5
        cell = torch.floor(pos.detach())
        w = (pos - cell).to(dtype)
        c0 = cell.to(torch.int64) & 0xFFFFFFFF
        acc = 0
        for corner in range(8):
            bits = [(corner >> d) & 1 for d in range(3)]
            cc = [(c0[:, d] + bits[d]) & 0xFFFFFFFF for d in range(3)]
            wt = 1
            for d in range(3):
                wt = wt * (w[:, d] if bits[d] else (1 - w[:, d]))
            idx = _grid_index(cc[0], cc[1], cc[2], res, size) + off
            acc = acc + wt[:, None] * tab[idx].to(dtype)
        outs.append(acc)
    return torch.cat(outs, -1)


# ----------------------------------------------------------------------------------------
# other encodings
# ----------------------------------------------------------------------------------------
def sh_encode(u, degree):
    """tcnn SphericalHarmonics: input u in [0,1]^3 -> xyz = 2u-1; degree 4 -> 16, 5 -> 25."""
    d = u * 2 - 1
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    x4, y4, z4 = x2 * x2, y2 * y2, z2 * z2
    o = [torch.full_like(x, 0.28209479177387814)]
    if degree > 1:
        o += [-0.48860251190291987 * y, 0.48860251190291987 * z, -0.48860251190291987 * x]
    if degree > 2:
        o += [1.0925484305920792 * xy, -1.0925484305920792 * yz,
              0.94617469575755997 * z2 - 0.31539156525251999,
              -1.0925484305920792 * xz, 0.54627421529603959 * x2 - 0.54627421529603959 * y2]
    if degree > 3:
        o += [0.59004358992664352 * y * (-3.0 * x2 + y2), 2.8906114426405538 * xy * z,
              0.45704579946446572 * y * (1.0 - 5.0 * z2), 0.3731763325901154 * z * (5.0 * z2 - 3.0),
              0.45704579946446572 * x * (1.0 - 5.0 * z2), 1.4453057213202769 * z * (x2 - y2),
              0.59004358992664352 * x * (-x2 + 3.0 * y2)]
    if degree > 4:
        o += [2.5033429417967046 * xy * (x2 - y2), 1.7701307697799304 * yz * (-3.0 * x2 + y2),
              0.94617469575756008 * xy * (7.0 * z2 - 1.0), 0.66904654355728921 * yz * (3.0 - 7.0 * z2),
              -3.1735664074561294 * z2 + 3.7024941420321507 * z4 + 0.31735664074561293,
              0.66904654355728921 * xz * (3.0 - 7.0 * z2), 0.47308734787878004 * (x2 - y2) * (7.0 * z2 - 1.0),
              1.7701307697799304 * xz * (-x2 + 3.0 * y2),
              -3.7550144126950569 * x2 * y2 + 0.62583573544917614 * x4 + 0.62583573544917614 * y4]
    return torch.stack(o, -1)


def frequency_encode(x, n_freqs, x_scale=1.0, x_offset=0.0):
    """models/network_utils.py:27-33: [sin(f0 x), cos(f0 x), sin(f1 x), ...], f_k = 2^k, no pi."""
    x = x * x_scale + x_offset
    out = []
    for k in range(n_freqs):
        f = float(2 ** k)
        out += [torch.sin(f * x), torch.cos(f * x)]
    return torch.cat(out, -1)


# ----------------------------------------------------------------------------------------
# MLP
# ----------------------------------------------------------------------------------------
def effective_weight(layer):
    """layer: dict with 'weight' or ('weight_g','weight_v'), and 'bias'.
    weight_norm(dim=0): w = g * v / ||v||_row."""
    if "weight" in layer:
        return layer["weight"]
    v, g = layer["weight_v"], layer["weight_g"]
    return v * (g / v.norm(dim=1, keepdim=True))


def vanilla_mlp(x, layers, activation):
    """activation: 'softplus100' (sphere_init nets) or 'relu'.  No output activation here."""
    h = x
    for i, layer in enumerate(layers):
        h = F.linear(h, effective_weight(layer), layer["bias"])
        if i + 1 < len(layers):
            h = F.softplus(h, beta=100) if activation == "softplus100" else F.relu(h)
    return h


def mlp_layers_from_state(state, prefix):
    """Pull [{'weight'|'weight_g','weight_v','bias'}] for `prefix`layers.N.* out of a state dict."""
    out = []
    for n in range(64):  # nn.Sequential indices: Linear at even N, activations at odd N
        keys = [k for k in state if k.startswith(f"{prefix}layers.{n}.")]
        if keys:
            out.append({k.split(".")[-1]: state[k] for k in keys})
    return out


# ----------------------------------------------------------------------------------------
# VolumeSDF / VolumeRadiance / alpha
# ----------------------------------------------------------------------------------------
def cuda_scalar_div(x, s):
    """`tensor / python_scalar` AS THE REFERENCE EXECUTES IT: the reference only ever runs on
    CUDA, where PyTorch evaluates a float32 tensor divided by a scalar as a multiplication by
    the float32 reciprocal (probed on the B200 box with scripts/probe_div.py: 25% of the
    results differ from true division by one ulp, 0% from x * fl32(1/s)).  The contraction
    `scale_anything` (models/utils.py:109-114) feeds the hash-grid cell lookup, where one ulp
    of x01 moves the fine-level interpolation weights by ~1e-4, so the oracle follows the
    CUDA semantics here."""
    if x.dtype == torch.float32 and CUDA_SCALAR_DIV:
        return x * float(np.float32(1.0) / np.float32(s))
    return x / s


# True: follow the CUDA execution of `tensor / scalar` (what every product comparison needs).  tests/test_reference_host_cpu.py
# clears it while it compares with the reference's Python running on the CPU, where the same line is a true division.
CUDA_SCALAR_DIV = True


def scale_to_unit(p, radius):
    return cuda_scalar_div(p - (-radius), radius - (-radius)) * (1 - 0) + 0
def sdf_field(points, table, meta, mlp, radius, level_mask=None, with_grad=True,
              create_graph=False, include_xyz=True, dtype=None):
    """VolumeSDF.forward, grad_type='analytic' (models/geometry.py:206-228).
    points [S,3] world.  Returns sdf [S], grad [S,3] (d sdf / d world point), feature [S,48]."""
    p = points
    if with_grad:
        p = points.detach().clone().requires_grad_(True) if not points.requires_grad else points
    # `dtype` = arithmetic type of everything downstream of the (always fp32-faithful) cell
    # lookup; float64 gives a rounding-free yardstick for gradient comparisons.
    dtype = dtype or points.dtype
    x01 = cuda_scalar_div(p - (-radius), radius - (-radius))
    x01 = x01 * (1 - 0) + 0
    enc = hash_encode(x01, table, meta, dtype=dtype)
    if level_mask is not None:
        enc = enc * level_mask
    h = torch.cat([(x01 * 2.0 + (-1.0)).to(dtype), enc], -1) if include_xyz else enc
    out = vanilla_mlp(h, mlp, "softplus100")
    sdf = out[:, 0]
    if not with_grad:
        return sdf, None, out
    (grad,) = torch.autograd.grad(sdf, p, torch.ones_like(sdf), create_graph=create_graph,
                                  retain_graph=True)
    return sdf, grad, out


def sdf_field_fd(points, table, meta, mlp, radius, eps, level_mask=None):
    """grad_type='finite_difference' (models/geometry.py:229-244)."""
    def f(pw):
        x01 = scale_to_unit(pw, radius)
        enc = hash_encode(x01, table, meta)
        if level_mask is not None:
            enc = enc * level_mask
        return vanilla_mlp(torch.cat([x01 * 2.0 - 1.0, enc], -1), mlp, "softplus100")
    out = f(points)
    offs = torch.tensor([[eps, 0, 0], [-eps, 0, 0], [0, eps, 0], [0, -eps, 0], [0, 0, eps], [0, 0, -eps]],
                        dtype=points.dtype)
    pd = (points[:, None, :] + offs).clamp(-radius, radius)
    sd = f(pd.view(-1, 3))[:, 0].view(-1, 6)
    grad = cuda_scalar_div(0.5 * (sd[:, 0::2] - sd[:, 1::2]), eps)
    return out[:, 0], grad, out


def radiance(feature, dirs, normal, mlp, sh_degree=4):
    """VolumeRadiance.forward: sigmoid(MLP(cat[feature, SH((d+1)/2), normal]))."""
    emb = sh_encode((dirs + 1.0) / 2.0, sh_degree)
    return torch.sigmoid(vanilla_mlp(torch.cat([feature, emb, normal], -1), mlp, "relu"))


def get_alpha(sdf, normal, dirs, dists, inv_s, cos_anneal_ratio):
    """models/neus.py:128-150."""
    inv_s = inv_s.clip(1e-6, 1e6)
    true_cos = (dirs * normal).sum(-1, keepdim=True)
    iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal_ratio)
                 + F.relu(-true_cos) * cos_anneal_ratio)
    nxt = sdf[..., None] + iter_cos * dists.reshape(-1, 1) * 0.5
    prv = sdf[..., None] - iter_cos * dists.reshape(-1, 1) * 0.5
    prev_cdf = torch.sigmoid(prv * inv_s)
    next_cdf = torch.sigmoid(nxt * inv_s)
    p = prev_cdf - next_cdf
    c = prev_cdf
    return ((p + 1e-5) / (c + 1e-5)).view(-1).clip(0.0, 1.0)


def occ_alpha(sdf, inv_s, render_step_size):
    """occ_eval_fn of models/neus.py:101-112."""
    inv_s = inv_s.clip(1e-6, 1e6)
    nxt = sdf[..., None] - render_step_size * 0.5
    prv = sdf[..., None] + render_step_size * 0.5
    prev_cdf = torch.sigmoid(prv * inv_s)
    next_cdf = torch.sigmoid(nxt * inv_s)
    return ((prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)).view(-1, 1).clip(0.0, 1.0)


# ----------------------------------------------------------------------------------------
# volume rendering (differentiable torch restatement; the serial-order C twin is march.py)
# ----------------------------------------------------------------------------------------
def render_weight_from_alpha(alphas, ray_indices, n_rays):
    """T_i = prod_{j<i in ray}(1-a_j), w = T*a  (lib/nerfacc/vol_rendering.py:396-449).
    Padded [n_rays, max_len] cumprod so autograd gives the backward."""
    S = alphas.shape[0]
    if S == 0:
        return alphas, alphas
    counts = torch.bincount(ray_indices, minlength=n_rays)
    starts = torch.cumsum(counts, 0) - counts
    pos = torch.arange(S) - starts[ray_indices]
    L = int(counts.max())
    pad = torch.ones(n_rays, L + 1, dtype=alphas.dtype)
    pad = pad.index_put((ray_indices, pos + 1), 1.0 - alphas)
    T = torch.cumprod(pad, 1)[ray_indices, pos]
    return T * alphas, T


def accumulate_along_rays(weights, values, ray_indices, n_rays):
    """lib/nerfacc/vol_rendering.py:132-198."""
    src = weights[:, None] * values if values is not None else weights[:, None]
    out = torch.zeros(n_rays, src.shape[-1], dtype=src.dtype)
    if src.shape[0] == 0:
        return out
    return out.index_add(0, ray_indices, src)


# ----------------------------------------------------------------------------------------
# parameter init (for generating seeded synthetic weights; mirrors VanillaMLP.make_linear)
# ----------------------------------------------------------------------------------------
def init_mlp(dim_in, dim_out, n_neurons, n_hidden, sphere_init, weight_norm, gen,
             sphere_init_radius=0.5):
    dims = [dim_in] + [n_neurons] * n_hidden + [dim_out]
    layers = []
    for i in range(len(dims) - 1):
        fi, fo = dims[i], dims[i + 1]
        is_first, is_last = i == 0, i == len(dims) - 2
        W = torch.empty(fo, fi)
        b = torch.zeros(fo)
        if sphere_init:
            if is_last:
                b.fill_(-sphere_init_radius)
                W.normal_(math.sqrt(math.pi) / math.sqrt(fi), 0.0001, generator=gen)
            elif is_first:
                W.zero_()
                W[:, :3].normal_(0.0, math.sqrt(2) / math.sqrt(fo), generator=gen)
            else:
                W.normal_(0.0, math.sqrt(2) / math.sqrt(fo), generator=gen)
        else:
            bound = math.sqrt(2.0) * math.sqrt(3.0 / fi)  # kaiming_uniform_(nonlinearity='relu')
            W.uniform_(-bound, bound, generator=gen)
        if weight_norm:
            layers.append({"weight_g": W.norm(dim=1, keepdim=True).clone(), "weight_v": W, "bias": b})
        else:
            layers.append({"weight": W, "bias": b})
    return layers
