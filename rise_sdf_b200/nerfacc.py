"""Drop-in for the `nerfacc` (0.5.3 API) and `lib.nerfacc` (vendored 0.3.5 API) symbols that
RISE-SDF imports on its render path, backed by librsdf_b200.so.

Import sites replaced:
    models/neus.py:11-18, models/split_mixed_occ.py:12-18, models/volrend.py:10-14,
    models/geometry.py:14
Reference semantics followed:
    lib/nerfacc/ray_marching.py:14-222, lib/nerfacc/grid.py:113-294,
    lib/nerfacc/vol_rendering.py:132-198,396-520, lib/nerfacc/intersection.py:13-49,
    lib/nerfacc/pack.py.
`OccGridEstimator.sampling` (nerfacc 0.5.3; source not in the reference tree) is defined as the
in-tree 0.3.5 `ray_marching` against the estimator's own AABB with outputs squeezed to [S]
(SURVEY.md Appendix A.5).
"""
from enum import Enum
from typing import Callable, Optional

import torch
from torch import Tensor, nn

from . import _lib as L


class ContractionType(Enum):
    """lib/nerfacc/contraction.py; only AABB is on the hot path."""
    AABB = 0
    UN_BOUNDED_TANH = 1
    UN_BOUNDED_SPHERE = 2


# ----------------------------------------------------------------------------------------
# marching
# ----------------------------------------------------------------------------------------
@torch.no_grad()
def ray_aabb_intersect(rays_o: Tensor, rays_d: Tensor, aabb: Tensor):
    """lib/nerfacc/intersection.py:13-49 -> (t_min, t_max); miss = (1e10, 1e10), t_min >= 0."""
    L.require_cuda(rays_o, rays_d)
    if isinstance(aabb, Tensor) and aabb.dim() == 2:
        raise NotImplementedError(
            "nerfacc>=0.5 multi-box ray_aabb_intersect is only used by the learned-background "
            "branch (models/neus.py:164), which is out of scope")
    rays_o, rays_d = rays_o.contiguous().float(), rays_d.contiguous().float()
    n = rays_o.shape[0]
    t_min = torch.empty(n, device=rays_o.device, dtype=torch.float32)
    t_max = torch.empty_like(t_min)
    L.call("rsdf_ray_aabb_intersect", L.ptr(rays_o), L.ptr(rays_d), L.host6(aabb), n,
           L.ptr(t_min), L.ptr(t_max), L.stream())
    return t_min, t_max


def pack_bits(grid_binary: Tensor) -> Tensor:
    g = grid_binary.contiguous().view(torch.uint8) if grid_binary.dtype == torch.bool else grid_binary.contiguous()
    n = g.numel()
    assert n % 32 == 0
    bits = torch.empty(n // 32, device=g.device, dtype=torch.int32)
    L.call("rsdf_grid_pack_bits", L.ptr(g), n, L.ptr(bits), L.stream())
    return bits


KEEP_BUDGET_BYTES = 2 << 30      # scratch of the one-march path: n_rays * cap * 8 bytes (131 072 rays x 1026 slots: 1.08 GB)


@torch.no_grad()
def _march(rays_o, rays_d, t_min, t_max, roi_host, grid_binary, grid_bits, step, cone, span=None):
    """== _C.ray_marching (lib/nerfacc/cuda/csrc/ray_marching.cu:194-289).
    Returns packed_info int32[R,2], ray_indices int64[S], t_starts[S], t_ends[S].

    `span`: an upper bound of t_max - t_min known on the host (AABB diagonal, far - near plane).  With it the
    count round keeps every ray's intervals (at most span/step + 2 of them) and the fill round is a copy; without
    it, when the scratch would exceed KEEP_BUDGET_BYTES, or if a ray overflows its slots after all, the fill
    round marches a second time like the reference.  Both orders produce the same bits (test_gpu_march)."""
    n = rays_o.shape[0]
    dev = rays_o.device
    rx, ry, rz = grid_binary.shape[-3:]
    packed = torch.empty(n, 2, device=dev, dtype=torch.int32)
    tmp = torch.empty(n + n // 1024 + 3, device=dev, dtype=torch.int32)
    total = torch.zeros(2, device=dev, dtype=torch.int32)
    gb = grid_binary.view(torch.uint8) if grid_binary.dtype == torch.bool else grid_binary
    st = L.stream()
    cap = 0
    if span is not None and step > 0 and n > 0 and span / step < 1e6:
        cap = int(span / step) + 2
        if cap * n * 8 > KEEP_BUDGET_BYTES:
            cap = 0
    if cap:
        keep = torch.empty(n, cap, 2, device=dev, dtype=torch.float32)
        L.call("rsdf_march_count_keep", L.ptr(rays_o), L.ptr(rays_d), L.ptr(t_min), L.ptr(t_max), roi_host,
               L.ptr(gb), L.ptr(grid_bits), rx, ry, rz, float(step), float(cone), n, L.ptr(packed),
               L.ptr(tmp), L.ptr(total), L.ptr(keep), cap, st)
    else:
        L.call("rsdf_march_count", L.ptr(rays_o), L.ptr(rays_d), L.ptr(t_min), L.ptr(t_max), roi_host,
               L.ptr(gb), L.ptr(grid_bits), rx, ry, rz, float(step), float(cone), n, L.ptr(packed),
               L.ptr(tmp), L.ptr(total), st)
    S, overflow = total.tolist()   # the one host sync of the march (reference: ray_marching.cu:261)
    ri = torch.empty(S, device=dev, dtype=torch.int64)
    ts = torch.empty(S, device=dev, dtype=torch.float32)
    te = torch.empty(S, device=dev, dtype=torch.float32)
    if S > 0:
        if cap and not overflow:
            L.call("rsdf_march_compact", L.ptr(packed), L.ptr(keep), cap, n, L.ptr(ri), L.ptr(ts), L.ptr(te), st)
        else:
            L.call("rsdf_march_fill", L.ptr(rays_o), L.ptr(rays_d), L.ptr(t_min), L.ptr(t_max), roi_host,
                   L.ptr(gb), L.ptr(grid_bits), rx, ry, rz, float(step), float(cone), n, L.ptr(packed),
                   L.ptr(ri), L.ptr(ts), L.ptr(te), st)
    return packed, ri, ts, te


@torch.no_grad()
def ray_marching(rays_o: Tensor, rays_d: Tensor, t_min: Optional[Tensor] = None,
                 t_max: Optional[Tensor] = None, scene_aabb: Optional[Tensor] = None,
                 grid=None, sigma_fn: Optional[Callable] = None, alpha_fn: Optional[Callable] = None,
                 early_stop_eps: float = 1e-4, alpha_thre: float = 0.0,
                 near_plane: Optional[float] = None, far_plane: Optional[float] = None,
                 render_step_size: float = 1e-3, stratified: bool = False, cone_angle: float = 0.0,
                 _return_packed: bool = False, _return_mask: bool = False):
    """lib/nerfacc/ray_marching.py:14-222 (vendored nerfacc 0.3.5 signature).
    Returns ray_indices int64[S], t_starts [S,1], t_ends [S,1].  `_return_mask` (ours) appends, for every
    surviving sample, its row in the concatenated outputs of the alpha_fn / sigma_fn calls (None without one), so
    a caller can reuse what the visibility pass computed instead of evaluating the survivors again."""
    L.require_cuda(rays_o, rays_d)
    if alpha_fn is not None and sigma_fn is not None:
        raise ValueError("Only one of `alpha_fn` and `sigma_fn` should be provided.")
    rays_o, rays_d = rays_o.contiguous().float(), rays_d.contiguous().float()
    span = None          # host-side upper bound of t_max - t_min (unit directions; _march re-checks on the device)
    if t_min is None or t_max is None:
        if scene_aabb is not None:
            box = grid.roi_host if (grid is not None and scene_aabb is getattr(grid, "_aabb0", None)) \
                else L.host6(scene_aabb)
            t_min, t_max = ray_aabb_intersect(rays_o, rays_d, box)
            span = sum((box[k + 3] - box[k]) ** 2 for k in range(3)) ** 0.5
        else:
            t_min = torch.zeros_like(rays_o[..., 0])
            t_max = torch.ones_like(rays_o[..., 0]) * 1e10
    if near_plane is not None:
        t_min = torch.clamp(t_min, min=near_plane)
    if far_plane is not None:
        t_max = torch.clamp(t_max, max=far_plane)
        if far_plane < 1e9:
            span = min(span if span is not None else 1e30, far_plane - max(near_plane or 0.0, 0.0))
    if stratified:
        t_min = t_min + torch.rand_like(t_min) * render_step_size
    if grid is not None:
        if grid.contraction_type != ContractionType.AABB:
            raise NotImplementedError("only ContractionType.AABB is on the hot path")
        roi_host, gbin, gbits = grid.roi_host, grid.binary, grid.bits
    else:
        roi_host = L.host6([-1e10] * 3 + [1e10] * 3)
        gbin = torch.ones(1, 1, 1, dtype=torch.bool, device=rays_o.device)
        gbits = None
    packed, ri, ts, te = _march(rays_o, rays_d, t_min.contiguous(), t_max.contiguous(), roi_host,
                                gbin.contiguous(), gbits, render_step_size, cone_angle, span)
    ts, te = ts[:, None], te[:, None]
    masks = rows = None
    if sigma_fn is not None or alpha_fn is not None:
        if sigma_fn is not None:
            sigmas = sigma_fn(ts, te, ri)
            assert sigmas.shape == ts.shape, f"sigmas must have shape of (N, 1)! Got {sigmas.shape}"
            alphas = 1.0 - torch.exp(-sigmas * (te - ts))
        elif VISIBILITY_CHUNKS and early_stop_eps > 0 and ts.shape[0] > 0:
            alphas, rows = _alphas_front_to_back(alpha_fn, packed, ri, ts, te, early_stop_eps)
        else:
            alphas = alpha_fn(ts, te, ri)
            assert alphas.shape == ts.shape, f"alphas must have shape of (N, 1)! Got {alphas.shape}"
        masks = render_visibility(alphas, packed_info=packed, early_stop_eps=early_stop_eps,
                                  alpha_thre=alpha_thre)
        # ONE compaction (one host read-back) shared by the four gathers: `x[bool_mask]` runs nonzero + a sync each time
        keep_idx = torch.nonzero(masks.reshape(-1))[:, 0]
        ri, ts, te = ri[keep_idx], ts[keep_idx], te[keep_idx]
        packed = None
    rv = (ri, ts, te) + ((packed,) if _return_packed else ())
    if _return_mask:
        # row of every surviving sample in the concatenation of alpha_fn's outputs over its calls
        rv += (None if masks is None else (rows[keep_idx] if rows is not None else keep_idx),)
    return rv


VISIBILITY_CHUNKS = (64, 128, 256)      # samples per ray in rounds 1, 2, 3; a last round takes the rest


def _transmittance(packed, alphas):
    """T [S] of the scan kernel behind render_weight_from_alpha / render_visibility (no autograd)."""
    T = torch.empty(alphas.shape[0], device=alphas.device, dtype=torch.float32)
    L.call("rsdf_weight_from_alpha_fwd", L.ptr(packed), L.ptr(alphas), packed.shape[0], None, L.ptr(T), L.stream())
    return T


def _round_plan_kernels(packed, T, done, chunk, eps, ts_flat, te_flat):
    """One visibility round on the device (csrc/glue.cu): who is still alive (1 launch), prefix sum, the round's candidate
    list with its gathered t-values and ray indices (1 launch) -- instead of the ~25 torch launches of
    `_round_plan_torch`, which sit on the critical path right after the round's read-back.
    -> (idx int64[total], t_starts[total,1], t_ends[total,1], ray_indices int64[total]) or None when nothing is left."""
    n_rays, S0, dev, st = packed.shape[0], ts_flat.shape[0], packed.device, L.stream()
    lens = torch.empty(n_rays, device=dev, dtype=torch.int64)
    L.call("rsdf_vis_round_lens", L.ptr(packed), L.ptr(T), done, min(chunk, 2 ** 30), float(eps), n_rays, S0, L.ptr(lens), st)
    csum = torch.cumsum(lens, 0)
    total = int(csum[-1])
    if total == 0:
        return None
    idx = torch.empty(total, device=dev, dtype=torch.int64)
    ts_sel = torch.empty(total, 1, device=dev, dtype=torch.float32)
    te_sel = torch.empty(total, 1, device=dev, dtype=torch.float32)
    ri_sel = torch.empty(total, device=dev, dtype=torch.int64)
    L.call("rsdf_vis_round_fill", L.ptr(packed), L.ptr(lens), L.ptr(csum), done, n_rays, L.ptr(ts_flat), L.ptr(te_flat),
           L.ptr(idx), L.ptr(ts_sel), L.ptr(te_sel), L.ptr(ri_sel), st)
    return idx, ts_sel, te_sel, ri_sel


def _round_plan_torch(packed, T, done, chunk, eps, ts_flat, te_flat):
    """The same bookkeeping as index arithmetic: the executable specification of the two kernels above
    (tests/test_host_logic.py pins it on the CPU, tests/test_gpu_glue.py compares the kernels with it).  `sampling`
    itself only accepts CUDA rays."""
    n_rays, S0, dev = packed.shape[0], ts_flat.shape[0], packed.device
    base, count = packed[:, 0].long(), packed[:, 1].long()
    active = count > done
    if done > 0:
        active &= T[(base + done).clamp(max=S0 - 1)] >= eps
    lens = torch.where(active, (count - done).clamp(max=chunk), torch.zeros_like(count))
    total = int(lens.sum())
    if total == 0:
        return None
    first = torch.cumsum(lens, 0) - lens
    ray_of = torch.repeat_interleave(torch.arange(n_rays, device=dev), lens, output_size=total)
    idx = (base + done)[ray_of] + (torch.arange(total, device=dev) - first[ray_of])
    return idx, ts_flat[idx][:, None], te_flat[idx][:, None], ray_of


def _alphas_front_to_back(alpha_fn, packed, ri, ts, te, early_stop_eps):
    """The visibility pass of lib/nerfacc/ray_marching.py:198-218 evaluates `alpha_fn` on every marched candidate
    and then masks the ones behind `T < early_stop_eps`.  A sample's transmittance depends only on the samples in
    front of it, so the candidates are evaluated front to back in rounds; a ray whose transmittance AT ITS NEXT
    SAMPLE -- computed by the same scan kernel, from the same alphas, as the final mask -- is already below the
    threshold is dropped from the later rounds.  Never-evaluated samples keep alpha 0 and are masked by their
    transmittance exactly as before: the surviving set and its alphas are bit-identical to the one-shot pass
    (tests/test_gpu_march.py::test_front_to_back_visibility_*).

    Returns alphas [S0, 1] and rows int64[S0]: the row of each sample in the concatenation of alpha_fn's outputs
    over the rounds (-1: never evaluated)."""
    S0, dev = ts.shape[0], ts.device
    n_rays = packed.shape[0]
    packed = packed.contiguous()
    ts_flat, te_flat = ts.reshape(-1).contiguous(), te.reshape(-1).contiguous()
    alphas = torch.zeros(S0, 1, device=dev, dtype=torch.float32)
    rows = torch.full((S0,), -1, device=dev, dtype=torch.int64)
    max_count = int(packed[:, 1].max())
    done = n_eval = k = 0
    while done < max_count:
        chunk = VISIBILITY_CHUNKS[k] if k < len(VISIBILITY_CHUNKS) else max_count
        T = _transmittance(packed, alphas) if done > 0 else None
        plan = (_round_plan_kernels if packed.is_cuda else _round_plan_torch)(packed, T, done, chunk, early_stop_eps,
                                                                             ts_flat, te_flat)
        if plan is None:
            break
        idx, ts_sel, te_sel, ri_sel = plan
        total = idx.shape[0]
        a = alpha_fn(ts_sel, te_sel, ri_sel)
        assert a.shape == (total, 1), f"alphas must have shape of (N, 1)! Got {a.shape}"
        if packed.is_cuda:
            L.call("rsdf_vis_round_scatter", L.ptr(idx), L.ptr(a.float().contiguous()), n_eval, total, L.ptr(alphas),
                   L.ptr(rows), L.stream())
        else:
            alphas[idx] = a.float()
            rows[idx] = torch.arange(n_eval, n_eval + total, device=dev)
        n_eval += total
        done += chunk
        k += 1
    return alphas, rows


@torch.no_grad()
def pack_info(ray_indices: Tensor, n_rays: int) -> Tensor:
    """lib/nerfacc/pack.py: sorted ray_indices -> packed_info int32[n_rays, 2] (base, count)."""
    L.require_cuda(ray_indices)
    ri = ray_indices.contiguous().long()
    packed = torch.empty(n_rays, 2, device=ri.device, dtype=torch.int32)
    L.call("rsdf_pack_info", L.ptr(ri), ri.numel(), n_rays, L.ptr(packed), L.stream())
    return packed


# ----------------------------------------------------------------------------------------
# scan + accumulate (autograd)
# ----------------------------------------------------------------------------------------
class _WeightFromAlpha(torch.autograd.Function):
    @staticmethod
    def forward(ctx, packed, alphas):
        alphas = alphas.contiguous()
        n_rays = packed.shape[0]
        w = torch.empty_like(alphas)
        T = torch.empty_like(alphas)
        L.call("rsdf_weight_from_alpha_fwd", L.ptr(packed), L.ptr(alphas), n_rays, L.ptr(w), L.ptr(T),
               L.stream())
        ctx.save_for_backward(packed, alphas, w, T)
        return w, T

    @staticmethod
    def backward(ctx, gw, gT):
        packed, alphas, w, T = ctx.saved_tensors
        gw = gw.contiguous() if gw is not None else None
        gT = gT.contiguous() if gT is not None else None
        ga = torch.zeros_like(alphas)
        L.call("rsdf_weight_from_alpha_bwd", L.ptr(packed), L.ptr(alphas), L.ptr(w), L.ptr(T), L.ptr(gw),
               L.ptr(gT), packed.shape[0], L.ptr(ga), L.stream())
        return None, ga


def _packed_for(ray_indices, packed_info, n_rays):
    if packed_info is not None:
        return packed_info.contiguous().int()
    if n_rays is None:
        n_rays = int(ray_indices.max()) + 1 if ray_indices.numel() else 0
    return pack_info(ray_indices, n_rays)


def render_weight_from_alpha(alphas: Tensor, *, packed_info: Optional[Tensor] = None,
                             ray_indices: Optional[Tensor] = None, n_rays: Optional[int] = None):
    """nerfacc 0.5.3 form used at models/neus.py:262, models/volrend.py:105,264,851:
    returns (weights, trans), both shaped like `alphas`, differentiable wrt alphas."""
    assert ray_indices is not None or packed_info is not None, \
        "Either ray_indices or packed_info should be provided."
    L.require_cuda(alphas)
    shape = alphas.shape
    if alphas.numel() == 0:
        return alphas, alphas
    packed = _packed_for(ray_indices, packed_info, n_rays)
    w, T = _WeightFromAlpha.apply(packed, alphas.reshape(-1).float())
    return w.view(shape), T.view(shape)


def render_transmittance_from_alpha(alphas, *, packed_info=None, ray_indices=None, n_rays=None):
    return render_weight_from_alpha(alphas, packed_info=packed_info, ray_indices=ray_indices, n_rays=n_rays)[1]


@torch.no_grad()
def render_visibility(alphas: Tensor, *, ray_indices: Optional[Tensor] = None,
                      packed_info: Optional[Tensor] = None, n_rays: Optional[int] = None,
                      early_stop_eps: float = 1e-4, alpha_thre: float = 0.0) -> Tensor:
    """lib/nerfacc/vol_rendering.py:453-520."""
    if alphas.numel() == 0:
        return torch.zeros(0, dtype=torch.bool, device=alphas.device)
    _, T = render_weight_from_alpha(alphas, packed_info=packed_info, ray_indices=ray_indices, n_rays=n_rays)
    vis = T >= early_stop_eps
    if alpha_thre > 0:
        vis = vis & (alphas >= alpha_thre)
    return vis.reshape(alphas.shape[0], -1)[:, 0] if alphas.dim() > 1 else vis


class _Accumulate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, packed, ray_indices, weights, values):
        n_rays = packed.shape[0]
        weights = weights.contiguous()
        D = 1 if values is None else values.shape[-1]
        if values is not None:
            values = values.contiguous()
        out = torch.empty(n_rays, D, device=weights.device, dtype=torch.float32)
        L.call("rsdf_accumulate_fwd", L.ptr(packed), L.ptr(weights), L.ptr(values), n_rays, D, L.ptr(out),
               L.stream())
        ctx.save_for_backward(ray_indices, weights, values)
        ctx.D = D
        return out

    @staticmethod
    def backward(ctx, go):
        ray_indices, weights, values = ctx.saved_tensors
        go = go.contiguous()
        S = weights.shape[0]
        gw = torch.empty_like(weights) if ctx.needs_input_grad[2] else None
        gv = torch.empty_like(values) if (values is not None and ctx.needs_input_grad[3]) else None
        if gw is not None or gv is not None:
            L.call("rsdf_accumulate_bwd", L.ptr(ray_indices), L.ptr(weights), L.ptr(values), L.ptr(go), S,
                   ctx.D, L.ptr(gw), L.ptr(gv), L.stream())
        return None, None, gw, gv


def accumulate_along_rays(weights: Tensor, values: Optional[Tensor] = None,
                          ray_indices: Optional[Tensor] = None, n_rays: Optional[int] = None,
                          packed_info: Optional[Tensor] = None) -> Tensor:
    """nerfacc 0.5.3 form (models/neus.py:265-276): weights [S] (or [S,1]), values [S,D] | None
    -> [n_rays, D|1].  Deterministic segmented reduction (no atomics)."""
    L.require_cuda(weights)
    assert ray_indices is not None
    w = weights.reshape(-1).float()
    if values is not None:
        assert values.dim() == 2 and values.shape[0] == w.shape[0], \
            f"Invalid shapes: {values.shape} vs {weights.shape}"
    if w.numel() == 0:
        assert n_rays is not None
        return torch.zeros((n_rays, 1 if values is None else values.shape[-1]), device=weights.device)
    ri = ray_indices.contiguous().long()
    packed = _packed_for(ri, packed_info, n_rays)
    return _Accumulate.apply(packed, ri, w, None if values is None else values.float())


def render_weight_from_density(*args, **kwargs):
    raise NotImplementedError("density rendering is only used by the learned-background branch "
                              "(models/neus.py:195-197): out of scope (SURVEY.md §8b)")


# ----------------------------------------------------------------------------------------
# occupancy grid
# ----------------------------------------------------------------------------------------
class OccGridEstimator(nn.Module):
    """nerfacc 0.5.3 `OccGridEstimator(roi_aabb, resolution=128)` as used at models/neus.py:75-78.
    Buffers keep the 0.5.3 names (`aabbs[1,6]`, `occs[R^3]`, `binaries[1,R,R,R]`) so reference
    checkpoints load; the update rule is lib/nerfacc/grid.py:196-239."""

    def __init__(self, roi_aabb, resolution: int = 128, levels: int = 1):
        super().__init__()
        if levels != 1:
            raise NotImplementedError("multi-level grids are not used by RISE-SDF")
        roi_aabb = torch.as_tensor(roi_aabb, dtype=torch.float32).flatten()
        assert roi_aabb.shape == (6,)
        self.n_cells = int(resolution) ** 3
        self.register_buffer("resolution", torch.tensor([resolution] * 3, dtype=torch.int32))
        self.register_buffer("aabbs", roi_aabb[None, :].clone())
        self.register_buffer("occs", torch.zeros(self.n_cells))
        self.register_buffer("binaries", torch.zeros(1, resolution, resolution, resolution, dtype=torch.bool))
        self._res = int(resolution)
        self.roi_host = L.host6(roi_aabb)
        self._bits = None
        self._bits_version = -1
        self.contraction_type = ContractionType.AABB

    # 0.3.5 `Grid` surface used by lib.nerfacc.ray_marching
    @property
    def binary(self):
        return self.binaries[0]

    @property
    def roi_aabb(self):
        return self.aabbs[0]

    @property
    def bits(self):
        key = (self.binaries.data_ptr(), self.binaries._version)
        if self._bits is None or self._bits_version != key:
            self._bits = pack_bits(self.binaries)
            self._bits_version = key
        return self._bits

    def _apply(self, fn, *a, **k):
        self._bits = None
        return super()._apply(fn, *a, **k)

    def _load_from_state_dict(self, *a, **k):
        super()._load_from_state_dict(*a, **k)
        self.roi_host = L.host6(self.aabbs[0])        # a checkpoint may carry another box
        self._bits = None

    @torch.no_grad()
    def sampling(self, rays_o: Tensor, rays_d: Tensor, sigma_fn: Optional[Callable] = None,
                 alpha_fn: Optional[Callable] = None, near_plane: float = 0.0, far_plane: float = 1e10,
                 t_min: Optional[Tensor] = None, t_max: Optional[Tensor] = None,
                 render_step_size: float = 1e-3, early_stop_eps: float = 1e-4, alpha_thre: float = 0.0,
                 stratified: bool = False, cone_angle: float = 0.0, _return_packed: bool = False,
                 _return_mask: bool = False):
        """-> (ray_indices int64[S], t_starts[S], t_ends[S]) sorted by ray then t (+ packed_info / the survivors'
        rows in the visibility pass's outputs when asked for, see `ray_marching`)."""
        def wrap(fn):
            if fn is None:
                return None
            return lambda ts, te, ri: fn(ts[:, 0], te[:, 0], ri).reshape(-1, 1)
        self._aabb0 = self.aabbs[0]          # lets ray_marching reuse roi_host instead of copying the box back
        out = ray_marching(rays_o, rays_d, t_min=t_min, t_max=t_max, scene_aabb=self._aabb0, grid=self,
                           sigma_fn=wrap(sigma_fn), alpha_fn=wrap(alpha_fn), early_stop_eps=early_stop_eps,
                           alpha_thre=min(alpha_thre, float(self.occs.mean())) if alpha_thre > 0 else 0.0,
                           near_plane=near_plane, far_plane=far_plane, render_step_size=render_step_size,
                           stratified=stratified, cone_angle=cone_angle, _return_packed=True, _return_mask=True)
        ri, ts, te, packed, masks = out
        return (ri, ts[:, 0], te[:, 0]) + ((packed,) if _return_packed else ()) + ((masks,) if _return_mask else ())

    @torch.no_grad()
    def _update(self, step: int, occ_eval_fn: Callable, occ_thre: float = 0.01, ema_decay: float = 0.95,
                warmup_steps: int = 256, jitter: Optional[Tensor] = None):
        dev = self.occs.device
        r = self._res
        if step < warmup_steps:
            indices, n = None, self.n_cells                    # every cell, in order
        else:
            n = self.n_cells // 4
            uniform = torch.randint(self.n_cells, (n,), device=dev)
            occupied = torch.nonzero(self.binaries.flatten())[:, 0]
            if n < len(occupied):
                occupied = occupied[torch.randint(len(occupied), (n,), device=dev)]
            indices = torch.cat([uniform, occupied], dim=0)
            n = indices.shape[0]
        if jitter is None:
            jitter = torch.rand(n, 3, device=dev)
        if not self.occs.is_cuda:
            raise NotImplementedError("Only support cuda inputs.")
        # csrc/glue.cu: cell -> jittered point (1 launch), EMA-max (2, deterministic under duplicate indices),
        # mean + threshold + bool grid + bit-packed grid (3) -- instead of ~25 torch launches and a re-pack
        idx = None if indices is None else indices.contiguous()
        jitter = jitter.to(dev).contiguous().float()
        x = torch.empty(n, 3, device=dev, dtype=torch.float32)
        L.call("rsdf_occ_points", L.ptr(idx), L.ptr(jitter), n, r, self.roi_host, L.ptr(x), L.stream())
        occ = occ_eval_fn(x).reshape(-1).contiguous().float()
        snapshot = self.occs.clone()
        L.call("rsdf_occ_update", L.ptr(self.occs), L.ptr(snapshot), L.ptr(idx), L.ptr(occ), n, float(ema_decay), L.stream())
        binaries = torch.empty(self.binaries.shape, device=dev, dtype=torch.bool)
        bits = torch.empty(self.n_cells // 32, device=dev, dtype=torch.int32)
        partials = torch.empty(4 * (L.LOSS_BLOCKS + 1), device=dev, dtype=torch.float32)
        L.call("rsdf_occ_threshold", L.ptr(self.occs), self.n_cells, float(occ_thre), L.ptr(binaries.view(torch.uint8)),
               L.ptr(bits), L.ptr(partials), L.stream())
        self.binaries = binaries
        self._bits, self._bits_version = bits, (binaries.data_ptr(), binaries._version)

    @torch.no_grad()
    def update_every_n_steps(self, step: int, occ_eval_fn: Callable, occ_thre: float = 1e-2,
                             ema_decay: float = 0.95, warmup_steps: int = 256, n: int = 16):
        if not self.training:
            raise RuntimeError("You should only call this function only during training. "
                               "Please call _update() directly if you want to update the "
                               "field during inference.")
        if step % n == 0 and self.training:
            self._update(step=step, occ_eval_fn=occ_eval_fn, occ_thre=occ_thre, ema_decay=ema_decay,
                         warmup_steps=warmup_steps)

    every_n_step = update_every_n_steps   # 0.3.5 name (lib/nerfacc/grid.py:241)

    @torch.no_grad()
    def query_occ(self, samples: Tensor) -> Tensor:
        s = samples.contiguous().float()
        out = torch.empty(s.shape[0], device=s.device, dtype=torch.uint8)
        g = self.binaries.contiguous().view(torch.uint8)
        r = self._res
        L.call("rsdf_grid_query", L.ptr(s), self.roi_host, L.ptr(g), r, r, r, s.shape[0], L.ptr(out), L.stream())
        return out.bool()


class OccupancyGrid(OccGridEstimator):
    """vendored-0.3.5 name (lib/nerfacc/grid.py:113); AABB contraction only."""

    def __init__(self, roi_aabb, resolution=128, contraction_type=ContractionType.AABB):
        if contraction_type != ContractionType.AABB:
            raise NotImplementedError("contracted (unbounded) grids belong to the learned-background "
                                      "branch: out of scope")
        super().__init__(roi_aabb, resolution if isinstance(resolution, int) else int(resolution[0]))
