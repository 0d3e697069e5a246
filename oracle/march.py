"""ctypes front-end of oracle/march_oracle.c.  TEST INFRASTRUCTURE ONLY.

Host-side logic restates lib/nerfacc/ray_marching.py:131-222 (t_min/t_max selection, near/far
clamps, two-pass count + cumsum + fill, packed_info = stack([cum - n, n])) and
lib/nerfacc/vol_rendering.py:453-520 (render_visibility).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmarch_oracle.so")
_SRC = os.path.join(_HERE, "march_oracle.c")
_lib = None


def build(force=False):
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", _SO, _SRC, "-lm"]
        )
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_ray_marching.restype = ctypes.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def ray_aabb_intersect(rays_o, rays_d, aabb):
    o, d, bb = _f32(rays_o), _f32(rays_d), _f32(aabb)
    n = o.shape[0]
    t_min = np.empty(n, np.float32)
    t_max = np.empty(n, np.float32)
    lib().oracle_ray_aabb_intersect(n, _p(o), _p(d), _p(bb), _p(t_min), _p(t_max))
    return t_min, t_max


def ray_marching_raw(rays_o, rays_d, t_min, t_max, roi, grid_binary, step_size, cone_angle=0.0):
    """== _C.ray_marching (ray_marching.cu:194-289): returns packed_info, ray_indices, t_starts, t_ends."""
    o, d = _f32(rays_o), _f32(rays_d)
    tmin, tmax, roi = _f32(t_min), _f32(t_max), _f32(roi)
    g = np.ascontiguousarray(np.asarray(grid_binary).astype(np.uint8))
    res = np.asarray(g.shape, np.int32)
    n = o.shape[0]
    num = np.zeros(n, np.int32)
    L = lib()
    L.oracle_ray_marching(n, _p(o), _p(d), _p(tmin), _p(tmax), _p(roi), _p(res), _p(g),
                          ctypes.c_float(step_size), ctypes.c_float(cone_angle),
                          None, _p(num), None, None, None)
    cum = np.cumsum(num, dtype=np.int32)
    packed = np.stack([cum - num, num], 1).astype(np.int32)
    total = int(cum[-1]) if n > 0 else 0
    ri = np.empty(total, np.int64)
    ts = np.empty(total, np.float32)
    te = np.empty(total, np.float32)
    L.oracle_ray_marching(n, _p(o), _p(d), _p(tmin), _p(tmax), _p(roi), _p(res), _p(g),
                          ctypes.c_float(step_size), ctypes.c_float(cone_angle),
                          _p(packed), None, _p(ri), _p(ts), _p(te))
    return packed, ri, ts, te


def grid_query(samples, roi, grid):
    s, roi = _f32(samples), _f32(roi)
    g = np.ascontiguousarray(np.asarray(grid).astype(np.uint8))
    res = np.asarray(g.shape, np.int32)
    out = np.empty(s.shape[0], np.uint8)
    lib().oracle_grid_query(s.shape[0], _p(s), _p(roi), _p(res), _p(g), _p(out))
    return out.astype(bool)


def weight_from_alpha(packed_info, alphas):
    pk = np.ascontiguousarray(packed_info, dtype=np.int32)
    a = _f32(alphas)
    w = np.empty_like(a)
    T = np.empty_like(a)
    lib().oracle_weight_from_alpha_forward(pk.shape[0], _p(pk), _p(a), _p(w), _p(T))
    return w, T


def weight_from_alpha_backward(packed_info, alphas, weights, grad_weights):
    pk = np.ascontiguousarray(packed_info, dtype=np.int32)
    a, w, gw = _f32(alphas), _f32(weights), _f32(grad_weights)
    ga = np.zeros_like(a)
    lib().oracle_weight_from_alpha_backward(pk.shape[0], _p(pk), _p(a), _p(w), _p(gw), _p(ga))
    return ga


def transmittance_from_alpha_backward(packed_info, alphas, trans, trans_grad):
    pk = np.ascontiguousarray(packed_info, dtype=np.int32)
    a, T, gT = _f32(alphas), _f32(trans), _f32(trans_grad)
    ga = np.zeros_like(a)
    lib().oracle_transmittance_from_alpha_backward(pk.shape[0], _p(pk), _p(a), _p(T), _p(gT), _p(ga))
    return ga


def pack_info(ray_indices, n_rays):
    """lib/nerfacc/pack.py: ray_indices (sorted) -> packed_info[n_rays, 2]."""
    num = np.bincount(np.asarray(ray_indices, np.int64), minlength=n_rays).astype(np.int32)
    cum = np.cumsum(num, dtype=np.int32)
    return np.stack([cum - num, num], 1).astype(np.int32)


def ray_marching(rays_o, rays_d, t_min=None, t_max=None, scene_aabb=None, grid_roi=None,
                 grid_binary=None, alpha_fn=None, early_stop_eps=1e-4, alpha_thre=0.0,
                 near_plane=None, far_plane=None, render_step_size=1e-3, jitter=None,
                 cone_angle=0.0):
    """lib/nerfacc/ray_marching.py:14-222 on numpy arrays.  `jitter` is the explicit
    per-ray U[0,1) draw that `stratified=True` would make (ray_marching.py:157-158)."""
    o, d = _f32(rays_o), _f32(rays_d)
    if t_min is None or t_max is None:
        if scene_aabb is not None:
            t_min, t_max = ray_aabb_intersect(o, d, scene_aabb)
        else:
            t_min = np.zeros(o.shape[0], np.float32)
            t_max = np.full(o.shape[0], 1e10, np.float32)
    t_min, t_max = _f32(t_min), _f32(t_max)
    if near_plane is not None:
        t_min = np.maximum(t_min, np.float32(near_plane))
    if far_plane is not None:
        t_max = np.minimum(t_max, np.float32(far_plane))
    if jitter is not None:
        t_min = (t_min + _f32(jitter) * np.float32(render_step_size)).astype(np.float32)
    if grid_binary is None:
        grid_roi = np.array([-1e10] * 3 + [1e10] * 3, np.float32)
        grid_binary = np.ones((1, 1, 1), bool)
    packed, ri, ts, te = ray_marching_raw(o, d, t_min, t_max, grid_roi, grid_binary,
                                          render_step_size, cone_angle)
    if alpha_fn is not None:
        alphas = _f32(alpha_fn(ts, te, ri))
        _, T = weight_from_alpha(packed, alphas)
        vis = T >= np.float32(early_stop_eps)
        if alpha_thre > 0:
            vis &= alphas >= np.float32(alpha_thre)
        ri, ts, te = ri[vis], ts[vis], te[vis]
    return ri, ts, te
