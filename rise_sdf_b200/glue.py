"""Autograd wrappers of the elementwise stages in csrc/glue.cu: each replaces a chain of ~5-40 torch launches of the
reference's host code by one kernel (and its backward by one more).

    freq_encode ......... VanillaFrequency.forward                models/network_utils.py:27-33
    composite ........... background blend (+ sRGB + clamp)       models/neus.py:307-311, models/split_mixed_occ.py:416-437
    ray_loss_terms ...... masked rgb MSE + mask BCE               systems/neus.py:98-107,123-125, systems/split_occ.py:163-184
    get_rays ............ per-ray (image, pixel) -> rays          systems/split_occ.py:58-103, models/ray_utils.py:32-56
    RaySampler .......... preprocess_data('train') on the device  systems/split_occ.py:58-131 (+ dynamic ray count :159-161)
    occupancy update .... see nerfacc.OccGridEstimator._update    lib/nerfacc/grid.py:196-239
"""
import ctypes

import torch

from . import _lib as L


def _mask_array(mask, n):
    if mask is None:
        return None
    vals = [float(v) for v in mask.tolist()] if isinstance(mask, torch.Tensor) else [float(v) for v in mask]
    return (ctypes.c_float * n)(*vals[:n])


class _FreqEncode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, n_freqs, x_scale, x_offset, mask):
        x = x.contiguous().float()
        n, c = x.shape
        out = torch.empty(n, 2 * n_freqs * c, device=x.device, dtype=torch.float32)
        marr = _mask_array(mask, n_freqs)
        L.call("rsdf_freq_encode_fwd", L.ptr(x), n, c, n_freqs, float(x_scale), float(x_offset), marr, L.ptr(out), L.stream())
        ctx.save_for_backward(x)
        ctx.args = (n_freqs, float(x_scale), float(x_offset), marr)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, go):
        (x,) = ctx.saved_tensors
        n_freqs, sc, off, marr = ctx.args
        gx = torch.empty_like(x)
        L.call("rsdf_freq_encode_bwd", L.ptr(x), L.ptr(go.contiguous().float()), x.shape[0], x.shape[1], n_freqs, sc, off,
               marr, L.ptr(gx), L.stream())
        return gx, None, None, None, None


def freq_encode(x, n_freqs, x_scale=1.0, x_offset=0.0, mask=None):
    """[..., C] -> [..., 2 * n_freqs * C] in the reference's band order (sin f0 | cos f0 | sin f1 | ...)."""
    L.require_cuda(x)
    lead = x.shape[:-1]
    out = _FreqEncode.apply(x.reshape(-1, x.shape[-1]), int(n_freqs), x_scale, x_offset, mask)
    return out.view(*lead, out.shape[-1])


class _Composite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb, opacity, bg, srgb):
        rgb, opacity, bg = rgb.contiguous().float(), opacity.reshape(-1).contiguous().float(), bg.reshape(3).contiguous().float()
        n = rgb.shape[0]
        out = torch.empty_like(rgb)
        L.call("rsdf_composite_fwd", L.ptr(rgb), L.ptr(opacity), L.ptr(bg), n, int(srgb), L.ptr(out), L.stream())
        ctx.save_for_backward(rgb, opacity, bg)
        ctx.srgb = int(srgb)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, go):
        rgb, opacity, bg = ctx.saved_tensors
        n = rgb.shape[0]
        g_rgb = torch.empty_like(rgb)
        g_op = torch.empty(n, 1, device=rgb.device, dtype=torch.float32) if ctx.needs_input_grad[1] else None
        L.call("rsdf_composite_bwd", L.ptr(rgb), L.ptr(opacity), L.ptr(bg), L.ptr(go.contiguous().float()), n, ctx.srgb,
               L.ptr(g_rgb), L.ptr(g_op), L.stream())
        return g_rgb, g_op, None, None


def composite(rgb, opacity, background, srgb=False):
    """rgb [R,3] + background[3] * (1 - opacity [R,1]); srgb: followed by rgb_to_srgb(...).clamp(0, 1)."""
    if not rgb.is_cuda or rgb.shape[0] == 0:
        out = rgb + background[None, :] * (1.0 - opacity)
        if srgb:
            from .light import rgb_to_srgb
            out = rgb_to_srgb(out).clamp(0, 1)
        return out
    return _Composite.apply(rgb, opacity, background, srgb)


class _RayLossTerms(torch.autograd.Function):
    @staticmethod
    def forward(ctx, full, opacity, target, fg):
        full, opacity = full.contiguous().float(), opacity.reshape(-1).contiguous().float()
        target, fg = target.contiguous().float(), fg.reshape(-1).contiguous().float()
        n = full.shape[0]
        sums = torch.empty(4, device=full.device, dtype=torch.float32)
        partials = torch.empty(4 * (L.LOSS_BLOCKS + 1), device=full.device, dtype=torch.float32)
        L.call("rsdf_neus_loss_fwd", L.ptr(full), L.ptr(opacity), L.ptr(target), L.ptr(fg), n, L.ptr(sums), L.ptr(partials),
               L.stream())
        ctx.save_for_backward(full, opacity, target, fg)
        ctx.op_shape = None
        return sums

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_sums):
        full, opacity, target, fg = ctx.saved_tensors
        n = full.shape[0]
        g_full = torch.empty_like(full)
        g_op = torch.empty(n, device=full.device, dtype=torch.float32)
        L.call("rsdf_neus_loss_bwd", L.ptr(full), L.ptr(opacity), L.ptr(target), L.ptr(fg), L.ptr(g_sums.contiguous().float()),
               n, L.ptr(g_full), L.ptr(g_op), L.stream())
        return g_full, g_op, None, None


def ray_loss_terms(comp_rgb_full, opacity, target_rgb, fg_mask):
    """-> (rgb_mse over rays with opacity > 0, mask BCE) exactly as systems/neus.py:103,123-125 define them:
    F.mse_loss(full[valid], rgb[valid]) and binary_cross_entropy(clamp(opacity, 1e-3, 1 - 1e-3), fg_mask)."""
    n = comp_rgb_full.shape[0]
    op = opacity.reshape(-1)
    sums = _RayLossTerms.apply(comp_rgb_full, op, target_rgb, fg_mask)
    return sums[0] / (sums[1] * 3.0), sums[2] / max(n, 1)


@torch.no_grad()
def get_rays(directions, c2w, index, px, py):
    """directions [H,W,3], c2w [n_img,3,4], per-ray index / px / py int64 [n] -> rays [n,6] = (origin, unit direction)."""
    L.require_cuda(directions, c2w, px, py)
    n = px.shape[0]
    H, W = directions.shape[:2]
    rays = torch.empty(n, 6, device=px.device, dtype=torch.float32)
    L.call("rsdf_get_rays", L.ptr(directions.contiguous().float()), L.ptr(c2w.contiguous().float()),
           L.ptr(None if index is None else index.contiguous().long()), L.ptr(px.contiguous().long()),
           L.ptr(py.contiguous().long()), n, W, H, c2w.shape[0], L.ptr(rays), L.stream())
    return rays


class RaySampler:
    """`preprocess_data(batch, 'train')` of systems/split_occ.py:58-131 with the dataset resident on the device (as the
    reference keeps it: datasets/tensoir_synthetic.py holds all images and `directions` on the GPU): per step draw
    (image, x, y) per ray, build the rays with ONE kernel, gather the targets, draw the background colour and apply
    the mask composite of :119-126.  Also owns the dynamic ray count of :159-161."""

    def __init__(self, directions, c2w, images, fg_masks, train_num_rays, max_train_num_rays=None, num_samples_per_ray=1024,
                 dynamic=False, background="random", apply_mask=True, srgb_background=True, generator=None):
        self.directions, self.c2w, self.images, self.fg_masks = directions, c2w, images, fg_masks
        self.train_num_rays = int(train_num_rays)
        self.max_train_num_rays = int(max_train_num_rays or train_num_rays)
        self.train_num_samples = int(train_num_rays) * int(num_samples_per_ray)
        self.dynamic, self.background, self.apply_mask, self.srgb_background = dynamic, background, apply_mask, srgb_background
        self.generator = generator

    @torch.no_grad()
    def sample(self):
        dev = self.images.device
        n, g = self.train_num_rays, self.generator
        H, W = self.directions.shape[:2]
        index = torch.randint(0, self.images.shape[0], (n,), device=dev, generator=g)
        x = torch.randint(0, W, (n,), device=dev, generator=g)
        y = torch.randint(0, H, (n,), device=dev, generator=g)
        rays = get_rays(self.directions, self.c2w, index, x, y)
        rgb = self.images[index, y, x].view(n, -1).float()
        fg = self.fg_masks[index, y, x].view(-1).float()
        if self.background == "white":
            bg = torch.ones(3, device=dev)
        elif self.background == "black":
            bg = torch.zeros(3, device=dev)
        else:
            bg = torch.rand(3, device=dev, generator=g)
        if self.apply_mask:
            b = bg
            if self.srgb_background:
                from .light import rgb_to_srgb
                b = rgb_to_srgb(bg)
            rgb = rgb * fg[:, None] + b * (1 - fg[:, None])
        return rays, rgb, fg, bg

    def update_ray_count(self, num_samples):
        """systems/split_occ.py:159-161 (`num_samples`: the step's sample count, already on the host after the march)."""
        if self.dynamic and num_samples > 0:
            want = int(self.train_num_rays * (self.train_num_samples / num_samples))
            self.train_num_rays = min(int(self.train_num_rays * 0.9 + want * 0.1), self.max_train_num_rays)
        return self.train_num_rays
