"""Host-side mirror of the environment light on the render path:
    EnvironmentLightMipCube ... lib/pbr/light.py:126-206 (learnable 6x512x512x3 cube base,
                                build_mips / get_mip / eval_mip)
    cubemap_mip ................ lib/pbr/utils/light_utils.py:94-109 (2x2 avg-pool mip + its backward)
    blender_latlong_to_cubemap . lib/pbr/utils/light_utils.py:126-139
    rgb_to_srgb ................ lib/pbr/utils/nvdiffrecmc_util.py:95-103
Lookups and prefilters run on librsdf_b200.so through the nvdiffrast / renderutils shims.
MC sampling helpers (pdf / sample / update_pdf) are out of scope (unused by split-sum).
"""
import numpy as np
import torch
import torch.nn as nn

from . import nvdiffrast as dr
from . import renderutils as ru
from .network_utils import Config


def _rgb_to_srgb(f):
    return torch.where(f <= 0.0031308, f * 12.92, torch.pow(torch.clamp(f, 0.0031308), 1.0 / 2.4) * 1.055 - 0.055)


def rgb_to_srgb(f):
    assert f.shape[-1] == 3 or f.shape[-1] == 4
    return torch.cat((_rgb_to_srgb(f[..., 0:3]), f[..., 3:4]), dim=-1) if f.shape[-1] == 4 else _rgb_to_srgb(f)


def avg_pool_nhwc(x, size):
    y = x.permute(0, 3, 1, 2)
    y = torch.nn.functional.avg_pool2d(y, size)
    return y.permute(0, 2, 3, 1).contiguous()


def safe_normalize(x, eps=1e-20):
    return x / torch.sqrt(torch.clamp(torch.sum(x * x, -1, keepdim=True), min=eps))


def cube_to_dir(s, x, y):
    one = torch.ones_like(x)
    if s == 0:
        rx, ry, rz = one, -y, -x
    elif s == 1:
        rx, ry, rz = -one, -y, x
    elif s == 2:
        rx, ry, rz = x, one, y
    elif s == 3:
        rx, ry, rz = x, -one, -y
    elif s == 4:
        rx, ry, rz = x, -y, one
    else:
        rx, ry, rz = -x, -y, -one
    return torch.stack((rx, ry, rz), dim=-1)


_FACE_DIRS = {}


def _face_dirs(res, device):
    """Unit directions of the texel centres of the six faces, [6, res, res, 3]: constants of (res, device), cached
    (the reference rebuilds them in every cubemap_mip backward, lib/pbr/utils/light_utils.py:99-109: ~56 launches)."""
    key = (int(res), str(device))
    if key not in _FACE_DIRS:
        lin = torch.linspace(-1.0 + 1.0 / res, 1.0 - 1.0 / res, res, device=device)
        gy, gx = torch.meshgrid(lin, lin, indexing="ij")
        _FACE_DIRS[key] = torch.stack([safe_normalize(cube_to_dir(s, gx, gy)) for s in range(6)]).contiguous()
    return _FACE_DIRS[key]


class cubemap_mip(torch.autograd.Function):
    """2x2 average-pool mip; the backward is the reference's own (a cube-filtered upsample of
    0.25*dout, lib/pbr/utils/light_utils.py:99-109), not the adjoint of avg-pool.  The six per-face lookups of the
    reference are one lookup over all 6*res^2 directions (same kernel, same per-texel arithmetic)."""

    @staticmethod
    def forward(ctx, cubemap):
        return avg_pool_nhwc(cubemap, (2, 2))

    @staticmethod
    def backward(ctx, dout):
        res = dout.shape[1] * 2
        dirs = _face_dirs(res, dout.device)
        out = dr.texture(dout[None, ...] * 0.25, dirs.view(1, 6 * res, res, 3), filter_mode="linear",
                         boundary_mode="cube")
        return out.view(6, res, res, dout.shape[-1])


def blender_latlong_to_cubemap(latlong_map, res):
    cubemap = torch.zeros(6, res[0], res[1], latlong_map.shape[-1], dtype=torch.float32, device=latlong_map.device)
    for s, v in enumerate(_face_dirs(res[0], latlong_map.device)):
        tu = torch.atan2(-v[..., 1:2], v[..., 0:1]) / (2 * np.pi) + 0.5
        tv = torch.acos(torch.clamp(v[..., 2:3], min=-1, max=1)) / np.pi
        texcoord = torch.cat((tu, tv), dim=-1)
        cubemap[s, ...] = dr.texture(latlong_map[None, ...], texcoord[None, ...], filter_mode="linear")[0]
    return cubemap


class EnvironmentLightMipCube(nn.Module):
    LIGHT_MIN_RES = 16
    MIN_ROUGHNESS = 0.08
    MAX_ROUGHNESS = 0.5

    def __init__(self, config, latlong=None, device="cuda"):
        """config.envlight_config: scale, bias, base_res (hdr_filepath is replaced by an in-memory
        `latlong` [H,W,3] tensor: there are no HDR files offline)."""
        super().__init__()
        self.config = config = Config(config)
        ec = config.envlight_config
        if latlong is None:
            base = torch.rand(6, ec.base_res, ec.base_res, 3, dtype=torch.float32, device=device) * ec.scale + ec.bias
        else:
            img = latlong.to(device).float()
            if ec.get("clamp", False):
                img = img.clamp(0, 1)
            base = blender_latlong_to_cubemap(img, [512, 512])
        self.register_parameter("base", nn.Parameter(base))
        self.specular, self.diffuse = None, None

    def relight(self, latlong):
        base = blender_latlong_to_cubemap(latlong.to(self.base.device).float(), [512, 512])
        self.register_parameter("base", nn.Parameter(base))

    def build_mips(self, cutoff=0.99):
        self.specular = [self.base]
        while self.specular[-1].shape[1] > self.LIGHT_MIN_RES:
            self.specular += [cubemap_mip.apply(self.specular[-1])]
        self.diffuse = ru.diffuse_cubemap(self.specular[-1])
        for idx in range(len(self.specular) - 1):
            roughness = (idx / (len(self.specular) - 2)) * (self.MAX_ROUGHNESS - self.MIN_ROUGHNESS) + self.MIN_ROUGHNESS
            self.specular[idx] = ru.specular_cubemap(self.specular[idx], roughness, cutoff)
        self.specular[-1] = ru.specular_cubemap(self.specular[-1], 1.0, cutoff)

    def get_mip(self, roughness):
        n = len(self.specular)
        return torch.where(
            roughness < self.MAX_ROUGHNESS,
            (torch.clamp(roughness, self.MIN_ROUGHNESS, self.MAX_ROUGHNESS) - self.MIN_ROUGHNESS)
            / (self.MAX_ROUGHNESS - self.MIN_ROUGHNESS) * (n - 2),
            (torch.clamp(roughness, self.MAX_ROUGHNESS, 1.0) - self.MAX_ROUGHNESS) / (1.0 - self.MAX_ROUGHNESS) + n - 2)

    def eval_mip(self, directions, specular=False, roughness=None):
        pn, bn = directions.shape[0], 1
        if specular:
            assert roughness is not None
            miplevel = self.get_mip(roughness)
            miplevel = miplevel.reshape(1, pn // bn, bn, miplevel.shape[-1])
            light = dr.texture(self.specular[0][None, ...],
                               directions.reshape(1, pn // bn, bn, directions.shape[-1]).contiguous(),
                               mip=list(m[None, ...] for m in self.specular[1:]), mip_level_bias=miplevel[..., 0],
                               filter_mode="linear-mipmap-linear", boundary_mode="cube")
        else:
            light = dr.texture(self.diffuse[None, ...],
                               directions.reshape(1, pn // bn, bn, directions.shape[-1]).contiguous(),
                               filter_mode="linear", boundary_mode="cube")
        return light.reshape(light.shape[1], -1)

    def parameters(self, recurse=True):
        return [self.base]
