"""Seeded synthetic inputs for the hot path: cameras, rays, targets,
analytic occupancy grids, HDR env maps and a split-sum BSDF LUT.  No dataset is available
offline, so bench.py, smoke() and the tests all draw from here.  Everything is generated on
the host with explicit generators so CPU oracle and CUDA path see identical bits.

Camera / ray conventions follow models/ray_utils.py:9-56 (pixel centres +0.5, OpenGL camera,
rays_d = directions @ c2w[:3,:3]^T) and systems/split_occ.py:63-103 (per-ray random image
index and pixel, rays = cat[o, normalize(d)]).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

W = H = 800
FOCAL = 0.5 * 800 / math.tan(0.5 * 0.6911112)  # 1111.11, the Blender/TensoIR camera_angle_x
RADIUS_CAM = 4.0


def ray_directions(w=W, h=H, focal=FOCAL):
    """models/ray_utils.py:9-29 (openGL_camera=True, pixel centres)."""
    i, j = np.meshgrid(np.arange(w, dtype=np.float32) + 0.5, np.arange(h, dtype=np.float32) + 0.5,
                       indexing="xy")
    i, j = torch.from_numpy(i), torch.from_numpy(j)
    return torch.stack([(i - w / 2) / focal, -(j - h / 2) / focal, -torch.ones_like(i)], -1)


def camera_poses(n=100, radius=RADIUS_CAM):
    """n look-at-origin poses on the upper hemisphere, Fibonacci spacing, OpenGL c2w [n,3,4]."""
    k = torch.arange(n, dtype=torch.float64) + 0.5
    z = k / n * 0.9 + 0.05                      # elevation in (0.05, 0.95)
    phi = k * math.pi * (3.0 - math.sqrt(5.0))
    r = torch.sqrt(1 - z * z)
    pos = torch.stack([r * torch.cos(phi), r * torch.sin(phi), z], -1) * radius
    fwd = -pos / pos.norm(dim=-1, keepdim=True)         # camera looks along -z_cam
    up = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64).expand_as(fwd)
    right = torch.cross(fwd, up, dim=-1)
    right = right / right.norm(dim=-1, keepdim=True)
    true_up = torch.cross(right, fwd, dim=-1)
    c2w = torch.stack([right, true_up, -fwd, pos], -1)  # columns: x, y, z, t
    return c2w.float()


def get_rays(directions, c2w):
    """models/ray_utils.py:32-56 for directions [N,3], c2w [N,3,4] (or [1,3,4])."""
    rays_d = (directions[:, None, :] * c2w[:, :3, :3]).sum(-1)
    rays_o = c2w[:, :, 3].expand(rays_d.shape)
    return rays_o, rays_d


def training_rays(n_rays, seed=42, rank=0, poses=None, directions=None):
    """systems/split_occ.py:63-103: index ~ randint(n_images), x,y ~ randint(800) per ray.
    Returns rays [n,6] float32 (o, unit d), plus synthetic target rgb [n,3] and fg mask [n]."""
    g = torch.Generator().manual_seed(seed + rank)
    poses = camera_poses() if poses is None else poses
    directions = ray_directions() if directions is None else directions
    index = torch.randint(0, poses.shape[0], (n_rays,), generator=g)
    x = torch.randint(0, W, (n_rays,), generator=g)
    y = torch.randint(0, H, (n_rays,), generator=g)
    d = directions[y, x]
    o, dd = get_rays(d, poses[index])
    rays = torch.cat([o, F.normalize(dd, p=2, dim=-1)], -1).contiguous()
    # synthetic supervision: a unit-sphere-ish object of radius 0.8 in front of random bg
    b = (rays[:, :3] * rays[:, 3:]).sum(-1)
    disc = b * b - ((rays[:, :3] ** 2).sum(-1) - 0.8 ** 2)
    fg = (disc > 0).float()
    rgb = torch.rand(n_rays, 3, generator=g) * 0.5 + 0.25
    bg = torch.rand(3, generator=g)
    rgb = rgb * fg[:, None] + bg * (1 - fg[:, None])
    return rays, rgb.contiguous(), fg, bg


def frame_rays(pose_index=0, poses=None, directions=None, w=W, h=H):
    """All w*h pixels of one pose, row-major (systems/split_occ.py test branch)."""
    poses = camera_poses() if poses is None else poses
    directions = ray_directions(w, h, FOCAL * w / W) if directions is None else directions
    d = directions.reshape(-1, 3)
    o, dd = get_rays(d, poses[pose_index:pose_index + 1])
    return torch.cat([o.expand_as(dd), F.normalize(dd, p=2, dim=-1)], -1).contiguous()


# ---- occupancy grids (SURVEY §8d "Occupancy grid" row, variant A) ---------------------------
def analytic_grid(kind="ball", res=128, radius=1.5):
    c = (torch.arange(res, dtype=torch.float32) + 0.5) / res * (2 * radius) - radius
    X, Y, Z = torch.meshgrid(c, c, c, indexing="ij")
    r = torch.sqrt(X * X + Y * Y + Z * Z)
    if kind == "ball":
        return r < 1.1
    if kind == "shell":
        return (r > 0.7) & (r < 0.8)
    if kind == "ones":
        return torch.ones(res, res, res, dtype=torch.bool)
    if kind == "zeros":
        return torch.zeros(res, res, res, dtype=torch.bool)
    if kind == "voxel":
        g = torch.zeros(res, res, res, dtype=torch.bool)
        g[res // 2, res // 2 + 3, res // 2 - 5] = True
        return g
    if kind == "random":
        gen = torch.Generator().manual_seed(7)
        return torch.rand(res, res, res, generator=gen) < 0.3
    raise ValueError(kind)


# ---- env maps / LUT (used by the split-sum rows) ---------------------------------------------
def env_latlong(kind="bridge", w=512, h=256, seed=1):
    """Two synthetic HDR lat-long maps [h,w,3] float32 with the statistics of the bridge/city
    maps the reference relights with (sky gradient + sun; ambient + bright rectangles)."""
    g = torch.Generator().manual_seed(seed)
    v = (torch.arange(h, dtype=torch.float32) + 0.5) / h
    u = (torch.arange(w, dtype=torch.float32) + 0.5) / w
    V, U = torch.meshgrid(v, u, indexing="ij")
    if kind == "bridge":
        sky = 0.3 + (2.0 - 0.3) * (1 - V * 2).clamp(0, 1)
        img = torch.where(V < 0.5, sky, torch.full_like(sky, 0.05))
        ang = torch.sqrt(((U - 0.3) * 2 * math.pi) ** 2 + ((V - 0.25) * math.pi) ** 2)
        img = img + 50.0 * torch.exp(-0.5 * (ang / math.radians(3.0)) ** 2)
        img = img[..., None] * torch.tensor([1.0, 0.95, 0.9])
    else:
        img = torch.full((h, w, 3), 0.1)
        for _ in range(200):
            x0 = int(torch.randint(0, w - 8, (1,), generator=g))
            y0 = int(torch.randint(0, h - 8, (1,), generator=g))
            ww = int(torch.randint(2, 24, (1,), generator=g))
            hh = int(torch.randint(2, 12, (1,), generator=g))
            val = torch.rand(3, generator=g) * 19 + 1
            img[y0:y0 + hh, x0:x0 + ww] = val
    return img.contiguous()


def bsdf_lut(res=256, n_samples=256):
    """Synthetic split-sum (A,B) table [1,res,res,2]: row = roughness, col = NoV
    (layout of np.fromfile(bsdf_256_256.bin).reshape(1,256,256,2), models/texture.py:285).
    Karis' GGX importance-sampled integration with a Hammersley set."""
    nov = ((torch.arange(res, dtype=torch.float64) + 0.5) / res)[None, :, None]
    rough = ((torch.arange(res, dtype=torch.float64) + 0.5) / res)[:, None, None]
    i = torch.arange(n_samples, dtype=torch.float64)
    bits = torch.zeros(n_samples, dtype=torch.float64)
    ii = torch.arange(n_samples)
    f = 0.5
    while int(ii.max()) > 0:
        bits += f * (ii % 2).double()
        ii = ii // 2
        f *= 0.5
    xi1, xi2 = ((i + 0.5) / n_samples)[None, None, :], bits[None, None, :]
    a = rough * rough
    phi = 2 * math.pi * xi1
    cos_t = torch.sqrt((1 - xi2) / (1 + (a * a - 1) * xi2))
    sin_t = torch.sqrt(1 - cos_t * cos_t)
    hx, hz = sin_t * torch.cos(phi), cos_t
    vx, vz = torch.sqrt(1 - nov * nov), nov
    voh = vx * hx + vz * hz
    lz = 2 * voh * hz - vz
    nol, noh, voh = lz.clamp(0, 1), hz.clamp(0, 1), voh.clamp(0, 1)
    k = a / 2
    gv = nov / (nov * (1 - k) + k)
    gl = nol / (nol * (1 - k) + k)
    gvis = gv * gl * voh / (noh * nov).clamp_min(1e-8)
    fc = (1 - voh) ** 5
    ok = (nol > 0).double()
    A = ((1 - fc) * gvis * ok).mean(-1)
    B = (fc * gvis * ok).mean(-1)
    return torch.stack([A, B], -1).float().view(1, res, res, 2).contiguous()
