"""CPU restatement of the split-sum lookups and the cube-map prefilter.  TEST INFRASTRUCTURE ONLY.

  tex2d / cube_sample ... nvdiffrast `texture()` (third-party, unpinned git HEAD in the reference's
                          README.md:45-51, absent from /root/reference): restated from its published
                          behaviour.  PARITY UNPINNED.  The cube face /
                          in-face convention IS pinned: it is the reference's own cube_to_dir /
                          dir_to_side (lib/renderutils/c_src/cubemap.cu:32-60,
                          lib/pbr/utils/light_utils.py:85-92).
  diffuse / specular_bounds / specular ... lib/renderutils/c_src/cubemap.cu:17-350 and
                          lib/renderutils/ops.py:391-458.  PINNED against the reference's own
                          compiled kernels (oracle/_ref/renderutils_plugin.so -> tests/golden/cubemap_*.npz).
Differentiable torch ops where gradients are needed (autograd supplies the backward).
"""
import numpy as np
import torch

# ------------------------------------------------------------------------------------ 2-D


def tex2d(tex, uv, wrap=False):
    """tex [H,W,C], uv [n,2] (u -> width, v -> height); bilinear, clamp-to-edge, or -- wrap=True, nvdiffrast's
    default boundary mode -- uv reduced to [0,1) with the taps wrapping around both axes."""
    H, W, _ = tex.shape
    if wrap:
        uv = uv - torch.floor(uv.detach())
    tx, ty = uv[:, 0] * W - 0.5, uv[:, 1] * H - 0.5
    x0f, y0f = torch.floor(tx.detach()), torch.floor(ty.detach())
    ax, ay = (tx - x0f)[:, None], (ty - y0f)[:, None]
    if wrap:
        x0, x1 = x0f.long() % W, (x0f.long() + 1) % W
        y0, y1 = y0f.long() % H, (y0f.long() + 1) % H
    else:
        x0, x1 = x0f.long().clamp(0, W - 1), (x0f.long() + 1).clamp(0, W - 1)
        y0, y1 = y0f.long().clamp(0, H - 1), (y0f.long() + 1).clamp(0, H - 1)
    top = tex[y0, x0] * (1 - ax) + tex[y0, x1] * ax
    bot = tex[y1, x0] * (1 - ax) + tex[y1, x1] * ax
    return top * (1 - ay) + bot * ay


# ------------------------------------------------------------------------------------ cube
def dir_to_face(d):
    """d [n,3] -> face [n], u, v in [-1,1]  (inverse of cube_to_dir, cubemap.cu:32-60)."""
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    ax, ay, az = x.abs(), y.abs(), z.abs()
    fx = (ax >= ay) & (ax >= az)
    fy = (~fx) & (ay >= az)
    fz = ~(fx | fy)
    face = torch.zeros_like(x, dtype=torch.long)
    face = torch.where(fx, torch.where(x > 0, 0, 1), face)
    face = torch.where(fy, torch.where(y > 0, 2, 3), face)
    face = torch.where(fz, torch.where(z > 0, 4, 5), face)
    m = torch.where(fx, ax, torch.where(fy, ay, az))
    U = [-z, z, x, x, x, -x]
    V = [-y, -y, z, -z, -y, -y]
    u = torch.zeros_like(x)
    v = torch.zeros_like(x)
    for f in range(6):
        u = torch.where(face == f, U[f] / m, u)
        v = torch.where(face == f, V[f] / m, v)
    return face, u, v


def face_to_dir(face, fx, fy):
    one = torch.ones_like(fx)
    X = [one, -one, fx, fx, fx, -fx]
    Y = [-fy, -fy, one, -one, -fy, -fy]
    Z = [-fx, fx, fy, -fy, one, -one]
    x = torch.zeros_like(fx); y = torch.zeros_like(fx); z = torch.zeros_like(fx)
    for f in range(6):
        x = torch.where(face == f, X[f], x)
        y = torch.where(face == f, Y[f], y)
        z = torch.where(face == f, Z[f], z)
    return torch.stack([x, y, z], -1)


def _texel_on_cube(face, ix, iy, N):
    """flat texel index with off-face taps redirected to the neighbouring face; -1 at cube corners."""
    ox = (ix < 0) | (ix >= N)
    oy = (iy < 0) | (iy >= N)
    inside = face * N * N + iy.clamp(0, N - 1) * N + ix.clamp(0, N - 1)
    fx = 2.0 * ((ix.float() + 0.5) / N) - 1.0
    fy = 2.0 * ((iy.float() + 0.5) / N) - 1.0
    f2, u, v = dir_to_face(face_to_dir(face, fx, fy))
    jx = torch.floor((u + 1.0) * 0.5 * N).long().clamp(0, N - 1)
    jy = torch.floor((v + 1.0) * 0.5 * N).long().clamp(0, N - 1)
    redirected = f2 * N * N + jy * N + jx
    idx = torch.where(ox | oy, redirected, inside)
    return torch.where(ox & oy, torch.full_like(idx, -1), idx)


def cube_linear(tex, d):
    """tex [6,N,N,C], d [n,3] -> [n,C]; bilinear with cross-face taps."""
    N, C = tex.shape[1], tex.shape[-1]
    face, u, v = dir_to_face(d)
    tx, ty = (u + 1.0) * 0.5 * N - 0.5, (v + 1.0) * 0.5 * N - 0.5
    x0f, y0f = torch.floor(tx.detach()), torch.floor(ty.detach())
    ax, ay = tx - x0f, ty - y0f
    ix, iy = x0f.long(), y0f.long()
    flat = tex.reshape(-1, C)
    ws = [(1 - ax) * (1 - ay), ax * (1 - ay), (1 - ax) * ay, ax * ay]
    taps = [(ix, iy), (ix + 1, iy), (ix, iy + 1), (ix + 1, iy + 1)]
    out, wsum = 0, 0
    for w, (jx, jy) in zip(ws, taps):
        idx = _texel_on_cube(face, jx, jy, N)
        ok = (idx >= 0).to(w.dtype)
        w = w * ok
        out = out + w[:, None] * flat[idx.clamp_min(0)]
        wsum = wsum + w
    renorm = torch.where((wsum < 1.0) & (wsum > 0.0), 1.0 / wsum.detach(), torch.ones_like(wsum))
    return out * renorm[:, None]


def cube_sample(levels, d, mip_level_bias=None):
    """'linear' (one level / no bias) or 'linear-mipmap-linear' with explicit mips + per-sample bias."""
    if len(levels) == 1 or mip_level_bias is None:
        return cube_linear(levels[0], d)
    n = len(levels)
    lv = mip_level_bias.clamp(0.0, float(n - 1))
    l0 = torch.floor(lv.detach()).long().clamp(max=n - 1)
    l1 = (l0 + 1).clamp(max=n - 1)
    f = (lv - l0.to(lv.dtype))[:, None]
    per_level = torch.stack([cube_linear(t, d) for t in levels], 0)       # [n, S, C]
    idx = torch.arange(d.shape[0])
    a, b = per_level[l0, idx], per_level[l1, idx]
    return a + f * (b - a)


# ------------------------------------------------------------------------------------ prefilter
def _cube_to_dir_np(x, y, side, N):
    fx = np.float32(2.0) * ((x.astype(np.float32) + np.float32(0.5)) / np.float32(N)) - np.float32(1.0)
    fy = np.float32(2.0) * ((y.astype(np.float32) + np.float32(0.5)) / np.float32(N)) - np.float32(1.0)
    one = np.ones_like(fx)
    v = [np.stack(c, -1) for c in ([one, -fy, -fx], [-one, -fy, fx], [fx, one, fy], [fx, -one, -fy],
                                   [fx, -fy, one], [-fx, -fy, -one])]
    out = np.zeros(fx.shape + (3,), np.float32)
    for s in range(6):
        out = np.where((side == s)[..., None], v[s], out)
    l = np.sqrt((out * out).sum(-1, keepdims=True))
    return (out / l).astype(np.float32)


def texel_dirs(N):
    s, y, x = np.meshgrid(np.arange(6), np.arange(N), np.arange(N), indexing="ij")
    return _cube_to_dir_np(x, y, s, N)                    # [6,N,N,3]


def pixel_area(N):
    """cubemap.cu:17-30 -> [N,N] (same for every face)."""
    if N <= 1:
        return np.ones((N, N), np.float32)
    H = N // 2
    i = np.abs(np.arange(N) - H).astype(np.float32)
    d = (np.arctan((i + 1) / np.float32(H)) - np.arctan(i / np.float32(H))).astype(np.float32)
    return (d[:, None] * d[None, :]).astype(np.float32)   # [y, x] = dy*dx


def diffuse_weights(N):
    """W[p, x] = clamp(N_p.L_x, 0, .999) * area(x) / 3.141592  (cubemap.cu:126-136)."""
    D = texel_dirs(N).reshape(-1, 3).astype(np.float64)
    cs = np.clip(D @ D.T, 0.0, 0.999)
    area = np.tile(pixel_area(N).reshape(-1), 6).astype(np.float64)
    return cs * area[None, :] / 3.141592


def diffuse_cubemap(cubemap):
    """torch [6,N,N,3] -> [6,N,N,3], differentiable."""
    N = cubemap.shape[1]
    W = torch.from_numpy(diffuse_weights(N)).to(cubemap.dtype)
    return (W @ cubemap.reshape(-1, 3)).reshape(6, N, N, 3)


def ndf_cutoff(roughness, cutoff=0.99):
    """lib/renderutils/ops.py:428-443."""
    def ndfGGX(alphaSqr, costheta):
        costheta = np.clip(costheta, 0.0, 1.0)
        d = (costheta * alphaSqr - costheta) * costheta + 1.0
        return alphaSqr / (d * d * np.pi)
    costheta = np.cos(np.linspace(0, np.pi / 2.0, 1000000))
    D = np.cumsum(ndfGGX(roughness ** 4, costheta))
    return float(costheta[np.argmax(D >= D[..., -1] * cutoff)])


def specular_bounds(N, cutoff):
    """cubemap.cu:181-244 incl. the 16x16-tile interval culling -> float32 [6,N,N,24]."""
    dirs = texel_dirs(N)                                   # [6,N,N,3]
    V = dirs.reshape(-1, 3)                                # output texels
    TILE = 16
    nt = (N + TILE - 1) // TILE
    out = np.zeros((6 * N * N, 24), np.float32)
    c = np.float32(cutoff)
    for s in range(6):
        mnx = np.full(len(V), N - 1, np.int64); mxx = np.zeros(len(V), np.int64)
        mny = np.full(len(V), N - 1, np.int64); mxy = np.zeros(len(V), np.int64)
        for tx in range(nt):
            for ty in range(nt):
                tsx, tsy = tx * TILE, ty * TILE
                tex, tey = min((tx + 1) * TILE, N), min((ty + 1) * TILE, N)
                cx = np.array([tsx, tex, tsx, tex]); cy = np.array([tsy, tsy, tey, tey])
                L = _cube_to_dir_np(cx, cy, np.full(4, s), N)          # [4,3]
                mn, mx = L.min(0), L.max(0)
                maxdp = (np.maximum(mn[0] * V[:, 0], mx[0] * V[:, 0]) + np.maximum(mn[1] * V[:, 1], mx[1] * V[:, 1])
                         + np.maximum(mn[2] * V[:, 2], mx[2] * V[:, 2])).astype(np.float32)
                live = np.nonzero(maxdp >= c)[0]
                if len(live) == 0:
                    continue
                Lt = dirs[s, tsy:tey, tsx:tex].reshape(-1, 3)           # tile texels
                ys, xs = np.meshgrid(np.arange(tsy, tey), np.arange(tsx, tex), indexing="ij")
                ys, xs = ys.reshape(-1), xs.reshape(-1)
                # dot in the kernel's order: L.x*V.x + L.y*V.y + L.z*V.z in fp32
                dp = (Lt[None, :, 0] * V[live, None, 0] + Lt[None, :, 1] * V[live, None, 1]).astype(np.float32)
                dp = (dp + Lt[None, :, 2] * V[live, None, 2]).astype(np.float32)
                hit = dp >= c
                big = 1 << 30
                x_min = np.where(hit, xs[None, :], big).min(1); x_max = np.where(hit, xs[None, :], -1).max(1)
                y_min = np.where(hit, ys[None, :], big).min(1); y_max = np.where(hit, ys[None, :], -1).max(1)
                mnx[live] = np.minimum(mnx[live], x_min); mxx[live] = np.maximum(mxx[live], x_max)
                mny[live] = np.minimum(mny[live], y_min); mxy[live] = np.maximum(mxy[live], y_max)
        out[:, s * 4 + 0], out[:, s * 4 + 1], out[:, s * 4 + 2], out[:, s * 4 + 3] = mnx, mxx, mny, mxy
    return out.reshape(6, N, N, 24)


def specular_weights(N, roughness, cutoff, bounds):
    """W[p, x] = (L.V) D_ggx(V.H) area(x)/4 inside the lobe box of p (cubemap.cu:277-292), float64."""
    D = texel_dirs(N).reshape(-1, 3).astype(np.float64)
    dp = D @ D.T
    area = np.tile(pixel_area(N).reshape(-1), 6).astype(np.float64)
    H = D[:, None, :] + D[None, :, :]
    H = H / np.maximum(np.linalg.norm(H, axis=-1, keepdims=True), 1e-30)
    vdh = np.clip((D[:, None, :] * H).sum(-1), 0.0, 1.0)
    a2 = float(roughness) ** 4
    d = (vdh * a2 - vdh) * vdh + 1.0
    ndf = a2 / (d * d * np.pi)
    W = np.maximum(dp, 0.0) * ndf * area[None, :] / 4.0
    # lobe box membership
    b = bounds.reshape(-1, 6, 4)
    xs = np.tile(np.tile(np.arange(N), N), 6); ys = np.tile(np.repeat(np.arange(N), N), 6)
    ss = np.repeat(np.arange(6), N * N)
    bx = b[:, ss, :]                                        # [P, X, 4]
    inbox = (xs[None] >= bx[..., 0]) & (xs[None] <= bx[..., 1]) & (ys[None] >= bx[..., 2]) & (ys[None] <= bx[..., 3])
    return W * (inbox & (dp >= cutoff))


def specular_cubemap(cubemap, roughness, cutoff=0.99, bounds=None):
    """lib/renderutils/ops.py:446-458 (normalised output), differentiable; small N only (dense pairs)."""
    N = cubemap.shape[1]
    c = ndf_cutoff(roughness, cutoff)
    if bounds is None:
        bounds = specular_bounds(N, c)
    if N > 32:
        return specular_cubemap_chunked(cubemap, roughness, c, bounds)
    W = torch.from_numpy(specular_weights(N, roughness, np.float32(c), bounds)).to(cubemap.dtype)
    col = (W @ cubemap.reshape(-1, 3)).reshape(6, N, N, 3)
    wsum = W.sum(1).reshape(6, N, N, 1)
    return col / wsum


def specular_cubemap_chunked(cubemap, roughness, c, bounds, chunk=512):
    """Same filter for N = 64: output texels processed `chunk` rows at a time (the weights are constants of the
    geometry, so every chunk is a differentiable `W_chunk @ cubemap`)."""
    N = cubemap.shape[1]
    D = texel_dirs(N).reshape(-1, 3).astype(np.float64)
    area = np.tile(pixel_area(N).reshape(-1), 6).astype(np.float64)
    xs = np.tile(np.tile(np.arange(N), N), 6); ys = np.tile(np.repeat(np.arange(N), N), 6)
    ss = np.repeat(np.arange(6), N * N)
    b = bounds.reshape(-1, 6, 4)
    cube = cubemap.reshape(-1, 3)
    a2 = float(roughness) ** 4
    out = []
    for p0 in range(0, len(D), chunk):
        V = D[p0:p0 + chunk]
        dp = V @ D.T
        H = V[:, None, :] + D[None, :, :]
        H /= np.maximum(np.linalg.norm(H, axis=-1, keepdims=True), 1e-30)
        vdh = np.clip((V[:, None, :] * H).sum(-1), 0.0, 1.0)
        d = (vdh * a2 - vdh) * vdh + 1.0
        W = np.maximum(dp, 0.0) * (a2 / (d * d * np.pi)) * area[None, :] / 4.0
        bx = b[p0:p0 + chunk][:, ss, :]
        inbox = (xs[None] >= bx[..., 0]) & (xs[None] <= bx[..., 1]) & (ys[None] >= bx[..., 2]) & (ys[None] <= bx[..., 3])
        W = W * (inbox & (dp >= np.float32(c)))
        Wt = torch.from_numpy(W / W.sum(1, keepdims=True))
        out.append((Wt @ cube.double()).to(cubemap.dtype))
    return torch.cat(out).reshape(6, N, N, 3)
