// K5a -- filtered texture lookups for split-sum shading: the three nvdiffrast `texture()` modes
// RISE-SDF uses (models/texture.py:340, lib/pbr/light.py:194-206):
//   * 2-D, linear, clamp                (FG / BSDF LUT)
//   * cube, linear                      (16^2 diffuse irradiance map)
//   * cube, linear-mipmap-linear with an explicit mip stack and a per-sample mip_level_bias
// One thread per sample; texels are read through the read-only path (the whole pyramid is
// L2-resident).  Cube face/in-face conventions are pinned by the reference's own cube_to_dir /
// dir_to_side (lib/renderutils/c_src/cubemap.cu:32-60).  Bilinear taps that fall off a face edge
// are redirected to the adjacent face by re-projecting the tap centre; at cube corners the
// doubly-out-of-range tap is dropped and the weights renormalised.
#include "common.cuh"

namespace {

struct CubeTap {
    int idx[4];     // flat texel index (face*N*N + y*N + x) or -1 when dropped
    float w[4];
};

// (face, in-face coords in [-1,1]) of a direction; dir need not be normalised
__device__ __forceinline__ void dir_to_face(float x, float y, float z, int &face, float &u, float &v) {
    const float ax = fabsf(x), ay = fabsf(y), az = fabsf(z);
    if (ax >= ay && ax >= az) {
        const float im = 1.0f / ax;
        if (x > 0) { face = 0; u = -z * im; v = -y * im; } else { face = 1; u = z * im; v = -y * im; }
    } else if (ay >= az) {
        const float im = 1.0f / ay;
        if (y > 0) { face = 2; u = x * im; v = z * im; } else { face = 3; u = x * im; v = -z * im; }
    } else {
        const float im = 1.0f / az;
        if (z > 0) { face = 4; u = x * im; v = -y * im; } else { face = 5; u = -x * im; v = -y * im; }
    }
}

// un-normalised direction of in-face coords (fx, fy) on `face` (cubemap.cu:32-46)
__device__ __forceinline__ void face_to_dir(int face, float fx, float fy, float &x, float &y, float &z) {
    switch (face) {
        case 0: x = 1.f; y = -fy; z = -fx; break;
        case 1: x = -1.f; y = -fy; z = fx; break;
        case 2: x = fx; y = 1.f; z = fy; break;
        case 3: x = fx; y = -1.f; z = -fy; break;
        case 4: x = fx; y = -fy; z = 1.f; break;
        default: x = -fx; y = -fy; z = -1.f; break;
    }
}

__device__ __forceinline__ int texel_on_cube(int face, int ix, int iy, int N) {
    const bool ox = ix < 0 || ix >= N, oy = iy < 0 || iy >= N;
    if (!ox && !oy) return face * N * N + iy * N + ix;
    if (ox && oy) return -1;                                  // cube corner: tap dropped
    const float fx = 2.0f * (((float)ix + 0.5f) / (float)N) - 1.0f;
    const float fy = 2.0f * (((float)iy + 0.5f) / (float)N) - 1.0f;
    float x, y, z, u, v;
    int f2;
    face_to_dir(face, fx, fy, x, y, z);
    dir_to_face(x, y, z, f2, u, v);
    int jx = (int)floorf((u + 1.0f) * 0.5f * (float)N), jy = (int)floorf((v + 1.0f) * 0.5f * (float)N);
    jx = min(max(jx, 0), N - 1); jy = min(max(jy, 0), N - 1);
    return f2 * N * N + jy * N + jx;
}

__device__ __forceinline__ CubeTap cube_taps(float dx, float dy, float dz, int N) {
    int face; float u, v;
    dir_to_face(dx, dy, dz, face, u, v);
    const float tx = (u + 1.0f) * 0.5f * (float)N - 0.5f, ty = (v + 1.0f) * 0.5f * (float)N - 0.5f;
    const float fx0 = floorf(tx), fy0 = floorf(ty);
    const int ix = (int)fx0, iy = (int)fy0;
    const float ax = tx - fx0, ay = ty - fy0;
    CubeTap t;
    t.idx[0] = texel_on_cube(face, ix, iy, N);         t.w[0] = (1.f - ax) * (1.f - ay);
    t.idx[1] = texel_on_cube(face, ix + 1, iy, N);     t.w[1] = ax * (1.f - ay);
    t.idx[2] = texel_on_cube(face, ix, iy + 1, N);     t.w[2] = (1.f - ax) * ay;
    t.idx[3] = texel_on_cube(face, ix + 1, iy + 1, N); t.w[3] = ax * ay;
    float ws = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) { if (t.idx[k] < 0) t.w[k] = 0.f; ws += t.w[k]; }
    if (ws < 1.0f && ws > 0.0f) {
        const float inv = 1.0f / ws;
#pragma unroll
        for (int k = 0; k < 4; ++k) t.w[k] *= inv;
    }
    return t;
}

template <int C>
__device__ __forceinline__ void cube_fetch(const float *__restrict__ tex, int N, float dx, float dy, float dz,
                                           float *out) {
    const CubeTap t = cube_taps(dx, dy, dz, N);
#pragma unroll
    for (int c = 0; c < C; ++c) out[c] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (t.idx[k] >= 0) {
#pragma unroll
            for (int c = 0; c < C; ++c) out[c] = fmaf(t.w[k], __ldg(tex + (size_t)t.idx[k] * C + c), out[c]);
        }
    }
}

struct MipStack {
    const float *level[8];
    float *grad[8];
    int res[8];
    int n_levels;
};

// cube, linear (n_levels == 1) or linear-mipmap-linear (level = clamp(bias, 0, n_levels-1))
__global__ void cube_sample_fwd_kernel(const MipStack m, const float *__restrict__ dirs,
                                       const float *__restrict__ bias, int n, float *__restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const float dx = dirs[3 * s], dy = dirs[3 * s + 1], dz = dirs[3 * s + 2];
    float r[3];
    if (m.n_levels == 1 || bias == nullptr) {
        cube_fetch<3>(m.level[0], m.res[0], dx, dy, dz, r);
    } else {
        const float lv = fminf(fmaxf(bias[s], 0.0f), (float)(m.n_levels - 1));
        const int l0 = min((int)floorf(lv), m.n_levels - 1), l1 = min(l0 + 1, m.n_levels - 1);
        const float f = lv - (float)l0;
        float a[3], b[3];
        cube_fetch<3>(m.level[l0], m.res[l0], dx, dy, dz, a);
        if (f > 0.0f && l1 != l0) {
            cube_fetch<3>(m.level[l1], m.res[l1], dx, dy, dz, b);
#pragma unroll
            for (int c = 0; c < 3; ++c) a[c] = fmaf(f, b[c] - a[c], a[c]);
        }
        r[0] = a[0]; r[1] = a[1]; r[2] = a[2];
    }
    out[3 * s] = r[0]; out[3 * s + 1] = r[1]; out[3 * s + 2] = r[2];
}

// backward wrt texels (atomic scatter into each level's grad) and wrt mip_level_bias
__global__ void cube_sample_bwd_kernel(const MipStack m, const float *__restrict__ dirs,
                                       const float *__restrict__ bias, const float *__restrict__ go, int n,
                                       float *__restrict__ g_bias, float *__restrict__ g_dirs) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const float dx = dirs[3 * s], dy = dirs[3 * s + 1], dz = dirs[3 * s + 2];
    const float g[3] = {go[3 * s], go[3 * s + 1], go[3 * s + 2]};
    float d_u = 0.f, d_v = 0.f;   // d<g,out>/d(in-face u, v), summed over the two mip levels
    int l0 = 0, l1 = 0;
    float f = 0.f;
    bool live = false;
    if (m.n_levels > 1 && bias != nullptr) {
        const float raw = bias[s];
        const float lv = fminf(fmaxf(raw, 0.0f), (float)(m.n_levels - 1));
        l0 = min((int)floorf(lv), m.n_levels - 1); l1 = min(l0 + 1, m.n_levels - 1);
        f = lv - (float)l0;
        live = raw >= 0.0f && raw <= (float)(m.n_levels - 1);
    }
    float a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
    for (int pass = 0; pass < 2; ++pass) {
        const int l = pass == 0 ? l0 : l1;
        const float wl = pass == 0 ? 1.0f - f : f;
        if (pass == 1 && (l1 == l0)) break;
        const CubeTap t = cube_taps(dx, dy, dz, m.res[l]);
        float *acc = pass == 0 ? a : b;
        if (g_dirs && wl != 0.0f) {
            // bilinear weight derivatives; taps dropped at cube corners contribute nothing
            int face; float u, v;
            dir_to_face(dx, dy, dz, face, u, v);
            const float Nf = (float)m.res[l];
            const float tx = (u + 1.0f) * 0.5f * Nf - 0.5f, ty = (v + 1.0f) * 0.5f * Nf - 0.5f;
            const float ax = tx - floorf(tx), ay = ty - floorf(ty);
            float gt[4];
            for (int k = 0; k < 4; ++k) {
                gt[k] = 0.f;
                if (t.idx[k] >= 0)
                    for (int c = 0; c < 3; ++c) gt[k] = fmaf(g[c], __ldg(m.level[l] + (size_t)t.idx[k] * 3 + c), gt[k]);
            }
            d_u += wl * 0.5f * Nf * ((1.f - ay) * (gt[1] - gt[0]) + ay * (gt[3] - gt[2]));
            d_v += wl * 0.5f * Nf * ((1.f - ax) * (gt[2] - gt[0]) + ax * (gt[3] - gt[1]));
        }
        for (int k = 0; k < 4; ++k) {
            if (t.idx[k] < 0) continue;
            for (int c = 0; c < 3; ++c) {
                if (g_bias) acc[c] = fmaf(t.w[k], __ldg(m.level[l] + (size_t)t.idx[k] * 3 + c), acc[c]);
                if (m.grad[l] && wl != 0.0f) atomicAdd(m.grad[l] + (size_t)t.idx[k] * 3 + c, wl * t.w[k] * g[c]);
            }
        }
    }
    if (g_bias) {
        float gb = 0.f;
        if (live && l1 != l0) gb = (b[0] - a[0]) * g[0] + (b[1] - a[1]) * g[1] + (b[2] - a[2]) * g[2];
        g_bias[s] = gb;
    }
    if (g_dirs) {
        // chain (u, v) = (cu*A, cv*B)/|M| back to the direction (face table: cubemap.cu:48-60)
        int face; float u, v;
        dir_to_face(dx, dy, dz, face, u, v);
        float gd[3] = {0.f, 0.f, 0.f};
        const float d[3] = {dx, dy, dz};
        const int Mx[6] = {0, 0, 1, 1, 2, 2}, Ax[6] = {2, 2, 0, 0, 0, 0}, Bx[6] = {1, 1, 2, 2, 1, 1};
        const float cu[6] = {-1.f, 1.f, 1.f, 1.f, 1.f, -1.f}, cv[6] = {-1.f, -1.f, 1.f, -1.f, -1.f, -1.f};
        const float am = fabsf(d[Mx[face]]), im = 1.0f / am;
        gd[Ax[face]] += cu[face] * im * d_u;
        gd[Bx[face]] += cv[face] * im * d_v;
        gd[Mx[face]] += -(u * d_u + v * d_v) * im * (d[Mx[face]] > 0.f ? 1.f : -1.f);
        g_dirs[3 * s] = gd[0]; g_dirs[3 * s + 1] = gd[1]; g_dirs[3 * s + 2] = gd[2];
    }
}

// 2-D, linear, clamp or wrap.  tex [H, W, C]; uv [n, 2] with uv.x -> width, uv.y -> height.
// wrap (nvdiffrast's default boundary mode, the one the lat-long -> cube conversion uses): the coordinate is
// reduced to [0,1) and a tap that falls off one edge comes back in at the opposite one.
__device__ __forceinline__ void tex2d_taps(float u, int N, bool wrap, int &i0, int &i1, float &a) {
    if (wrap) u -= floorf(u);
    const float t = u * (float)N - 0.5f;
    const float f0 = floorf(t);
    a = t - f0;
    i0 = (int)f0;
    i1 = i0 + 1;
    if (wrap) {
        if (i0 < 0) i0 += N;
        if (i1 >= N) i1 -= N;
    }
    i0 = min(max(i0, 0), N - 1);
    i1 = min(max(i1, 0), N - 1);
}

template <int C>
__global__ void tex2d_fwd_kernel(const float *__restrict__ tex, int H, int W, bool wrap,
                                 const float *__restrict__ uv, int n, float *__restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int x0, x1, y0, y1;
    float ax, ay;
    tex2d_taps(uv[2 * s], W, wrap, x0, x1, ax);
    tex2d_taps(uv[2 * s + 1], H, wrap, y0, y1, ay);
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float v00 = __ldg(tex + ((size_t)y0 * W + x0) * C + c), v10 = __ldg(tex + ((size_t)y0 * W + x1) * C + c);
        const float v01 = __ldg(tex + ((size_t)y1 * W + x0) * C + c), v11 = __ldg(tex + ((size_t)y1 * W + x1) * C + c);
        const float top = fmaf(ax, v10 - v00, v00), bot = fmaf(ax, v11 - v01, v01);
        out[(size_t)s * C + c] = fmaf(ay, bot - top, top);
    }
}

template <int C>
__global__ void tex2d_bwd_kernel(const float *__restrict__ tex, int H, int W, bool wrap,
                                 const float *__restrict__ uv, const float *__restrict__ go, int n, float *__restrict__ g_tex,
                                 float *__restrict__ g_uv) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int x0, x1, y0, y1;
    float ax, ay;
    tex2d_taps(uv[2 * s], W, wrap, x0, x1, ax);
    tex2d_taps(uv[2 * s + 1], H, wrap, y0, y1, ay);
    float gu = 0.f, gv = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float g = go[(size_t)s * C + c];
        const float v00 = __ldg(tex + ((size_t)y0 * W + x0) * C + c), v10 = __ldg(tex + ((size_t)y0 * W + x1) * C + c);
        const float v01 = __ldg(tex + ((size_t)y1 * W + x0) * C + c), v11 = __ldg(tex + ((size_t)y1 * W + x1) * C + c);
        gu += g * ((1.f - ay) * (v10 - v00) + ay * (v11 - v01));
        gv += g * ((1.f - ax) * (v01 - v00) + ax * (v11 - v10));
        if (g_tex) {
            atomicAdd(g_tex + ((size_t)y0 * W + x0) * C + c, g * (1.f - ax) * (1.f - ay));
            atomicAdd(g_tex + ((size_t)y0 * W + x1) * C + c, g * ax * (1.f - ay));
            atomicAdd(g_tex + ((size_t)y1 * W + x0) * C + c, g * (1.f - ax) * ay);
            atomicAdd(g_tex + ((size_t)y1 * W + x1) * C + c, g * ax * ay);
        }
    }
    if (g_uv) { g_uv[2 * s] = gu * (float)W; g_uv[2 * s + 1] = gv * (float)H; }
}

}  // namespace

extern "C" {

int rsdf_tex2d_fwd(const float *tex, int H, int W, int C, int wrap, const float *uv, int n, float *out,
                   void *stream) {
    if (n == 0) return 0;
    if (!tex || !uv || !out) return RSDF_EBADARG;
    const int blocks = rsdf_div_up(n, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 2) tex2d_fwd_kernel<2><<<blocks, 256, 0, st>>>(tex, H, W, wrap != 0, uv, n, out);
    else if (C == 3) tex2d_fwd_kernel<3><<<blocks, 256, 0, st>>>(tex, H, W, wrap != 0, uv, n, out);
    else if (C == 1) tex2d_fwd_kernel<1><<<blocks, 256, 0, st>>>(tex, H, W, wrap != 0, uv, n, out);
    else return RSDF_EBADARG;
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_tex2d_bwd(const float *tex, int H, int W, int C, int wrap, const float *uv, const float *grad_out, int n,
                   float *grad_tex, float *grad_uv, void *stream) {
    if (n == 0) return 0;
    if (!tex || !uv || !grad_out) return RSDF_EBADARG;
    const int blocks = rsdf_div_up(n, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 2) tex2d_bwd_kernel<2><<<blocks, 256, 0, st>>>(tex, H, W, wrap != 0, uv, grad_out, n, grad_tex, grad_uv);
    else if (C == 3) tex2d_bwd_kernel<3><<<blocks, 256, 0, st>>>(tex, H, W, wrap != 0, uv, grad_out, n, grad_tex, grad_uv);
    else if (C == 1) tex2d_bwd_kernel<1><<<blocks, 256, 0, st>>>(tex, H, W, wrap != 0, uv, grad_out, n, grad_tex, grad_uv);
    else return RSDF_EBADARG;
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_cube_sample_fwd(const float *const *levels_host, const int *res_host, int n_levels, const float *dirs,
                         const float *mip_level_bias, int n, float *out, void *stream) {
    if (n == 0) return 0;
    if (!levels_host || !res_host || n_levels < 1 || n_levels > 8 || !dirs || !out) return RSDF_EBADARG;
    MipStack m;
    m.n_levels = n_levels;
    for (int l = 0; l < 8; ++l) {
        m.level[l] = l < n_levels ? levels_host[l] : nullptr;
        m.grad[l] = nullptr;
        m.res[l] = l < n_levels ? res_host[l] : 0;
    }
    cube_sample_fwd_kernel<<<rsdf_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(m, dirs, mip_level_bias, n, out);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_cube_sample_bwd(const float *const *levels_host, float *const *grad_levels_host, const int *res_host,
                         int n_levels, const float *dirs, const float *mip_level_bias, const float *grad_out,
                         int n, float *grad_bias, float *grad_dirs, void *stream) {
    if (n == 0) return 0;
    if (!levels_host || !res_host || n_levels < 1 || n_levels > 8 || !dirs || !grad_out) return RSDF_EBADARG;
    MipStack m;
    m.n_levels = n_levels;
    for (int l = 0; l < 8; ++l) {
        m.level[l] = l < n_levels ? levels_host[l] : nullptr;
        m.grad[l] = (l < n_levels && grad_levels_host) ? grad_levels_host[l] : nullptr;
        m.res[l] = l < n_levels ? res_host[l] : 0;
    }
    cube_sample_bwd_kernel<<<rsdf_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(m, dirs, mip_level_bias,
                                                                                 grad_out, n, grad_bias, grad_dirs);
    RSDF_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
