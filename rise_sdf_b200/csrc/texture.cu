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
#include "texture.cuh"

namespace {
using namespace rsdf_tex;

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
        cube_fetch_mip<3>(m, bias[s], dx, dy, dz, r);
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
    if (m.n_levels > 1 && bias != nullptr) mip_levels(bias[s], m.n_levels, l0, l1, f, live);
    float a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
    cube_level_bwd(m.level[l0], m.grad[l0], m.res[l0], dx, dy, dz, 1.0f - f, g, g_dirs != nullptr, a, d_u, d_v);
    if (l1 != l0) cube_level_bwd(m.level[l1], m.grad[l1], m.res[l1], dx, dy, dz, f, g, g_dirs != nullptr, b, d_u, d_v);
    if (g_bias) {
        float gb = 0.f;
        if (live && l1 != l0) gb = (b[0] - a[0]) * g[0] + (b[1] - a[1]) * g[1] + (b[2] - a[2]) * g[2];
        g_bias[s] = gb;
    }
    if (g_dirs) {
        float gd[3] = {0.f, 0.f, 0.f};
        face_uv_grad_to_dir(dx, dy, dz, d_u, d_v, gd);
        g_dirs[3 * s] = gd[0]; g_dirs[3 * s + 1] = gd[1]; g_dirs[3 * s + 2] = gd[2];
    }
}

// 2-D, linear, clamp or wrap (tex2d_taps in texture.cuh).  tex [H, W, C]; uv [n, 2] with uv.x -> width, uv.y -> height.
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
