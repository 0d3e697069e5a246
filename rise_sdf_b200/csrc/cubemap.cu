// K5b -- cube-map prefilter for the split-sum environment light (runs EVERY training step because
// the 6x512x512 base map is learnable: systems/split_occ.py:151-152 -> lib/pbr/light.py:169-180).
// Replaces lib/renderutils/c_src/cubemap.cu:110-350 (diffuse cosine convolution, GGX-lobe bounds,
// GGX specular prefilter; forward + backward).  Same arithmetic per (output texel, source texel)
// pair as the reference, restructured for B200:
//   * per-texel directions and solid angles come from tables built once per resolution and shared
//     by every pair (the reference recomputes 2 atan + 1 rsqrt per pair);
//   * the backward passes are GATHERS, not atomic scatters: the pair weight is
//     s(p,x) * area(x) with s symmetric in (p,x), and the set {p : L_x . N_p >= cutoff} is bounded by
//     the same lobe box as the forward pass, so grad_in[x] = area(x) * sum_p s(p,x) grad_out[p] is a
//     deterministic per-texel reduction (no fp32 atomics, bit-reproducible).
#include <float.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ float pixel_area(int x, int y, int N) {
    if (N > 1) {
        const int H = N / 2;
        x = abs(x - H);
        y = abs(y - H);
        const float dx = atanf((float)(x + 1) / (float)H) - atanf((float)x / (float)H);
        const float dy = atanf((float)(y + 1) / (float)H) - atanf((float)y / (float)H);
        return dx * dy;
    }
    return 1.0f;
}

__device__ __forceinline__ float3 cube_to_dir(int x, int y, int side, int N) {
    const float fx = 2.0f * (((float)x + 0.5f) / (float)N) - 1.0f;
    const float fy = 2.0f * (((float)y + 0.5f) / (float)N) - 1.0f;
    float3 v;
    switch (side) {
        case 0: v = make_float3(1.f, -fy, -fx); break;
        case 1: v = make_float3(-1.f, -fy, fx); break;
        case 2: v = make_float3(fx, 1.f, fy); break;
        case 3: v = make_float3(fx, -1.f, -fy); break;
        case 4: v = make_float3(fx, -fy, 1.f); break;
        default: v = make_float3(-fx, -fy, -1.f); break;
    }
    const float l = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    return l > 0.0f ? make_float3(v.x / l, v.y / l, v.z / l) : make_float3(0.f, 0.f, 0.f);
}
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// table[(s*N + y)*N + x] = (dir.xyz, pixel_area)
__global__ void texel_table_kernel(int N, float4 *__restrict__ table) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 6 * N * N) return;
    const int x = i % N, y = (i / N) % N, s = i / (N * N);
    const float3 d = cube_to_dir(x, y, s, N);
    table[i] = make_float4(d.x, d.y, d.z, pixel_area(x, y, N));
}

// ---- diffuse: out[p] = sum_x in[x] * clamp(N_p.L_x, 0, .999) * area(x) / pi  -----------------
// TRANSPOSED == false: forward.  true: backward (grad_in[x] = area(x)/pi * sum_p grad_out[p] * clamp(..))
template <bool TRANSPOSED>
__global__ void __launch_bounds__(128)
diffuse_kernel(int N, const float4 *__restrict__ table, const float *__restrict__ src, float *__restrict__ dst) {
    extern __shared__ float4 sh[];               // tile of the source side: (dir or value) staging
    const int n_tex = 6 * N * N;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    float4 me = make_float4(0, 0, 0, 0);
    if (p < n_tex) me = table[p];
    float cx = 0.f, cy = 0.f, cz = 0.f;
    for (int base = 0; base < n_tex; base += blockDim.x) {
        const int q = base + threadIdx.x;
        __syncthreads();
        if (q < n_tex) {
            sh[threadIdx.x] = table[q];
            sh[blockDim.x + threadIdx.x] = make_float4(src[3 * q], src[3 * q + 1], src[3 * q + 2], 0.f);
        }
        __syncthreads();
        const int lim = min((int)blockDim.x, n_tex - base);
        for (int k = 0; k < lim; ++k) {
            const float4 o = sh[k], v = sh[blockDim.x + k];
            const float cs = fminf(fmaxf(me.x * o.x + me.y * o.y + me.z * o.z, 0.0f), 0.999f);
            const float w = TRANSPOSED ? cs * me.w / 3.141592f : cs * o.w / 3.141592f;
            cx += v.x * w; cy += v.y * w; cz += v.z * w;
        }
    }
    if (p < n_tex) { dst[3 * p] = cx; dst[3 * p + 1] = cy; dst[3 * p + 2] = cz; }
}

// ---- GGX lobe bounds (cubemap.cu:181-244): per output texel, per face, the box of texels with
// L.VNR >= cutoff.  One warp per output texel; the 16x16-tile culling test is kept, lanes split
// the tiles and the min/max are shuffle-reduced.
// corner[(s*(nt+1) + j)*(nt+1) + i] = cube_to_dir(min(i*16,N), min(j*16,N), s, N): the tile-corner
// directions every output texel's culling test re-derives in the reference
__global__ void tile_corner_kernel(int N, float4 *__restrict__ corner) {
    const int nt = (N + 15) / 16, nc = nt + 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 6 * nc * nc) return;
    const int cx = i % nc, cy = (i / nc) % nc, s = i / (nc * nc);
    const float3 d = cube_to_dir(min(cx * 16, N), min(cy * 16, N), s, N);
    corner[i] = make_float4(d.x, d.y, d.z, 0.f);
}

__global__ void specular_bounds_kernel(int N, float cutoff, const float4 *__restrict__ table,
                                       const float4 *__restrict__ corner, float *__restrict__ out) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gw >= 6 * N * N) return;
    const float4 vn = table[gw];
    const int TILE = 16, nt = (N + TILE - 1) / TILE, nc = nt + 1;
    for (int s = 0; s < 6; ++s) {
        int mnx = N - 1, mxx = 0, mny = N - 1, mxy = 0;
        for (int t = lane; t < nt * nt; t += 32) {
            const int tx = t / nt, ty = t % nt;
            const int tsx = tx * TILE, tsy = ty * TILE, tex = min((tx + 1) * TILE, N), tey = min((ty + 1) * TILE, N);
            const float4 L0 = corner[(s * nc + ty) * nc + tx], L1 = corner[(s * nc + ty) * nc + tx + 1];
            const float4 L2 = corner[(s * nc + ty + 1) * nc + tx], L3 = corner[(s * nc + ty + 1) * nc + tx + 1];
            const float minx = fminf(fminf(L0.x, L1.x), fminf(L2.x, L3.x)), maxx = fmaxf(fmaxf(L0.x, L1.x), fmaxf(L2.x, L3.x));
            const float miny = fminf(fminf(L0.y, L1.y), fminf(L2.y, L3.y)), maxy = fmaxf(fmaxf(L0.y, L1.y), fmaxf(L2.y, L3.y));
            const float minz = fminf(fminf(L0.z, L1.z), fminf(L2.z, L3.z)), maxz = fmaxf(fmaxf(L0.z, L1.z), fmaxf(L2.z, L3.z));
            const float maxdp = fmaxf(minx * vn.x, maxx * vn.x) + fmaxf(miny * vn.y, maxy * vn.y) +
                                fmaxf(minz * vn.z, maxz * vn.z);
            if (maxdp >= cutoff) {
                for (int y = tsy; y < tey; ++y)
                    for (int x = tsx; x < tex; ++x) {
                        const float4 L = table[(s * N + y) * N + x];
                        if (L.x * vn.x + L.y * vn.y + L.z * vn.z >= cutoff) {
                            mnx = min(mnx, x); mxx = max(mxx, x); mny = min(mny, y); mxy = max(mxy, y);
                        }
                    }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
            mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o)); mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        }
        if (lane == 0) {
            float *o = out + (size_t)gw * 24 + s * 4;
            o[0] = (float)mnx; o[1] = (float)mxx; o[2] = (float)mny; o[3] = (float)mxy;
        }
    }
}

__device__ __forceinline__ float ndf_ggx(float alphaSqr, float cosTheta) {
    const float c = fminf(fmaxf(cosTheta, 0.0f), 1.0f);
    const float d = (c * alphaSqr - c) * c + 1.0f;
    return alphaSqr / (d * d * 3.14159265358979323846f);
}

// ---- specular (cubemap.cu:246-350).  Forward: out[p] = (sum_x in[x] w, sum_x w),
// w = (L.V) D_ggx(V.H) area(x)/4 over the lobe box.  Backward (TRANSPOSED): gather over p in the
// lobe box of x: grad_in[x] = area(x)/4 * sum_p (L.V) D_ggx grad_out[p].
template <bool TRANSPOSED>
__global__ void specular_kernel(int N, float roughness, float cutoff, const float4 *__restrict__ table,
                                const float *__restrict__ bounds, const float *__restrict__ src,
                                float *__restrict__ dst) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= 6 * N * N) return;
    const float4 me = table[p];
    const float alpha = roughness * roughness, alphaSqr = alpha * alpha;
    float cx = 0.f, cy = 0.f, cz = 0.f, wsum = 0.f;
    for (int s = 0; s < 6; ++s) {
        const float *b = bounds + (size_t)p * 24 + s * 4;
        const int xmin = (int)b[0], xmax = (int)b[1], ymin = (int)b[2], ymax = (int)b[3];
        if (xmin > xmax) continue;
        for (int y = ymin; y <= ymax; ++y) {
            for (int x = xmin; x <= xmax; ++x) {
                const int q = (s * N + y) * N + x;
                const float4 o = table[q];
                const float dp = o.x * me.x + o.y * me.y + o.z * me.z;
                if (dp >= cutoff) {
                    float hx = o.x + me.x, hy = o.y + me.y, hz = o.z + me.z;
                    const float hl = sqrtf(hx * hx + hy * hy + hz * hz);
                    if (hl > 0.0f) { hx /= hl; hy /= hl; hz /= hl; } else { hx = hy = hz = 0.f; }
                    const float wiDotN = fmaxf(dp, 0.0f);
                    const float vDotH = fmaxf(me.x * hx + me.y * hy + me.z * hz, 0.0f);
                    const float w = wiDotN * ndf_ggx(alphaSqr, vDotH) * (TRANSPOSED ? me.w : o.w) / 4.0f;
                    cx += src[3 * q] * w; cy += src[3 * q + 1] * w; cz += src[3 * q + 2] * w;
                    wsum += w;
                }
            }
        }
    }
    if (TRANSPOSED) {
        dst[3 * p] = cx; dst[3 * p + 1] = cy; dst[3 * p + 2] = cz;
    } else {
        dst[4 * p] = cx; dst[4 * p + 1] = cy; dst[4 * p + 2] = cz; dst[4 * p + 3] = wsum;
    }
}

}  // namespace

extern "C" {

int rsdf_cubemap_texel_table(int res, float *table, void *stream) {
    if (!table || res < 1) return RSDF_EBADARG;
    texel_table_kernel<<<rsdf_div_up(6 * res * res, 256), 256, 0, (cudaStream_t)stream>>>(res, (float4 *)table);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_diffuse_cubemap(const float *table, const float *src, int res, int transposed, float *dst, void *stream) {
    if (!table || !src || !dst) return RSDF_EBADARG;
    const int n = 6 * res * res, T = 128;
    const size_t sm = 2 * T * sizeof(float4);
    if (transposed)
        diffuse_kernel<true><<<rsdf_div_up(n, T), T, sm, (cudaStream_t)stream>>>(res, (const float4 *)table, src, dst);
    else
        diffuse_kernel<false><<<rsdf_div_up(n, T), T, sm, (cudaStream_t)stream>>>(res, (const float4 *)table, src, dst);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_specular_bounds(const float *table, int res, float costheta_cutoff, float *corner_scratch,
                         float *bounds, void *stream) {
    if (!table || !bounds || !corner_scratch) return RSDF_EBADARG;
    const int nc = (res + 15) / 16 + 1;
    tile_corner_kernel<<<rsdf_div_up(6 * nc * nc, 256), 256, 0, (cudaStream_t)stream>>>(res, (float4 *)corner_scratch);
    const long long threads = 6LL * res * res * 32;
    specular_bounds_kernel<<<rsdf_div_up(threads, 256), 256, 0, (cudaStream_t)stream>>>(
        res, costheta_cutoff, (const float4 *)table, (const float4 *)corner_scratch, bounds);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_specular_cubemap(const float *table, const float *bounds, const float *src, int res, float roughness,
                          float costheta_cutoff, int transposed, float *dst, void *stream) {
    if (!table || !bounds || !src || !dst) return RSDF_EBADARG;
    const int n = 6 * res * res;
    if (transposed)
        specular_kernel<true><<<rsdf_div_up(n, 128), 128, 0, (cudaStream_t)stream>>>(
            res, roughness, costheta_cutoff, (const float4 *)table, bounds, src, dst);
    else
        specular_kernel<false><<<rsdf_div_up(n, 128), 128, 0, (cudaStream_t)stream>>>(
            res, roughness, costheta_cutoff, (const float4 *)table, bounds, src, dst);
    RSDF_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
