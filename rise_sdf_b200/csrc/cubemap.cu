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
                    // V = the direction of the OUTPUT texel of the forward pass: `me` forward, `o` in the gather
                    // backward (the reference's scatter evaluates the weight from the output texel's side; at
                    // roughness 0.08 the two sides differ by 5e-3 through the fp32 cancellation in D_ggx)
                    const float vDotH = TRANSPOSED ? fmaxf(o.x * hx + o.y * hy + o.z * hz, 0.0f)
                                                   : fmaxf(me.x * hx + me.y * hy + me.z * hz, 0.0f);
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

// ---- specular prefilter as a CACHED SPARSE OPERATOR ---------------------------------------------------------------
// The pair weight s(p,x) = (L_x.N_p) D_ggx(N_p.H) / 4 depends on (resolution, roughness, cutoff) only -- not on the
// cube map -- and the prefilter runs every training step (forward + backward) on a map that is the only thing that
// changes.  With 180 GB of HBM per GPU the whole operator fits: the weights of every (output texel, lobe-box texel)
// pair are evaluated ONCE, with exactly the arithmetic of specular_kernel above (IEEE sqrt / divisions: the GGX lobe of
// the finest level is sharp enough that one ulp in N.H moves D by 1e-3), and stored as a dense run per output texel:
//     weights[offset[p] + sum of the box areas of faces < s + (y - ymin) * box_w + (x - xmin)]
// (0 outside the cutoff circle).  4.9 GB for the 512 ... 16 pyramid of the config.  A training step then streams them:
// one warp per output texel, lanes over the run (coalesced 128-byte reads), the source texel gathered from L2,
// three FMAs per pair and a shuffle reduction.  The backward is the same gather with the roles of the two texels
// swapped (the lobe box is symmetric), scaled by the receiving texel's solid angle -- the semantics of
// specular_kernel<true>.  HBM-bound: ~1 ms per pass instead of ~12 ms of fp32 divisions.
// `me` = the texel that owns the run, `o` = the other one; V (in V.H) is the forward pass's OUTPUT texel: `me` for the
// forward operator, `o` for the transposed one (see specular_kernel)
template <bool TRANSPOSED>
__device__ __forceinline__ float pair_weight(const float4 me, const float4 o, float alphaSqr, float cutoff) {
    const float dp = o.x * me.x + o.y * me.y + o.z * me.z;
    if (!(dp >= cutoff)) return 0.0f;
    float hx = o.x + me.x, hy = o.y + me.y, hz = o.z + me.z;
    const float hl = sqrtf(hx * hx + hy * hy + hz * hz);
    if (hl > 0.0f) { hx /= hl; hy /= hl; hz /= hl; } else { hx = hy = hz = 0.f; }
    const float wiDotN = fmaxf(dp, 0.0f);
    const float vDotH = TRANSPOSED ? fmaxf(o.x * hx + o.y * hy + o.z * hz, 0.0f)
                                   : fmaxf(me.x * hx + me.y * hy + me.z * hz, 0.0f);
    return wiDotN * ndf_ggx(alphaSqr, vDotH);          // x area / 4 when applied (as in specular_kernel)
}

// warp per output texel p: writes its run of pair weights and (forward operator) wsum[p] = sum_x s(p,x) area(x) / 4
template <bool TRANSPOSED>
__global__ void __launch_bounds__(256)
specular_build_kernel(int N, float roughness, float cutoff, const float4 *__restrict__ table,
                      const float *__restrict__ bounds, const long long *__restrict__ offset,
                      float *__restrict__ weights, float *__restrict__ wsum) {
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= 6 * N * N) return;
    const float4 me = table[p];
    const float alpha = roughness * roughness, alphaSqr = alpha * alpha;
    float *run = weights + offset[p];
    float acc = 0.f;
    for (int s = 0; s < 6; ++s) {
        const float *b = bounds + (size_t)p * 24 + s * 4;
        const int xmin = (int)b[0], xmax = (int)b[1], ymin = (int)b[2], ymax = (int)b[3];
        if (xmin > xmax || ymin > ymax) continue;
        const int bw = xmax - xmin + 1, n = bw * (ymax - ymin + 1);
        const float inv_bw = 1.0f / (float)bw;
        for (int k = lane; k < n; k += 32) {
            int dy = (int)(((float)k + 0.5f) * inv_bw);
            int dx = k - dy * bw;
            if (dx < 0) { --dy; dx += bw; } else if (dx >= bw) { ++dy; dx -= bw; }
            const float4 o = table[(s * N + ymin + dy) * N + xmin + dx];
            const float w = pair_weight<TRANSPOSED>(me, o, alphaSqr, cutoff);
            run[k] = w;
            acc += w * o.w / 4.0f;
        }
        run += n;
    }
    if (!TRANSPOSED) {
        acc = warp_sum(acc);
        if (lane == 0) wsum[p] = acc;
    }
}

// out[p] = scale_p * sum over p's run of weights * src[q]     (src: [6N^2, 4] -- rgb padded to 16 bytes)
//   forward : src = cubemap * area / 4 (pre-multiplied by the caller), scale_p = 1
//   backward: src = grad_out (already divided by wsum through autograd), scale_p = area(p) / 4
// Branch-free inner loop, four pairs per lane and trip: the weight loads (coalesced, streaming) and the 16-byte
// gathers of one trip are all in flight before the first FMA.
template <bool TRANSPOSED>
__global__ void __launch_bounds__(256)
specular_apply_kernel(int N, const float4 *__restrict__ table, const float4 *__restrict__ bounds,
                      const long long *__restrict__ offset, const float *__restrict__ weights,
                      const float4 *__restrict__ src, float *__restrict__ dst) {
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= 6 * N * N) return;
    const float *run = weights + offset[p];
    // lane s < 6 fetches the box of face s; everybody reads it back with a shuffle
    float4 mine = make_float4(1.f, 0.f, 1.f, 0.f);
    if (lane < 6) mine = __ldg(bounds + (size_t)p * 6 + lane);
    float cx = 0.f, cy = 0.f, cz = 0.f;
#pragma unroll 1
    for (int s = 0; s < 6; ++s) {
        const int xmin = (int)__shfl_sync(0xffffffffu, mine.x, s), xmax = (int)__shfl_sync(0xffffffffu, mine.y, s);
        const int ymin = (int)__shfl_sync(0xffffffffu, mine.z, s), ymax = (int)__shfl_sync(0xffffffffu, mine.w, s);
        if (xmin > xmax || ymin > ymax) continue;
        const int bw = xmax - xmin + 1, n = bw * (ymax - ymin + 1);
        const float inv_bw = 1.0f / (float)bw;
        const float4 *face = src + (size_t)(s * N + ymin) * N + xmin;
        for (int k0 = lane; k0 < n; k0 += 128) {
            float w[4];
            int q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = k0 + 32 * u;
                const bool in = k < n;
                w[u] = in ? __ldcs(run + k) : 0.0f;
                const int kk = in ? k : 0;
                int dy = (int)(((float)kk + 0.5f) * inv_bw);
                int dx = kk - dy * bw;
                if (dx < 0) { --dy; dx += bw; } else if (dx >= bw) { ++dy; dx -= bw; }
                q[u] = dy * N + dx;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 v = __ldg(face + q[u]);
                cx = fmaf(v.x, w[u], cx); cy = fmaf(v.y, w[u], cy); cz = fmaf(v.z, w[u], cz);
            }
        }
        run += n;
    }
    cx = warp_sum(cx); cy = warp_sum(cy); cz = warp_sum(cz);
    if (lane == 0) {
        const float sc = TRANSPOSED ? table[p].w / 4.0f : 1.0f;
        dst[3 * p] = cx * sc; dst[3 * p + 1] = cy * sc; dst[3 * p + 2] = cz * sc;
    }
}

}  // namespace

extern "C" {

int rsdf_specular_build(const float *table, const float *bounds, const long long *offset, int res, float roughness,
                        float costheta_cutoff, int transposed, float *weights, float *wsum, void *stream) {
    if (!table || !bounds || !offset || !weights || (!transposed && !wsum) || res < 1) return RSDF_EBADARG;
    const long long threads = 6LL * res * res * 32;
    if (transposed)
        specular_build_kernel<true><<<rsdf_div_up(threads, 256), 256, 0, (cudaStream_t)stream>>>(
            res, roughness, costheta_cutoff, (const float4 *)table, bounds, offset, weights, wsum);
    else
        specular_build_kernel<false><<<rsdf_div_up(threads, 256), 256, 0, (cudaStream_t)stream>>>(
            res, roughness, costheta_cutoff, (const float4 *)table, bounds, offset, weights, wsum);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_specular_apply(const float *table, const float *bounds, const long long *offset, const float *weights,
                        const float *src, int res, int transposed, float *dst, void *stream) {
    if (!table || !bounds || !offset || !weights || !src || !dst || res < 1) return RSDF_EBADARG;
    const long long threads = 6LL * res * res * 32;
    if (transposed)
        specular_apply_kernel<true><<<rsdf_div_up(threads, 256), 256, 0, (cudaStream_t)stream>>>(
            res, (const float4 *)table, (const float4 *)bounds, offset, weights, (const float4 *)src, dst);
    else
        specular_apply_kernel<false><<<rsdf_div_up(threads, 256), 256, 0, (cudaStream_t)stream>>>(
            res, (const float4 *)table, (const float4 *)bounds, offset, weights, (const float4 *)src, dst);
    RSDF_LAUNCH_CHECK();
    return 0;
}

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
