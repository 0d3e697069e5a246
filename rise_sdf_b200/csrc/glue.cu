// Elementwise stages either side of the big kernels, each ONE launch instead of a chain of torch ops:
//   * VanillaFrequency position encoding                          models/network_utils.py:14-40
//   * ray generation from (image, pixel) draws                    systems/split_occ.py:58-103, models/ray_utils.py:32-56
//   * background composite + sRGB + clamp epilogue                models/neus.py:307-311, models/split_mixed_occ.py:416-437,
//                                                                 lib/pbr/utils/nvdiffrecmc_util.py:95-103
//   * loss block (masked rgb MSE, mask BCE)                       systems/neus.py:98-135, systems/criterions.py:155-159
//   * occupancy-grid update: cell -> jittered point, EMA-max, threshold + bool grid + bit-packed grid
//                                                                 lib/nerfacc/grid.py:196-239
// All HBM-bound streaming kernels; every reduction is two-stage with a fixed order (bit-reproducible, no float atomics).
#include "common.cuh"

namespace {

constexpr int MAX_FREQS = 16;
struct FreqMask {
    float m[MAX_FREQS];
};

// out[s, (2k + f) * C + c] = (f ? cos : sin)(2^k * (x[s,c] * scale + offset)) * mask[k]
// thread per (sample, k, c): sincosf shares the range reduction; a block stages [256 / (F C) samples] rows and writes
// them as one contiguous, coalesced range
__global__ void __launch_bounds__(256)
freq_fwd_kernel(const float *__restrict__ x, int n, int C, int F, float scale, float offset, const FreqMask mask,
                float *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over n * F * C
    const int per = F * C;
    if (i >= (long long)n * per) return;
    const long long s = i / per;
    const int r = (int)(i - s * per), k = r / C, c = r - k * C;
    const float v = __fadd_rn(__fmul_rn(x[s * C + c], scale), offset);      // two roundings, like `x * x_scale + x_offset` in torch
    const float f = (float)(1 << k);
    float sn, cs;
    sincosf(f * v, &sn, &cs);
    float *o = out + s * (2 * per) + (size_t)(2 * k) * C + c;
    o[0] = sn * mask.m[k];
    o[C] = cs * mask.m[k];
}
// grad_x[s,c] = scale * sum_k 2^k mask[k] (cos(.) g_sin - sin(.) g_cos)
__global__ void __launch_bounds__(256)
freq_bwd_kernel(const float *__restrict__ x, const float *__restrict__ go, int n, int C, int F, float scale, float offset,
                const FreqMask mask, float *__restrict__ gx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over n * C
    if (i >= (long long)n * C) return;
    const long long s = i / C;
    const int c = (int)(i - s * C);
    const float v = __fadd_rn(__fmul_rn(x[i], scale), offset);
    const float *g = go + s * (2 * F * C);
    float acc = 0.f;
    for (int k = 0; k < F; ++k) {
        const float f = (float)(1 << k);
        float sn, cs;
        sincosf(f * v, &sn, &cs);
        acc += f * mask.m[k] * (cs * g[(2 * k) * C + c] - sn * g[(2 * k + 1) * C + c]);
    }
    gx[i] = acc * scale;
}

// rays[i] = (c2w[index_i][:, 3], normalize(directions[y_i, x_i] @ c2w[index_i][:3,:3]^T))
__global__ void __launch_bounds__(256)
get_rays_kernel(const float *__restrict__ directions, const float *__restrict__ c2w, const long long *__restrict__ index,
                const long long *__restrict__ px, const long long *__restrict__ py, int n, int W, int H, int n_img,
                float *__restrict__ rays) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long im = n_img == 1 ? 0 : index[i];
    const float *d = directions + ((size_t)py[i] * W + px[i]) * 3;
    const float *m = c2w + (size_t)im * 12;
    const float dx = d[0], dy = d[1], dz = d[2];
    const float rx = dx * m[0] + dy * m[1] + dz * m[2];
    const float ry = dx * m[4] + dy * m[5] + dz * m[6];
    const float rz = dx * m[8] + dy * m[9] + dz * m[10];
    const float inv = 1.0f / fmaxf(sqrtf(rx * rx + ry * ry + rz * rz), 1e-12f);      // F.normalize(p=2, eps=1e-12)
    float *o = rays + (size_t)i * 6;
    o[0] = m[3]; o[1] = m[7]; o[2] = m[11];
    o[3] = rx * inv; o[4] = ry * inv; o[5] = rz * inv;
}

__device__ __forceinline__ float srgb_of(float f) {
    return f <= 0.0031308f ? f * 12.92f : powf(fmaxf(f, 0.0031308f), 1.0f / 2.4f) * 1.055f - 0.055f;
}
__device__ __forceinline__ float dsrgb_of(float f) {     // d srgb / d f as torch.where + pow + clamp differentiate it
    return f <= 0.0031308f ? 12.92f : (1.055f / 2.4f) * powf(f, 1.0f / 2.4f - 1.0f);
}
// out = rgb + bg (1 - opacity);  SRGB: out = clamp(rgb_to_srgb(out), 0, 1)
template <bool SRGB>
__global__ void __launch_bounds__(256)
composite_fwd_kernel(const float *__restrict__ rgb, const float *__restrict__ op, const float *__restrict__ bg, int n,
                     float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // over n * 3
    if (i >= 3 * n) return;
    const int r = i / 3, c = i - 3 * r;
    float v = __fadd_rn(rgb[i], __fmul_rn(bg[c], __fsub_rn(1.0f, op[r])));       // op by op, like the torch expression
    if (SRGB) v = fminf(fmaxf(srgb_of(v), 0.0f), 1.0f);
    out[i] = v;
}
template <bool SRGB>
__global__ void __launch_bounds__(256)
composite_bwd_kernel(const float *__restrict__ rgb, const float *__restrict__ op, const float *__restrict__ bg,
                     const float *__restrict__ go, int n, float *__restrict__ g_rgb, float *__restrict__ g_op) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;      // over rays
    if (r >= n) return;
    const float o = op[r];
    float gop = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float g = go[3 * r + c];
        if (SRGB) {
            const float lin = rgb[3 * r + c] + bg[c] * (1.0f - o);
            const float s = srgb_of(lin);
            g = (s >= 0.0f && s <= 1.0f) ? g * dsrgb_of(lin) : 0.0f;
        }
        g_rgb[3 * r + c] = g;
        gop -= g * bg[c];
    }
    if (g_op) g_op[r] = gop;
}

// ---- neus loss block ------------------------------------------------------------------------------------------------
// per ray: valid = op > 0;  sq = sum_c (full - target)^2 (valid rays);
//          o = clamp(op, 1e-3, 1 - 1e-3);  bce = -(fg log o + (1 - fg) log(1 - o))
// partials[b] = (sum sq, count valid, sum bce, 0) per block; a second one-block launch adds them in order.
__global__ void __launch_bounds__(256)
neus_loss_fwd_kernel(const float *__restrict__ full, const float *__restrict__ op, const float *__restrict__ target,
                     const float *__restrict__ fg, int n, float4 *__restrict__ partials) {
    float sq = 0.f, cnt = 0.f, bce = 0.f;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const float o = op[r];
        float e = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float d = full[3 * r + c] - target[3 * r + c];
            e = fmaf(d, d, e);
        }
        if (o > 0.0f) { sq += e; cnt += 1.0f; }
        const float oc = fminf(fmaxf(o, 1e-3f), 1.0f - 1e-3f), t = fg[r];
        bce -= t * logf(oc) + (1.0f - t) * logf(1.0f - oc);
    }
    sq = warp_sum(sq); cnt = warp_sum(cnt); bce = warp_sum(bce);
    __shared__ float4 sh[8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = make_float4(sq, cnt, bce, 0.f);
    __syncthreads();
    if (threadIdx.x == 0) {
        float4 a = sh[0];
        for (int k = 1; k < 8; ++k) { a.x += sh[k].x; a.y += sh[k].y; a.z += sh[k].z; }
        partials[blockIdx.x] = a;
    }
}
__global__ void sum4_finish_kernel(const float4 *__restrict__ partials, int nb, float4 *__restrict__ out) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = threadIdx.x; i < nb; i += 32) { a.x += partials[i].x; a.y += partials[i].y; a.z += partials[i].z; a.w += partials[i].w; }
    a.x = warp_sum(a.x); a.y = warp_sum(a.y); a.z = warp_sum(a.z); a.w = warp_sum(a.w);
    if (threadIdx.x == 0) *out = a;
}
// cot = (d loss / d sq_sum, -, d loss / d bce_sum): grads w.r.t. full [n,3] and (the BCE leg of) opacity [n]
__global__ void __launch_bounds__(256)
neus_loss_bwd_kernel(const float *__restrict__ full, const float *__restrict__ op, const float *__restrict__ target,
                     const float *__restrict__ fg, const float *__restrict__ cot, int n, float *__restrict__ g_full,
                     float *__restrict__ g_op) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const float c_sq = cot[0], c_bce = cot[2];
    const float o = op[r];
    const bool valid = o > 0.0f;
    float gop = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c)
        g_full[3 * r + c] = valid ? 2.0f * c_sq * (full[3 * r + c] - target[3 * r + c]) : 0.0f;
    if (o >= 1e-3f && o <= 1.0f - 1e-3f) {               // torch.clamp passes the gradient on the closed interval
        const float t = fg[r];
        gop += c_bce * (-(t / o) + (1.0f - t) / (1.0f - o));
    }
    g_op[r] = gop;
}

// ---- occupancy grid update ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
occ_points_kernel(const long long *__restrict__ indices, const float *__restrict__ jitter, long long n, int res,
                  const float3 lo, const float3 ext, float *__restrict__ x) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long cell = indices ? indices[i] : i;
    const float cx = (float)(cell / ((long long)res * res)), cy = (float)((cell / res) % res), cz = (float)(cell % res);
    const float r = (float)res;
    // torch evaluates ((coords + jitter) / r) * ext + lo op by op: no fused multiply-add
    x[3 * i] = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(cx, jitter[3 * i]), r), ext.x), lo.x);
    x[3 * i + 1] = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(cy, jitter[3 * i + 1]), r), ext.y), lo.y);
    x[3 * i + 2] = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(cz, jitter[3 * i + 2]), r), ext.z), lo.z);
}
// phase A: occs[idx] = snapshot[idx] * decay (duplicates write the same value); phase B: max with the new evaluations.
// Non-negative floats order like their bit patterns, so atomicMax on the bits is an exact float max: the result does not
// depend on the order in which duplicates arrive (torch's gather/scatter form leaves that to chance).
__global__ void __launch_bounds__(256)
occ_decay_kernel(float *__restrict__ occs, const float *__restrict__ snapshot, const long long *__restrict__ indices,
                 long long n, float decay) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long c = indices ? indices[i] : i;
    occs[c] = snapshot[c] * decay;
}
__global__ void __launch_bounds__(256)
occ_max_kernel(float *__restrict__ occs, const long long *__restrict__ indices, const float *__restrict__ occ, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long c = indices ? indices[i] : i;
    atomicMax(reinterpret_cast<unsigned int *>(occs + c), __float_as_uint(fmaxf(occ[i], 0.0f)));
}
__global__ void __launch_bounds__(256)
sum_partials_kernel(const float *__restrict__ v, long long n, float4 *__restrict__ partials) {
    float a = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a += v[i];
    a = warp_sum(a);
    __shared__ float sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = sh[0];
        for (int k = 1; k < 8; ++k) t += sh[k];
        partials[blockIdx.x] = make_float4(t, 0.f, 0.f, 0.f);
    }
}
// binary = occs > min(mean(occs), thre); one thread per 32 cells writes the bool bytes and the packed word
__global__ void __launch_bounds__(256)
occ_threshold_kernel(const float *__restrict__ occs, long long n_cells, const float4 *__restrict__ total, float thre,
                     uint8_t *__restrict__ binaries, uint32_t *__restrict__ bits) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w * 32 >= n_cells) return;
    const float t = fminf(total->x / (float)n_cells, thre);
    uint32_t word = 0;
#pragma unroll 8
    for (int b = 0; b < 32; ++b) {
        const long long c = w * 32 + b;
        const bool on = c < n_cells && occs[c] > t;
        if (c < n_cells) binaries[c] = on ? 1 : 0;
        word |= (on ? 1u : 0u) << b;
    }
    if (bits) bits[w] = word;
}

// ---- NeuS alpha for the no-grad passes (visibility filter of `sampling`, eval) -------------------------------
// models/neus.py:128-150 == models/split_mixed_occ.py:151-177 as ONE launch instead of ~20, in the op order of the
// torch chain (no contraction: every product and sum rounded on its own, IEEE division, accurate expf), so that the
// visibility masks are the ones the op-by-op path produces.
__device__ __forceinline__ float sigmoid_rn(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }
__global__ void neus_alpha_kernel(const float *__restrict__ sdf, const float *__restrict__ normals,
                                  const float *__restrict__ dirs, const float *__restrict__ dists,
                                  const float *__restrict__ inv_s_ptr, float ratio, float one_minus_ratio, int n,
                                  float *__restrict__ alpha) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float inv_s = fminf(fmaxf(__ldg(inv_s_ptr), 1e-6f), 1e6f);
    const float p0 = __fmul_rn(dirs[3 * (size_t)i], normals[3 * (size_t)i]);
    const float p1 = __fmul_rn(dirs[3 * (size_t)i + 1], normals[3 * (size_t)i + 1]);
    const float p2 = __fmul_rn(dirs[3 * (size_t)i + 2], normals[3 * (size_t)i + 2]);
    const float c = __fadd_rn(__fadd_rn(p0, p1), p2);
    const float t1 = fmaxf(__fadd_rn(__fmul_rn(-c, 0.5f), 0.5f), 0.0f);
    const float t2 = fmaxf(-c, 0.0f);
    const float iter_cos = -__fadd_rn(__fmul_rn(t1, one_minus_ratio), __fmul_rn(t2, ratio));
    const float h = __fmul_rn(__fmul_rn(iter_cos, dists[i]), 0.5f);
    const float s = sdf[i];
    const float prev_cdf = sigmoid_rn(__fmul_rn(__fsub_rn(s, h), inv_s));
    const float next_cdf = sigmoid_rn(__fmul_rn(__fadd_rn(s, h), inv_s));
    const float a = __fdiv_rn(__fadd_rn(__fsub_rn(prev_cdf, next_cdf), 1e-5f), __fadd_rn(prev_cdf, 1e-5f));
    alpha[i] = fminf(fmaxf(a, 0.0f), 1.0f);
}

// ---- front-to-back visibility rounds (nerfacc._alphas_front_to_back) -----------------------------------------
// lens[r] = number of candidates of ray r evaluated in this round: the next `chunk` samples after the first `done`,
// unless the ray has fewer, or its transmittance at sample `done` (T from the scan kernel) is already below eps.
__global__ void vis_round_lens_kernel(const int32_t *__restrict__ packed, const float *__restrict__ T, int done,
                                      int chunk, float eps, int n_rays, long long S0, long long *__restrict__ lens) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const int base = packed[2 * r], count = packed[2 * r + 1];
    bool active = count > done;
    if (active && done > 0 && T) {
        long long j = (long long)base + done;
        if (j > S0 - 1) j = S0 - 1;
        active = T[j] >= eps;
    }
    lens[r] = active ? (long long)min(count - done, chunk) : 0;
}
// warp per ray: the round's candidates of ray r, packed in ray order at first[r] = csum[r] - lens[r]
__global__ void vis_round_fill_kernel(const int32_t *__restrict__ packed, const long long *__restrict__ lens,
                                      const long long *__restrict__ csum, int done, int n_rays,
                                      const float *__restrict__ ts, const float *__restrict__ te,
                                      long long *__restrict__ idx, float *__restrict__ ts_sel,
                                      float *__restrict__ te_sel, long long *__restrict__ ri_sel) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= n_rays) return;
    const int len = (int)lens[r];
    if (len == 0) return;
    const long long first = csum[r] - len, src0 = (long long)packed[2 * r] + done;
    for (int j = lane; j < len; j += 32) {
        idx[first + j] = src0 + j;
        ts_sel[first + j] = ts[src0 + j];
        te_sel[first + j] = te[src0 + j];
        ri_sel[first + j] = r;
    }
}
__global__ void vis_round_scatter_kernel(const long long *__restrict__ idx, const float *__restrict__ a,
                                         long long n_eval, long long total, float *__restrict__ alphas,
                                         long long *__restrict__ rows) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long d = idx[i];
    alphas[d] = a[i];
    rows[d] = n_eval + i;
}

}  // namespace

extern "C" {

int rsdf_freq_encode_fwd(const float *x, int n, int c, int n_freqs, float x_scale, float x_offset, const float *mask_host,
                         float *out, void *stream) {
    if (n == 0) return 0;
    if (!x || !out || c < 1 || n_freqs < 1 || n_freqs > MAX_FREQS) return RSDF_EBADARG;
    FreqMask m;
    for (int k = 0; k < MAX_FREQS; ++k) m.m[k] = (mask_host && k < n_freqs) ? mask_host[k] : 1.0f;
    const long long total = (long long)n * n_freqs * c;
    freq_fwd_kernel<<<rsdf_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(x, n, c, n_freqs, x_scale, x_offset, m, out);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_freq_encode_bwd(const float *x, const float *grad_out, int n, int c, int n_freqs, float x_scale, float x_offset,
                         const float *mask_host, float *grad_x, void *stream) {
    if (n == 0) return 0;
    if (!x || !grad_out || !grad_x || c < 1 || n_freqs < 1 || n_freqs > MAX_FREQS) return RSDF_EBADARG;
    FreqMask m;
    for (int k = 0; k < MAX_FREQS; ++k) m.m[k] = (mask_host && k < n_freqs) ? mask_host[k] : 1.0f;
    freq_bwd_kernel<<<rsdf_div_up((long long)n * c, 256), 256, 0, (cudaStream_t)stream>>>(x, grad_out, n, c, n_freqs, x_scale,
                                                                                           x_offset, m, grad_x);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_get_rays(const float *directions, const float *c2w, const long long *index, const long long *px,
                  const long long *py, int n, int width, int height, int n_images, float *rays, void *stream) {
    if (n == 0) return 0;
    if (!directions || !c2w || !px || !py || !rays || (n_images > 1 && !index) || n_images < 1) return RSDF_EBADARG;
    get_rays_kernel<<<rsdf_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(directions, c2w, index, px, py, n, width, height,
                                                                            n_images, rays);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_composite_fwd(const float *rgb, const float *opacity, const float *bg, int n, int srgb, float *out, void *stream) {
    if (n == 0) return 0;
    if (!rgb || !opacity || !bg || !out) return RSDF_EBADARG;
    if (srgb) composite_fwd_kernel<true><<<rsdf_div_up(3 * n, 256), 256, 0, (cudaStream_t)stream>>>(rgb, opacity, bg, n, out);
    else composite_fwd_kernel<false><<<rsdf_div_up(3 * n, 256), 256, 0, (cudaStream_t)stream>>>(rgb, opacity, bg, n, out);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_composite_bwd(const float *rgb, const float *opacity, const float *bg, const float *grad_out, int n, int srgb,
                       float *grad_rgb, float *grad_opacity, void *stream) {
    if (n == 0) return 0;
    if (!rgb || !opacity || !bg || !grad_out || !grad_rgb) return RSDF_EBADARG;
    if (srgb) composite_bwd_kernel<true><<<rsdf_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(rgb, opacity, bg, grad_out, n, grad_rgb, grad_opacity);
    else composite_bwd_kernel<false><<<rsdf_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(rgb, opacity, bg, grad_out, n, grad_rgb, grad_opacity);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_neus_loss_fwd(const float *comp_rgb_full, const float *opacity, const float *target_rgb, const float *fg_mask,
                       int n, float *sums4, float *partials, void *stream) {
    if (!sums4 || !partials) return RSDF_EBADARG;
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(sums4, 0, 4 * sizeof(float), (cudaStream_t)stream);
        return e == cudaSuccess ? 0 : (int)e;
    }
    if (!comp_rgb_full || !opacity || !target_rgb || !fg_mask) return RSDF_EBADARG;
    const int nb = rsdf_div_up(n, 256) < RSDF_LOSS_BLOCKS ? rsdf_div_up(n, 256) : RSDF_LOSS_BLOCKS;
    neus_loss_fwd_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(comp_rgb_full, opacity, target_rgb, fg_mask, n,
                                                               (float4 *)partials);
    sum4_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const float4 *)partials, nb, (float4 *)sums4);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_neus_loss_bwd(const float *comp_rgb_full, const float *opacity, const float *target_rgb, const float *fg_mask,
                       const float *cot4, int n, float *grad_comp_rgb_full, float *grad_opacity, void *stream) {
    if (n == 0) return 0;
    if (!comp_rgb_full || !opacity || !target_rgb || !fg_mask || !cot4 || !grad_comp_rgb_full || !grad_opacity)
        return RSDF_EBADARG;
    neus_loss_bwd_kernel<<<rsdf_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(comp_rgb_full, opacity, target_rgb, fg_mask,
                                                                                cot4, n, grad_comp_rgb_full, grad_opacity);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_occ_points(const long long *indices, const float *jitter, long long n, int res, const float *roi, float *x,
                    void *stream) {
    if (n == 0) return 0;
    if (!jitter || !roi || !x || res < 1) return RSDF_EBADARG;
    const float3 lo = make_float3(roi[0], roi[1], roi[2]);
    const float3 ext = make_float3(roi[3] - roi[0], roi[4] - roi[1], roi[5] - roi[2]);
    occ_points_kernel<<<rsdf_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(indices, jitter, n, res, lo, ext, x);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_occ_update(float *occs, const float *snapshot, const long long *indices, const float *occ, long long n,
                    float ema_decay, void *stream) {
    if (n == 0) return 0;
    if (!occs || !snapshot || !occ) return RSDF_EBADARG;
    occ_decay_kernel<<<rsdf_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(occs, snapshot, indices, n, ema_decay);
    occ_max_kernel<<<rsdf_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(occs, indices, occ, n);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_occ_threshold(const float *occs, long long n_cells, float occ_thre, uint8_t *binaries, uint32_t *bits,
                       float *partials, void *stream) {
    if (n_cells == 0) return 0;
    if (!occs || !binaries || !partials) return RSDF_EBADARG;
    float4 *p4 = (float4 *)partials;
    const int nb = RSDF_LOSS_BLOCKS;
    sum_partials_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(occs, n_cells, p4 + 1);
    sum4_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p4 + 1, nb, p4);
    const long long words = (n_cells + 31) / 32;
    occ_threshold_kernel<<<rsdf_div_up(words, 256), 256, 0, (cudaStream_t)stream>>>(occs, n_cells, p4, occ_thre, binaries, bits);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_neus_alpha(const float *sdf, const float *normals, const float *dirs, const float *dists, const float *inv_s,
                    float cos_anneal_ratio, int n, float *alpha, void *stream) {
    if (n == 0) return 0;
    if (!sdf || !normals || !dirs || !dists || !inv_s || !alpha) return RSDF_EBADARG;
    // (1.0 - ratio) is a Python double in the reference, rounded to fp32 when it meets the tensor
    const float omr = (float)(1.0 - (double)cos_anneal_ratio);
    neus_alpha_kernel<<<rsdf_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(sdf, normals, dirs, dists, inv_s,
                                                                             cos_anneal_ratio, omr, n, alpha);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_vis_round_lens(const int32_t *packed_info, const float *trans, int done, int chunk, float early_stop_eps,
                        int n_rays, long long n_samples, long long *lens, void *stream) {
    if (n_rays == 0) return 0;
    if (!packed_info || !lens || chunk < 1 || done < 0) return RSDF_EBADARG;
    vis_round_lens_kernel<<<rsdf_div_up(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(packed_info, trans, done, chunk,
                                                                                    early_stop_eps, n_rays, n_samples, lens);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_vis_round_fill(const int32_t *packed_info, const long long *lens, const long long *lens_cumsum, int done,
                        int n_rays, const float *t_starts, const float *t_ends, long long *idx, float *t_starts_sel,
                        float *t_ends_sel, long long *ray_indices_sel, void *stream) {
    if (n_rays == 0) return 0;
    if (!packed_info || !lens || !lens_cumsum || !t_starts || !t_ends || !idx || !t_starts_sel || !t_ends_sel ||
        !ray_indices_sel)
        return RSDF_EBADARG;
    vis_round_fill_kernel<<<rsdf_div_up(n_rays, 8), 256, 0, (cudaStream_t)stream>>>(
        packed_info, lens, lens_cumsum, done, n_rays, t_starts, t_ends, idx, t_starts_sel, t_ends_sel, ray_indices_sel);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_vis_round_scatter(const long long *idx, const float *alphas_round, long long n_evaluated_before, long long total,
                           float *alphas, long long *rows, void *stream) {
    if (total == 0) return 0;
    if (!idx || !alphas_round || !alphas || !rows) return RSDF_EBADARG;
    vis_round_scatter_kernel<<<rsdf_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(idx, alphas_round,
                                                                                       n_evaluated_before, total, alphas, rows);
    RSDF_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
