// K4 -- NeuS alpha + segmented per-ray transmittance scan + accumulation (fwd + bwd).
//
// Replaces nerfacc.render_weight_from_alpha / accumulate_along_rays (models/neus.py:262-276,
// models/volrend.py:851-885; in-tree twins lib/nerfacc/cuda/csrc/render_weight.cu:86-154,
// render_transmittance.cu:85-145) and the ~20 elementwise launches of get_alpha
// (models/neus.py:128-150).  One warp owns one ray: samples are consumed 32 at a time with a
// shuffle-based inclusive product scan, the running transmittance is carried in a register,
// and the per-ray sums are reduced with shuffles -> no atomics, bit-reproducible run to run.
// The scan reassociates the product (tree order inside a 32-chunk), so T differs from the
// reference's serial loop by a few ulp; tests state the tolerance.
#include "common.cuh"

namespace {

constexpr int WARPS_PER_BLOCK = 4;

__device__ __forceinline__ float warp_incl_prod(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v *= t;
    }
    return v;
}
__device__ __forceinline__ float warp_incl_sum(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}
// inclusive suffix sum (lane i gets sum over lanes >= i)
__device__ __forceinline__ float warp_suffix_sum(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_down_sync(0xffffffffu, v, o);
        if (lane + o < 32) v += t;
    }
    return v;
}

__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK)
weight_fwd_kernel(const int32_t *__restrict__ packed, const float *__restrict__ alphas, int n_rays,
                  float *__restrict__ weights, float *__restrict__ trans) {
    const int ray = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (ray >= n_rays) return;
    const int base = packed[2 * ray], steps = packed[2 * ray + 1];
    float carry = 1.0f;
    for (int c = 0; c < steps; c += 32) {
        const int j = c + lane;
        const float a = j < steps ? alphas[base + j] : 0.0f;
        const float incl = warp_incl_prod(1.0f - a, lane);
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0f;
        const float T = carry * excl;
        if (j < steps) {
            if (weights) weights[base + j] = a * T;
            if (trans) trans[base + j] = T;
        }
        carry *= __shfl_sync(0xffffffffu, incl, 31);
    }
}

// grad_alpha_j = (gw_j T_j - sum_{i>=j} gw_i w_i - sum_{i>j} gT_i T_i) / max(1-a_j, 1e-10)
// (render_weight.cu:137-151 + render_transmittance.cu:137-143), chunks walked back to front.
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK)
weight_bwd_kernel(const int32_t *__restrict__ packed, const float *__restrict__ alphas,
                  const float *__restrict__ weights, const float *__restrict__ trans,
                  const float *__restrict__ gw, const float *__restrict__ gT, int n_rays,
                  float *__restrict__ galpha) {
    const int ray = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (ray >= n_rays) return;
    const int base = packed[2 * ray], steps = packed[2 * ray + 1];
    float carry = 0.0f;  // sum over samples in later chunks of (gw w + gT T)
    const int nchunks = (steps + 31) >> 5;
    for (int ch = nchunks - 1; ch >= 0; --ch) {
        const int j = ch * 32 + lane;
        const bool ok = j < steps;
        const float a = ok ? alphas[base + j] : 0.0f;
        const float T = ok ? trans[base + j] : 0.0f;
        const float w = weights ? (ok ? weights[base + j] : 0.0f) : a * T;
        const float gwj = (gw && ok) ? gw[base + j] : 0.0f;
        const float gTj = (gT && ok) ? gT[base + j] : 0.0f;
        const float sw = warp_suffix_sum(gwj * w, lane);          // inclusive suffix
        const float sT = warp_suffix_sum(gTj * T, lane) - gTj * T;  // exclusive suffix
        if (ok) galpha[base + j] = (gwj * T - (sw + sT + carry)) / fmaxf(1.0f - a, 1e-10f);
        carry += __shfl_sync(0xffffffffu, sw, 0) + __shfl_sync(0xffffffffu, sT + gTj * T, 0);
    }
}

// out[ray, c] = sum_j w_j v[j, c]; channels processed 8 at a time.
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK)
accumulate_fwd_kernel(const int32_t *__restrict__ packed, const float *__restrict__ weights,
                      const float *__restrict__ values, int n_rays, int D, float *__restrict__ out) {
    const int ray = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (ray >= n_rays) return;
    const int base = packed[2 * ray], steps = packed[2 * ray + 1];
    for (int c0 = 0; c0 < D; c0 += 8) {
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.0f;
        for (int j = lane; j < steps; j += 32) {
            const float w = weights[base + j];
            if (values) {
                const float *v = values + (size_t)(base + j) * D + c0;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (c0 + k < D) acc[k] = fmaf(w, v[k], acc[k]);
            } else {
                acc[0] += w;
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (c0 + k < D) {
                const float s = warp_sum(acc[k]);
                if (lane == 0) out[(size_t)ray * D + c0 + k] = s;
            }
        }
    }
}

__global__ void accumulate_bwd_kernel(const int64_t *__restrict__ ri, const float *__restrict__ weights,
                                      const float *__restrict__ values, const float *__restrict__ go,
                                      int n_samples, int D, float *__restrict__ gw,
                                      float *__restrict__ gv) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_samples) return;
    const int r = (int)ri[s];
    const float *g = go + (size_t)r * D;
    if (!values) {
        if (gw) gw[s] = g[0];
        return;
    }
    const float w = weights[s];
    float acc = 0.0f;
    for (int c = 0; c < D; ++c) {
        const float gc = __ldg(g + c);
        acc = fmaf(gc, values[(size_t)s * D + c], acc);
        if (gv) gv[(size_t)s * D + c] = w * gc;
    }
    if (gw) gw[s] = acc;
}

// ---------------------------------------------------------------------------- fused NeuS
struct AlphaTerms {
    float alpha, P, N, e_p, e_n, praw;  // praw = (p+eps)/(P+eps) before clip
};

__device__ __forceinline__ AlphaTerms neus_alpha(float sdf, float cosv, float dist, float inv_s,
                                                 float ratio) {
    // models/neus.py:133-150
    AlphaTerms t;
    const float ic = -(fmaxf(-cosv * 0.5f + 0.5f, 0.0f) * (1.0f - ratio) + fmaxf(-cosv, 0.0f) * ratio);
    t.e_n = sdf + ic * dist * 0.5f;
    t.e_p = sdf - ic * dist * 0.5f;
    t.P = sigmoid_acc(t.e_p * inv_s);
    t.N = sigmoid_acc(t.e_n * inv_s);
    const float p = t.P - t.N;
    t.praw = (p + 1e-5f) / (t.P + 1e-5f);
    t.alpha = fminf(fmaxf(t.praw, 0.0f), 1.0f);
    return t;
}

// One kernel family for both renderers:
//   CD        colour channels accumulated per ray (3: neus radiance; 7 / 24: split-sum stage 0 / 1, models/texture.py:345)
//   NORMALIZE true : `vec` is the raw sdf gradient, n = vec / max(|vec|, eps) inside (models/neus.py:253)
//             false: `vec` is the normal the caller already normalised (models/split_mixed_occ.py:247, needed earlier by
//                    the shading networks); its gradient is returned as is
//   ORIENT    one more output column: sum_i w_i relu(d . n_i), the normal-orientation map of
//             models/split_mixed_occ.py:384-394
// out[n_rays, CD + 5 (+ 1)] = (colours, normal3 (un-normalised sum), opacity, depth (, orientation))
template <int CD, bool NORMALIZE, bool ORIENT>
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK)
sdf_render_fwd_kernel(const int32_t *__restrict__ packed, const float *__restrict__ rays_d,
                      const float *__restrict__ t_starts, const float *__restrict__ t_ends,
                      const float *__restrict__ sdf, const float *__restrict__ vec,
                      const float *__restrict__ rgb, const float *__restrict__ inv_s_ptr,
                      float ratio, float eps, int n_rays, float *__restrict__ alpha_out,
                      float *__restrict__ w_out, float *__restrict__ T_out,
                      float *__restrict__ out) {
    constexpr int NO = CD + 5 + (ORIENT ? 1 : 0);
    const int ray = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (ray >= n_rays) return;
    const int base = packed[2 * ray], steps = packed[2 * ray + 1];
    const float inv_s = fminf(fmaxf(__ldg(inv_s_ptr), 1e-6f), 1e6f);
    const float dx = rays_d[3 * ray], dy = rays_d[3 * ray + 1], dz = rays_d[3 * ray + 2];
    float acc[NO];
#pragma unroll
    for (int k = 0; k < NO; ++k) acc[k] = 0.0f;
    float carry = 1.0f;
    for (int c = 0; c < steps; c += 32) {
        const int j = c + lane;
        const bool ok = j < steps;
        const int s = base + j;
        float a = 0.0f, nx = 0.f, ny = 0.f, nz = 0.f, mid = 0.f, cosv = 0.f;
        if (ok) {
            nx = vec[3 * s]; ny = vec[3 * s + 1]; nz = vec[3 * s + 2];
            if (NORMALIZE) {
                const float inv_n = 1.0f / fmaxf(sqrtf(nx * nx + ny * ny + nz * nz), eps);
                nx *= inv_n; ny *= inv_n; nz *= inv_n;
            }
            const float t0 = t_starts[s], t1 = t_ends[s];
            mid = (t0 + t1) / 2.0f;
            cosv = dx * nx + dy * ny + dz * nz;
            a = neus_alpha(sdf[s], cosv, t1 - t0, inv_s, ratio).alpha;
        }
        const float incl = warp_incl_prod(1.0f - a, lane);
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0f;
        const float T = carry * excl;
        const float w = a * T;
        carry *= __shfl_sync(0xffffffffu, incl, 31);
        if (ok) {
            alpha_out[s] = a;
            w_out[s] = w;
            T_out[s] = T;
#pragma unroll
            for (int k = 0; k < CD; ++k) acc[k] = fmaf(w, rgb[(size_t)CD * s + k], acc[k]);
            acc[CD] = fmaf(w, nx, acc[CD]);
            acc[CD + 1] = fmaf(w, ny, acc[CD + 1]);
            acc[CD + 2] = fmaf(w, nz, acc[CD + 2]);
            acc[CD + 3] += w;
            acc[CD + 4] = fmaf(w, mid, acc[CD + 4]);
            if (ORIENT) acc[CD + 5] = fmaf(w, fmaxf(cosv, 0.0f), acc[CD + 5]);
        }
    }
#pragma unroll
    for (int k = 0; k < NO; ++k) {
        const float s = warp_sum(acc[k]);
        if (lane == 0) out[(size_t)ray * NO + k] = s;
    }
}

template <int CD, bool NORMALIZE, bool ORIENT>
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK)
sdf_render_bwd_kernel(const int32_t *__restrict__ packed, const float *__restrict__ rays_d,
                      const float *__restrict__ t_starts, const float *__restrict__ t_ends,
                      const float *__restrict__ sdf, const float *__restrict__ vec,
                      const float *__restrict__ rgb, const float *__restrict__ alpha_in,
                      const float *__restrict__ w_in, const float *__restrict__ T_in,
                      const float *__restrict__ go,
                      const float *__restrict__ gw_extra, const float *__restrict__ inv_s_ptr,
                      float ratio, float eps, int n_rays, float *__restrict__ g_sdf,
                      float *__restrict__ g_vec, float *__restrict__ g_rgb,
                      float *__restrict__ g_inv_s_ray) {
    constexpr int NO = CD + 5 + (ORIENT ? 1 : 0);
    const int ray = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (ray >= n_rays) return;
    const int base = packed[2 * ray], steps = packed[2 * ray + 1];
    const float inv_s_raw = __ldg(inv_s_ptr);
    const float inv_s = fminf(fmaxf(inv_s_raw, 1e-6f), 1e6f);
    const bool s_live = inv_s_raw >= 1e-6f && inv_s_raw <= 1e6f;
    const float dx = rays_d[3 * ray], dy = rays_d[3 * ray + 1], dz = rays_d[3 * ray + 2];
    float g[NO];
#pragma unroll
    for (int k = 0; k < NO; ++k) g[k] = go[(size_t)ray * NO + k];
    float carry = 0.0f, ginv = 0.0f;
    const int nchunks = (steps + 31) >> 5;
    for (int ch = nchunks - 1; ch >= 0; --ch) {
        const int j = ch * 32 + lane;
        const bool ok = j < steps;
        const int s = base + j;
        float a = 0.f, w = 0.f, gwj = 0.f, T = 0.f;
        float nx = 0.f, ny = 0.f, nz = 0.f, inv_n = 1.f, dist = 0.f, sd = 0.f, cosv = 0.f, vlen = 1.f;
        if (ok) {
            a = alpha_in[s];
            w = w_in[s];
            nx = vec[3 * s]; ny = vec[3 * s + 1]; nz = vec[3 * s + 2];
            if (NORMALIZE) {
                vlen = sqrtf(nx * nx + ny * ny + nz * nz);
                inv_n = 1.0f / fmaxf(vlen, eps);
                nx *= inv_n; ny *= inv_n; nz *= inv_n;
            }
            const float t0 = t_starts[s], t1 = t_ends[s];
            dist = t1 - t0;
            sd = sdf[s];
            const float mid = (t0 + t1) / 2.0f;
            cosv = dx * nx + dy * ny + dz * nz;
            gwj = g[CD] * nx + g[CD + 1] * ny + g[CD + 2] * nz + g[CD + 3] + g[CD + 4] * mid;
#pragma unroll
            for (int k = 0; k < CD; ++k) gwj = fmaf(g[k], rgb[(size_t)CD * s + k], gwj);
            if (ORIENT) gwj = fmaf(g[CD + 5], fmaxf(cosv, 0.0f), gwj);
            if (gw_extra) gwj += gw_extra[s];
            T = T_in[s];
        }
        const float sw = warp_suffix_sum(gwj * w, lane);
        float ga = 0.0f;
        if (ok) ga = (gwj * T - (sw + carry)) / fmaxf(1.0f - a, 1e-10f);
        carry += __shfl_sync(0xffffffffu, sw, 0);
        if (ok) {
            // values' own grads
#pragma unroll
            for (int k = 0; k < CD; ++k) g_rgb[(size_t)CD * s + k] = w * g[k];
            float gnx = w * g[CD], gny = w * g[CD + 1], gnz = w * g[CD + 2];
            if (ORIENT && cosv > 0.0f) {
                const float go_ = w * g[CD + 5];
                gnx = fmaf(go_, dx, gnx); gny = fmaf(go_, dy, gny); gnz = fmaf(go_, dz, gnz);
            }
            // alpha backward (models/neus.py:133-150)
            const AlphaTerms t = neus_alpha(sd, cosv, dist, inv_s, ratio);
            float gsd = 0.0f;
            if (t.praw >= 0.0f && t.praw <= 1.0f) {
                const float den = 1.0f / (t.P + 1e-5f);
                const float dP = ga * (den - t.praw * den);
                const float dN = -ga * den;
                const float sP = dP * t.P * (1.0f - t.P), sN = dN * t.N * (1.0f - t.N);
                const float de_p = sP * inv_s, de_n = sN * inv_s;
                if (s_live) ginv += sP * t.e_p + sN * t.e_n;
                gsd = de_p + de_n;
                const float dic = (de_n - de_p) * dist * 0.5f;
                const float dcos = dic * (((-cosv * 0.5f + 0.5f) > 0.0f ? 0.5f * (1.0f - ratio) : 0.0f) +
                                          ((-cosv) > 0.0f ? ratio : 0.0f));
                gnx = fmaf(dcos, dx, gnx);
                gny = fmaf(dcos, dy, gny);
                gnz = fmaf(dcos, dz, gnz);
            }
            g_sdf[s] = gsd;
            if (NORMALIZE) {
                // normalize backward: n = v / max(|v|, eps)
                if (vlen > eps) {
                    const float ndot = nx * gnx + ny * gny + nz * gnz;
                    gnx = (gnx - nx * ndot) * inv_n; gny = (gny - ny * ndot) * inv_n; gnz = (gnz - nz * ndot) * inv_n;
                } else {
                    gnx *= inv_n; gny *= inv_n; gnz *= inv_n;
                }
            }
            g_vec[3 * s] = gnx; g_vec[3 * s + 1] = gny; g_vec[3 * s + 2] = gnz;
        }
    }
    ginv = warp_sum(ginv);
    if (lane == 0 && g_inv_s_ray) g_inv_s_ray[ray] = ginv;
}

}  // namespace

// ---- per-sample set-up and normal normalisation (the elementwise glue of models/neus.py:247-256) ----------
// positions = o[ray] + d[ray] * (t0 + t1) / 2  with the reference's operation order (add, divide by 2 == * 0.5
// exactly, multiply, add: no contraction), plus the gathered directions, midpoints and interval lengths.
__global__ void sample_setup_kernel(const float *__restrict__ rays_o, const float *__restrict__ rays_d,
                                    const long long *__restrict__ ray_indices, const float *__restrict__ t_starts,
                                    const float *__restrict__ t_ends, int n, float *__restrict__ positions,
                                    float *__restrict__ dirs, float *__restrict__ midpoints, float *__restrict__ dists) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long r = ray_indices[i];
    const float t0 = t_starts[i], t1 = t_ends[i];
    const float mid = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
    midpoints[i] = mid;
    dists[i] = __fsub_rn(t1, t0);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float o = __ldg(rays_o + 3 * r + d), v = __ldg(rays_d + 3 * r + d);
        dirs[3 * (size_t)i + d] = v;
        positions[3 * (size_t)i + d] = __fadd_rn(o, __fmul_rn(v, mid));
    }
}

// F.normalize(g, p=2, dim=-1, eps): n = g / max(|g|, eps)
__global__ void normalize3_fwd_kernel(const float *__restrict__ g, int n, float eps, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = g[3 * (size_t)i], y = g[3 * (size_t)i + 1], z = g[3 * (size_t)i + 2];
    const float inv = 1.0f / fmaxf(sqrtf(x * x + y * y + z * z), eps);
    out[3 * (size_t)i] = x * inv; out[3 * (size_t)i + 1] = y * inv; out[3 * (size_t)i + 2] = z * inv;
}
// backward: |g| > eps: (gn - n (n . gn)) / |g|;  otherwise gn / eps
__global__ void normalize3_bwd_kernel(const float *__restrict__ g, const float *__restrict__ gn, int n, float eps,
                                      float *__restrict__ gg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = g[3 * (size_t)i], y = g[3 * (size_t)i + 1], z = g[3 * (size_t)i + 2];
    const float a = gn[3 * (size_t)i], b = gn[3 * (size_t)i + 1], c = gn[3 * (size_t)i + 2];
    const float len = sqrtf(x * x + y * y + z * z);
    if (len > eps) {
        const float inv = 1.0f / len;
        const float nx = x * inv, ny = y * inv, nz = z * inv;
        const float dot = nx * a + ny * b + nz * c;
        gg[3 * (size_t)i] = (a - nx * dot) * inv; gg[3 * (size_t)i + 1] = (b - ny * dot) * inv;
        gg[3 * (size_t)i + 2] = (c - nz * dot) * inv;
    } else {
        const float inv = 1.0f / eps;
        gg[3 * (size_t)i] = a * inv; gg[3 * (size_t)i + 1] = b * inv; gg[3 * (size_t)i + 2] = c * inv;
    }
}

// ---- eikonal + sparsity regularisers (systems/neus.py:117-131): sums over the samples in one pass ----------
//   out[0] += sum_i (|g_i| - 1)^2        out[1] += sum_i exp(-scale * |sdf_i|)
__global__ void sdf_reg_fwd_kernel(const float *__restrict__ g, const float *__restrict__ sdf, int n, float scale,
                                   float *__restrict__ out) {
    float e = 0.f, sp = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = g[3 * (size_t)i], y = g[3 * (size_t)i + 1], z = g[3 * (size_t)i + 2];
        const float d = sqrtf(x * x + y * y + z * z) - 1.0f;
        e = fmaf(d, d, e);
        sp += expf(-scale * fabsf(sdf[i]));
    }
    e = warp_sum(e); sp = warp_sum(sp);
    __shared__ float se[32], ss[32];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { se[w] = e; ss[w] = sp; }
    __syncthreads();
    if (w == 0) {
        e = l < (blockDim.x >> 5) ? se[l] : 0.f; sp = l < (blockDim.x >> 5) ? ss[l] : 0.f;
        e = warp_sum(e); sp = warp_sum(sp);
        if (l == 0) { out[2 * blockIdx.x] = e; out[2 * blockIdx.x + 1] = sp; }     // per-block partials, no atomics
    }
}
// second stage: one block adds the per-block partials in a fixed order -> bit-reproducible loss terms
__global__ void sdf_reg_finish_kernel(const float *__restrict__ partials, int n_blocks, float *__restrict__ out) {
    float e = 0.f, sp = 0.f;
    for (int i = threadIdx.x; i < n_blocks; i += blockDim.x) { e += partials[2 * i]; sp += partials[2 * i + 1]; }
    e = warp_sum(e); sp = warp_sum(sp);
    __shared__ float se[32], ss[32];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { se[w] = e; ss[w] = sp; }
    __syncthreads();
    if (w == 0) {
        e = l < (blockDim.x >> 5) ? se[l] : 0.f; sp = l < (blockDim.x >> 5) ? ss[l] : 0.f;
        e = warp_sum(e); sp = warp_sum(sp);
        if (l == 0) { out[0] = e; out[1] = sp; }
    }
}
// cotangents ce, cs (device scalars, already divided by n by the caller's mean):
//   gg_i = ce * 2 (|g_i| - 1) g_i / |g_i|      gs_i = -cs * scale * sign(sdf_i) exp(-scale |sdf_i|)
__global__ void sdf_reg_bwd_kernel(const float *__restrict__ g, const float *__restrict__ sdf, int n, float scale,
                                   const float *__restrict__ cot, float *__restrict__ gg, float *__restrict__ gs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float ce = cot[0], cs = cot[1];
    const float x = g[3 * (size_t)i], y = g[3 * (size_t)i + 1], z = g[3 * (size_t)i + 2];
    const float len = sqrtf(x * x + y * y + z * z);
    const float k = len > 0.0f ? ce * 2.0f * (len - 1.0f) / len : 0.0f;
    gg[3 * (size_t)i] = k * x; gg[3 * (size_t)i + 1] = k * y; gg[3 * (size_t)i + 2] = k * z;
    const float v = sdf[i];
    const float sgn = v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f);
    gs[i] = -cs * scale * sgn * expf(-scale * fabsf(v));
}

extern "C" {

int rsdf_sdf_reg_fwd(const float *sdf_grad, const float *sdf, int n, float sparsity_scale, float *out2,
                     float *partials, void *stream) {
    if (!out2 || !partials) return RSDF_EBADARG;
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(out2, 0, 2 * sizeof(float), (cudaStream_t)stream);
        return e == cudaSuccess ? 0 : (int)e;
    }
    if (!sdf_grad || !sdf) return RSDF_EBADARG;
    const int blocks = rsdf_div_up(n, 256) < RSDF_SDF_REG_BLOCKS ? rsdf_div_up(n, 256) : RSDF_SDF_REG_BLOCKS;
    sdf_reg_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(sdf_grad, sdf, n, sparsity_scale, partials);
    sdf_reg_finish_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partials, blocks, out2);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_sdf_reg_bwd(const float *sdf_grad, const float *sdf, int n, float sparsity_scale, const float *cot2,
                     float *grad_sdf_grad, float *grad_sdf, void *stream) {
    if (n == 0) return 0;
    if (!sdf_grad || !sdf || !cot2 || !grad_sdf_grad || !grad_sdf) return RSDF_EBADARG;
    sdf_reg_bwd_kernel<<<rsdf_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(sdf_grad, sdf, n, sparsity_scale, cot2,
                                                                              grad_sdf_grad, grad_sdf);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_sample_setup(const float *rays_o, const float *rays_d, const long long *ray_indices, const float *t_starts,
                      const float *t_ends, int n_samples, float *positions, float *dirs, float *midpoints,
                      float *dists, void *stream) {
    if (n_samples == 0) return 0;
    if (!rays_o || !rays_d || !ray_indices || !t_starts || !t_ends || !positions || !dirs || !midpoints || !dists)
        return RSDF_EBADARG;
    sample_setup_kernel<<<rsdf_div_up(n_samples, 256), 256, 0, (cudaStream_t)stream>>>(
        rays_o, rays_d, ray_indices, t_starts, t_ends, n_samples, positions, dirs, midpoints, dists);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_normalize3_fwd(const float *g, int n, float eps, float *out, void *stream) {
    if (n == 0) return 0;
    if (!g || !out) return RSDF_EBADARG;
    normalize3_fwd_kernel<<<rsdf_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(g, n, eps, out);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_normalize3_bwd(const float *g, const float *grad_out, int n, float eps, float *grad_g, void *stream) {
    if (n == 0) return 0;
    if (!g || !grad_out || !grad_g) return RSDF_EBADARG;
    normalize3_bwd_kernel<<<rsdf_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(g, grad_out, n, eps, grad_g);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_weight_from_alpha_fwd(const int32_t *packed_info, const float *alphas, int n_rays,
                               float *weights, float *trans, void *stream) {
    if (n_rays == 0) return 0;
    if (!packed_info || !alphas) return RSDF_EBADARG;
    weight_fwd_kernel<<<rsdf_div_up(n_rays, WARPS_PER_BLOCK), 32 * WARPS_PER_BLOCK, 0,
                        (cudaStream_t)stream>>>(packed_info, alphas, n_rays, weights, trans);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_weight_from_alpha_bwd(const int32_t *packed_info, const float *alphas,
                               const float *weights, const float *trans,
                               const float *grad_weights, const float *grad_trans, int n_rays,
                               float *grad_alphas, void *stream) {
    if (n_rays == 0) return 0;
    if (!packed_info || !alphas || !trans || !grad_alphas) return RSDF_EBADARG;
    weight_bwd_kernel<<<rsdf_div_up(n_rays, WARPS_PER_BLOCK), 32 * WARPS_PER_BLOCK, 0,
                        (cudaStream_t)stream>>>(packed_info, alphas, weights, trans, grad_weights,
                                                grad_trans, n_rays, grad_alphas);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_accumulate_fwd(const int32_t *packed_info, const float *weights, const float *values,
                        int n_rays, int D, float *out, void *stream) {
    if (n_rays == 0) return 0;
    if (!packed_info || !weights || !out || D < 1 || (!values && D != 1)) return RSDF_EBADARG;
    accumulate_fwd_kernel<<<rsdf_div_up(n_rays, WARPS_PER_BLOCK), 32 * WARPS_PER_BLOCK, 0,
                            (cudaStream_t)stream>>>(packed_info, weights, values, n_rays, D, out);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_accumulate_bwd(const int64_t *ray_indices, const float *weights, const float *values,
                        const float *grad_out, int n_samples, int D, float *grad_weights,
                        float *grad_values, void *stream) {
    if (n_samples == 0) return 0;
    if (!ray_indices || !grad_out || D < 1 || (values && !weights)) return RSDF_EBADARG;
    accumulate_bwd_kernel<<<rsdf_div_up(n_samples, 256), 256, 0, (cudaStream_t)stream>>>(
        ray_indices, weights, values, grad_out, n_samples, D, grad_weights, grad_values);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_neus_render_fwd(const int32_t *packed_info, const float *rays_d, const float *t_starts,
                         const float *t_ends, const float *sdf, const float *sdf_grad,
                         const float *rgb, const float *inv_s, float cos_anneal_ratio, int n_rays,
                         float *alpha, float *weights, float *trans, float *out, void *stream) {
    if (n_rays == 0) return 0;
    if (!packed_info || !rays_d || !inv_s || !out || !alpha || !weights || !trans)
        return RSDF_EBADARG;
    sdf_render_fwd_kernel<3, true, false><<<rsdf_div_up(n_rays, WARPS_PER_BLOCK), 32 * WARPS_PER_BLOCK, 0,
                                            (cudaStream_t)stream>>>(packed_info, rays_d, t_starts, t_ends, sdf,
                                                                    sdf_grad, rgb, inv_s, cos_anneal_ratio, 1e-12f,
                                                                    n_rays, alpha, weights, trans, out);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_neus_render_bwd(const int32_t *packed_info, const float *rays_d, const float *t_starts,
                         const float *t_ends, const float *sdf, const float *sdf_grad,
                         const float *rgb, const float *alpha, const float *weights,
                         const float *trans, const float *grad_out,
                         const float *grad_weights_extra,
                         const float *inv_s, float cos_anneal_ratio, int n_rays, float *grad_sdf,
                         float *grad_sdf_grad, float *grad_rgb, float *grad_inv_s_per_ray,
                         void *stream) {
    if (n_rays == 0) return 0;
    if (!packed_info || !rays_d || !inv_s || !grad_out || !trans) return RSDF_EBADARG;
    sdf_render_bwd_kernel<3, true, false><<<rsdf_div_up(n_rays, WARPS_PER_BLOCK), 32 * WARPS_PER_BLOCK, 0,
                                            (cudaStream_t)stream>>>(
        packed_info, rays_d, t_starts, t_ends, sdf, sdf_grad, rgb, alpha, weights, trans, grad_out,
        grad_weights_extra, inv_s, cos_anneal_ratio, 1e-12f, n_rays, grad_sdf, grad_sdf_grad, grad_rgb,
        grad_inv_s_per_ray);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_split_render_fwd(const int32_t *packed_info, const float *rays_d, const float *t_starts, const float *t_ends,
                          const float *sdf, const float *normals, const float *colors, int color_dim, const float *inv_s,
                          float cos_anneal_ratio, int n_rays, float *alpha, float *weights, float *trans, float *out,
                          void *stream) {
    if (n_rays == 0) return 0;
    if (!packed_info || !rays_d || !inv_s || !out || !alpha || !weights || !trans) return RSDF_EBADARG;
    const int blocks = rsdf_div_up(n_rays, WARPS_PER_BLOCK), threads = 32 * WARPS_PER_BLOCK;
    cudaStream_t st = (cudaStream_t)stream;
#define RSDF_SR(CD)                                                                                                   \
    sdf_render_fwd_kernel<CD, false, true><<<blocks, threads, 0, st>>>(packed_info, rays_d, t_starts, t_ends, sdf,    \
                                                                       normals, colors, inv_s, cos_anneal_ratio, 0.f, \
                                                                       n_rays, alpha, weights, trans, out)
    if (color_dim == 24) RSDF_SR(24);
    else if (color_dim == 7) RSDF_SR(7);
    else return RSDF_EBADARG;
#undef RSDF_SR
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_split_render_bwd(const int32_t *packed_info, const float *rays_d, const float *t_starts, const float *t_ends,
                          const float *sdf, const float *normals, const float *colors, int color_dim, const float *alpha,
                          const float *weights, const float *trans, const float *grad_out,
                          const float *grad_weights_extra, const float *inv_s, float cos_anneal_ratio, int n_rays,
                          float *grad_sdf, float *grad_normals, float *grad_colors, float *grad_inv_s_per_ray,
                          void *stream) {
    if (n_rays == 0) return 0;
    if (!packed_info || !rays_d || !inv_s || !grad_out || !trans) return RSDF_EBADARG;
    const int blocks = rsdf_div_up(n_rays, WARPS_PER_BLOCK), threads = 32 * WARPS_PER_BLOCK;
    cudaStream_t st = (cudaStream_t)stream;
#define RSDF_SR(CD)                                                                                                     \
    sdf_render_bwd_kernel<CD, false, true><<<blocks, threads, 0, st>>>(                                                  \
        packed_info, rays_d, t_starts, t_ends, sdf, normals, colors, alpha, weights, trans, grad_out,                  \
        grad_weights_extra, inv_s, cos_anneal_ratio, 0.f, n_rays, grad_sdf, grad_normals, grad_colors, grad_inv_s_per_ray)
    if (color_dim == 24) RSDF_SR(24);
    else if (color_dim == 7) RSDF_SR(7);
    else return RSDF_EBADARG;
#undef RSDF_SR
    RSDF_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
