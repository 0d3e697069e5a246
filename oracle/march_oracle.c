/*
 * oracle/march_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked by the product).
 *
 * Plain-C, single-thread restatement of the reference's occupancy-grid ray march and
 * per-ray transmittance scan, written from the algorithm in
 *   /root/reference/lib/nerfacc/cuda/csrc/ray_marching.cu:9-192   (march, DDA skip)
 *   /root/reference/lib/nerfacc/cuda/csrc/intersection.cu:16-91   (ray/AABB slab test)
 *   /root/reference/lib/nerfacc/cuda/csrc/include/helpers_contraction.h:16-21 (roi_to_unit)
 *   /root/reference/lib/nerfacc/cuda/csrc/render_weight.cu:86-154 (w = alpha*T, backward)
 *   /root/reference/lib/nerfacc/cuda/csrc/render_transmittance.cu:85-145
 *
 * Floating point: the reference is compiled by nvcc with the default --fmad=true, so
 * `origin + t_mid*dir` (ray_marching.cu:148) is a fused multiply-add on the device.  Every
 * other expression on the path is either a lone add/mul/div (correctly rounded on both
 * sides) or a multiply by +-0.5 (exact), so it is contraction-invariant.  This file is
 * built with -ffp-contract=off and spells the one fused op with fmaf().
 *
 * Parity pin: tests/golden/march_*.npz are outputs of the reference's own kernel
 * (oracle/_ref/nerfacc_cuda.so, built from the reference sources) run on a B200; this
 * restatement is checked against them bit-for-bit in tests/test_oracle_march.py.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* helpers_math.h:1167-1170: clamp(f,a,b) = fmaxf(a, fminf(f,b)) */
static inline float clampf(float x, float lo, float hi) { return fmaxf(lo, fminf(x, hi)); }
static inline float signf_(float x) { return copysignf(1.0f, x); }

/* intersection.cu:16-91 : near/far of ray vs aabb, miss -> (1e10,1e10), near clamped >= 0 */
void oracle_ray_aabb_intersect(int n, const float *o, const float *d, const float *aabb,
                               float *t_min, float *t_max) {
    for (int i = 0; i < n; ++i) {
        const float *ro = o + 3 * i, *rd = d + 3 * i;
        float nearv, farv;
        float tmin = (aabb[0] - ro[0]) / rd[0];
        float tmax = (aabb[3] - ro[0]) / rd[0];
        int miss = 0;
        if (tmin > tmax) { float c = tmin; tmin = tmax; tmax = c; }
        float tymin = (aabb[1] - ro[1]) / rd[1];
        float tymax = (aabb[4] - ro[1]) / rd[1];
        if (tymin > tymax) { float c = tymin; tymin = tymax; tymax = c; }
        if (tmin > tymax || tymin > tmax) miss = 1;
        if (!miss) {
            if (tymin > tmin) tmin = tymin;
            if (tymax < tmax) tmax = tymax;
            float tzmin = (aabb[2] - ro[2]) / rd[2];
            float tzmax = (aabb[5] - ro[2]) / rd[2];
            if (tzmin > tzmax) { float c = tzmin; tzmin = tzmax; tzmax = c; }
            if (tmin > tzmax || tzmin > tmax) miss = 1;
            if (!miss) {
                if (tzmin > tmin) tmin = tzmin;
                if (tzmax < tmax) tmax = tzmax;
            }
        }
        if (miss) { nearv = 1e10f; farv = 1e10f; } else { nearv = tmin; farv = tmax; }
        t_min[i] = nearv > 0.0f ? nearv : 0.0f;
        t_max[i] = farv;
    }
}

/* ray_marching.cu:16-45 (AABB contraction only) */
static inline int occupied_at(const float x[3], const float *roi, const int res[3],
                              const uint8_t *grid) {
    for (int k = 0; k < 3; ++k)
        if (x[k] < roi[k] || x[k] > roi[3 + k]) return 0;
    int ixyz[3];
    for (int k = 0; k < 3; ++k) {
        float u = (x[k] - roi[k]) / (roi[3 + k] - roi[k]);
        int iv = (int)(u * (float)res[k]);
        iv = iv < 0 ? 0 : (iv > res[k] - 1 ? res[k] - 1 : iv);
        ixyz[k] = iv;
    }
    int idx = ixyz[0] * res[1] * res[2] + ixyz[1] * res[2] + ixyz[2];
    return grid[idx] != 0;
}

/* ray_marching.cu:48-75 */
static inline float advance_to_next_voxel(float t, float dt_min, const float x[3],
                                          const float dir[3], const float inv_dir[3],
                                          const float *roi, const int res[3], float far) {
    float tx[3];
    for (int k = 0; k < 3; ++k) {
        float r = (float)res[k];
        float u = ((x[k] - roi[k]) / (roi[3 + k] - roi[k])) * r;
        float f = floorf(u + 0.5f + 0.5f * signf_(dir[k]));
        tx[k] = (((f - u) * inv_dir[k]) / r) * (roi[3 + k] - roi[k]);
    }
    float tt = fminf(fminf(tx[0], tx[1]), tx[2]);
    float t_target = t + fmaxf(tt, 0.0f);
    t_target = fminf(t_target, far);
    float _t = t;
    do { _t += dt_min; } while (_t < t_target);
    return _t;
}

/* ray_marching.cu:81-192.  pass 0: counts only (num_steps[n]); pass 1: fill using
 * packed_info (base, count).  Returns total steps of the rays processed. */
int64_t oracle_ray_marching(int n_rays, const float *rays_o, const float *rays_d,
                            const float *t_min, const float *t_max, const float *roi,
                            const int *res, const uint8_t *grid, float step_size,
                            float cone_angle, const int *packed_info, int *num_steps,
                            int64_t *ray_indices, float *t_starts, float *t_ends) {
    int64_t total = 0;
    for (int i = 0; i < n_rays; ++i) {
        const float *o = rays_o + 3 * i, *d = rays_d + 3 * i;
        float inv_dir[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
        const float near = t_min[i], far = t_max[i];
        const float dt_min = step_size, dt_max = 1e10f;
        int base = packed_info ? packed_info[2 * i] : 0;
        int j = 0;
        float t0 = near;
        float dt = clampf(t0 * cone_angle, dt_min, dt_max);
        float t1 = t0 + dt;
        float t_mid = (t0 + t1) * 0.5f;
        while (t_mid < far) {
            float x[3];
            for (int k = 0; k < 3; ++k) x[k] = fmaf(t_mid, d[k], o[k]);
            if (occupied_at(x, roi, res, grid)) {
                if (packed_info) {
                    t_starts[base + j] = t0;
                    t_ends[base + j] = t1;
                    ray_indices[base + j] = i;
                }
                ++j;
                t0 = t1;
                t1 = t0 + clampf(t0 * cone_angle, dt_min, dt_max);
                t_mid = (t0 + t1) * 0.5f;
            } else {
                t_mid = advance_to_next_voxel(t_mid, dt_min, x, d, inv_dir, roi, res, far);
                dt = clampf(t_mid * cone_angle, dt_min, dt_max);
                t0 = t_mid - dt * 0.5f;
                t1 = t_mid + dt * 0.5f;
            }
        }
        if (!packed_info) num_steps[i] = j;
        total += j;
    }
    return total;
}

/* ray_marching.cu:295-320 */
void oracle_grid_query(int n, const float *samples, const float *roi, const int *res,
                       const uint8_t *grid, uint8_t *out) {
    for (int i = 0; i < n; ++i) out[i] = (uint8_t)occupied_at(samples + 3 * i, roi, res, grid);
}

/* render_weight.cu:86-111 + render_transmittance.cu:85-112 (serial order per ray) */
void oracle_weight_from_alpha_forward(int n_rays, const int *packed_info, const float *alphas,
                                      float *weights, float *trans) {
    for (int i = 0; i < n_rays; ++i) {
        int base = packed_info[2 * i], steps = packed_info[2 * i + 1];
        float T = 1.0f;
        for (int j = 0; j < steps; ++j) {
            float a = alphas[base + j];
            if (weights) weights[base + j] = a * T;
            if (trans) trans[base + j] = T;
            T *= (1.0f - a);
        }
    }
}

/* render_weight.cu:113-154 */
void oracle_weight_from_alpha_backward(int n_rays, const int *packed_info, const float *alphas,
                                       const float *weights, const float *grad_weights,
                                       float *grad_alphas) {
    for (int i = 0; i < n_rays; ++i) {
        int base = packed_info[2 * i], steps = packed_info[2 * i + 1];
        float accum = 0.0f;
        for (int j = 0; j < steps; ++j) accum += grad_weights[base + j] * weights[base + j];
        float T = 1.0f;
        for (int j = 0; j < steps; ++j) {
            float a = alphas[base + j];
            grad_alphas[base + j] = (grad_weights[base + j] * T - accum) / fmaxf(1.0f - a, 1e-10f);
            accum -= grad_weights[base + j] * weights[base + j];
            T *= (1.0f - a);
        }
    }
}

/* render_transmittance.cu:114-145 */
void oracle_transmittance_from_alpha_backward(int n_rays, const int *packed_info,
                                              const float *alphas, const float *trans,
                                              const float *trans_grad, float *alphas_grad) {
    for (int i = 0; i < n_rays; ++i) {
        int base = packed_info[2 * i], steps = packed_info[2 * i + 1];
        float cumsum = 0.0f;
        for (int j = steps - 1; j >= 0; --j) {
            alphas_grad[base + j] = cumsum / fmaxf(1.0f - alphas[base + j], 1e-10f);
            cumsum += -trans_grad[base + j] * trans[base + j];
        }
    }
}
