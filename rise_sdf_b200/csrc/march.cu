// K1 -- occupancy-grid ray marching + sample compaction (integer/index path, bit-exact).
//
// Replaces lib/nerfacc/cuda/csrc/ray_marching.cu:81-289 + intersection.cu:69-145.
// Same two-round structure (count, prefix, fill) because the packed, ray-ordered layout is
// part of the contract, but: the prefix sum runs on the device in the same stream (no
// cumsum/stack/.item() on the host), the grid can be read bit-packed (256 KiB for 128^3, so
// it stays in L1/L2), and blocks are small so that 8192 rays still cover all 148 SMs.
//
// Floating-point contract: every operation below is spelled with a
// round-to-nearest intrinsic so the compiler can neither add nor drop a contraction; the one
// FMA the reference's build performs (origin + t_mid*dir under nvcc's default --fmad=true)
// is written as __fmaf_rn.
#include "common.cuh"

namespace {

struct MarchGrid {
    float roi[6];
    int rx, ry, rz;
    const uint8_t *cells;   // bool grid, or
    const uint32_t *bits;   // bit-packed grid (preferred)
};

__device__ __forceinline__ float clamp_ref(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }

__device__ __forceinline__ bool occupied_at(const float x, const float y, const float z,
                                            const MarchGrid &g) {
    if (x < g.roi[0] || x > g.roi[3] || y < g.roi[1] || y > g.roi[4] || z < g.roi[2] || z > g.roi[5])
        return false;
    const float ux = __fdiv_rn(__fsub_rn(x, g.roi[0]), __fsub_rn(g.roi[3], g.roi[0]));
    const float uy = __fdiv_rn(__fsub_rn(y, g.roi[1]), __fsub_rn(g.roi[4], g.roi[1]));
    const float uz = __fdiv_rn(__fsub_rn(z, g.roi[2]), __fsub_rn(g.roi[5], g.roi[2]));
    int ix = (int)__fmul_rn(ux, (float)g.rx);
    int iy = (int)__fmul_rn(uy, (float)g.ry);
    int iz = (int)__fmul_rn(uz, (float)g.rz);
    ix = min(max(ix, 0), g.rx - 1);
    iy = min(max(iy, 0), g.ry - 1);
    iz = min(max(iz, 0), g.rz - 1);
    const int idx = ix * g.ry * g.rz + iy * g.rz + iz;
    if (g.bits) return (__ldg(g.bits + (idx >> 5)) >> (idx & 31)) & 1u;
    return __ldg(g.cells + idx) != 0;
}

__device__ __forceinline__ float axis_exit(float x, float lo, float hi, float r, float dir,
                                           float inv_dir) {
    const float u = __fmul_rn(__fdiv_rn(__fsub_rn(x, lo), __fsub_rn(hi, lo)), r);
    const float f = floorf(__fadd_rn(__fadd_rn(u, 0.5f), __fmul_rn(0.5f, copysignf(1.0f, dir))));
    return __fmul_rn(__fdiv_rn(__fmul_rn(__fsub_rn(f, u), inv_dir), r), __fsub_rn(hi, lo));
}

// MODE 0: count; 1: fill (second march); 2: count and keep the first `cap` intervals of every ray in
// keep[ray][cap] so that the fill round is a copy (compact_kernel) instead of a second march.
template <int MODE>
__global__ void __launch_bounds__(64)
march_kernel(const int n_rays, const float *__restrict__ rays_o, const float *__restrict__ rays_d,
             const float *__restrict__ t_min, const float *__restrict__ t_max, const MarchGrid g,
             const float step_size, const float cone_angle,
             const int32_t *__restrict__ packed_info, int32_t *__restrict__ num_steps,
             int64_t *__restrict__ ray_indices, float *__restrict__ t_starts,
             float *__restrict__ t_ends, float2 *__restrict__ keep = nullptr, const int cap = 0,
             int32_t *__restrict__ overflow = nullptr) {
    constexpr bool FILL = MODE == 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays) return;
    const float ox = rays_o[3 * i], oy = rays_o[3 * i + 1], oz = rays_o[3 * i + 2];
    const float dx = rays_d[3 * i], dy = rays_d[3 * i + 1], dz = rays_d[3 * i + 2];
    const float ix = __fdiv_rn(1.0f, dx), iy = __fdiv_rn(1.0f, dy), iz = __fdiv_rn(1.0f, dz);
    const float near = t_min[i], far = t_max[i];
    const float dt_min = step_size, dt_max = 1e10f;
    int base = 0;
    if (FILL) base = packed_info[2 * i];

    int j = 0;
    float t0 = near;
    float dt = clamp_ref(__fmul_rn(t0, cone_angle), dt_min, dt_max);
    float t1 = __fadd_rn(t0, dt);
    float t_mid = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
    while (t_mid < far) {
        const float x = __fmaf_rn(t_mid, dx, ox);
        const float y = __fmaf_rn(t_mid, dy, oy);
        const float z = __fmaf_rn(t_mid, dz, oz);
        if (occupied_at(x, y, z, g)) {
            if (FILL) {
                t_starts[base + j] = t0;
                t_ends[base + j] = t1;
                ray_indices[base + j] = i;
            }
            if (MODE == 2 && j < cap) __stcs(keep + (size_t)i * cap + j, make_float2(t0, t1));
            ++j;
            t0 = t1;
            t1 = __fadd_rn(t0, clamp_ref(__fmul_rn(t0, cone_angle), dt_min, dt_max));
            t_mid = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
        } else {
            const float tx = axis_exit(x, g.roi[0], g.roi[3], (float)g.rx, dx, ix);
            const float ty = axis_exit(y, g.roi[1], g.roi[4], (float)g.ry, dy, iy);
            const float tz = axis_exit(z, g.roi[2], g.roi[5], (float)g.rz, dz, iz);
            const float tt = fmaxf(fminf(fminf(tx, ty), tz), 0.0f);
            const float t_target = fminf(__fadd_rn(t_mid, tt), far);
            float _t = t_mid;
            do { _t = __fadd_rn(_t, dt_min); } while (_t < t_target);
            t_mid = _t;
            dt = clamp_ref(__fmul_rn(t_mid, cone_angle), dt_min, dt_max);
            t0 = __fsub_rn(t_mid, __fmul_rn(dt, 0.5f));
            t1 = __fadd_rn(t_mid, __fmul_rn(dt, 0.5f));
        }
    }
    if (!FILL) num_steps[i] = j;
    if (MODE == 2 && j > cap) *overflow = 1;
}

// Count round that keeps its intervals, ONE WARP PER RAY (the thread-per-ray kernel above runs 8192 rays as 8192
// threads: 3 % of the warp slots, every iteration a dependent chain of IEEE divisions).  Inside an occupied run the
// march is a pure recurrence (t0 <- t1, t1 <- t0 + dt), so lane k evaluates the state after k occupied steps: the
// same fp32 additions in the same order as the sequential loop, hence the same bits.  A ballot finds the first lane
// whose sample is empty or beyond t_max; the lanes before it emit their intervals (coalesced), and the skip to the
// next voxel -- the only step whose outcome the following state depends on -- is done once, warp-uniformly.
// While a ray crosses empty space the window shrinks to one lane (no speculation to throw away).
template <bool CONE0>
__global__ void __launch_bounds__(128)
march_warp_kernel(const int n_rays, const float *__restrict__ rays_o, const float *__restrict__ rays_d,
                  const float *__restrict__ t_min, const float *__restrict__ t_max, const MarchGrid g,
                  const float step_size, const float cone_angle, int32_t *__restrict__ num_steps,
                  float2 *__restrict__ keep, const int cap, int32_t *__restrict__ overflow) {
    const int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (ray >= n_rays) return;
    const float ox = rays_o[3 * ray], oy = rays_o[3 * ray + 1], oz = rays_o[3 * ray + 2];
    const float dx = rays_d[3 * ray], dy = rays_d[3 * ray + 1], dz = rays_d[3 * ray + 2];
    const float ix = __fdiv_rn(1.0f, dx), iy = __fdiv_rn(1.0f, dy), iz = __fdiv_rn(1.0f, dz);
    const float near = t_min[ray], far = t_max[ray];
    const float dt_min = step_size, dt_max = 1e10f;
    float2 *my_keep = keep + (size_t)ray * cap;

    int j = 0;
    int W = 32;                                            // speculation window: 32 lanes, or 1 while crossing empty space
    float t0 = near;
    float t1 = __fadd_rn(t0, clamp_ref(__fmul_rn(t0, cone_angle), dt_min, dt_max));
    float t_mid = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
    while (t_mid < far) {                                  // warp-uniform state (t0, t1, t_mid, j, W)
        float a0 = t0, a1 = t1;                            // lane k: the state after k occupied steps
#pragma unroll 4
        for (int k = 0; k < W - 1; ++k) {
            const float n1 = __fadd_rn(a1, CONE0 ? dt_min : clamp_ref(__fmul_rn(a1, cone_angle), dt_min, dt_max));
            if (k < lane) { a0 = a1; a1 = n1; }
        }
        const float am = lane == 0 ? t_mid : __fmul_rn(__fadd_rn(a0, a1), 0.5f);
        const bool valid = lane < W && am < far;
        float x = 0.f, y = 0.f, z = 0.f;
        bool occ = false;
        if (valid) {
            x = __fmaf_rn(am, dx, ox);
            y = __fmaf_rn(am, dy, oy);
            z = __fmaf_rn(am, dz, oz);
            occ = occupied_at(x, y, z, g);
        }
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        const unsigned stop = ~__ballot_sync(0xffffffffu, occ);
        const int f = min(stop ? __ffs(stop) - 1 : 32, W);   // lanes < f continue the occupied run
        if (lane < f && j + lane < cap) __stcs(my_keep + j + lane, make_float2(a0, a1));
        j += f;
        if (f == W) {                                      // whole window occupied: carry on from its last lane's successor
            t0 = __shfl_sync(0xffffffffu, a1, W - 1);
            t1 = __fadd_rn(t0, clamp_ref(__fmul_rn(t0, cone_angle), dt_min, dt_max));
            t_mid = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
            W = 32;
            continue;
        }
        if (!((vmask >> f) & 1u)) break;                   // lane f is beyond t_max: the ray is done
        // lane f sits in an empty cell: advance to the next voxel exactly like the sequential loop
        const float fm = __shfl_sync(0xffffffffu, am, f);
        const float fx = __shfl_sync(0xffffffffu, x, f), fy = __shfl_sync(0xffffffffu, y, f),
                    fz = __shfl_sync(0xffffffffu, z, f);
        const float tx = axis_exit(fx, g.roi[0], g.roi[3], (float)g.rx, dx, ix);
        const float ty = axis_exit(fy, g.roi[1], g.roi[4], (float)g.ry, dy, iy);
        const float tz = axis_exit(fz, g.roi[2], g.roi[5], (float)g.rz, dz, iz);
        const float tt = fmaxf(fminf(fminf(tx, ty), tz), 0.0f);
        const float t_target = fminf(__fadd_rn(fm, tt), far);
        float _t = fm;
        do { _t = __fadd_rn(_t, dt_min); } while (_t < t_target);
        t_mid = _t;
        const float dt = clamp_ref(__fmul_rn(t_mid, cone_angle), dt_min, dt_max);
        t0 = __fsub_rn(t_mid, __fmul_rn(dt, 0.5f));
        t1 = __fadd_rn(t_mid, __fmul_rn(dt, 0.5f));
        W = f == 0 ? 1 : 32;                               // still in empty space: probe one sample before fanning out
    }
    if (lane == 0) {
        num_steps[ray] = j;
        if (j > cap) *overflow = 1;
    }
}

// fill round of MODE 2: one warp per ray copies its kept intervals to their packed position.
__global__ void __launch_bounds__(256)
compact_kernel(const int n_rays, const int32_t *__restrict__ packed_info, const float2 *__restrict__ keep,
               const int cap, int64_t *__restrict__ ray_indices, float *__restrict__ t_starts,
               float *__restrict__ t_ends) {
    const int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (ray >= n_rays) return;
    const int base = packed_info[2 * ray], count = packed_info[2 * ray + 1];
    const float2 *src = keep + (size_t)ray * cap;
    for (int j = lane; j < count; j += 32) {
        const float2 v = __ldcs(src + j);
        t_starts[base + j] = v.x;
        t_ends[base + j] = v.y;
        ray_indices[base + j] = ray;
    }
}

// ---- int32 exclusive scan of per-ray counts -> packed_info (base, count) -----------------
// Phase A: per-1024 block sums; phase B: one block scans the sums; phase C: block-local scan
// + offset.  n_rays <= 640k -> <= 625 block sums: three tiny launches, no host round trip.
constexpr int SCAN_T = 256, SCAN_E = 4, SCAN_B = SCAN_T * SCAN_E;

__device__ __forceinline__ int block_exclusive_scan(int v, int *smem, int &block_total) {
    // v: thread value; returns exclusive prefix within the block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < (SCAN_T / 32) ? smem[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < (SCAN_T / 32)) smem[lane] = winc - w;
        if (lane == (SCAN_T / 32) - 1) smem[32] = winc;
    }
    __syncthreads();
    block_total = smem[32];
    const int r = smem[warp] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_T) scan_block_sums(const int32_t *num, int n, int32_t *sums) {
    __shared__ int smem[33];
    const int b0 = blockIdx.x * SCAN_B;
    int v = 0;
#pragma unroll
    for (int e = 0; e < SCAN_E; ++e) {
        int k = b0 + threadIdx.x * SCAN_E + e;
        if (k < n) v += num[k];
    }
    int tot;
    block_exclusive_scan(v, smem, tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_T) scan_sums(int32_t *sums, int nb, int32_t *total) {
    __shared__ int smem[33];
    int carry = 0;
    for (int b0 = 0; b0 < nb; b0 += SCAN_T) {
        int k = b0 + threadIdx.x;
        int v = k < nb ? sums[k] : 0;
        int tot;
        int ex = block_exclusive_scan(v, smem, tot);
        if (k < nb) sums[k] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(SCAN_T)
scan_finalize(const int32_t *num, int n, const int32_t *sums, int32_t *packed_info) {
    __shared__ int smem[33];
    const int b0 = blockIdx.x * SCAN_B;
    int vals[SCAN_E];
    int v = 0;
#pragma unroll
    for (int e = 0; e < SCAN_E; ++e) {
        int k = b0 + threadIdx.x * SCAN_E + e;
        vals[e] = k < n ? num[k] : 0;
        v += vals[e];
    }
    int tot;
    int ex = block_exclusive_scan(v, smem, tot) + sums[blockIdx.x];
#pragma unroll
    for (int e = 0; e < SCAN_E; ++e) {
        int k = b0 + threadIdx.x * SCAN_E + e;
        if (k < n) {
            packed_info[2 * k] = ex;
            packed_info[2 * k + 1] = vals[e];
        }
        ex += vals[e];
    }
}

template <typename scalar_t>
__global__ void aabb_kernel(int n, const scalar_t *__restrict__ rays_o,
                            const scalar_t *__restrict__ rays_d, float a0, float a1, float a2,
                            float a3, float a4, float a5, scalar_t *__restrict__ t_min,
                            scalar_t *__restrict__ t_max) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float ox = rays_o[3 * i], oy = rays_o[3 * i + 1], oz = rays_o[3 * i + 2];
    const float dx = rays_d[3 * i], dy = rays_d[3 * i + 1], dz = rays_d[3 * i + 2];
    float tmin = __fdiv_rn(__fsub_rn(a0, ox), dx), tmax = __fdiv_rn(__fsub_rn(a3, ox), dx);
    if (tmin > tmax) { float c = tmin; tmin = tmax; tmax = c; }
    float tymin = __fdiv_rn(__fsub_rn(a1, oy), dy), tymax = __fdiv_rn(__fsub_rn(a4, oy), dy);
    if (tymin > tymax) { float c = tymin; tymin = tymax; tymax = c; }
    bool miss = (tmin > tymax || tymin > tmax);
    if (!miss) {
        if (tymin > tmin) tmin = tymin;
        if (tymax < tmax) tmax = tymax;
        float tzmin = __fdiv_rn(__fsub_rn(a2, oz), dz), tzmax = __fdiv_rn(__fsub_rn(a5, oz), dz);
        if (tzmin > tzmax) { float c = tzmin; tzmin = tzmax; tzmax = c; }
        miss = (tmin > tzmax || tzmin > tmax);
        if (!miss) {
            if (tzmin > tmin) tmin = tzmin;
            if (tzmax < tmax) tmax = tzmax;
        }
    }
    if (miss) { tmin = 1e10f; tmax = 1e10f; }
    t_min[i] = tmin > 0.0f ? tmin : 0.0f;
    t_max[i] = tmax;
}

__global__ void pack_bits_kernel(const uint8_t *__restrict__ cells, int n_words, uint32_t *bits) {
    // one warp per 32 cells -> one word, via ballot (coalesced byte reads)
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= n_words) return;
    const unsigned m = __ballot_sync(0xffffffffu, cells[gw * 32 + lane] != 0);
    if (lane == 0) bits[gw] = m;
}

__global__ void grid_query_kernel(int n, const float *__restrict__ s, const MarchGrid g,
                                  uint8_t *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = occupied_at(s[3 * i], s[3 * i + 1], s[3 * i + 2], g) ? 1 : 0;
}

__global__ void pack_info_kernel(const int64_t *__restrict__ ri, int n_samples, int n_rays,
                                 int32_t *packed) {
    // one thread per ray: lower/upper bound in the sorted ray_indices (no races, and empty
    // rays get base = insertion point, count = 0 exactly like pack_info's cumsum - n).
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    int lo = 0, hi = n_samples;
    while (lo < hi) { int m = (lo + hi) >> 1; if (ri[m] < r) lo = m + 1; else hi = m; }
    const int first = lo;
    hi = n_samples;
    while (lo < hi) { int m = (lo + hi) >> 1; if (ri[m] <= r) lo = m + 1; else hi = m; }
    packed[2 * r] = first;
    packed[2 * r + 1] = lo - first;
}

MarchGrid make_grid(const float *roi, const uint8_t *cells, const uint32_t *bits, int rx, int ry,
                    int rz) {
    MarchGrid g;
    for (int k = 0; k < 6; ++k) g.roi[k] = roi[k];
    g.rx = rx; g.ry = ry; g.rz = rz;
    g.cells = cells; g.bits = bits;
    return g;
}

}  // namespace

extern "C" {

int rsdf_ray_aabb_intersect(const float *rays_o, const float *rays_d, const float *aabb,
                            int n_rays, float *t_min, float *t_max, void *stream) {
    if (n_rays == 0) return 0;
    if (!rays_o || !rays_d || !aabb || !t_min || !t_max) return RSDF_EBADARG;
    aabb_kernel<float><<<rsdf_div_up(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(
        n_rays, rays_o, rays_d, aabb[0], aabb[1], aabb[2], aabb[3], aabb[4], aabb[5], t_min, t_max);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_grid_pack_bits(const uint8_t *grid_binary, int n_cells, uint32_t *bits, void *stream) {
    if (!grid_binary || !bits || n_cells % 32) return RSDF_EBADARG;
    const int n_words = n_cells / 32;
    pack_bits_kernel<<<rsdf_div_up((long long)n_words * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        grid_binary, n_words, bits);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_march_count(const float *rays_o, const float *rays_d, const float *t_min,
                     const float *t_max, const float *roi, const uint8_t *grid_binary,
                     const uint32_t *grid_bits, int rx, int ry, int rz, float step_size,
                     float cone_angle, int n_rays, int32_t *packed_info, int32_t *scan_tmp,
                     int32_t *total, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!total) return RSDF_EBADARG;
    if (n_rays == 0) { return (int)cudaMemsetAsync(total, 0, sizeof(int32_t), st); }
    if (!rays_o || !rays_d || !t_min || !t_max || !roi || (!grid_binary && !grid_bits) ||
        !packed_info || !scan_tmp)
        return RSDF_EBADARG;
    const MarchGrid g = make_grid(roi, grid_binary, grid_bits, rx, ry, rz);
    // layout of scan_tmp: [nb+1 block sums][n_rays counts]
    const int nb = rsdf_div_up(n_rays, SCAN_B);
    int32_t *sums = scan_tmp;
    int32_t *num = scan_tmp + nb + 1;
    march_kernel<0><<<rsdf_div_up(n_rays, 64), 64, 0, st>>>(
        n_rays, rays_o, rays_d, t_min, t_max, g, step_size, cone_angle, nullptr, num, nullptr,
        nullptr, nullptr);
    RSDF_LAUNCH_CHECK();
    scan_block_sums<<<nb, SCAN_T, 0, st>>>(num, n_rays, sums);
    scan_sums<<<1, SCAN_T, 0, st>>>(sums, nb, total);
    scan_finalize<<<nb, SCAN_T, 0, st>>>(num, n_rays, sums, packed_info);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_march_count_keep(const float *rays_o, const float *rays_d, const float *t_min,
                          const float *t_max, const float *roi, const uint8_t *grid_binary,
                          const uint32_t *grid_bits, int rx, int ry, int rz, float step_size,
                          float cone_angle, int n_rays, int32_t *packed_info, int32_t *scan_tmp,
                          int32_t *total2, float *keep, int cap, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!total2) return RSDF_EBADARG;
    cudaError_t e = cudaMemsetAsync(total2, 0, 2 * sizeof(int32_t), st);
    if (e != cudaSuccess) return (int)e;
    if (n_rays == 0) return 0;
    if (!rays_o || !rays_d || !t_min || !t_max || !roi || (!grid_binary && !grid_bits) ||
        !packed_info || !scan_tmp || !keep || cap < 1 || (((uintptr_t)keep) & 7))
        return RSDF_EBADARG;
    const MarchGrid g = make_grid(roi, grid_binary, grid_bits, rx, ry, rz);
    const int nb = rsdf_div_up(n_rays, SCAN_B);
    int32_t *sums = scan_tmp;
    int32_t *num = scan_tmp + nb + 1;
    const int wblocks = rsdf_div_up((long long)n_rays * 32, 128);
    if (cone_angle == 0.0f)
        march_warp_kernel<true><<<wblocks, 128, 0, st>>>(n_rays, rays_o, rays_d, t_min, t_max, g, step_size,
                                                         cone_angle, num, reinterpret_cast<float2 *>(keep), cap,
                                                         total2 + 1);
    else
        march_warp_kernel<false><<<wblocks, 128, 0, st>>>(n_rays, rays_o, rays_d, t_min, t_max, g, step_size,
                                                          cone_angle, num, reinterpret_cast<float2 *>(keep), cap,
                                                          total2 + 1);
    RSDF_LAUNCH_CHECK();
    scan_block_sums<<<nb, SCAN_T, 0, st>>>(num, n_rays, sums);
    scan_sums<<<1, SCAN_T, 0, st>>>(sums, nb, total2);
    scan_finalize<<<nb, SCAN_T, 0, st>>>(num, n_rays, sums, packed_info);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_march_compact(const int32_t *packed_info, const float *keep, int cap, int n_rays,
                       int64_t *ray_indices, float *t_starts, float *t_ends, void *stream) {
    if (n_rays == 0) return 0;
    if (!packed_info || !keep || cap < 1 || !ray_indices || !t_starts || !t_ends) return RSDF_EBADARG;
    compact_kernel<<<rsdf_div_up((long long)n_rays * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        n_rays, packed_info, reinterpret_cast<const float2 *>(keep), cap, ray_indices, t_starts, t_ends);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_march_fill(const float *rays_o, const float *rays_d, const float *t_min,
                    const float *t_max, const float *roi, const uint8_t *grid_binary,
                    const uint32_t *grid_bits, int rx, int ry, int rz, float step_size,
                    float cone_angle, int n_rays, const int32_t *packed_info,
                    int64_t *ray_indices, float *t_starts, float *t_ends, void *stream) {
    if (n_rays == 0) return 0;
    if (!rays_o || !rays_d || !t_min || !t_max || !roi || (!grid_binary && !grid_bits) ||
        !packed_info)
        return RSDF_EBADARG;
    const MarchGrid g = make_grid(roi, grid_binary, grid_bits, rx, ry, rz);
    march_kernel<1><<<rsdf_div_up(n_rays, 64), 64, 0, (cudaStream_t)stream>>>(
        n_rays, rays_o, rays_d, t_min, t_max, g, step_size, cone_angle, packed_info, nullptr,
        ray_indices, t_starts, t_ends);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_grid_query(const float *samples, const float *roi, const uint8_t *grid_binary, int rx,
                    int ry, int rz, int n_samples, uint8_t *out, void *stream) {
    if (n_samples == 0) return 0;
    if (!samples || !roi || !grid_binary || !out) return RSDF_EBADARG;
    const MarchGrid g = make_grid(roi, grid_binary, nullptr, rx, ry, rz);
    grid_query_kernel<<<rsdf_div_up(n_samples, 256), 256, 0, (cudaStream_t)stream>>>(
        n_samples, samples, g, out);
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_pack_info(const int64_t *ray_indices, int n_samples, int n_rays, int32_t *packed_info,
                   void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!packed_info) return RSDF_EBADARG;
    if (n_rays == 0) return 0;
    if (n_samples > 0 && !ray_indices) return RSDF_EBADARG;
    pack_info_kernel<<<rsdf_div_up(n_rays, 256), 256, 0, st>>>(ray_indices, n_samples, n_rays,
                                                               packed_info);
    RSDF_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
