// Adam over ONE flat fp32 parameter buffer: the optimizer of
// systems/utils.py:309-320 (`torch.optim.Adam` with per-group learning rates,
// configs/neus-blender.yaml:92-104, configs/split-mixed-occ-tensoir.yaml:153-166) as a single
// HBM-bound launch.  Parameters, gradients (the data-parallel exchange bucket) and both moments are
// flat buffers in the same order, so one pass reads p, g, m, v and writes p, m, v: 28 B/parameter
// (+4 when the gradient bucket is cleared for the next step in the same pass).
#include "common.cuh"

namespace {

struct AdamElem {
    float step_size, omb1, b2, omb2, eps, bc2_sqrt, wd;
    bool skip;
};

__device__ __forceinline__ int find_seg(const rsdf_adam_groups &G, long long i) {
    int s = 0;
#pragma unroll
    for (int k = 0; k < RSDF_ADAM_MAX_GROUPS - 1; ++k) s += (k < G.n_groups - 1 && i >= G.end[k]) ? 1 : 0;
    return s;
}

__device__ __forceinline__ AdamElem load_seg(const rsdf_adam_groups &G, int s) {
    AdamElem e;
    e.step_size = G.step_size[s];
    e.omb1 = G.one_minus_beta1[s];
    e.b2 = G.beta2[s];
    e.omb2 = G.one_minus_beta2[s];
    e.eps = G.eps[s];
    e.bc2_sqrt = G.bias2_sqrt[s];
    e.wd = G.weight_decay[s];
    e.skip = G.skip[s] != 0;
    return e;
}

// torch/optim/adam.py `_single_tensor_adam` (amsgrad off, maximize off): the operation order of
// lerp_ / mul_.addcmul_ / sqrt / div / add / addcdiv_ kept, every product rounded on its own.
__device__ __forceinline__ void adam_one(float &p, float g, float &m, float &v, const AdamElem &e) {
    if (e.skip) return;                                                // a group that is off the graph this step
    if (e.wd != 0.0f) g = __fmaf_rn(e.wd, p, g);                       // grad.add(param, alpha=wd)
    m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), e.omb1));            // exp_avg.lerp_(grad, 1-b1)
    v = __fadd_rn(__fmul_rn(v, e.b2), __fmul_rn(__fmul_rn(e.omb2, g), g));  // mul_(b2).addcmul_(g,g,1-b2)
    float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), e.bc2_sqrt), e.eps);
    p = __fsub_rn(p, __fmul_rn(e.step_size, __fdiv_rn(m, denom)));
}

template <bool kZeroGrad>
__global__ void __launch_bounds__(256) adam_flat_kernel(float *__restrict__ p, float *__restrict__ g,
                                                       float *__restrict__ m, float *__restrict__ v, long long n,
                                                       rsdf_adam_groups G) {
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
        float4 P = reinterpret_cast<float4 *>(p)[q];
        float4 Gr = __ldcs(reinterpret_cast<const float4 *>(g) + q);
        float4 M = reinterpret_cast<float4 *>(m)[q];
        float4 V = reinterpret_cast<float4 *>(v)[q];
        const long long i = q << 2;
        const int s0 = find_seg(G, i), s3 = find_seg(G, i + 3);
        AdamElem e = load_seg(G, s0);
        if (s0 == s3) {
            adam_one(P.x, Gr.x, M.x, V.x, e);
            adam_one(P.y, Gr.y, M.y, V.y, e);
            adam_one(P.z, Gr.z, M.z, V.z, e);
            adam_one(P.w, Gr.w, M.w, V.w, e);
        } else {            // a group boundary inside the vector
            adam_one(P.x, Gr.x, M.x, V.x, e);
            e = load_seg(G, find_seg(G, i + 1));
            adam_one(P.y, Gr.y, M.y, V.y, e);
            e = load_seg(G, find_seg(G, i + 2));
            adam_one(P.z, Gr.z, M.z, V.z, e);
            e = load_seg(G, s3);
            adam_one(P.w, Gr.w, M.w, V.w, e);
        }
        reinterpret_cast<float4 *>(p)[q] = P;
        reinterpret_cast<float4 *>(m)[q] = M;
        reinterpret_cast<float4 *>(v)[q] = V;
        if (kZeroGrad) reinterpret_cast<float4 *>(g)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // tail (n % 4 elements)
    const long long t = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) {
        AdamElem e = load_seg(G, find_seg(G, t));
        float P = p[t], M = m[t], V = v[t];
        adam_one(P, g[t], M, V, e);
        p[t] = P;
        m[t] = M;
        v[t] = V;
        if (kZeroGrad) g[t] = 0.f;
    }
}

}  // namespace

extern "C" int rsdf_adam_step(float *params, float *grads, float *exp_avg, float *exp_avg_sq, long long n,
                              const rsdf_adam_groups *groups, int zero_grad, void *stream) {
    if (n == 0) return 0;
    if (!params || !grads || !exp_avg || !exp_avg_sq || !groups) return RSDF_EBADARG;
    if (groups->n_groups < 1 || groups->n_groups > RSDF_ADAM_MAX_GROUPS) return RSDF_EBADARG;
    if (groups->end[groups->n_groups - 1] != n) return RSDF_EBADARG;
    if ((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) != 0) return RSDF_EBADARG;
    long long want = ((n >> 2) + 255) / 256;
    int grid = (int)(want < 1 ? 1 : (want > 8LL * RSDF_NUM_SMS ? 8 * RSDF_NUM_SMS : want));
    if (zero_grad)
        adam_flat_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, *groups);
    else
        adam_flat_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, *groups);
    RSDF_LAUNCH_CHECK();
    return 0;
}
