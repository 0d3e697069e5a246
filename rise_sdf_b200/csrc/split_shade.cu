// K5c -- the split-sum combine of models/texture.py:330-377 (`VolumeMixedMipSplitOcc.forward` after the four material
// networks) as ONE kernel per direction, one thread per sample:
//   sigmoid of the 12 raw network outputs, blend mixes, diffuse/specular albedo, reflection direction, the FG LUT
//   lookup (2-D linear clamp, models/texture.py:340), the diffuse irradiance lookup (cube linear,
//   lib/pbr/light.py:203-206), get_mip + the prefiltered specular lookup (cube linear-mipmap-linear,
//   lib/pbr/light.py:168-202) and the 24-channel packing of models/texture.py:345:
//   [diff3, spec3, blend1, diff_pbr3, spec_pbr3, spec_ref3, spec_light3, albedo3, metallic1, roughness1]
//   (stage 0: the first 7).
// The op-by-op path is ~45 torch launches forward and ~110 backward over [S, 1..24] tensors; here every sample's
// intermediate lives in registers.  The backward recomputes the forward from the raw outputs (cheaper than storing
// 40 floats per sample) and scatters the emitter gradients with atomics exactly like rsdf_cube_sample_bwd.
#include "texture.cuh"

namespace {
using namespace rsdf_tex;

struct ShadeArgs {
    const float *raw_albedo;   // [S,6]  (diff_rgb 3 | albedo 3)   albedo_network
    const float *raw_rough;    // [S,1]                            roughness_network
    const float *raw_metal;    // [S,2]  (blend | metallic)        metallic_network
    const float *raw_env;      // [S,3]                            env_network
    const float *normals;      // [S,3]
    const float *dirs;         // [S,3]  ray directions (wi = -dirs)
    const float *lut;          // [H,W,2]
    int lut_h, lut_w;
    const float *diffuse;      // [6,Nd,Nd,3]
    int diffuse_res;
    float min_rough, max_rough;
    int n;
};

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

// lib/pbr/light.py:168-176 get_mip and its derivative (the clamp masks of torch.clamp's backward are inclusive)
__device__ __forceinline__ float get_mip(float R, float lo, float hi, int n_levels, float &dR) {
    if (R < hi) {
        const float sc = (float)(n_levels - 2) / (hi - lo);
        dR = (R >= lo && R <= hi) ? sc : 0.0f;
        return (fminf(fmaxf(R, lo), hi) - lo) / (hi - lo) * (float)(n_levels - 2);
    }
    dR = (R >= hi && R <= 1.0f) ? 1.0f / (1.0f - hi) : 0.0f;
    return (fminf(fmaxf(R, hi), 1.0f) - hi) / (1.0f - hi) + (float)(n_levels - 2);
}

struct Mat {
    float D[3], A[3], E[3], R, B, M;
    float n[3], wi[3], wo[3], NoV, win;
};

__device__ __forceinline__ Mat load_material(const ShadeArgs &a, int s) {
    Mat m;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        m.D[c] = sigmoidf(a.raw_albedo[6 * (size_t)s + c]);
        m.A[c] = sigmoidf(a.raw_albedo[6 * (size_t)s + 3 + c]);
        m.E[c] = sigmoidf(a.raw_env[3 * (size_t)s + c]);
        m.n[c] = a.normals[3 * (size_t)s + c];
        m.wi[c] = -a.dirs[3 * (size_t)s + c];
    }
    m.R = sigmoidf(a.raw_rough[s]);
    m.B = sigmoidf(a.raw_metal[2 * (size_t)s]);
    m.M = sigmoidf(a.raw_metal[2 * (size_t)s + 1]);
    m.win = m.wi[0] * m.n[0] + m.wi[1] * m.n[1] + m.wi[2] * m.n[2];
#pragma unroll
    for (int c = 0; c < 3; ++c) m.wo[c] = m.win * m.n[c] * 2.0f - m.wi[c];
    m.NoV = m.n[0] * m.wi[0] + m.n[1] * m.wi[1] + m.n[2] * m.wi[2];
    return m;
}

template <bool STAGE1>
__global__ void __launch_bounds__(256)
split_shade_fwd_kernel(const ShadeArgs a, const MipStack mips, float *__restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n) return;
    const Mat m = load_material(a, s);
    constexpr int CD = STAGE1 ? 24 : 7;
    float o[CD];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        o[c] = (1.0f - m.B) * m.D[c];
        o[3 + c] = m.B * m.E[c];
    }
    o[6] = m.B;
    if (STAGE1) {
        // FG LUT (uv.x = NoV, uv.y = roughness; both clamped to [0,1])
        int x0, x1, y0, y1;
        float ax, ay;
        tex2d_taps(fminf(fmaxf(m.NoV, 0.0f), 1.0f), a.lut_w, false, x0, x1, ax);
        tex2d_taps(fminf(fmaxf(m.R, 0.0f), 1.0f), a.lut_h, false, y0, y1, ay);
        float fg[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float v00 = __ldg(a.lut + ((size_t)y0 * a.lut_w + x0) * 2 + c), v10 = __ldg(a.lut + ((size_t)y0 * a.lut_w + x1) * 2 + c);
            const float v01 = __ldg(a.lut + ((size_t)y1 * a.lut_w + x0) * 2 + c), v11 = __ldg(a.lut + ((size_t)y1 * a.lut_w + x1) * 2 + c);
            const float top = fmaf(ax, v10 - v00, v00), bot = fmaf(ax, v11 - v01, v01);
            fg[c] = fmaf(ay, bot - top, top);
        }
        float DL[3], SL[3], dR;
        cube_fetch<3>(a.diffuse, a.diffuse_res, m.n[0], m.n[1], m.n[2], DL);
        const float lv = get_mip(m.R, a.min_rough, a.max_rough, mips.n_levels, dR);
        cube_fetch_mip<3>(mips, lv, m.wo[0], m.wo[1], m.wo[2], SL);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float da = (1.0f - m.M) * m.A[c];
            const float sa = 0.04f * (1.0f - m.M) + m.M * m.A[c];
            const float sr = sa * fg[0] + fg[1];
            o[7 + c] = da * DL[c];
            o[10 + c] = sr * SL[c];
            o[13 + c] = sr;
            o[16 + c] = SL[c];
            o[19 + c] = m.A[c];
        }
        o[22] = m.M;
        o[23] = m.R;
    }
    float *dst = out + (size_t)s * CD;
    if (STAGE1) {
#pragma unroll
        for (int c = 0; c < CD; c += 4) *reinterpret_cast<float4 *>(dst + c) = make_float4(o[c], o[c + 1], o[c + 2], o[c + 3]);
    } else {
#pragma unroll
        for (int c = 0; c < CD; ++c) dst[c] = o[c];
    }
}

struct ShadeGrads {
    float *raw_albedo, *raw_rough, *raw_metal, *raw_env, *normals;   // same shapes as the inputs
    float *lut;        // [H,W,2] or null
    float *diffuse;    // [6,Nd,Nd,3] or null
};

template <bool STAGE1>
__global__ void __launch_bounds__(256)
split_shade_bwd_kernel(const ShadeArgs a, const MipStack mips, const float *__restrict__ go, const ShadeGrads G) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n) return;
    const Mat m = load_material(a, s);
    constexpr int CD = STAGE1 ? 24 : 7;
    float g[CD];
#pragma unroll
    for (int c = 0; c < CD; ++c) g[c] = go[(size_t)s * CD + c];
    float gD[3], gA[3] = {0.f, 0.f, 0.f}, gE[3], gB = g[6], gM = 0.f, gR = 0.f, gn[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        gD[c] = (1.0f - m.B) * g[c];
        gE[c] = m.B * g[3 + c];
        gB += -m.D[c] * g[c] + m.E[c] * g[3 + c];
    }
    if (STAGE1) {
        int x0, x1, y0, y1;
        float ax, ay;
        tex2d_taps(fminf(fmaxf(m.NoV, 0.0f), 1.0f), a.lut_w, false, x0, x1, ax);
        tex2d_taps(fminf(fmaxf(m.R, 0.0f), 1.0f), a.lut_h, false, y0, y1, ay);
        float fg[2], dfg_u[2], dfg_v[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float v00 = __ldg(a.lut + ((size_t)y0 * a.lut_w + x0) * 2 + c), v10 = __ldg(a.lut + ((size_t)y0 * a.lut_w + x1) * 2 + c);
            const float v01 = __ldg(a.lut + ((size_t)y1 * a.lut_w + x0) * 2 + c), v11 = __ldg(a.lut + ((size_t)y1 * a.lut_w + x1) * 2 + c);
            const float top = fmaf(ax, v10 - v00, v00), bot = fmaf(ax, v11 - v01, v01);
            fg[c] = fmaf(ay, bot - top, top);
            dfg_u[c] = ((1.f - ay) * (v10 - v00) + ay * (v11 - v01)) * (float)a.lut_w;
            dfg_v[c] = ((1.f - ax) * (v01 - v00) + ax * (v11 - v10)) * (float)a.lut_h;
        }
        float DL[3], dR;
        cube_fetch<3>(a.diffuse, a.diffuse_res, m.n[0], m.n[1], m.n[2], DL);
        const float lv = get_mip(m.R, a.min_rough, a.max_rough, mips.n_levels, dR);
        float SL[3];
        cube_fetch_mip<3>(mips, lv, m.wo[0], m.wo[1], m.wo[2], SL);

        float gDL[3], gSL[3], g_fg0 = 0.f, g_fg1 = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float da = (1.0f - m.M) * m.A[c];
            const float sa = 0.04f * (1.0f - m.M) + m.M * m.A[c];
            const float sr = sa * fg[0] + fg[1];
            const float g_da = DL[c] * g[7 + c];
            gDL[c] = da * g[7 + c];
            const float g_sr = g[13 + c] + SL[c] * g[10 + c];
            gSL[c] = g[16 + c] + sr * g[10 + c];
            const float g_sa = fg[0] * g_sr;
            g_fg0 = fmaf(sa, g_sr, g_fg0);
            g_fg1 += g_sr;
            gA[c] = g[19 + c] + (1.0f - m.M) * g_da + m.M * g_sa;
            gM += -m.A[c] * g_da + (m.A[c] - 0.04f) * g_sa;
        }
        gM += g[22];
        gR = g[23];
        // FG LUT backward
        const float g_u = g_fg0 * dfg_u[0] + g_fg1 * dfg_u[1], g_v = g_fg0 * dfg_v[0] + g_fg1 * dfg_v[1];
        const float gNoV = (m.NoV >= 0.0f && m.NoV <= 1.0f) ? g_u : 0.0f;
        if (m.R >= 0.0f && m.R <= 1.0f) gR += g_v;
        if (G.lut) {
            const float gf[2] = {g_fg0, g_fg1};
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                atomicAdd(G.lut + ((size_t)y0 * a.lut_w + x0) * 2 + c, gf[c] * (1.f - ax) * (1.f - ay));
                atomicAdd(G.lut + ((size_t)y0 * a.lut_w + x1) * 2 + c, gf[c] * ax * (1.f - ay));
                atomicAdd(G.lut + ((size_t)y1 * a.lut_w + x0) * 2 + c, gf[c] * (1.f - ax) * ay);
                atomicAdd(G.lut + ((size_t)y1 * a.lut_w + x1) * 2 + c, gf[c] * ax * ay);
            }
        }
        // diffuse irradiance lookup backward: texels + direction (the normal)
        {
            float val[3], d_u = 0.f, d_v = 0.f;
            cube_level_bwd(a.diffuse, G.diffuse, a.diffuse_res, m.n[0], m.n[1], m.n[2], 1.0f, gDL, true, val, d_u, d_v);
            face_uv_grad_to_dir(m.n[0], m.n[1], m.n[2], d_u, d_v, gn);
        }
        // specular lookup backward: texels of the two levels, the level (-> roughness), the direction (wo -> normal)
        {
            int l0, l1; float f; bool live;
            mip_levels(lv, mips.n_levels, l0, l1, f, live);
            float va[3], vb[3] = {0.f, 0.f, 0.f}, d_u = 0.f, d_v = 0.f;
            cube_level_bwd(mips.level[l0], mips.grad[l0], mips.res[l0], m.wo[0], m.wo[1], m.wo[2], 1.0f - f, gSL, true, va, d_u, d_v);
            if (l1 != l0)
                cube_level_bwd(mips.level[l1], mips.grad[l1], mips.res[l1], m.wo[0], m.wo[1], m.wo[2], f, gSL, true, vb, d_u, d_v);
            if (live && l1 != l0)
                gR += dR * ((vb[0] - va[0]) * gSL[0] + (vb[1] - va[1]) * gSL[1] + (vb[2] - va[2]) * gSL[2]);
            float gwo[3] = {0.f, 0.f, 0.f};
            face_uv_grad_to_dir(m.wo[0], m.wo[1], m.wo[2], d_u, d_v, gwo);
            // wo = 2 (wi.n) n - wi
            const float gwon = gwo[0] * m.n[0] + gwo[1] * m.n[1] + gwo[2] * m.n[2];
#pragma unroll
            for (int c = 0; c < 3; ++c) gn[c] += 2.0f * (gwon * m.wi[c] + m.win * gwo[c]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) gn[c] = fmaf(gNoV, m.wi[c], gn[c]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        G.raw_albedo[6 * (size_t)s + c] = gD[c] * m.D[c] * (1.0f - m.D[c]);
        G.raw_albedo[6 * (size_t)s + 3 + c] = gA[c] * m.A[c] * (1.0f - m.A[c]);
        G.raw_env[3 * (size_t)s + c] = gE[c] * m.E[c] * (1.0f - m.E[c]);
        if (G.normals) G.normals[3 * (size_t)s + c] = gn[c];
    }
    G.raw_rough[s] = gR * m.R * (1.0f - m.R);
    G.raw_metal[2 * (size_t)s] = gB * m.B * (1.0f - m.B);
    G.raw_metal[2 * (size_t)s + 1] = gM * m.M * (1.0f - m.M);
}

int fill_mips(MipStack &m, const float *const *levels, float *const *grads, const int *res, int n_levels) {
    if (n_levels < 2 || n_levels > 8 || !levels || !res) return RSDF_EBADARG;
    m.n_levels = n_levels;
    for (int l = 0; l < 8; ++l) {
        m.level[l] = l < n_levels ? levels[l] : nullptr;
        m.grad[l] = (l < n_levels && grads) ? grads[l] : nullptr;
        m.res[l] = l < n_levels ? res[l] : 0;
    }
    return 0;
}

ShadeArgs pack(const rsdf_split_shade_args *p) {
    ShadeArgs a;
    a.raw_albedo = p->raw_albedo; a.raw_rough = p->raw_roughness; a.raw_metal = p->raw_metallic; a.raw_env = p->raw_env;
    a.normals = p->normals; a.dirs = p->dirs;
    a.lut = p->fg_lut; a.lut_h = p->lut_h; a.lut_w = p->lut_w;
    a.diffuse = p->diffuse; a.diffuse_res = p->diffuse_res;
    a.min_rough = p->min_roughness; a.max_rough = p->max_roughness;
    a.n = p->n;
    return a;
}

}  // namespace

extern "C" {

int rsdf_split_shade_fwd(const rsdf_split_shade_args *p, float *colors, void *stream) {
    if (!p) return RSDF_EBADARG;
    if (p->n == 0) return 0;
    if (!p->raw_albedo || !p->raw_roughness || !p->raw_metallic || !p->raw_env || !p->normals || !p->dirs || !colors)
        return RSDF_EBADARG;
    const ShadeArgs a = pack(p);
    MipStack m = {};
    cudaStream_t st = (cudaStream_t)stream;
    if (p->stage == 0) {
        split_shade_fwd_kernel<false><<<rsdf_div_up(p->n, 256), 256, 0, st>>>(a, m, colors);
    } else {
        if (!p->fg_lut || !p->diffuse) return RSDF_EBADARG;
        if (int e = fill_mips(m, p->specular_levels, nullptr, p->specular_res, p->n_levels)) return e;
        split_shade_fwd_kernel<true><<<rsdf_div_up(p->n, 256), 256, 0, st>>>(a, m, colors);
    }
    RSDF_LAUNCH_CHECK();
    return 0;
}

int rsdf_split_shade_bwd(const rsdf_split_shade_args *p, const float *grad_colors, float *grad_raw_albedo,
                         float *grad_raw_roughness, float *grad_raw_metallic, float *grad_raw_env, float *grad_normals,
                         float *grad_fg_lut, float *grad_diffuse, float *const *grad_specular_levels, void *stream) {
    if (!p) return RSDF_EBADARG;
    if (p->n == 0) return 0;
    if (!p->raw_albedo || !p->raw_roughness || !p->raw_metallic || !p->raw_env || !p->normals || !p->dirs ||
        !grad_colors || !grad_raw_albedo || !grad_raw_roughness || !grad_raw_metallic || !grad_raw_env)
        return RSDF_EBADARG;
    const ShadeArgs a = pack(p);
    ShadeGrads G = {grad_raw_albedo, grad_raw_roughness, grad_raw_metallic, grad_raw_env, grad_normals, grad_fg_lut, grad_diffuse};
    MipStack m = {};
    cudaStream_t st = (cudaStream_t)stream;
    if (p->stage == 0) {
        split_shade_bwd_kernel<false><<<rsdf_div_up(p->n, 256), 256, 0, st>>>(a, m, grad_colors, G);
    } else {
        if (!p->fg_lut || !p->diffuse) return RSDF_EBADARG;
        if (int e = fill_mips(m, p->specular_levels, grad_specular_levels, p->specular_res, p->n_levels)) return e;
        split_shade_bwd_kernel<true><<<rsdf_div_up(p->n, 256), 256, 0, st>>>(a, m, grad_colors, G);
    }
    RSDF_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
