// Device-side texture sampling shared by texture.cu (the nvdiffrast-shaped lookups) and split_shade.cu (the fused
// split-sum shading kernel).  Cube face / in-face conventions: lib/renderutils/c_src/cubemap.cu:32-60.
#pragma once
#include "common.cuh"

namespace rsdf_tex {

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

// One mip level's share of a cube lookup backward: adds wl * w_k * g to the level's texel gradients (atomics), returns
// the filtered value in val[3] and accumulates d<g, out>/d(in-face u, v) into (d_u, d_v).
__device__ __forceinline__ void cube_level_bwd(const float *__restrict__ level, float *__restrict__ grad, int N, float dx,
                                               float dy, float dz, float wl, const float g[3], bool want_dir,
                                               float val[3], float &d_u, float &d_v) {
    const CubeTap t = cube_taps(dx, dy, dz, N);
    float tex[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) tex[k][c] = t.idx[k] >= 0 ? __ldg(level + (size_t)t.idx[k] * 3 + c) : 0.f;
    if (want_dir && wl != 0.0f) {
        // bilinear weight derivatives; taps dropped at cube corners contribute nothing
        int face; float u, v;
        dir_to_face(dx, dy, dz, face, u, v);
        const float Nf = (float)N;
        const float tx = (u + 1.0f) * 0.5f * Nf - 0.5f, ty = (v + 1.0f) * 0.5f * Nf - 0.5f;
        const float ax = tx - floorf(tx), ay = ty - floorf(ty);
        float gt[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            gt[k] = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) gt[k] = fmaf(g[c], tex[k][c], gt[k]);
        }
        d_u += wl * 0.5f * Nf * ((1.f - ay) * (gt[1] - gt[0]) + ay * (gt[3] - gt[2]));
        d_v += wl * 0.5f * Nf * ((1.f - ax) * (gt[2] - gt[0]) + ax * (gt[3] - gt[1]));
    }
    val[0] = val[1] = val[2] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (t.idx[k] < 0) continue;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            val[c] = fmaf(t.w[k], tex[k][c], val[c]);
            if (grad && wl != 0.0f) atomicAdd(grad + (size_t)t.idx[k] * 3 + c, wl * t.w[k] * g[c]);
        }
    }
}

// chain (u, v) = (cu*A, cv*B)/|M| back to the direction (face table: cubemap.cu:48-60); ADDS into gd
__device__ __forceinline__ void face_uv_grad_to_dir(float dx, float dy, float dz, float d_u, float d_v, float gd[3]) {
    int face; float u, v;
    dir_to_face(dx, dy, dz, face, u, v);
    const float d[3] = {dx, dy, dz};
    const int Mx[6] = {0, 0, 1, 1, 2, 2}, Ax[6] = {2, 2, 0, 0, 0, 0}, Bx[6] = {1, 1, 2, 2, 1, 1};
    const float cu[6] = {-1.f, 1.f, 1.f, 1.f, 1.f, -1.f}, cv[6] = {-1.f, -1.f, 1.f, -1.f, -1.f, -1.f};
    const float am = fabsf(d[Mx[face]]), im = 1.0f / am;
    gd[Ax[face]] += cu[face] * im * d_u;
    gd[Bx[face]] += cv[face] * im * d_v;
    gd[Mx[face]] += -(u * d_u + v * d_v) * im * (d[Mx[face]] > 0.f ? 1.f : -1.f);
}

// linear-mipmap-linear level selection shared by the forward and the backward
__device__ __forceinline__ void mip_levels(float raw, int n_levels, int &l0, int &l1, float &f, bool &live) {
    const float lv = fminf(fmaxf(raw, 0.0f), (float)(n_levels - 1));
    l0 = min((int)floorf(lv), n_levels - 1);
    l1 = min(l0 + 1, n_levels - 1);
    f = lv - (float)l0;
    live = raw >= 0.0f && raw <= (float)(n_levels - 1);
}

template <int C>
__device__ __forceinline__ void cube_fetch_mip(const MipStack &m, float raw_level, float dx, float dy, float dz, float *r) {
    int l0, l1; float f; bool live;
    mip_levels(raw_level, m.n_levels, l0, l1, f, live);
    cube_fetch<C>(m.level[l0], m.res[l0], dx, dy, dz, r);
    if (f > 0.0f && l1 != l0) {
        float b[C];
        cube_fetch<C>(m.level[l1], m.res[l1], dx, dy, dz, b);
#pragma unroll
        for (int c = 0; c < C; ++c) r[c] = fmaf(f, b[c] - r[c], r[c]);
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

}  // namespace rsdf_tex
