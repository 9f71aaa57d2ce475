// Soft ray/rectangle blocking (artist/raytracing/blocking.py:212-354) evaluated per ray inside the trace kernels.
//
// The reference evaluates EVERY ray against the set of primitives that its LBVH filter kept for the whole batch
// (O(rays x K)).  Here each heliostat-sample carries a short candidate list (the primitives whose bounding sphere
// touches the tapered capsule between the heliostat and its target, built by ab200_blocking_candidates), each
// surface point culls that list against the undistorted reflection with a margin for the sun-shape scatter, and a
// ray evaluates the five sigmoids only when it is within 3 % of a rectangle (a primitive dropped this way
// contributes < 1e-9 to the optical depth, far below fp32 resolution of 1 - blocked).
#pragma once
#include "common.cuh"

namespace ab200 {

constexpr int kMaxBlockCandidates = 64;

// packed primitive: 16 floats
struct BlockPrim {
    float c0[3], su[3], sv[3], n[3];
    float uu, vv, uv, det;  // |su|^2, |sv|^2, su.sv, safe determinant
};

struct BlockParams {
    float softness, alpha, offset, epsilon;
    float cull_angle;  // bound on the scatter angle used by the per-point cull (<= 0: no per-point culling)
};

__device__ __forceinline__ float sigmoidf_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

struct BlockGeom {
    float t, u, v, den;
    float off[3];
};

__device__ __forceinline__ void block_geometry(BlockGeom& g, const BlockPrim& p, const BlockParams& bp, float o0, float o1,
                                               float o2, float d0, float d1, float d2) {
    float den = d0 * p.n[0] + d1 * p.n[1] + d2 * p.n[2];
    if (fabsf(den) < bp.epsilon) den = den >= 0.0f ? bp.epsilon : -bp.epsilon;
    const float num = (p.c0[0] - o0) * p.n[0] + (p.c0[1] - o1) * p.n[1] + (p.c0[2] - o2) * p.n[2];
    const float t = num / den;
    g.off[0] = o0 + t * d0 - p.c0[0]; g.off[1] = o1 + t * d1 - p.c0[1]; g.off[2] = o2 + t * d2 - p.c0[2];
    const float pu = g.off[0] * p.su[0] + g.off[1] * p.su[1] + g.off[2] * p.su[2];
    const float pv = g.off[0] * p.sv[0] + g.off[1] * p.sv[1] + g.off[2] * p.sv[2];
    g.u = (pu * p.vv - pv * p.uv) / p.det;
    g.v = (pv * p.uu - pu * p.uv) / p.det;
    g.t = t; g.den = den;
}

// a ray this far outside the rectangle (or this close to / behind its origin) contributes < 1e-9
__device__ __forceinline__ bool block_relevant(const BlockGeom& g, const BlockParams& bp) {
    const float m = 20.8f / bp.softness;
    return (g.t > bp.offset - m) && (g.u > -m) && (g.u < 1.0f + m) && (g.v > -m) && (g.v < 1.0f + m);
}

// per-point cull of the candidate list against the undistorted reflection direction: bit c of the result is set if
// candidate c may matter for some ray of this point
static __device__ __noinline__ unsigned long long block_point_mask(const BlockPrim* prims, int n_cand, const BlockParams& bp,
                                                               float o0, float o1, float o2, float r0, float r1, float r2) {
    unsigned long long mask = 0ull;
    for (int c = 0; c < n_cand; ++c) {
        const BlockPrim& p = prims[c];
        bool keep = true;
        if (bp.cull_angle > 0.0f) {
            BlockGeom g;
            block_geometry(g, p, bp, o0, o1, o2, r0, r1, r2);
            const float cosi = fabsf(g.den);
            if (cosi > 0.05f) {
                const float delta = fabsf(g.t) * bp.cull_angle / cosi * 1.5f + 0.02f;   // lateral reach of the scatter (m)
                const float mu = 20.8f / bp.softness + delta * rsqrtf(p.uu), mv = 20.8f / bp.softness + delta * rsqrtf(p.vv);
                keep = (g.t > -delta - 0.1f) && (g.u > -mu) && (g.u < 1.0f + mu) && (g.v > -mv) && (g.v < 1.0f + mv);
            }
        }
        if (keep) mask |= (1ull << c);
    }
    return mask;
}

// blocked = 1 - exp(-alpha * sum_k sigma_k)
static __device__ __noinline__ float block_eval(const BlockPrim* prims, unsigned long long mask, const BlockParams& bp, float o0,
                                            float o1, float o2, float d0, float d1, float d2) {
    float sum = 0.0f;
    while (mask) {
        const int c = __ffsll((long long)mask) - 1;
        mask &= mask - 1;
        BlockGeom g;
        block_geometry(g, prims[c], bp, o0, o1, o2, d0, d1, d2);
        if (!block_relevant(g, bp)) continue;
        const float k = bp.softness;
        const float inside = sigmoidf_fast(k * g.u) * sigmoidf_fast(k * (1.0f - g.u)) * sigmoidf_fast(k * g.v) *
                             sigmoidf_fast(k * (1.0f - g.v));
        const float front = sigmoidf_fast(k * (g.t - bp.offset));
        sum += fminf(fmaxf(inside * front, 0.0f), 1.0f);
    }
    return 1.0f - __expf(-bp.alpha * sum);
}

// backward: given dL/dblocked, return blocked (recomputed) and d/d(origin), d/d(direction); d/d(primitive geometry) is
// accumulated atomically into grad_prims (rare: only rays inside a sigmoid transition contribute).  Out of line on
// purpose: the shadowed-ray path must not cost the unshadowed hot loop any registers.
struct BlockBack {
    float blocked, go0, go1, go2, gd0, gd1, gd2;
};

static __device__ __noinline__ BlockBack block_backward(const BlockPrim* prims, const int* cand_rows, unsigned long long mask,
                                                 const BlockParams bp, float o0, float o1, float o2, float d0, float d1,
                                                 float d2, float g_blocked_scale /* dL/dblocked */, float* grad_prims) {
    float go[3] = {0.f, 0.f, 0.f}, gd[3] = {0.f, 0.f, 0.f};
    BlockBack out;
    // first pass: optical depth
    float sum = 0.0f;
    unsigned long long m = mask;
    while (m) {
        const int c = __ffsll((long long)m) - 1;
        m &= m - 1;
        BlockGeom g;
        block_geometry(g, prims[c], bp, o0, o1, o2, d0, d1, d2);
        if (!block_relevant(g, bp)) continue;
        const float k = bp.softness;
        const float inside = sigmoidf_fast(k * g.u) * sigmoidf_fast(k * (1.0f - g.u)) * sigmoidf_fast(k * g.v) *
                             sigmoidf_fast(k * (1.0f - g.v));
        sum += fminf(fmaxf(inside * sigmoidf_fast(k * (g.t - bp.offset)), 0.0f), 1.0f);
    }
    const float transmittance = __expf(-bp.alpha * sum);
    out.blocked = 1.0f - transmittance;
    out.go0 = out.go1 = out.go2 = out.gd0 = out.gd1 = out.gd2 = 0.0f;
    // dL/dsum = dL/dblocked * alpha * transmittance
    const float g_sum = g_blocked_scale * bp.alpha * transmittance;
    if (g_sum == 0.0f) return out;
    m = mask;
    while (m) {
        const int c = __ffsll((long long)m) - 1;
        m &= m - 1;
        const BlockPrim& p = prims[c];
        BlockGeom g;
        block_geometry(g, p, bp, o0, o1, o2, d0, d1, d2);
        if (!block_relevant(g, bp)) continue;
        const float k = bp.softness;
        const float su0 = sigmoidf_fast(k * g.u), su1 = sigmoidf_fast(k * (1.0f - g.u));
        const float sv0 = sigmoidf_fast(k * g.v), sv1 = sigmoidf_fast(k * (1.0f - g.v));
        const float fr = sigmoidf_fast(k * (g.t - bp.offset));
        const float au = su0 * su1, av = sv0 * sv1;
        const float sigma = au * av * fr;
        if (!(sigma > 0.0f && sigma < 1.0f)) continue;   // clamp(0,1) passes gradient strictly inside
        const float g_u = g_sum * k * au * (su1 - su0) * av * fr;
        const float g_v = g_sum * k * av * (sv1 - sv0) * au * fr;
        const float g_t0 = g_sum * k * fr * (1.0f - fr) * au * av;
        if (fabsf(g_u) + fabsf(g_v) + fabsf(g_t0) == 0.0f) continue;
        const float idet = 1.0f / p.det;
        const float pu = g.off[0] * p.su[0] + g.off[1] * p.su[1] + g.off[2] * p.su[2];
        const float pv = g.off[0] * p.sv[0] + g.off[1] * p.sv[1] + g.off[2] * p.sv[2];
        const float g_pu = (g_u * p.vv - g_v * p.uv) * idet, g_pv = (g_v * p.uu - g_u * p.uv) * idet;
        float g_off[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) g_off[q] = g_pu * p.su[q] + g_pv * p.sv[q];
        const float dv[3] = {d0, d1, d2};
        const float g_t = g_t0 + g_off[0] * d0 + g_off[1] * d1 + g_off[2] * d2;
        const float iden = 1.0f / g.den;
        // t = ((c0 - o).n) / (d.n)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            go[q] += g_off[q] - g_t * p.n[q] * iden;
            gd[q] += g.t * g_off[q] - g_t * g.t * p.n[q] * iden;
        }
        if (grad_prims) {
            float* gp = grad_prims + (size_t)cand_rows[c] * 12;
            const float g_vv = g_u * pu * idet, g_uu = g_v * pv * idet, g_det = -(g_u * g.u + g_v * g.v) * idet;
            const float g_uv = -(g_u * pv + g_v * pu) * idet - 2.0f * p.uv * g_det;
            const float g_uu2 = g_uu + g_det * p.vv, g_vv2 = g_vv + g_det * p.uu;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                atomicAdd(gp + q, -g_off[q] + g_t * p.n[q] * iden);                                   // corner 0
                atomicAdd(gp + 3 + q, g_pu * g.off[q] + 2.0f * g_uu2 * p.su[q] + g_uv * p.sv[q]);      // span u
                atomicAdd(gp + 6 + q, g_pv * g.off[q] + 2.0f * g_vv2 * p.sv[q] + g_uv * p.su[q]);      // span v
                atomicAdd(gp + 9 + q, -g_t * g.off[q] * iden);                                        // normal
            }
            (void)dv;
        }
    }
    out.go0 = go[0]; out.go1 = go[1]; out.go2 = go[2];
    out.gd0 = gd[0]; out.gd1 = gd[1]; out.gd2 = gd[2];
    return out;
}

}  // namespace ab200
