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

// packed primitive: 16 floats in global memory (ab200_blocking_pack) + 8 derived ones filled in when a CTA copies its
// candidates to shared memory (block_finish_prim)
struct BlockPrimPacked {    // row of the global primitive table
    float c0[3], su[3], sv[3], n[3];
    float uu, vv, uv, det;  // |su|^2, |sv|^2, su.sv, safe determinant
};
// shared-memory copy: what the per-point cull and the per-ray classification read comes first, as four 16-byte words
// (c0 | n | U | V | ilu ilv det uv), the rest is only touched by the out-of-line soft-mask evaluation
struct __align__(16) BlockPrim {
    float c0[3], n[3];
    float U[3], V[3];       // rectangle coordinates as dot products: u = off . U, v = off . V
    float ilu, ilv;         // 1 / |su|, 1 / |sv|
    float det, uv;
    float su[3], uu, sv[3], vv;
};
constexpr int kPrimFloats = 16, kPrimFloatsShared = 24;
static_assert(sizeof(BlockPrimPacked) == kPrimFloats * 4 && sizeof(BlockPrim) == kPrimFloatsShared * 4, "BlockPrim layout");

__device__ __forceinline__ void block_finish_prim(BlockPrim& p) {
    const float idet = 1.0f / p.det;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        p.U[q] = (p.su[q] * p.vv - p.sv[q] * p.uv) * idet;
        p.V[q] = (p.sv[q] * p.uu - p.su[q] * p.uv) * idet;
    }
    p.ilu = rsqrtf(p.uu); p.ilv = rsqrtf(p.vv);
}

// One CTA's candidate list -> shared memory (all threads of the CTA call this; ends with a barrier).
template <int THREADS>
__device__ __forceinline__ int block_load_candidates(BlockPrim* blk_sh, int* rows_sh, const ab200_blockers& B, int h, int tid) {
    const int n = min(B.cand_count[h], kMaxBlockCandidates);
    const int* cand = B.cand_idx + (size_t)h * B.max_candidates;
    if (rows_sh) for (int i = tid; i < n; i += THREADS) rows_sh[i] = cand[i];
    // packed row (c0 su sv n uu vv uv det) -> float index in BlockPrim
    for (int i = tid; i < n * kPrimFloats; i += THREADS) {
        const int c = i / kPrimFloats, k = i % kPrimFloats;
        const int dst = k < 3 ? k : k < 6 ? 16 + (k - 3) : k < 9 ? 20 + (k - 6) : k < 12 ? 3 + (k - 9) : k == 12 ? 19 : k == 13 ? 23 : k == 14 ? 15 : 14;
        reinterpret_cast<float*>(blk_sh)[c * kPrimFloatsShared + dst] = B.prims[(size_t)cand[c] * kPrimFloats + k];
    }
    __syncthreads();
    for (int c = tid; c < n; c += THREADS) block_finish_prim(blk_sh[c]);
    __syncthreads();
    return n;
}

// explicit shared-memory loads of a primitive's hot words (the pointer travels through context structs and out-of-line
// calls, where the compiler falls back to generic loads)
__device__ __forceinline__ float4 lds128(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
struct BlockHot {   // c0, n, U, V
    float c0x, c0y, c0z, nx, ny, nz, Ux, Uy, Uz, Vx, Vy, Vz;
};
__device__ __forceinline__ BlockHot block_hot(unsigned prim_addr) {
    const float4 a = lds128(prim_addr), b = lds128(prim_addr + 16), c = lds128(prim_addr + 32);
    return BlockHot{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
}
// ray / rectangle-plane intersection in rectangle coordinates: t, u, v and d . n
__device__ __forceinline__ void block_tuv(const BlockHot& p, float eps, float o0, float o1, float o2, float d0, float d1, float d2,
                                          float& t, float& u, float& v, float& den) {
    den = d0 * p.nx + d1 * p.ny + d2 * p.nz;
    if (fabsf(den) < eps) den = den >= 0.0f ? eps : -eps;
    const float b0 = o0 - p.c0x, b1 = o1 - p.c0y, b2 = o2 - p.c0z;
    t = __fdividef(-(b0 * p.nx + b1 * p.ny + b2 * p.nz), den);
    const float f0 = fmaf(t, d0, b0), f1 = fmaf(t, d1, b1), f2 = fmaf(t, d2, b2);
    u = f0 * p.Ux + f1 * p.Uy + f2 * p.Uz;
    v = f0 * p.Vx + f1 * p.Vy + f2 * p.Vz;
}

// The same intersection in the reference's own operation order (blocking.py:318-346: every product and sum its own
// rounding, sums over the three components left to right, IEEE divisions) - torch's CPU evaluation reproduced bit for bit
// (checked op by op against torch 2.11 CPU).  The soft mask multiplies the rectangle coordinates by softness (1000) and
// the optical depth by alpha (100): one ulp of u or v is up to 1e-4 of `blocked`, so the rays that are actually
// evaluated (block_eval / block_backward, a few per cent) take this path; culling and classification, which only
// compare against margins, keep the short form above.
__device__ __forceinline__ float dot3_strict(float a0, float a1, float a2, float b0, float b1, float b2) {
    return sadd(sadd(smul(a0, b0), smul(a1, b1)), smul(a2, b2));
}
__device__ __forceinline__ void block_tuv_strict(const BlockPrim& p, float eps, float o0, float o1, float o2, float d0, float d1,
                                                 float d2, float& t, float& u, float& v, float& den) {
    den = dot3_strict(d0, d1, d2, p.n[0], p.n[1], p.n[2]);
    if (fabsf(den) < eps) den = den >= 0.0f ? eps : -eps;
    const float num = dot3_strict(ssub(p.c0[0], o0), ssub(p.c0[1], o1), ssub(p.c0[2], o2), p.n[0], p.n[1], p.n[2]);
    t = sdiv(num, den);
    const float f0 = ssub(sadd(o0, smul(t, d0)), p.c0[0]), f1 = ssub(sadd(o1, smul(t, d1)), p.c0[1]),
                f2 = ssub(sadd(o2, smul(t, d2)), p.c0[2]);
    const float pu = dot3_strict(f0, f1, f2, p.su[0], p.su[1], p.su[2]), pv = dot3_strict(f0, f1, f2, p.sv[0], p.sv[1], p.sv[2]);
    u = sdiv(ssub(smul(pu, p.vv), smul(pv, p.uv)), p.det);
    v = sdiv(ssub(smul(pv, p.uu), smul(pu, p.uv)), p.det);
}

struct BlockParams {
    float softness, alpha, offset, epsilon;
    float cull_angle;  // bound on the scatter angle used by the per-point cull (<= 0: no per-point culling)
};

__device__ __forceinline__ float sigmoidf_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
// torch.sigmoid as the CPU evaluates it, 1 / (1 + exp(-x)), with the accurate expf and an IEEE division (within an ulp or
// two of torch's vectorised exp)
#ifdef AB200_BLOCK_SIGMOID_IEEE
__device__ __forceinline__ float sigmoidf_ref(float x) { return sdiv(1.0f, sadd(1.0f, expf(-x))); }
#else
__device__ __forceinline__ float sigmoidf_ref(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
#endif

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

// Cheap per-ray classification of the point's candidates (the hot loops call this, inline; everything else in this file is
// out of line and only runs for the few rays inside a sigmoid transition):
//   0  no candidate matters: blocked == 0
//   1  some candidate shadows the ray completely: every sigmoid argument is >= 20.8, each factor rounds to 1.0f, the optical
//      depth is >= 1 and blocked = 1 - exp(-alpha) == 1.0f in fp32 (alpha = 100: exp(-100) is below half an ulp of 1) - the
//      ray carries no intensity and no gradient, exactly as in the reference's fp32 evaluation
//   2  the ray is inside a transition band of some candidate: evaluate block_eval / block_backward
// q = min(u, 1-u, v, 1-v, t - offset) is the ray's depth inside the (rectangle x half-line): relevant iff q > -m, saturated
// iff q > m, with m = 20.8 / softness.
__device__ __forceinline__ int block_classify(const BlockPrim* prims, unsigned long long mask, const BlockParams& bp, float o0,
                                              float o1, float o2, float d0, float d1, float d2) {
    const float m = 20.8f / bp.softness;
    const bool can_saturate = bp.alpha >= 17.0f;   // exp(-alpha) < 2^-25
    const unsigned base = (unsigned)__cvta_generic_to_shared(prims);
    int cls = 0;
    while (mask) {
        const int c = __ffsll((long long)mask) - 1;
        mask &= mask - 1;
        const BlockHot p = block_hot(base + (unsigned)c * (kPrimFloatsShared * 4));
        float t, u, v, den;
        block_tuv(p, bp.epsilon, o0, o1, o2, d0, d1, d2, t, u, v, den);
        const float q = fminf(fminf(fminf(u, 1.0f - u), fminf(v, 1.0f - v)), t - bp.offset);
        if (q > -m) {
            if (q > m && can_saturate) return 1;
            cls = 2;
        }
    }
    return cls;
}

// a ray this far outside the rectangle (or this close to / behind its origin) contributes < 1e-9
__device__ __forceinline__ bool block_relevant(const BlockGeom& g, const BlockParams& bp) {
    const float m = 20.8f / bp.softness;
    return (g.t > bp.offset - m) && (g.u > -m) && (g.u < 1.0f + m) && (g.v > -m) && (g.v < 1.0f + m);
}

// per-point cull of the candidate list against the undistorted reflection direction: bit c of the result is set if
// candidate c may matter for some ray of this point
// `dead` (optional): set when some candidate certainly shadows EVERY ray of the point completely - the undistorted
// reflection hits it deeper inside than the scatter can reach (same margins as the keep test, mirrored), at a decent
// incidence - so that each ray would classify as 1 (blocked == 1.0f): no intensity, no gradient.
__device__ __forceinline__ unsigned long long block_point_mask(const BlockPrim* prims, int n_cand, const BlockParams& bp,
                                                               float o0, float o1, float o2, float r0, float r1, float r2,
                                                               bool* dead = nullptr) {
    unsigned long long mask = 0ull;
    const unsigned base = (unsigned)__cvta_generic_to_shared(prims);
    const float m = 20.8f / bp.softness;
    bool all_dead = false;
    for (int c = 0; c < n_cand; ++c) {
        bool keep = true;
        if (bp.cull_angle > 0.0f) {
            const unsigned addr = base + (unsigned)c * (kPrimFloatsShared * 4);
            const BlockHot p = block_hot(addr);
            float t, u, v, den;
            block_tuv(p, bp.epsilon, o0, o1, o2, r0, r1, r2, t, u, v, den);
            const float cosi = fabsf(den);
            if (cosi > 0.05f) {
                const float4 w = lds128(addr + 48);   // ilu, ilv, det, uv
                const float delta = fabsf(t) * bp.cull_angle * __fdividef(1.5f, cosi) + 0.02f;   // lateral reach of the scatter (m)
                const float mu = fmaf(delta, w.x, m), mv = fmaf(delta, w.y, m);
                keep = (t > -delta - 0.1f) && (u > -mu) && (u < 1.0f + mu) && (v > -mv) && (v < 1.0f + mv);
                all_dead |= (cosi > 0.2f) && (bp.alpha >= 17.0f) && (t > bp.offset + m + 0.15f + delta) &&
                            (fminf(u, 1.0f - u) > mu) && (fminf(v, 1.0f - v) > mv);
            }
        }
        if (keep) mask |= (1ull << c);
    }
    if (dead) *dead = all_dead;
    return mask;
}

// blocked = 1 - exp(-alpha * sum_k sigma_k)
static __device__ __noinline__ float block_eval(const BlockPrim* prims, unsigned long long mask, const BlockParams& bp, float o0,
                                            float o1, float o2, float d0, float d1, float d2) {
    const unsigned base = (unsigned)__cvta_generic_to_shared(prims);
    const float k = bp.softness, m = 20.8f / k;
    float sum = 0.0f;
    while (mask) {
        const int c = __ffsll((long long)mask) - 1;
        mask &= mask - 1;
        float t, u, v, den;
        block_tuv_strict(prims[c], bp.epsilon, o0, o1, o2, d0, d1, d2, t, u, v, den);
        if (!(fminf(fminf(fminf(u, 1.0f - u), fminf(v, 1.0f - v)), t - bp.offset) > -m)) continue;   // contributes < 1e-9
        // blocking.py:347-353 in its order: ((s(ku) s(k(1-u))) s(kv)) s(k(1-v)), times the front gate, clamped, summed over k
        const float inside = smul(smul(smul(sigmoidf_ref(smul(k, u)), sigmoidf_ref(smul(k, ssub(1.0f, u)))), sigmoidf_ref(smul(k, v))),
                                  sigmoidf_ref(smul(k, ssub(1.0f, v))));
        const float front = sigmoidf_ref(smul(k, ssub(t, bp.offset)));
        sum = sadd(sum, fminf(fmaxf(smul(inside, front), 0.0f), 1.0f));
    }
    return ssub(1.0f, expf(-smul(bp.alpha, sum)));
}

// ---------------------------------------------------------------------------------------------------------------------
// Deferral of shadow-affected points.  The ray loops of the trace kernels run at the register budget of the unblocked
// kernels only if they contain no blocking code at all (measured: 68 local loads + 35 local stores per ray and a 8x slower
// backward with the soft mask inline, profiles/r02_motor_ncu_before.txt).  So a CTA traces its sample in two passes:
//   1. the ordinary loops (compiled without blocking) over the points whose candidate mask is EMPTY - the per-point cull
//      (block_point_mask, out of line, once per point) decides; the other points are only recorded in a shared bit set;
//   2. the recorded points, spread evenly over the CTA's threads whatever their position on the mirror (shadows are
//      compact blobs: the round-robin point -> thread map of pass 1 would leave most lanes idle), through the generic
//      loops with the per-ray classification / soft mask.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kDeferWords = 512;   // 16384 points per CTA (chunk); larger chunks evaluate the mask inline

struct DeferCtx {
    unsigned* bits;     // shared [kDeferWords]: bit (p - p_begin) set = point deferred to pass 2
    int* prefix;        // shared [kDeferWords]: exclusive prefix sum of the words' popcounts (defer_scan)
    int n_words, count; // words in use; number of deferred points (after defer_scan)
    const BlockPrim* blk;
    int n_blk;
    BlockParams bp;
};

// pass 1, per point: kPointClear = trace it here; kPointDeferred = recorded, the caller skips the point; kPointDead = every
// ray is completely shadowed (the caller traces it without taps / writes zero gradients)
enum { kPointClear = 0, kPointDeferred = 1, kPointDead = 2 };
__device__ __forceinline__ int defer_point(const DeferCtx* dc, int local_index, float o0, float o1, float o2, float r0, float r1,
                                           float r2) {
    if (dc == nullptr || dc->n_blk == 0) return kPointClear;
    bool dead;
    if (block_point_mask(dc->blk, dc->n_blk, dc->bp, o0, o1, o2, r0, r1, r2, &dead) == 0ull) return kPointClear;
    if (dead) return kPointDead;
    atomicOr(dc->bits + (local_index >> 5), 1u << (local_index & 31));
    return kPointDeferred;
}

// all threads of the CTA; returns the number of deferred points
template <int THREADS>
__device__ __forceinline__ int defer_scan(DeferCtx& dc, int* total_sh) {
    __syncthreads();   // pass 1 complete
    if (threadIdx.x < 32) {
        const int per_lane = (dc.n_words + 31) / 32, w0 = threadIdx.x * per_lane;
        int sum = 0;
        for (int w = w0; w < min(w0 + per_lane, dc.n_words); ++w) sum += __popc(dc.bits[w]);
        int incl = sum;
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if ((int)threadIdx.x >= d) incl += v;
        }
        int run = incl - sum;
        for (int w = w0; w < min(w0 + per_lane, dc.n_words); ++w) { dc.prefix[w] = run; run += __popc(dc.bits[w]); }
        if (threadIdx.x == 31) *total_sh = incl;
    }
    __syncthreads();
    dc.count = *total_sh;
    return dc.count;
}

// local index of the k-th deferred point (0 <= k < count)
__device__ __forceinline__ int defer_lookup(const DeferCtx& dc, int k) {
    int lo = 0, hi = dc.n_words - 1;
    while (lo < hi) {   // last word with prefix <= k
        const int mid = (lo + hi + 1) >> 1;
        if (dc.prefix[mid] <= k) lo = mid; else hi = mid - 1;
    }
    return lo * 32 + (int)__fns(dc.bits[lo], 0, k - dc.prefix[lo] + 1);
}

// backward: given dL/dblocked, return blocked (recomputed) and d/d(origin), d/d(direction); d/d(primitive geometry) is
// accumulated atomically into grad_prims (rare: only rays inside a sigmoid transition contribute).  Out of line on
// purpose: the shadowed-ray path must not cost the unshadowed hot loop any registers.
struct BlockBack {
    float blocked, go0, go1, go2, gd0, gd1, gd2;
};

// `gacc` (shared, may be NULL): the CTA's accumulators of d/d(corner0, span_u, span_v, normal) per candidate SLOT
// ([n_cand][12] doubles; double so that the order of the adds - the one order-dependent step - practically never reaches
// the float the sums are rounded to).
static __device__ __noinline__ BlockBack block_backward(const BlockPrim* prims, unsigned long long mask,
                                                 const BlockParams bp, float o0, float o1, float o2, float d0, float d1,
                                                 float d2, float g_blocked_scale /* dL/dblocked */, double* gacc,
                                                 int gacc_copies, int gacc_stride) {
    float go[3] = {0.f, 0.f, 0.f}, gd[3] = {0.f, 0.f, 0.f};
    BlockBack out;
    // first pass: optical depth (cheap rectangle coordinates; the candidates that matter are remembered)
    const unsigned base = (unsigned)__cvta_generic_to_shared(prims);
    const float k = bp.softness, m_rel = 20.8f / k;
    float sum = 0.0f;
    unsigned long long m = mask, relevant = 0ull;
    while (m) {
        const int c = __ffsll((long long)m) - 1;
        m &= m - 1;
        float t, u, v, den;
        block_tuv_strict(prims[c], bp.epsilon, o0, o1, o2, d0, d1, d2, t, u, v, den);   // exactly block_eval's sequence
        if (!(fminf(fminf(fminf(u, 1.0f - u), fminf(v, 1.0f - v)), t - bp.offset) > -m_rel)) continue;
        relevant |= 1ull << c;
        const float inside = smul(smul(smul(sigmoidf_ref(smul(k, u)), sigmoidf_ref(smul(k, ssub(1.0f, u)))), sigmoidf_ref(smul(k, v))),
                                  sigmoidf_ref(smul(k, ssub(1.0f, v))));
        sum = sadd(sum, fminf(fmaxf(smul(inside, sigmoidf_ref(smul(k, ssub(t, bp.offset)))), 0.0f), 1.0f));
    }
    const float transmittance = expf(-smul(bp.alpha, sum));
    out.blocked = ssub(1.0f, transmittance);
    out.go0 = out.go1 = out.go2 = out.gd0 = out.gd1 = out.gd2 = 0.0f;
    // dL/dsum = dL/dblocked * alpha * transmittance
    const float g_sum = g_blocked_scale * bp.alpha * transmittance;
    if (g_sum == 0.0f) return out;
    m = relevant;
    while (m) {
        const int c = __ffsll((long long)m) - 1;
        m &= m - 1;
        const BlockPrim& p = prims[c];
        BlockGeom g;
        {   // rectangle coordinates as in the first pass (and in block_eval), plus the offset vector
            const BlockHot hp = block_hot(base + (unsigned)c * (kPrimFloatsShared * 4));
            block_tuv_strict(p, bp.epsilon, o0, o1, o2, d0, d1, d2, g.t, g.u, g.v, g.den);   // the forward's values
            g.off[0] = fmaf(g.t, d0, o0 - hp.c0x); g.off[1] = fmaf(g.t, d1, o1 - hp.c0y); g.off[2] = fmaf(g.t, d2, o2 - hp.c0z);
        }
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
        const float idet = __fdividef(1.0f, p.det);
        const float pu = g.off[0] * p.su[0] + g.off[1] * p.su[1] + g.off[2] * p.su[2];
        const float pv = g.off[0] * p.sv[0] + g.off[1] * p.sv[1] + g.off[2] * p.sv[2];
        const float g_pu = (g_u * p.vv - g_v * p.uv) * idet, g_pv = (g_v * p.uu - g_u * p.uv) * idet;
        float g_off[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) g_off[q] = g_pu * p.su[q] + g_pv * p.sv[q];
        const float dv[3] = {d0, d1, d2};
        const float g_t = g_t0 + g_off[0] * d0 + g_off[1] * d1 + g_off[2] * d2;
        const float iden = __fdividef(1.0f, g.den);
        // t = ((c0 - o).n) / (d.n)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            go[q] += g_off[q] - g_t * p.n[q] * iden;
            gd[q] += g.t * g_off[q] - g_t * g.t * p.n[q] * iden;
        }
        if (gacc) {
            // `gacc` holds `gacc_copies` interleaved copies of the [n_slots][12] accumulators (the 64 x 12 cells are shared out
            // among as many copies as fit): threads spread over the copies, because every band ray of a CTA adds to the same
            // dozen cells and shared-memory double adds are compare-and-swap loops
            double* gp = gacc + ((threadIdx.x ^ (threadIdx.x >> 5)) & (gacc_copies - 1)) * gacc_stride + c * 12;
            const float g_vv = g_u * pu * idet, g_uu = g_v * pv * idet, g_det = -(g_u * g.u + g_v * g.v) * idet;
            const float g_uv = -(g_u * pv + g_v * pu) * idet - 2.0f * p.uv * g_det;
            const float g_uu2 = g_uu + g_det * p.vv, g_vv2 = g_vv + g_det * p.uu;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                atomicAdd(gp + q, (double)(-g_off[q] + g_t * p.n[q] * iden));                                   // corner 0
                atomicAdd(gp + 3 + q, (double)(g_pu * g.off[q] + 2.0f * g_uu2 * p.su[q] + g_uv * p.sv[q]));      // span u
                atomicAdd(gp + 6 + q, (double)(g_pv * g.off[q] + 2.0f * g_vv2 * p.sv[q] + g_uv * p.su[q]));      // span v
                atomicAdd(gp + 9 + q, (double)(-g_t * g.off[q] * iden));                                        // normal
            }
            (void)dv;
        }
    }
    out.go0 = go[0]; out.go1 = go[1]; out.go2 = go[2];
    out.gd0 = gd[0]; out.gd1 = gd[1]; out.gd2 = gd[2];
    return out;
}

}  // namespace ab200
