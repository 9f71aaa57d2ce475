// Per-ray device math shared by the forward and backward trace kernels.
//
// Everything that decides WHICH pixel a ray lands in (reflect -> scatter -> intersection ->
// bitmap coordinates) is written with the strict one-rounding-per-op helpers of common.cuh in
// the evaluation order of the reference's eager ops (verified bit-equal
// against the reference on CPU by tests/golden).  Gradient math (backward only) is free to use
// FMAs.
#pragma once
#include "common.cuh"

namespace ab200 {

// atan2 in double precision, rounded once to float (the correctly rounded value)
__device__ __forceinline__ float atan2_rounded_once(float y, float x) { return (float)atan2((double)y, (double)x); }

// Per-CTA target constants (one heliostat-sample aims at one target area).
struct TargetCtx {
    int planar;
    float em1, um1;  // (E-1), (U-1) as float
    // planar: artist/raytracing/geometry.py:100-204
    float n0, n1, n2, c0, c1, c2, w, h, half_w, half_h;
    // cylindrical: artist/raytracing/geometry.py:287-445 (local frame rows u = n x axis, n, axis)
    float ux, uy, uz, ax0, ax1, ax2, rad2, opn, ang0;
    // pixel-per-metre scale used only for sizing the shared-memory window
    float px_per_m_e, px_per_m_u;
    // correctly rounded reciprocals of the two per-target divisors (planar: w, h; cylinder: opening, h) and whether the
    // 3-instruction exact quotient (const_div) may be used for them
    float rw, rh;
    int fastdiv;
};

// a / b for a divisor b that is constant per CTA, given rb = RN(1/b): q0 = RN(a*rb); r = a - q0*b (exact, FMA);
// q = RN(q0 + r*rb).  By Markstein's theorem q is the correctly rounded quotient for every a unless b's significand is
// all ones (checked once per target -> IEEE division path); inputs outside the normal range give the same inf/NaN
// classification, which is all the validity test needs.  tests/test_gpu_ops_parity.py checks it against __fdiv_rn.
__device__ __forceinline__ float const_div(float a, float b, float rb) {
    const float q0 = __fmul_rn(a, rb);
    const float r = __fmaf_rn(-q0, b, a);
    return __fmaf_rn(r, rb, q0);
}
__device__ __forceinline__ bool const_div_ok(float b) {
    const unsigned u = __float_as_uint(b);
    const unsigned ex = (u >> 23) & 0xffu;
    return ((u & 0x7fffffu) != 0x7fffffu) && ex > 40u && ex < 210u;
}

__device__ inline void load_target(TargetCtx& T, const ab200_targets& tg, int tidx, int res_e, int res_u) {
    T.em1 = (float)(res_e - 1);
    T.um1 = (float)(res_u - 1);
    // memory safety never depends on the caller's index values: out-of-range indices are clamped here (the host layer
    // raises IndexError for them, possibly one call later - ops._check_index_range)
    tidx = min(max(tidx, 0), tg.n_planar + tg.n_cyl - 1);
    if (tidx < tg.n_planar) {
        T.planar = 1;
        const float* n = tg.planar_normals + 4 * tidx;
        const float* c = tg.planar_centers + 4 * tidx;
        T.n0 = n[0]; T.n1 = n[1]; T.n2 = n[2];
        T.c0 = c[0]; T.c1 = c[1]; T.c2 = c[2];
        T.w = tg.planar_dims[2 * tidx];
        T.h = tg.planar_dims[2 * tidx + 1];
        T.half_w = sdiv(T.w, 2.0f);
        T.half_h = sdiv(T.h, 2.0f);
        T.px_per_m_e = T.em1 / T.w;
        T.px_per_m_u = T.um1 / T.h;
        T.rw = __frcp_rn(T.w);
        T.rh = __frcp_rn(T.h);
        T.fastdiv = const_div_ok(T.w) && const_div_ok(T.h);
    } else {
        const int k = tidx - tg.n_planar;
        T.planar = 0;
        const float* n = tg.cyl_normals + 4 * k;
        const float* a = tg.cyl_axes + 4 * k;
        const float* c = tg.cyl_centers + 4 * k;
        T.n0 = n[0]; T.n1 = n[1]; T.n2 = n[2];
        T.ax0 = a[0]; T.ax1 = a[1]; T.ax2 = a[2];
        T.c0 = c[0]; T.c1 = c[1]; T.c2 = c[2];
        // u = cross(normal, axis).  torch's CPU cross kernel evaluates a1*b2 - a2*b1 as fma(a1, b2, -RN(a2*b1)) (the compiler
        // contracts the subtraction in the AVX2/AVX-512 builds; checked against torch.cross on 3e5 random triples, 0 mismatches;
        // the all-unfused form differs in 25 % of them - one ulp of u moves every cylinder hit, tools/diag_cylinder_parity.py)
        T.ux = cross_comp(T.n1, T.ax2, T.n2, T.ax1);
        T.uy = cross_comp(T.n2, T.ax0, T.n0, T.ax2);
        T.uz = cross_comp(T.n0, T.ax1, T.n1, T.ax0);
        const float rad = tg.cyl_radii[k];
        T.rad2 = smul(rad, rad);
        T.h = tg.cyl_heights[k];
        T.half_h = sdiv(T.h, 2.0f);
        T.opn = tg.cyl_opening[k];
        T.w = T.opn;  // "width" of the unwrapped sector is the opening angle
        T.half_w = 0.f;
        T.ang0 = ssub(atan2_rounded_once(T.n1, T.n0), sdiv(T.opn, 2.0f));   // once per CTA: correctly rounded
        T.px_per_m_e = T.em1 / fmaxf(rad * T.opn, 1e-6f);
        T.px_per_m_u = T.um1 / T.h;
        T.rw = __frcp_rn(T.opn);
        T.rh = __frcp_rn(T.h);
        T.fastdiv = const_div_ok(T.opn) && const_div_ok(T.h);
    }
}

// Per-point context: origin, preferred reflection direction and the ray-independent terms.
struct PointCtx {
    float o0, o1, o2;   // ray origin (aligned surface point)
    float r0, r1, r2;   // preferred reflection direction (geometry.py:32-41)
    float dot;          // incident . normal (needed by the backward of reflect)
    float num;          // planar: (c - o) . n_t      (geometry.py:126-128)
    float ox, oy, oz;   // cylindrical: origin in the cylinder frame
    float cc;           // cylindrical: ox^2 + oy^2 - r^2
};

// origin-dependent, ray-independent terms (numerator of the plane distance / origin in the cylinder frame)
__device__ __forceinline__ void make_origin(PointCtx& pc, const TargetCtx& T, const float4 o) {
    pc.o0 = o.x; pc.o1 = o.y; pc.o2 = o.z;
    if (T.planar) {
        pc.num = sadd(sadd(smul(ssub(T.c0, o.x), T.n0), smul(ssub(T.c1, o.y), T.n1)), smul(ssub(T.c2, o.z), T.n2));
    } else {
        const float q0 = ssub(o.x, T.c0), q1 = ssub(o.y, T.c1), q2 = ssub(o.z, T.c2);
        // [P,3] @ [3,3]: the CPU GEMM accumulates as an FMA chain over k (measured)
        pc.ox = fmaf(q2, T.uz, fmaf(q1, T.uy, smul(q0, T.ux)));
        pc.oy = fmaf(q2, T.n2, fmaf(q1, T.n1, smul(q0, T.n0)));
        pc.oz = fmaf(q2, T.ax2, fmaf(q1, T.ax1, smul(q0, T.ax0)));
        pc.cc = ssub(sadd(smul(pc.ox, pc.ox), smul(pc.oy, pc.oy)), T.rad2);
    }
}

__device__ __forceinline__ void make_point(PointCtx& pc, const TargetCtx& T, float i0, float i1, float i2,
                                           const float4 o, const float4 n) {
    // dot over the 4 homogeneous components; incident.w * normal.w = 0 contributes +0
    const float dot = sadd(sadd(smul(i0, n.x), smul(i1, n.y)), smul(i2, n.z));
    pc.dot = dot;
    const float two_dot = smul(2.0f, dot);
    pc.r0 = ssub(i0, smul(two_dot, n.x));
    pc.r1 = ssub(i1, smul(two_dot, n.y));
    pc.r2 = ssub(i2, smul(two_dot, n.z));
    make_origin(pc, T, o);
}

// Where a CTA's surface points come from: already aligned [P] float4 rows, or - when the caller passes the sample's
// orientation (ab200_trace_args::orientations) - the un-aligned rows, which are rotated here exactly as
// ab200_align_fwd does (out_j = sum_k d_k * O[j][k] accumulated as an FMA chain over k, like the CPU GEMM of
// heliostat_group_rigid_body.py:217-222), so the aligned [N,P,4] tensors never exist in HBM.
struct PointSrc {
    const float4* pts;
    const float4* nrm;
    const float* O;   // shared memory, row-major 4x4, or nullptr
    // rows of the sample the CTA one wave later will start with (or nullptr): the threads that will read them for the
    // window placement pull them into L2 during their last point, so that CTA's start-up waits for L2, not for DRAM
    // (kept in shared memory: {points, normals} of that sample; no registers inside the ray loops)
    const float4* const* next = nullptr;
};

__device__ __forceinline__ float4 apply_orientation(const float* O, const float4 d) {
    float4 r;
    r.x = fmaf(d.w, O[3], fmaf(d.z, O[2], fmaf(d.y, O[1], smul(d.x, O[0]))));
    r.y = fmaf(d.w, O[7], fmaf(d.z, O[6], fmaf(d.y, O[5], smul(d.x, O[4]))));
    r.z = fmaf(d.w, O[11], fmaf(d.z, O[10], fmaf(d.y, O[9], smul(d.x, O[8]))));
    r.w = 0.f;   // the trace never reads the homogeneous component
    return r;
}

__device__ __forceinline__ void orient_point(const PointSrc& s, float4& o, float4& n) {
    if (s.O) { o = apply_orientation(s.O, o); n = apply_orientation(s.O, n); }
}

// Backward of orient_point for one surface point: (gp, gn) arrive as gradients w.r.t. the ALIGNED point / normal and
// leave as gradients w.r.t. the un-aligned rows (g @ O); `acc` (12 registers, rows 0..2 of dL/dO, or nullptr) collects
// sum_p g_j * d_k over the thread's points.  d, e are the un-aligned point / normal.
__device__ __forceinline__ void orient_point_backward(const PointSrc& s, const float4 d, const float4 e, float4& gp,
                                                      float4& gn, float* acc) {
    if (!s.O) return;
    if (acc) {
        const float g[3] = {gp.x, gp.y, gp.z}, m[3] = {gn.x, gn.y, gn.z};
        const float dd[4] = {d.x, d.y, d.z, d.w}, ee[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[j * 4 + k] += g[j] * dd[k] + m[j] * ee[k];
    }
    const float* O = s.O;
    const float4 a = gp, b = gn;
    gp = make_float4(a.x * O[0] + a.y * O[4] + a.z * O[8], a.x * O[1] + a.y * O[5] + a.z * O[9],
                     a.x * O[2] + a.y * O[6] + a.z * O[10], a.x * O[3] + a.y * O[7] + a.z * O[11]);
    gn = make_float4(b.x * O[0] + b.y * O[4] + b.z * O[8], b.x * O[1] + b.y * O[5] + b.z * O[9],
                     b.x * O[2] + b.y * O[6] + b.z * O[10], b.x * O[3] + b.y * O[7] + b.z * O[11]);
}

// Scatter: d = M(e,u) * r   (geometry/transforms.py:67-74, heliostat_ray_tracer.py:547-552)
struct Scatter {
    float cu, su, ce, se;
    float m10, m11, m20, m21;
    float dx, dy, dz;
};

__device__ __forceinline__ void scatter(Scatter& s, const PointCtx& pc) {
    s.m10 = smul(s.ce, s.su);
    s.m11 = smul(s.ce, s.cu);
    s.m20 = smul(s.se, s.su);
    s.m21 = smul(s.se, s.cu);
    s.dx = sadd(smul(s.cu, pc.r0), smul(-s.su, pc.r1));
    s.dy = sadd(sadd(smul(s.m10, pc.r0), smul(s.m11, pc.r1)), smul(-s.se, pc.r2));
    s.dz = sadd(sadd(smul(s.m20, pc.r0), smul(s.m21, pc.r1)), smul(s.ce, pc.r2));
}

// Result of one ray / target intersection in bitmap coordinates.
struct Hit {
    float be, bu;    // continuous bitmap coordinates (be already flipped for planar targets)
    float t;         // intersection distance
    float lam;       // ray_magnitude * Lambert cosine (0 if invalid)
    bool valid;
    // planar intermediates for the backward
    float a;
    // cylindrical intermediates for the backward
    float dlx, dly, dlz, x, y, qa, qb, sq, nrm;
    int near_root;
};

template <bool FASTDIV>
__device__ __forceinline__ void hit_planar(Hit& h, const TargetCtx& T, const PointCtx& pc, const Scatter& s, float mag) {
    const float a = sadd(sadd(smul(s.dx, T.n0), smul(s.dy, T.n1)), smul(s.dz, T.n2));
    const bool ff = a < 0.0f;
    const float t = ff ? sdiv(pc.num, a) : 0.0f;
    const float X = sadd(pc.o0, smul(s.dx, t));
    const float Z = sadd(pc.o2, smul(s.dz, t));
    const float te = ssub(sadd(X, T.half_w), T.c0);
    const float tu = ssub(sadd(Z, T.half_h), T.c2);
    const float be0 = smul(FASTDIV ? const_div(te, T.w, T.rw) : sdiv(te, T.w), T.em1);
    const float bu0 = smul(FASTDIV ? const_div(tu, T.h, T.rh) : sdiv(tu, T.h), T.um1);
    const bool valid = ff && (0.0f <= be0) && (be0 <= T.em1) && (0.0f <= bu0) && (bu0 <= T.um1);
    h.a = a;
    h.valid = valid;
    h.t = valid ? t : 0.0f;
    h.be = ssub(T.em1, valid ? be0 : 0.0f);
    h.bu = valid ? bu0 : 0.0f;
    h.lam = valid ? smul(mag, -a) : 0.0f;
}

// raw (unmasked) planar coordinates of the undistorted ray, used only to place the smem window
__device__ __forceinline__ bool centre_planar(const TargetCtx& T, const PointCtx& pc, float& be, float& bu, float& t,
                                              float& cosi) {
    const float a = pc.r0 * T.n0 + pc.r1 * T.n1 + pc.r2 * T.n2;
    if (!(a < 0.0f)) return false;
    t = pc.num / a;
    be = T.em1 - ((pc.o0 + pc.r0 * t) + T.half_w - T.c0) / T.w * T.em1;
    bu = ((pc.o2 + pc.r2 * t) + T.half_h - T.c2) / T.h * T.um1;
    cosi = -a;
    return true;
}

// CRATAN: the sector angle through a double-precision atan2 rounded once (the correctly rounded float, which torch's CPU
// atan2 returns for 97.7 % of its arguments; atan2f agrees with it for about half).  Per-ray double arithmetic: the strict
// parity mode (AB200_TRIG_TABLE) pays it, the device-trig modes keep atan2f.
template <bool FASTDIV, bool CRATAN = false>
__device__ __forceinline__ void hit_cylinder(Hit& h, const TargetCtx& T, const PointCtx& pc, const Scatter& s, float mag) {
    // directions @ rot^T (FMA chain over k, as the CPU GEMM does)
    const float dlx = fmaf(s.dz, T.uz, fmaf(s.dy, T.uy, smul(s.dx, T.ux)));
    const float dly = fmaf(s.dz, T.n2, fmaf(s.dy, T.n1, smul(s.dx, T.n0)));
    const float dlz = fmaf(s.dz, T.ax2, fmaf(s.dy, T.ax1, smul(s.dx, T.ax0)));
    const float qa = sadd(smul(dlx, dlx), smul(dly, dly));
    const float qb = smul(2.0f, sadd(smul(pc.ox, dlx), smul(pc.oy, dly)));
    const float disc = ssub(smul(qb, qb), smul(smul(4.0f, qa), pc.cc));
    const bool hits = (disc >= 0.0f) && (fabsf(qa) > 1e-8f);
    const float sq = sqrtf(sadd(hits ? disc : smul(disc, 0.0f), 1e-12f));
    const float two_a = smul(2.0f, qa);
    float t1 = sdiv(ssub(-qb, sq), two_a);
    float t2 = sdiv(sadd(-qb, sq), two_a);
    const float inf = __int_as_float(0x7f800000);
    t1 = (t1 > 0.0f) ? t1 : inf;
    t2 = (t2 > 0.0f) ? t2 : inf;
    float t = fminf(t1, t2);
    const int near_root = (t1 <= t2);
    const bool vd = hits && (t < inf) && (t == t);
    t = vd ? t : 0.0f;
    const float x = sadd(pc.ox, smul(t, dlx));
    const float y = sadd(pc.oy, smul(t, dly));
    float z = sadd(pc.oz, smul(t, dlz));
    const float nrm = sqrtf(fmaf(y, y, smul(x, x)));   // torch.norm over (x, y, 0): the FMA chain fma(0,0, fma(y,y, x*x))
    const float nlx = sdiv(x, nrm), nly = sdiv(y, nrm);
    float lam = sadd(smul(-dlx, nlx), smul(-dly, nly));
    lam = fmaxf(lam, 0.0f);
    z = sadd(z, T.half_h);
    const float ang = ssub(CRATAN ? atan2_rounded_once(y, x) : atan2f(y, x), T.ang0);
    const bool on = (z >= 0.0f) && (z <= T.h) && (ang >= 0.0f) && (ang <= T.opn);
    const bool valid = on && vd;
    h.valid = valid;
    h.bu = valid ? smul(FASTDIV ? const_div(z, T.h, T.rh) : sdiv(z, T.h), T.um1) : 0.0f;
    h.be = valid ? smul(FASTDIV ? const_div(ang, T.opn, T.rw) : sdiv(ang, T.opn), T.em1) : 0.0f;
    h.t = valid ? t : 0.0f;
    h.lam = valid ? smul(mag, lam) : 0.0f;
    h.dlx = dlx; h.dly = dly; h.dlz = dlz; h.x = x; h.y = y; h.qa = qa; h.qb = qb; h.sq = sq; h.nrm = nrm;
    h.near_root = near_root;
    h.a = 0.f;
}

__device__ __forceinline__ bool centre_cylinder(const TargetCtx& T, const PointCtx& pc, float& be, float& bu, float& t,
                                                float& cosi) {
    Scatter s;
    s.dx = pc.r0; s.dy = pc.r1; s.dz = pc.r2;
    Hit h;
    hit_cylinder<false>(h, T, pc, s, 1.0f);
    if (!h.valid) return false;
    be = h.be; bu = h.bu; t = h.t; cosi = fmaxf(h.lam, 0.05f);
    return true;
}

// Bilinear splat weights (heliostat_ray_tracer.py:674-728).  Precondition: the ray is valid, so 0 <= be <= E-1 and
// 0 <= bu <= U-1 and the lower-bound tests of the reference's on-target mask are always true.
struct Splat {
    int ie, iu;
    bool on;
    float wle, whe, wlu, whu;
};

__device__ __forceinline__ void splat_weights(Splat& sp, float be, float bu, int res_e, int res_u) {
    const float fe = truncf(be), fu = truncf(bu);  // == tensor.long() for non-negative coordinates
    sp.ie = __float2int_rz(be);
    sp.iu = __float2int_rz(bu);
    sp.on = (sp.ie + 1 < res_e) && (sp.iu + 1 < res_u);
    sp.wle = ssub(sadd(fe, 1.0f), be);   // (ie + 1) - be: ie + 1 is exact in fp32
    sp.wlu = ssub(sadd(fu, 1.0f), bu);
    sp.whe = ssub(be, fe);
    sp.whu = ssub(bu, fu);
}

}  // namespace ab200
