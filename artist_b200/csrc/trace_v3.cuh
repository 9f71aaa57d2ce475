// "v3" trace kernels: the benchmark-shaped fast path (planar target with the exact constant-divisor quotient,
// polynomial trig, fixed-point bitmap, no blocking, bitmap width <= 256, one CTA per heliostat-sample).
// Included by trace.cu (inside namespace ab200, after TraceParams / Window / FwdCtx / the generic ray loops).
//
// What differs from the fast2 loops of trace.cu (results are bit-identical, asserted by the parity tests):
//  * POINT-pair packing: the two fp32x2 lanes are two ADJACENT surface points (p, p+1) tracing the same ray index, not
//    two rays of one point.  One 16-byte load brings both rays' (u, e); the per-point prologue (orientation, reflection,
//    plane numerator) and, in the backward, the per-point epilogue run packed too; no odd-R / tail special cases.
//  * FULL-ROW window: the shared-memory window is `wh` complete bitmap rows with a fixed pitch of 260 cells, so a tap
//    address is  base + (iu - u0) * 1040 + ie * 4  = one multiply-add and one shift-add on the bit patterns the floor()
//    trick leaves behind (no index subtraction, the row above at an immediate +1040), the column test is the on-bitmap test,
//    and the flush / staging move whole rows.  224 KB hold 220 rows: bitmaps up to 220 rows never leave shared memory.
//  * Four rays per iteration (two ray indices x two points), the rays-per-point loop unrolled at compile time for the
//    reference's default R = 10, distortion rows prefetched two iterations ahead, the fast taps predicated instead of
//    branched, one rarely taken region per ray row for everything else (invalid / irregular rays, taps outside the window).
#pragma once

namespace v3 {

constexpr int kMaxE = 256;             // widest bitmap the full-row window supports
// Window row pitch in cells.  NOT a multiple of 32: neighbouring surface points of a warp reflect to a (near) vertical
// line of pixels, and with a 1 KB pitch all of a warp's taps would hit one bank.  260 = 4 banks per row keeps rows
// 16-byte aligned (quad flush, TMA staging) at <= 4-way conflicts for such lines.
constexpr int kPitch = 260;
constexpr unsigned kPitchBytes = kPitch * 4u;
constexpr unsigned kIdxBits = 0x4B400000u;   // bit pattern of 1.5 * 2^23: floor(x) + offset lives in the low mantissa bits

// ---- packed fp32x2 values as ONE 64-bit register pair -------------------------------------------------------------------
// (float2 structs are split into two independent 32-bit registers by the compiler and re-paired with MOVs in front of every
// packed instruction when their halves come from different places - measured: 10 MOV per ray; a 64-bit value stays paired)
typedef unsigned long long P2;
__device__ __forceinline__ P2 pk(float x, float y) { P2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ P2 pb(float x) { return pk(x, x); }
__device__ __forceinline__ void unpk(P2 v, float& x, float& y) { asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ void unpk_bits(P2 v, unsigned& x, unsigned& y) { asm("mov.b64 {%0,%1}, %2;" : "=r"(x), "=r"(y) : "l"(v)); }
__device__ __forceinline__ P2 fma2(P2 a, P2 b, P2 c) { P2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ P2 add_rm2(P2 a, P2 b) { P2 d; asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// strict packed product / sum / difference: identity-operand FMAs (see common.cuh, Packed)
struct PK {
    float one, nz, m1;
    __device__ __forceinline__ explicit PK(const PackedIdentities& k) : one(k.one), nz(k.negzero), m1(k.negone) {}
    __device__ __forceinline__ P2 mul(P2 a, P2 b) const { return fma2(a, b, pb(nz)); }    // RN(a*b)
    __device__ __forceinline__ P2 add(P2 a, P2 b) const { return fma2(a, pb(one), b); }   // RN(a+b)
    __device__ __forceinline__ P2 sub(P2 a, P2 b) const { return fma2(b, pb(m1), a); }    // RN(a-b)
};

// packed apply_orientation for the two points of a pair: the FMA chain over k of the scalar form, per element
__device__ __forceinline__ void orient2(const float* O, const float4 a, const float4 b, P2& x, P2& y, P2& z, const PK& K) {
    const P2 dx = pk(a.x, b.x), dy = pk(a.y, b.y), dz = pk(a.z, b.z), dw = pk(a.w, b.w);
    x = fma2(dw, pb(O[3]), fma2(dz, pb(O[2]), fma2(dy, pb(O[1]), K.mul(dx, pb(O[0])))));
    y = fma2(dw, pb(O[7]), fma2(dz, pb(O[6]), fma2(dy, pb(O[5]), K.mul(dx, pb(O[4])))));
    z = fma2(dw, pb(O[11]), fma2(dz, pb(O[10]), fma2(dy, pb(O[9]), K.mul(dx, pb(O[8])))));
}

// Per point PAIR: origins, preferred reflection directions and plane numerators of the two points (one per lane),
// computed with exactly the operations of make_point / make_origin (planar).
struct PairCtx {
    P2 o0, o2, r0, r1, r2, nr1, nr2, num;
};

// returns false if one of the two points is irregular (|numerator| outside [1e-18, 1e18]); that lane's numerator is then
// NaN, so every ray of it fails the validity tests of the row code
__device__ __forceinline__ bool make_pair(PairCtx& pc, const TargetCtx& T, const PointSrc& src, float i0, float i1, float i2,
                                          const float4 oa, const float4 na, const float4 ob, const float4 nb, const PK& K) {
    P2 ox, oy, oz, nx, ny, nz;
    if (src.O) {
        orient2(src.O, oa, ob, ox, oy, oz, K);
        orient2(src.O, na, nb, nx, ny, nz, K);
    } else {
        ox = pk(oa.x, ob.x); oy = pk(oa.y, ob.y); oz = pk(oa.z, ob.z);
        nx = pk(na.x, nb.x); ny = pk(na.y, nb.y); nz = pk(na.z, nb.z);
    }
    const P2 dot = K.add(K.add(K.mul(pb(i0), nx), K.mul(pb(i1), ny)), K.mul(pb(i2), nz));
    const P2 two_dot = K.mul(pb(2.0f), dot);
    pc.r0 = K.sub(pb(i0), K.mul(two_dot, nx));
    pc.r1 = K.sub(pb(i1), K.mul(two_dot, ny));
    pc.r2 = K.sub(pb(i2), K.mul(two_dot, nz));
    pc.nr1 = pc.r1 ^ 0x8000000080000000ull;
    pc.nr2 = pc.r2 ^ 0x8000000080000000ull;
    pc.o0 = ox; pc.o2 = oz;
    pc.num = K.add(K.add(K.mul(K.sub(pb(T.c0), ox), pb(T.n0)), K.mul(K.sub(pb(T.c1), oy), pb(T.n1))),
                   K.mul(K.sub(pb(T.c2), oz), pb(T.n2)));
    float na_, nb_;
    unpk(pc.num, na_, nb_);
    const bool ok_a = fabsf(na_) > 1e-18f && fabsf(na_) < 1e18f, ok_b = fabsf(nb_) > 1e-18f && fabsf(nb_) < 1e18f;
    if (!(ok_a && ok_b)) {   // never with physical inputs
        const float nan = __int_as_float(0x7fc00000);
        pc.num = pk(ok_a ? na_ : nan, ok_b ? nb_ : nan);
        return false;
    }
    return true;
}

// sin / cos of the angles of one ray row.  xu = (u of point a, u of point b), xe likewise.
struct RowTrig {
    P2 cu, su, ce, se;
};
__device__ __forceinline__ void sincos_tiny_p(P2 x, P2& s, P2& c, const PK& K) {   // = sincos_tiny2
    const P2 z = K.mul(x, x);
    s = fma2(K.mul(x, z), pb(-1.6666667163e-1f), x);
#ifndef AB200_COS_CORRECTLY_ROUNDED
    const P2 ax = x & 0x7fffffff7fffffffull;   // |x| in both lanes
    c = fma2(pb(-1.0f), fma2(pb(kCosBiasB), ax, fma2(pb(0.5f), z, pb(-kCosBiasA))), pb(1.0f));
#else
    c = fma2(pb(-0.5f), z, pb(1.0f));
#endif
}
__device__ __forceinline__ void sincos_poly_p(P2 x, P2& s, P2& c, const PK& K) {   // = sincos_poly_core2
    const P2 z = K.mul(x, x);
    P2 ps = fma2(pb(-1.9515295891e-4f), z, pb(8.3321608736e-3f));
    ps = fma2(ps, z, pb(-1.6666654611e-1f));
    s = fma2(K.mul(ps, z), x, x);
    P2 pc = fma2(pb(2.443315711809948e-5f), z, pb(-1.388731625493765e-3f));
    pc = fma2(pc, z, pb(4.166664568298827e-2f));
    c = fma2(K.mul(pc, z), z, fma2(pb(-0.5f), z, pb(1.0f)));
}
__device__ __forceinline__ P2 sel2(bool ta, bool tb, P2 t, P2 p) {
    float tx, ty, px, py;
    unpk(t, tx, ty); unpk(p, px, py);
    return pk(ta ? tx : px, tb ? ty : py);
}
// the rare mixed case: per lane the small-angle form or the polynomial; a lane whose angle lies beyond the polynomial's
// range is poisoned (NaN cosine): it is invalid here and the generic loop re-traces it.  Out of line, and the result comes
// back in REGISTERS (one packed value per call): a by-reference result would live on the stack, and ptxas puts stack loads
// on the same scoreboard as the distortion prefetch - the first use after the join then waits for the prefetch.
__device__ __noinline__ P2 trig_mixed(const P2 x, const int tiny_mask, const int ok_mask, const int want_sin,
                                      const PackedIdentities ident) {
    const PK K(ident);
    P2 ts, tc, ps, pc;
    sincos_tiny_p(x, ts, tc, K);
    sincos_poly_p(x, ps, pc, K);
    const bool tiny_a = tiny_mask & 1, tiny_b = tiny_mask & 2;
    if (want_sin) return sel2(tiny_a, tiny_b, ts, ps);
    const P2 c = sel2(tiny_a, tiny_b, tc, pc);
    float cx, cy;
    unpk(c, cx, cy);
    const float nan = __int_as_float(0x7fc00000);
    return pk((ok_mask & 1) ? cx : nan, (ok_mask & 2) ? cy : nan);
}

struct FwdV3 {
    unsigned win_c;      // window byte address with the floor-trick offsets folded in (see tap_address)
    unsigned win_s;      // shared-memory byte address of the window's first cell
    unsigned col_lim;    // E - 1: (me bits - kIdxBits) below this <=> 0 <= ie and ie + 1 < E
    unsigned whm1;       // window rows - 1
    int u0, wh, E, U;
    unsigned* out_u;     // this sample's [U,E] output row (integer taps outside the window)
};

__device__ __forceinline__ unsigned tap_address(const FwdV3& f, unsigned meb, unsigned mub) {
    // base + ((iu - u0) * pitch + ie) * 4 with ie = meb - kIdxBits, iu - u0 = mub - kIdxBits (mod 2^32)
    return mub * kPitchBytes + ((meb << 2) + f.win_c);
}

// the four taps of a ray into the shared-memory window (ptxas turns predicated reductions into one branch region per
// instruction, so this stays an ordinary if-block: one region per ray)
__device__ __forceinline__ void window_taps(const bool on, const unsigned lo, unsigned q1, unsigned q2, unsigned q3, unsigned q4) {
    if (on) {
        asm volatile("red.shared.add.u32 [%0+1040], %1;" ::"r"(lo), "r"(q1) : "memory");
        asm volatile("red.shared.add.u32 [%0+1044], %1;" ::"r"(lo), "r"(q2) : "memory");
        asm volatile("red.shared.add.u32 [%0+4], %1;" ::"r"(lo), "r"(q3) : "memory");
        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(lo), "r"(q4) : "memory");
    }
}

// taps of one valid ray that is NOT in the window interior (rare): per-tap routing, window rows in shared memory, other
// rows as integer REDs on the (pre-zeroed) output row; rays off the bitmap's last row / column are dropped
__device__ __noinline__ bool slow_taps(const unsigned win_s, const int u0, const int wh, const int E, const int U, unsigned* out_u,
                                       unsigned meb, unsigned mub, unsigned q1, unsigned q2, unsigned q3, unsigned q4) {
    const int ie = (int)(meb - kIdxBits), cu = (int)(mub - kIdxBits), iu = cu + u0;
    bool fell_back = false;
    if (ie + 1 >= E || iu + 1 >= U || ie < 0 || iu < 0) return false;
    const bool u_in0 = (unsigned)cu < (unsigned)wh, u_in1 = (unsigned)(cu + 1) < (unsigned)wh;
    unsigned* g_hi = out_u + (size_t)(U - 1 - (iu + 1)) * E + ie;
    unsigned* g_lo = g_hi + E;
    const unsigned s_lo = win_s + (unsigned)(cu * kPitch + ie) * 4u, s_hi = s_lo + kPitch * 4u;
    if (u_in1) {
        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(s_hi), "r"(q1) : "memory");
        asm volatile("red.shared.add.u32 [%0+4], %1;" ::"r"(s_hi), "r"(q2) : "memory");
    } else { fell_back = true; atomicAdd(g_hi, q1); atomicAdd(g_hi + 1, q2); }
    if (u_in0) {
        asm volatile("red.shared.add.u32 [%0+4], %1;" ::"r"(s_lo), "r"(q3) : "memory");
        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(s_lo), "r"(q4) : "memory");
    } else { fell_back = true; atomicAdd(g_lo + 1, q3); atomicAdd(g_lo, q4); }
    return fell_back;
}

// Constants of the ray rows (scalars: packed instructions take them as broadcast operands).
struct RowConst {
    float n0, n1, n2, half_w, half_h, c0, c2, rw, rh, nw, nh, em1, um1, nmag, ome, refl, magic_e, magic_u, fxs_sub;
    unsigned em1_bits, um1_bits;
};

__device__ __forceinline__ void make_row_const(RowConst& c, const TraceParams& prm, const TargetCtx& T, int u0) {
    c.n0 = T.n0; c.n1 = T.n1; c.n2 = T.n2;
    c.half_w = T.half_w; c.half_h = T.half_h; c.c0 = T.c0; c.c2 = T.c2;
    c.rw = T.rw; c.rh = T.rh; c.nw = -T.w; c.nh = -T.h; c.em1 = T.em1; c.um1 = T.um1;
    c.nmag = -prm.a.ray_magnitude; c.ome = prm.a.one_minus_extinction; c.refl = prm.a.reflectivity;
    c.magic_e = 12582912.0f; c.magic_u = 12582912.0f - (float)u0;
    c.fxs_sub = prm.fx_scale * 0x1p-100f;
    c.em1_bits = __float_as_uint(T.em1); c.um1_bits = __float_as_uint(T.um1);
}

__device__ __forceinline__ P2 div_regular_p(P2 a, P2 b, const PK& K) {   // = div_regular2
    float bx, by, rx, ry;
    unpk(b, bx, by);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rx) : "f"(bx));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ry) : "f"(by));
    P2 r = pk(rx, ry);
    const P2 nb = K.mul(b, pb(K.m1));                 // -b (exact)
    const P2 e = fma2(nb, r, pb(K.one));
    r = fma2(r, e, r);
    const P2 q = fma2(a, r, pb(0.0f));
    const P2 rem = fma2(nb, q, a);
    return fma2(r, rem, q);
}

// One ray row of a point pair: strict coordinate path (operation for operation that of fwd_rays_planar_fast2), splat
// weights, taps.  Lanes that must not produce a ray here (irregular point, angle beyond the polynomial's range) arrive
// POISONED (NaN numerator / NaN trig) and fail every test below.
// The common case is decided by THREE unsigned compares per ray on bit patterns: the ray is front-facing and regular
// (a in (-1e18, -1e-18)), its column index ie = floor(be) lies in [0, E-2] and its window row iu - u0 in [0, wh-2] -
// which implies that it is valid (0 <= be <= E-1, 0 <= bu <= U-1) and on the bitmap.  Such rays are counted and tapped
// into the window; everything else (invalid rays, the bitmap's last row / column, rows outside the window, back-facing
// or irregular rays) goes through one rarely needed region that applies the reference's exact tests.
// AXIS_N: the target normal is exactly +-north (n_e = n_u = 0, every scenario of the reference): a = RN(d_n * n_n).
struct RowOut {   // what the taps of a ray row need (packed: lane a = first point, lane b = second point)
    P2 a, me, mu, v1, v2, v3, v4, be0, bu0;
};

template <bool AXIS_N>
__device__ __forceinline__ void fwd_row_math(RowOut& o, const RowConst& c, const PairCtx& pc, const RowTrig& tr, const PK& K) {
    const P2 m10 = K.mul(tr.ce, tr.su), m11 = K.mul(tr.ce, tr.cu), m20 = K.mul(tr.se, tr.su), m21 = K.mul(tr.se, tr.cu);
    const P2 dx = K.add(K.mul(tr.cu, pc.r0), K.mul(tr.su, pc.nr1));
    const P2 dy = K.add(K.add(K.mul(m10, pc.r0), K.mul(m11, pc.r1)), K.mul(tr.se, pc.nr2));
    const P2 dz = K.add(K.add(K.mul(m20, pc.r0), K.mul(m21, pc.r1)), K.mul(tr.ce, pc.r2));
    const P2 a = AXIS_N ? K.mul(dy, pb(c.n1)) : K.add(K.add(K.mul(dx, pb(c.n0)), K.mul(dy, pb(c.n1))), K.mul(dz, pb(c.n2)));
    const P2 t = div_regular_p(pc.num, a, K);
    const P2 X = K.add(pc.o0, K.mul(dx, t));
    const P2 Z = K.add(pc.o2, K.mul(dz, t));
    const P2 te = K.sub(K.add(X, pb(c.half_w)), pb(c.c0));
    const P2 tu = K.sub(K.add(Z, pb(c.half_h)), pb(c.c2));
    const P2 qe0 = K.mul(te, pb(c.rw)), qu0 = K.mul(tu, pb(c.rh));
    const P2 be0 = fma2(fma2(fma2(qe0, pb(c.nw), te), pb(c.rw), qe0), pb(c.em1), pb(0.0f));
    const P2 bu0 = fma2(fma2(fma2(qu0, pb(c.nh), tu), pb(c.rh), qu0), pb(c.um1), pb(0.0f));
    const P2 inten = K.mul(K.mul(K.mul(a, pb(c.nmag)), pb(c.ome)), pb(c.refl));   // mag * (-a) * (1 - extinction) * reflectivity
    const P2 be = K.sub(pb(c.em1), be0), bu = bu0;
    const P2 me = add_rm2(be, pb(c.magic_e)), mu = add_rm2(bu, pb(c.magic_u));
    const P2 whe = K.sub(be, K.sub(me, pb(c.magic_e))), whu = K.sub(bu, K.sub(mu, pb(c.magic_u)));
    const P2 wlu = K.sub(pb(K.one), whu);
    const P2 sc = K.mul(inten, pb(c.fxs_sub));                 // >= 0 for valid rays (magnitude, 1 - extinction, reflectivity > 0)
    const P2 ahi = K.mul(whu, sc), alo = K.mul(wlu, sc);
    const P2 whes = K.mul(whe, pb(0x1p-49f)), wles = K.sub(pb(0x1p-49f), whes);
    o.v1 = fma2(wles, ahi, pb(0.0f)); o.v2 = fma2(whes, ahi, pb(0.0f));
    o.v3 = fma2(whes, alo, pb(0.0f)); o.v4 = fma2(wles, alo, pb(0.0f));
    o.a = a; o.me = me; o.mu = mu; o.be0 = be0; o.bu0 = bu0;
}

// Taps of one ray row.  The common case is decided by THREE unsigned compares per ray on bit patterns: the ray is
// front-facing and regular (a in (-1e18, -1e-18)), its column index ie = floor(be) lies in [0, E-2] and its window row
// iu - u0 in [0, wh-2] - which implies that it is valid (0 <= be <= E-1, 0 <= bu <= U-1) and on the bitmap.  When both
// rays of the row pass, eight reductions go to the window; otherwise (invalid rays, the bitmap's last row / column, rows
// outside the window, back-facing or irregular rays, poisoned lanes) the same region applies the reference's exact tests
// per ray.
__device__ __forceinline__ void fwd_row_taps(const FwdV3& f, const RowConst& c, const RowOut& o, int& cnt_valid, bool& susp,
                                             bool& fell_back) {
    float a_a, a_b;
    unpk(o.a, a_a, a_b);
    unsigned meb_a, meb_b, mub_a, mub_b, q1a, q1b, q2a, q2b, q3a, q3b, q4a, q4b;
    unpk_bits(o.me, meb_a, meb_b); unpk_bits(o.mu, mub_a, mub_b);
    unpk_bits(o.v1, q1a, q1b); unpk_bits(o.v2, q2a, q2b); unpk_bits(o.v3, q3a, q3b); unpk_bits(o.v4, q4a, q4b);
    const bool fr_a = front_regular(a_a), fr_b = front_regular(a_b);
    const bool fin_a = fr_a && ((meb_a - kIdxBits) < f.col_lim) && ((mub_a - kIdxBits) < f.whm1);
    const bool fin_b = fr_b && ((meb_b - kIdxBits) < f.col_lim) && ((mub_b - kIdxBits) < f.whm1);
    const unsigned lo_a = tap_address(f, meb_a, mub_a), lo_b = tap_address(f, meb_b, mub_b);
    if (fin_a && fin_b) {
        cnt_valid += 2;
        asm volatile("red.shared.add.u32 [%0+1040], %1;" ::"r"(lo_a), "r"(q1a) : "memory");
        asm volatile("red.shared.add.u32 [%0+1044], %1;" ::"r"(lo_a), "r"(q2a) : "memory");
        asm volatile("red.shared.add.u32 [%0+4], %1;" ::"r"(lo_a), "r"(q3a) : "memory");
        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(lo_a), "r"(q4a) : "memory");
        asm volatile("red.shared.add.u32 [%0+1040], %1;" ::"r"(lo_b), "r"(q1b) : "memory");
        asm volatile("red.shared.add.u32 [%0+1044], %1;" ::"r"(lo_b), "r"(q2b) : "memory");
        asm volatile("red.shared.add.u32 [%0+4], %1;" ::"r"(lo_b), "r"(q3b) : "memory");
        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(lo_b), "r"(q4b) : "memory");
    } else {
        unsigned be0_a, be0_b, bu0_a, bu0_b;
        unpk_bits(o.be0, be0_a, be0_b); unpk_bits(o.bu0, bu0_a, bu0_b);
        susp |= !(fr_a && fr_b);
        if (fin_a) {
            ++cnt_valid;
            window_taps(true, lo_a, q1a, q2a, q3a, q4a);
        } else if (fr_a && (be0_a <= c.em1_bits) && (bu0_a <= c.um1_bits)) {
            ++cnt_valid;
            fell_back |= slow_taps(f.win_s, f.u0, f.wh, f.E, f.U, f.out_u, meb_a, mub_a, q1a, q2a, q3a, q4a);
        }
        if (fin_b) {
            ++cnt_valid;
            window_taps(true, lo_b, q1b, q2b, q3b, q4b);
        } else if (fr_b && (be0_b <= c.em1_bits) && (bu0_b <= c.um1_bits)) {
            ++cnt_valid;
            fell_back |= slow_taps(f.win_s, f.u0, f.wh, f.E, f.U, f.out_u, meb_b, mub_b, q1b, q2b, q3b, q4b);
        }
    }
}

// the two ray rows (r, r + 1) of a point pair; xu = (u of point a, u of point b) of a row, xe likewise
struct GroupAngles {
    P2 xu0, xe0, xu1, xe1;
};
__device__ __forceinline__ void group_magnitudes(const GroupAngles& g, float& ma0, float& ma1, float& mb0, float& mb1) {
    float ua, ub, ea, eb;
    unpk(g.xu0, ua, ub); unpk(g.xe0, ea, eb);
    ma0 = fmaxf(fabsf(ua), fabsf(ea)); mb0 = fmaxf(fabsf(ub), fabsf(eb));
    unpk(g.xu1, ua, ub); unpk(g.xe1, ea, eb);
    ma1 = fmaxf(fabsf(ua), fabsf(ea)); mb1 = fmaxf(fabsf(ub), fabsf(eb));
}
__device__ __forceinline__ void group_trig(RowTrig& t0, RowTrig& t1, const GroupAngles& g, const float ma0, const float ma1,
                                           const float mb0, const float mb1, const PK& K, const PackedIdentities& ident) {
    const float ma = fmaxf(ma0, ma1), mb = fmaxf(mb0, mb1);
    if (fmaxf(ma, mb) <= kTinyAngle) {   // all but ~1e-5 of the groups of a physical sun shape
        sincos_tiny_p(g.xu0, t0.su, t0.cu, K);
        sincos_tiny_p(g.xe0, t0.se, t0.ce, K);
        sincos_tiny_p(g.xu1, t1.su, t1.cu, K);
        sincos_tiny_p(g.xe1, t1.se, t1.ce, K);
    } else {   // the pair rule of the fast2 loops, per point (= per lane)
        const bool tiny_a = ma <= kTinyAngle, tiny_b = mb <= kTinyAngle;
        const int tm = (int)tiny_a | ((int)tiny_b << 1);
        const int ok0 = (int)(tiny_a || ma0 <= 0.785f) | ((int)(tiny_b || mb0 <= 0.785f) << 1);
        const int ok1 = (int)(tiny_a || ma1 <= 0.785f) | ((int)(tiny_b || mb1 <= 0.785f) << 1);
        t0.cu = trig_mixed(g.xu0, tm, ok0, 0, ident); t0.su = trig_mixed(g.xu0, tm, 3, 1, ident);
        t0.ce = trig_mixed(g.xe0, tm, 3, 0, ident);   t0.se = trig_mixed(g.xe0, tm, 3, 1, ident);
        t1.cu = trig_mixed(g.xu1, tm, ok1, 0, ident); t1.su = trig_mixed(g.xu1, tm, 3, 1, ident);
        t1.ce = trig_mixed(g.xe1, tm, 3, 0, ident);   t1.se = trig_mixed(g.xe1, tm, 3, 1, ident);
    }
}
template <bool AXIS_N>
__device__ __forceinline__ void fwd_group(const FwdV3& f, const RowConst& c, const PairCtx& pc, const GroupAngles& g,
                                          const float ma0, const float ma1, const float mb0, const float mb1, const PK& K,
                                          const PackedIdentities& ident, int& cnt_valid, bool& susp, bool& fell_back) {
    RowTrig t0, t1;
    group_trig(t0, t1, g, ma0, ma1, mb0, mb1, K, ident);
    // the math of both rows first (one basic block: their two dependency chains interleave), then the taps
    RowOut o0, o1;
    fwd_row_math<AXIS_N>(o0, c, pc, t0, K);
    fwd_row_math<AXIS_N>(o1, c, pc, t1, K);
    fwd_row_taps(f, c, o0, cnt_valid, susp, fell_back);
    fwd_row_taps(f, c, o1, cnt_valid, susp, fell_back);
}

// The ray loop of one CTA: the thread's (point pair, ray pair) groups as ONE flat sequence, two groups per iteration with
// ping-pong distortion registers (group j is traced while j + 1 is in registers and j + 2 in flight - no register shifting,
// and a loop body of eight rays that stays in the instruction cache, unlike the fully unrolled per-point loop).
template <int THREADS, bool AXIS_N>
__device__ __forceinline__ void fwd_rays_v3(const TraceParams& prm, const TargetCtx& T, const FwdV3& f, const PointSrc& src, int h,
                                            float i0, float i1, float i2, int& cnt_valid_out, bool& fell_back_out, bool& any_irr_out) {
    const int tid = threadIdx.x;
    const int P = prm.a.n_points, R = prm.a.n_rays;
    const int NG = R >> 1;                       // ray pairs per point (R even, >= 4)
    const int P2n = P >> 1;                      // point pairs (P even)
    const PK K(prm.ident);
    RowConst c;
    make_row_const(c, prm, T, f.u0);
    asm volatile("" : "+f"(c.magic_u), "+f"(c.fxs_sub));   // loop invariants: kept in registers, not re-derived per row
    // planar distortions [2,N,R,P]: one 8-byte element per point pair and plane = the (point a, point b) register pair
    const P2* dist_u = reinterpret_cast<const P2*>(prm.a.distortions_planar) + (size_t)h * R * P2n;   // [R, P/2]
    const long long plane = (long long)prm.a.n_samples * R * P2n;                                      // u plane -> e plane
    int cnt_valid = 0;
    bool any_irr = false, fell_back = false;
    const long long row2 = 2 * (long long)P2n;   // element stride between two ray pairs
    // equal work per thread: with `rounds` = ceil(P2n / THREADS) passes the first `stride` = ceil(P2n / rounds) threads take
    // `rounds` point pairs each (5000 pairs, 512 threads: 500 threads x 10) instead of a last pass that keeps a quarter of
    // the warps waiting at the barrier
    const int rounds = (P2n + THREADS - 1) / THREADS;
    const int stride = (P2n + rounds - 1) / rounds;
    const int n_my = tid < stride ? (P2n - 1 - tid) / stride + 1 : 0;
    const int n_tot = n_my * NG;
    // Load stream: ONE group ahead of the compute stream, into the slot the previous group has just left (A / B alternate,
    // two groups per loop iteration: no register shifting).  Deeper prefetch would not help: ptxas puts every global load
    // of the loop on one scoreboard, so the first use of a group's data waits for everything in flight - the new loads are
    // therefore issued only AFTER that wait, at the start of each group, and get a whole group (four rays of four warps per
    // scheduler, ~2000 cycles) to arrive.
    const P2* lptr = dist_u + tid;
    const long long pp_step = (long long)stride - (long long)NG * row2;
    int lg = 0;
    auto load_next = [&](GroupAngles& g) {
        g.xu0 = __ldcs(lptr); g.xe0 = __ldcs(lptr + plane);
        g.xu1 = __ldcs(lptr + P2n); g.xe1 = __ldcs(lptr + plane + P2n);
        lptr += row2;
        if (++lg == NG) { lg = 0; lptr += pp_step; }
    };
    GroupAngles sA, sB;
    sA.xu0 = sA.xe0 = sA.xu1 = sA.xe1 = 0ull;
    sB = sA;
    int pp = tid, cg = 0;
    float4 oa, na, ob, nb;
    oa = na = ob = nb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n_tot > 0) {
        load_next(sA);
        oa = __ldg(src.pts + 2 * pp); ob = __ldg(src.pts + 2 * pp + 1);
        na = __ldg(src.nrm + 2 * pp); nb = __ldg(src.nrm + 2 * pp + 1);
    }
    PairCtx pc;
    // one group: `cur` holds its distortions, `nxt` receives those of the next group
    auto half = [&](const GroupAngles& cur, GroupAngles& nxt, const int j) {
        // (1) everything that waits for loaded data: the group's angle magnitudes and, at a point pair's first ray pair, the
        //     pair's rows (fetched during the previous pair's last group)
        float ma0, ma1, mb0, mb1;
        group_magnitudes(cur, ma0, ma1, mb0, mb1);
        if (cg == 0) {
            if (!make_pair(pc, T, src, i0, i1, i2, oa, na, ob, nb, K)) any_irr = true;
            const int ppn = pp + stride;
            if (ppn < P2n) { prefetch_l2(src.pts + 2 * ppn); prefetch_l2(src.nrm + 2 * ppn); }
            else if (tid * kWindowSampleStride < P && src.next && src.next[0]) {
                prefetch_l2(src.next[0] + tid * kWindowSampleStride); prefetch_l2(src.next[1] + tid * kWindowSampleStride);
            }
        }
        // (2) the next loads
        if (j + 1 < n_tot) load_next(nxt);
        if (cg == NG - 1 && pp + stride < P2n) {   // last ray pair: fetch the next point pair's rows (L2 hits)
            const int ppn = pp + stride;
            oa = __ldg(src.pts + 2 * ppn); ob = __ldg(src.pts + 2 * ppn + 1);
            na = __ldg(src.nrm + 2 * ppn); nb = __ldg(src.nrm + 2 * ppn + 1);
        }
        // (3) the math
        fwd_group<AXIS_N>(f, c, pc, cur, ma0, ma1, mb0, mb1, K, prm.ident, cnt_valid, any_irr, fell_back);
        if (++cg == NG) { cg = 0; pp += stride; }
    };
    for (int j = 0; j < n_tot; j += 2) {
        half(sA, sB, j);
        if (j + 1 < n_tot) half(sB, sA, j + 1);
    }
    cnt_valid_out = cnt_valid;
    fell_back_out = fell_back;
    any_irr_out = any_irr;
}

// Row range of the bitmap that the window holds: centred on the undistorted reflections of every 32nd surface point.
template <int THREADS>
__device__ int place_rows(const TraceParams& prm, const TargetCtx& T, const PointSrc& src, const WindowSamples& ws, int P, int wh,
                          float i0, float i1, float i2, float* red /* [64] */, int* u0_sh) {
    const int tid = threadIdx.x, U = prm.a.res_u;
    const float inf = __int_as_float(0x7f800000);
    float umin = inf, umax = -inf;
    if (U > wh) {
        const int pa = tid * kWindowSampleStride, pb = pa + THREADS * kWindowSampleStride;
        float4 oa = ws.oa, na = ws.na, ob = ws.ob, nb = ws.nb;
        orient_point(src, oa, na);
        orient_point(src, ob, nb);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (!((k == 0 ? pa : pb) < P)) continue;
            PointCtx pc;
            make_point(pc, T, i0, i1, i2, k == 0 ? oa : ob, k == 0 ? na : nb);
            float be, bu, t, cosi;
            const bool ok = T.planar ? centre_planar(T, pc, be, bu, t, cosi) : centre_cylinder(T, pc, be, bu, t, cosi);
            if (ok && bu == bu && fabsf(bu) < 1e6f) { umin = fminf(umin, bu); umax = fmaxf(umax, bu); }
        }
        for (int p = pb + THREADS * kWindowSampleStride; p < P; p += THREADS * kWindowSampleStride) {   // very large P only
            PointCtx pc;
            float4 o4 = __ldg(src.pts + p), n4 = __ldg(src.nrm + p);
            orient_point(src, o4, n4);
            make_point(pc, T, i0, i1, i2, o4, n4);
            float be, bu, t, cosi;
            const bool ok = T.planar ? centre_planar(T, pc, be, bu, t, cosi) : centre_cylinder(T, pc, be, bu, t, cosi);
            if (ok && bu == bu && fabsf(bu) < 1e6f) { umin = fminf(umin, bu); umax = fmaxf(umax, bu); }
        }
    }
    umin = warp_min(umin); umax = warp_max(umax);
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0) { red[warp] = umin; red[32 + warp] = umax; }
    __syncthreads();
    if (warp == 0) {
        const int nw = THREADS / 32;
        umin = lane < nw ? red[lane] : inf;
        umax = lane < nw ? red[32 + lane] : -inf;
        umin = warp_min(umin); umax = warp_max(umax);
        if (lane == 0) {
            int u0 = 0;
            if (U > wh && umin <= umax) {
                const int centre = (int)floorf(0.5f * (umin + umax));
                u0 = min(max(centre - wh / 2, 0), U - wh);
            }
            *u0_sh = u0;
        }
    }
    __syncthreads();
    return *u0_sh;
}

#ifndef AB200_V3_FWD_THREADS
#define AB200_V3_FWD_THREADS 512
#endif
constexpr int kFwdThreads = AB200_V3_FWD_THREADS;

}  // namespace v3

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
trace_fwd_v3_kernel(const TraceParams prm) {
    using namespace v3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned* win_u = reinterpret_cast<unsigned*>(smem_raw);
    __shared__ TargetCtx T_sh;
    __shared__ float red[64];
    __shared__ int cnt_sh[3];
    __shared__ int fallback_sh, u0_sh;
    __shared__ int fb_box[4];
    __shared__ float O_sh[16];
    __shared__ const float4* next_sh[2];
    __shared__ BlockPrim blk_dummy[1];

    const int tid = threadIdx.x;
    const int li = blockIdx.x;
    const int h = prm.a.local_rows ? min(max(prm.a.local_rows[li], 0), prm.a.n_samples - 1) : li;
    const int P = prm.a.n_points, R = prm.a.n_rays, E = prm.a.res_e, U = prm.a.res_u;
    const int wh = min(U, prm.win_cap / kPitch);

    long long t_phase = (prm.a.stats && tid == 0) ? clock64() : 0;
    const int hs = h;   // (no activation map on this path: v3_eligible)
    const float4* pts_h = reinterpret_cast<const float4*>(prm.a.points) + (size_t)hs * P;
    const float4* nrm_h = reinterpret_cast<const float4*>(prm.a.normals) + (size_t)hs * P;
    WindowSamples ws;
    load_window_samples<THREADS>(ws, pts_h, nrm_h, 0, U > wh ? P : 0);
    const float i0 = __ldg(prm.a.incident + 4 * h), i1 = __ldg(prm.a.incident + 4 * h + 1), i2 = __ldg(prm.a.incident + 4 * h + 2);
    if (tid == 0) {
        load_target(T_sh, prm.a.targets, prm.a.target_idx[h], E, U);
        cnt_sh[0] = 0; cnt_sh[1] = 0; cnt_sh[2] = 0; fallback_sh = 0;
        fb_box[0] = 1 << 30; fb_box[1] = -1; fb_box[2] = 1 << 30; fb_box[3] = -1;
    }
    if (prm.a.orientations && tid >= 32 && tid < 48) O_sh[tid - 32] = __ldg(prm.a.orientations + (size_t)h * 16 + (tid - 32));
    if (tid == 64) set_next_sample<false>(next_sh, prm, li);
    {
        uint4* w4 = reinterpret_cast<uint4*>(smem_raw);
        const int n4 = (wh * kPitch) >> 2;
        for (int i = tid; i < n4; i += THREADS) w4[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    AB200_PHASE(4, 0);   // start-up loads + window clear
    PointSrc src;
    src.pts = pts_h; src.nrm = nrm_h; src.O = prm.a.orientations ? O_sh : nullptr; src.next = next_sh;
    const TargetCtx T = T_sh;
    const int u0 = place_rows<THREADS>(prm, T, src, ws, P, wh, i0, i1, i2, red, &u0_sh);
    if (prm.a.windows && tid == 0) reinterpret_cast<int4*>(prm.a.windows)[h] = make_int4(0, u0, kPitch, wh);
    AB200_PHASE(4, 1);   // window placement
    float* out_f = prm.a.flux + (size_t)h * U * E;
    unsigned* out_u = reinterpret_cast<unsigned*>(out_f);
    const int e4 = E >> 2;
    const int r_lo = U - u0 - wh;                  // first output row of the window (output rows are flipped)
    const int n_out4 = (U - wh) * e4;              // quads of the output rows outside the window
    if (prm.self_zero) {
        float4* o4 = reinterpret_cast<float4*>(out_f);
        for (int i = tid; i < n_out4; i += THREADS) {
            int row = i / e4;
            const int q = i - row * e4;
            if (row >= r_lo) row += wh;
            o4[row * e4 + q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
    }
    AB200_PHASE(4, 2);   // clearing the bitmap rows outside the window

    FwdV3 f;
    f.win_s = (unsigned)__cvta_generic_to_shared(win_u);
    f.win_c = f.win_s - kIdxBits * kPitchBytes - (kIdxBits << 2);
    f.col_lim = (unsigned)(E - 1);
    f.whm1 = (unsigned)(wh - 1);
    f.u0 = u0; f.wh = wh; f.E = E; f.U = U; f.out_u = out_u;
    asm volatile("" : "+r"(f.win_c), "+r"(f.col_lim), "+r"(f.whm1));   // loop invariants: kept in registers

    FwdCtx fc;   // the same window for the generic loops (cylindrical targets, irregular rays)
    fc.win_u = win_u; fc.win_f = reinterpret_cast<float*>(win_u); fc.out_f = out_f;
    fc.e0 = 0; fc.u0 = u0; fc.ww = kPitch; fc.wh = wh; fc.wwm1 = E - 1; fc.whm1 = wh - 1;
    fc.fb_box = fb_box; fc.blk = blk_dummy; fc.n_blk = 0;

    int cnt_lam = 0, cnt_int = 0, cnt_blk = 0;
    bool fell_back = false;
    if (T.planar && T.fastdiv) {
        bool any_irr = false;
        int cnt_valid = 0;
        if (T.n0 == 0.0f && T.n2 == 0.0f) fwd_rays_v3<THREADS, true>(prm, T, f, src, h, i0, i1, i2, cnt_valid, fell_back, any_irr);
        else fwd_rays_v3<THREADS, false>(prm, T, f, src, h, i0, i1, i2, cnt_valid, fell_back, any_irr);
        cnt_lam = cnt_int = cnt_valid;
        if (__syncthreads_or(any_irr)) {   // never with physical inputs: the generic loop picks up the skipped rays
            int c1 = 0, c2 = 0, c3 = 0;
            bool fb = false;
            fwd_rays<THREADS, AB200_TRIG_POLY, false, false, true, false, true, false>(prm, T, fc, src, h, 0, P, i0, i1, i2, c1, c2, c3, fb);
            cnt_lam += c1; cnt_int += c2; fell_back = fell_back || fb;
        }
    } else if (T.planar) {
        fwd_rays<THREADS, AB200_TRIG_POLY, false, false, true, false, false, false>(prm, T, fc, src, h, 0, P, i0, i1, i2, cnt_lam, cnt_int, cnt_blk, fell_back);
    } else {
        fwd_rays<THREADS, AB200_TRIG_POLY, false, false, false, false, false, false>(prm, T, fc, src, h, 0, P, i0, i1, i2, cnt_lam, cnt_int, cnt_blk, fell_back);
    }
    AB200_PHASE(4, 3);   // ray loop (thread 0)
    cnt_lam = warp_sum(cnt_lam);
    cnt_int = warp_sum(cnt_int);
    if ((tid & 31) == 0) {
        if (cnt_lam) atomicAdd(&cnt_sh[0], cnt_lam);
        if (cnt_int) atomicAdd(&cnt_sh[1], cnt_int);
    }
    if (fell_back) fallback_sh = 1;
    if (prm.a.stats && tid == 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(prm.a.stats) + 1, (unsigned long long)(wh * E));
        atomicAdd(reinterpret_cast<unsigned long long*>(prm.a.stats) + 2, 1ull);
    }
    if (prm.a.stats && fell_back) atomicAdd(reinterpret_cast<unsigned long long*>(prm.a.stats) + 0, 1ull);
    __threadfence();   // integer REDs of this thread are performed before the barrier
    __syncthreads();
    AB200_PHASE(4, 4);   // waiting for the CTA's last warp
    if (tid == 0) {
        const float rp = (float)(R * P);
        prm.a.on_target[h] = sdiv((float)cnt_sh[0], rp);
        prm.a.intercept[h] = sdiv((float)cnt_sh[1], rp);
        prm.a.blocking[h] = 1.0f;
    }
    const float inv = prm.fx_inv;
    {   // window rows -> output rows, one quad (uint4 -> float4) per thread and iteration
        const int n4 = wh * e4;
        const int d_r = THREADS / e4, d_q = THREADS - d_r * e4;
        int r = tid / e4, q = tid - r * e4;
        for (int idx = tid; idx < n4; idx += THREADS) {
            const uint4 v = *reinterpret_cast<const uint4*>(win_u + r * kPitch + (q << 2));
            float4* orow = reinterpret_cast<float4*>(out_f + (size_t)(U - 1 - (u0 + r)) * E);
            __stcs(orow + q, make_float4(__uint2float_rn(v.x) * inv, __uint2float_rn(v.y) * inv, __uint2float_rn(v.z) * inv,
                                         __uint2float_rn(v.w) * inv));
            r += d_r; q += d_q;
            if (q >= e4) { q -= e4; ++r; }
        }
    }
    AB200_PHASE(4, 5);   // window flush (thread 0)
    if (fallback_sh != 0) {   // integer taps on the rows outside the window: convert in place
        uint4* o4 = reinterpret_cast<uint4*>(out_f);
        for (int i = tid; i < n_out4; i += THREADS) {
            int row = i / e4;
            const int q = i - row * e4;
            if (row >= r_lo) row += wh;
            const uint4 v = __ldcg(o4 + row * e4 + q);
            if (v.x | v.y | v.z | v.w)
                reinterpret_cast<float4*>(o4)[row * e4 + q] = make_float4(__uint2float_rn(v.x) * inv, __uint2float_rn(v.y) * inv,
                                                                         __uint2float_rn(v.z) * inv, __uint2float_rn(v.w) * inv);
        }
    }
    AB200_PHASE(4, 6);   // conversion of the out-of-window taps (thread 0)
}
