// Rigid-body kinematics (actuators -> joint angles -> 4x4 orientation), the iterative
// alignment to incident ray directions, and the application of the orientation to the surface
// points/normals, each with an explicit backward.
// Reference: artist/field/kinematics_rigid_body.py:194-324 (forward kinematics), :326-508 (closed
// form inverse), :540-634 (fixed-point alignment); artist/field/actuators_linear.py:79-370;
// artist/field/heliostat_group_rigid_body.py:217-222 (apply); artist/geometry/transforms.py:86-273.
#include <cooperative_groups.h>
#include <cstdlib>
#include "common.cuh"

namespace ab200 {

// Transcendental functions of the kinematics.  torch's CPU implementations (SLEEF, <= 1 ulp) return the correctly rounded
// float for 91 % (acos) ... 99.9 % (log1p) of their arguments; CUDA's float versions (<= 2 ulp) hit it less often.  These
// are per-heliostat evaluations (a few dozen per heliostat and step, nothing per ray), so they are taken in double
// precision and rounded once: the correctly rounded value in all but ~1e-8 of the cases - the closest a device
// function gets to the reference's last bit (tools/diag_kinematics_parity.py; -DAB200_KIN_FLOAT_TRIG switches back).
#ifdef AB200_KIN_FLOAT_TRIG
__device__ __forceinline__ void k_sincos(float x, float* s, float* c) { sincosf(x, s, c); }
__device__ __forceinline__ float k_sin(float x) { return sinf(x); }
__device__ __forceinline__ float k_cos(float x) { return cosf(x); }
__device__ __forceinline__ float k_acos(float x) { return acosf(x); }
__device__ __forceinline__ float k_asin(float x) { return asinf(x); }
__device__ __forceinline__ float k_atan2(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ float k_exp(float x) { return expf(x); }
__device__ __forceinline__ float k_log1p(float x) { return log1pf(x); }
#else
__device__ __forceinline__ void k_sincos(float x, float* s, float* c) {
    double sd, cd;
    sincos((double)x, &sd, &cd);
    *s = (float)sd; *c = (float)cd;
}
__device__ __forceinline__ float k_sin(float x) { return (float)sin((double)x); }
__device__ __forceinline__ float k_cos(float x) { return (float)cos((double)x); }
__device__ __forceinline__ float k_acos(float x) { return (float)acos((double)x); }
__device__ __forceinline__ float k_asin(float x) { return (float)asin((double)x); }
__device__ __forceinline__ float k_atan2(float y, float x) { return (float)atan2((double)y, (double)x); }
__device__ __forceinline__ float k_exp(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float k_log1p(float x) { return (float)log1p((double)x); }
#endif

struct M4 {
    float m[16];
};

__device__ inline M4 m4_identity() {
    M4 r;
#pragma unroll
    for (int i = 0; i < 16; ++i) r.m[i] = (i % 5 == 0) ? 1.f : 0.f;
    return r;
}
// Product of two per-heliostat 4x4 matrices as torch's CPU bmm evaluates it for these tiny batched operands (also when
// one side is a broadcast [1,4,4]): every product rounded, added left to right over k - NOT the FMA chain of the large
// GEMMs (checked against torch 2.11 CPU, numpy emulation == torch bit for bit on [N,4,4] @ [N,4,4], @ [1,4,4] and
// [N,4,4]^T @ [N,4,1]; the FMA chain differs in 38 % of the entries).
__device__ inline M4 m4_mul(const M4& a, const M4& b) {
    M4 r;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float acc = smul(a.m[i * 4], b.m[j]);
#pragma unroll
            for (int k = 1; k < 4; ++k) acc = sadd(acc, smul(a.m[i * 4 + k], b.m[k * 4 + j]));
            r.m[i * 4 + j] = acc;
        }
    return r;
}
__device__ inline M4 m4_transpose(const M4& a) {
    M4 r;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) r.m[i * 4 + j] = a.m[j * 4 + i];
    return r;
}
// axis: 0 = east, 1 = north, 2 = up  (transforms.py:86-212)
__device__ inline M4 m4_rot(int axis, float ang) {
    M4 r = m4_identity();
    float s, c;
    k_sincos(ang, &s, &c);
    if (axis == 0) { r.m[5] = c; r.m[6] = -s; r.m[9] = s; r.m[10] = c; }
    else if (axis == 1) { r.m[0] = c; r.m[2] = -s; r.m[8] = s; r.m[10] = c; }
    else { r.m[0] = c; r.m[1] = -s; r.m[4] = s; r.m[5] = c; }
    return r;
}
// d/d(angle) of the rotation matrix
__device__ inline M4 m4_drot(int axis, float ang) {
    M4 r;
#pragma unroll
    for (int i = 0; i < 16; ++i) r.m[i] = 0.f;
    float s, c;
    k_sincos(ang, &s, &c);
    if (axis == 0) { r.m[5] = -s; r.m[6] = -c; r.m[9] = c; r.m[10] = -s; }
    else if (axis == 1) { r.m[0] = -s; r.m[2] = -c; r.m[8] = c; r.m[10] = -s; }
    else { r.m[0] = -s; r.m[1] = -c; r.m[4] = c; r.m[5] = -s; }
    return r;
}
__device__ inline M4 m4_trans(float e, float n, float u) {
    M4 r = m4_identity();
    r.m[3] = e; r.m[7] = n; r.m[11] = u;
    return r;
}

// ---- actuators ----------------------------------------------------------------------------------
__device__ inline float softplus100(float x) {  // torch softplus(beta=100, threshold=20)
    const float bx = smul(x, 100.0f);
    return bx > 20.0f ? x : sdiv(k_log1p(k_exp(bx)), 100.0f);
}
__device__ inline float dsoftplus100(float x) {
    const float bx = x * 100.0f;
    return bx > 20.0f ? 1.0f : 1.0f / (1.0f + k_exp(-bx));
}

struct LinAct {
    float inc, off, rad, a0, s0, ds0, cw, lo, hi;
    float ab0, g0;   // absolute angle at motor position 0 and its stroke derivative: constants of the actuator
};
__device__ inline float lin_abs_angle(const LinAct& a, float motor, float* dabs_dstroke);

__device__ inline LinAct load_lin(const ab200_kinematics_args& k, int i, int j) {
    const float* no = k.actuator_non_opt + (size_t)i * 14;  // [7,2]
    const float* op = k.actuator_opt + (size_t)i * 4;       // [2,2]
    LinAct a;
    a.cw = no[1 * 2 + j]; a.lo = no[2 * 2 + j]; a.hi = no[3 * 2 + j];
    a.inc = sadd(softplus100(no[4 * 2 + j]), 1e-6f);
    a.off = sadd(softplus100(no[5 * 2 + j]), 1e-6f);
    a.rad = sadd(softplus100(no[6 * 2 + j]), 1e-6f);
    a.a0 = op[0 * 2 + j];
    a.s0 = sadd(softplus100(op[1 * 2 + j]), 1e-6f);
    a.ds0 = dsoftplus100(op[1 * 2 + j]);
    a.ab0 = lin_abs_angle(a, 0.0f, &a.g0);
    return a;
}

// absolute angle from the law of cosines; also d(angle)/d(stroke before clamp)
__device__ inline float lin_abs_angle(const LinAct& a, float motor, float* dabs_dstroke) {
    const float eps = 1e-6f;
    // actuators_linear.py:218-229 in its operation order (one rounding per product / sum / quotient)
    float stroke = sadd(sdiv(motor, a.inc), a.s0);
    const float lo = sadd(fabsf(ssub(a.off, a.rad)), eps), hi = ssub(sadd(a.off, a.rad), eps);
    const bool clamped_s = (stroke < lo) || (stroke > hi);
    stroke = fminf(fmaxf(stroke, lo), hi);
    const float div = sdiv(ssub(sadd(smul(a.off, a.off), smul(a.rad, a.rad)), smul(stroke, stroke)), smul(smul(2.0f, a.off), a.rad));
    const bool clamped_d = (div < -1.0f + 1e-6f) || (div > 1.0f - 1e-6f);
    const float dc = fminf(fmaxf(div, -1.0f + 1e-6f), 1.0f - 1e-6f);
    if (dabs_dstroke) {
        float g = 0.f;
        if (!clamped_s && !clamped_d) g = (-1.0f / sqrtf(1.0f - dc * dc)) * (-stroke / (a.off * a.rad));
        *dabs_dstroke = g;
    }
    return k_acos(dc);
}

__device__ inline float lin_motor_to_angle(const LinAct& a, float motor, float* dang_dmotor, float* dang_da0, float* dang_ds0raw) {
    float g;
    const float g0 = a.g0;
    const float ab = lin_abs_angle(a, motor, &g);
    const float delta = ssub(a.ab0, ab);
    const float sign = (a.cw == 1.0f) ? 1.0f : ((a.cw == 0.0f) ? -1.0f : 0.0f);
    if (dang_dmotor) *dang_dmotor = sign * (-g) / a.inc;
    if (dang_da0) *dang_da0 = 1.0f;
    if (dang_ds0raw) *dang_ds0raw = sign * (g0 - g) * a.ds0;
    return sadd(a.a0, sign * delta);   // sign * delta is exact
}

__device__ inline float lin_angle_to_motor(const LinAct& a, float angle) {
    const float eps = 1e-6f;
    // actuators_linear.py:331-370 in its operation order
    const float delta = (a.cw == 1.0f) ? ssub(angle, a.a0) : ssub(a.a0, angle);
    const float ia = ssub(a.ab0, delta);
    const float cv = fminf(fmaxf(k_cos(ia), -1.0f + 1e-6f), 1.0f - 1e-6f);
    float stroke = sqrtf(ssub(sadd(smul(a.off, a.off), smul(a.rad, a.rad)), smul(smul(smul(2.0f, a.off), a.rad), cv)));
    stroke = fminf(fmaxf(stroke, sadd(fabsf(ssub(a.off, a.rad)), eps)), ssub(sadd(a.off, a.rad), eps));
    return smul(ssub(stroke, a.s0), a.inc);
}

// ---- forward kinematics chain -----------------------------------------------------------------
constexpr int kFactors = 11;
// factor k: 0 T(pos) 1 Rn(t1n) 2 Ru(t1u) 3 T(t1) 4 Re(theta1) 5 Re(t2e) 6 Rn(t2n) 7 T(t2) 8 Ru(theta2) 9 T(tc) 10 offset

// `pre` (optional): the heliostat's two actuators already loaded (the alignment loop loads them once for all sweeps)
__device__ inline void joint_angles(const ab200_kinematics_args& k, int i, const float* motor, float* th, float* dth_dm,
                                    float* dth_da0, float* dth_ds0, const LinAct* pre = nullptr) {
    for (int j = 0; j < 2; ++j) {
        const float mp = motor[(size_t)i * 2 + j];
        if (k.linear_actuators) {
            const LinAct a = pre ? pre[j] : load_lin(k, i, j);
            th[j] = lin_motor_to_angle(a, mp, dth_dm ? dth_dm + j : nullptr, dth_da0 ? dth_da0 + j : nullptr,
                                       dth_ds0 ? dth_ds0 + j : nullptr);
        } else {
            th[j] = mp;
            if (dth_dm) dth_dm[j] = 1.0f;
            if (dth_da0) dth_da0[j] = 0.0f;
            if (dth_ds0) dth_ds0[j] = 0.0f;
        }
    }
}

__device__ inline void build_factors(M4* F, const ab200_kinematics_args& k, int i, const float* th) {
    const float* pos = k.positions + (size_t)i * 4;
    const float* td = k.translation_dev + (size_t)i * 9;
    const float* rd = k.rotation_dev + (size_t)i * 4;
    F[0] = m4_trans(pos[0], pos[1], pos[2]);
    F[1] = m4_rot(1, rd[0]);
    F[2] = m4_rot(2, rd[1]);
    F[3] = m4_trans(td[0], td[1], td[2]);
    F[4] = m4_rot(0, th[0]);
    F[5] = m4_rot(0, rd[2]);
    F[6] = m4_rot(1, rd[3]);
    F[7] = m4_trans(td[3], td[4], td[5]);
    F[8] = m4_rot(2, th[1]);
    F[9] = m4_trans(td[6], td[7], td[8]);
    for (int q = 0; q < 16; ++q) F[10].m[q] = k.orientation_offset[q];
}

// orientation WITHOUT the final offset, associated like the reference: ((T J1) J2) Tc with
// J1 = ((Rn Ru) T1) Re, J2 = ((Re Rn) T2) Ru
__device__ inline M4 raw_orientation(const M4* F) {
    const M4 j1 = m4_mul(m4_mul(m4_mul(F[1], F[2]), F[3]), F[4]);
    const M4 j2 = m4_mul(m4_mul(m4_mul(F[5], F[6]), F[7]), F[8]);
    return m4_mul(m4_mul(m4_mul(F[0], j1), j2), F[9]);
}

__global__ void kinematics_fwd_kernel(const ab200_kinematics_args k, const float* __restrict__ motor, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k.n) return;
    float th[2];
    joint_angles(k, i, motor, th, nullptr, nullptr, nullptr);
    M4 F[kFactors];
    build_factors(F, k, i, th);
    const M4 o = m4_mul(raw_orientation(F), F[10]);
    for (int q = 0; q < 16; ++q) out[(size_t)i * 16 + q] = o.m[q];
}

__device__ inline float m4_dot(const M4& a, const M4& b) {
    float s = 0.f;
    for (int q = 0; q < 16; ++q) s = fmaf(a.m[q], b.m[q], s);
    return s;
}

__global__ void kinematics_bwd_kernel(const ab200_kinematics_args k, const float* __restrict__ motor,
                                      const float* __restrict__ gout, float* g_motor, float* g_rot, float* g_trans,
                                      float* g_act, float* g_pos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k.n) return;
    float th[2], dm[2], da0[2], ds0[2];
    joint_angles(k, i, motor, th, dm, da0, ds0);
    M4 F[kFactors];
    build_factors(F, k, i, th);
    // prefix[k] = F0..F(k-1), suffix[k] = F(k+1)..F(K-1)
    M4 prefix[kFactors], suffix[kFactors];
    prefix[0] = m4_identity();
    for (int q = 1; q < kFactors; ++q) prefix[q] = m4_mul(prefix[q - 1], F[q - 1]);
    suffix[kFactors - 1] = m4_identity();
    for (int q = kFactors - 2; q >= 0; --q) suffix[q] = m4_mul(F[q + 1], suffix[q + 1]);
    M4 G;
    for (int q = 0; q < 16; ++q) G.m[q] = gout[(size_t)i * 16 + q];
    const float* rd = k.rotation_dev + (size_t)i * 4;
    auto factor_grad = [&](int q) { return m4_mul(m4_mul(m4_transpose(prefix[q]), G), m4_transpose(suffix[q])); };
    const float g_t1n = m4_dot(factor_grad(1), m4_drot(1, rd[0]));
    const float g_t1u = m4_dot(factor_grad(2), m4_drot(2, rd[1]));
    const float g_th1 = m4_dot(factor_grad(4), m4_drot(0, th[0]));
    const float g_t2e = m4_dot(factor_grad(5), m4_drot(0, rd[2]));
    const float g_t2n = m4_dot(factor_grad(6), m4_drot(1, rd[3]));
    const float g_th2 = m4_dot(factor_grad(8), m4_drot(2, th[1]));
    if (g_rot) {
        g_rot[(size_t)i * 4 + 0] = g_t1n; g_rot[(size_t)i * 4 + 1] = g_t1u;
        g_rot[(size_t)i * 4 + 2] = g_t2e; g_rot[(size_t)i * 4 + 3] = g_t2n;
    }
    if (g_trans) {
        const M4 a = factor_grad(3), b = factor_grad(7), c = factor_grad(9);
        float* o = g_trans + (size_t)i * 9;
        o[0] = a.m[3]; o[1] = a.m[7]; o[2] = a.m[11];
        o[3] = b.m[3]; o[4] = b.m[7]; o[5] = b.m[11];
        o[6] = c.m[3]; o[7] = c.m[7]; o[8] = c.m[11];
    }
    if (g_pos) {
        const M4 a = factor_grad(0);
        float* o = g_pos + (size_t)i * 4;
        o[0] = a.m[3]; o[1] = a.m[7]; o[2] = a.m[11]; o[3] = 0.f;
    }
    if (g_motor) {
        g_motor[(size_t)i * 2 + 0] = g_th1 * dm[0];
        g_motor[(size_t)i * 2 + 1] = g_th2 * dm[1];
    }
    if (g_act) {  // [2,2]: row 0 initial angle, row 1 initial stroke length (raw, before softplus)
        float* o = g_act + (size_t)i * 4;
        o[0] = g_th1 * da0[0]; o[1] = g_th2 * da0[1];
        o[2] = g_th1 * ds0[0]; o[3] = g_th2 * ds0[1];
    }
}

// ---- inverse kinematics + fixed-point alignment (2 x max_iterations small multi-block launches; the convergence vote
// of all heliostats stays on the device, no host sync) ------------------
__device__ inline void motor_from_normal(const ab200_kinematics_args& k, int i, const float* nrm /*3*/, float* motor_out,
                                         const LinAct* pre = nullptr) {
    const float eps = 1e-8f;
    const float* rd = k.rotation_dev + (size_t)i * 4;
    const M4 f1 = m4_mul(m4_rot(1, rd[0]), m4_rot(2, rd[1]));
    const M4 f2 = m4_mul(m4_rot(0, rd[2]), m4_rot(1, rd[3]));
    // n' = F1^T n
    float np[3];
    // (kinematics_rigid_body.py:420-470 in its operation order; the batched mat-vec adds rounded products left to right,
    // the 4th term is f1[3][j] * 0 = +-0 and changes nothing)
    for (int j = 0; j < 3; ++j) np[j] = sadd(sadd(smul(f1.m[j], nrm[0]), smul(f1.m[4 + j], nrm[1])), smul(f1.m[8 + j], nrm[2]));
    const float f00 = f2.m[0], f01 = f2.m[1];
    const float den = sqrtf(sadd(smul(f00, f00), smul(f01, f01)));
    const float phi = k_atan2(-f01, f00);
    const float ratio = fminf(fmaxf(sdiv(np[0], sadd(den, eps)), -1.0f + eps), 1.0f - eps);
    const float pi = 3.14159265358979323846f;
    const float asr = k_asin(ratio);
    float s[2] = {ssub(asr, phi), ssub(ssub(pi, asr), phi)};
    float mot[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    LinAct la[2];
    if (k.linear_actuators) { la[0] = pre ? pre[0] : load_lin(k, i, 0); la[1] = pre ? pre[1] : load_lin(k, i, 1); }
    const float* no = k.actuator_non_opt + (size_t)i * 14;
    bool ok1 = true;
    // the reference evaluates both solutions and keeps the first when it lies inside the motor limits: the second is only
    // computed when it is needed (same result)
    for (int c = 0; c < 2; ++c) {
        if (c == 1 && ok1) break;
        float ss, cs;
        k_sincos(s[c], &ss, &cs);
        s[c] = k_atan2(ss, cs);
        // v = F2 Ru(s) (0,-1,0,0)
        k_sincos(s[c], &ss, &cs);
        const float w0 = ss, w1 = -cs;  // Ru(s) * (0,-1,0): (-(-sin), -cos, 0) = (sin, -cos, 0)
        const float v1 = sadd(smul(f2.m[4], w0), smul(f2.m[5], w1)), v2 = sadd(smul(f2.m[8], w0), smul(f2.m[9], w1));
        float a = k_atan2(ssub(smul(v1, np[2]), smul(v2, np[1])), sadd(smul(v1, np[1]), smul(v2, np[2])));
        float sa, ca;
        k_sincos(a, &sa, &ca);
        a = k_atan2(sa, ca);
        if (k.linear_actuators) {
            mot[c][0] = lin_angle_to_motor(la[0], a);
            mot[c][1] = lin_angle_to_motor(la[1], s[c]);
        } else {
            mot[c][0] = a; mot[c][1] = s[c];
        }
        if (c == 0) ok1 = (mot[0][0] >= no[4] && mot[0][0] <= no[6]) && (mot[0][1] >= no[5] && mot[0][1] <= no[7]);
    }
    motor_out[0] = ok1 ? mot[0][0] : mot[1][0];
    motor_out[1] = ok1 ? mot[0][1] : mot[1][1];
}

__device__ inline void normalize3(float* v, float eps) {
    const float n = fmaxf(sqrtf(fmaf(v[2], v[2], fmaf(v[1], v[1], smul(v[0], v[0])))), eps);   // torch's norm: the FMA chain
    v[0] = sdiv(v[0], n); v[1] = sdiv(v[1], n); v[2] = sdiv(v[2], n);
}

// The alignment loop: per iteration two small kernels (forward kinematics + convergence vote, then inverse
// kinematics), one thread per heliostat.  The reference stops when ALL heliostats converged
// (kinematics_rigid_body.py:621-625); the vote lives in device memory (`flags`), so there is no host sync:
// later launches simply return once `done` is set.
//   scratch layout (floats): [0,n) last loss | [n,4n) desired normal | then 8 ints: open[0..5], done
struct AlignFlags {
    int open[6];
    int done;
    int pad;
};

// forward kinematics of heliostat i at its current motor positions: orientation out, loss and desired normal to scratch;
// returns whether the heliostat is still "open" (first iteration, or its loss moved by more than min_eps)
__device__ __forceinline__ bool align_forward_one(const ab200_kinematics_args& k, const float* __restrict__ incident,
                                                  const float* __restrict__ aim, int i, int it, float min_eps,
                                                  float* __restrict__ out, float* __restrict__ motor_io,
                                                  float* __restrict__ scratch, const LinAct* pre = nullptr) {
    if (it == 0) { motor_io[(size_t)i * 2] = 0.f; motor_io[(size_t)i * 2 + 1] = 0.f; }
    float th[2];
    joint_angles(k, i, motor_io, th, nullptr, nullptr, nullptr, pre);
    M4 F[kFactors];
    build_factors(F, k, i, th);
    const M4 o = raw_orientation(F);
    const M4 fin = m4_mul(o, F[10]);
    for (int q = 0; q < 16; ++q) out[(size_t)i * 16 + q] = fin.m[q];
    const float cn[4] = {-o.m[1], -o.m[5], -o.m[9], -o.m[13]};   // O (0,-1,0,0)
    float wr[3] = {ssub(aim[(size_t)i * 4], o.m[3]), ssub(aim[(size_t)i * 4 + 1], o.m[7]), ssub(aim[(size_t)i * 4 + 2], o.m[11])};
    normalize3(wr, 1e-8f);
    float wn[3] = {sadd(-incident[(size_t)i * 4], wr[0]), sadd(-incident[(size_t)i * 4 + 1], wr[1]), sadd(-incident[(size_t)i * 4 + 2], wr[2])};
    normalize3(wn, 1e-8f);
    const float loss = (fabsf(wn[0] - cn[0]) + fabsf(wn[1] - cn[1]) + fabsf(wn[2] - cn[2]) + fabsf(cn[3])) / 4.0f;
    const bool open = it == 0 || !(fabsf(scratch[i] - loss) <= min_eps);
    scratch[i] = loss;
    scratch[(size_t)k.n + 3 * i] = wn[0]; scratch[(size_t)k.n + 3 * i + 1] = wn[1]; scratch[(size_t)k.n + 3 * i + 2] = wn[2];
    return open;
}

__device__ __forceinline__ void align_inverse_one(const ab200_kinematics_args& k, int i, float* __restrict__ motor_io,
                                                  const float* __restrict__ scratch, const LinAct* pre = nullptr) {
    const float wn[3] = {scratch[(size_t)k.n + 3 * i], scratch[(size_t)k.n + 3 * i + 1], scratch[(size_t)k.n + 3 * i + 2]};
    float mo[2];
    motor_from_normal(k, i, wn, mo, pre);
    motor_io[(size_t)i * 2] = mo[0];
    motor_io[(size_t)i * 2 + 1] = mo[1];
}

__global__ void kin_align_forward_kernel(const ab200_kinematics_args k, const float* __restrict__ incident,
                                         const float* __restrict__ aim, int it, float min_eps, float* __restrict__ out,
                                         float* __restrict__ motor_io, float* __restrict__ scratch) {
    AlignFlags* fl = reinterpret_cast<AlignFlags*>(scratch + (size_t)4 * k.n);
    if (fl->done) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k.n) return;
    if (align_forward_one(k, incident, aim, i, it, min_eps, out, motor_io, scratch)) fl->open[it] = 1;   // benign race: everybody writes 1
}

__global__ void kin_align_inverse_kernel(const ab200_kinematics_args k, int it, float* __restrict__ motor_io,
                                         float* __restrict__ scratch) {
    AlignFlags* fl = reinterpret_cast<AlignFlags*>(scratch + (size_t)4 * k.n);
    if (fl->done) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (it > 0 && fl->open[it] == 0) {   // torch.all(eps <= min_eps): stop, keep this iteration's orientations
        __syncthreads();                 // every thread of the block has read the vote before it is overwritten
        if (i == 0) fl->done = 1;
        return;
    }
    if (i >= k.n) return;
    align_inverse_one(k, i, motor_io, scratch);
}

// The whole loop in ONE launch for fields of up to kAlignClusterMax heliostats: one thread-block CLUSTER of 8 CTAs x 512
// threads, one heliostat per thread through all sweeps (forward and inverse kinematics of a heliostat never need another
// thread's data); the reference's "stop when ALL heliostats converged" (kinematics_rigid_body.py:621-625) is a
// __syncthreads_or per CTA plus a vote over the cluster through distributed shared memory - two cluster barriers per sweep
// instead of two kernel launches.  Same per-heliostat code as the two kernels above (bit-identical results).  (One CTA alone
// was slower than the eight launches: the per-heliostat chain of asin / atan2 / divisions is latency-bound on one SM.)
constexpr int kAlignClusterCtas = 8, kAlignClusterThreads = 512, kAlignClusterMax = kAlignClusterCtas * kAlignClusterThreads;
__global__ void __cluster_dims__(kAlignClusterCtas, 1, 1) __launch_bounds__(kAlignClusterThreads)
kin_align_cluster_kernel(const ab200_kinematics_args k, const float* __restrict__ incident, const float* __restrict__ aim,
                         int max_iterations, float min_eps, float* __restrict__ out, float* __restrict__ motor_io,
                         float* __restrict__ scratch) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ int open_sh;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool mine = i < k.n;
    LinAct la[2];   // the heliostat's actuators: loaded (softplus, law of cosines at rest) once for all sweeps
    const LinAct* pre = nullptr;
    if (mine && k.linear_actuators) { la[0] = load_lin(k, i, 0); la[1] = load_lin(k, i, 1); pre = la; }
    for (int it = 0; it < max_iterations; ++it) {
        const bool open = mine && align_forward_one(k, incident, aim, i, it, min_eps, out, motor_io, scratch, pre);
        const int cta_open = __syncthreads_or(open);
        if (threadIdx.x == 0) open_sh = cta_open;
        cluster.sync();                                   // every CTA's vote is visible
        int any_open = 0;
        for (int r = 0; r < kAlignClusterCtas; ++r) any_open |= *cluster.map_shared_rank(&open_sh, r);
        cluster.sync();                                   // ... and read, before the next sweep overwrites it
        if (!any_open && it > 0) break;                   // uniform over the cluster
        if (mine) align_inverse_one(k, i, motor_io, scratch, pre);
    }
}

__global__ void __launch_bounds__(256) align_fwd_kernel(const float4* __restrict__ pts, const float4* __restrict__ nrm,
                                                        const float* __restrict__ ori, const int* __restrict__ src_row,
                                                        int n_points, float4* __restrict__ out_p, float4* __restrict__ out_n) {
    __shared__ float O[16];
    const int n = blockIdx.y;
    if (threadIdx.x < 16) O[threadIdx.x] = ori[(size_t)n * 16 + threadIdx.x];
    __syncthreads();
    const int src = src_row ? src_row[n] : n;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_points; p += gridDim.x * blockDim.x) {
        const float4 d = __ldg(pts + (size_t)src * n_points + p);
        const float4 e = __ldg(nrm + (size_t)src * n_points + p);
        float r[4], s[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // out_j = sum_k d_k * O[j][k]  (FMA chain over k, like the CPU GEMM)
            r[j] = fmaf(d.w, O[j * 4 + 3], fmaf(d.z, O[j * 4 + 2], fmaf(d.y, O[j * 4 + 1], smul(d.x, O[j * 4]))));
            s[j] = fmaf(e.w, O[j * 4 + 3], fmaf(e.z, O[j * 4 + 2], fmaf(e.y, O[j * 4 + 1], smul(e.x, O[j * 4]))));
        }
        out_p[(size_t)n * n_points + p] = make_float4(r[0], r[1], r[2], r[3]);
        out_n[(size_t)n * n_points + p] = make_float4(s[0], s[1], s[2], s[3]);
    }
}

// grad wrt the un-aligned data (g @ O) and wrt O (sum_p g_j d_k), ordered block reduction
__global__ void __launch_bounds__(256) align_bwd_kernel(const float4* __restrict__ pts, const float4* __restrict__ nrm,
                                                        const float* __restrict__ ori, const int* __restrict__ src_row,
                                                        int n_points, const float4* __restrict__ g_op,
                                                        const float4* __restrict__ g_on, float4* __restrict__ g_p,
                                                        float4* __restrict__ g_n, float* __restrict__ g_ori) {
    __shared__ float O[16];
    __shared__ float red[8][16];
    const int n = blockIdx.x;
    if (threadIdx.x < 16) O[threadIdx.x] = ori[(size_t)n * 16 + threadIdx.x];
    __syncthreads();
    const int src = src_row ? src_row[n] : n;
    float acc[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = 0.f;
    for (int p = threadIdx.x; p < n_points; p += blockDim.x) {
        const float4 gp = g_op ? g_op[(size_t)n * n_points + p] : make_float4(0, 0, 0, 0);
        const float4 gn = g_on ? g_on[(size_t)n * n_points + p] : make_float4(0, 0, 0, 0);
        const float a[4] = {gp.x, gp.y, gp.z, gp.w}, b[4] = {gn.x, gn.y, gn.z, gn.w};
        if (g_p) {
            float r[4];
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) r[kx] = a[0] * O[kx] + a[1] * O[4 + kx] + a[2] * O[8 + kx] + a[3] * O[12 + kx];
            g_p[(size_t)n * n_points + p] = make_float4(r[0], r[1], r[2], r[3]);
        }
        if (g_n) {
            float r[4];
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) r[kx] = b[0] * O[kx] + b[1] * O[4 + kx] + b[2] * O[8 + kx] + b[3] * O[12 + kx];
            g_n[(size_t)n * n_points + p] = make_float4(r[0], r[1], r[2], r[3]);
        }
        if (g_ori) {
            const float4 d4 = __ldg(pts + (size_t)src * n_points + p), e4 = __ldg(nrm + (size_t)src * n_points + p);
            const float d[4] = {d4.x, d4.y, d4.z, d4.w}, e[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int kx = 0; kx < 4; ++kx) acc[j * 4 + kx] += a[j] * d[kx] + b[j] * e[kx];
        }
    }
    if (!g_ori) return;
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = warp_sumf(acc[q]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
        for (int q = 0; q < 16; ++q) red[warp][q] = acc[q];
    __syncthreads();
    if (threadIdx.x < 16) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
        g_ori[(size_t)n * 16 + threadIdx.x] = s;
    }
}

// ---- per-target reduction ---------------------------------------------------------------------
// 256 threads = 8 sample lanes (one warp each) x 32 pixel quads: every warp streams 512 contiguous bytes of one
// sample's bitmap; the 8 partial sums are combined in a fixed order (bit-reproducible).
__global__ void __launch_bounds__(256) bitmaps_per_target_kernel(const float4* __restrict__ bm, const int* __restrict__ tidx,
                                                                 int n_samples, int ue4, float4* __restrict__ out) {
    __shared__ float4 part[8][32];
    const int t = blockIdx.y;
    const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    float4 acc = make_float4(0, 0, 0, 0);
    if (i < ue4) {
        for (int n = sl; n < n_samples; n += 8) {
            if (__ldg(tidx + n) != t) continue;  // uniform across the warp
            const float4 v = __ldcs(bm + (size_t)n * ue4 + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    part[sl][lane] = acc;
    __syncthreads();
    if (sl == 0 && i < ue4) {
        float4 s = part[0][lane];
#pragma unroll
        for (int k = 1; k < 8; ++k) { s.x += part[k][lane].x; s.y += part[k][lane].y; s.z += part[k][lane].z; s.w += part[k][lane].w; }
        out[(size_t)t * ue4 + i] = s;
    }
}

__global__ void bitmaps_per_target_scalar_kernel(const float* __restrict__ bm, const int* __restrict__ tidx, int n_samples,
                                                 int ue, float* __restrict__ out) {
    const int t = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ue) return;
    float acc = 0.f;
    for (int n = 0; n < n_samples; ++n) {
        if (__ldg(tidx + n) != t) continue;
        acc += bm[(size_t)n * ue + i];
    }
    out[(size_t)t * ue + i] = acc;
}

static int32_t validate_kin(const ab200_kinematics_args* k) {
    AB200_REQUIRE(k != nullptr, AB200_EINVAL, "args is NULL");
    AB200_REQUIRE(k->abi_version == AB200_ABI_VERSION, AB200_EINVAL, "abi_version mismatch");
    AB200_REQUIRE(k->n >= 0, AB200_EINVAL, "n < 0");
    AB200_REQUIRE(k->positions && k->translation_dev && k->rotation_dev && k->actuator_non_opt && k->orientation_offset,
                  AB200_EINVAL, "NULL input pointer");
    AB200_REQUIRE(!k->linear_actuators || k->actuator_opt, AB200_EINVAL, "linear actuators need actuator_opt");
    return AB200_OK;
}

}  // namespace ab200

using namespace ab200;

extern "C" int32_t ab200_kinematics_fwd(const ab200_kinematics_args* k, const float* motor, float* out, void* stream) {
    int32_t rc = validate_kin(k);
    if (rc != AB200_OK) return rc;
    AB200_REQUIRE(motor && out, AB200_EINVAL, "NULL pointer");
    if (k->n == 0) return AB200_OK;
    kinematics_fwd_kernel<<<(k->n + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(*k, motor, out);
    note_launch();
    AB200_CUDA_TRY(cudaGetLastError());
    return AB200_OK;
}

extern "C" int32_t ab200_kinematics_bwd(const ab200_kinematics_args* k, const float* motor, const float* gout, float* g_motor,
                                        float* g_rot, float* g_trans, float* g_act, float* g_pos, void* stream) {
    int32_t rc = validate_kin(k);
    if (rc != AB200_OK) return rc;
    AB200_REQUIRE(motor && gout, AB200_EINVAL, "NULL pointer");
    if (k->n == 0) return AB200_OK;
    kinematics_bwd_kernel<<<(k->n + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(*k, motor, gout, g_motor, g_rot, g_trans,
                                                                                         g_act, g_pos);
    note_launch();
    AB200_CUDA_TRY(cudaGetLastError());
    return AB200_OK;
}

extern "C" int32_t ab200_kinematics_align_incident(const ab200_kinematics_args* k, const float* incident, const float* aim,
                                                   int32_t max_iterations, float min_eps, float* orientations,
                                                   float* motor_positions, float* scratch, void* stream) {
    int32_t rc = validate_kin(k);
    if (rc != AB200_OK) return rc;
    AB200_REQUIRE(incident && aim && orientations && motor_positions && scratch, AB200_EINVAL, "NULL pointer");
    AB200_REQUIRE(max_iterations >= 1, AB200_EINVAL, "max_iterations < 1");
    if (k->n == 0) return AB200_OK;
    AB200_REQUIRE(max_iterations <= 6, AB200_ELIMIT, "max_iterations > 6");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (k->n <= kAlignClusterMax && !getenv("AB200_ALIGN_MULTI_KERNEL")) {
        kin_align_cluster_kernel<<<kAlignClusterCtas, kAlignClusterThreads, 0, st>>>(*k, incident, aim, max_iterations, min_eps,
                                                                                  orientations, motor_positions, scratch);
        note_launch();
        AB200_CUDA_TRY(cudaGetLastError());
        return AB200_OK;
    }
    AB200_CUDA_TRY(cudaMemsetAsync(scratch + (size_t)4 * k->n, 0, sizeof(AlignFlags), st));
    const int blocks = (k->n + 63) / 64;
    for (int it = 0; it < max_iterations; ++it) {
        kin_align_forward_kernel<<<blocks, 64, 0, st>>>(*k, incident, aim, it, min_eps, orientations, motor_positions, scratch);
        kin_align_inverse_kernel<<<blocks, 64, 0, st>>>(*k, it, motor_positions, scratch);
        note_launch(2);
    }
    AB200_CUDA_TRY(cudaGetLastError());
    return AB200_OK;
}

extern "C" int32_t ab200_align_fwd(const float* points, const float* normals, const float* orientations, const int32_t* src_row,
                                   int32_t n_samples, int32_t n_points, float* out_points, float* out_normals, void* stream) {
    AB200_REQUIRE(points && normals && orientations && out_points && out_normals, AB200_EINVAL, "NULL pointer");
    AB200_REQUIRE(n_samples >= 0 && n_points > 0, AB200_EINVAL, "bad sizes");
    if (n_samples == 0) return AB200_OK;
    AB200_REQUIRE(n_samples <= 65535, AB200_ELIMIT, "more than 65535 samples per call");
    dim3 grid((unsigned)((n_points + 255) / 256 < 64 ? (n_points + 255) / 256 : 64), (unsigned)n_samples);
    align_fwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(points), reinterpret_cast<const float4*>(normals), orientations, src_row, n_points,
        reinterpret_cast<float4*>(out_points), reinterpret_cast<float4*>(out_normals));
    note_launch();
    AB200_CUDA_TRY(cudaGetLastError());
    return AB200_OK;
}

extern "C" int32_t ab200_align_bwd(const float* points, const float* normals, const float* orientations, const int32_t* src_row,
                                   int32_t n_samples, int32_t n_points, const float* g_op, const float* g_on, float* g_p,
                                   float* g_n, float* g_ori, void* stream) {
    AB200_REQUIRE(points && normals && orientations, AB200_EINVAL, "NULL pointer");
    AB200_REQUIRE(g_op || g_on, AB200_EINVAL, "no upstream gradient given");
    AB200_REQUIRE(n_samples >= 0 && n_points > 0, AB200_EINVAL, "bad sizes");
    if (n_samples == 0) return AB200_OK;
    align_bwd_kernel<<<n_samples, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(points), reinterpret_cast<const float4*>(normals), orientations, src_row, n_points,
        reinterpret_cast<const float4*>(g_op), reinterpret_cast<const float4*>(g_on), reinterpret_cast<float4*>(g_p),
        reinterpret_cast<float4*>(g_n), g_ori);
    note_launch();
    AB200_CUDA_TRY(cudaGetLastError());
    return AB200_OK;
}

extern "C" int32_t ab200_bitmaps_per_target(const float* bitmaps, const int32_t* target_idx, int32_t n_samples, int32_t n_targets,
                                            int32_t res_u, int32_t res_e, float* out, void* stream) {
    AB200_REQUIRE(bitmaps && target_idx && out, AB200_EINVAL, "NULL pointer");
    AB200_REQUIRE(n_samples >= 0 && n_targets > 0 && res_u > 0 && res_e > 0, AB200_EINVAL, "bad sizes");
    const int ue = res_u * res_e;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (ue % 4 == 0) {
        dim3 grid((unsigned)((ue / 4 + 31) / 32), (unsigned)n_targets);
        bitmaps_per_target_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(bitmaps), target_idx, n_samples, ue / 4,
                                                        reinterpret_cast<float4*>(out));
    } else {
        dim3 grid((unsigned)((ue + 255) / 256), (unsigned)n_targets);
        bitmaps_per_target_scalar_kernel<<<grid, 256, 0, st>>>(bitmaps, target_idx, n_samples, ue, out);
    }
    note_launch();
    AB200_CUDA_TRY(cudaGetLastError());
    return AB200_OK;
}

extern "C" int32_t ab200_trace_host(const ab200_host_trace_args* h, void* stream) {
    AB200_REQUIRE(h != nullptr, AB200_EINVAL, "args is NULL");
    const ab200_trace_args* a = &h->dev;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    AB200_REQUIRE(h->h_incident && h->h_target_idx && h->d_target_bitmaps && h->h_target_bitmaps, AB200_EINVAL, "NULL pointer");
    const size_t np4 = (size_t)a->n_samples * a->n_points * 4 * sizeof(float);
    if (h->h_points) AB200_CUDA_TRY(cudaMemcpyAsync(const_cast<float*>(a->points), h->h_points, np4, cudaMemcpyHostToDevice, st));
    if (h->h_normals) AB200_CUDA_TRY(cudaMemcpyAsync(const_cast<float*>(a->normals), h->h_normals, np4, cudaMemcpyHostToDevice, st));
    AB200_CUDA_TRY(cudaMemcpyAsync(const_cast<float*>(a->incident), h->h_incident, (size_t)a->n_samples * 4 * sizeof(float),
                                   cudaMemcpyHostToDevice, st));
    AB200_CUDA_TRY(cudaMemcpyAsync(const_cast<int32_t*>(a->target_idx), h->h_target_idx, (size_t)a->n_samples * sizeof(int32_t),
                                   cudaMemcpyHostToDevice, st));
    int32_t rc = ab200_trace_fwd(a, stream);
    if (rc != AB200_OK) return rc;
    const int n_targets = a->targets.n_planar + a->targets.n_cyl;
    rc = ab200_bitmaps_per_target(a->flux, a->target_idx, a->n_samples, n_targets, a->res_u, a->res_e, h->d_target_bitmaps, stream);
    if (rc != AB200_OK) return rc;
    AB200_CUDA_TRY(cudaMemcpyAsync(h->h_target_bitmaps, h->d_target_bitmaps,
                                   (size_t)n_targets * a->res_u * a->res_e * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (h->h_factors) {
        AB200_CUDA_TRY(cudaMemcpyAsync(h->h_factors, a->intercept, (size_t)a->n_samples * sizeof(float), cudaMemcpyDeviceToHost, st));
        AB200_CUDA_TRY(cudaMemcpyAsync(h->h_factors + a->n_samples, a->on_target, (size_t)a->n_samples * sizeof(float),
                                       cudaMemcpyDeviceToHost, st));
        AB200_CUDA_TRY(cudaMemcpyAsync(h->h_factors + 2 * (size_t)a->n_samples, a->blocking, (size_t)a->n_samples * sizeof(float),
                                       cudaMemcpyDeviceToHost, st));
    }
    AB200_CUDA_TRY(cudaStreamSynchronize(st));
    return AB200_OK;
}
