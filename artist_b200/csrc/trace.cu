// Fused heliostat ray tracer: forward (flux bitmaps + factors) and backward (gradients to the
// aligned surface points / normals).  One CTA per heliostat-sample (or per chunk of its surface
// points when there are few samples), thread <-> surface point, loop over the R rays of the point
// in registers.  See DESIGN.md "Kernels" for the data layout and the roofline of each kernel.
//
// Flux accumulation: a window of the U x E bitmap around the heliostat's focal spot lives in
// shared memory as a 32-bit FIXED-POINT histogram updated with native integer atomics
// (ATOMS.ADD; fp32 shared atomics are a CAS loop on sm_100a - measured 3.4x slower, see
// profiles/r01_atomics_microbench.txt).  Integer adds are order independent, so the bitmap is
// bit-reproducible run to run.  Rays that fall outside the window use integer REDG atomics on
// the (pre-zeroed) output row itself, which is converted to fp32 in place afterwards.
#include <cstdint>
#ifndef AB200_TU_BWD
#define AB200_TU_FWD 1   // this file alone = the forward half (see trace_bwd.cu)
#endif
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <cstdlib>
#include "trace_device.cuh"
#include "blocking_device.cuh"

namespace ab200 {

#ifndef AB200_TU_BWD   // (defined once: the forward translation unit; trace_bwd.cu compiles this file with AB200_TU_BWD)
static thread_local char g_error_detail[512] = "";

void set_error_detail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error_detail, sizeof(g_error_detail), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_kernel_launches{0};
void note_launch(int n) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }
#endif

#ifndef AB200_TU_BWD
int sm_count() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}
#endif

struct TraceParams {
    ab200_trace_args a;
    int split;          // CTAs per heliostat-sample
    int pts_per_chunk;  // surface points per CTA
    int win_cap;        // shared-memory window capacity (32-bit cells)
    float fx_scale;     // fixed-point scale (counts per unit of flux)
    float fx_inv;       // 1 / fx_scale (signed)
    float sigma;        // scatter sigma used for the window margin
    int simple_counts;  // 1: magnitude, (1 - extinction), reflectivity >= 1e-6: lambert > 0 and intensity > 0 hold exactly for valid rays
    int self_zero;      // 1: every CTA clears the part of its bitmap row outside the window itself (no memset pass)
    int wave;           // CTAs per wave (number of SMs) in the one-CTA-per-sample mode, else 0: L2 prefetch distance in samples
    int quad;           // 1: E % 4 == 0 and the bitmap rows are 16-byte aligned: the window is placed on 4-column quads, so
                        //    that clearing, staging and flushing move float4 / uint4
    PackedIdentities ident;  // 1, -0, -1 as run-time values (see common.cuh, packed arithmetic)
};

// Window of the bitmap held in shared memory, in (iu, ie) index space.
struct Window {
    int e0, u0, ww, wh;
};

#ifndef AB200_WIN_SAMPLE_STRIDE
#define AB200_WIN_SAMPLE_STRIDE 32
#endif
// every 32nd surface point places the window: each sampled row is its own 32-byte DRAM sector, and a CTA that floods the
// load queue with thousands of scattered sector requests spends ~10 % of its life waiting for them (measured with
// AB200_PHASE: stride 8 -> 19 K cycles before the first barrier).  Neighbouring points reflect to within a few pixels
// of each other, far inside the 4-sigma margin the window gets anyway.
constexpr int kWindowSampleStride = AB200_WIN_SAMPLE_STRIDE;
// diagnostics (only when ab200_trace_args::stats is given): cycles thread 0 spends in each phase of a CTA, summed over
// the CTAs of a launch: stats[4 + k] forward, stats[12 + k] backward
#define AB200_PHASE(base, k)                                                                                   \
    do {                                                                                                        \
        if (prm.a.stats && threadIdx.x == 0) {                                                                  \
            const long long t_now = clock64();                                                                  \
            atomicAdd(reinterpret_cast<unsigned long long*>(prm.a.stats) + (base) + (k), (unsigned long long)(t_now - t_phase)); \
            t_phase = t_now;                                                                                    \
        }                                                                                                       \
    } while (0)
#ifndef AB200_FWD_THREADS
#define AB200_FWD_THREADS 1024
#endif
#ifndef AB200_BWD_THREADS
#define AB200_BWD_THREADS 768
#endif
#ifndef AB200_WIN_KB
#define AB200_WIN_KB 224   // shared-memory bitmap window per CTA in the one-CTA-per-sample mode
#endif
#ifndef AB200_BWD_TMA_STAGE
#define AB200_BWD_TMA_STAGE 1   // 1: the backward stages its gradient window with TMA bulk copies (cp.async.bulk + mbarrier)
#endif
#ifndef AB200_BWD_SPLIT_TAIL
#define AB200_BWD_SPLIT_TAIL 1   // 1: the backward's last, partial round of points is split by ray pair (see bwd_rays_planar_fast2)
#endif
#ifndef AB200_FLUSH_ST
#define AB200_FLUSH_ST 0   // store flavour of the window flush: 0 st.cs (streaming), 1 plain, 2 st.cg, 3 st.wt (tuning)
#endif
#ifndef AB200_PACKED_RAYS
#define AB200_PACKED_RAYS 1     // fast loops process two rays per iteration with fp32x2 (FFMA2) arithmetic
#endif
#ifndef AB200_RAY_UNROLL
#define AB200_RAY_UNROLL 1
#endif
constexpr int kFwdThreadsLarge = AB200_FWD_THREADS;  // one CTA per SM with a ~200 KB bitmap window
constexpr int kBwdThreadsLarge = AB200_BWD_THREADS;
constexpr int kRayUnroll = AB200_RAY_UNROLL;

// ---------------------------------------------------------------------------------------------
// window placement: bounding box of the undistorted reflections of a subset of the CTA's points,
// grown by ~4 sigma of the sun shape projected onto the target.  Correctness never depends on
// the window (misses take the global path); it only decides how many rays take the fast path.
// ---------------------------------------------------------------------------------------------
// the (un-oriented) rows of the two window-sample points of a thread: loaded early, so that their DRAM latency
// overlaps the CTA's other start-up loads (target constants, orientation, incident direction)
struct WindowSamples {
    float4 oa, na, ob, nb;
};
template <int THREADS>
__device__ __forceinline__ void load_window_samples(WindowSamples& ws, const float4* pts, const float4* nrm, int p_begin, int p_end) {
    const int pa = p_begin + threadIdx.x * kWindowSampleStride, pb = pa + THREADS * kWindowSampleStride;
    ws.oa = ws.na = ws.ob = ws.nb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pa < p_end) { ws.oa = __ldg(pts + pa); ws.na = __ldg(nrm + pa); }
    if (pb < p_end) { ws.ob = __ldg(pts + pb); ws.nb = __ldg(nrm + pb); }
}

template <int THREADS>
__device__ void place_window(Window& win_out, const TraceParams& prm, const TargetCtx& T, const PointSrc& src,
                             const WindowSamples& ws, int p_begin, int p_end, float i0, float i1, float i2,
                             float* red /* [6*32] */, Window* win_sh, const int4* recorded = nullptr) {
    const int tid = threadIdx.x;
    const float4* pts = src.pts;
    const float4* nrm = src.nrm;
    const float inf = __int_as_float(0x7f800000);
    float emin = inf, emax = -inf, umin = inf, umax = -inf, tmax = 0.f, cmin = inf;
    {
        // every kWindowSampleStride-th point; the two candidate rows were loaded by the caller before its first barrier
        const int pa = p_begin + tid * kWindowSampleStride, pb = pa + THREADS * kWindowSampleStride;
        const bool has_a = pa < p_end, has_b = pb < p_end;
        float4 oa = ws.oa, na = ws.na, ob = ws.ob, nb = ws.nb;
        orient_point(src, oa, na);
        orient_point(src, ob, nb);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (!(k == 0 ? has_a : has_b)) continue;
            PointCtx pc;
            make_point(pc, T, i0, i1, i2, k == 0 ? oa : ob, k == 0 ? na : nb);
            float be, bu, t, cosi;
            const bool ok = T.planar ? centre_planar(T, pc, be, bu, t, cosi) : centre_cylinder(T, pc, be, bu, t, cosi);
            if (ok && be == be && bu == bu && fabsf(be) < 1e6f && fabsf(bu) < 1e6f) {
                emin = fminf(emin, be); emax = fmaxf(emax, be);
                umin = fminf(umin, bu); umax = fmaxf(umax, bu);
                tmax = fmaxf(tmax, t);  cmin = fminf(cmin, cosi);
            }
        }
        for (int p = pb + THREADS * kWindowSampleStride; p < p_end; p += THREADS * kWindowSampleStride) {  // very large P only
            PointCtx pc;
            float4 o4 = __ldg(pts + p), n4 = __ldg(nrm + p);
            orient_point(src, o4, n4);
            make_point(pc, T, i0, i1, i2, o4, n4);
            float be, bu, t, cosi;
            const bool ok = T.planar ? centre_planar(T, pc, be, bu, t, cosi) : centre_cylinder(T, pc, be, bu, t, cosi);
            if (ok && be == be && bu == bu && fabsf(be) < 1e6f && fabsf(bu) < 1e6f) {
                emin = fminf(emin, be); emax = fmaxf(emax, be);
                umin = fminf(umin, bu); umax = fmaxf(umax, bu);
                tmax = fmaxf(tmax, t);  cmin = fminf(cmin, cosi);
            }
        }
    }
    emin = warp_min(emin); emax = warp_max(emax); umin = warp_min(umin); umax = warp_max(umax);
    tmax = warp_max(tmax); cmin = warp_min(cmin);
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0) {
        red[warp] = emin; red[32 + warp] = emax; red[64 + warp] = umin; red[96 + warp] = umax;
        red[128 + warp] = tmax; red[160 + warp] = cmin;
    }
    __syncthreads();
    if (warp == 0) {
        const int nw = THREADS / 32;
        emin = lane < nw ? red[lane] : inf;         emax = lane < nw ? red[32 + lane] : -inf;
        umin = lane < nw ? red[64 + lane] : inf;    umax = lane < nw ? red[96 + lane] : -inf;
        tmax = lane < nw ? red[128 + lane] : 0.f;   cmin = lane < nw ? red[160 + lane] : inf;
        emin = warp_min(emin); emax = warp_max(emax); umin = warp_min(umin); umax = warp_max(umax);
        tmax = warp_max(tmax); cmin = warp_min(cmin);
        if (lane == 0) {
            Window w;
            const int E = prm.a.res_e, U = prm.a.res_u;
            if (!(emin <= emax) || prm.win_cap < 4) {
                w.e0 = 1 << 28; w.u0 = 1 << 28; w.ww = 0; w.wh = 0;  // nothing takes the fast path
            } else {
                const float m = 4.0f * prm.sigma * tmax / fmaxf(cmin, 0.1f);
                const float me = m * T.px_per_m_e + 2.0f, mu = m * T.px_per_m_u + 2.0f;
                int e_lo = max(0, (int)floorf(emin - me)), e_hi = min(E - 1, (int)ceilf(emax + me) + 1);
                int u_lo = max(0, (int)floorf(umin - mu)), u_hi = min(U - 1, (int)ceilf(umax + mu) + 1);
                int ww = max(0, e_hi - e_lo + 1), wh = max(0, u_hi - u_lo + 1);
                if ((long long)ww * wh > prm.win_cap) {
                    // keep the aspect ratio, centre on the bounding box
                    const float sc = sqrtf((float)prm.win_cap / ((float)ww * (float)wh));
                    int nww = max(2, (int)(ww * sc)), nwh = max(2, (int)(wh * sc));
                    while ((long long)nww * nwh > prm.win_cap) { if (nww > nwh) --nww; else --nwh; }
                    e_lo += (ww - nww) / 2; u_lo += (wh - nwh) / 2;
                    ww = nww; wh = nwh;
                }
                if (prm.quad && ww >= 2 && wh >= 2) {   // whole quads: widen to multiples of 4 columns, give rows back if needed
                    const int a_lo = e_lo & ~3, a_hi = min(E, (e_lo + ww + 3) & ~3);
                    const int aw = a_hi - a_lo;
                    if ((long long)aw * wh > prm.win_cap) { const int nwh = prm.win_cap / aw; u_lo += (wh - nwh) / 2; wh = nwh; }
                    e_lo = a_lo; ww = aw;
                }
                if (ww < 2 || wh < 2) { w.e0 = 1 << 28; w.u0 = 1 << 28; w.ww = 0; w.wh = 0; }
                else { w.e0 = e_lo; w.u0 = u_lo; w.ww = ww; w.wh = wh; }
            }
            if (recorded) {   // the window the forward of the same call placed (trusted only as far as it is harmless); the
                              // rest of the placement is kept as it is: any early exit here costs the ray loop its
                              // register allocation (spills inside the loop, measured +25 %)
                const int4 r = __ldg(recorded);
                const int E = prm.a.res_e, U = prm.a.res_u;
                if (r.z >= 2 && r.w >= 2 && r.x >= 0 && r.y >= 0 && r.x + r.z <= E && r.y + r.w <= U &&
                    (long long)r.z * r.w <= prm.win_cap && (!prm.quad || ((r.x | r.z) & 3) == 0)) {
                    w.e0 = r.x; w.u0 = r.y; w.ww = r.z; w.wh = r.w;
                }
            }
            *win_sh = w;
        }
    }
    __syncthreads();
    win_out = *win_sh;
}

// The sample the CTA one wave later will trace (one-CTA-per-sample mode): its rows for the window placement are pulled into
// L2 by this CTA's threads during their last surface point (fwd_/bwd_rays_planar_fast2), its per-sample scalars here.
// MAP (compile time): the kernel reads its surface rows through ab200_trace_args::src_rows.  Looked up at run time behind a
// NULL test, the row index - and every address derived from it - left the uniform datapath, and the forward of the plain
// kernels was 2.6 % slower even without a map (found by bisection, tools/time_fwd.py): so the map is its own instantiation.
template <bool MAP>
__device__ inline void set_next_sample(const float4** next_sh, const TraceParams& prm, int li) {
    next_sh[0] = nullptr; next_sh[1] = nullptr;
    if (!MAP && prm.wave > 0 && li + prm.wave < prm.a.n_local) {   // (with a map the next sample's rows are not prefetched)
        const int hn = prm.a.local_rows ? prm.a.local_rows[li + prm.wave] : li + prm.wave;
        const int sn = hn;
        next_sh[0] = reinterpret_cast<const float4*>(prm.a.points) + (size_t)sn * prm.a.n_points;
        next_sh[1] = reinterpret_cast<const float4*>(prm.a.normals) + (size_t)sn * prm.a.n_points;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(prm.a.incident + 4 * hn));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(prm.a.target_idx + hn));
        if (prm.a.orientations) asm volatile("prefetch.global.L2 [%0];" ::"l"(prm.a.orientations + (size_t)hn * 16));
    }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
// The quads (4 columns x 1 row, 16 bytes) of a [U,E] bitmap row that lie OUTSIDE a quad-aligned window, enumerated
// without gaps: first the full rows below / above the window, then the strips left and right of it.
struct OutsideQuads {
    int e4, q_lo, qw, side, r_lo, wh, n_full, n_out;
    __device__ OutsideQuads(const Window& W, int E, int U) {
        e4 = E >> 2; q_lo = W.e0 >> 2; qw = W.ww >> 2; side = e4 - qw; wh = W.wh;
        if (W.ww <= 0) { q_lo = 0; qw = 0; side = e4; wh = 0; }
        r_lo = U - W.u0 - W.wh;   // first output row of the window (output rows are flipped)
        n_full = (U - wh) * e4; n_out = n_full + wh * side;
    }
    __device__ __forceinline__ int at(int i) const {   // quad index (row * e4 + q) of the i-th outside quad
        int row, q;
        if (i < n_full) { row = i / e4; q = i - row * e4; if (row >= r_lo) row += wh; }
        else { const int j = i - n_full, rr = j / side, qq = j - rr * side; row = r_lo + rr; q = qq < q_lo ? qq : qq + qw; }
        return row * e4 + q;
    }
};

struct FwdCtx {
    unsigned* win_u;     // shared-memory window (fixed point) ...
    float* win_f;        // ... or fp32 (AB200_FLAG_FP32_ACCUM)
    float* out_f;        // this sample's [U,E] output row
    int e0, u0, ww, wwm1, whm1;  // window origin, row pitch, (width-1), (height-1); an empty window has wwm1 = whm1 = 0
    int wh;
    int* fb_box;         // shared: bounding box (output rows/cols) of the taps that took the global path
    const BlockPrim* blk; // shared: candidate blocking primitives of this sample
    int n_blk;
    BlockParams bp;
};

// The per-ray loop of one CTA, specialised on the target type and on whether the exact constant-divisor quotient
// may be used, so that the hot loop carries no target-type branch.
// "Irregular" rays are the ones the branch-free fast loops skip: scatter angles beyond the polynomial's range,
// a grazing ray (|d.n_t| < 1e-18) or a degenerate point (|(c-o).n_t| outside [1e-18, 1e18]) - none of which occur
// with physical inputs.  A CTA that saw any re-runs the generic loop for just those rays (ONLY_IRREGULAR).
__device__ __forceinline__ bool point_regular(const PointCtx& pc) {
    const float n = fabsf(pc.num);
    return n > 1e-18f && n < 1e18f;
}
template <int TRIG>
__device__ __forceinline__ bool angles_regular(float u, float e) {
    return TRIG == AB200_TRIG_TABLE ? true : (fmaxf(fabsf(u), fabsf(e)) <= 0.785f);
}
__device__ __forceinline__ bool cosine_regular(float a) { return !(a < 0.0f) || (a < -1e-18f && a > -1e18f); }

__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }
enum { kDeferOff = 0, kDeferRecord = 1, kDeferList = 2, kDeferSkip = 3 };

template <int THREADS, int TRIG, bool DBG, bool FP32ACC, bool PLANAR, bool FASTDIV, bool ONLY_IRREGULAR, bool BLK>
__device__ __forceinline__ void fwd_rays(const TraceParams& prm, const TargetCtx& T, const FwdCtx& fc, const PointSrc& src, int h, int p_begin,
                                         int p_end, float i0, float i1, float i2, int& cnt_lam_out, int& cnt_int_out,
                                         int& cnt_blk_out, bool& fell_back_out, const DeferCtx* dc = nullptr,
                                         int dmode = kDeferOff) {
    const int tid = threadIdx.x;
    const int P = prm.a.n_points, R = prm.a.n_rays, E = prm.a.res_e, U = prm.a.res_u;
    const float4* pts = src.pts;
    const float4* nrm = src.nrm;
    const float2* dist = reinterpret_cast<const float2*>(prm.a.distortions) + (size_t)h * R * P;
    const float4* trig = (TRIG == AB200_TRIG_TABLE) ? reinterpret_cast<const float4*>(prm.a.trig) + (size_t)h * R * P : nullptr;
    unsigned* out_u = reinterpret_cast<unsigned*>(fc.out_f);
    const float mag = prm.a.ray_magnitude, ome = prm.a.one_minus_extinction, refl = prm.a.reflectivity;
    const float fxs = prm.fx_scale;
    int cnt_lam = 0, cnt_int = 0, cnt_blk = 0;
    bool fell_back = false;

    // (blocking_device.cuh, "Deferral"): kDeferRecord = pass 1 (this instantiation has no blocking code; shadow-affected
    // points are recorded and skipped), kDeferList = pass 2 over the recorded points, kDeferSkip = the irregular-ray pass
    // behind a deferring fast loop (recorded points are traced completely by pass 2)
    const int n_items = dmode == kDeferList ? dc->count : p_end - p_begin;
    int p_list = (dmode == kDeferList && tid < n_items) ? p_begin + defer_lookup(*dc, tid) : 0;
    for (int it = tid; it < n_items; it += THREADS) {
        const int p = dmode == kDeferList ? p_list : p_begin + it;
        if (dmode == kDeferList) {
            // the recorded points are scattered over the surface: no coalescing, no streaming pattern - the rows and the R
            // distortion pairs of the thread's NEXT point are pulled into L2 while this one is traced
            if (it + THREADS < n_items) {
                p_list = p_begin + defer_lookup(*dc, it + THREADS);
                prefetch_l2(pts + p_list); prefetch_l2(nrm + p_list);
                for (int r = 0; r < R; ++r) prefetch_l2(dist + (size_t)r * P + p_list);
            }
        }
        PointCtx pc;
        {
            float4 o4 = __ldg(pts + p), n4 = __ldg(nrm + p);
            orient_point(src, o4, n4);
            make_point(pc, T, i0, i1, i2, o4, n4);
        }
        bool point_dead = false;   // pass 1: every ray of the point is completely shadowed (counted, not splatted)
        if (!BLK && dmode == kDeferRecord) {
            const int pcls = defer_point(dc, p - p_begin, pc.o0, pc.o1, pc.o2, pc.r0, pc.r1, pc.r2);
            if (pcls == kPointDeferred) continue;
            point_dead = pcls == kPointDead;
        }
        const unsigned long long bmask = (BLK && fc.n_blk) ? block_point_mask(fc.blk, fc.n_blk, fc.bp, pc.o0, pc.o1, pc.o2, pc.r0, pc.r1, pc.r2) : 0ull;
        if (BLK && dmode == kDeferSkip && bmask) continue;
        const float2* dp = dist + p;
        float2 d_next = __ldcs(dp);
        for (int r = 0; r < R; ++r) {
            const float2 d = d_next;
            dp += P;
            if (r + 1 < R) d_next = __ldcs(dp);
            Scatter s;
            ray_trig<TRIG>(d.x, d.y, trig, (size_t)r * P + p, s.cu, s.su, s.ce, s.se);
            scatter(s, pc);
            Hit hit;
            if (PLANAR) hit_planar<FASTDIV>(hit, T, pc, s, mag); else hit_cylinder<FASTDIV, TRIG == AB200_TRIG_TABLE>(hit, T, pc, s, mag);
            if (ONLY_IRREGULAR && point_regular(pc) && angles_regular<TRIG>(d.x, d.y) && cosine_regular(hit.a)) continue;
            float blocked = 0.0f;
            if (BLK) {
                if (bmask) {
                    const int cls = block_classify(fc.blk, bmask, fc.bp, pc.o0, pc.o1, pc.o2, s.dx, s.dy, s.dz);
                    if (cls) blocked = cls == 1 ? 1.0f : block_eval(fc.blk, bmask, fc.bp, pc.o0, pc.o1, pc.o2, s.dx, s.dy, s.dz);
                }
                cnt_blk += (blocked < 1e-3f);
            } else if (dmode == kDeferRecord) {
                if (point_dead) { cnt_lam += (hit.lam > 0.0f); continue; }   // blocked == 1: no intensity
                cnt_blk += 1;   // an unshadowed point of a blocking trace: blocked == 0
            }
            // intensities = lambert * (1 - blocked) * (1 - extinction) * reflectivity   (:482-487); lambert * 1 is exact
            const float inten = BLK ? smul(smul(smul(hit.lam, ssub(1.0f, blocked)), ome), refl) : smul(smul(hit.lam, ome), refl);
            if (DBG) {
                const size_t q = ((size_t)h * R + r) * P + p;
                if (prm.a.dbg_be) prm.a.dbg_be[q] = hit.be;
                if (prm.a.dbg_bu) prm.a.dbg_bu[q] = hit.bu;
                if (prm.a.dbg_t) prm.a.dbg_t[q] = hit.t;
                if (prm.a.dbg_lambert) prm.a.dbg_lambert[q] = hit.lam;
            }
            cnt_lam += (hit.lam > 0.0f);
            cnt_int += (inten > 0.0f);
            if (!hit.valid) continue;
            Splat sp;
            splat_weights(sp, hit.be, hit.bu, E, U);
            if (!sp.on) continue;
            const int ce = sp.ie - fc.e0, cu = sp.iu - fc.u0;
            const bool fast = ((unsigned)ce < (unsigned)fc.wwm1) && ((unsigned)cu < (unsigned)fc.whm1);
            if (FP32ACC) {
                const float v1 = smul(smul(sp.wle, sp.whu), inten), v2 = smul(smul(sp.whe, sp.whu), inten);
                const float v3 = smul(smul(sp.whe, sp.wlu), inten), v4 = smul(smul(sp.wle, sp.wlu), inten);
                if (fast) {
                    float* b = fc.win_f + cu * fc.ww + ce;
                    atomicAdd(b + fc.ww, v1); atomicAdd(b + fc.ww + 1, v2); atomicAdd(b + 1, v3); atomicAdd(b, v4);
                } else {
                    float* row_hi = fc.out_f + (size_t)(U - 1 - (sp.iu + 1)) * E + sp.ie;
                    float* row_lo = row_hi + E;
                    atomicAdd(row_hi, v1); atomicAdd(row_hi + 1, v2); atomicAdd(row_lo + 1, v3); atomicAdd(row_lo, v4);
                }
            } else {
                const float ahi = fabsf(sp.whu * inten) * fxs, alo = fabsf(sp.wlu * inten) * fxs;
                const unsigned q1 = __float2uint_rn(sp.wle * ahi), q2 = __float2uint_rn(sp.whe * ahi);
                const unsigned q3 = __float2uint_rn(sp.whe * alo), q4 = __float2uint_rn(sp.wle * alo);
                if (fast) {
                    unsigned* b = fc.win_u + cu * fc.ww + ce;
                    atomicAdd(b + fc.ww, q1); atomicAdd(b + fc.ww + 1, q2); atomicAdd(b + 1, q3); atomicAdd(b, q4);
                } else {
                    // per-tap routing: a tap inside the window region must go to shared memory so that window pixels
                    // are owned by shared memory only
                    fell_back = true;
                    atomicMin(fc.fb_box + 0, U - 2 - sp.iu); atomicMax(fc.fb_box + 1, U - 1 - sp.iu);
                    atomicMin(fc.fb_box + 2, sp.ie); atomicMax(fc.fb_box + 3, sp.ie + 1);
                    const bool e_in0 = (unsigned)ce < (unsigned)fc.ww, e_in1 = (unsigned)(ce + 1) < (unsigned)fc.ww;
                    const bool u_in0 = (unsigned)cu < (unsigned)fc.wh, u_in1 = (unsigned)(cu + 1) < (unsigned)fc.wh;
                    unsigned* g_hi = out_u + (size_t)(U - 1 - (sp.iu + 1)) * E + sp.ie;
                    unsigned* g_lo = g_hi + E;
                    unsigned* b = fc.win_u + cu * fc.ww + ce;
                    if (u_in1 && e_in0) atomicAdd(b + fc.ww, q1); else atomicAdd(g_hi, q1);
                    if (u_in1 && e_in1) atomicAdd(b + fc.ww + 1, q2); else atomicAdd(g_hi + 1, q2);
                    if (u_in0 && e_in1) atomicAdd(b + 1, q3); else atomicAdd(g_lo + 1, q3);
                    if (u_in0 && e_in0) atomicAdd(b, q4); else atomicAdd(g_lo, q4);
                }
            }
        }
    }
    cnt_lam_out = cnt_lam; cnt_int_out = cnt_int; cnt_blk_out = cnt_blk; fell_back_out = fell_back;
}

// Branch-free fast loop for planar targets: polynomial (or table) trig, range-guarded exact divisions, everything
// predicated; irregular rays are only counted (n_irregular) and left to the generic loop.
template <int THREADS, int TRIG, bool DBG, bool FP32ACC, bool BLK>
__device__ __forceinline__ void fwd_rays_planar_fast(const TraceParams& prm, const TargetCtx& T, const FwdCtx& fc, const PointSrc& src, int h,
                                                     int p_begin, int p_end, float i0, float i1, float i2, int& cnt_lam_out,
                                                     int& cnt_int_out, int& cnt_blk_out, bool& fell_back_out,
                                                     int& n_irregular_out) {
    const int tid = threadIdx.x;
    const int P = prm.a.n_points, R = prm.a.n_rays, E = prm.a.res_e, U = prm.a.res_u;
    const float4* pts = src.pts;
    const float4* nrm = src.nrm;
    const float2* dist = reinterpret_cast<const float2*>(prm.a.distortions) + (size_t)h * R * P;
    const float4* trig = (TRIG == AB200_TRIG_TABLE) ? reinterpret_cast<const float4*>(prm.a.trig) + (size_t)h * R * P : nullptr;
    unsigned* out_u = reinterpret_cast<unsigned*>(fc.out_f);
    const float mag = prm.a.ray_magnitude, ome = prm.a.one_minus_extinction, refl = prm.a.reflectivity;
    const float fxs = prm.fx_scale;
    const float e_lim = (float)E, u_lim = (float)U;
    int cnt_lam = 0, cnt_int = 0, cnt_blk = 0, n_irr = 0;
    bool fell_back = false;

    for (int p = p_begin + tid; p < p_end; p += THREADS) {
        PointCtx pc;
        {
            float4 o4 = __ldg(pts + p), n4 = __ldg(nrm + p);
            orient_point(src, o4, n4);
            make_point(pc, T, i0, i1, i2, o4, n4);
        }
        if (!point_regular(pc)) { n_irr += R; continue; }
        const unsigned long long bmask = (BLK && fc.n_blk) ? block_point_mask(fc.blk, fc.n_blk, fc.bp, pc.o0, pc.o1, pc.o2, pc.r0, pc.r1, pc.r2) : 0ull;
        const float2* dp = dist + p;
        float2 d_next = __ldcs(dp);
#pragma unroll kRayUnroll
        for (int r = 0; r < R; ++r) {
            const float2 d = d_next;
            dp += P;
            if (r + 1 < R) d_next = __ldcs(dp);
            Scatter s;
            if (TRIG == AB200_TRIG_TABLE) {
                const float4 t4 = __ldg(trig + (size_t)r * P + p);
                s.cu = t4.x; s.su = t4.y; s.ce = t4.z; s.se = t4.w;
            } else {
                sincos_poly_core(d.x, &s.su, &s.cu);
                sincos_poly_core(d.y, &s.se, &s.ce);
            }
            scatter(s, pc);
            const float a = sadd(sadd(smul(s.dx, T.n0), smul(s.dy, T.n1)), smul(s.dz, T.n2));
            if (!(angles_regular<TRIG>(d.x, d.y) && cosine_regular(a))) { ++n_irr; continue; }
            const bool ff = a < 0.0f;
            const float t = div_regular(pc.num, a);                       // == __fdiv_rn for regular operands
            const float X = sadd(pc.o0, smul(s.dx, t));
            const float Z = sadd(pc.o2, smul(s.dz, t));
            const float te = ssub(sadd(X, T.half_w), T.c0);
            const float tu = ssub(sadd(Z, T.half_h), T.c2);
            const float be0 = smul(const_div(te, T.w, T.rw), T.em1);
            const float bu0 = smul(const_div(tu, T.h, T.rh), T.um1);
            const bool valid = ff && (0.0f <= be0) && (be0 <= T.em1) && (0.0f <= bu0) && (bu0 <= T.um1);
            const float lam = valid ? smul(mag, -a) : 0.0f;
            float blocked = 0.0f;
            if (BLK) {
                if (bmask) {
                    const int cls = block_classify(fc.blk, bmask, fc.bp, pc.o0, pc.o1, pc.o2, s.dx, s.dy, s.dz);
                    if (cls) blocked = cls == 1 ? 1.0f : block_eval(fc.blk, bmask, fc.bp, pc.o0, pc.o1, pc.o2, s.dx, s.dy, s.dz);
                }
                cnt_blk += (blocked < 1e-3f);
            }
            const float inten = BLK ? smul(smul(smul(lam, ssub(1.0f, blocked)), ome), refl) : smul(smul(lam, ome), refl);
            const float be = ssub(T.em1, valid ? be0 : 0.0f);
            const float bu = valid ? bu0 : 0.0f;
            if (DBG) {
                const size_t q = ((size_t)h * R + r) * P + p;
                if (prm.a.dbg_be) prm.a.dbg_be[q] = be;
                if (prm.a.dbg_bu) prm.a.dbg_bu[q] = bu;
                if (prm.a.dbg_t) prm.a.dbg_t[q] = valid ? t : 0.0f;
                if (prm.a.dbg_lambert) prm.a.dbg_lambert[q] = lam;
            }
            cnt_lam += (lam > 0.0f);
            cnt_int += (inten > 0.0f);
            // splat weights (valid rays have 0 <= be <= E-1, 0 <= bu <= U-1)
            const float fe = truncf(be), fu = truncf(bu);
            const float fe1 = sadd(fe, 1.0f), fu1 = sadd(fu, 1.0f);
            const int ie = __float2int_rz(be), iu = __float2int_rz(bu);
            const bool go = valid && (fe1 < e_lim) && (fu1 < u_lim);
            const float wle = ssub(fe1, be), wlu = ssub(fu1, bu), whe = ssub(be, fe), whu = ssub(bu, fu);
            const int ce = ie - fc.e0, cu = iu - fc.u0;
            const bool fast = ((unsigned)ce < (unsigned)fc.wwm1) && ((unsigned)cu < (unsigned)fc.whm1);
            if (FP32ACC) {
                if (go) {
                    const float v1 = smul(smul(wle, whu), inten), v2 = smul(smul(whe, whu), inten);
                    const float v3 = smul(smul(whe, wlu), inten), v4 = smul(smul(wle, wlu), inten);
                    if (fast) {
                        float* b = fc.win_f + cu * fc.ww + ce;
                        atomicAdd(b + fc.ww, v1); atomicAdd(b + fc.ww + 1, v2); atomicAdd(b + 1, v3); atomicAdd(b, v4);
                    } else {
                        float* row_hi = fc.out_f + (size_t)(U - 1 - (iu + 1)) * E + ie;
                        float* row_lo = row_hi + E;
                        atomicAdd(row_hi, v1); atomicAdd(row_hi + 1, v2); atomicAdd(row_lo + 1, v3); atomicAdd(row_lo, v4);
                    }
                }
            } else {
                const float ahi = fabsf(whu * inten) * fxs, alo = fabsf(wlu * inten) * fxs;
                const unsigned q1 = __float2uint_rn(wle * ahi), q2 = __float2uint_rn(whe * ahi);
                const unsigned q3 = __float2uint_rn(whe * alo), q4 = __float2uint_rn(wle * alo);
                if (go && fast) {
                    unsigned* b = fc.win_u + cu * fc.ww + ce;
                    atomicAdd(b + fc.ww, q1); atomicAdd(b + fc.ww + 1, q2); atomicAdd(b + 1, q3); atomicAdd(b, q4);
                } else if (go) {
                    fell_back = true;
                    atomicMin(fc.fb_box + 0, U - 2 - iu); atomicMax(fc.fb_box + 1, U - 1 - iu);
                    atomicMin(fc.fb_box + 2, ie); atomicMax(fc.fb_box + 3, ie + 1);
                    const bool e_in0 = (unsigned)ce < (unsigned)fc.ww, e_in1 = (unsigned)(ce + 1) < (unsigned)fc.ww;
                    const bool u_in0 = (unsigned)cu < (unsigned)fc.wh, u_in1 = (unsigned)(cu + 1) < (unsigned)fc.wh;
                    unsigned* g_hi = out_u + (size_t)(U - 1 - (iu + 1)) * E + ie;
                    unsigned* g_lo = g_hi + E;
                    unsigned* b = fc.win_u + cu * fc.ww + ce;
                    if (u_in1 && e_in0) atomicAdd(b + fc.ww, q1); else atomicAdd(g_hi, q1);
                    if (u_in1 && e_in1) atomicAdd(b + fc.ww + 1, q2); else atomicAdd(g_hi + 1, q2);
                    if (u_in0 && e_in1) atomicAdd(b + 1, q3); else atomicAdd(g_lo + 1, q3);
                    if (u_in0 && e_in0) atomicAdd(b, q4); else atomicAdd(g_lo, q4);
                }
            }
        }
    }
    cnt_lam_out = cnt_lam; cnt_int_out = cnt_int; cnt_blk_out = cnt_blk; fell_back_out = fell_back; n_irregular_out = n_irr;
}


// a in (-1e18, -1e-18), i.e. front-facing AND regular: ONE unsigned range test on the bit pattern
// (bits(-1e-18f) = 0xA19392EF, bits(-1e18f) = 0xDD5E0B6B; negative floats order like their bit patterns)
__device__ __forceinline__ bool front_regular(float a) {
    return (__float_as_uint(a) - 0xA19392F0u) < (0xDD5E0B6Bu - 0xA19392F0u);
}

// Packed variant of the fast loop: two rays of a point per iteration in fp32x2 registers (FFMA2), which halves the
// issue slots of the floating-point part.  Per element the operations and roundings are exactly those of the scalar
// loop (strict products/sums are identity-FMAs, see common.cuh), so the results are bit-identical to it.
// Preconditions (checked by the caller): R even; and - unless BLK - ray_magnitude, (1 - extinction) and reflectivity
// all >= 1e-6, so that "lambert > 0" and "intensity > 0" hold exactly for the valid rays (the factor counters then
// count valid rays instead of comparing two floats per ray).
// Integer ALU work is kept off the half-rate pipe where possible: validity is an unsigned compare of bit patterns
// (0 <= x <= lim  <=>  bits(x) <= bits(lim) once -0 is excluded, which the +0 addend of the last product does),
// pixel indices come from the magic-number add, and the on-bitmap test is implied by the window test.
// The distortion pair of the NEXT iteration (also across the thread's point boundary) is always in flight, and the
// next point's rows are prefetched into L2 one point ahead, so a point start costs one L2 hit instead of two serial
// DRAM round trips.
// REVEN: the number of rays per point is even (every scenario of the reference: 4, 10, 100, 170, 180, 200), so both
// lanes of every pair are live and the loads of the software pipeline need no per-lane conditions.
// The loop of the plain (no blocking) kernels, exactly as round 1 left it.  The blocking trace uses the variant below
// (fwd_rays_planar_fast2: deferral hooks of pass 1, list walk of pass 2); keeping the two apart keeps the register
// allocation and address arithmetic of the headline kernel out of reach of the blocking bookkeeping (with the hooks merged
// in behind compile-time switches the loop still came out 2.6 instructions per ray longer: 1.055 instead of 1.02 ms).
template <int THREADS, int TRIG, bool DBG, bool FP32ACC, bool BLK, bool REVEN>
__device__ __forceinline__ void fwd_rays_planar_fast2_plain(const TraceParams& prm, const TargetCtx& T, const FwdCtx& fc, const PointSrc& src, int h,
                                                      int p_begin, int p_end, float i0, float i1, float i2, int& cnt_lam_out,
                                                      int& cnt_int_out, int& cnt_blk_out, bool& fell_back_out,
                                                      int& n_irregular_out) {
    const int tid = threadIdx.x;
    const int P = prm.a.n_points, R = prm.a.n_rays, E = prm.a.res_e, U = prm.a.res_u;
    const float4* pts = src.pts;
    const float4* nrm = src.nrm;
    const float2* dist = reinterpret_cast<const float2*>(prm.a.distortions) + (size_t)h * R * P;
    const float4* trig = (TRIG == AB200_TRIG_TABLE) ? reinterpret_cast<const float4*>(prm.a.trig) + (size_t)h * R * P : nullptr;
    unsigned* out_u = reinterpret_cast<unsigned*>(fc.out_f);
    const Packed K(prm.ident);
    const float mag = prm.a.ray_magnitude, ome = prm.a.one_minus_extinction, refl = prm.a.reflectivity;
    const float fxs_sub = prm.fx_scale * 0x1p-100f;   // (exact) see the tap rounding below
    const float kTwoM49 = 0x1p-49f;
    const unsigned em1_bits = __float_as_uint(T.em1), um1_bits = __float_as_uint(T.um1);
    const bool axis_n = (T.n0 == 0.0f) && (T.n2 == 0.0f);
    // Pixel indices relative to the window come straight out of the floor() trick: be + (1.5 * 2^23 - e0), rounded down,
    // has the bit pattern 0x4B400000 + (floor(be) - e0) (two's complement wrap-around for negative differences), so
    // cex = bits - 0x4B400000 costs ONE integer instruction with an immediate operand.  The window's shared-memory
    // address is kept opaque so that it stays in a register instead of being re-derived for every tap.
    const int e0w = fc.ww > 0 ? fc.e0 : 0, u0w = fc.ww > 0 ? fc.u0 : 0;   // (an empty window has a huge origin)
    float magic_e = 12582912.0f - (float)e0w, magic_u = 12582912.0f - (float)u0w;
    asm volatile("" : "+f"(magic_e), "+f"(magic_u));   // loop invariants: kept in registers, not re-derived per pair
    const unsigned kIdxBits = 0x4B400000u;
    unsigned wwm1 = (unsigned)max(fc.ww, 1) - 1u, whm1 = (unsigned)max(fc.wh, 1) - 1u;   // empty window: nothing is "inside"
    asm volatile("" : "+r"(wwm1), "+r"(whm1));
    unsigned win_base = (unsigned)__cvta_generic_to_shared(fc.win_u);
    asm volatile("" : "+r"(win_base));
    const unsigned row_bytes = (unsigned)fc.ww * 4u;
    // factor counters: valid rays = R * (regular points) - cnt_bad; invalid rays are rare, so the bookkeeping (and the
    // irregularity flag, which implies invalid) lives in a rarely taken branch
    int cnt_bad = 0, n_reg_points = 0, cnt_lam = 0, cnt_int = 0, cnt_blk = 0;
    bool fell_back = false, any_irr = false;
    // the next distortion pair of this thread: +2P inside a point, then over to the first pair of its next point
    const int n_pairs = (R + 1) >> 1;
    const long long step_inner = 2 * (long long)P, step_last = (long long)THREADS - (long long)(n_pairs - 1) * 2 * (long long)P;

    int p = p_begin + tid;
    float2 da = make_float2(0.f, 0.f), db = da;
    const float2* nx = dist + p;      // address of the pair loaded NEXT (always one pair ahead of the math)
    if (p < p_end) { da = __ldcs(nx); if (R > 1) db = __ldcs(nx + P); }
    for (; p < p_end; p += THREADS) {
        const int pn = p + THREADS;
        const bool more = pn < p_end;
        if (more) { prefetch_l2(pts + pn); prefetch_l2(nrm + pn); }
        else if (tid * kWindowSampleStride < P && src.next && src.next[0]) {
            prefetch_l2(src.next[0] + tid * kWindowSampleStride); prefetch_l2(src.next[1] + tid * kWindowSampleStride);
        }
        PointCtx pc;
        {
            float4 o4 = __ldg(pts + p), n4 = __ldg(nrm + p);
            orient_point(src, o4, n4);
            make_point(pc, T, i0, i1, i2, o4, n4);
        }
        if (!point_regular(pc)) {   // never with physical inputs: the generic loop re-traces the whole point
            any_irr = true;
            nx += THREADS;
            if (more) { da = __ldcs(nx); if (R > 1) db = __ldcs(nx + P); }
            continue;
        }
        const unsigned long long bmask = (BLK && fc.n_blk) ? block_point_mask(fc.blk, fc.n_blk, fc.bp, pc.o0, pc.o1, pc.o2, pc.r0, pc.r1, pc.r2) : 0ull;
        ++n_reg_points;
        for (int r = 0; r < R; r += 2) {
            const bool two = REVEN || (r + 1 < R);   // an odd R leaves the second lane of the last pair idle
            const float2 d0 = da, d1 = db;
            {
                const bool inner = r + 2 < R;
                nx += inner ? step_inner : step_last;
                if (inner || more) {
                    da = __ldcs(nx);
                    if (REVEN || (inner ? (r + 3 < R) : (R > 1))) db = __ldcs(nx + P);
                }
            }
            float2 cu, su, ce, se;
            bool ang0 = true, ang1 = true;
            if (TRIG == AB200_TRIG_TABLE) {
                const float4 ta = __ldg(trig + (size_t)r * P + p);
                const float4 tb = two ? __ldg(trig + (size_t)(r + 1) * P + p) : ta;
                cu = make_float2(ta.x, tb.x); su = make_float2(ta.y, tb.y); ce = make_float2(ta.z, tb.z); se = make_float2(ta.w, tb.w);
            } else {
                const float m0 = fmaxf(fabsf(d0.x), fabsf(d0.y)), m1 = fmaxf(fabsf(d1.x), fabsf(d1.y));
                if (fmaxf(m0, m1) <= kTinyAngle) {   // always, for a physical sun shape (decided per PAIR, identically
                    sincos_tiny2(make_float2(d0.x, d1.x), &su, &cu, K);   // in the forward and the backward kernel)
                    sincos_tiny2(make_float2(d0.y, d1.y), &se, &ce, K);
                } else {
                    sincos_poly_core2(make_float2(d0.x, d1.x), &su, &cu, K);
                    sincos_poly_core2(make_float2(d0.y, d1.y), &se, &ce, K);
                    ang0 = m0 <= 0.785f;
                    ang1 = m1 <= 0.785f;
                }
            }
            // scatter: d = M(e,u) r   ((-su) * r1 == su * (-r1), (-se) * r2 == se * (-r2) exactly)
            const float2 m10 = K.mul(ce, su), m11 = K.mul(ce, cu), m20 = K.mul(se, su), m21 = K.mul(se, cu);
            const float2 dx = K.add(K.mul(cu, bc2(pc.r0)), K.mul(su, bc2(-pc.r1)));
            const float2 dy = K.add(K.add(K.mul(m10, bc2(pc.r0)), K.mul(m11, bc2(pc.r1))), K.mul(se, bc2(-pc.r2)));
            const float2 dz = K.add(K.add(K.mul(m20, bc2(pc.r0)), K.mul(m21, bc2(pc.r1))), K.mul(ce, bc2(pc.r2)));
            // a = d . n_t; for a target facing exactly +-north (n_e = n_u = 0, every scenario of the reference) the two
            // zero products only add +-0, so the strict sum IS RN(dy * n_n) whenever it is non-zero (a == 0 is invalid anyway)
            const float2 a = axis_n ? K.mul(dy, bc2(T.n1))
                                    : K.add(K.add(K.mul(dx, bc2(T.n0)), K.mul(dy, bc2(T.n1))), K.mul(dz, bc2(T.n2)));
            // regular = angles in the polynomial's range and (a >= 0  or  a in (-1e18, -1e-18)); irregular rays are
            // left to the generic loop (same predicate there: angles_regular && cosine_regular)
            const bool fr0 = front_regular(a.x), fr1 = front_regular(a.y);
            const bool front0 = ang0 && fr0, front1 = two && ang1 && fr1;   // regular and front-facing: may be valid
            const float2 t = div_regular2(bc2(pc.num), a, K);
            const float2 X = K.add(bc2(pc.o0), K.mul(dx, t));
            const float2 Z = K.add(bc2(pc.o2), K.mul(dz, t));
            const float2 te = K.sub(K.add(X, bc2(T.half_w)), bc2(T.c0));
            const float2 tu = K.sub(K.add(Z, bc2(T.half_h)), bc2(T.c2));
            // exact constant-divisor quotients (const_div): q0 = te*rw; r = te - q0*w; q = q0 + r*rw.  The last
            // product adds +0 instead of -0: identical except that a -0 product becomes +0 (same validity, see above)
            const float2 qe0 = K.mul(te, bc2(T.rw)), qu0 = K.mul(tu, bc2(T.rh));
            const float2 be0 = pfma(pfma(pfma(qe0, bc2(-T.w), te), bc2(T.rw), qe0), bc2(T.em1), bc2(0.0f));
            const float2 bu0 = pfma(pfma(pfma(qu0, bc2(-T.h), tu), bc2(T.rh), qu0), bc2(T.um1), bc2(0.0f));
            const bool valid0 = front0 && (__float_as_uint(be0.x) <= em1_bits) && (__float_as_uint(bu0.x) <= um1_bits);
            const bool valid1 = front1 && (__float_as_uint(be0.y) <= em1_bits) && (__float_as_uint(bu0.y) <= um1_bits);
            float2 lam = K.mul(a, bc2(-mag));                       // mag * (-a)
            lam.x = valid0 ? lam.x : 0.0f; lam.y = valid1 ? lam.y : 0.0f;
            float2 inten;
            if (BLK) {
                float2 blocked = make_float2(0.f, 0.f);
                const bool reg0 = ang0 && (fr0 || !(a.x < 0.0f)), reg1 = two && ang1 && (fr1 || !(a.y < 0.0f));
                if (bmask) {
                    if (reg0) blocked.x = block_eval(fc.blk, bmask, fc.bp, pc.o0, pc.o1, pc.o2, dx.x, dy.x, dz.x);
                    if (reg1) blocked.y = block_eval(fc.blk, bmask, fc.bp, pc.o0, pc.o1, pc.o2, dx.y, dy.y, dz.y);
                }
                cnt_blk += (reg0 && blocked.x < 1e-3f) + (reg1 && blocked.y < 1e-3f);
                inten = K.mul(K.mul(K.mul(lam, K.sub(K.one, blocked)), bc2(ome)), bc2(refl));
                cnt_lam += (lam.x > 0.0f) + (lam.y > 0.0f);
                cnt_int += (inten.x > 0.0f) + (inten.y > 0.0f);
            } else {
                inten = K.mul(K.mul(lam, bc2(ome)), bc2(refl));
            }
            if (!(valid0 && valid1)) {   // rare
                cnt_bad += (int)!valid0 + (int)(two && !valid1);
                any_irr |= !ang0 || ((a.x < 0.0f) && !fr0) || (two && (!ang1 || ((a.y < 0.0f) && !fr1)));
            }
            float2 be = K.sub(bc2(T.em1), be0), bu = bu0;
            if (DBG) {
                const size_t q = ((size_t)h * R + r) * P + p;
                const bool reg0 = ang0 && (fr0 || !(a.x < 0.0f)), reg1 = two && ang1 && (fr1 || !(a.y < 0.0f));
                if (reg0) {
                    if (prm.a.dbg_be) prm.a.dbg_be[q] = valid0 ? be.x : T.em1;
                    if (prm.a.dbg_bu) prm.a.dbg_bu[q] = valid0 ? bu.x : 0.0f;
                    if (prm.a.dbg_t) prm.a.dbg_t[q] = valid0 ? t.x : 0.0f;
                    if (prm.a.dbg_lambert) prm.a.dbg_lambert[q] = lam.x;
                }
                if (reg1) {
                    if (prm.a.dbg_be) prm.a.dbg_be[q + P] = valid1 ? be.y : T.em1;
                    if (prm.a.dbg_bu) prm.a.dbg_bu[q + P] = valid1 ? bu.y : 0.0f;
                    if (prm.a.dbg_t) prm.a.dbg_t[q + P] = valid1 ? t.y : 0.0f;
                    if (prm.a.dbg_lambert) prm.a.dbg_lambert[q + P] = lam.y;
                }
            }
            // splat weights (valid rays have 0 <= be <= E-1, 0 <= bu <= U-1).  floor() and the integer pixel index come
            // from one round-down add of 2^23 (FMA pipe) instead of FRND + F2I (XU pipe): for 0 <= x < 2^23 the sum's
            // low mantissa bits ARE floor(x).
            const float2 me = make_float2(__fadd_rd(be.x, magic_e), __fadd_rd(be.y, magic_e));
            const float2 mu = make_float2(__fadd_rd(bu.x, magic_u), __fadd_rd(bu.y, magic_u));
            // high weights = fractional parts (exact); low weights = 1 - high, the same real number as the reference's
            // (ie + 1) - be with its single rounding placed differently (<= 1 ulp of the weight, 1e-7 of a tap)
            const float2 whe = K.sub(be, K.sub(me, bc2(magic_e))), whu = K.sub(bu, K.sub(mu, bc2(magic_u)));
            const float2 wlu = K.sub(K.one, whu);
            float2 v1, v2, v3, v4;   // tap values (fp32 accumulate) or the rounded scaled tap value as a SUBNORMAL float (fixed point)
            if (FP32ACC) {
                const float2 wle = K.sub(K.one, whe);
                v1 = K.mul(K.mul(wle, whu), inten); v2 = K.mul(K.mul(whe, whu), inten);
                v3 = K.mul(K.mul(whe, wlu), inten); v4 = K.mul(K.mul(wle, wlu), inten);
            } else {
                // round(w_e * w_u * scaled intensity) to an integer WITHOUT a float->int step: the factors carry 2^-49 and
                // 2^-100 (exact: powers of two, no underflow before a tap is far below one count), so the product of the
                // last multiplication is tap * 2^-149 - a subnormal float, whose bit pattern IS the integer, rounded to
                // nearest even by the multiplier itself (the same rounding as adding 2^23 to the unscaled product)
                float2 sc = K.mul(inten, bc2(fxs_sub));
                sc.x = fabsf(sc.x); sc.y = fabsf(sc.y);
                const float2 ahi = K.mul(whu, sc), alo = K.mul(wlu, sc);
                const float2 whes = K.mul(whe, bc2(kTwoM49)), wles = K.sub(bc2(kTwoM49), whes);
                v1 = pfma(wles, ahi, bc2(0.0f)); v2 = pfma(whes, ahi, bc2(0.0f));
                v3 = pfma(whes, alo, bc2(0.0f)); v4 = pfma(wles, alo, bc2(0.0f));
            }
#pragma unroll
            for (int lane = 0; lane < 2; ++lane) {
                if (!(lane ? valid1 : valid0)) continue;
                const int cex = (int)(__float_as_uint(lane ? me.y : me.x) - kIdxBits);   // ie - e0
                const int cux = (int)(__float_as_uint(lane ? mu.y : mu.x) - kIdxBits);   // iu - u0
                const int ie = cex + e0w, iu = cux + u0w;   // only the slow paths need the absolute pixel index
                // inside the window interior: implies ie + 1 < E and iu + 1 < U (the window lies on the bitmap)
                const bool fast = ((unsigned)cex < wwm1) && ((unsigned)cux < whm1);
                const float a1 = lane ? v1.y : v1.x, a2 = lane ? v2.y : v2.x, a3 = lane ? v3.y : v3.x, a4 = lane ? v4.y : v4.x;
                if (FP32ACC) {
                    if (fast) {
                        float* b = fc.win_f + cux * fc.ww + cex;
                        atomicAdd(b + fc.ww, a1); atomicAdd(b + fc.ww + 1, a2); atomicAdd(b + 1, a3); atomicAdd(b, a4);
                    } else if (ie + 1 < E && iu + 1 < U) {
                        float* row_hi = fc.out_f + (size_t)(U - 1 - (iu + 1)) * E + ie;
                        float* row_lo = row_hi + E;
                        atomicAdd(row_hi, a1); atomicAdd(row_hi + 1, a2); atomicAdd(row_lo + 1, a3); atomicAdd(row_lo, a4);
                    }
                } else {
                    const unsigned q1 = __float_as_uint(a1), q2 = __float_as_uint(a2), q3 = __float_as_uint(a3), q4 = __float_as_uint(a4);
                    if (fast) {
                        // 32-bit shared-window addresses + red.shared: no generic->shared conversion per tap
                        const unsigned lo = win_base + (unsigned)(cux * fc.ww + cex) * 4u, hi = lo + row_bytes;
                        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(hi), "r"(q1) : "memory");
                        asm volatile("red.shared.add.u32 [%0+4], %1;" ::"r"(hi), "r"(q2) : "memory");
                        asm volatile("red.shared.add.u32 [%0+4], %1;" ::"r"(lo), "r"(q3) : "memory");
                        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(lo), "r"(q4) : "memory");
                    } else if (ie + 1 < E && iu + 1 < U) {
                        fell_back = true;
                        atomicMin(fc.fb_box + 0, U - 2 - iu); atomicMax(fc.fb_box + 1, U - 1 - iu);
                        atomicMin(fc.fb_box + 2, ie); atomicMax(fc.fb_box + 3, ie + 1);
                        const bool e_in0 = (unsigned)cex < (unsigned)fc.ww, e_in1 = (unsigned)(cex + 1) < (unsigned)fc.ww;
                        const bool u_in0 = (unsigned)cux < (unsigned)fc.wh, u_in1 = (unsigned)(cux + 1) < (unsigned)fc.wh;
                        unsigned* g_hi = out_u + (size_t)(U - 1 - (iu + 1)) * E + ie;
                        unsigned* g_lo = g_hi + E;
                        unsigned* b = fc.win_u + cux * fc.ww + cex;
                        if (u_in1 && e_in0) atomicAdd(b + fc.ww, q1); else atomicAdd(g_hi, q1);
                        if (u_in1 && e_in1) atomicAdd(b + fc.ww + 1, q2); else atomicAdd(g_hi + 1, q2);
                        if (u_in0 && e_in1) atomicAdd(b + 1, q3); else atomicAdd(g_lo + 1, q3);
                        if (u_in0 && e_in0) atomicAdd(b, q4); else atomicAdd(g_lo, q4);
                    }
                }
            }
        }
    }
    const int cnt_valid = n_reg_points * R - cnt_bad;
    cnt_lam_out = BLK ? cnt_lam : cnt_valid; cnt_int_out = BLK ? cnt_int : cnt_valid; cnt_blk_out = cnt_blk; fell_back_out = fell_back;
    n_irregular_out = any_irr ? 1 : 0;
}

// DEFER: pass 1 of a blocking trace (BLK is false then): shadow-affected points are recorded in `dc` and skipped,
// completely shadowed ones counted without taps.  A compile-time switch so that the plain kernels carry none of it.
template <int THREADS, int TRIG, bool DBG, bool FP32ACC, bool BLK, bool REVEN, bool DEFER = false>
__device__ __forceinline__ void fwd_rays_planar_fast2(const TraceParams& prm, const TargetCtx& T, const FwdCtx& fc, const PointSrc& src, int h,
                                                      int p_begin, int p_end, float i0, float i1, float i2, int& cnt_lam_out,
                                                      int& cnt_int_out, int& cnt_blk_out, bool& fell_back_out,
                                                      int& n_irregular_out, const DeferCtx* dc = nullptr) {
    const int tid = threadIdx.x;
    const int P = prm.a.n_points, R = prm.a.n_rays, E = prm.a.res_e, U = prm.a.res_u;
    const float4* pts = src.pts;
    const float4* nrm = src.nrm;
    const float2* dist = reinterpret_cast<const float2*>(prm.a.distortions) + (size_t)h * R * P;
    const float4* trig = (TRIG == AB200_TRIG_TABLE) ? reinterpret_cast<const float4*>(prm.a.trig) + (size_t)h * R * P : nullptr;
    unsigned* out_u = reinterpret_cast<unsigned*>(fc.out_f);
    const Packed K(prm.ident);
    const float mag = prm.a.ray_magnitude, ome = prm.a.one_minus_extinction, refl = prm.a.reflectivity;
    const float fxs_sub = prm.fx_scale * 0x1p-100f;   // (exact) see the tap rounding below
    const float kTwoM49 = 0x1p-49f;
    const unsigned em1_bits = __float_as_uint(T.em1), um1_bits = __float_as_uint(T.um1);
    const bool axis_n = (T.n0 == 0.0f) && (T.n2 == 0.0f);
    // Pixel indices relative to the window come straight out of the floor() trick: be + (1.5 * 2^23 - e0), rounded down,
    // has the bit pattern 0x4B400000 + (floor(be) - e0) (two's complement wrap-around for negative differences), so
    // cex = bits - 0x4B400000 costs ONE integer instruction with an immediate operand.  The window's shared-memory
    // address is kept opaque so that it stays in a register instead of being re-derived for every tap.
    const int e0w = fc.ww > 0 ? fc.e0 : 0, u0w = fc.ww > 0 ? fc.u0 : 0;   // (an empty window has a huge origin)
    float magic_e = 12582912.0f - (float)e0w, magic_u = 12582912.0f - (float)u0w;
    asm volatile("" : "+f"(magic_e), "+f"(magic_u));   // loop invariants: kept in registers, not re-derived per pair
    const unsigned kIdxBits = 0x4B400000u;
    unsigned wwm1 = (unsigned)max(fc.ww, 1) - 1u, whm1 = (unsigned)max(fc.wh, 1) - 1u;   // empty window: nothing is "inside"
    asm volatile("" : "+r"(wwm1), "+r"(whm1));
    unsigned win_base = (unsigned)__cvta_generic_to_shared(fc.win_u);
    asm volatile("" : "+r"(win_base));
    const unsigned row_bytes = (unsigned)fc.ww * 4u;
    // factor counters: valid rays = R * (regular points) - cnt_bad; invalid rays are rare, so the bookkeeping (and the
    // irregularity flag, which implies invalid) lives in a rarely taken branch
    int cnt_bad = 0, n_reg_points = 0, cnt_lam = 0, cnt_int = 0, cnt_blk = 0, cnt_irr = 0, n_dead_points = 0, cnt_dead_bad = 0;
    bool fell_back = false, any_irr = false;
    // the next distortion pair of this thread: +2P inside a point, then over to the first pair of its next point
    const int n_pairs = (R + 1) >> 1;
    const long long step_inner = 2 * (long long)P, step_last = (long long)THREADS - (long long)(n_pairs - 1) * 2 * (long long)P;

    // BLK instantiation = pass 2 of a blocking trace: the points are the recorded (shadow-edge) ones, looked up in the
    // deferral list `dc` instead of walked round-robin; everything else is the same software pipeline
    int it = tid;
    const int n_items = BLK ? dc->count : 0;
    int p = BLK ? (it < n_items ? p_begin + defer_lookup(*dc, it) : p_end) : p_begin + tid;
    int pn = 0;
    float2 da = make_float2(0.f, 0.f), db = da;
    const float2* nx = dist + p;      // address of the pair loaded NEXT (always one pair ahead of the math)
    if (p < p_end) { da = __ldcs(nx); if (R > 1) db = __ldcs(nx + P); }
    for (; p < p_end; p = BLK ? pn : p + THREADS) {   // (plain kernels: no `pn` alive across the body)
        bool more;
        if (BLK) {
            it += THREADS;
            more = it < n_items;
            pn = more ? p_begin + defer_lookup(*dc, it) : p_end;
        } else {
            pn = p + THREADS;
            more = pn < p_end;
        }
        if (more) { prefetch_l2(pts + pn); prefetch_l2(nrm + pn); }
        else if (tid * kWindowSampleStride < P && src.next && src.next[0]) {
            prefetch_l2(src.next[0] + tid * kWindowSampleStride); prefetch_l2(src.next[1] + tid * kWindowSampleStride);
        }
        PointCtx pc;
        {
            float4 o4 = __ldg(pts + p), n4 = __ldg(nrm + p);
            orient_point(src, o4, n4);
            make_point(pc, T, i0, i1, i2, o4, n4);
        }
        if (!point_regular(pc)) {   // never with physical inputs: the generic loop re-traces the whole point
            any_irr = true;
            if (BLK) nx = dist + pn; else nx += THREADS;
            if (more) { da = __ldcs(nx); if (R > 1) db = __ldcs(nx + P); }
            continue;
        }
        bool point_dead = false;   // pass 1 of a blocking trace: every ray of the point is completely shadowed - traced for the
        if (DEFER) {               // on-target counter, no taps
            const int pcls = defer_point(dc, p - p_begin, pc.o0, pc.o1, pc.o2, pc.r0, pc.r1, pc.r2);
            if (pcls == kPointDeferred) {   // shadow-affected: pass 2
                nx += THREADS;
                if (more) { da = __ldcs(nx); if (R > 1) db = __ldcs(nx + P); }
                continue;
            }
            point_dead = pcls == kPointDead;
            n_dead_points += (int)point_dead;
        }
        const unsigned long long bmask = (BLK && fc.n_blk) ? block_point_mask(fc.blk, fc.n_blk, fc.bp, pc.o0, pc.o1, pc.o2, pc.r0, pc.r1, pc.r2) : 0ull;
        ++n_reg_points;
        for (int r = 0; r < R; r += 2) {
            const bool two = REVEN || (r + 1 < R);   // an odd R leaves the second lane of the last pair idle
            const float2 d0 = da, d1 = db;
            {
                const bool inner = r + 2 < R;
                if (BLK) nx = inner ? nx + step_inner : dist + pn; else nx += inner ? step_inner : step_last;
                if (inner || more) {
                    da = __ldcs(nx);
                    if (REVEN || (inner ? (r + 3 < R) : (R > 1))) db = __ldcs(nx + P);
                }
            }
            float2 cu, su, ce, se;
            bool ang0 = true, ang1 = true;
            if (TRIG == AB200_TRIG_TABLE) {
                const float4 ta = __ldg(trig + (size_t)r * P + p);
                const float4 tb = two ? __ldg(trig + (size_t)(r + 1) * P + p) : ta;
                cu = make_float2(ta.x, tb.x); su = make_float2(ta.y, tb.y); ce = make_float2(ta.z, tb.z); se = make_float2(ta.w, tb.w);
            } else {
                const float m0 = fmaxf(fabsf(d0.x), fabsf(d0.y)), m1 = fmaxf(fabsf(d1.x), fabsf(d1.y));
                if (fmaxf(m0, m1) <= kTinyAngle) {   // always, for a physical sun shape (decided per PAIR, identically
                    sincos_tiny2(make_float2(d0.x, d1.x), &su, &cu, K);   // in the forward and the backward kernel)
                    sincos_tiny2(make_float2(d0.y, d1.y), &se, &ce, K);
                } else {
                    sincos_poly_core2(make_float2(d0.x, d1.x), &su, &cu, K);
                    sincos_poly_core2(make_float2(d0.y, d1.y), &se, &ce, K);
                    ang0 = m0 <= 0.785f;
                    ang1 = m1 <= 0.785f;
                }
            }
            // scatter: d = M(e,u) r   ((-su) * r1 == su * (-r1), (-se) * r2 == se * (-r2) exactly)
            const float2 m10 = K.mul(ce, su), m11 = K.mul(ce, cu), m20 = K.mul(se, su), m21 = K.mul(se, cu);
            const float2 dx = K.add(K.mul(cu, bc2(pc.r0)), K.mul(su, bc2(-pc.r1)));
            const float2 dy = K.add(K.add(K.mul(m10, bc2(pc.r0)), K.mul(m11, bc2(pc.r1))), K.mul(se, bc2(-pc.r2)));
            const float2 dz = K.add(K.add(K.mul(m20, bc2(pc.r0)), K.mul(m21, bc2(pc.r1))), K.mul(ce, bc2(pc.r2)));
            // a = d . n_t; for a target facing exactly +-north (n_e = n_u = 0, every scenario of the reference) the two
            // zero products only add +-0, so the strict sum IS RN(dy * n_n) whenever it is non-zero (a == 0 is invalid anyway)
            const float2 a = axis_n ? K.mul(dy, bc2(T.n1))
                                    : K.add(K.add(K.mul(dx, bc2(T.n0)), K.mul(dy, bc2(T.n1))), K.mul(dz, bc2(T.n2)));
            // regular = angles in the polynomial's range and (a >= 0  or  a in (-1e18, -1e-18)); irregular rays are
            // left to the generic loop (same predicate there: angles_regular && cosine_regular)
            const bool fr0 = front_regular(a.x), fr1 = front_regular(a.y);
            const bool front0 = ang0 && fr0, front1 = two && ang1 && fr1;   // regular and front-facing: may be valid
            const float2 t = div_regular2(bc2(pc.num), a, K);
            const float2 X = K.add(bc2(pc.o0), K.mul(dx, t));
            const float2 Z = K.add(bc2(pc.o2), K.mul(dz, t));
            const float2 te = K.sub(K.add(X, bc2(T.half_w)), bc2(T.c0));
            const float2 tu = K.sub(K.add(Z, bc2(T.half_h)), bc2(T.c2));
            // exact constant-divisor quotients (const_div): q0 = te*rw; r = te - q0*w; q = q0 + r*rw.  The last
            // product adds +0 instead of -0: identical except that a -0 product becomes +0 (same validity, see above)
            const float2 qe0 = K.mul(te, bc2(T.rw)), qu0 = K.mul(tu, bc2(T.rh));
            const float2 be0 = pfma(pfma(pfma(qe0, bc2(-T.w), te), bc2(T.rw), qe0), bc2(T.em1), bc2(0.0f));
            const float2 bu0 = pfma(pfma(pfma(qu0, bc2(-T.h), tu), bc2(T.rh), qu0), bc2(T.um1), bc2(0.0f));
            const bool valid0 = front0 && (__float_as_uint(be0.x) <= em1_bits) && (__float_as_uint(bu0.x) <= um1_bits);
            const bool valid1 = front1 && (__float_as_uint(be0.y) <= em1_bits) && (__float_as_uint(bu0.y) <= um1_bits);
            float2 lam = K.mul(a, bc2(-mag));                       // mag * (-a)
            lam.x = valid0 ? lam.x : 0.0f; lam.y = valid1 ? lam.y : 0.0f;
            float2 inten;
            bool shadow0 = false, shadow1 = false;   // fully shadowed rays: zero intensity, no taps
            if (BLK) {
                float2 blocked = make_float2(0.f, 0.f);
                const bool reg0 = ang0 && (fr0 || !(a.x < 0.0f)), reg1 = two && ang1 && (fr1 || !(a.y < 0.0f));
                if (bmask) {
                    const int cls0 = reg0 ? block_classify(fc.blk, bmask, fc.bp, pc.o0, pc.o1, pc.o2, dx.x, dy.x, dz.x) : 0;
                    const int cls1 = reg1 ? block_classify(fc.blk, bmask, fc.bp, pc.o0, pc.o1, pc.o2, dx.y, dy.y, dz.y) : 0;
                    if (cls0) blocked.x = cls0 == 1 ? 1.0f : block_eval(fc.blk, bmask, fc.bp, pc.o0, pc.o1, pc.o2, dx.x, dy.x, dz.x);
                    if (cls1) blocked.y = cls1 == 1 ? 1.0f : block_eval(fc.blk, bmask, fc.bp, pc.o0, pc.o1, pc.o2, dx.y, dy.y, dz.y);
                    shadow0 = cls0 == 1; shadow1 = cls1 == 1;
                }
                cnt_blk += (reg0 && blocked.x < 1e-3f) + (reg1 && blocked.y < 1e-3f);
                inten = K.mul(K.mul(K.mul(lam, K.sub(K.one, blocked)), bc2(ome)), bc2(refl));
                cnt_lam += (lam.x > 0.0f) + (lam.y > 0.0f);
                cnt_int += (inten.x > 0.0f) + (inten.y > 0.0f);
            } else {
                inten = K.mul(K.mul(lam, bc2(ome)), bc2(refl));
            }
            if (!(valid0 && valid1)) {   // rare
                cnt_bad += (int)!valid0 + (int)(two && !valid1);
                const bool irr0 = !ang0 || ((a.x < 0.0f) && !fr0), irr1 = two && (!ang1 || ((a.y < 0.0f) && !fr1));
                any_irr |= irr0 || irr1;
                if (DEFER) {
                    if (!point_dead) cnt_irr += (int)irr0 + (int)irr1;
                    else cnt_dead_bad += (int)!valid0 + (int)(two && !valid1);
                }
            }
            if (DEFER && point_dead) continue;   // counted above, nothing to splat
            float2 be = K.sub(bc2(T.em1), be0), bu = bu0;
            if (DBG) {
                const size_t q = ((size_t)h * R + r) * P + p;
                const bool reg0 = ang0 && (fr0 || !(a.x < 0.0f)), reg1 = two && ang1 && (fr1 || !(a.y < 0.0f));
                if (reg0) {
                    if (prm.a.dbg_be) prm.a.dbg_be[q] = valid0 ? be.x : T.em1;
                    if (prm.a.dbg_bu) prm.a.dbg_bu[q] = valid0 ? bu.x : 0.0f;
                    if (prm.a.dbg_t) prm.a.dbg_t[q] = valid0 ? t.x : 0.0f;
                    if (prm.a.dbg_lambert) prm.a.dbg_lambert[q] = lam.x;
                }
                if (reg1) {
                    if (prm.a.dbg_be) prm.a.dbg_be[q + P] = valid1 ? be.y : T.em1;
                    if (prm.a.dbg_bu) prm.a.dbg_bu[q + P] = valid1 ? bu.y : 0.0f;
                    if (prm.a.dbg_t) prm.a.dbg_t[q + P] = valid1 ? t.y : 0.0f;
                    if (prm.a.dbg_lambert) prm.a.dbg_lambert[q + P] = lam.y;
                }
            }
            // splat weights (valid rays have 0 <= be <= E-1, 0 <= bu <= U-1).  floor() and the integer pixel index come
            // from one round-down add of 2^23 (FMA pipe) instead of FRND + F2I (XU pipe): for 0 <= x < 2^23 the sum's
            // low mantissa bits ARE floor(x).
            const float2 me = make_float2(__fadd_rd(be.x, magic_e), __fadd_rd(be.y, magic_e));
            const float2 mu = make_float2(__fadd_rd(bu.x, magic_u), __fadd_rd(bu.y, magic_u));
            // high weights = fractional parts (exact); low weights = 1 - high, the same real number as the reference's
            // (ie + 1) - be with its single rounding placed differently (<= 1 ulp of the weight, 1e-7 of a tap)
            const float2 whe = K.sub(be, K.sub(me, bc2(magic_e))), whu = K.sub(bu, K.sub(mu, bc2(magic_u)));
            const float2 wlu = K.sub(K.one, whu);
            float2 v1, v2, v3, v4;   // tap values (fp32 accumulate) or the rounded scaled tap value as a SUBNORMAL float (fixed point)
            if (FP32ACC) {
                const float2 wle = K.sub(K.one, whe);
                v1 = K.mul(K.mul(wle, whu), inten); v2 = K.mul(K.mul(whe, whu), inten);
                v3 = K.mul(K.mul(whe, wlu), inten); v4 = K.mul(K.mul(wle, wlu), inten);
            } else {
                // round(w_e * w_u * scaled intensity) to an integer WITHOUT a float->int step: the factors carry 2^-49 and
                // 2^-100 (exact: powers of two, no underflow before a tap is far below one count), so the product of the
                // last multiplication is tap * 2^-149 - a subnormal float, whose bit pattern IS the integer, rounded to
                // nearest even by the multiplier itself (the same rounding as adding 2^23 to the unscaled product)
                float2 sc = K.mul(inten, bc2(fxs_sub));
                sc.x = fabsf(sc.x); sc.y = fabsf(sc.y);
                const float2 ahi = K.mul(whu, sc), alo = K.mul(wlu, sc);
                const float2 whes = K.mul(whe, bc2(kTwoM49)), wles = K.sub(bc2(kTwoM49), whes);
                v1 = pfma(wles, ahi, bc2(0.0f)); v2 = pfma(whes, ahi, bc2(0.0f));
                v3 = pfma(whes, alo, bc2(0.0f)); v4 = pfma(wles, alo, bc2(0.0f));
            }
#pragma unroll
            for (int lane = 0; lane < 2; ++lane) {
                if (!(lane ? valid1 : valid0)) continue;
                if (BLK && (lane ? shadow1 : shadow0)) continue;
                const int cex = (int)(__float_as_uint(lane ? me.y : me.x) - kIdxBits);   // ie - e0
                const int cux = (int)(__float_as_uint(lane ? mu.y : mu.x) - kIdxBits);   // iu - u0
                const int ie = cex + e0w, iu = cux + u0w;   // only the slow paths need the absolute pixel index
                // inside the window interior: implies ie + 1 < E and iu + 1 < U (the window lies on the bitmap)
                const bool fast = ((unsigned)cex < wwm1) && ((unsigned)cux < whm1);
                const float a1 = lane ? v1.y : v1.x, a2 = lane ? v2.y : v2.x, a3 = lane ? v3.y : v3.x, a4 = lane ? v4.y : v4.x;
                if (FP32ACC) {
                    if (fast) {
                        float* b = fc.win_f + cux * fc.ww + cex;
                        atomicAdd(b + fc.ww, a1); atomicAdd(b + fc.ww + 1, a2); atomicAdd(b + 1, a3); atomicAdd(b, a4);
                    } else if (ie + 1 < E && iu + 1 < U) {
                        float* row_hi = fc.out_f + (size_t)(U - 1 - (iu + 1)) * E + ie;
                        float* row_lo = row_hi + E;
                        atomicAdd(row_hi, a1); atomicAdd(row_hi + 1, a2); atomicAdd(row_lo + 1, a3); atomicAdd(row_lo, a4);
                    }
                } else {
                    const unsigned q1 = __float_as_uint(a1), q2 = __float_as_uint(a2), q3 = __float_as_uint(a3), q4 = __float_as_uint(a4);
                    if (fast) {
                        // 32-bit shared-window addresses + red.shared: no generic->shared conversion per tap
                        const unsigned lo = win_base + (unsigned)(cux * fc.ww + cex) * 4u, hi = lo + row_bytes;
                        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(hi), "r"(q1) : "memory");
                        asm volatile("red.shared.add.u32 [%0+4], %1;" ::"r"(hi), "r"(q2) : "memory");
                        asm volatile("red.shared.add.u32 [%0+4], %1;" ::"r"(lo), "r"(q3) : "memory");
                        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(lo), "r"(q4) : "memory");
                    } else if (ie + 1 < E && iu + 1 < U) {
                        fell_back = true;
                        atomicMin(fc.fb_box + 0, U - 2 - iu); atomicMax(fc.fb_box + 1, U - 1 - iu);
                        atomicMin(fc.fb_box + 2, ie); atomicMax(fc.fb_box + 3, ie + 1);
                        const bool e_in0 = (unsigned)cex < (unsigned)fc.ww, e_in1 = (unsigned)(cex + 1) < (unsigned)fc.ww;
                        const bool u_in0 = (unsigned)cux < (unsigned)fc.wh, u_in1 = (unsigned)(cux + 1) < (unsigned)fc.wh;
                        unsigned* g_hi = out_u + (size_t)(U - 1 - (iu + 1)) * E + ie;
                        unsigned* g_lo = g_hi + E;
                        unsigned* b = fc.win_u + cux * fc.ww + cex;
                        if (u_in1 && e_in0) atomicAdd(b + fc.ww, q1); else atomicAdd(g_hi, q1);
                        if (u_in1 && e_in1) atomicAdd(b + fc.ww + 1, q2); else atomicAdd(g_hi + 1, q2);
                        if (u_in0 && e_in1) atomicAdd(b + 1, q3); else atomicAdd(g_lo + 1, q3);
                        if (u_in0 && e_in0) atomicAdd(b, q4); else atomicAdd(g_lo, q4);
                    }
                }
            }
        }
    }
    const int cnt_valid = n_reg_points * R - cnt_bad;
    // (dead points of pass 1: their valid rays are on target but carry no intensity and count as blocked)
    cnt_lam_out = BLK ? cnt_lam : cnt_valid; cnt_int_out = BLK ? cnt_int : (DEFER ? cnt_valid - (n_dead_points * R - cnt_dead_bad) : cnt_valid);
    fell_back_out = fell_back;
    // pass 1 of a blocking trace: every regular ray traced here is unshadowed (blocked == 0 < 1e-3); the irregular ones are
    // counted by the generic pass that re-traces them
    cnt_blk_out = DEFER ? (n_reg_points - n_dead_points) * R - cnt_irr : cnt_blk;
    n_irregular_out = any_irr ? 1 : 0;
}

template <int THREADS, int TRIG, bool DBG, bool FP32ACC, bool BLK, bool MAP = false>
__global__ void __launch_bounds__(THREADS, (THREADS > 512 ? 1 : 2))
trace_fwd_kernel(const TraceParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned* win_u = reinterpret_cast<unsigned*>(smem_raw);
    float* win_f = reinterpret_cast<float*>(smem_raw);
    __shared__ TargetCtx T_sh;
    __shared__ Window win_sh;
    __shared__ float red[6 * 32];
    __shared__ int cnt_sh[3];
    __shared__ int fallback_sh;
    __shared__ int fb_box[4];
    __shared__ BlockPrim blk_sh[BLK ? kMaxBlockCandidates : 1];
    __shared__ unsigned defer_bits_sh[BLK ? kDeferWords : 1];
    __shared__ int defer_prefix_sh[BLK ? kDeferWords : 1];
    __shared__ int defer_total_sh;
    __shared__ float O_sh[16];
    __shared__ const float4* next_sh[2];

    const int tid = threadIdx.x;
    const int li = blockIdx.x / prm.split;
    const int chunk = blockIdx.x - li * prm.split;
    const int h = prm.a.local_rows ? min(max(prm.a.local_rows[li], 0), prm.a.n_samples - 1) : li;   // (clamped: memory safety)
    const int P = prm.a.n_points, R = prm.a.n_rays, E = prm.a.res_e, U = prm.a.res_u;
    const int p_begin = chunk * prm.pts_per_chunk;
    const int p_end = min(P, p_begin + prm.pts_per_chunk);

    // start-up loads, all issued before the first barrier so that their latencies overlap: target constants (thread
    // 0), orientation (threads 32..47), incident direction and the window-sample rows (every thread); meanwhile the
    // whole shared-memory window is cleared
    long long t_phase = (prm.a.stats && tid == 0) ? clock64() : 0;
    const int hs = MAP ? __ldg(prm.a.src_rows + h) : h;   // activation index map: surface row of this sample
    const float4* pts_h = reinterpret_cast<const float4*>(prm.a.points) + (size_t)hs * P;
    const float4* nrm_h = reinterpret_cast<const float4*>(prm.a.normals) + (size_t)hs * P;
    WindowSamples ws;
    load_window_samples<THREADS>(ws, pts_h, nrm_h, p_begin, p_end);
    const float i0 = __ldg(prm.a.incident + 4 * h), i1 = __ldg(prm.a.incident + 4 * h + 1),
                i2 = __ldg(prm.a.incident + 4 * h + 2);
    if (tid == 0) {
        load_target(T_sh, prm.a.targets, prm.a.target_idx[h], E, U);
        cnt_sh[0] = 0; cnt_sh[1] = 0; cnt_sh[2] = 0; fallback_sh = 0;
        // output rows / columns touched by global-path taps: row min, row max, col min, col max
        fb_box[0] = 1 << 30; fb_box[1] = -1; fb_box[2] = 1 << 30; fb_box[3] = -1;
    }
    if (prm.a.orientations && tid >= 32 && tid < 48) O_sh[tid - 32] = __ldg(prm.a.orientations + (size_t)h * 16 + (tid - 32));
    if (tid == 64) set_next_sample<MAP>(next_sh, prm, li);
    {
        uint4* w4 = reinterpret_cast<uint4*>(smem_raw);
        const int n4 = prm.win_cap >> 2;
        for (int i = tid; i < n4; i += THREADS) w4[i] = make_uint4(0u, 0u, 0u, 0u);
        for (int i = (n4 << 2) + tid; i < prm.win_cap; i += THREADS) win_u[i] = 0u;
    }
    __syncthreads();
    AB200_PHASE(4, 0);   // start-up loads + window clear
    PointSrc src;
    src.pts = pts_h;
    src.nrm = nrm_h;
    src.O = prm.a.orientations ? O_sh : nullptr;
    src.next = next_sh;
    const TargetCtx T = T_sh;

    Window W;
    place_window<THREADS>(W, prm, T, src, ws, p_begin, p_end, i0, i1, i2, red, &win_sh);
    if (prm.a.windows && prm.split == 1 && tid == 0)   // for the backward of the same call
        reinterpret_cast<int4*>(prm.a.windows)[h] = make_int4(W.e0, W.u0, W.ww, W.wh);
    AB200_PHASE(4, 1);   // window placement
    float* out_f = prm.a.flux + (size_t)h * U * E;
    unsigned* out_u = reinterpret_cast<unsigned*>(out_f);
    if (prm.self_zero && prm.quad) {
        // clear the quads the window flush will not overwrite (the global-path taps need zeros to add to): a handful
        // of 16-byte stores per thread (store bandwidth of an SM is ~32 B/clk: clearing the whole row costs 8 K cycles)
        const OutsideQuads oq(W, E, U);
        float4* o4 = reinterpret_cast<float4*>(out_f);
        for (int i = tid; i < oq.n_out; i += THREADS) o4[oq.at(i)] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else if (prm.self_zero) {
        // clear the pixels the window flush will not overwrite (the global-path taps need zeros to add to); output
        // rows of the window: [U - u0 - wh, U - 1 - u0], columns [e0, e0 + ww)
        const int r_lo = U - W.u0 - W.wh, r_hi = U - 1 - W.u0, c_lo = W.e0, c_hi = W.e0 + W.ww;
        const int warp_z = tid >> 5, lane_z = tid & 31;
        if ((E & 3) == 0) {
            const int e4 = E >> 2;
            for (int row = warp_z; row < U; row += THREADS / 32) {
                float4* o4 = reinterpret_cast<float4*>(out_f + (size_t)row * E);
                const bool row_in = row >= r_lo && row <= r_hi;
                for (int q = lane_z; q < e4; q += 32) {
                    const int c = q << 2;
                    if (!row_in || c + 4 <= c_lo || c >= c_hi) o4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                    else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) if (c + k < c_lo || c + k >= c_hi) out_f[(size_t)row * E + c + k] = 0.f;
                    }
                }
            }
        } else {
            for (int row = warp_z; row < U; row += THREADS / 32) {
                const bool row_in = row >= r_lo && row <= r_hi;
                for (int c = lane_z; c < E; c += 32)
                    if (!row_in || c < c_lo || c >= c_hi) out_f[(size_t)row * E + c] = 0.f;
            }
        }
    }
    __syncthreads();
    AB200_PHASE(4, 2);   // clearing the bitmap row outside the window

    FwdCtx fc;
    fc.win_u = win_u; fc.win_f = win_f; fc.out_f = out_f;
    fc.e0 = W.e0; fc.u0 = W.u0; fc.ww = W.ww; fc.wh = W.wh;
    fc.wwm1 = W.ww > 1 ? W.ww - 1 : 0; fc.whm1 = W.wh > 1 ? W.wh - 1 : 0;
    fc.fb_box = fb_box;
    fc.blk = blk_sh;
    fc.n_blk = 0;
    if (BLK && prm.a.blockers.n_blockers > 0) {
        const ab200_blockers& B = prm.a.blockers;
        fc.bp.softness = B.softness; fc.bp.alpha = B.alpha; fc.bp.offset = B.ray_origin_offset; fc.bp.epsilon = B.epsilon;
        fc.bp.cull_angle = B.cull_angle;
        fc.n_blk = block_load_candidates<THREADS>(blk_sh, nullptr, B, h, tid);
    }

    int cnt_lam = 0, cnt_int = 0, cnt_blk = 0;
    bool fell_back = false;
    // Blocking: two passes (blocking_device.cuh, "Deferral") - the loops below run WITHOUT blocking code over the points no
    // candidate can shadow and record the others, which the generic loop then traces with the soft mask, spread evenly
    // over the CTA.  Chunks beyond the bit set's capacity evaluate the mask inline in the generic loops.
    const bool blocking = BLK && fc.n_blk > 0;
    const bool defer = BLK && (p_end - p_begin) <= kDeferWords * 32;
    DeferCtx dc;
    dc.bits = defer_bits_sh; dc.prefix = defer_prefix_sh; dc.n_words = (p_end - p_begin + 31) >> 5; dc.count = 0;
    dc.blk = fc.blk; dc.n_blk = fc.n_blk; dc.bp = fc.bp;
    if (BLK && defer) {
        for (int w = tid; w < dc.n_words; w += THREADS) defer_bits_sh[w] = 0u;
        __syncthreads();
    }
    const DeferCtx* rec = (BLK && defer) ? &dc : nullptr;
    const int rec_mode = (BLK && defer) ? kDeferRecord : kDeferOff;
    // (fx_scale >= 1e-6: the packed loop's subnormal tap rounding scales it by 2^-100 and needs the result to stay normal)
    if (BLK && !defer) {
        if (T.planar)
            fwd_rays<THREADS, TRIG, DBG, FP32ACC, true, false, false, BLK>(prm, T, fc, src, h, p_begin, p_end, i0, i1, i2, cnt_lam, cnt_int, cnt_blk, fell_back);
        else
            fwd_rays<THREADS, TRIG, DBG, FP32ACC, false, false, false, BLK>(prm, T, fc, src, h, p_begin, p_end, i0, i1, i2, cnt_lam, cnt_int, cnt_blk, fell_back);
    } else if (T.planar && T.fastdiv && TRIG != AB200_TRIG_SINCOSF && prm.simple_counts && (FP32ACC || prm.fx_scale >= 1e-6f)) {
        int n_irr = 0;
#if AB200_PACKED_RAYS
        if ((R & 1) == 0)
            if (BLK) fwd_rays_planar_fast2<THREADS, TRIG, DBG, FP32ACC, false, true, true>(prm, T, fc, src, h, p_begin, p_end, i0, i1, i2, cnt_lam, cnt_int, cnt_blk, fell_back, n_irr, rec);
            else fwd_rays_planar_fast2_plain<THREADS, TRIG, DBG, FP32ACC, false, true>(prm, T, fc, src, h, p_begin, p_end, i0, i1, i2, cnt_lam, cnt_int, cnt_blk, fell_back, n_irr);
        else
            if (BLK) fwd_rays_planar_fast2<THREADS, TRIG, DBG, FP32ACC, false, false, true>(prm, T, fc, src, h, p_begin, p_end, i0, i1, i2, cnt_lam, cnt_int, cnt_blk, fell_back, n_irr, rec);
            else fwd_rays_planar_fast2_plain<THREADS, TRIG, DBG, FP32ACC, false, false>(prm, T, fc, src, h, p_begin, p_end, i0, i1, i2, cnt_lam, cnt_int, cnt_blk, fell_back, n_irr);
#else   // tuning build without the packed loops: the scalar fast loop evaluates the mask inline, nothing is deferred
        fwd_rays_planar_fast<THREADS, TRIG, DBG, FP32ACC, BLK>(prm, T, fc, src, h, p_begin, p_end, i0, i1, i2, cnt_lam, cnt_int, cnt_blk, fell_back, n_irr);
#endif
        if (__syncthreads_or(n_irr != 0)) {   // never with physical inputs: the generic loop picks up the skipped rays
            int c1 = 0, c2 = 0, c3 = 0;
            bool fb = false;
            const bool skip_masked = AB200_PACKED_RAYS && rec != nullptr;
            fwd_rays<THREADS, TRIG, DBG, FP32ACC, true, false, true, BLK>(prm, T, fc, src, h, p_begin, p_end, i0, i1, i2, c1, c2, c3, fb,
                                                                         rec, skip_masked ? kDeferSkip : kDeferOff);
            cnt_lam += c1; cnt_int += c2; cnt_blk += c3; fell_back = fell_back || fb;
        }
    } else if (T.planar) {
        fwd_rays<THREADS, TRIG, DBG, FP32ACC, true, false, false, false>(prm, T, fc, src, h, p_begin, p_end, i0, i1, i2, cnt_lam, cnt_int, cnt_blk, fell_back, rec, rec_mode);
    } else {
        fwd_rays<THREADS, TRIG, DBG, FP32ACC, false, false, false, false>(prm, T, fc, src, h, p_begin, p_end, i0, i1, i2, cnt_lam, cnt_int, cnt_blk, fell_back, rec, rec_mode);
    }
    if (BLK && defer && blocking) {   // pass 2 (uniform over the CTA)
        AB200_PHASE(4, 3);   // ray loop, pass 1
        if (defer_scan<THREADS>(dc, &defer_total_sh) > 0) {
            if (prm.a.stats && tid == 0) atomicAdd(reinterpret_cast<unsigned long long*>(prm.a.stats) + 17, (unsigned long long)dc.count);
            int c1 = 0, c2 = 0, c3 = 0;
            bool fb = false;
#if AB200_PACKED_RAYS
            if (T.planar && T.fastdiv && TRIG != AB200_TRIG_SINCOSF && (FP32ACC || prm.fx_scale >= 1e-6f)) {
                // the packed loop with the per-ray classification inline, over the recorded points only
                int n_irr = 0;
                if ((R & 1) == 0)
                    fwd_rays_planar_fast2<THREADS, TRIG, DBG, FP32ACC, BLK, true>(prm, T, fc, src, h, p_begin, p_end, i0, i1, i2, c1, c2, c3, fb, n_irr, &dc);
                else
                    fwd_rays_planar_fast2<THREADS, TRIG, DBG, FP32ACC, BLK, false>(prm, T, fc, src, h, p_begin, p_end, i0, i1, i2, c1, c2, c3, fb, n_irr, &dc);
                if (__syncthreads_or(n_irr != 0)) {   // never with physical inputs
                    int d1 = 0, d2 = 0, d3 = 0;
                    bool fb2 = false;
                    fwd_rays<THREADS, TRIG, DBG, FP32ACC, true, false, true, BLK>(prm, T, fc, src, h, p_begin, p_end, i0, i1, i2, d1, d2, d3, fb2, &dc, kDeferList);
                    c1 += d1; c2 += d2; c3 += d3; fb = fb || fb2;
                }
            } else
#endif
            if (T.planar)
                fwd_rays<THREADS, TRIG, DBG, FP32ACC, true, false, false, BLK>(prm, T, fc, src, h, p_begin, p_end, i0, i1, i2, c1, c2, c3, fb, &dc, kDeferList);
            else
                fwd_rays<THREADS, TRIG, DBG, FP32ACC, false, false, false, BLK>(prm, T, fc, src, h, p_begin, p_end, i0, i1, i2, c1, c2, c3, fb, &dc, kDeferList);
            cnt_lam += c1; cnt_int += c2; cnt_blk += c3; fell_back = fell_back || fb;
        }
        AB200_PHASE(14, 4);   // stats[18]: pass 2 (deferred points)
    }

    AB200_PHASE(4, 3);   // ray loop (thread 0)
    // ---- epilogue: counters, window flush -----------------------------------------------------
    cnt_lam = warp_sum(cnt_lam);
    cnt_int = warp_sum(cnt_int);
    cnt_blk = warp_sum(cnt_blk);
    if ((tid & 31) == 0) {
        if (cnt_lam) atomicAdd(&cnt_sh[0], cnt_lam);
        if (cnt_int) atomicAdd(&cnt_sh[1], cnt_int);
        if (cnt_blk) atomicAdd(&cnt_sh[2], cnt_blk);
    }
    if (!FP32ACC && fell_back) fallback_sh = 1;
    if (prm.a.stats && tid == 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(prm.a.stats) + 1, (unsigned long long)(W.ww * W.wh));
        atomicAdd(reinterpret_cast<unsigned long long*>(prm.a.stats) + 2, 1ull);
    }
    if (prm.a.stats && fell_back) atomicAdd(reinterpret_cast<unsigned long long*>(prm.a.stats) + 0, 1ull);
    if (!FP32ACC) __threadfence();  // integer REDs of this thread are performed before the barrier
    __syncthreads();
    AB200_PHASE(4, 4);   // waiting for the CTA's last warp

    const bool single = (prm.split == 1);
    if (tid == 0) {
        if (single) {
            const float rp = (float)(R * P);
            prm.a.on_target[h] = sdiv((float)cnt_sh[0], rp);
            prm.a.intercept[h] = sdiv((float)cnt_sh[1], rp);
            prm.a.blocking[h] = BLK ? sdiv((float)cnt_sh[2], rp) : 1.0f;  // (blocked < 1e-3).sum() / (R*P)
        } else {
            atomicAdd(reinterpret_cast<int*>(prm.a.on_target) + h, cnt_sh[0]);
            atomicAdd(reinterpret_cast<int*>(prm.a.intercept) + h, cnt_sh[1]);
            atomicAdd(reinterpret_cast<int*>(prm.a.blocking) + h, BLK ? cnt_sh[2] : (p_end - p_begin) * R);
        }
    }

    const int warp = tid >> 5, lane = tid & 31, nwarps = THREADS / 32;
    if (FP32ACC) {
        for (int r = warp; r < W.wh; r += nwarps) {
            float* orow = out_f + (size_t)(U - 1 - (W.u0 + r)) * E + W.e0;
            for (int c = lane; c < W.ww; c += 32) {
                const float v = win_f[r * W.ww + c];
                if (v != 0.0f) atomicAdd(orow + c, v);
            }
        }
        return;
    }
    if (single && prm.quad && W.ww > 0) {
        const float inv = prm.fx_inv;
        const bool any_fb = fallback_sh != 0;
        {   // window -> output rows, one quad (uint4 -> float4) per thread and iteration
            const int qw = W.ww >> 2, n4 = W.wh * qw;
            const int d_r = THREADS / qw, d_q = THREADS - d_r * qw;
            const uint4* w4 = reinterpret_cast<const uint4*>(win_u);
            int r = tid / qw, q = tid - r * qw;
            for (int idx = tid; idx < n4; idx += THREADS) {
                const uint4 v = w4[idx];
                float4* orow = reinterpret_cast<float4*>(out_f + (size_t)(U - 1 - (W.u0 + r)) * E + W.e0);
                const float4 o4v = make_float4(__uint2float_rn(v.x) * inv, __uint2float_rn(v.y) * inv, __uint2float_rn(v.z) * inv,
                                               __uint2float_rn(v.w) * inv);
#if AB200_FLUSH_ST == 1
                orow[q] = o4v;
#elif AB200_FLUSH_ST == 2
                __stcg(orow + q, o4v);
#elif AB200_FLUSH_ST == 3
                __stwt(orow + q, o4v);
#else
                __stcs(orow + q, o4v);
#endif
                r += d_r; q += d_q;
                if (q >= qw) { q -= qw; ++r; }
            }
        }
        AB200_PHASE(4, 5);   // window flush (thread 0)
        if (any_fb) {
            // integer taps that landed outside the window (on the cleared part of the row): convert in place.  Only the
            // out-of-window quads are enumerated: the full rows below / above the window, then the strips left and right
            // of it - a handful per thread, all of whose L2 reads are in flight together
            const OutsideQuads oq(W, E, U);
            uint4* o4 = reinterpret_cast<uint4*>(out_f);
            constexpr int kBatch = 4;
            for (int idx0 = tid; idx0 < oq.n_out; idx0 += kBatch * THREADS) {
                uint4 v[kBatch];
                int at[kBatch];
#pragma unroll
                for (int k = 0; k < kBatch; ++k) {
                    const int i = idx0 + k * THREADS;
                    v[k] = make_uint4(0u, 0u, 0u, 0u);
                    at[k] = 0;
                    if (i < oq.n_out) { at[k] = oq.at(i); v[k] = __ldcg(o4 + at[k]); }
                }
#pragma unroll
                for (int k = 0; k < kBatch; ++k)
                    if (v[k].x | v[k].y | v[k].z | v[k].w)
                        reinterpret_cast<float4*>(o4)[at[k]] =
                            make_float4(__uint2float_rn(v[k].x) * inv, __uint2float_rn(v[k].y) * inv,
                                        __uint2float_rn(v[k].z) * inv, __uint2float_rn(v[k].w) * inv);
            }
        }
        AB200_PHASE(4, 6);   // conversion of the out-of-window taps (thread 0)
    } else if (single) {
        const float inv = prm.fx_inv;
        const bool any_fb = fallback_sh != 0;
        for (int r = warp; r < W.wh; r += nwarps) {
            float* orow = out_f + (size_t)(U - 1 - (W.u0 + r)) * E + W.e0;
            for (int c = lane; c < W.ww; c += 32) orow[c] = __uint2float_rn(win_u[r * W.ww + c]) * inv;
        }
        if (any_fb) {
            // rare: convert the integer taps that landed outside the window, in place
            const int r_lo = max(fb_box[0], 0), r_hi = min(fb_box[1], U - 1), c_lo = max(fb_box[2], 0), c_hi = min(fb_box[3], E - 1);
            const int wr_lo = U - W.u0 - W.wh, wr_hi = U - 1 - W.u0;     // output rows of the window
            constexpr int kRows = 4;                                     // independent L2 reads per thread and iteration
            for (int row0 = r_lo + warp * kRows; row0 <= r_hi; row0 += nwarps * kRows) {
                for (int c = c_lo + lane; c <= c_hi; c += 32) {
                    const bool col_in = c >= W.e0 && c < W.e0 + W.ww;
                    unsigned q[kRows];
#pragma unroll
                    for (int k = 0; k < kRows; ++k) {
                        const int row = row0 + k;
                        q[k] = 0u;
                        if (row <= r_hi && !(col_in && row >= wr_lo && row <= wr_hi)) q[k] = __ldcg(out_u + (size_t)row * E + c);
                    }
#pragma unroll
                    for (int k = 0; k < kRows; ++k)
                        if (q[k]) out_f[(size_t)(row0 + k) * E + c] = __uint2float_rn(q[k]) * inv;
                }
            }
        }
        AB200_PHASE(4, 5);   // window flush (thread 0)
    } else {
        for (int r = warp; r < W.wh; r += nwarps) {
            unsigned* orow = out_u + (size_t)(U - 1 - (W.u0 + r)) * E + W.e0;
            for (int c = lane; c < W.ww; c += 32) {
                const unsigned q = win_u[r * W.ww + c];
                if (q) atomicAdd(orow + c, q);
            }
        }
    }
}

#ifndef AB200_TU_BWD   // (forward half only)
// split mode only: integer bitmap / counters -> fp32, in place
__global__ void finalize_split_kernel(const TraceParams prm) {
    const int li = blockIdx.y;
    const int h = prm.a.local_rows ? prm.a.local_rows[li] : li;
    const int UE = prm.a.res_u * prm.a.res_e;
    float* out_f = prm.a.flux + (size_t)h * UE;
    unsigned* out_u = reinterpret_cast<unsigned*>(out_f);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < UE; i += gridDim.x * blockDim.x) {
        const unsigned q = out_u[i];
        if (q) out_f[i] = __uint2float_rn(q) * prm.fx_inv;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const float rp = (float)(prm.a.n_rays * prm.a.n_points);
        const int c0 = reinterpret_cast<int*>(prm.a.on_target)[h];
        const int c1 = reinterpret_cast<int*>(prm.a.intercept)[h];
        const int c2 = reinterpret_cast<int*>(prm.a.blocking)[h];
        prm.a.on_target[h] = sdiv((float)c0, rp);
        prm.a.intercept[h] = sdiv((float)c1, rp);
        prm.a.blocking[h] = sdiv((float)c2, rp);
    }
}

__global__ void finalize_split_fp32_kernel(const TraceParams prm) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= prm.a.n_local) return;
    const int h = prm.a.local_rows ? prm.a.local_rows[li] : li;
    const float rp = (float)(prm.a.n_rays * prm.a.n_points);
    const int c0 = reinterpret_cast<int*>(prm.a.on_target)[h];
    const int c1 = reinterpret_cast<int*>(prm.a.intercept)[h];
    const int c2 = reinterpret_cast<int*>(prm.a.blocking)[h];
    prm.a.on_target[h] = sdiv((float)c0, rp);
    prm.a.intercept[h] = sdiv((float)c1, rp);
    prm.a.blocking[h] = sdiv((float)c2, rp);
}
#endif  // !AB200_TU_BWD

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
struct BwdCtx {
    const float* win_g;  // shared-memory window of the bitmap gradient
    const float* gf;     // this sample's [U,E] gradient in global memory
    int e0, u0, ww, wwm1, whm1;
    const BlockPrim* blk; // shared: candidate blocking primitives of this sample
    const int* blk_rows;  // shared: their rows in the primitive table
    int n_blk;
    BlockParams bp;
    double* gacc;         // shared: the CTA's blocker-gradient accumulators per candidate slot (NULL: not wanted)
    int gacc_copies, gacc_stride;   // ... as `gacc_copies` (a power of two) copies, `gacc_stride` doubles apart
};

template <int THREADS, int TRIG, bool PLANAR, bool FASTDIV, bool ONLY_IRREGULAR, bool BLK>
__device__ __forceinline__ void bwd_rays(const TraceParams& prm, const TargetCtx& T, const BwdCtx& bc, const PointSrc& src, int h,
                                         int p_begin, int p_end, float i0, float i1, float i2, float* __restrict__ grad_points,
                                         float* __restrict__ grad_normals, float* gori_acc, const DeferCtx* dc = nullptr,
                                         int dmode = kDeferOff) {
    const int tid = threadIdx.x;
    const int P = prm.a.n_points, R = prm.a.n_rays, E = prm.a.res_e, U = prm.a.res_u;
    const float4* pts = src.pts;
    const float4* nrm = src.nrm;
    const float2* dist = reinterpret_cast<const float2*>(prm.a.distortions) + (size_t)h * R * P;
    const float4* trig = (TRIG == AB200_TRIG_TABLE) ? reinterpret_cast<const float4*>(prm.a.trig) + (size_t)h * R * P : nullptr;
    const float mag = prm.a.ray_magnitude;
    const float k_or = prm.a.one_minus_extinction * prm.a.reflectivity;
    const float k_int = mag * k_or;                 // d intensity / d lambert-cosine
    const float k_e = T.em1 / T.w, k_u = T.um1 / T.h;

    // dmode: see fwd_rays
    const int n_items = dmode == kDeferList ? dc->count : p_end - p_begin;
    int p_list = (dmode == kDeferList && tid < n_items) ? p_begin + defer_lookup(*dc, tid) : 0;
    for (int it = tid; it < n_items; it += THREADS) {
        const int p = dmode == kDeferList ? p_list : p_begin + it;
        if (dmode == kDeferList && it + THREADS < n_items) {   // scattered points: prefetch the next one's data (see fwd_rays)
            p_list = p_begin + defer_lookup(*dc, it + THREADS);
            prefetch_l2(pts + p_list); prefetch_l2(nrm + p_list);
            for (int r = 0; r < R; ++r) prefetch_l2(dist + (size_t)r * P + p_list);
        }
        const float4 o_raw = __ldg(pts + p), n_raw = __ldg(nrm + p);
        float4 o4 = o_raw, n4 = n_raw;
        orient_point(src, o4, n4);
        PointCtx pc;
        make_point(pc, T, i0, i1, i2, o4, n4);
        bool point_dead = false;   // pass 1: every ray completely shadowed - zero gradient rows
        if (!BLK && dmode == kDeferRecord) {
            const int pcls = defer_point(dc, p - p_begin, pc.o0, pc.o1, pc.o2, pc.r0, pc.r1, pc.r2);
            if (pcls == kPointDeferred) continue;
            point_dead = pcls == kPointDead;
        }
        float go0 = 0.f, go1 = 0.f, go2 = 0.f;   // planar: grad origin (world); cylindrical: grad origin (cylinder frame)
        float gr0 = 0.f, gr1 = 0.f, gr2 = 0.f;   // grad preferred reflection direction
        float gow0 = 0.f, gow1 = 0.f, gow2 = 0.f;  // grad origin through the blocking term (always world frame)
        bool touched = false;
        const unsigned long long bmask = (BLK && bc.n_blk) ? block_point_mask(bc.blk, bc.n_blk, bc.bp, pc.o0, pc.o1, pc.o2, pc.r0, pc.r1, pc.r2) : 0ull;
        if (BLK && dmode == kDeferSkip && bmask) continue;
        const float2* dp = dist + p;
        float2 d_next = __ldcs(dp);
        for (int r = 0; r < (point_dead ? 0 : R); ++r) {
            const float2 d = d_next;
            dp += P;
            if (r + 1 < R) d_next = __ldcs(dp);
            Scatter s;
            ray_trig<TRIG>(d.x, d.y, trig, (size_t)r * P + p, s.cu, s.su, s.ce, s.se);
            scatter(s, pc);
            int bcls = 0;   // 1: fully shadowed - no intensity, no gradient (most rays of a deferred point); 2: inside a
            if (BLK && bmask) {   // sigmoid transition
                bcls = block_classify(bc.blk, bmask, bc.bp, pc.o0, pc.o1, pc.o2, s.dx, s.dy, s.dz);
                if (bcls == 1) continue;
            }
            Hit hit;
            if (PLANAR) hit_planar<FASTDIV>(hit, T, pc, s, mag); else hit_cylinder<FASTDIV, TRIG == AB200_TRIG_TABLE>(hit, T, pc, s, mag);
            if (ONLY_IRREGULAR && point_regular(pc) && angles_regular<TRIG>(d.x, d.y) && cosine_regular(hit.a)) continue;
            if (!hit.valid) continue;
            Splat sp;
            splat_weights(sp, hit.be, hit.bu, E, U);
            if (!sp.on) continue;
            touched = true;
            // gather the four gradient taps
            float g1, g2, g3, g4;
            const int ce = sp.ie - bc.e0, cu = sp.iu - bc.u0;
            if (((unsigned)ce < (unsigned)bc.wwm1) && ((unsigned)cu < (unsigned)bc.whm1)) {
                const float* b = bc.win_g + cu * bc.ww + ce;
                g1 = b[bc.ww]; g2 = b[bc.ww + 1]; g3 = b[1]; g4 = b[0];
            } else {
                const float* row_hi = bc.gf + (size_t)(U - 1 - (sp.iu + 1)) * E + sp.ie;
                const float* row_lo = row_hi + E;
                g1 = __ldg(row_hi); g2 = __ldg(row_hi + 1); g3 = __ldg(row_lo + 1); g4 = __ldg(row_lo);
            }
            const float g_int0 = sp.whu * (sp.wle * g1 + sp.whe * g2) + sp.wlu * (sp.whe * g3 + sp.wle * g4);
            float gdb0 = 0.f, gdb1 = 0.f, gdb2 = 0.f;
            float unblocked = 1.0f;
            if (BLK && bcls == 2) {   // intensity = lambert * (1 - blocked) * k_or
                const BlockBack bb = block_backward(bc.blk, bmask, bc.bp, pc.o0, pc.o1, pc.o2, s.dx, s.dy, s.dz,
                                                    -g_int0 * hit.lam * k_or, bc.gacc, bc.gacc_copies, bc.gacc_stride);
                unblocked = 1.0f - bb.blocked;
                gow0 += bb.go0; gow1 += bb.go1; gow2 += bb.go2;
                gdb0 = bb.gd0; gdb1 = bb.gd1; gdb2 = bb.gd2;
            }
            const float g_int = g_int0 * unblocked;
            const float inten = hit.lam * k_or * unblocked;
            const float g_be = inten * (sp.whu * (g2 - g1) + sp.wlu * (g3 - g4));
            const float g_bu = inten * (sp.wle * (g1 - g4) + sp.whe * (g2 - g3));
            float gdx, gdy, gdz;
            if (PLANAR) {
                const float g_a0 = -g_int * k_int;                 // lam = mag * (-a)
                const float gX = -g_be * k_e, gZ = g_bu * k_u;     // be = (E-1) - te/w*(E-1)
                go0 += gX; go2 += gZ;
                const float gt = gX * s.dx + gZ * s.dz;
                const float gnum = __fdividef(gt, hit.a);          // t = num / a
                const float g_a = g_a0 - gnum * hit.t;
                go0 -= gnum * T.n0; go1 -= gnum * T.n1; go2 -= gnum * T.n2;
                gdx = gX * hit.t + g_a * T.n0;
                gdy = g_a * T.n1;
                gdz = gZ * hit.t + g_a * T.n2;
            } else {
                const float g_ang = g_be * (T.em1 / T.opn), g_z = g_bu * k_u;
                const float inv_n = 1.0f / hit.nrm, inv_n2 = inv_n * inv_n;
                const float sdot = hit.dlx * hit.x + hit.dly * hit.y;    // lam = -sdot / nrm  (if positive)
                const float g_lam = (sdot < 0.0f) ? g_int * k_int : 0.0f;
                const float gs = -g_lam * inv_n, gn = g_lam * sdot * inv_n2;
                float gx = gs * hit.dlx + gn * hit.x * inv_n - g_ang * hit.y * inv_n2;
                float gy = gs * hit.dly + gn * hit.y * inv_n + g_ang * hit.x * inv_n2;
                float gdlx = gs * hit.x + gx * hit.t;
                float gdly = gs * hit.y + gy * hit.t;
                float gdlz = g_z * hit.t;
                const float gt = gx * hit.dlx + gy * hit.dly + g_z * hit.dlz;
                const float inv_2a = 0.5f / hit.qa;
                float gqb = -gt * inv_2a;
                const float gsq = (hit.near_root ? -gt : gt) * inv_2a;
                float gqa = -gt * hit.t / hit.qa;
                const float gdisc = gsq * 0.5f / hit.sq;
                gqb += 2.0f * hit.qb * gdisc;
                gqa -= 4.0f * pc.cc * gdisc;
                const float gcc = -4.0f * hit.qa * gdisc;
                go0 += gx + 2.0f * pc.ox * gcc + 2.0f * hit.dlx * gqb;
                go1 += gy + 2.0f * pc.oy * gcc + 2.0f * hit.dly * gqb;
                go2 += g_z;
                gdlx += 2.0f * pc.ox * gqb + 2.0f * hit.dlx * gqa;
                gdly += 2.0f * pc.oy * gqb + 2.0f * hit.dly * gqa;
                gdx = T.ux * gdlx + T.n0 * gdly + T.ax0 * gdlz;
                gdy = T.uy * gdlx + T.n1 * gdly + T.ax1 * gdlz;
                gdz = T.uz * gdlx + T.n2 * gdly + T.ax2 * gdlz;
            }
            gdx += gdb0; gdy += gdb1; gdz += gdb2;
            // d = M r  ->  grad r += M^T grad d
            gr0 += s.cu * gdx + s.m10 * gdy + s.m20 * gdz;
            gr1 += -s.su * gdx + s.m11 * gdy + s.m21 * gdz;
            gr2 += -s.se * gdy + s.ce * gdz;
        }
        float gp0, gp1, gp2;
        if (PLANAR) { gp0 = go0; gp1 = go1; gp2 = go2; }
        else {  // origin_local = Rot (o - c)
            gp0 = T.ux * go0 + T.n0 * go1 + T.ax0 * go2;
            gp1 = T.uy * go0 + T.n1 * go1 + T.ax1 * go2;
            gp2 = T.uz * go0 + T.n2 * go1 + T.ax2 * go2;
        }
        gp0 += gow0; gp1 += gow1; gp2 += gow2;
        // r = i - 2 (i.n) n   ->   grad n = -2 [ (i.n) grad r + (grad r . n) i ]
        const float grn = gr0 * n4.x + gr1 * n4.y + gr2 * n4.z;
        const float gn0 = -2.0f * (pc.dot * gr0 + grn * i0);
        const float gn1 = -2.0f * (pc.dot * gr1 + grn * i1);
        const float gn2 = -2.0f * (pc.dot * gr2 + grn * i2);
        float4* gpp = reinterpret_cast<float4*>(grad_points) + (size_t)h * P + p;
        float4* gnp = reinterpret_cast<float4*>(grad_normals) + (size_t)h * P + p;
        float4 gp4 = make_float4(gp0, gp1, gp2, 0.f), gn4 = make_float4(gn0, gn1, gn2, 0.f);
        if (ONLY_IRREGULAR) {   // add to what the fast loop wrote for this point
            if (touched) {
                orient_point_backward(src, o_raw, n_raw, gp4, gn4, gori_acc);
                if (grad_points) {
                    const float4 a4 = *gpp, b4 = *gnp;
                    *gpp = make_float4(a4.x + gp4.x, a4.y + gp4.y, a4.z + gp4.z, a4.w + gp4.w);
                    *gnp = make_float4(b4.x + gn4.x, b4.y + gn4.y, b4.z + gn4.z, b4.w + gn4.w);
                }
            }
        } else {
            orient_point_backward(src, o_raw, n_raw, gp4, gn4, gori_acc);
            if (grad_points) {
                *gpp = gp4;
                *gnp = gn4;
            }
        }
    }
}

// Branch-light fast loop for planar targets (see fwd_rays_planar_fast).
template <int THREADS, int TRIG, bool BLK>
__device__ __forceinline__ void bwd_rays_planar_fast(const TraceParams& prm, const TargetCtx& T, const BwdCtx& bc,
                                                     const PointSrc& src, int h, int p_begin, int p_end, float i0, float i1,
                                                     float i2, float* __restrict__ grad_points,
                                                     float* __restrict__ grad_normals, float* gori_acc, int& n_irregular_out) {
    const int tid = threadIdx.x;
    const int P = prm.a.n_points, R = prm.a.n_rays, E = prm.a.res_e, U = prm.a.res_u;
    const float4* pts = src.pts;
    const float4* nrm = src.nrm;
    const float2* dist = reinterpret_cast<const float2*>(prm.a.distortions) + (size_t)h * R * P;
    const float4* trig = (TRIG == AB200_TRIG_TABLE) ? reinterpret_cast<const float4*>(prm.a.trig) + (size_t)h * R * P : nullptr;
    const float mag = prm.a.ray_magnitude;
    const float k_or = prm.a.one_minus_extinction * prm.a.reflectivity;
    const float k_lam = mag * k_or;                         // intensity = k_lam * (-a)
    const float k_e = T.em1 / T.w, k_u = T.um1 / T.h;
    const float e_lim = (float)E, u_lim = (float)U;
    int n_irr = 0;

    for (int p = p_begin + tid; p < p_end; p += THREADS) {
        const float4 o_raw = __ldg(pts + p), n_raw = __ldg(nrm + p);
        float4 o4 = o_raw, n4 = n_raw;
        orient_point(src, o4, n4);
        PointCtx pc;
        make_point(pc, T, i0, i1, i2, o4, n4);
        float go0 = 0.f, go1 = 0.f, go2 = 0.f, gr0 = 0.f, gr1 = 0.f, gr2 = 0.f;
        if (point_regular(pc)) {
            const unsigned long long bmask = (BLK && bc.n_blk) ? block_point_mask(bc.blk, bc.n_blk, bc.bp, pc.o0, pc.o1, pc.o2, pc.r0, pc.r1, pc.r2) : 0ull;
            const float2* dp = dist + p;
            float2 d_next = __ldcs(dp);
#pragma unroll kRayUnroll
            for (int r = 0; r < R; ++r) {
                const float2 d = d_next;
                dp += P;
                if (r + 1 < R) d_next = __ldcs(dp);
                Scatter s;
                if (TRIG == AB200_TRIG_TABLE) {
                    const float4 t4 = __ldg(trig + (size_t)r * P + p);
                    s.cu = t4.x; s.su = t4.y; s.ce = t4.z; s.se = t4.w;
                } else {
                    sincos_poly_core(d.x, &s.su, &s.cu);
                    sincos_poly_core(d.y, &s.se, &s.ce);
                }
                scatter(s, pc);
                const float a = sadd(sadd(smul(s.dx, T.n0), smul(s.dy, T.n1)), smul(s.dz, T.n2));
                if (!(angles_regular<TRIG>(d.x, d.y) && cosine_regular(a))) { ++n_irr; continue; }
                const float t = div_regular(pc.num, a);
                const float X = sadd(pc.o0, smul(s.dx, t));
                const float Z = sadd(pc.o2, smul(s.dz, t));
                const float te = ssub(sadd(X, T.half_w), T.c0);
                const float tu = ssub(sadd(Z, T.half_h), T.c2);
                const float be0 = smul(const_div(te, T.w, T.rw), T.em1);
                const float bu0 = smul(const_div(tu, T.h, T.rh), T.um1);
                const bool valid = (a < 0.0f) && (0.0f <= be0) && (be0 <= T.em1) && (0.0f <= bu0) && (bu0 <= T.um1);
                const float be = ssub(T.em1, be0), bu = bu0;
                const float fe = truncf(be), fu = truncf(bu);
                const float fe1 = sadd(fe, 1.0f), fu1 = sadd(fu, 1.0f);
                if (!(valid && (fe1 < e_lim) && (fu1 < u_lim))) continue;
                int bcls = 0;   // 1: fully shadowed - no intensity, no gradient; 2: inside a sigmoid transition
                if (BLK && bmask) {
                    bcls = block_classify(bc.blk, bmask, bc.bp, pc.o0, pc.o1, pc.o2, s.dx, s.dy, s.dz);
                    if (bcls == 1) continue;
                }
                const int ie = __float2int_rz(be), iu = __float2int_rz(bu);
                const float wle = ssub(fe1, be), wlu = ssub(fu1, bu), whe = ssub(be, fe), whu = ssub(bu, fu);
                float g1, g2, g3, g4;
                const int ce = ie - bc.e0, cu = iu - bc.u0;
                if (((unsigned)ce < (unsigned)bc.wwm1) && ((unsigned)cu < (unsigned)bc.whm1)) {
                    const float* b = bc.win_g + cu * bc.ww + ce;
                    g1 = b[bc.ww]; g2 = b[bc.ww + 1]; g3 = b[1]; g4 = b[0];
                } else {
                    const float* row_hi = bc.gf + (size_t)(U - 1 - (iu + 1)) * E + ie;
                    const float* row_lo = row_hi + E;
                    g1 = __ldg(row_hi); g2 = __ldg(row_hi + 1); g3 = __ldg(row_lo + 1); g4 = __ldg(row_lo);
                }
                float g_int = whu * fmaf(wle, g1, whe * g2) + wlu * fmaf(whe, g3, wle * g4);
                float inten = -a * k_lam;
                float gdb0 = 0.f, gdb1 = 0.f, gdb2 = 0.f;
                if (BLK && bcls == 2) {   // intensity = lambert * (1 - blocked) * k_or  (rare: rays inside a sigmoid transition)
                    const BlockBack bb = block_backward(bc.blk, bmask, bc.bp, pc.o0, pc.o1, pc.o2, s.dx, s.dy, s.dz,
                                                        -g_int * inten, bc.gacc, bc.gacc_copies, bc.gacc_stride);
                    go0 += bb.go0; go1 += bb.go1; go2 += bb.go2;
                    gdb0 = bb.gd0; gdb1 = bb.gd1; gdb2 = bb.gd2;
                    g_int *= 1.0f - bb.blocked;
                    inten *= 1.0f - bb.blocked;
                }
                const float gX = -k_e * inten * fmaf(whu, g2 - g1, wlu * (g3 - g4));   // d/dX through be = (E-1) - te/w*(E-1)
                const float gZ = k_u * inten * fmaf(wle, g1 - g4, whe * (g2 - g3));
                const float gt = fmaf(gX, s.dx, gZ * s.dz);
                const float gnum = __fdividef(gt, a);                                   // t = num / a
                const float g_a = fmaf(-gnum, t, -g_int * k_lam);                       // lam = mag * (-a)
                go0 += fmaf(-gnum, T.n0, gX);
                go1 = fmaf(-gnum, T.n1, go1);
                go2 += fmaf(-gnum, T.n2, gZ);
                const float gdx = fmaf(gX, t, g_a * T.n0) + gdb0, gdy = g_a * T.n1 + gdb1, gdz = fmaf(gZ, t, g_a * T.n2) + gdb2;
                gr0 += fmaf(s.cu, gdx, fmaf(s.m10, gdy, s.m20 * gdz));
                gr1 += fmaf(-s.su, gdx, fmaf(s.m11, gdy, s.m21 * gdz));
                gr2 += fmaf(-s.se, gdy, s.ce * gdz);
            }
        } else {
            n_irr += R;
        }
        const float grn = gr0 * n4.x + gr1 * n4.y + gr2 * n4.z;
        const float gn0 = -2.0f * (pc.dot * gr0 + grn * i0);
        const float gn1 = -2.0f * (pc.dot * gr1 + grn * i1);
        const float gn2 = -2.0f * (pc.dot * gr2 + grn * i2);
        float4 gp4 = make_float4(go0, go1, go2, 0.f), gn4 = make_float4(gn0, gn1, gn2, 0.f);
        orient_point_backward(src, o_raw, n_raw, gp4, gn4, gori_acc);
        if (grad_points) {   // NULL: the caller wants the orientation / blocker gradients only (motor-position optimisation)
            reinterpret_cast<float4*>(grad_points)[(size_t)h * P + p] = gp4;
            reinterpret_cast<float4*>(grad_normals)[(size_t)h * P + p] = gn4;
        }
    }
    n_irregular_out = n_irr;
}

// Packed variant of the backward fast loop: two rays per iteration in fp32x2 registers.  The recomputation of the
// ray (which decides the pixels) repeats the forward's strict operation sequence exactly; the gradient math behind it
// is ordinary packed FMA arithmetic.  A pair with one dead lane (invalid / off-bitmap ray) zeroes that lane's inputs,
// so it contributes exact zeros; a pair with two dead lanes is skipped.
template <int THREADS, int TRIG, bool BLK, bool REVEN, bool DEFER = false>   // DEFER: see fwd_rays_planar_fast2
__device__ __forceinline__ void bwd_rays_planar_fast2(const TraceParams& prm, const TargetCtx& T, const BwdCtx& bc,
                                                      const PointSrc& src, int h, int p_begin, int p_end, float i0, float i1,
                                                      float i2, float* __restrict__ grad_points,
                                                      float* __restrict__ grad_normals, float* gori_acc, int& n_irregular_out,
                                                      const DeferCtx* dc = nullptr) {
    const int tid = threadIdx.x;
    const int P = prm.a.n_points, R = prm.a.n_rays, E = prm.a.res_e, U = prm.a.res_u;
    const float4* pts = src.pts;
    const float4* nrm = src.nrm;
    const float2* dist = reinterpret_cast<const float2*>(prm.a.distortions) + (size_t)h * R * P;
    const float4* trig = (TRIG == AB200_TRIG_TABLE) ? reinterpret_cast<const float4*>(prm.a.trig) + (size_t)h * R * P : nullptr;
    const Packed K(prm.ident);
    const float mag = prm.a.ray_magnitude;
    const float k_or = prm.a.one_minus_extinction * prm.a.reflectivity;
    const float k_lam = mag * k_or;                         // intensity = k_lam * (-a)
    const float k_e = T.em1 / T.w, k_u = T.um1 / T.h;
    const unsigned em1_bits = __float_as_uint(T.em1), um1_bits = __float_as_uint(T.um1);
    const bool axis_n = (T.n0 == 0.0f) && (T.n2 == 0.0f);
    const float2 zero2 = make_float2(0.f, 0.f);
    // window-relative pixel indices straight from the floor() trick (see fwd_rays_planar_fast2)
    const int e0w = bc.ww > 0 ? bc.e0 : 0, u0w = bc.ww > 0 ? bc.u0 : 0;
    float magic_e = 12582912.0f - (float)e0w, magic_u = 12582912.0f - (float)u0w;
    asm volatile("" : "+f"(magic_e), "+f"(magic_u));   // loop invariants: kept in registers, not re-derived per pair
    const unsigned kIdxBits = 0x4B400000u;
    bool any_irr = false;
    const int n_pairs = (R + 1) >> 1;
    const long long step_inner = 2 * (long long)P, step_last = (long long)THREADS - (long long)(n_pairs - 1) * 2 * (long long)P;

    // one pair of rays of a surface point: strict recomputation, tap gather, gradient math; accumulates into go*, gr*
    auto pair_body = [&](const float2 d0, const float2 d1, const bool two, const int r, const int p, const PointCtx& pc,
                         const unsigned long long bmask, float2& go0, float2& go1, float2& go2, float2& gr0, float2& gr1,
                         float2& gr2) {
                float2 cu, su, ce, se;
                bool ang0 = true, ang1 = true;
                if (TRIG == AB200_TRIG_TABLE) {
                    const float4 ta = __ldg(trig + (size_t)r * P + p);
                    const float4 tb = two ? __ldg(trig + (size_t)(r + 1) * P + p) : ta;
                    cu = make_float2(ta.x, tb.x); su = make_float2(ta.y, tb.y); ce = make_float2(ta.z, tb.z); se = make_float2(ta.w, tb.w);
                } else {
                    const float m0 = fmaxf(fabsf(d0.x), fabsf(d0.y)), m1 = fmaxf(fabsf(d1.x), fabsf(d1.y));
                    if (fmaxf(m0, m1) <= kTinyAngle) {   // always, for a physical sun shape (decided per PAIR, identically
                        sincos_tiny2(make_float2(d0.x, d1.x), &su, &cu, K);   // in the forward and the backward kernel)
                        sincos_tiny2(make_float2(d0.y, d1.y), &se, &ce, K);
                    } else {
                        sincos_poly_core2(make_float2(d0.x, d1.x), &su, &cu, K);
                        sincos_poly_core2(make_float2(d0.y, d1.y), &se, &ce, K);
                        ang0 = m0 <= 0.785f;
                        ang1 = m1 <= 0.785f;
                    }
                }
                // ---- strict recomputation (same operations as fwd_rays_planar_fast2) ----
                float2 dx, dz, a, dy_keep;
                {
                const float2 m10 = K.mul(ce, su), m11 = K.mul(ce, cu), m20 = K.mul(se, su), m21 = K.mul(se, cu);
                dx = K.add(K.mul(cu, bc2(pc.r0)), K.mul(su, bc2(-pc.r1)));
                const float2 dy = K.add(K.add(K.mul(m10, bc2(pc.r0)), K.mul(m11, bc2(pc.r1))), K.mul(se, bc2(-pc.r2)));
                dz = K.add(K.add(K.mul(m20, bc2(pc.r0)), K.mul(m21, bc2(pc.r1))), K.mul(ce, bc2(pc.r2)));
                a = axis_n ? K.mul(dy, bc2(T.n1))
                           : K.add(K.add(K.mul(dx, bc2(T.n0)), K.mul(dy, bc2(T.n1))), K.mul(dz, bc2(T.n2)));
                dy_keep = dy;
                }
                const bool fr0 = front_regular(a.x), fr1 = front_regular(a.y);
                float2 rinv;
                float2 t = div_regular2_r(bc2(pc.num), a, K, rinv);
                const float2 X = K.add(bc2(pc.o0), K.mul(dx, t));
                const float2 Z = K.add(bc2(pc.o2), K.mul(dz, t));
                const float2 te = K.sub(K.add(X, bc2(T.half_w)), bc2(T.c0));
                const float2 tu = K.sub(K.add(Z, bc2(T.half_h)), bc2(T.c2));
                const float2 qe0 = K.mul(te, bc2(T.rw)), qu0 = K.mul(tu, bc2(T.rh));
                const float2 be0 = pfma(pfma(pfma(qe0, bc2(-T.w), te), bc2(T.rw), qe0), bc2(T.em1), zero2);
                const float2 bu0 = pfma(pfma(pfma(qu0, bc2(-T.h), tu), bc2(T.rh), qu0), bc2(T.um1), zero2);
                const bool hit0 = ang0 && fr0 && (__float_as_uint(be0.x) <= em1_bits) && (__float_as_uint(bu0.x) <= um1_bits);
                const bool hit1 = two && ang1 && fr1 && (__float_as_uint(be0.y) <= em1_bits) && (__float_as_uint(bu0.y) <= um1_bits);
                // blocking class per lane (block_classify): a fully shadowed ray (1) is treated like an invalid one from here on
                int bcls0 = 0, bcls1 = 0;
                if (BLK && bmask) {
                    if (hit0) bcls0 = block_classify(bc.blk, bmask, bc.bp, pc.o0, pc.o1, pc.o2, dx.x, dy_keep.x, dz.x);
                    if (hit1) bcls1 = block_classify(bc.blk, bmask, bc.bp, pc.o0, pc.o1, pc.o2, dx.y, dy_keep.y, dz.y);
                    if (bcls0 == 1 && bcls1 == 1) return;
                }
                const bool valid0 = hit0 && !(BLK && bcls0 == 1), valid1 = hit1 && !(BLK && bcls1 == 1);
                const float2 be = K.sub(bc2(T.em1), be0), bu = bu0;
                const float2 me = make_float2(__fadd_rd(be.x, magic_e), __fadd_rd(be.y, magic_e));
                const float2 mu = make_float2(__fadd_rd(bu.x, magic_u), __fadd_rd(bu.y, magic_u));
                const int cex0 = (int)(__float_as_uint(me.x) - kIdxBits), cux0 = (int)(__float_as_uint(mu.x) - kIdxBits);   // ie - e0, iu - u0
                const int cex1 = (int)(__float_as_uint(me.y) - kIdxBits), cux1 = (int)(__float_as_uint(mu.y) - kIdxBits);
                if (!(hit0 && hit1))   // rare; irregular implies invalid
                    any_irr |= !ang0 || ((a.x < 0.0f) && !fr0) || (two && (!ang1 || ((a.y < 0.0f) && !fr1)));
                // ---- gather the four gradient taps per live lane ----
                // window interior (implies on the bitmap: ie + 1 < E and iu + 1 < U)
                const bool in0 = valid0 && ((unsigned)cex0 < (unsigned)bc.wwm1) && ((unsigned)cux0 < (unsigned)bc.whm1);
                const bool in1 = valid1 && ((unsigned)cex1 < (unsigned)bc.wwm1) && ((unsigned)cux1 < (unsigned)bc.whm1);
                float2 g1, g2, g3, g4;
                bool live0 = true, live1 = true;
                if (in0 && in1) {   // the common case: eight shared-memory loads
                    const float* b0 = bc.win_g + cux0 * bc.ww + cex0;
                    const float* b1 = bc.win_g + cux1 * bc.ww + cex1;
                    g1.x = b0[bc.ww]; g2.x = b0[bc.ww + 1]; g3.x = b0[1]; g4.x = b0[0];
                    g1.y = b1[bc.ww]; g2.y = b1[bc.ww + 1]; g3.y = b1[1]; g4.y = b1[0];
                } else {
                    g1 = g2 = g3 = g4 = zero2;
                    // on the bitmap: ie + 1 < E and iu + 1 < U (the lower bounds hold for valid rays)
                    const int ie0 = cex0 + e0w, iu0 = cux0 + u0w, ie1 = cex1 + e0w, iu1 = cux1 + u0w;
                    live0 = valid0 && ((unsigned)ie0 < (unsigned)(E - 1)) && ((unsigned)iu0 < (unsigned)(U - 1));
                    live1 = valid1 && ((unsigned)ie1 < (unsigned)(E - 1)) && ((unsigned)iu1 < (unsigned)(U - 1));
                    if (!(live0 || live1)) return;
#pragma unroll
                    for (int lane = 0; lane < 2; ++lane) {
                        if (!(lane ? live1 : live0)) continue;
                        const int ie = lane ? ie1 : ie0, iu = lane ? iu1 : iu0;
                        const int cex = lane ? cex1 : cex0, cux = lane ? cux1 : cux0;
                        float t1, t2, t3, t4;
                        if (lane ? in1 : in0) {
                            const float* b = bc.win_g + cux * bc.ww + cex;
                            t1 = b[bc.ww]; t2 = b[bc.ww + 1]; t3 = b[1]; t4 = b[0];
                        } else {
                            const float* row_hi = bc.gf + (size_t)(U - 1 - (iu + 1)) * E + ie;
                            const float* row_lo = row_hi + E;
                            t1 = __ldg(row_hi); t2 = __ldg(row_hi + 1); t3 = __ldg(row_lo + 1); t4 = __ldg(row_lo);
                        }
                        if (lane) { g1.y = t1; g2.y = t2; g3.y = t3; g4.y = t4; } else { g1.x = t1; g2.x = t2; g3.x = t3; g4.x = t4; }
                    }
                }
                // fractional parts (exact); the low weights are 1 - high here - the last-bit difference to the forward's
                // (ie + 1) - be only touches the gradient VALUE, never the pixel choice
                float2 whe = K.sub(be, K.sub(me, bc2(magic_e))), whu = K.sub(bu, K.sub(mu, bc2(magic_u)));
                float2 dead_scale = make_float2(1.f, 1.f);
                if (!(live0 && live1)) {   // rare: silence the dead lane (its values may be anything, even NaN)
                    if (!live0) {
                        cu.x = su.x = ce.x = se.x = dx.x = dz.x = a.x = t.x = rinv.x = 0.f;
                        whe.x = whu.x = 0.f; dead_scale.x = 0.f;
                    } else {
                        cu.y = su.y = ce.y = se.y = dx.y = dz.y = a.y = t.y = rinv.y = 0.f;
                        whe.y = whu.y = 0.f; dead_scale.y = 0.f;
                    }
                }
                // ---- gradient math (packed, FMA contraction allowed) ----
                const float2 wle = lsub(dead_scale, whe), wlu = lsub(dead_scale, whu);   // 1 - high (0 on a dead lane)
                float2 g_int = lfma(whu, lfma(wle, g1, lmul(whe, g2)), lmul(wlu, lfma(whe, g3, lmul(wle, g4))));
                const float2 dgx = lfma(whu, lsub(g2, g1), lmul(wlu, lsub(g3, g4)));
                const float2 dgz = lfma(wle, lsub(g1, g4), lmul(whe, lsub(g2, g3)));
                float2 inten = lmul(a, bc2(-k_lam));
                float2 gdb0 = zero2, gdb1 = zero2, gdb2 = zero2;
                if (BLK && (bcls0 == 2 || bcls1 == 2)) {   // intensity = lambert * (1 - blocked) * k_or  (rare: sigmoid transitions)
                    float2 unblocked = make_float2(1.f, 1.f);
#pragma unroll
                    for (int lane = 0; lane < 2; ++lane) {
                        if (!(lane ? live1 : live0) || (lane ? bcls1 : bcls0) != 2) continue;
                        const BlockBack bb = block_backward(bc.blk, bmask, bc.bp, pc.o0, pc.o1, pc.o2,
                                                            lane ? dx.y : dx.x, lane ? dy_keep.y : dy_keep.x, lane ? dz.y : dz.x,
                                                            -(lane ? g_int.y : g_int.x) * (lane ? inten.y : inten.x), bc.gacc, bc.gacc_copies, bc.gacc_stride);
                        if (lane) { go0.y += bb.go0; go1.y += bb.go1; go2.y += bb.go2; gdb0.y = bb.gd0; gdb1.y = bb.gd1; gdb2.y = bb.gd2; unblocked.y = 1.0f - bb.blocked; }
                        else      { go0.x += bb.go0; go1.x += bb.go1; go2.x += bb.go2; gdb0.x = bb.gd0; gdb1.x = bb.gd1; gdb2.x = bb.gd2; unblocked.x = 1.0f - bb.blocked; }
                    }
                    g_int = lmul(g_int, unblocked);
                    inten = lmul(inten, unblocked);
                }
                // d/dX through be = (E-1) - te/w*(E-1), d/dZ through bu = tu/h*(U-1)
                const float2 gX = lmul(lmul(inten, bc2(-k_e)), dgx);
                const float2 gZ = lmul(lmul(inten, bc2(k_u)), dgz);
                const float2 gt = lfma(gX, dx, lmul(gZ, dz));
                const float2 ngnum = lmul(gt, make_float2(-rinv.x, -rinv.y));        // -(gt / a):  t = num / a
                const float2 g_a = lfma(ngnum, t, lmul(g_int, bc2(-k_lam)));         // lam = mag * (-a)
                go0 = lfma(ngnum, bc2(T.n0), ladd(go0, gX));
                go1 = lfma(ngnum, bc2(T.n1), go1);
                go2 = lfma(ngnum, bc2(T.n2), ladd(go2, gZ));
                float2 gdx = lfma(gX, t, lmul(g_a, bc2(T.n0))), gdy = lmul(g_a, bc2(T.n1)), gdz = lfma(gZ, t, lmul(g_a, bc2(T.n2)));
                if (BLK) { gdx = ladd(gdx, gdb0); gdy = ladd(gdy, gdb1); gdz = ladd(gdz, gdb2); }
                // d = M r  ->  grad r += M^T grad d
                //   M^T = [cu, ce su, se su; -su, ce cu, se cu; 0, -se, ce]: grouped as ce * gdy + se * gdz and -se * gdy + ce * gdz
                const float2 q = lfma(ce, gdy, lmul(se, gdz));
                gr0 = lfma(cu, gdx, lfma(su, q, gr0));
                gr1 = lfma(make_float2(-su.x, -su.y), gdx, lfma(cu, q, gr1));
                gr2 = lfma(make_float2(-se.x, -se.y), gdy, lfma(ce, gdz, gr2));
    };
    // per-point epilogue: gradients w.r.t. the (un-aligned) point and normal rows from the sums over the point's rays
    auto point_epilogue = [&](const int p, const float s_go0, const float s_go1, const float s_go2, const float s_gr0,
                              const float s_gr1, const float s_gr2) {
        // the point's rows again (L1/L2 hit): cheaper than keeping 11 registers alive across the ray loop
        const float4 o_raw = __ldg(pts + p), n_raw = __ldg(nrm + p);
        float4 o4 = o_raw, n4 = n_raw;
        orient_point(src, o4, n4);
        const float dot_in = sadd(sadd(smul(i0, n4.x), smul(i1, n4.y)), smul(i2, n4.z));
        // r = i - 2 (i.n) n   ->   grad n = -2 [ (i.n) grad r + (grad r . n) i ]
        const float grn = s_gr0 * n4.x + s_gr1 * n4.y + s_gr2 * n4.z;
        const float gn0 = -2.0f * (dot_in * s_gr0 + grn * i0);
        const float gn1 = -2.0f * (dot_in * s_gr1 + grn * i1);
        const float gn2 = -2.0f * (dot_in * s_gr2 + grn * i2);
        float4 gp4 = make_float4(s_go0, s_go1, s_go2, 0.f), gn4 = make_float4(gn0, gn1, gn2, 0.f);
        orient_point_backward(src, o_raw, n_raw, gp4, gn4, gori_acc);
        if (grad_points) {   // NULL: the caller wants the orientation / blocker gradients only (motor-position optimisation)
            reinterpret_cast<float4*>(grad_points)[(size_t)h * P + p] = gp4;
            reinterpret_cast<float4*>(grad_normals)[(size_t)h * P + p] = gn4;
        }
    };
    // The CTA's last, partial round (npts mod THREADS points - 16 of 10000 with 768 threads) would keep one warp busy for a
    // whole point while all others wait at the barrier: those points are split by ray PAIR instead - `slots` (a power of
    // two) lanes per point, one pair each, partial sums combined by a shuffle tree in a fixed order.
    const int npts = p_end - p_begin, n_full = (npts / THREADS) * THREADS, n_tail = npts - n_full;
    int slots = 1;
    while (slots < n_pairs) slots <<= 1;
    const bool split_tail = AB200_BWD_SPLIT_TAIL && REVEN && TRIG != AB200_TRIG_TABLE && n_full > 0 && n_tail > 0 && slots <= 32 && n_tail * slots <= THREADS &&
                            !DEFER && !BLK;   // (the passes of a blocking trace handle whole points)
    const int p_end_main = split_tail ? p_begin + n_full : p_end;

    // BLK instantiation = pass 2 of a blocking trace over the deferral list `dc` (see fwd_rays_planar_fast2)
    int it = tid;
    const int n_items = BLK ? dc->count : 0;
    int p = BLK ? (it < n_items ? p_begin + defer_lookup(*dc, it) : p_end_main) : p_begin + tid;
    int pn = 0;
    float2 da = zero2, db = zero2;
    const float2* nx = dist + p;      // address of the pair loaded NEXT (always one pair ahead of the math)
    if (p < p_end_main) { da = __ldcs(nx); if (R > 1) db = __ldcs(nx + P); }
    for (; p < p_end_main; p = BLK ? pn : p + THREADS) {
        bool more;
        if (BLK) {
            it += THREADS;
            more = it < n_items;
            pn = more ? p_begin + defer_lookup(*dc, it) : p_end_main;
        } else {
            pn = p + THREADS;
            more = pn < p_end_main;
        }
        if (more) { prefetch_l2(pts + pn); prefetch_l2(nrm + pn); }
        else if (tid * kWindowSampleStride < P && src.next && src.next[0]) {
            prefetch_l2(src.next[0] + tid * kWindowSampleStride); prefetch_l2(src.next[1] + tid * kWindowSampleStride);
        }
        PointCtx pc;
        {
            float4 o4 = __ldg(pts + p), n4 = __ldg(nrm + p);
            orient_point(src, o4, n4);
            make_point(pc, T, i0, i1, i2, o4, n4);
        }
        float2 go0 = zero2, go1 = zero2, go2 = zero2, gr0 = zero2, gr1 = zero2, gr2 = zero2;   // one partial sum per lane
        if (point_regular(pc)) {
            const int pcls = DEFER ? defer_point(dc, p - p_begin, pc.o0, pc.o1, pc.o2, pc.r0, pc.r1, pc.r2) : kPointClear;
            if (pcls != kPointClear) {   // shadow-affected: pass 2 traces the point and writes its rows; completely shadowed:
                nx += THREADS;           // zero rows from the epilogue below
                if (more) { da = __ldcs(nx); if (R > 1) db = __ldcs(nx + P); }
                if (pcls == kPointDeferred) continue;
            }
            const unsigned long long bmask = (BLK && bc.n_blk) ? block_point_mask(bc.blk, bc.n_blk, bc.bp, pc.o0, pc.o1, pc.o2, pc.r0, pc.r1, pc.r2) : 0ull;
            for (int r = 0; r < (pcls == kPointClear ? R : 0); r += 2) {
                const bool two = REVEN || (r + 1 < R);
                const float2 d0 = da, d1 = db;
                {
                    const bool inner = r + 2 < R;
                    if (BLK) nx = inner ? nx + step_inner : dist + pn; else nx += inner ? step_inner : step_last;
                    if (inner || more) {
                        da = __ldcs(nx);
                        if (REVEN || (inner ? (r + 3 < R) : (R > 1))) db = __ldcs(nx + P);
                    }
                }
                pair_body(d0, d1, two, r, p, pc, bmask, go0, go1, go2, gr0, gr1, gr2);
            }
        } else {
            any_irr = true;
            if (BLK) nx = dist + pn; else nx += THREADS;
            if (more) { da = __ldcs(nx); if (R > 1) db = __ldcs(nx + P); }
        }
        point_epilogue(p, go0.x + go0.y, go1.x + go1.y, go2.x + go2.y, gr0.x + gr0.y, gr1.x + gr1.y, gr2.x + gr2.y);
    }
    if (split_tail && (tid & ~31) < n_tail * slots) {   // (whole warps, so that the shuffles below are convergent)
        const int lp = tid / slots, k = tid - lp * slots;
        const bool in_range = lp < n_tail, act = in_range && k < n_pairs;
        const int pt = p_begin + n_full + (in_range ? lp : 0);
        float2 go0 = zero2, go1 = zero2, go2 = zero2, gr0 = zero2, gr1 = zero2, gr2 = zero2;
        if (act) {
            PointCtx pc;
            float4 o4 = __ldg(pts + pt), n4 = __ldg(nrm + pt);
            orient_point(src, o4, n4);
            make_point(pc, T, i0, i1, i2, o4, n4);
            if (point_regular(pc)) {
                const unsigned long long bmask = (BLK && bc.n_blk) ? block_point_mask(bc.blk, bc.n_blk, bc.bp, pc.o0, pc.o1, pc.o2, pc.r0, pc.r1, pc.r2) : 0ull;
                const float2 d0 = __ldcs(dist + (size_t)(2 * k) * P + pt), d1 = __ldcs(dist + (size_t)(2 * k + 1) * P + pt);
                pair_body(d0, d1, true, 2 * k, pt, pc, bmask, go0, go1, go2, gr0, gr1, gr2);
            } else {
                any_irr = true;
            }
        }
        float s[6] = {go0.x + go0.y, go1.x + go1.y, go2.x + go2.y, gr0.x + gr0.y, gr1.x + gr1.y, gr2.x + gr2.y};
        for (int off = slots >> 1; off > 0; off >>= 1) {
#pragma unroll
            for (int q = 0; q < 6; ++q) s[q] += __shfl_xor_sync(0xffffffffu, s[q], off);
        }
        if (in_range && k == 0) point_epilogue(pt, s[0], s[1], s[2], s[3], s[4], s[5]);
    }
    n_irregular_out = any_irr ? 1 : 0;
}

template <int THREADS, int TRIG, bool BLK, bool MAP = false>
__global__ void __launch_bounds__(THREADS, (THREADS > 512 ? 1 : 2))
trace_bwd_kernel(const TraceParams prm, const float* __restrict__ grad_flux, const long long grad_stride,
                 float* __restrict__ grad_points, float* __restrict__ grad_normals, float* __restrict__ grad_prims,
                 float* __restrict__ grad_orientations, float* __restrict__ grad_prims_scratch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* win_g = reinterpret_cast<float*>(smem_raw);
    __shared__ TargetCtx T_sh;
    __shared__ Window win_sh;
    __shared__ float red[6 * 32];
    __shared__ float O_sh[16];
    __shared__ BlockPrim blk_sh[BLK ? kMaxBlockCandidates : 1];
    __shared__ int blk_rows_sh[BLK ? kMaxBlockCandidates : 1];
    __shared__ double gacc_sh[BLK ? kMaxBlockCandidates * 12 : 1];
    __shared__ unsigned defer_bits_sh[BLK ? kDeferWords : 1];
    __shared__ int defer_prefix_sh[BLK ? kDeferWords : 1];
    __shared__ int defer_total_sh;
    __shared__ const float4* next_sh[2];
    __shared__ __align__(8) unsigned long long stage_bar;   // mbarrier of the gradient-window staging (TMA bulk copies)

    const int tid = threadIdx.x;
    const int li = blockIdx.x / prm.split;
    const int chunk = blockIdx.x - li * prm.split;
    const int h = prm.a.local_rows ? min(max(prm.a.local_rows[li], 0), prm.a.n_samples - 1) : li;
    const int P = prm.a.n_points, E = prm.a.res_e, U = prm.a.res_u;
    const int p_begin = chunk * prm.pts_per_chunk;
    const int p_end = min(P, p_begin + prm.pts_per_chunk);

    long long t_phase = (prm.a.stats && tid == 0) ? clock64() : 0;
    // start-up loads issued together before the first barrier (see trace_fwd_kernel)
    PointSrc src;
    const int hs = MAP ? __ldg(prm.a.src_rows + h) : h;   // activation index map (the GRADIENT rows stay per sample)
    src.pts = reinterpret_cast<const float4*>(prm.a.points) + (size_t)hs * P;
    src.nrm = reinterpret_cast<const float4*>(prm.a.normals) + (size_t)hs * P;
    src.O = prm.a.orientations ? O_sh : nullptr;
    // the forward of the same call recorded its window (ab200_trace_args::windows): no sampling, no placement
    const bool have_window = prm.a.windows != nullptr && prm.split == 1;
    const int p_end_samples = have_window ? p_begin : p_end;   // (no surface samples are read then)
    WindowSamples ws;
    load_window_samples<THREADS>(ws, src.pts, src.nrm, p_begin, p_end_samples);
    const float i0 = __ldg(prm.a.incident + 4 * h), i1 = __ldg(prm.a.incident + 4 * h + 1),
                i2 = __ldg(prm.a.incident + 4 * h + 2);
    if (tid == 0) load_target(T_sh, prm.a.targets, prm.a.target_idx[h], E, U);
    if (prm.a.orientations && tid >= 32 && tid < 48) O_sh[tid - 32] = __ldg(prm.a.orientations + (size_t)h * 16 + (tid - 32));
    if (tid == 64) set_next_sample<MAP>(next_sh, prm, li);
    const unsigned stage_bar_s = (unsigned)__cvta_generic_to_shared(&stage_bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(stage_bar_s), "r"(THREADS) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    AB200_PHASE(12, 0);
    src.next = next_sh;
    const TargetCtx T = T_sh;
    float gori[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) gori[q] = 0.f;
    float* gori_acc = (grad_orientations && src.O) ? gori : nullptr;

    Window W;
    place_window<THREADS>(W, prm, T, src, ws, p_begin, p_end_samples, i0, i1, i2, red, &win_sh,
                          have_window ? reinterpret_cast<const int4*>(prm.a.windows) + h : nullptr);
    const bool quad_w = prm.quad;
    AB200_PHASE(12, 1);
    const float* gf = grad_flux + (size_t)h * grad_stride;
    {   // stage the gradient window: win_g[(iu-u0)*ww + (ie-e0)] = grad_flux[h, U-1-iu, ie]
        const int warp = tid >> 5, lane = tid & 31, nwarps = THREADS / 32;
        if (AB200_BWD_TMA_STAGE && quad_w && W.ww > 0) {
            // TMA: one 1-D bulk copy (cp.async.bulk, global -> shared) per window row, completion counted in bytes on an
            // mbarrier every thread arrives on - no registers, no LSU instructions per element, and the wait doubles as
            // the CTA barrier in front of the ray loop
            const unsigned row_bytes = (unsigned)W.ww * 4u;
            const unsigned win_s = (unsigned)__cvta_generic_to_shared(win_g);
            unsigned my_bytes = 0;
            for (int r = tid; r < W.wh; r += THREADS) my_bytes += row_bytes;
            unsigned long long state;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 %0, [%1], %2;" : "=l"(state) : "r"(stage_bar_s), "r"(my_bytes) : "memory");
            (void)state;
            for (int r = tid; r < W.wh; r += THREADS) {
                const float* grow = gf + (size_t)(U - 1 - (W.u0 + r)) * E + W.e0;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(win_s + (unsigned)r * row_bytes), "l"(grow), "r"(row_bytes), "r"(stage_bar_s) : "memory");
            }
            asm volatile("{\n"
                         ".reg .pred p;\n"
                         "STAGE_WAIT_%=:\n"
                         "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
                         "@!p bra STAGE_WAIT_%=;\n"
                         "}\n" ::"r"(stage_bar_s) : "memory");
        } else if (quad_w && W.ww > 0) {
            const int qw = W.ww >> 2, n4 = W.wh * qw;
            const int d_r = THREADS / qw, d_q = THREADS - d_r * qw;
            float4* w4 = reinterpret_cast<float4*>(win_g);
            int r = tid / qw, q = tid - r * qw;
            for (int idx = tid; idx < n4; idx += THREADS) {
                w4[idx] = __ldg(reinterpret_cast<const float4*>(gf + (size_t)(U - 1 - (W.u0 + r)) * E + W.e0) + q);
                r += d_r; q += d_q;
                if (q >= qw) { q -= qw; ++r; }
            }
        } else
        for (int r = warp; r < W.wh; r += nwarps) {
            const float* grow = gf + (size_t)(U - 1 - (W.u0 + r)) * E + W.e0;
            for (int c = lane; c < W.ww; c += 32) win_g[r * W.ww + c] = __ldg(grow + c);
        }
    }
    if (!(AB200_BWD_TMA_STAGE && quad_w && W.ww > 0)) __syncthreads();
    AB200_PHASE(12, 2);  // staging the gradient window
    BwdCtx bc;
    bc.win_g = win_g; bc.gf = gf; bc.e0 = W.e0; bc.u0 = W.u0; bc.ww = W.ww;
    bc.wwm1 = W.ww > 1 ? W.ww - 1 : 0; bc.whm1 = W.wh > 1 ? W.wh - 1 : 0;
    bc.blk = blk_sh; bc.blk_rows = blk_rows_sh; bc.n_blk = 0; bc.gacc = (BLK && grad_prims) ? gacc_sh : nullptr;
    bc.gacc_copies = 1; bc.gacc_stride = 0;
    if (BLK && grad_prims)
        for (int i = tid; i < kMaxBlockCandidates * 12; i += THREADS) gacc_sh[i] = 0.0;   // (barriers follow before any use)
    if (BLK && prm.a.blockers.n_blockers > 0) {
        const ab200_blockers& B = prm.a.blockers;
        bc.bp.softness = B.softness; bc.bp.alpha = B.alpha; bc.bp.offset = B.ray_origin_offset; bc.bp.epsilon = B.epsilon;
        bc.bp.cull_angle = B.cull_angle;
        bc.n_blk = block_load_candidates<THREADS>(blk_sh, blk_rows_sh, B, h, tid);
        int slots = 1;
        while (slots < bc.n_blk) slots <<= 1;
        bc.gacc_copies = kMaxBlockCandidates / slots;    // 4 candidates: 16 copies
        bc.gacc_stride = slots * 12;
    }
    // Blocking: two passes, see trace_fwd_kernel and blocking_device.cuh ("Deferral")
    const bool blocking = BLK && bc.n_blk > 0;
    const bool defer = BLK && (p_end - p_begin) <= kDeferWords * 32;
    DeferCtx dc;
    dc.bits = defer_bits_sh; dc.prefix = defer_prefix_sh; dc.n_words = (p_end - p_begin + 31) >> 5; dc.count = 0;
    dc.blk = bc.blk; dc.n_blk = bc.n_blk; dc.bp = bc.bp;
    if (BLK && defer) {
        for (int w = tid; w < dc.n_words; w += THREADS) defer_bits_sh[w] = 0u;
        __syncthreads();
    }
    const DeferCtx* rec = (BLK && defer) ? &dc : nullptr;
    const int rec_mode = (BLK && defer) ? kDeferRecord : kDeferOff;
    if (BLK && !defer) {
        if (T.planar)
            bwd_rays<THREADS, TRIG, true, false, false, BLK>(prm, T, bc, src, h, p_begin, p_end, i0, i1, i2, grad_points, grad_normals, gori_acc);
        else
            bwd_rays<THREADS, TRIG, false, false, false, BLK>(prm, T, bc, src, h, p_begin, p_end, i0, i1, i2, grad_points, grad_normals, gori_acc);
    } else if (T.planar && T.fastdiv && TRIG != AB200_TRIG_SINCOSF) {
        int n_irr = 0;
#if AB200_PACKED_RAYS
        if ((prm.a.n_rays & 1) == 0)
            bwd_rays_planar_fast2<THREADS, TRIG, false, true, BLK>(prm, T, bc, src, h, p_begin, p_end, i0, i1, i2, grad_points, grad_normals, gori_acc, n_irr, rec);
        else
            bwd_rays_planar_fast2<THREADS, TRIG, false, false, BLK>(prm, T, bc, src, h, p_begin, p_end, i0, i1, i2, grad_points, grad_normals, gori_acc, n_irr, rec);
#else   // tuning build without the packed loops: inline mask, nothing deferred
        bwd_rays_planar_fast<THREADS, TRIG, BLK>(prm, T, bc, src, h, p_begin, p_end, i0, i1, i2, grad_points, grad_normals, gori_acc, n_irr);
#endif
        if (__syncthreads_or(n_irr != 0)) {   // never with physical inputs (each thread re-reads only its own points)
            const bool skip_masked = AB200_PACKED_RAYS && rec != nullptr;
            bwd_rays<THREADS, TRIG, true, false, true, BLK>(prm, T, bc, src, h, p_begin, p_end, i0, i1, i2, grad_points, grad_normals, gori_acc,
                                                            rec, skip_masked ? kDeferSkip : kDeferOff);
        }
    } else if (T.planar) {
        bwd_rays<THREADS, TRIG, true, false, false, false>(prm, T, bc, src, h, p_begin, p_end, i0, i1, i2, grad_points, grad_normals, gori_acc, rec, rec_mode);
    } else {
        bwd_rays<THREADS, TRIG, false, false, false, false>(prm, T, bc, src, h, p_begin, p_end, i0, i1, i2, grad_points, grad_normals, gori_acc, rec, rec_mode);
    }
    if (BLK && defer && blocking) {   // pass 2 (uniform over the CTA)
        AB200_PHASE(12, 3);  // ray loop, pass 1
        const int n_def = defer_scan<THREADS>(dc, &defer_total_sh);
        if (n_def > 0) {
#if AB200_PACKED_RAYS
            if (T.planar && T.fastdiv && TRIG != AB200_TRIG_SINCOSF) {
                int n_irr = 0;
                if ((prm.a.n_rays & 1) == 0)
                    bwd_rays_planar_fast2<THREADS, TRIG, BLK, true>(prm, T, bc, src, h, p_begin, p_end, i0, i1, i2, grad_points, grad_normals, gori_acc, n_irr, &dc);
                else
                    bwd_rays_planar_fast2<THREADS, TRIG, BLK, false>(prm, T, bc, src, h, p_begin, p_end, i0, i1, i2, grad_points, grad_normals, gori_acc, n_irr, &dc);
                if (__syncthreads_or(n_irr != 0))   // never with physical inputs
                    bwd_rays<THREADS, TRIG, true, false, true, BLK>(prm, T, bc, src, h, p_begin, p_end, i0, i1, i2, grad_points, grad_normals, gori_acc, &dc, kDeferList);
            } else
#endif
            if (T.planar)
                bwd_rays<THREADS, TRIG, true, false, false, BLK>(prm, T, bc, src, h, p_begin, p_end, i0, i1, i2, grad_points, grad_normals, gori_acc, &dc, kDeferList);
            else
                bwd_rays<THREADS, TRIG, false, false, false, BLK>(prm, T, bc, src, h, p_begin, p_end, i0, i1, i2, grad_points, grad_normals, gori_acc, &dc, kDeferList);
        }
        AB200_PHASE(14, 5);  // stats[19]: pass 2 (deferred points)
    }
    if (BLK && grad_prims && bc.n_blk > 0) {
        // blocker gradients of this CTA: one row per candidate slot into the caller's scratch (summed over the samples in a
        // fixed order by blocker_grad_reduce_kernel), or - without scratch - float atomics on the primitive table's rows
        __syncthreads();
        const int max_cand = prm.a.blockers.max_candidates;
        for (int i = tid; i < bc.n_blk * 12; i += THREADS) {
            double sum = 0.0;
            for (int k = 0; k < bc.gacc_copies; ++k) sum += gacc_sh[k * bc.gacc_stride + i];   // fixed order
            const float v = (float)sum;
            if (grad_prims_scratch) grad_prims_scratch[((size_t)blockIdx.x * max_cand + i / 12) * 12 + i % 12] = v;
            else if (v != 0.0f) atomicAdd(grad_prims + (size_t)blk_rows_sh[i / 12] * 12 + i % 12, v);
        }
    }
    AB200_PHASE(12, 3);  // ray loop (thread 0)
    if (gori_acc) {
        // dL/dO rows 0..2 of this sample: warp shuffle tree, then the warps' partial sums in index order (fixed order);
        // the gradient window in shared memory is dead by now and serves as scratch
        __syncthreads();
        const int warp = tid >> 5, lane = tid & 31, nwarps = THREADS / 32;
#pragma unroll
        for (int q = 0; q < 12; ++q) {
            const float v = warp_sumf(gori[q]);
            if (lane == 0) win_g[warp * 12 + q] = v;
        }
        __syncthreads();
        if (tid < 12) {
            float sum = 0.f;
            for (int w = 0; w < nwarps; ++w) sum += win_g[w * 12 + tid];
            float* dst = grad_orientations + (size_t)h * 16 + tid;
            if (prm.split == 1) *dst = sum; else atomicAdd(dst, sum);   // caller zeroes the buffer (row 3 stays 0)
        }
        AB200_PHASE(12, 4);  // waiting for the last warp + dL/dO reduction
    }
}

#ifndef AB200_TU_FWD
// Blocker gradients, second stage: row k of the primitive table = the sum of the per-CTA rows of every CTA whose sample has
// primitive k in its candidate list, in a fixed order (one warp per primitive: lane l takes CTAs l, l + 32, ... in
// ascending order, then a shuffle tree) - no atomics, reproducible.  OVERWRITES grad_prims.
__global__ void __launch_bounds__(256) blocker_grad_reduce_kernel(const float* __restrict__ partial, const int* __restrict__ cand_idx,
                                                                  const int* __restrict__ cand_count,
                                                                  const int* __restrict__ local_rows, int n_ctas, int split,
                                                                  int max_cand, int n_prims, float* __restrict__ grad_prims) {
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (k >= n_prims) return;
    float acc[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) acc[q] = 0.0f;
    for (int cta = lane; cta < n_ctas; cta += 32) {
        const int li = cta / split;
        const int h = local_rows ? local_rows[li] : li;
        const int cnt = min(cand_count[h], max_cand);
        const int* cand = cand_idx + (size_t)h * max_cand;
        for (int c = 0; c < cnt; ++c) {
            if (__ldg(cand + c) != k) continue;
            const float* row = partial + ((size_t)cta * max_cand + c) * 12;
#pragma unroll
            for (int q = 0; q < 12; ++q) acc[q] += row[q];
        }
    }
#pragma unroll
    for (int q = 0; q < 12; ++q) {
        float v = acc[q];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) grad_prims[(size_t)k * 12 + q] = v;
    }
}

#endif  // !AB200_TU_FWD

#ifndef AB200_TU_BWD
#include "trace_v3.cuh"
#endif

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int32_t validate(const ab200_trace_args* a) {
    AB200_REQUIRE(a != nullptr, AB200_EINVAL, "args is NULL");
    AB200_REQUIRE(a->abi_version == AB200_ABI_VERSION, AB200_EINVAL, "abi_version %d != %d", a->abi_version, AB200_ABI_VERSION);
    AB200_REQUIRE(a->n_samples >= 0 && a->n_points > 0 && a->n_rays > 0, AB200_EINVAL, "bad sizes N=%d P=%d R=%d",
                  a->n_samples, a->n_points, a->n_rays);
    AB200_REQUIRE(a->res_e >= 2 && a->res_u >= 2, AB200_EINVAL, "bitmap resolution must be >= 2x2 (got %dx%d)", a->res_e, a->res_u);
    AB200_REQUIRE(a->n_local >= 0 && a->n_local <= a->n_samples, AB200_EINVAL, "n_local %d out of range", a->n_local);
    AB200_REQUIRE((long long)a->n_points * a->n_rays < (1ll << 30), AB200_ELIMIT, "P*R too large");
    AB200_REQUIRE(a->points && a->normals && a->incident && a->distortions && a->target_idx, AB200_EINVAL, "NULL input pointer");
    AB200_REQUIRE(a->trig_mode != AB200_TRIG_TABLE || a->trig, AB200_EINVAL, "trig_mode TABLE needs args->trig");
    AB200_REQUIRE(a->trig_mode >= 0 && a->trig_mode <= 2, AB200_EINVAL, "bad trig_mode %d", a->trig_mode);
    AB200_REQUIRE(a->targets.n_planar + a->targets.n_cyl > 0, AB200_EINVAL, "no target areas");
    AB200_REQUIRE(a->blockers.n_blockers == 0 || (a->blockers.prims && a->blockers.cand_idx && a->blockers.cand_count &&
                                                   a->blockers.max_candidates >= 1 && a->blockers.max_candidates <= kMaxBlockCandidates),
                  AB200_EINVAL, "blocking needs prims, cand_idx, cand_count and 1 <= max_candidates <= %d", kMaxBlockCandidates);
    return AB200_OK;
}

struct LaunchPlan {
    int threads, split, pts_per_chunk, smem_bytes;
};

// dynamic shared memory available to one CTA next to the kernels' static shared data (target, window, reduction
// scratch, 64 blocking primitives: < 6 KB)
static int max_window_bytes(bool blocking) {
    int dev = 0, optin = 227 * 1024;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    return optin - (blocking ? 20 : 6) * 1024;   // blocking: + 64 candidates x 96 B, the deferral bit set and its prefix sums,
                                                 // 64 x 12 double gradient accumulators (backward)
}

static LaunchPlan make_plan(int n_local, int n_points, int max_threads_large, bool force_large = false, bool blocking = false) {
    LaunchPlan pl;
    const int sms = sm_count();
    // one CTA per sample from half a wave of samples on: measured on a B200 (bench field, 10^4 points x 10 rays), forward /
    // backward ms with several CTAs per sample vs one: 64 samples 0.084 / 0.103 vs 0.079 / 0.097, 148: 0.124 / 0.160 vs
    // 0.086 / 0.102, 256: 0.208 / 0.293 vs 0.152 / 0.188 (profiles/r02_mode_threshold.txt)
    if (2 * n_local >= sms || force_large) {
        pl.threads = max_threads_large; pl.split = 1;
        pl.smem_bytes = AB200_WIN_KB * 1024 < max_window_bytes(blocking) ? AB200_WIN_KB * 1024 : max_window_bytes(blocking);
    } else {
        pl.threads = 512;
        const int want = (4 * sms + (n_local > 0 ? n_local : 1) - 1) / (n_local > 0 ? n_local : 1);
        const int max_split = (n_points + pl.threads - 1) / pl.threads;
        pl.split = want < 1 ? 1 : (want > max_split ? max_split : want);
        pl.smem_bytes = 100 * 1024;
    }
    pl.pts_per_chunk = (n_points + pl.split - 1) / pl.split;
    return pl;
}

static void fill_params(TraceParams& prm, const ab200_trace_args* a, const LaunchPlan& pl) {
    prm.a = *a;
    prm.split = pl.split;
    prm.pts_per_chunk = pl.pts_per_chunk;
    prm.win_cap = pl.smem_bytes / 4;
    const double k = (double)a->ray_magnitude * (double)a->one_minus_extinction * (double)a->reflectivity;
    const double imax = (k < 0 ? -k : k) * 1.01;   // |lambert cosine| <= 1 for unit vectors (+1% slack)
    const double rays = (double)a->n_points * (double)a->n_rays;
    if (imax > 0.0) {
        double s = (4294967295.0 - 2.0 * rays) / (rays * imax) * (1.0 - 1e-6);
        if (s * imax > 4194304.0) s = 4194304.0 / imax;   // one tap < 2^22: lets the packed loop round with a magic-number FMA
        prm.fx_scale = (float)s;
        prm.fx_inv = (float)((k < 0 ? -1.0 : 1.0) / (double)prm.fx_scale);
    } else {
        prm.fx_scale = 0.f;
        prm.fx_inv = 0.f;
    }
    prm.sigma = a->scatter_sigma > 0.f ? a->scatter_sigma : 2.5e-3f;
    prm.ident.one = 1.0f; prm.ident.negzero = -0.0f; prm.ident.negone = -1.0f;
    prm.self_zero = 0;
    prm.quad = 0;
    prm.wave = pl.split == 1 ? sm_count() : 0;
    prm.simple_counts = (a->ray_magnitude >= 1e-6f && a->one_minus_extinction >= 1e-6f && a->reflectivity >= 1e-6f &&
                         a->ray_magnitude <= 1e6f && a->one_minus_extinction <= 1e6f && a->reflectivity <= 1e6f) ? 1 : 0;
}

#ifndef AB200_TU_BWD
template <int THREADS, int TRIG>
static cudaError_t launch_fwd(const TraceParams& prm, const LaunchPlan& pl, cudaStream_t st, bool dbg, bool fp32acc) {
    const int grid = prm.a.n_local * pl.split;
#define AB200_LAUNCH_FWD(DBG, ACC, BLK)                                                                      \
    do {                                                                                                     \
        auto kern = trace_fwd_kernel<THREADS, TRIG, DBG, ACC, BLK>;                                          \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem_bytes); \
        if (e != cudaSuccess) return e;                                                                      \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
        if (e != cudaSuccess) return e;                                                                      \
        kern<<<grid, THREADS, pl.smem_bytes, st>>>(prm);                                                     \
        note_launch();                                                                                       \
    } while (0)
    if (prm.a.src_rows != nullptr) {   // activation map: production variants with the polynomial trig only
        if (dbg || fp32acc || TRIG != AB200_TRIG_POLY) return cudaErrorNotSupported;
        if (TRIG == AB200_TRIG_POLY) {   // (compile-time guard: no MAP instantiations for the other trig modes)
            auto launch_map = [&](auto kern) -> cudaError_t {
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem_bytes);
                if (e != cudaSuccess) return e;
                e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                if (e != cudaSuccess) return e;
                kern<<<grid, THREADS, pl.smem_bytes, st>>>(prm);
                note_launch();
                return cudaGetLastError();
            };
            constexpr int T2 = TRIG == AB200_TRIG_POLY ? TRIG : AB200_TRIG_POLY;
            return prm.a.blockers.n_blockers > 0 ? launch_map(trace_fwd_kernel<THREADS, T2, false, false, true, true>)
                                                 : launch_map(trace_fwd_kernel<THREADS, T2, false, false, false, true>);
        }
    }
    if (prm.a.blockers.n_blockers > 0) {   // blocking: production variant only
        if (dbg || fp32acc) return cudaErrorNotSupported;
        AB200_LAUNCH_FWD(false, false, true);
    } else if (dbg) { if (fp32acc) AB200_LAUNCH_FWD(true, true, false); else AB200_LAUNCH_FWD(true, false, false); }
    else     { if (fp32acc) AB200_LAUNCH_FWD(false, true, false); else AB200_LAUNCH_FWD(false, false, false); }
#undef AB200_LAUNCH_FWD
    return cudaGetLastError();
}

template <int THREADS>
static cudaError_t launch_fwd_trig(const TraceParams& prm, const LaunchPlan& pl, cudaStream_t st, bool dbg, bool fp32acc) {
    switch (prm.a.trig_mode) {
        case AB200_TRIG_TABLE: return launch_fwd<THREADS, AB200_TRIG_TABLE>(prm, pl, st, dbg, fp32acc);
        case AB200_TRIG_POLY: return launch_fwd<THREADS, AB200_TRIG_POLY>(prm, pl, st, dbg, fp32acc);
        default: return launch_fwd<THREADS, AB200_TRIG_SINCOSF>(prm, pl, st, dbg, fp32acc);
    }
}

#endif  // !AB200_TU_BWD

#ifndef AB200_TU_FWD
template <int THREADS, int TRIG>
static cudaError_t launch_bwd(const TraceParams& prm, const LaunchPlan& pl, cudaStream_t st, const float* gflux,
                              long long gstride, float* gpts, float* gnrm, float* gprims, float* gori, float* gscratch) {
    if (prm.a.src_rows != nullptr) {   // activation map: polynomial trig only (see launch_fwd)
        if (TRIG != AB200_TRIG_POLY) return cudaErrorNotSupported;
        constexpr int T2 = TRIG == AB200_TRIG_POLY ? TRIG : AB200_TRIG_POLY;
        auto launch_map = [&](auto kern) -> cudaError_t {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem_bytes);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (e != cudaSuccess) return e;
            kern<<<prm.a.n_local * pl.split, THREADS, pl.smem_bytes, st>>>(prm, gflux, gstride, gpts, gnrm, gprims, gori, gscratch);
            note_launch();
            return cudaGetLastError();
        };
        return prm.a.blockers.n_blockers > 0 ? launch_map(trace_bwd_kernel<THREADS, T2, true, true>)
                                             : launch_map(trace_bwd_kernel<THREADS, T2, false, true>);
    }
    if (prm.a.blockers.n_blockers > 0) {
        auto kern = trace_bwd_kernel<THREADS, TRIG, true>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem_bytes);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        kern<<<prm.a.n_local * pl.split, THREADS, pl.smem_bytes, st>>>(prm, gflux, gstride, gpts, gnrm, gprims, gori, gscratch);
    } else {
        auto kern = trace_bwd_kernel<THREADS, TRIG, false>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem_bytes);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        kern<<<prm.a.n_local * pl.split, THREADS, pl.smem_bytes, st>>>(prm, gflux, gstride, gpts, gnrm, gprims, gori, gscratch);
    }
    note_launch();
    return cudaGetLastError();
}

template <int THREADS>
static cudaError_t launch_bwd_trig(const TraceParams& prm, const LaunchPlan& pl, cudaStream_t st, const float* gflux,
                                   long long gstride, float* gpts, float* gnrm, float* gprims, float* gori, float* gscratch) {
    switch (prm.a.trig_mode) {
        case AB200_TRIG_TABLE: return launch_bwd<THREADS, AB200_TRIG_TABLE>(prm, pl, st, gflux, gstride, gpts, gnrm, gprims, gori, gscratch);
        case AB200_TRIG_POLY: return launch_bwd<THREADS, AB200_TRIG_POLY>(prm, pl, st, gflux, gstride, gpts, gnrm, gprims, gori, gscratch);
        default: return launch_bwd<THREADS, AB200_TRIG_SINCOSF>(prm, pl, st, gflux, gstride, gpts, gnrm, gprims, gori, gscratch);
    }
}

#endif  // !AB200_TU_FWD

#ifndef AB200_TU_BWD
// The v3 kernels (trace_v3.cuh) cover the benchmark-shaped case; everything else runs the general kernels above.
static bool v3_enabled() {
    static const bool on = [] { const char* v = getenv("AB200_TRACE_V3"); return !(v && v[0] == '0'); }();
    return on;
}
static bool v3_eligible(const ab200_trace_args* a, const TraceParams& prm, const LaunchPlan& pl, bool dbg, bool fp32acc) {
    return v3_enabled() && pl.split == 1 && !dbg && !fp32acc && a->blockers.n_blockers == 0 && a->src_rows == nullptr && a->trig_mode == AB200_TRIG_POLY &&
           a->n_rays >= 4 && (a->n_rays & 1) == 0 && (a->n_points & 1) == 0 && a->res_e % 4 == 0 && a->res_e <= v3::kMaxE &&
           a->distortions_planar != nullptr && reinterpret_cast<uintptr_t>(a->distortions_planar) % 8 == 0 &&
           reinterpret_cast<uintptr_t>(a->points) % 16 == 0 &&
           reinterpret_cast<uintptr_t>(a->normals) % 16 == 0 && prm.simple_counts && prm.fx_scale >= 1e-6f &&
           pl.smem_bytes / 4 / v3::kPitch >= 2;
}

static cudaError_t launch_fwd_v3(const TraceParams& prm, const LaunchPlan& pl, cudaStream_t st) {
    auto kern = trace_fwd_v3_kernel<v3::kFwdThreads>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    kern<<<prm.a.n_local, v3::kFwdThreads, pl.smem_bytes, st>>>(prm);
    note_launch();
    return cudaGetLastError();
}

#endif  // !AB200_TU_BWD

}  // namespace ab200

using namespace ab200;

#ifndef AB200_TU_BWD
extern "C" int32_t ab200_trace_fwd(const ab200_trace_args* a, void* stream) {
    int32_t rc = validate(a);
    if (rc != AB200_OK) return rc;
    AB200_REQUIRE(a->flux && a->intercept && a->on_target && a->blocking, AB200_EINVAL, "NULL output pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t ue = (size_t)a->res_u * a->res_e;
    // rows not traced by this rank stay zero; traced rows are accumulated into / overwritten
    const LaunchPlan pl = make_plan(a->n_local, a->n_points, kFwdThreadsLarge, (a->flags & AB200_FLAG_ONE_CTA_PER_SAMPLE) != 0, a->blockers.n_blockers > 0);
    const bool dbg = a->dbg_be || a->dbg_bu || a->dbg_t || a->dbg_lambert;
    const bool fp32acc = (a->flags & AB200_FLAG_FP32_ACCUM) != 0;
    // one CTA per sample and every sample traced here: each CTA clears its own bitmap row (no N*U*E memset pass)
    const bool self_zero = pl.split == 1 && a->n_local == a->n_samples && !fp32acc && a->n_samples > 0;
    if (!self_zero) {
        AB200_CUDA_TRY(cudaMemsetAsync(a->flux, 0, (size_t)a->n_samples * ue * sizeof(float), st));
        AB200_CUDA_TRY(cudaMemsetAsync(a->intercept, 0, (size_t)a->n_samples * sizeof(float), st));
        AB200_CUDA_TRY(cudaMemsetAsync(a->on_target, 0, (size_t)a->n_samples * sizeof(float), st));
        AB200_CUDA_TRY(cudaMemsetAsync(a->blocking, 0, (size_t)a->n_samples * sizeof(float), st));
    }
    if (a->n_local == 0 || a->n_samples == 0) return AB200_OK;
    TraceParams prm;
    fill_params(prm, a, pl);
    prm.self_zero = self_zero ? 1 : 0;
    prm.quad = (pl.split == 1 && !fp32acc && a->res_e % 4 == 0 && reinterpret_cast<uintptr_t>(a->flux) % 16 == 0) ? 1 : 0;
    cudaError_t e;
    if (v3_eligible(a, prm, pl, dbg, fp32acc) && prm.quad) {
        e = launch_fwd_v3(prm, pl, st);
    } else {
        e = (pl.threads == kFwdThreadsLarge) ? launch_fwd_trig<kFwdThreadsLarge>(prm, pl, st, dbg, fp32acc)
                                             : launch_fwd_trig<512>(prm, pl, st, dbg, fp32acc);
    }
    AB200_REQUIRE(e == cudaSuccess, AB200_ECUDA, "trace_fwd launch failed: %s", cudaGetErrorString(e));
    if (pl.split > 1) {
        if (!fp32acc) {
            dim3 grid((unsigned)((ue + 1023) / 1024 < 64 ? (ue + 1023) / 1024 : 64), (unsigned)a->n_local);
            finalize_split_kernel<<<grid, 256, 0, st>>>(prm);
            note_launch();
        } else {
            finalize_split_fp32_kernel<<<(a->n_local + 127) / 128, 128, 0, st>>>(prm);
            note_launch();
        }
        AB200_CUDA_TRY(cudaGetLastError());
    }
    return AB200_OK;
}

#endif  // !AB200_TU_BWD

#ifndef AB200_TU_FWD
extern "C" int32_t ab200_trace_bwd(const ab200_trace_bwd_args* b, void* stream) {
    AB200_REQUIRE(b != nullptr, AB200_EINVAL, "args is NULL");
    const ab200_trace_args* a = &b->fwd;
    int32_t rc = validate(a);
    if (rc != AB200_OK) return rc;
    AB200_REQUIRE(b->grad_flux, AB200_EINVAL, "NULL gradient pointer");
    AB200_REQUIRE((b->grad_points == nullptr) == (b->grad_normals == nullptr), AB200_EINVAL,
                  "grad_points and grad_normals must both be given or both be NULL");
    AB200_REQUIRE(b->grad_points || b->grad_orientations || b->grad_prims, AB200_EINVAL, "no gradient output requested");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t np4 = (size_t)a->n_samples * a->n_points * 4;
    if (a->n_local < a->n_samples && b->grad_points) {
        AB200_CUDA_TRY(cudaMemsetAsync(b->grad_points, 0, np4 * sizeof(float), st));
        AB200_CUDA_TRY(cudaMemsetAsync(b->grad_normals, 0, np4 * sizeof(float), st));
    }
    if (a->n_local == 0 || a->n_samples == 0) return AB200_OK;
    const LaunchPlan pl = make_plan(a->n_local, a->n_points, kBwdThreadsLarge, (a->flags & AB200_FLAG_ONE_CTA_PER_SAMPLE) != 0, a->blockers.n_blockers > 0);
    TraceParams prm;
    fill_params(prm, a, pl);
    const long long gstride = b->grad_flux_stride >= 0 ? b->grad_flux_stride : (long long)a->res_u * a->res_e;
    prm.quad = (pl.split == 1 && a->res_e % 4 == 0 && gstride % 4 == 0 && reinterpret_cast<uintptr_t>(b->grad_flux) % 16 == 0) ? 1 : 0;
    // blocker gradients: per-CTA rows in the caller's scratch + an ordered reduction, if the scratch is large enough
    const long long n_ctas = (long long)a->n_local * pl.split;
    const bool blk_grad = a->blockers.n_blockers > 0 && b->grad_prims != nullptr;
    float* scratch = (blk_grad && b->grad_prims_scratch &&
                      b->grad_prims_scratch_floats >= n_ctas * a->blockers.max_candidates * 12) ? b->grad_prims_scratch : nullptr;
    cudaError_t e = (pl.threads == kBwdThreadsLarge)
                        ? launch_bwd_trig<kBwdThreadsLarge>(prm, pl, st, b->grad_flux, gstride, b->grad_points, b->grad_normals, b->grad_prims, b->grad_orientations, scratch)
                        : launch_bwd_trig<512>(prm, pl, st, b->grad_flux, gstride, b->grad_points, b->grad_normals, b->grad_prims, b->grad_orientations, scratch);
    AB200_REQUIRE(e == cudaSuccess, AB200_ECUDA, "trace_bwd launch failed: %s", cudaGetErrorString(e));
    if (scratch) {
        const int warps_per_block = 8;
        const unsigned blocks = (unsigned)((a->blockers.n_blockers + warps_per_block - 1) / warps_per_block);
        blocker_grad_reduce_kernel<<<blocks, warps_per_block * 32, 0, st>>>(scratch, a->blockers.cand_idx, a->blockers.cand_count,
                                                                            a->local_rows, (int)n_ctas, pl.split,
                                                                            a->blockers.max_candidates, a->blockers.n_blockers,
                                                                            b->grad_prims);
        note_launch();
        AB200_CUDA_TRY(cudaGetLastError());
    }
    return AB200_OK;
}

#endif  // !AB200_TU_FWD

#ifndef AB200_TU_BWD
extern "C" int32_t ab200_abi_version(void) { return AB200_ABI_VERSION; }

extern "C" int64_t ab200_kernel_launch_count(void) { return ab200::g_kernel_launches.load(); }

extern "C" const char* ab200_error_string(int32_t code) {
    switch (code) {
        case AB200_OK: return "ok";
        case AB200_EINVAL: return "invalid argument";
        case AB200_ECUDA: return "CUDA failure";
        case AB200_ELIMIT: return "size limit exceeded";
        default: return "unknown error";
    }
}

extern "C" const char* ab200_last_error_detail(void) { return ab200::g_error_detail; }

namespace ab200 {
__global__ void debug_trig_kernel(const float* x, int n, int mode, float* s, float* c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float ss, cc;
    if (mode == AB200_TRIG_POLY) sincos_poly(x[i], &ss, &cc); else sincosf(x[i], &ss, &cc);
    s[i] = ss; c[i] = cc;
}
}  // namespace ab200

namespace ab200 {
__global__ void debug_const_div_kernel(const float* a, int n, float b, float* q_fast, float* q_ieee) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float rb = __frcp_rn(b);
    q_fast[i] = const_div(a[i], b, rb);
    q_ieee[i] = __fdiv_rn(a[i], b);
}
}  // namespace ab200

namespace ab200 {
__global__ void debug_div_regular_kernel(const float* a, const float* b, int n, float* q_fast, float* q_ieee) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    q_fast[i] = div_regular(a[i], b[i]);
    q_ieee[i] = __fdiv_rn(a[i], b[i]);
}
}  // namespace ab200

extern "C" int32_t ab200_debug_div_regular(const float* a, const float* b, int32_t n, float* q_fast, float* q_ieee, void* stream) {
    AB200_REQUIRE(a && b && q_fast && q_ieee && n >= 0, AB200_EINVAL, "bad arguments");
    if (n == 0) return AB200_OK;
    debug_div_regular_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, n, q_fast, q_ieee);
    note_launch();
    AB200_CUDA_TRY(cudaGetLastError());
    return AB200_OK;
}

extern "C" int32_t ab200_debug_const_div(const float* a, int32_t n, float b, float* q_fast, float* q_ieee, void* stream) {
    AB200_REQUIRE(a && q_fast && q_ieee && n >= 0, AB200_EINVAL, "bad arguments");
    if (n == 0) return AB200_OK;
    debug_const_div_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, n, b, q_fast, q_ieee);
    note_launch();
    AB200_CUDA_TRY(cudaGetLastError());
    return AB200_OK;
}

extern "C" int32_t ab200_debug_trig(const float* angles, int32_t n, int32_t mode, float* out_sin, float* out_cos, void* stream) {
    AB200_REQUIRE(angles && out_sin && out_cos && n >= 0, AB200_EINVAL, "bad arguments");
    if (n == 0) return AB200_OK;
    debug_trig_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(angles, n, mode, out_sin, out_cos);
    note_launch();
    AB200_CUDA_TRY(cudaGetLastError());
    return AB200_OK;
}
#endif  // !AB200_TU_BWD
