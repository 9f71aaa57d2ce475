// Step after the ray tracer: centre of mass and centre-of-mass crop of flux bitmaps
// (artist/flux/bitmap.py:12-55 get_center_of_mass, :121-246 crop_flux_distributions_around_center), forward and
// backward.  The reference expresses the crop with affine_grid + grid_sample (bilinear, align_corners=True, zero
// padding) after two full-bitmap reductions; here one launch computes the moments of every bitmap and one launch
// resamples it, and the backward is a deterministic gather (no atomics): d/d(bitmap) = transpose of the bilinear
// resampling + the path through the centre of mass.
#include <climits>
#include <cstdlib>
#include "common.cuh"

namespace ab200 {

constexpr float kMassEps = 1e-8f;   // bitmap.py:41,163

// block-wide sum of three values in a fixed order (warp shuffles, then warp partials in index order)
__device__ __forceinline__ void block_sum3(float& a, float& b, float& c, float* red /* [3*32] */) {
    a = warp_sumf(a); b = warp_sumf(b); c = warp_sumf(c);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) { red[warp] = a; red[32 + warp] = b; red[64 + warp] = c; }
    __syncthreads();
    float x = 0.f, y = 0.f, z = 0.f;
    for (int w = 0; w < nw; ++w) { x += red[w]; y += red[32 + w]; z += red[64 + w]; }
    a = x; b = y; c = z;
}

// moments[b] = (S, sum_q x_q v / (S + eps), sum_p y_p v / (S + eps)) with pixel-index coordinates (normalised == 0:
// x_q = q, get_center_of_mass) or normalised ones (normalised == 1: linspace(-1, 1), the crop's centre)
__global__ void __launch_bounds__(256) flux_moments_kernel(const float* __restrict__ bitmaps, int res_u, int res_e, int normalised,
                                                           float* __restrict__ moments) {
    __shared__ float red[96];
    const int b = blockIdx.x;
    const float* v = bitmaps + (size_t)b * res_u * res_e;
    const float sx = normalised ? 2.0f / (float)(res_e - 1) : 1.0f, ox = normalised ? -1.0f : 0.0f;
    const float sy = normalised ? 2.0f / (float)(res_u - 1) : 1.0f, oy = normalised ? -1.0f : 0.0f;
    float s = 0.f, mx = 0.f, my = 0.f;
    for (int p = threadIdx.x >> 5; p < res_u; p += blockDim.x >> 5) {          // one warp per row: coalesced
        const float y = fmaf((float)p, sy, oy);
        float rs = 0.f, rx = 0.f;
        for (int q = threadIdx.x & 31; q < res_e; q += 32) {
            const float val = __ldg(v + (size_t)p * res_e + q);
            rs += val;
            rx = fmaf(fmaf((float)q, sx, ox), val, rx);
        }
        s += rs; mx += rx; my = fmaf(y, rs, my);
    }
    block_sum3(s, mx, my, red);
    if (threadIdx.x == 0) {
        const float inv = 1.0f / (s + kMassEps);
        moments[3 * b] = s; moments[3 * b + 1] = mx * inv; moments[3 * b + 2] = my * inv;
    }
}

// backward of the two centre-of-mass outputs: d c / d v[p,q] = (coord - c) / (S + eps)
__global__ void __launch_bounds__(256) flux_moments_bwd_kernel(const float* __restrict__ moments, const float* __restrict__ grad_centre,
                                                               int res_u, int res_e, int normalised, float* __restrict__ grad_in) {
    const int b = blockIdx.y;
    const float sx = normalised ? 2.0f / (float)(res_e - 1) : 1.0f, ox = normalised ? -1.0f : 0.0f;
    const float sy = normalised ? 2.0f / (float)(res_u - 1) : 1.0f, oy = normalised ? -1.0f : 0.0f;
    const float inv_mass = 1.0f / (moments[3 * b] + kMassEps), cx = moments[3 * b + 1], cy = moments[3 * b + 2];
    const float gx = grad_centre[2 * b], gy = grad_centre[2 * b + 1];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < res_u * res_e; k += gridDim.x * blockDim.x) {
        const int p = k / res_e, q = k - p * res_e;
        grad_in[(size_t)b * res_u * res_e + k] = (gx * (fmaf((float)q, sx, ox) - cx) + gy * (fmaf((float)p, sy, oy) - cy)) * inv_mass;
    }
}

struct CropMap {   // source pixel coordinate of output column j / row i:  ix = ax * j + bx,  iy = ay * i + by
    float ax, bx, ay, by;
};
__device__ __forceinline__ CropMap crop_map(float scale_x, float scale_y, float cx, float cy, int res_u, int res_e) {
    // grid x = scale_x * (-1 + 2 j / (E-1)) + cx  ->  ix = (x + 1) / 2 * (E-1)   (affine_grid / grid_sample, align_corners)
    CropMap m;
    m.ax = scale_x; m.bx = (cx + 1.0f - scale_x) * 0.5f * (float)(res_e - 1);
    m.ay = scale_y; m.by = (cy + 1.0f - scale_y) * 0.5f * (float)(res_u - 1);
    return m;
}

__device__ __forceinline__ float tap(const float* v, int p, int q, int res_u, int res_e) {
    return (p >= 0 && p < res_u && q >= 0 && q < res_e) ? __ldg(v + (size_t)p * res_e + q) : 0.0f;
}

__global__ void __launch_bounds__(256) flux_crop_fwd_kernel(const float* __restrict__ bitmaps, const float* __restrict__ scale,
                                                            const float* __restrict__ moments, int res_u, int res_e,
                                                            float* __restrict__ out) {
    const int b = blockIdx.y;
    const float* v = bitmaps + (size_t)b * res_u * res_e;
    const CropMap m = crop_map(scale[2 * b], scale[2 * b + 1], moments[3 * b + 1], moments[3 * b + 2], res_u, res_e);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < res_u * res_e; k += gridDim.x * blockDim.x) {
        const int i = k / res_e, j = k - i * res_e;
        const float ix = fmaf(m.ax, (float)j, m.bx), iy = fmaf(m.ay, (float)i, m.by);
        const float fx = floorf(ix), fy = floorf(iy);
        const int q0 = (int)fx, p0 = (int)fy;
        const float wx = ix - fx, wy = iy - fy;
        const float top = fmaf(wx, tap(v, p0, q0 + 1, res_u, res_e), (1.0f - wx) * tap(v, p0, q0, res_u, res_e));
        const float bot = fmaf(wx, tap(v, p0 + 1, q0 + 1, res_u, res_e), (1.0f - wx) * tap(v, p0 + 1, q0, res_u, res_e));
        out[(size_t)b * res_u * res_e + k] = fmaf(wy, bot, (1.0f - wy) * top);
    }
}

// g_centre[b] = (dL/dcx, dL/dcy): the resampling grid moves rigidly with the centre of mass
__global__ void __launch_bounds__(256) flux_crop_centre_grad_kernel(const float* __restrict__ bitmaps, const float* __restrict__ scale,
                                                                    const float* __restrict__ moments,
                                                                    const float* __restrict__ grad_out, int res_u, int res_e,
                                                                    float* __restrict__ g_centre) {
    __shared__ float red[96];
    const int b = blockIdx.x;
    const float* v = bitmaps + (size_t)b * res_u * res_e;
    const float* g = grad_out + (size_t)b * res_u * res_e;
    const CropMap m = crop_map(scale[2 * b], scale[2 * b + 1], moments[3 * b + 1], moments[3 * b + 2], res_u, res_e);
    float gx = 0.f, gy = 0.f, unused = 0.f;
    for (int k = threadIdx.x; k < res_u * res_e; k += blockDim.x) {
        const int i = k / res_e, j = k - i * res_e;
        const float ix = fmaf(m.ax, (float)j, m.bx), iy = fmaf(m.ay, (float)i, m.by);
        const float fx = floorf(ix), fy = floorf(iy);
        const int q0 = (int)fx, p0 = (int)fy;
        const float wx = ix - fx, wy = iy - fy;
        const float v00 = tap(v, p0, q0, res_u, res_e), v01 = tap(v, p0, q0 + 1, res_u, res_e);
        const float v10 = tap(v, p0 + 1, q0, res_u, res_e), v11 = tap(v, p0 + 1, q0 + 1, res_u, res_e);
        const float go = __ldg(g + k);
        gx = fmaf(go, fmaf(wy, v11 - v10, (1.0f - wy) * (v01 - v00)), gx);     // d out / d ix
        gy = fmaf(go, fmaf(wx, v11 - v01, (1.0f - wx) * (v10 - v00)), gy);     // d out / d iy
    }
    block_sum3(gx, gy, unused, red);
    if (threadIdx.x == 0) {
        g_centre[2 * b] = gx * 0.5f * (float)(res_e - 1);      // d ix / d cx
        g_centre[2 * b + 1] = gy * 0.5f * (float)(res_u - 1);
    }
}

// grad_in[p,q] = sum_{i,j} grad_out[i,j] hat(iy_i - p) hat(ix_j - q)  +  (g_cx (x_q - cx) + g_cy (y_p - cy)) / (S + eps)
__device__ __forceinline__ float crop_bwd_pixel(const float* __restrict__ g, const CropMap& m, int p, int q, int res_u, int res_e,
                                                float gcx, float gcy, float cx, float cy, float inv_mass) {
    float acc = 0.f;
    if (m.ax > 0.f && m.ay > 0.f) {
        // output columns j with |ax j + bx - q| < 1, rows i with |ay i + by - p| < 1
        const int j_lo = max(0, (int)ceilf(((float)q - 1.0f - m.bx) / m.ax)), j_hi = min(res_e - 1, (int)floorf(((float)q + 1.0f - m.bx) / m.ax));
        const int i_lo = max(0, (int)ceilf(((float)p - 1.0f - m.by) / m.ay)), i_hi = min(res_u - 1, (int)floorf(((float)p + 1.0f - m.by) / m.ay));
        for (int i = i_lo; i <= i_hi; ++i) {
            const float hy = fmaxf(0.0f, 1.0f - fabsf(fmaf(m.ay, (float)i, m.by) - (float)p));
            if (hy == 0.0f) continue;
            float row = 0.f;
            for (int j = j_lo; j <= j_hi; ++j) {
                const float hx = fmaxf(0.0f, 1.0f - fabsf(fmaf(m.ax, (float)j, m.bx) - (float)q));
                row = fmaf(hx, __ldg(g + (size_t)i * res_e + j), row);
            }
            acc = fmaf(hy, row, acc);
        }
    }
    const float xq = fmaf((float)q, 2.0f / (float)(res_e - 1), -1.0f), yp = fmaf((float)p, 2.0f / (float)(res_u - 1), -1.0f);
    return acc + (gcx * (xq - cx) + gcy * (yp - cy)) * inv_mass;
}

__global__ void __launch_bounds__(256) flux_crop_bwd_kernel(const float* __restrict__ scale, const float* __restrict__ moments,
                                                            const float* __restrict__ g_centre, const float* __restrict__ grad_out,
                                                            int res_u, int res_e, float* __restrict__ grad_in) {
    const int b = blockIdx.y;
    const float* g = grad_out + (size_t)b * res_u * res_e;
    const float S = moments[3 * b], cx = moments[3 * b + 1], cy = moments[3 * b + 2];
    const CropMap m = crop_map(scale[2 * b], scale[2 * b + 1], cx, cy, res_u, res_e);
    const float inv_mass = 1.0f / (S + kMassEps);
    const float gcx = g_centre[2 * b], gcy = g_centre[2 * b + 1];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < res_u * res_e; k += gridDim.x * blockDim.x) {
        const int p = k / res_e, q = k - p * res_e;
        grad_in[(size_t)b * res_u * res_e + k] = crop_bwd_pixel(g, m, p, q, res_u, res_e, gcx, gcy, cx, cy, inv_mass);
    }
}

// ---------------------------------------------------------------------------------------------
// Column-marching forms of the three crop kernels (round 2).  The resampling map is separable - the source column of an
// output pixel depends on its column only, the source row on its row only - so one thread owns a COLUMN and walks the
// rows: everything per column (source column, horizontal weights, bounds) is computed once per thread, everything per row is
// uniform over the warp, neighbouring rows share their source rows (kept in registers while the walk advances), a warp's
// accesses are 32 consecutive floats, and no pixel pays an integer division.  Measured at [2048,256,256] on a B200 (per-pixel
// kernels above -> these): forward 0.56 -> see profiles/r02_flux_epilogue.txt.  Same formulas per output value as the
// per-pixel kernels (the forward is bit-identical, the reductions sum in a different fixed order).
// One CTA per bitmap, blockDim.x threads walk columns tid, tid + blockDim.x, ...
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ldz(const float* row, int q, int res_e) { return (q >= 0 && q < res_e) ? __ldg(row + q) : 0.0f; }

// Per-row half of the separable map, the same for every column: source row p0 and weight wy of output row i, computed once
// per CTA into shared memory (kMaxRowTable rows; taller bitmaps compute it per pixel).
constexpr int kMaxRowTable = 1024;
struct RowMap { int p0; float wy; };
__device__ __forceinline__ RowMap row_map(const CropMap& m, int i) {
    const float iy = fmaf(m.ay, (float)i, m.by), fy = floorf(iy);
    return RowMap{(int)fy, iy - fy};
}
__device__ __forceinline__ void fill_row_table(RowMap* tab, const CropMap& m, int res_u) {
    for (int i = threadIdx.x; i < min(res_u, kMaxRowTable); i += blockDim.x) tab[i] = row_map(m, i);
    __syncthreads();
}
__device__ __forceinline__ RowMap row_of(const RowMap* tab, const CropMap& m, int i) { return i < kMaxRowTable ? tab[i] : row_map(m, i); }

__global__ void __launch_bounds__(256) flux_crop_fwd_cols_kernel(const float* __restrict__ bitmaps, const float* __restrict__ scale,
                                                                 const float* __restrict__ moments, int res_u, int res_e,
                                                                 int rows_per_cta, float* __restrict__ out) {
    __shared__ RowMap tab[kMaxRowTable];
    const int b = blockIdx.y;
    const float* v = bitmaps + (size_t)b * res_u * res_e;
    float* o = out + (size_t)b * res_u * res_e;
    const CropMap m = crop_map(scale[2 * b], scale[2 * b + 1], moments[3 * b + 1], moments[3 * b + 2], res_u, res_e);
    fill_row_table(tab, m, res_u);
    const int i_begin = blockIdx.x * rows_per_cta, i_end = min(res_u, i_begin + rows_per_cta);
    for (int j = threadIdx.x; j < res_e; j += blockDim.x) {
        const float ix = fmaf(m.ax, (float)j, m.bx), fx = floorf(ix), wx = ix - fx;
        const int q0 = (int)fx;
        const bool in0 = q0 >= 0 && q0 < res_e, in1 = q0 + 1 >= 0 && q0 + 1 < res_e;
        // horizontally interpolated source rows: h(p) = wx v[p][q0+1] + (1 - wx) v[p][q0]  (0 outside the bitmap)
        auto hrow = [&](int p) -> float {
            if (p < 0 || p >= res_u) return 0.0f;
            const float* row = v + (size_t)p * res_e + q0;
            return fmaf(wx, in1 ? __ldg(row + 1) : 0.0f, (1.0f - wx) * (in0 ? __ldg(row) : 0.0f));
        };
        int p_have = INT_MIN;   // source row held in `top` (bot = the next one)
        float top = 0.f, bot = 0.f;
        float* oc = o + j;
        for (int i = i_begin; i < i_end; ++i) {
            const RowMap r = row_of(tab, m, i);
            if (r.p0 != p_have) {
                top = (r.p0 == p_have + 1) ? bot : hrow(r.p0);
                bot = hrow(r.p0 + 1);
                p_have = r.p0;
            }
            oc[(size_t)i * res_e] = fmaf(r.wy, bot, (1.0f - r.wy) * top);
        }
    }
}

__global__ void __launch_bounds__(256) flux_crop_centre_grad_cols_kernel(const float* __restrict__ bitmaps, const float* __restrict__ scale,
                                                                         const float* __restrict__ moments,
                                                                         const float* __restrict__ grad_out, int res_u, int res_e,
                                                                         float* __restrict__ g_centre) {
    __shared__ float red[96];
    __shared__ RowMap tab[kMaxRowTable];
    const int b = blockIdx.x;
    const float* v = bitmaps + (size_t)b * res_u * res_e;
    const float* g = grad_out + (size_t)b * res_u * res_e;
    const CropMap m = crop_map(scale[2 * b], scale[2 * b + 1], moments[3 * b + 1], moments[3 * b + 2], res_u, res_e);
    fill_row_table(tab, m, res_u);
    float gx = 0.f, gy = 0.f, unused = 0.f;
    for (int j = threadIdx.x; j < res_e; j += blockDim.x) {
        const float ix = fmaf(m.ax, (float)j, m.bx), fx = floorf(ix), wx = ix - fx;
        const int q0 = (int)fx;
        const bool in0 = q0 >= 0 && q0 < res_e, in1 = q0 + 1 >= 0 && q0 + 1 < res_e;
        int p_have = INT_MIN;
        float t0 = 0.f, t1 = 0.f, b0 = 0.f, b1 = 0.f;   // v[p0][q0], v[p0][q0+1], v[p0+1][q0], v[p0+1][q0+1]
        auto load_row = [&](int p, float& a0, float& a1) {
            if (p < 0 || p >= res_u) { a0 = 0.f; a1 = 0.f; return; }
            const float* row = v + (size_t)p * res_e + q0;
            a0 = in0 ? __ldg(row) : 0.0f; a1 = in1 ? __ldg(row + 1) : 0.0f;
        };
        const float* gc = g + j;
        for (int i = 0; i < res_u; ++i) {
            const RowMap r = row_of(tab, m, i);
            if (r.p0 != p_have) {
                if (r.p0 == p_have + 1) { t0 = b0; t1 = b1; } else load_row(r.p0, t0, t1);
                load_row(r.p0 + 1, b0, b1);
                p_have = r.p0;
            }
            const float go = __ldg(gc + (size_t)i * res_e);
            gx = fmaf(go, fmaf(r.wy, b1 - b0, (1.0f - r.wy) * (t1 - t0)), gx);     // d out / d ix
            gy = fmaf(go, fmaf(wx, b1 - t1, (1.0f - wx) * (b0 - t0)), gy);         // d out / d iy
        }
    }
    block_sum3(gx, gy, unused, red);
    if (threadIdx.x == 0) {
        g_centre[2 * b] = gx * 0.5f * (float)(res_e - 1);      // d ix / d cx
        g_centre[2 * b + 1] = gy * 0.5f * (float)(res_u - 1);
    }
}

// Transposed resampling as a per-column march over the OUTPUT rows: output row i adds (1 - wy_i) R_i to input row p0_i and
// wy_i R_i to row p0_i + 1, with R_i = sum_j hat(ix_j - q) grad_out[i, j] over the <= kCropTaps output columns j whose source
// column is within one pixel of q.  p0_i never decreases, so two running accumulators per thread suffice and every input
// pixel is written exactly once, by its own thread: deterministic, no atomics.  Needs scale_x >= 0.4 (2 / scale_x + 3 <=
// kCropTaps), else the per-pixel gather runs for that bitmap.
constexpr int kCropTaps = 8;
__global__ void __launch_bounds__(256) flux_crop_bwd_cols_kernel(const float* __restrict__ scale, const float* __restrict__ moments,
                                                                 const float* __restrict__ g_centre, const float* __restrict__ grad_out,
                                                                 int res_u, int res_e, float* __restrict__ grad_in) {
    __shared__ RowMap tab[kMaxRowTable];
    const int b = blockIdx.x;
    const float* g = grad_out + (size_t)b * res_u * res_e;
    float* o = grad_in + (size_t)b * res_u * res_e;
    const float S = moments[3 * b], cx = moments[3 * b + 1], cy = moments[3 * b + 2];
    const CropMap m = crop_map(scale[2 * b], scale[2 * b + 1], cx, cy, res_u, res_e);
    const float inv_mass = 1.0f / (S + kMassEps);
    const float gcx = g_centre[2 * b], gcy = g_centre[2 * b + 1];
    if (!(m.ay > 0.f && m.ax > 0.f && 2.0f / m.ax + 3.0f <= (float)kCropTaps)) {   // strong magnification: per-pixel gather
        for (int k = threadIdx.x; k < res_u * res_e; k += blockDim.x) {
            const int p = k / res_e, q = k - p * res_e;
            o[k] = crop_bwd_pixel(g, m, p, q, res_u, res_e, gcx, gcy, cx, cy, inv_mass);
        }
        return;
    }
    fill_row_table(tab, m, res_u);
    const float inv_ax = 1.0f / m.ax;
    const float ky = 2.0f / (float)(res_u - 1), gyk = gcy * inv_mass;
    for (int q = threadIdx.x; q < res_e; q += blockDim.x) {
        // output columns j with |ax j + bx - q| < 1 (one extra on either side against the rounding of the reciprocal; a
        // column that does not contribute gets weight 0), compacted to the first contributing one
        int j_lo = max(0, (int)ceilf(((float)q - 1.0f - m.bx) * inv_ax) - 1);
        const int j_hi = min(res_e - 1, (int)floorf(((float)q + 1.0f - m.bx) * inv_ax) + 1);
        auto hat = [&](int j) { return (j <= j_hi) ? fmaxf(0.0f, 1.0f - fabsf(fmaf(m.ax, (float)j, m.bx) - (float)q)) : 0.0f; };
        while (j_lo <= j_hi && hat(j_lo) == 0.0f) ++j_lo;
        float hx[kCropTaps];
        int n_taps = 0;
#pragma unroll
        for (int t = 0; t < kCropTaps; ++t) {
            hx[t] = hat(j_lo + t);
            if (hx[t] != 0.0f) n_taps = t + 1;
        }
        // centre-of-mass term of input pixel (p, q): c0 + p * c1
        const float c0 = (gcx * (fmaf((float)q, 2.0f / (float)(res_e - 1), -1.0f) - cx) + gcy * (-1.0f - cy)) * inv_mass, c1 = gyk * ky;
        const float* gq = g + j_lo;
        float* oq = o + q;
        float acc = 0.f, acc_next = 0.f;
        int p = -1;    // input row `acc` belongs to (row -1 catches output rows that start above the bitmap)
        for (int i = 0; i < res_u; ++i) {
            const RowMap rm = row_of(tab, m, i);
            if (rm.p0 >= res_u) break;
            if (rm.p0 + 1 < 0) continue;
            while (p < rm.p0) {   // rows up to p0 - 1 are complete
                if (p >= 0) oq[(size_t)p * res_e] = acc + fmaf((float)p, c1, c0);
                acc = acc_next; acc_next = 0.f; ++p;
            }
            const float* row = gq + (size_t)i * res_e;
            float r = 0.f;
#pragma unroll
            for (int t = 0; t < kCropTaps; ++t) {
                if (t >= n_taps) break;
                r = fmaf(hx[t], __ldg(row + t), r);
            }
            acc = fmaf(1.0f - rm.wy, r, acc);
            acc_next = fmaf(rm.wy, r, acc_next);
        }
        while (p < res_u) {
            if (p >= 0) oq[(size_t)p * res_e] = acc + fmaf((float)p, c1, c0);
            acc = acc_next; acc_next = 0.f; ++p;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Losses on flux bitmaps (artist/optim/loss.py): PixelLoss :251-319 and KLDivergenceLoss :322-410, reduced over the
// whole bitmap (reduction_dimensions = (1, 2), what every caller passes).  One CTA per sample; the bitmap pair is
// read once from HBM (the second pass of the KL loss hits L1/L2), the reference re-reads both [N,U,E] tensors once per
// eager op (normalise, add, log, KLDiv, sum: ~10 passes).
//   kind 0: loss = sum (p - g)^2 / sum g
//   kind 1: P = g / max(sum |g|, 1e-12), Q = p / max(sum |p|, 1e-12), lP = log(P + 1e-12), lQ = log(Q + 1e-12),
//           loss = sum exp(lP) * (lP - lQ)                       (KLDivLoss(reduction="none", log_target=True))
// aux[b] = (sum g | max(sum|p|,eps), max(sum|g|,eps), sum_x A_x Q_x with A_x = -exp(lP_x) / (Q_x + eps), sum|p| > eps)
// ---------------------------------------------------------------------------------------------
constexpr float kLossEps = 1e-12f;

// 1024 threads and (through a dynamic shared-memory request it does not use) ONE resident CTA per SM: the 148 bitmap pairs in
// flight (512 KB each at 256 x 256) then fit the 126 MB L2 and the KL loss's second pass never goes back to HBM (measured:
// 2.1 GB -> 1.1 GB of DRAM reads per launch at [2048,256,256]).  16-byte loads where the row length allows.
template <bool VEC>
__global__ void __launch_bounds__(1024, 1) flux_loss_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int n_px,
                                                                int kind, float* __restrict__ loss, float* __restrict__ aux) {
    __shared__ float red[96];
    const int b = blockIdx.x;
    const float* p = pred + (size_t)b * n_px;
    const float* g = gt + (size_t)b * n_px;
    const int n_it = VEC ? n_px / 4 : n_px;
    auto load = [&](const float* base, int k, float (&v)[4]) {
        if (VEC) { const float4 t = __ldg(reinterpret_cast<const float4*>(base) + k); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
        else { v[0] = __ldg(base + k); v[1] = v[2] = v[3] = 0.f; }
    };
    constexpr int W = VEC ? 4 : 1;
    if (kind == 0) {
        float se = 0.f, sg = 0.f, unused = 0.f;
        for (int k = threadIdx.x; k < n_it; k += blockDim.x) {
            float pv[4], gv[4];
            load(p, k, pv); load(g, k, gv);
#pragma unroll
            for (int c = 0; c < W; ++c) { const float d = pv[c] - gv[c]; se = fmaf(d, d, se); sg += gv[c]; }
        }
        block_sum3(se, sg, unused, red);
        if (threadIdx.x == 0) { loss[b] = se / sg; aux[4 * b] = sg; aux[4 * b + 1] = 0.f; aux[4 * b + 2] = 0.f; aux[4 * b + 3] = 0.f; }
        return;
    }
    // (the normalisers too: every Q = p / dp inherits their rounding)
    __shared__ double red_d[64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    auto block_sum2d = [&](double& x, double& y) {
        for (int off = 16; off > 0; off >>= 1) { x += __shfl_xor_sync(0xffffffffu, x, off); y += __shfl_xor_sync(0xffffffffu, y, off); }
        __syncthreads();
        if (lane == 0) { red_d[warp] = x; red_d[32 + warp] = y; }
        __syncthreads();
        double a = 0.0, c = 0.0;
        for (int w = 0; w < nw; ++w) { a += red_d[w]; c += red_d[32 + w]; }
        x = a; y = c;
    };
    double sp_d = 0.0, sg_d = 0.0;
    for (int k = threadIdx.x; k < n_it; k += blockDim.x) {
        float pv[4], gv[4];
        load(p, k, pv); load(g, k, gv);
        float a4 = 0.f, c4 = 0.f;
#pragma unroll
        for (int c = 0; c < W; ++c) { a4 += fabsf(pv[c]); c4 += fabsf(gv[c]); }
        sp_d += (double)a4; sg_d += (double)c4;
    }
    block_sum2d(sp_d, sg_d);
    const float sp = (float)sp_d, sg = (float)sg_d;
    const float dp = fmaxf(sp, kLossEps), dg = fmaxf(sg, kLossEps);
    // the KL terms have both signs (the sum cancels when prediction ~ ground truth): the per-thread sums run in double, so
    // the result does not depend on how the pixels are dealt to the threads (the reference's pairwise fp32 sum is good to ~1e-7)
    double kl_d = 0.0, saq_d = 0.0;
    for (int k = threadIdx.x; k < n_it; k += blockDim.x) {
        float pv[4], gv[4];
        load(p, k, pv); load(g, k, gv);
        float kl4 = 0.f, saq4 = 0.f;
#pragma unroll
        for (int c = 0; c < W; ++c) {
            const float P = gv[c] / dg, Q = pv[c] / dp;
            const float lP = logf(P + kLossEps), lQ = logf(Q + kLossEps);
            const float eP = expf(lP);
            kl4 = fmaf(eP, lP - lQ, kl4);
            saq4 = fmaf(-eP / (Q + kLossEps), Q, saq4);
        }
        kl_d += (double)kl4; saq_d += (double)saq4;
    }
    block_sum2d(kl_d, saq_d);   // shuffles, then the warps' partials in index order
    if (threadIdx.x == 0) {
        const float kl = (float)kl_d, saq = (float)saq_d;
        loss[b] = kl;
        aux[4 * b] = dp; aux[4 * b + 1] = dg; aux[4 * b + 2] = saq; aux[4 * b + 3] = sp > kLossEps ? 1.0f : 0.0f;
    }
}

// d loss[b] / d prediction[b, :, :] * grad_loss[b]
template <bool VEC>
__global__ void __launch_bounds__(256) flux_loss_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                            const float* __restrict__ aux, const float* __restrict__ grad_loss,
                                                            int n_px, int kind, float* __restrict__ grad_pred) {
    const int b = blockIdx.y;
    const float* p = pred + (size_t)b * n_px;
    const float* g = gt + (size_t)b * n_px;
    float* o = grad_pred + (size_t)b * n_px;
    const float gl = grad_loss[b];
    constexpr int W = VEC ? 4 : 1;
    const int n_it = VEC ? n_px / 4 : n_px;
    const float k2 = kind == 0 ? 2.0f * gl / aux[4 * b] : 0.f;
    const float dp = aux[4 * b], dg = aux[4 * b + 1], saq = aux[4 * b + 2], through_norm = aux[4 * b + 3];
    const float inv_dp = gl / dp;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_it; k += gridDim.x * blockDim.x) {
        float pv[4], gv[4], ov[4];
        if (VEC) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p) + k), c = __ldg(reinterpret_cast<const float4*>(g) + k);
            pv[0] = a.x; pv[1] = a.y; pv[2] = a.z; pv[3] = a.w; gv[0] = c.x; gv[1] = c.y; gv[2] = c.z; gv[3] = c.w;
        } else { pv[0] = __ldg(p + k); gv[0] = __ldg(g + k); }
#pragma unroll
        for (int c = 0; c < W; ++c) {
            if (kind == 0) { ov[c] = k2 * (pv[c] - gv[c]); continue; }
            const float P = gv[c] / dg, Q = pv[c] / dp;
            const float A = -expf(logf(P + kLossEps)) / (Q + kLossEps);
            const float sgn = pv[c] > 0.f ? 1.0f : (pv[c] < 0.f ? -1.0f : 0.0f);
            ov[c] = inv_dp * (A - through_norm * sgn * saq);
        }
        if (VEC) reinterpret_cast<float4*>(o)[k] = make_float4(ov[0], ov[1], ov[2], ov[3]);
        else o[k] = ov[0];
    }
}

}  // namespace ab200

using namespace ab200;

extern "C" int32_t ab200_flux_moments(const float* bitmaps, int32_t n_bitmaps, int32_t res_u, int32_t res_e, int32_t normalised,
                                      float* moments, void* stream) {
    AB200_REQUIRE(bitmaps && moments, AB200_EINVAL, "NULL pointer");
    AB200_REQUIRE(n_bitmaps >= 0 && res_u >= 2 && res_e >= 2, AB200_EINVAL, "bad sizes");
    if (n_bitmaps == 0) return AB200_OK;
    flux_moments_kernel<<<n_bitmaps, 256, 0, static_cast<cudaStream_t>(stream)>>>(bitmaps, res_u, res_e, normalised, moments);
    note_launch();
    AB200_CUDA_TRY(cudaGetLastError());
    return AB200_OK;
}

extern "C" int32_t ab200_flux_moments_bwd(const float* moments, const float* grad_centre, int32_t n_bitmaps, int32_t res_u,
                                          int32_t res_e, int32_t normalised, float* grad_bitmaps, void* stream) {
    AB200_REQUIRE(moments && grad_centre && grad_bitmaps, AB200_EINVAL, "NULL pointer");
    AB200_REQUIRE(n_bitmaps >= 0 && n_bitmaps <= 65535 && res_u >= 2 && res_e >= 2, AB200_EINVAL, "bad sizes");
    if (n_bitmaps == 0) return AB200_OK;
    dim3 grid((unsigned)((res_u * res_e + 1023) / 1024), (unsigned)n_bitmaps);
    flux_moments_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(moments, grad_centre, res_u, res_e, normalised,
                                                                                 grad_bitmaps);
    note_launch();
    AB200_CUDA_TRY(cudaGetLastError());
    return AB200_OK;
}

extern "C" int32_t ab200_flux_crop_fwd(const float* bitmaps, const float* scale, int32_t n_bitmaps, int32_t res_u, int32_t res_e,
                                       float* moments, float* out, void* stream) {
    AB200_REQUIRE(bitmaps && scale && moments && out, AB200_EINVAL, "NULL pointer");
    AB200_REQUIRE(n_bitmaps >= 0 && n_bitmaps <= 65535 && res_u >= 2 && res_e >= 2, AB200_EINVAL, "bad sizes");
    if (n_bitmaps == 0) return AB200_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    flux_moments_kernel<<<n_bitmaps, 256, 0, st>>>(bitmaps, res_u, res_e, 1, moments);
    if (getenv("AB200_FLUX_PER_PIXEL")) {   // (the round-1 kernels, kept for the parity tests)
        dim3 grid((unsigned)((res_u * res_e + 1023) / 1024), (unsigned)n_bitmaps);
        flux_crop_fwd_kernel<<<grid, 256, 0, st>>>(bitmaps, scale, moments, res_u, res_e, out);
    } else {
        const int chunks = res_u >= 128 ? 4 : 1, rows = (res_u + chunks - 1) / chunks;
        flux_crop_fwd_cols_kernel<<<dim3((unsigned)chunks, (unsigned)n_bitmaps), 256, 0, st>>>(bitmaps, scale, moments, res_u, res_e, rows, out);
    }
    note_launch(2);
    AB200_CUDA_TRY(cudaGetLastError());
    return AB200_OK;
}

extern "C" int32_t ab200_flux_crop_bwd(const float* bitmaps, const float* scale, const float* moments, const float* grad_out,
                                       int32_t n_bitmaps, int32_t res_u, int32_t res_e, float* scratch, float* grad_in,
                                       void* stream) {
    AB200_REQUIRE(bitmaps && scale && moments && grad_out && scratch && grad_in, AB200_EINVAL, "NULL pointer");
    AB200_REQUIRE(n_bitmaps >= 0 && n_bitmaps <= 65535 && res_u >= 2 && res_e >= 2, AB200_EINVAL, "bad sizes");
    if (n_bitmaps == 0) return AB200_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (getenv("AB200_FLUX_PER_PIXEL")) {
        flux_crop_centre_grad_kernel<<<n_bitmaps, 256, 0, st>>>(bitmaps, scale, moments, grad_out, res_u, res_e, scratch);
        dim3 grid((unsigned)((res_u * res_e + 1023) / 1024), (unsigned)n_bitmaps);
        flux_crop_bwd_kernel<<<grid, 256, 0, st>>>(scale, moments, scratch, grad_out, res_u, res_e, grad_in);
    } else {
        flux_crop_centre_grad_cols_kernel<<<n_bitmaps, 256, 0, st>>>(bitmaps, scale, moments, grad_out, res_u, res_e, scratch);
        // (the column march falls back to the per-pixel gather per bitmap for strong magnifications, scale_x < 0.4)
        flux_crop_bwd_cols_kernel<<<n_bitmaps, 256, 0, st>>>(scale, moments, scratch, grad_out, res_u, res_e, grad_in);
    }
    note_launch(2);
    AB200_CUDA_TRY(cudaGetLastError());
    return AB200_OK;
}

extern "C" int32_t ab200_flux_loss_fwd(const float* prediction, const float* ground_truth, int32_t n_bitmaps, int32_t res_u,
                                       int32_t res_e, int32_t kind, float* loss, float* aux, void* stream) {
    AB200_REQUIRE(prediction && ground_truth && loss && aux, AB200_EINVAL, "NULL pointer");
    AB200_REQUIRE(n_bitmaps >= 0 && res_u >= 1 && res_e >= 1, AB200_EINVAL, "bad sizes");
    AB200_REQUIRE(kind == AB200_LOSS_PIXEL || kind == AB200_LOSS_KL_DIVERGENCE, AB200_EINVAL, "unknown loss kind %d", kind);
    if (n_bitmaps == 0) return AB200_OK;
    const int n_px = res_u * res_e;
    const bool vec = n_px % 4 == 0 && reinterpret_cast<uintptr_t>(prediction) % 16 == 0 && reinterpret_cast<uintptr_t>(ground_truth) % 16 == 0;
    // one resident CTA per SM (see the kernel): ask for more than half of the SM's shared memory
    static const int kSoloSmem = 120 * 1024;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (vec) {
        auto kern = flux_loss_fwd_kernel<true>;
        AB200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSoloSmem));
        kern<<<n_bitmaps, 1024, kSoloSmem, st>>>(prediction, ground_truth, n_px, kind, loss, aux);
    } else {
        auto kern = flux_loss_fwd_kernel<false>;
        AB200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSoloSmem));
        kern<<<n_bitmaps, 1024, kSoloSmem, st>>>(prediction, ground_truth, n_px, kind, loss, aux);
    }
    note_launch();
    AB200_CUDA_TRY(cudaGetLastError());
    return AB200_OK;
}

extern "C" int32_t ab200_flux_loss_bwd(const float* prediction, const float* ground_truth, const float* aux, const float* grad_loss,
                                       int32_t n_bitmaps, int32_t res_u, int32_t res_e, int32_t kind, float* grad_prediction,
                                       void* stream) {
    AB200_REQUIRE(prediction && ground_truth && aux && grad_loss && grad_prediction, AB200_EINVAL, "NULL pointer");
    AB200_REQUIRE(n_bitmaps >= 0 && n_bitmaps <= 65535 && res_u >= 1 && res_e >= 1, AB200_EINVAL, "bad sizes");
    AB200_REQUIRE(kind == AB200_LOSS_PIXEL || kind == AB200_LOSS_KL_DIVERGENCE, AB200_EINVAL, "unknown loss kind %d", kind);
    if (n_bitmaps == 0) return AB200_OK;
    const int n_px = res_u * res_e;
    const bool vec = n_px % 4 == 0 && reinterpret_cast<uintptr_t>(prediction) % 16 == 0 && reinterpret_cast<uintptr_t>(ground_truth) % 16 == 0 &&
                     reinterpret_cast<uintptr_t>(grad_prediction) % 16 == 0;
    dim3 grid((unsigned)((n_px + 1023) / 1024), (unsigned)n_bitmaps);
    if (vec)
        flux_loss_bwd_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(prediction, ground_truth, aux, grad_loss, n_px, kind, grad_prediction);
    else
        flux_loss_bwd_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(prediction, ground_truth, aux, grad_loss, n_px, kind, grad_prediction);
    note_launch();
    AB200_CUDA_TRY(cudaGetLastError());
    return AB200_OK;
}
