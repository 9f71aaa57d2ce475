// Step after the ray tracer: centre of mass and centre-of-mass crop of flux bitmaps
// (artist/flux/bitmap.py:12-55 get_center_of_mass, :121-246 crop_flux_distributions_around_center), forward and
// backward.  The reference expresses the crop with affine_grid + grid_sample (bilinear, align_corners=True, zero
// padding) after two full-bitmap reductions; here one launch computes the moments of every bitmap and one launch
// resamples it, and the backward is a deterministic gather (no atomics): d/d(bitmap) = transpose of the bilinear
// resampling + the path through the centre of mass.
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
        grad_in[(size_t)b * res_u * res_e + k] = acc + (gcx * (xq - cx) + gcy * (yp - cy)) * inv_mass;
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

__global__ void __launch_bounds__(256) flux_loss_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int n_px,
                                                            int kind, float* __restrict__ loss, float* __restrict__ aux) {
    __shared__ float red[96];
    const int b = blockIdx.x;
    const float* p = pred + (size_t)b * n_px;
    const float* g = gt + (size_t)b * n_px;
    if (kind == 0) {
        float se = 0.f, sg = 0.f, unused = 0.f;
        for (int k = threadIdx.x; k < n_px; k += blockDim.x) {
            const float pv = __ldg(p + k), gv = __ldg(g + k), d = pv - gv;
            se = fmaf(d, d, se); sg += gv;
        }
        block_sum3(se, sg, unused, red);
        if (threadIdx.x == 0) { loss[b] = se / sg; aux[4 * b] = sg; aux[4 * b + 1] = 0.f; aux[4 * b + 2] = 0.f; aux[4 * b + 3] = 0.f; }
        return;
    }
    float sp = 0.f, sg = 0.f, unused = 0.f;
    for (int k = threadIdx.x; k < n_px; k += blockDim.x) { sp += fabsf(__ldg(p + k)); sg += fabsf(__ldg(g + k)); }
    block_sum3(sp, sg, unused, red);
    const float dp = fmaxf(sp, kLossEps), dg = fmaxf(sg, kLossEps);
    float kl = 0.f, saq = 0.f;
    for (int k = threadIdx.x; k < n_px; k += blockDim.x) {
        const float P = __ldg(g + k) / dg, Q = __ldg(p + k) / dp;
        const float lP = logf(P + kLossEps), lQ = logf(Q + kLossEps);
        const float eP = expf(lP);
        kl = fmaf(eP, lP - lQ, kl);
        saq = fmaf(-eP / (Q + kLossEps), Q, saq);
    }
    unused = 0.f;
    block_sum3(kl, saq, unused, red);
    if (threadIdx.x == 0) {
        loss[b] = kl;
        aux[4 * b] = dp; aux[4 * b + 1] = dg; aux[4 * b + 2] = saq; aux[4 * b + 3] = sp > kLossEps ? 1.0f : 0.0f;
    }
}

// d loss[b] / d prediction[b, :, :] * grad_loss[b]
__global__ void __launch_bounds__(256) flux_loss_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                            const float* __restrict__ aux, const float* __restrict__ grad_loss,
                                                            int n_px, int kind, float* __restrict__ grad_pred) {
    const int b = blockIdx.y;
    const float* p = pred + (size_t)b * n_px;
    const float* g = gt + (size_t)b * n_px;
    float* o = grad_pred + (size_t)b * n_px;
    const float gl = grad_loss[b];
    if (kind == 0) {
        const float k2 = 2.0f * gl / aux[4 * b];
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_px; k += gridDim.x * blockDim.x) o[k] = k2 * (__ldg(p + k) - __ldg(g + k));
        return;
    }
    const float dp = aux[4 * b], dg = aux[4 * b + 1], saq = aux[4 * b + 2], through_norm = aux[4 * b + 3];
    const float inv_dp = gl / dp;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_px; k += gridDim.x * blockDim.x) {
        const float pv = __ldg(p + k);
        const float P = __ldg(g + k) / dg, Q = pv / dp;
        const float A = -expf(logf(P + kLossEps)) / (Q + kLossEps);
        const float sgn = pv > 0.f ? 1.0f : (pv < 0.f ? -1.0f : 0.0f);
        o[k] = inv_dp * (A - through_norm * sgn * saq);
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
    dim3 grid((unsigned)((res_u * res_e + 1023) / 1024), (unsigned)n_bitmaps);
    flux_crop_fwd_kernel<<<grid, 256, 0, st>>>(bitmaps, scale, moments, res_u, res_e, out);
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
    flux_crop_centre_grad_kernel<<<n_bitmaps, 256, 0, st>>>(bitmaps, scale, moments, grad_out, res_u, res_e, scratch);
    dim3 grid((unsigned)((res_u * res_e + 1023) / 1024), (unsigned)n_bitmaps);
    flux_crop_bwd_kernel<<<grid, 256, 0, st>>>(scale, moments, scratch, grad_out, res_u, res_e, grad_in);
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
    flux_loss_fwd_kernel<<<n_bitmaps, 256, 0, static_cast<cudaStream_t>(stream)>>>(prediction, ground_truth, res_u * res_e, kind, loss, aux);
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
    dim3 grid((unsigned)((res_u * res_e + 1023) / 1024), (unsigned)n_bitmaps);
    flux_loss_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(prediction, ground_truth, aux, grad_loss, res_u * res_e, kind,
                                                                            grad_prediction);
    note_launch();
    AB200_CUDA_TRY(cudaGetLastError());
    return AB200_OK;
}
