// Sun-shape distortion sampling (a8): the [N,R,P,2] buffer that `Sun.get_distortions` (artist/scene/sun.py:199-234)
// draws with `MultivariateNormal(mean, cov * I).sample((N, R, P))` on a CUDA device, written by ONE kernel.
//
// What the reference's eager path does on CUDA (torch 2.11): `torch.empty(shape).normal_()` - a grid-stride Philox4x32-10
// kernel, four Box-Muller normals per counter - followed by `loc + scale_tril @ eps`, which torch dispatches as one cuBLAS
// 2x2 gemv per 65535 rays (3125 launches, 233 ms at the benchmark size).  With a diagonal `scale_tril` that product is
// exactly `fl(sigma * eps)` per element (the second term of every dot product is an exact zero), so the whole sample is
//     out[li] = mean + fl(sigma * z[li]),   z = torch's normal_ stream for (seed, offset)
// and this kernel reproduces it bit for bit: same counter layout (element li <- thread (li mod T), iteration li / (4T),
// component (li / T) mod 4, T = torch's grid size x 256), same Philox rounds, same Box-Muller arithmetic
// (cuRAND's `_curand_box_muller`: fused `x * 2^-32 + 2^-33`, libdevice logf, IEEE sqrt, MUFU sin/cos).
// tests/test_gpu_sampling.py compares with `torch.distributions.MultivariateNormal.sample` element for element.
#include <cstdint>
#include "common.cuh"

namespace ab200 {

__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// cuRAND's Box-Muller (curand_normal.h:70-87), operation for operation
__device__ __forceinline__ float2 box_muller(uint32_t x, uint32_t y) {
    constexpr float k2Pow32Inv = 2.3283064e-10f;
    constexpr float k2Pow32Inv2Pi = 2.3283064e-10f * 6.2831855f;
    const float u = __fmaf_rn(__uint2float_rn(x), k2Pow32Inv, k2Pow32Inv / 2);
    const float v = __fmaf_rn(__uint2float_rn(y), k2Pow32Inv2Pi, k2Pow32Inv2Pi / 2);
    const float s = sqrtf(__fmul_rn(-2.0f, logf(u)));
    float sn, cs;
    __sincosf(v, &sn, &cs);
    return make_float2(__fmul_rn(sn, s), __fmul_rn(cs, s));
}

// One thread per Philox counter (torch thread `idx`, torch loop iteration `k`): four normals to four elements T apart.
// `sigma[c]`, `mean[c]` belong to component c = li & 1 of the (u, e) pair (the buffer's last dimension has size 2 and
// every launch starts on an even element).
__global__ void __launch_bounds__(256)
sample_distortions_kernel(float* __restrict__ out, long long numel, long long T, unsigned long long seed,
                          unsigned long long offset4 /* philox offset / 4 */, float sigma_u, float sigma_e, float mean_u,
                          float mean_e, long long n_work, float* __restrict__ planar, long long planar_first, long long plane_stride) {
    // T is a multiple of the block size: one (uniform, 32-bit) division per block
    const unsigned blocks_per_iter = (unsigned)(T >> 8);
    const unsigned kb = blockIdx.x / blocks_per_iter;
    const long long k = kb, idx = (long long)(blockIdx.x - kb * blocks_per_iter) * 256 + threadIdx.x;
    (void)n_work;
    const unsigned long long ctr_lo = offset4 + (unsigned long long)k;
    const uint4 r = philox4x32_10((uint32_t)ctr_lo, (uint32_t)(ctr_lo >> 32), (uint32_t)idx, (uint32_t)((unsigned long long)idx >> 32),
                                  (uint32_t)seed, (uint32_t)(seed >> 32));
    const float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
    const float z[4] = {a.x, a.y, b.x, b.y};
    const long long base = k * 4 * T + idx;
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
        const long long li = base + (long long)ii * T;
        if (li < numel) {
            const bool e = (li & 1) != 0;
            // torch: normal_ stores fl(z * 1 + 0) = z; then loc + fl(sigma * z)
            const float v = __fadd_rn(e ? mean_e : mean_u, __fmul_rn(e ? sigma_e : sigma_u, z[ii]));
            __stcs(out + li, v);
            if (planar) {   // global element index g = planar_first + li: pair g / 2 of plane g & 1
                const long long g = planar_first + li;
                __stcs(planar + (g & 1) * plane_stride + (g >> 1), v);
            }
        }
    }
}

}  // namespace ab200

using namespace ab200;

// torch's `calc_execution_policy` (ATen/native/cuda/DistributionTemplates.h:50-62) for one launch of `numel` elements
static void torch_policy(long long numel, int sms, int max_threads_per_sm, long long* threads_total, unsigned long long* counter_offset) {
    const long long block = 256;
    long long grid = (numel + block - 1) / block;
    const long long cap = (long long)sms * (max_threads_per_sm / block);
    if (grid > cap) grid = cap;
    *threads_total = grid * block;
    *counter_offset = (unsigned long long)((numel - 1) / (block * grid * 4) + 1) * 4ull;
}

// TensorIterator's 32-bit-indexing split (`with_32bit_indexing`): a 1-D fp32 tensor is launched whole if numel and its last
// byte offset fit int32, else its first half (numel / 2), then the rest, recursively - each launch drawing its own
// Philox offset from the generator in that order.
static int32_t sample_range(float* out, long long numel, unsigned long long seed, unsigned long long* offset, int sms,
                            int max_threads_per_sm, const float* sg, const float* mn, long long global_start, cudaStream_t st,
                            float* planar, long long plane_stride) {
    if (numel <= 0) return AB200_OK;
    if (numel > 536870912ll) {
        const long long first = numel / 2;
        int32_t rc = sample_range(out, first, seed, offset, sms, max_threads_per_sm, sg, mn, global_start, st, planar, plane_stride);
        if (rc != AB200_OK) return rc;
        return sample_range(out + first, numel - first, seed, offset, sms, max_threads_per_sm, sg, mn, global_start + first, st,
                            planar, plane_stride);
    }
    long long T;
    unsigned long long inc;
    torch_policy(numel, sms, max_threads_per_sm, &T, &inc);
    const long long iters = (numel - 1) / (4 * T) + 1, n_work = iters * T;
    const long long blocks = (n_work + 255) / 256;
    AB200_REQUIRE(blocks < (1ll << 31), AB200_ELIMIT, "too many sampling blocks");
    // a sub-range that starts on an odd element swaps the (u, e) roles of even/odd local indices
    const bool odd = (global_start & 1) != 0;
    sample_distortions_kernel<<<(unsigned)blocks, 256, 0, st>>>(out, numel, T, seed, *offset / 4, odd ? sg[1] : sg[0],
                                                                odd ? sg[0] : sg[1], odd ? mn[1] : mn[0], odd ? mn[0] : mn[1], n_work,
                                                                planar, global_start, plane_stride);
    note_launch();
    AB200_CUDA_TRY(cudaGetLastError());
    *offset += inc;
    return AB200_OK;
}

extern "C" int32_t ab200_sample_distortions(float* out, int64_t n_pairs, uint64_t seed, uint64_t philox_offset, float sigma_u,
                                            float sigma_e, float mean_u, float mean_e, int32_t sm_count_override,
                                            int32_t max_threads_per_sm_override, uint64_t* philox_offset_after, float* out_planar,
                                            void* stream) {
    AB200_REQUIRE(out != nullptr && n_pairs >= 0, AB200_EINVAL, "bad arguments");
    AB200_REQUIRE(philox_offset % 4 == 0, AB200_EINVAL, "philox_offset must be a multiple of 4 (torch's generator always is)");
    int dev = 0, sms = 148, mtps = 2048;
    AB200_CUDA_TRY(cudaGetDevice(&dev));
    AB200_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    AB200_CUDA_TRY(cudaDeviceGetAttribute(&mtps, cudaDevAttrMaxThreadsPerMultiProcessor, dev));
    if (sm_count_override > 0) sms = sm_count_override;
    if (max_threads_per_sm_override > 0) mtps = max_threads_per_sm_override;
    const float sg[2] = {sigma_u, sigma_e}, mn[2] = {mean_u, mean_e};
    unsigned long long off = philox_offset;
    const long long numel_all = (long long)n_pairs * 2;
    if (numel_all > 536870912ll) {
        // torch draws the generator offset for the WHOLE tensor before it notices that the tensor needs splitting
        // (distribution_nullary_kernel: philox_cuda_state() precedes the can_use_32bit_indexing() test): that increment
        // is consumed and never used
        long long T;
        unsigned long long inc;
        torch_policy(numel_all, sms, mtps, &T, &inc);
        off += inc;
    }
    const int32_t rc = sample_range(out, (long long)n_pairs * 2, seed, &off, sms, mtps, sg, mn, 0, static_cast<cudaStream_t>(stream),
                                    out_planar, (long long)n_pairs);
    if (philox_offset_after) *philox_offset_after = off;
    return rc;
}

namespace ab200 {
__global__ void __launch_bounds__(256) deinterleave_kernel(const float2* __restrict__ in, long long n_pairs, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n_pairs) return;
    const float2 v = __ldcs(in + i);
    out[i] = v.x;
    out[n_pairs + i] = v.y;
}
}  // namespace ab200

extern "C" int32_t ab200_deinterleave_distortions(const float* interleaved, int64_t n_pairs, float* out_planar, void* stream) {
    AB200_REQUIRE(interleaved && out_planar && n_pairs >= 0, AB200_EINVAL, "bad arguments");
    AB200_REQUIRE(reinterpret_cast<uintptr_t>(interleaved) % 8 == 0, AB200_EINVAL, "interleaved distortions must be 8-byte aligned");
    if (n_pairs == 0) return AB200_OK;
    const long long blocks = (n_pairs + 255) / 256;
    AB200_REQUIRE(blocks < (1ll << 31), AB200_ELIMIT, "too many pairs");
    deinterleave_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float2*>(interleaved),
                                                                                       n_pairs, out_planar);
    note_launch();
    AB200_CUDA_TRY(cudaGetLastError());
    return AB200_OK;
}
