// Blocking support kernels: pack the rectangle primitives of all heliostats and build, per active heliostat-sample,
// the ordered list of primitives that can shadow its rays (conservative tapered-capsule test).  The per-ray soft mask
// itself lives in blocking_device.cuh and is evaluated inside the trace kernels.
// Reference: artist/raytracing/blocking.py:123-209 (primitives), :832-995 (LBVH filter - replaced by the candidate
// lists), :212-354 (soft mask).
#include "blocking_device.cuh"

namespace ab200 {

__global__ void blocking_pack_kernel(const float* __restrict__ corners, const float* __restrict__ spans,
                                     const float* __restrict__ normals, int n, float epsilon, float* __restrict__ prims) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    BlockPrimPacked p;
    for (int q = 0; q < 3; ++q) {
        p.c0[q] = corners[(size_t)i * 12 + q];
        p.su[q] = spans[(size_t)i * 6 + q];
        p.sv[q] = spans[(size_t)i * 6 + 3 + q];
        p.n[q] = normals[(size_t)i * 3 + q];
    }
    // blocking.py:333-341 in its operation order (each product and sum rounded, left to right): block_tuv_strict
    p.uu = dot3_strict(p.su[0], p.su[1], p.su[2], p.su[0], p.su[1], p.su[2]);
    p.vv = dot3_strict(p.sv[0], p.sv[1], p.sv[2], p.sv[0], p.sv[1], p.sv[2]);
    p.uv = dot3_strict(p.su[0], p.su[1], p.su[2], p.sv[0], p.sv[1], p.sv[2]);
    float det = ssub(smul(p.uu, p.vv), smul(p.uv, p.uv));
    if (fabsf(det) < epsilon) det = (det > 0.f ? 1.f : (det < 0.f ? -1.f : 0.f)) * epsilon;   // torch.sign(det) * eps
    p.det = det;
    reinterpret_cast<BlockPrimPacked*>(prims)[i] = p;
}

// one warp per sample; primitives are scanned in index order so the list (and the summation order of the optical
// depth) is deterministic
__global__ void __launch_bounds__(128) blocking_candidates_kernel(const float* __restrict__ prims, int n_prims,
                                                                  const int* __restrict__ sample_to_blocker,
                                                                  const float* __restrict__ aim, const float* __restrict__ target_radius,
                                                                  int n_samples, float spread_angle, int max_cand,
                                                                  int* __restrict__ cand_idx, int* __restrict__ cand_count,
                                                                  int* __restrict__ overflow) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_samples) return;
    const BlockPrimPacked* P = reinterpret_cast<const BlockPrimPacked*>(prims);
    const int self = sample_to_blocker[warp];
    const BlockPrimPacked me = P[self];
    float c[3], a[3], d[3];
    for (int q = 0; q < 3; ++q) {
        c[q] = me.c0[q] + 0.5f * (me.su[q] + me.sv[q]);
        a[q] = aim[(size_t)warp * 4 + q];
        d[q] = a[q] - c[q];
    }
    const float len2 = fmaxf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2], 1e-12f);
    const float len = sqrtf(len2);
    const float r_h = 0.5f * sqrtf(me.uu + me.vv + 2.0f * fabsf(me.uv)) + 0.1f;
    const float r_t = fmaxf(target_radius[warp], r_h) + spread_angle * len;
    int count = 0;
    for (int base = 0; base < n_prims; base += 32) {
        const int k = base + lane;
        bool hit = false;
        if (k < n_prims && k != self) {
            const BlockPrimPacked p = P[k];
            float m[3];
            for (int q = 0; q < 3; ++q) m[q] = p.c0[q] + 0.5f * (p.su[q] + p.sv[q]) - c[q];
            const float rho = 0.5f * sqrtf(p.uu + p.vv + 2.0f * fabsf(p.uv)) + 0.1f;
            const float s_raw = (m[0] * d[0] + m[1] * d[1] + m[2] * d[2]) / len2;
            const float s = fminf(fmaxf(s_raw, 0.0f), 1.0f);
            const float e0 = m[0] - s * d[0], e1 = m[1] - s * d[1], e2 = m[2] - s * d[2];
            const float dist = sqrtf(e0 * e0 + e1 * e1 + e2 * e2);
            hit = dist <= r_h + s * (r_t - r_h) + rho;
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            const int slot = count + __popc(ballot & ((1u << lane) - 1u));
            if (slot < max_cand) cand_idx[(size_t)warp * max_cand + slot] = k;
        }
        count += __popc(ballot);
    }
    if (lane == 0) {
        cand_count[warp] = count < max_cand ? count : max_cand;
        if (count > max_cand) atomicAdd(overflow, 1);
    }
}

}  // namespace ab200

using namespace ab200;

extern "C" int32_t ab200_blocking_pack(const float* corners, const float* spans, const float* normals, int32_t n_prims,
                                       float epsilon, float* prims, void* stream) {
    AB200_REQUIRE(corners && spans && normals && prims && n_prims >= 0, AB200_EINVAL, "bad arguments");
    if (n_prims == 0) return AB200_OK;
    blocking_pack_kernel<<<(n_prims + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(corners, spans, normals, n_prims,
                                                                                              epsilon, prims);
    note_launch();
    AB200_CUDA_TRY(cudaGetLastError());
    return AB200_OK;
}

extern "C" int32_t ab200_blocking_candidates(const float* prims, int32_t n_prims, const int32_t* sample_to_blocker,
                                             const float* aim_points, const float* target_radius, int32_t n_samples,
                                             float spread_angle, int32_t max_candidates, int32_t* cand_idx,
                                             int32_t* cand_count, int32_t* overflow, void* stream) {
    AB200_REQUIRE(prims && sample_to_blocker && aim_points && target_radius && cand_idx && cand_count && overflow, AB200_EINVAL,
                  "NULL pointer");
    AB200_REQUIRE(max_candidates >= 1 && max_candidates <= kMaxBlockCandidates, AB200_EINVAL, "max_candidates must be 1..%d",
                  kMaxBlockCandidates);
    if (n_samples == 0) return AB200_OK;
    const int warps_per_block = 4;
    blocking_candidates_kernel<<<(n_samples + warps_per_block - 1) / warps_per_block, 32 * warps_per_block, 0,
                                 static_cast<cudaStream_t>(stream)>>>(prims, n_prims, sample_to_blocker, aim_points, target_radius,
                                                                      n_samples, spread_angle, max_candidates, cand_idx, cand_count,
                                                                      overflow);
    note_launch();
    AB200_CUDA_TRY(cudaGetLastError());
    return AB200_OK;
}
