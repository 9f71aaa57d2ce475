// Stand-alone forms of the ray-geometry steps the reference exports as free functions / methods
// (artist/raytracing/__init__.py:1-12): reflect (geometry.py:11-41), line_plane_intersections (:44-204),
// line_cylinder_intersections (:207-445) and HeliostatRayTracer.bilinear_splatting (heliostat_ray_tracer.py:610-778).
// They take and return the MATERIALISED per-ray tensors of the reference API ([B,R,P,...]) and exist for API parity and
// for step-wise checks against the reference's unit tests; trace_rays never goes through them (the fused kernels of
// trace.cu keep all of this in registers).  The per-ray math is the same device code the fused kernels inline
// (trace_device.cuh), so coordinates are bit-identical to the fused path and to the reference's eager ops.
#include "trace_device.cuh"

namespace ab200 {

__global__ void __launch_bounds__(256) reflect_kernel(const float4* __restrict__ incident, const float4* __restrict__ normals,
                                                      int n_points, long long total, float4* __restrict__ out) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const float4 i = __ldg(incident + k / n_points), n = __ldg(normals + k);
        // torch.sum over the 4 components adds left to right (checked against torch-CPU)
        const float dot = sadd(sadd(sadd(smul(i.x, n.x), smul(i.y, n.y)), smul(i.z, n.z)), smul(i.w, n.w));
        const float two_dot = smul(2.0f, dot);
        out[k] = make_float4(ssub(i.x, smul(two_dot, n.x)), ssub(i.y, smul(two_dot, n.y)), ssub(i.z, smul(two_dot, n.z)),
                             ssub(i.w, smul(two_dot, n.w)));
    }
}

// scatter (a9): d = M(e,u) r per ray, M from geometry/transforms.py:52-83 - the strict product/sum order of the fused
// kernels (trace_device.cuh::scatter); trig = the kernels' polynomial / libdevice sincos (<= 1 ulp)
__global__ void __launch_bounds__(256) scatter_kernel(const float* __restrict__ dist_u, const float* __restrict__ dist_e,
                                                      const float4* __restrict__ reflected, int n_rays, int n_points,
                                                      long long total, float4* __restrict__ out) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(k % n_points);
        const long long b = k / ((long long)n_rays * n_points);
        const float4 r = __ldg(reflected + b * n_points + p);
        Scatter s;
        sincos_poly(__ldg(dist_u + k), &s.su, &s.cu);
        sincos_poly(__ldg(dist_e + k), &s.se, &s.ce);
        PointCtx pc;
        pc.r0 = r.x; pc.r1 = r.y; pc.r2 = r.z;
        scatter(s, pc);
        out[k] = make_float4(s.dx, s.dy, s.dz, r.w);
    }
}

// one CTA row (blockIdx.y) per sample: the target constants are per sample
template <bool PLANAR>
__global__ void __launch_bounds__(256) intersections_kernel(const float4* __restrict__ dirs, const float* __restrict__ magnitudes,
                                                            const float4* __restrict__ origins, const ab200_targets targets,
                                                            const int* __restrict__ target_idx, int first_index, int n_rays,
                                                            int n_points, int res_e, int res_u, float* __restrict__ be,
                                                            float* __restrict__ bu, float* __restrict__ dist,
                                                            float* __restrict__ inten) {
    __shared__ TargetCtx T_sh;
    const int b = blockIdx.y;
    if (threadIdx.x == 0) load_target(T_sh, targets, first_index + (target_idx ? target_idx[b] : 0), res_e, res_u);
    __syncthreads();
    const TargetCtx T = T_sh;
    const long long per_sample = (long long)n_rays * n_points;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < per_sample; q += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(q % n_points);
        const long long k = (long long)b * per_sample + q;
        PointCtx pc;
        make_origin(pc, T, __ldg(origins + (long long)b * n_points + p));
        const float4 d = __ldg(dirs + k);
        Scatter s;
        s.dx = d.x; s.dy = d.y; s.dz = d.z;
        Hit h;
        if (PLANAR) hit_planar<false>(h, T, pc, s, __ldg(magnitudes + k)); else hit_cylinder<false>(h, T, pc, s, __ldg(magnitudes + k));
        be[k] = h.be; bu[k] = h.bu; dist[k] = h.t; inten[k] = h.lam;
    }
}

// out[b, U-1-iu(-1), ie(+1)] += weight * intensity (fp32 atomics; `out` zeroed by the entry point)
__global__ void __launch_bounds__(256) splat_kernel(const float* __restrict__ be, const float* __restrict__ bu,
                                                    const float* __restrict__ inten, long long per_sample, int res_e, int res_u,
                                                    float* __restrict__ out) {
    const int b = blockIdx.y;
    float* o = out + (size_t)b * res_u * res_e;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < per_sample; q += (long long)gridDim.x * blockDim.x) {
        const long long k = (long long)b * per_sample + q;
        const float e = be[k], u = bu[k], v = inten[k];
        const float fe = truncf(e), fu = truncf(u);                    // tensor.long() truncates towards zero
        if (!(fe >= 0.0f && fu >= 0.0f && fe + 1.0f < (float)res_e && fu + 1.0f < (float)res_u)) continue;
        const int ie = (int)fe, iu = (int)fu;
        const float wle = ssub(sadd(fe, 1.0f), e), wlu = ssub(sadd(fu, 1.0f), u), whe = ssub(e, fe), whu = ssub(u, fu);
        float* row_hi = o + (size_t)(res_u - 1 - (iu + 1)) * res_e + ie;   // final row flip (:778)
        float* row_lo = row_hi + res_e;
        atomicAdd(row_hi, smul(smul(wle, whu), v)); atomicAdd(row_hi + 1, smul(smul(whe, whu), v));
        atomicAdd(row_lo + 1, smul(smul(whe, wlu), v)); atomicAdd(row_lo, smul(smul(wle, wlu), v));
    }
}

static unsigned blocks_for(long long items) {
    const long long b = (items + 255) / 256;
    return (unsigned)(b < 1 ? 1 : (b > 1184 ? 1184 : b));   // 8 x 148 CTAs of 256 threads keep every SM busy
}
}  // namespace ab200

using namespace ab200;

extern "C" int32_t ab200_reflect(const float* incident, const float* normals, int32_t n_samples, int32_t n_points, float* out,
                                 void* stream) {
    AB200_REQUIRE(incident && normals && out, AB200_EINVAL, "NULL pointer");
    AB200_REQUIRE(n_samples >= 0 && n_points > 0, AB200_EINVAL, "bad sizes");
    if (n_samples == 0) return AB200_OK;
    const long long total = (long long)n_samples * n_points;
    reflect_kernel<<<blocks_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(incident), reinterpret_cast<const float4*>(normals), n_points, total,
        reinterpret_cast<float4*>(out));
    note_launch();
    AB200_CUDA_TRY(cudaGetLastError());
    return AB200_OK;
}

extern "C" int32_t ab200_scatter_rays(const float* distortions_u, const float* distortions_e, const float* reflected,
                                      int32_t n_samples, int32_t n_rays, int32_t n_points, float* out, void* stream) {
    AB200_REQUIRE(distortions_u && distortions_e && reflected && out, AB200_EINVAL, "NULL pointer");
    AB200_REQUIRE(n_samples >= 0 && n_rays > 0 && n_points > 0, AB200_EINVAL, "bad sizes");
    if (n_samples == 0) return AB200_OK;
    const long long total = (long long)n_samples * n_rays * n_points;
    scatter_kernel<<<blocks_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        distortions_u, distortions_e, reinterpret_cast<const float4*>(reflected), n_rays, n_points, total,
        reinterpret_cast<float4*>(out));
    note_launch();
    AB200_CUDA_TRY(cudaGetLastError());
    return AB200_OK;
}

extern "C" int32_t ab200_line_intersections(const float* ray_directions, const float* ray_magnitudes, const float* ray_origins,
                                            const ab200_targets* targets, const int32_t* target_idx, int32_t cylindrical,
                                            int32_t n_samples, int32_t n_rays, int32_t n_points, int32_t res_e, int32_t res_u,
                                            float* be, float* bu, float* distances, float* intensities, void* stream) {
    AB200_REQUIRE(ray_directions && ray_magnitudes && ray_origins && targets && be && bu && distances && intensities,
                  AB200_EINVAL, "NULL pointer");
    AB200_REQUIRE(n_samples >= 0 && n_rays > 0 && n_points > 0 && res_e >= 2 && res_u >= 2, AB200_EINVAL, "bad sizes");
    AB200_REQUIRE(cylindrical ? targets->n_cyl > 0 : targets->n_planar > 0, AB200_EINVAL, "no target area of the requested type");
    AB200_REQUIRE(n_samples <= 65535, AB200_ELIMIT, "more than 65535 samples per call");
    if (n_samples == 0) return AB200_OK;
    const long long per_sample = (long long)n_rays * n_points;
    dim3 grid(blocks_for(per_sample) / 8 + 1, (unsigned)n_samples);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float4* d4 = reinterpret_cast<const float4*>(ray_directions);
    const float4* o4 = reinterpret_cast<const float4*>(ray_origins);
    if (cylindrical)
        intersections_kernel<false><<<grid, 256, 0, st>>>(d4, ray_magnitudes, o4, *targets, target_idx, targets->n_planar, n_rays,
                                                          n_points, res_e, res_u, be, bu, distances, intensities);
    else
        intersections_kernel<true><<<grid, 256, 0, st>>>(d4, ray_magnitudes, o4, *targets, target_idx, 0, n_rays, n_points, res_e,
                                                         res_u, be, bu, distances, intensities);
    note_launch();
    AB200_CUDA_TRY(cudaGetLastError());
    return AB200_OK;
}

extern "C" int32_t ab200_bilinear_splatting(const float* be, const float* bu, const float* intensities, int32_t n_samples,
                                            int64_t rays_per_sample, int32_t res_e, int32_t res_u, float* out, void* stream) {
    AB200_REQUIRE(be && bu && intensities && out, AB200_EINVAL, "NULL pointer");
    AB200_REQUIRE(n_samples >= 0 && rays_per_sample >= 0 && res_e >= 2 && res_u >= 2, AB200_EINVAL, "bad sizes");
    AB200_REQUIRE(n_samples <= 65535, AB200_ELIMIT, "more than 65535 samples per call");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    AB200_CUDA_TRY(cudaMemsetAsync(out, 0, (size_t)n_samples * res_u * res_e * sizeof(float), st));
    if (n_samples == 0 || rays_per_sample == 0) return AB200_OK;
    dim3 grid(blocks_for(rays_per_sample) / 8 + 1, (unsigned)n_samples);
    splat_kernel<<<grid, 256, 0, st>>>(be, bu, intensities, rays_per_sample, res_e, res_u, out);
    note_launch();
    AB200_CUDA_TRY(cudaGetLastError());
    return AB200_OK;
}

// ---------------------------------------------------------------------------------------------
// activation index map (HeliostatGroup.activate_heliostats, artist/field/heliostat_group.py:256-315): the reference
// replicates every per-heliostat tensor with repeat_interleave(mask); here the trace kernels read surface row
// src_rows[sample] in place (ab200_trace_args::src_rows) and the autograd backward of that gather - the sum of the replicas'
// gradient rows - is this kernel.  Replicas of one heliostat are contiguous (repeat_interleave), so the sum runs over a
// row range in a fixed order: deterministic, no atomics.
// ---------------------------------------------------------------------------------------------
namespace ab200 {
template <typename V>
__global__ void __launch_bounds__(256) replica_sum_kernel(const V* __restrict__ in, const int* __restrict__ row_start, long long row_vecs,
                                                          V* __restrict__ out) {
    const int s = blockIdx.y;
    const int k0 = __ldg(row_start + s), k1 = __ldg(row_start + s + 1);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < row_vecs; i += (long long)gridDim.x * blockDim.x) {
        V acc = V();
        for (int k = k0; k < k1; ++k) {
            const V v = __ldcs(in + (long long)k * row_vecs + i);
            if constexpr (sizeof(V) == 16) { acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
            else acc += v;
        }
        out[(long long)s * row_vecs + i] = acc;
    }
}
}  // namespace ab200

extern "C" int32_t ab200_replica_sum(const float* in, const int32_t* row_start, int32_t n_src, int64_t row_elems, float* out,
                                     void* stream) {
    AB200_REQUIRE(in && row_start && out && n_src >= 0 && row_elems >= 0, AB200_EINVAL, "bad arguments");
    if (n_src == 0 || row_elems == 0) return AB200_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec = row_elems % 4 == 0 && reinterpret_cast<uintptr_t>(in) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0;
    const long long row_vecs = vec ? row_elems / 4 : row_elems;
    const unsigned gx = (unsigned)std::min<long long>((row_vecs + 255) / 256, 64);
    for (int32_t s0 = 0; s0 < n_src; s0 += 65535) {      // gridDim.y limit
        const dim3 grid(gx, (unsigned)std::min<int32_t>(65535, n_src - s0));
        if (vec)
            ab200::replica_sum_kernel<float4><<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(in), row_start + s0, row_vecs,
                                                                   reinterpret_cast<float4*>(out) + (long long)s0 * row_vecs);
        else
            ab200::replica_sum_kernel<float><<<grid, 256, 0, st>>>(in, row_start + s0, row_vecs, out + (long long)s0 * row_vecs);
        ab200::note_launch();
        AB200_CUDA_TRY(cudaGetLastError());
    }
    return AB200_OK;
}
