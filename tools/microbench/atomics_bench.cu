// Micro-benchmark: which accumulate primitive should the flux splat use on B200?
// Measures lane-updates/s for the 4-tap bilinear deposit pattern (idx, idx+1, idx+row, idx+row+1)
// with pseudo-random idx inside a WxW window, for: native smem int32 ATOMS.ADD, smem fp32 CAS
// atomicAdd, global REDG.F32 (per-CTA private 256x256 bitmap, L2 resident), REDG.F32x2, REDG.F32x4.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomics_bench atomics_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

__device__ __forceinline__ uint32_t lcg(uint32_t& s){ s = s*1664525u + 1013904223u; return s>>8; }

template<int MODE>
__global__ void __launch_bounds__(256) bench(float* gbuf, int iters, int win, int row /*bitmap row pitch*/, unsigned long long* sink)
{
    extern __shared__ int smem[];
    float* sf = reinterpret_cast<float*>(smem);
    const int tid = threadIdx.x;
    // per-CTA private global bitmap of row*row floats
    float* g = gbuf + (size_t)blockIdx.x * row * row;
    if (MODE <= 1 || MODE == 5) { for (int i = tid; i < win*win; i += blockDim.x) smem[i] = 0; __syncthreads(); }
    uint32_t s = (blockIdx.x * 256 + tid) * 2654435761u + 12345u;
    for (int it = 0; it < iters; ++it) {
        uint32_t rnd = lcg(s);
        int x = rnd % (win - 1);
        int y = (rnd >> 12) % (win - 1);
        float v = (float)(rnd & 255) * 0.001f;
        if (MODE == 0) {            // smem int32 native
            int q = __float2int_rn(v * 1024.f);
            int b = y*win + x;
            atomicAdd(&smem[b], q); atomicAdd(&smem[b+1], q); atomicAdd(&smem[b+win], q); atomicAdd(&smem[b+win+1], q);
        } else if (MODE == 1) {     // smem fp32 (CAS loop)
            int b = y*win + x;
            atomicAdd(&sf[b], v); atomicAdd(&sf[b+1], v); atomicAdd(&sf[b+win], v); atomicAdd(&sf[b+win+1], v);
        } else if (MODE == 2) {     // global REDG.F32 x4 scalar
            int b = y*row + x;
            atomicAdd(&g[b], v); atomicAdd(&g[b+1], v); atomicAdd(&g[b+row], v); atomicAdd(&g[b+row+1], v);
        } else if (MODE == 3) {     // global REDG.F32x2 (aligned pair)
            int b = y*row + (x & ~1);
            atomicAdd(reinterpret_cast<float2*>(&g[b]), make_float2(v, v));
            atomicAdd(reinterpret_cast<float2*>(&g[b+row]), make_float2(v, v));
        } else if (MODE == 4) {     // global REDG.F32x4 (aligned quad, two lanes zero)
            int b = y*row + (x & ~3);
            atomicAdd(reinterpret_cast<float4*>(&g[b]), make_float4(v, v, 0.f, 0.f));
            atomicAdd(reinterpret_cast<float4*>(&g[b+row]), make_float4(v, v, 0.f, 0.f));
        } else if (MODE == 5) {     // hybrid: upper row smem int, lower row global x2
            int q = __float2int_rn(v * 1024.f);
            int b = y*win + x;
            atomicAdd(&smem[b], q); atomicAdd(&smem[b+1], q);
            int gb = y*row + (x & ~1);
            atomicAdd(reinterpret_cast<float2*>(&g[gb+row]), make_float2(v, v));
        } else if (MODE == 6) {     // no accumulate at all: loop overhead baseline
            if (v < -1.f) sink[0] = rnd;
        }
    }
    if (MODE <= 1 || MODE == 5) { __syncthreads(); unsigned long long acc = 0; for (int i = tid; i < win*win; i += blockDim.x) acc += (unsigned)smem[i]; if (acc == 0xdeadbeefULL) sink[1] = acc; }
}

template<int MODE>
void run(const char* name, int win, int taps_per_iter, float* gbuf, unsigned long long* sink, int ctas, int iters)
{
    size_t smem_bytes = (MODE <= 1 || MODE == 5) ? (size_t)win*win*4 : 0;
    CK(cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    bench<MODE><<<ctas, 256, smem_bytes>>>(gbuf, iters/8, win, 256, sink);   // warm-up
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        bench<MODE><<<ctas, 256, smem_bytes>>>(gbuf, iters, win, 256, sink);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    double rays = (double)ctas * 256 * iters;
    printf("%-34s win=%3d ctas=%5d  %8.3f ms  %8.2f Grays/s  %8.2f Gtaps/s (%d instr-taps/ray)\n", name, win, ctas, best, rays/best*1e-6, rays*4/best*1e-6, taps_per_iter);
}

int main()
{
    int dev = 0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    printf("device %s SMs=%d smem/block optin=%zu\n", p.name, p.multiProcessorCount, (size_t)p.sharedMemPerBlockOptin);
    const int sms = p.multiProcessorCount;
    const int max_ctas = sms * 8;
    float* gbuf; CK(cudaMalloc(&gbuf, (size_t)max_ctas * 256 * 256 * 4)); CK(cudaMemset(gbuf, 0, (size_t)max_ctas * 256 * 256 * 4));
    unsigned long long* sink; CK(cudaMalloc(&sink, 64));
    const int iters = 2048;
    for (int win : {32, 64, 104}) {
        int ctas = (win == 104) ? sms * 4 : sms * 8;   // 104*104*4 = 43 KB -> 4-5 CTAs/SM
        run<0>("smem int32 ATOMS.ADD x4", win, 4, gbuf, sink, ctas, iters);
        run<1>("smem fp32 CAS atomicAdd x4", win, 4, gbuf, sink, ctas, iters);
    }
    for (int win : {32, 64, 128, 255}) {
        run<2>("global REDG.F32 x4", win, 4, gbuf, sink, sms * 8, iters);
        run<3>("global REDG.F32x2 x2", win, 2, gbuf, sink, sms * 8, iters);
        run<4>("global REDG.F32x4 x2", win, 2, gbuf, sink, sms * 8, iters);
    }
    run<5>("hybrid smem-int x2 + REDG.F32x2 x1", 64, 3, gbuf, sink, sms * 8, iters);
    run<6>("loop overhead only", 64, 0, gbuf, sink, sms * 8, iters);
    // fewer CTAs (1 per SM) to see per-SM limits vs chip limits
    run<2>("global REDG.F32 x4 (1 CTA/SM)", 128, 4, gbuf, sink, sms, iters);
    run<3>("global REDG.F32x2 x2 (1 CTA/SM)", 128, 2, gbuf, sink, sms, iters);
    run<0>("smem int32 ATOMS.ADD x4 (1 CTA/SM)", 104, 4, gbuf, sink, sms, iters);
    return 0;
}
