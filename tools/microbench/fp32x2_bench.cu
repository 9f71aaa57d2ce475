// Micro-benchmark: does Blackwell's packed fp32x2 arithmetic (FMUL2/FADD2/FFMA2, IEEE rn per element) free issue
// slots?  Compares scalar and packed mul/add chains, alone and mixed with integer work.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2_bench fp32x2_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;} }while(0)

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, int iters, float m, float a, int salt) {
    float x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 0.001f + i;
    int z[4] = {threadIdx.x, salt, 3, 5};
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {       // scalar: 8 x (mul, add)
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = __fadd_rn(__fmul_rn(x[i], m), a);
        } else {                            // packed: 4 x (mul2, add2)
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                float2 v = make_float2(x[i], x[i + 1]);
                v = __fadd2_rn(__fmul2_rn(v, make_float2(m, m)), make_float2(a, a));
                x[i] = v.x; x[i + 1] = v.y;
            }
        }
        if (MODE >= 2) {                    // + 8 integer ops
#pragma unroll
            for (int i = 0; i < 4; ++i) { z[i] = z[i] * 3 + salt; z[i] ^= (z[i] >> 3); }
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(z[0] ^ z[1] ^ z[2] ^ z[3]);
}

template <int MODE>
float run(float* out, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, 1024>>>(out, 64, 0.999f, 0.001f, 7);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<MODE><<<148, 1024>>>(out, iters, 0.999f, 0.001f, 7);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    float* out; CK(cudaMalloc(&out, 148 * 1024 * 4));
    const int iters = 1 << 16;
    const double flops = 148.0 * 1024 * iters * 16;
    float t0 = run<0>(out, iters), t1 = run<1>(out, iters), t2 = run<2>(out, iters), t3 = run<3>(out, iters);
    printf("scalar mul+add          : %8.3f ms  %7.2f Tflop/s\n", t0, flops / t0 * 1e-9);
    printf("packed mul2+add2        : %8.3f ms  %7.2f Tflop/s\n", t1, flops / t1 * 1e-9);
    printf("scalar + 8 int ops/iter : %8.3f ms\n", t2);
    printf("packed + 8 int ops/iter : %8.3f ms\n", t3);
    return 0;
}
