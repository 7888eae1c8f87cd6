// micro-benchmark: does fma.rn.f32x2 double FP32 FMA throughput per issue slot on sm_100a?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_scalar(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long x, unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__global__ void k_packed(float* out, int iters, float a, float b) {
    float t = threadIdx.x;
    unsigned long long x0 = pack(t, t + 1), x1 = pack(t + 2, t + 3), x2 = pack(t + 4, t + 5), x3 = pack(t + 6, t + 7);
    unsigned long long A = pack(a, a), B = pack(b, b);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { x0 = fma2(x0, A, B); x1 = fma2(x1, A, B); x2 = fma2(x2, A, B); x3 = fma2(x3, A, B); }
    }
    unsigned long long s = x0 ^ x1 ^ x2 ^ x3;
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)(s & 0xffffffffu)) + __uint_as_float((unsigned)(s >> 32));
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000; float ms;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); k_scalar<<<148 * 8, 256>>>(out, iters, 0.999f, 0.001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 148 * 8 * 256 * (double)iters * 64;
        printf("scalar FFMA : %.3f ms  %.1f TFLOP/s\n", ms, flops / ms * 1e-9);
        cudaEventRecord(e0); k_packed<<<148 * 8, 256>>>(out, iters, 0.999f, 0.001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("packed FFMA2: %.3f ms  %.1f TFLOP/s (same flop count)\n", ms, flops / ms * 1e-9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
