// Microbenchmark: does FFMA2 (fma.rn.f32x2) save issue slots on sm_100a?
// Per step 8 fp32 FMAs (as 8 FFMA or 4 FFMA2) plus K single-instruction integer ops (LOP3) or K shared-memory loads.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <bool PACKED, int K, bool LDS>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b, unsigned m) {
    __shared__ float sh[256 * 9];
    float x[8];
    unsigned u[8];
    float acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 0.001f + i, u[i] = threadIdx.x * 2654435761u + i;
    for (int i = threadIdx.x; i < 256 * 9; i += 256) sh[i] = i;
    __syncthreads();
    const float* sp = sh + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (!PACKED) {
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
            } else {
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                    float2 v = __ffma2_rn(make_float2(x[i], x[i + 1]), make_float2(a, a), make_float2(b, b));
                    x[i] = v.x, x[i + 1] = v.y;
                }
            }
#pragma unroll
            for (int i = 0; i < K; ++i) {
                if (LDS) acc += sp[((it + i) & 7) * 256];   // 1 LDS + 1 FADD (+ address math hoisted by unrolling)
                else u[i & 7] = (u[i & 7] ^ m) & u[(i + 1) & 7];
            }
        }
    }
    float s = acc;
    unsigned t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i], t += u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)t;
}

template <bool PACKED, int K, bool LDS>
void run(float* out, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    k<PACKED, K, LDS><<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f, 0x55555555u);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) k<PACKED, K, LDS><<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f, 0x55555555u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    const double steps = 148.0 * 8 * 8 * iters * 4.0;  // warp-steps
    // cycles per warp-step per scheduler: 148 SMs x 4 schedulers at ~1.965 GHz
    printf("{\"fma\": \"%s\", \"extra\": \"%d %s\", \"ms\": %.4f, \"sched_cycles_per_step\": %.2f}\n", PACKED ? "4 FFMA2" : "8 FFMA", K, LDS ? "LDS+FADD" : "LOP3",
           ms, ms * 1e-3 * 1.965e9 * 148 * 4 / steps);
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 4096;
    run<false, 0, false>(out, iters), run<true, 0, false>(out, iters);
    run<false, 2, false>(out, iters), run<true, 2, false>(out, iters);
    run<false, 4, false>(out, iters), run<true, 4, false>(out, iters);
    run<false, 8, false>(out, iters), run<true, 8, false>(out, iters);
    run<false, 2, true>(out, iters), run<true, 2, true>(out, iters);
    run<false, 4, true>(out, iters), run<true, 4, true>(out, iters);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
