// Microbenchmark: what does the memory system deliver for the ACCESS PATTERN of the streamed row kernels, with no compute?
// Persistent CTAs (CPS per SM), one producer lane issuing cp.async.bulk row copies into a ring of NST stages x HS planes,
// consumer warps only wait on the full barrier, touch one value per row and release the stage.  Row (b, y) of plane n sits at
// ((b * N + n) * H + y) * W floats: consecutive planes of one row are H * W floats apart, exactly as in [B, N, H, W] logits.
// usage: tma_stream_bench [rows_per_copy] [cps] [nst] [hs]
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream_bench tma_stream_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* q) { return (uint32_t)__cvta_generic_to_shared(q); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.b32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_row(float* dst, const float* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}

constexpr int MAXST = 8;

__global__ void __launch_bounds__(192) stream_kernel(const float* __restrict__ logits, float* __restrict__ out, int B, int N, int H, int W, int rpc, int nst,
                                                     int hs) {
    extern __shared__ __align__(128) unsigned char raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(raw);
    uint64_t* empty = full + MAXST;
    float* ring = reinterpret_cast<float*>(raw + 128);
    const int ncw = blockDim.x / 32 - 1;  // consumer warps
    if (threadIdx.x == 0) {
        for (int i = 0; i < nst; ++i) mbar_init(full + i, 1), mbar_init(empty + i, ncw);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int ngroups = B * H / rpc, nblk = (N + hs - 1) / hs;
    const int nit = (ngroups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t rowf = (size_t)rpc * W;
    int stage = 0, use = 0;
    if (warp == ncw) {  // producer
        for (int it = 0; it < nit; ++it) {
            const int g = blockIdx.x + it * gridDim.x, row0 = g * rpc, b = row0 / H, y = row0 - b * H;
            for (int j = 0; j < nblk; ++j) {
                if (use > 0) mbar_wait(empty + stage, (use - 1) & 1);
                const int n0 = j * hs, n1 = min(N, n0 + hs);
                if (lane == 0) {
                    mbar_expect_tx(full + stage, (uint32_t)((n1 - n0) * rowf * 4));
                    for (int n = n0; n < n1; ++n)
                        tma_row(ring + ((size_t)stage * hs + (n - n0)) * rowf, logits + (((size_t)b * N + n) * H + y) * W, (uint32_t)(rowf * 4), full + stage);
                }
                if (++stage == nst) stage = 0, ++use;
            }
        }
        return;
    }
    float acc = 0.0f;
    uint32_t ph = 0;
    for (int it = 0; it < nit; ++it) {
        for (int j = 0; j < nblk; ++j) {
            mbar_wait(full + stage, ph);
            acc += ring[(size_t)stage * hs * rowf + threadIdx.x];
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + stage);
            if (++stage == nst) stage = 0, ph ^= 1;
        }
    }
    if (acc == 123.456f) out[threadIdx.x] = acc;
}

int main(int argc, char** argv) {
    const int rpc = argc > 1 ? atoi(argv[1]) : 1, cps = argc > 2 ? atoi(argv[2]) : 4, nst = argc > 3 ? atoi(argv[3]) : 3, hs = argc > 4 ? atoi(argv[4]) : 4;
    const int B = 12, N = 49, H = 192, W = 640;
    const size_t n = (size_t)B * N * H * W;
    float *d, *out;
    cudaMalloc(&d, n * 4);
    cudaMalloc(&out, 4096);
    cudaMemset(d, 0, n * 4);
    const size_t smem = 128 + (size_t)nst * hs * rpc * W * 4;
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    const int grid = 148 * cps;
    for (int i = 0; i < 3; ++i) stream_kernel<<<grid, 192, smem>>>(d, out, B, N, H, W, rpc, nst, hs);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) stream_kernel<<<grid, 192, smem>>>(d, out, B, N, H, W, rpc, nst, hs);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    printf("{\"rows_per_copy\": %d, \"ctas_per_sm\": %d, \"nst\": %d, \"hs\": %d, \"smem_kb\": %.1f, \"ms\": %.4f, \"GBps\": %.0f, \"err\": \"%s\"}\n", rpc, cps, nst, hs,
           smem / 1024.0, ms, n * 4 / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
