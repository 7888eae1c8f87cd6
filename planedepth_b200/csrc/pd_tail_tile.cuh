// Decoder tail (networks/depth_decoder.py:258-291), TMA-tile kernels.
//
// Same arithmetic as pd_tail.cuh (one thread per pixel, LDG plane loops: latency-bound, 0.23 - 0.34 of the HBM roofline), laid
// out for memory-level parallelism: a CTA owns one (image row, column tile); one elected thread issues a bulk copy
// (cp.async.bulk, SASS UBLKCP) for EVERY plane row of the tile up front — N x tile bytes in flight per CTA instead of a few
// 4-byte loads per thread — with one mbarrier per group of planes so that the first pass starts when the first group has
// landed; the softmax over planes then runs in passes over shared memory (max, sum, probabilities) and every output row leaves
// as 64-bit stores, 256 contiguous bytes per warp.  Two pixels per thread.
//
//   forward : tile = raw logits [N][tw] (+ raw sigma [N][tw] with the mixture); in place: l2 = logit * log2(e), then weights
//   backward: tile = saved logits [N][tw] -> softmax probabilities in place; g_logits / sigma / masks stream through LDG.64
#pragma once
#include "pd_tail.cuh"

namespace pd {
namespace tl {

constexpr int TT_PX = 2;        // pixels per thread
constexpr int TT_GROUP = 8;     // plane rows per mbarrier
constexpr int TT_MAXG = 32;     // groups (N <= 256)

struct TileCfg {
    int tw;        // tile width in pixels (divides W, multiple of 4)
    int tiles;     // tiles per image row
    int threads;   // CTA size (multiple of 32, >= tw / 2)
    size_t smem;   // dynamic shared memory
};

__device__ __forceinline__ uint32_t tt_smem_u32(const void* q) { return (uint32_t)__cvta_generic_to_shared(q); }
__device__ __forceinline__ void tt_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tt_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tt_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tt_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tt_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    const uint32_t a = tt_smem_u32(bar);
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(a), "r"(parity), "r"(20000u)
            : "memory");
    }
}
__device__ __forceinline__ void tt_tma_row(float* dst, const float* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tt_smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(tt_smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ float2 ldg2(const float* q) { return __ldg(reinterpret_cast<const float2*>(q)); }
__device__ __forceinline__ void stg2(float* q, float a, float b) { *reinterpret_cast<float2*>(q) = make_float2(a, b); }

// mask values of this thread's two pixels: staged row value, one 64-bit load (dense fp32 mask, unit x stride), or two scalar loads
__device__ __forceinline__ float2 tt_mask2(const TailParams& p, const float* mrow, int mode, int b, int n, int y, int x) {
    if (mode == 0) return make_float2(mrow[n], mrow[n]);
    if (mode == 1) return ldg2(reinterpret_cast<const float*>(p.mask) + soff(p.ms, b, n, y, x));
    return make_float2(load_mask(p.mask, p.mask_dtype, soff(p.ms, b, n, y, x)), load_mask(p.mask, p.mask_dtype, soff(p.ms, b, n, y, x + 1)));
}
__device__ __forceinline__ int tt_mask_mode(const TailParams& p) {
    if (p.mask_dtype == PD_MASK_NONE || p.ms.x == 0) return 0;
    const bool v2 = p.mask_dtype == PD_MASK_F32 && p.ms.x == 1 && ((p.ms.y | p.ms.n | p.ms.b) & 1) == 0 && (reinterpret_cast<uintptr_t>(p.mask) & 7) == 0;
    return v2 ? 1 : 2;
}

// per-plane scalars of the tile's row when they do not vary along x (row mask, decoder disparities): staged once per CTA
__device__ __forceinline__ void tt_stage_row_scalars(const TailParams& p, int b, int y, float* mrow, float* drow) {
    for (int n = threadIdx.x; n < p.N; n += blockDim.x) {
        mrow[n] = (p.ms.x == 0 || p.mask_dtype == PD_MASK_NONE) ? load_mask(p.mask, p.mask_dtype, soff(p.ms, b, n, y, 0)) : 1.0f;
        drow[n] = (p.ds.x == 0) ? __ldg(p.disp_layered + soff(p.ds, b, n, y, 0)) : 0.0f;
    }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <bool MIX>
__global__ void __launch_bounds__(160) tail_fwd_tile_kernel(const TailParams p, const int tw, const int tiles) {
    extern __shared__ __align__(128) unsigned char tt_raw[];
    const int N = p.N, W = p.W, H = p.H;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tt_raw);                       // [TT_MAXG]
    float* mrow = reinterpret_cast<float*>(tt_raw + TT_MAXG * 8);                // [N] row mask values
    float* drow = mrow + N;                                                     // [N] row disparities
    float* tile = reinterpret_cast<float*>(tt_raw + TT_MAXG * 8 + (((size_t)2 * N * 4 + 15) & ~(size_t)15));  // [N][tw] (+ [N][tw] sigma)
    float* stile = tile + (size_t)N * tw;
    const int t = blockIdx.x % tiles;
    const int row = blockIdx.x / tiles;
    const int b = row / H, y = row - b * H;
    const int x0 = t * tw;
    const int ng = (N + TT_GROUP - 1) / TT_GROUP;
    if (threadIdx.x == 0) {
        for (int g = 0; g < ng; ++g) tt_mbar_init(bars + g, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tt_stage_row_scalars(p, b, y, mrow, drow);
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t rb = (uint32_t)tw * 4u;
        const float* src = p.raw + (((int64_t)b * N) * H + y) * W + x0;
        const float* ssrc = MIX ? p.sraw + (((int64_t)b * N) * H + y) * W + x0 : nullptr;
        for (int g = 0; g < ng; ++g) {
            const int n0 = g * TT_GROUP, n1 = min(N, n0 + TT_GROUP);
            tt_mbar_expect_tx(bars + g, (uint32_t)(n1 - n0) * rb * (MIX ? 2u : 1u));
            for (int n = n0; n < n1; ++n) {
                tt_tma_row(tile + (size_t)n * tw, src + (int64_t)n * p.hw, rb, bars + g);
                if (MIX) tt_tma_row(stile + (size_t)n * tw, ssrc + (int64_t)n * p.hw, rb, bars + g);
            }
        }
    }
    const int px = threadIdx.x * TT_PX;
    const bool active = px < tw;
    const int x = x0 + px;
    const int64_t rem = (int64_t)y * W + x;
    const int64_t base = (int64_t)b * N * p.hw + rem;
    const int mmode = tt_mask_mode(p);  // 0: row value staged in shared memory, 1: 64-bit loads, 2: scalar loads
    const bool disp_px = p.ds.x != 0;
    // pass 1: masked logits out, l2 = logit * log2(e) in place, running maximum
    float M0 = -INFINITY, M1 = -INFINITY;
    for (int g = 0; g < ng; ++g) {
        tt_mbar_wait(bars + g, 0);
        if (!active) continue;
        const int n0 = g * TT_GROUP, n1 = min(N, n0 + TT_GROUP);
        // the group's mask values first (independent loads, all in flight together), then the arithmetic
        float2 mk[TT_GROUP];
#pragma unroll
        for (int i = 0; i < TT_GROUP; ++i) mk[i] = (n0 + i < n1) ? tt_mask2(p, mrow, mmode, b, n0 + i, y, x) : make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < TT_GROUP; ++i) {
            const int n = n0 + i;
            if (n < n1) {
                float2 v = *reinterpret_cast<const float2*>(tile + (size_t)n * tw + px);
                v.x *= mk[i].x, v.y *= mk[i].y;
                stg2(p.logits + base + (int64_t)n * p.hw, v.x, v.y);
                v.x *= kLog2e, v.y *= kLog2e;
                *reinterpret_cast<float2*>(tile + (size_t)n * tw + px) = v;
                M0 = fmaxf(M0, v.x), M1 = fmaxf(M1, v.y);
            }
        }
    }
    if (!active) return;
    // pass 2: e = exp2(l2 - M) in place, S = sum e
    float S0 = 0.0f, S1 = 0.0f;
#pragma unroll 4
    for (int n = 0; n < N; ++n) {
        float2 v = *reinterpret_cast<const float2*>(tile + (size_t)n * tw + px);
        v.x = fast_exp2(v.x - M0), v.y = fast_exp2(v.y - M1);
        S0 += v.x, S1 += v.y;
        *reinterpret_cast<float2*>(tile + (size_t)n * tw + px) = v;
    }
    const float iS0 = 1.0f / S0, iS1 = 1.0f / S1;
    float Z0 = 1.0f, Z1 = 1.0f;
    if (MIX) {
        // mixture: sigma = clamp(sigmoid(raw)), w = pi / sigma * mask in place, Z = sum w
        Z0 = Z1 = 0.0f;
#pragma unroll 4
        for (int n = 0; n < N; ++n) {
            float2 e = *reinterpret_cast<const float2*>(tile + (size_t)n * tw + px);
            const float2 sr = *reinterpret_cast<const float2*>(stile + (size_t)n * tw + px);
            const float2 mk = tt_mask2(p, mrow, mmode, b, n, y, x);
            const float m0 = mk.x, m1 = mk.y;
            const float pi0 = e.x * iS0, pi1 = e.y * iS1;
            const float sg0 = sigmoid_clamped(sr.x), sg1 = sigmoid_clamped(sr.y);
            stg2(p.sigma + base + (int64_t)n * p.hw, sg0, sg1);
            if (p.pi) stg2(p.pi + base + (int64_t)n * p.hw, pi0, pi1);
            e.x = pi0 / sg0 * m0, e.y = pi1 / sg1 * m1;
            Z0 += e.x, Z1 += e.y;
            *reinterpret_cast<float2*>(tile + (size_t)n * tw + px) = e;
        }
    }
    // pass 3: probabilities out, composited disparity
    const float k0 = MIX ? 1.0f / Z0 : iS0, k1 = MIX ? 1.0f / Z1 : iS1;
    float d0 = 0.0f, d1 = 0.0f;
#pragma unroll 4
    for (int n = 0; n < N; ++n) {
        const float2 e = *reinterpret_cast<const float2*>(tile + (size_t)n * tw + px);
        const float pr0 = e.x * k0, pr1 = e.y * k1;
        stg2(p.prob + base + (int64_t)n * p.hw, pr0, pr1);
        float dl0 = drow[n], dl1 = dl0;
        if (disp_px) {
            dl0 = __ldg(p.disp_layered + soff(p.ds, b, n, y, x));
            dl1 = __ldg(p.disp_layered + soff(p.ds, b, n, y, x + 1));
        }
        d0 = fmaf(pr0, dl0, d0), d1 = fmaf(pr1, dl1, d1);
    }
    const int64_t pix = (int64_t)b * p.hw + rem;
    stg2(p.disp + pix, d0, d1);
    if (p.depth) stg2(p.depth + pix, p.depth_c / d0, p.depth_c / d1);
    float* st = p.stats + (int64_t)b * 3 * p.hw + rem;
    stg2(st, M0, M1);
    stg2(st + p.hw, S0, S1);
    stg2(st + 2 * p.hw, Z0, Z1);
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
constexpr int TT_BWD_WARPS = 5;               // launch bound of the backward tile kernel: 160 threads
constexpr int TT_BWD_HDR = 2 + TT_BWD_WARPS;  // per-plane header floats: row mask, row disparity, one gradient sum per warp

template <bool MIX>
__global__ void __launch_bounds__(160) tail_bwd_tile_kernel(const TailParams p, const int tw, const int tiles) {
    extern __shared__ __align__(128) unsigned char tt_raw[];
    const int N = p.N, W = p.W, H = p.H;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tt_raw);
    float* mrow = reinterpret_cast<float*>(tt_raw + TT_MAXG * 8);
    float* drow = mrow + N;
    float* gacc0 = drow + N;  // [TT_BWD_WARPS][N] sums of the compact disparity gradient, one private row per warp (a shared
    float* gacc = gacc0 + (threadIdx.x >> 5) * N;  // float atomicAdd compiles to a compare-and-swap loop)
    float* tile = reinterpret_cast<float*>(tt_raw + TT_MAXG * 8 + (((size_t)TT_BWD_HDR * N * 4 + 15) & ~(size_t)15));  // [N][tw] logits -> pi
    const int t = blockIdx.x % tiles;
    const int row = blockIdx.x / tiles;
    const int b = row / H, y = row - b * H;
    const int x0 = t * tw;
    const int ng = (N + TT_GROUP - 1) / TT_GROUP;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int g = 0; g < ng; ++g) tt_mbar_init(bars + g, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tt_stage_row_scalars(p, b, y, mrow, drow);
    for (int n = threadIdx.x; n < TT_BWD_WARPS * N; n += blockDim.x) gacc0[n] = 0.0f;
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t rb = (uint32_t)tw * 4u;
        const float* src = p.logits + (((int64_t)b * N) * H + y) * W + x0;
        for (int g = 0; g < ng; ++g) {
            const int n0 = g * TT_GROUP, n1 = min(N, n0 + TT_GROUP);
            tt_mbar_expect_tx(bars + g, (uint32_t)(n1 - n0) * rb);
            for (int n = n0; n < n1; ++n) tt_tma_row(tile + (size_t)n * tw, src + (int64_t)n * p.hw, rb, bars + g);
        }
    }
    const int px = threadIdx.x * TT_PX;
    const bool active = px < tw;
    const int x = x0 + (active ? px : 0);
    const int64_t rem = (int64_t)y * W + x;
    const int64_t base = (int64_t)b * N * p.hw + rem;
    const int64_t pix = (int64_t)b * p.hw + rem;
    const int mmode = tt_mask_mode(p);
    const bool disp_px = p.ds.x != 0;
    const float* st = p.stats + (int64_t)b * 3 * p.hw + rem;
    float2 Mv = make_float2(0.f, 0.f), iS = make_float2(1.f, 1.f), iZ = make_float2(1.f, 1.f), gd = make_float2(0.f, 0.f);
    if (active) {
        Mv = ldg2(st);
        const float2 Sv = ldg2(st + p.hw), Zv = ldg2(st + 2 * p.hw);
        iS = make_float2(1.0f / Sv.x, 1.0f / Sv.y), iZ = make_float2(1.0f / Zv.x, 1.0f / Zv.y);
        if (p.g_disp) gd = ldg2(p.g_disp + pix);
        if (p.g_depth) {
            const float2 dv = ldg2(p.disp + pix), gz = ldg2(p.g_depth + pix);
            gd.x -= gz.x * p.depth_c / (dv.x * dv.x), gd.y -= gz.y * p.depth_c / (dv.y * dv.y);  // depth = c / disp
        }
    }
    const bool has_gp = p.g_prob != nullptr;
    // pass 1: softmax probabilities in place; dotp = sum_k probability_k * gp_k,  gp_k = g_prob_k + gd * disp_layered_k
    float dot0 = 0.0f, dot1 = 0.0f;
    for (int g = 0; g < ng; ++g) {
        tt_mbar_wait(bars + g, 0);
        if (!active) continue;
        const int n0 = g * TT_GROUP, n1 = min(N, n0 + TT_GROUP);
#pragma unroll 4
        for (int n = n0; n < n1; ++n) {
            const int64_t o = base + (int64_t)n * p.hw;
            float2 v = *reinterpret_cast<const float2*>(tile + (size_t)n * tw + px);
            v.x = fast_exp2(fmaf(v.x, kLog2e, -Mv.x)) * iS.x, v.y = fast_exp2(fmaf(v.y, kLog2e, -Mv.y)) * iS.y;
            *reinterpret_cast<float2*>(tile + (size_t)n * tw + px) = v;
            float pr0 = v.x, pr1 = v.y;
            if (MIX) {
                const float2 sg = ldg2(p.sigma + o);
                const float2 mk = tt_mask2(p, mrow, mmode, b, n, y, x);
                pr0 = v.x / sg.x * mk.x * iZ.x, pr1 = v.y / sg.y * mk.y * iZ.y;
            }
            float dl0 = drow[n], dl1 = dl0;
            if (disp_px) {
                dl0 = __ldg(p.disp_layered + soff(p.ds, b, n, y, x));
                dl1 = __ldg(p.disp_layered + soff(p.ds, b, n, y, x + 1));
            }
            float gp0 = gd.x * dl0, gp1 = gd.y * dl1;
            if (has_gp) {
                const float2 q = ldg2(p.g_prob + o);
                gp0 += q.x, gp1 += q.y;
            }
            dot0 = fmaf(pr0, gp0, dot0), dot1 = fmaf(pr1, gp1, dot1);
        }
    }
    // pass 2: gradients (the softmax-normaliser term of the mixture vanishes identically, see pd_tail.cuh)
    constexpr int CH = MIX ? 4 : 8;  // planes per chunk: every global operand of the chunk is requested before the first one is used
    const bool xred = p.g_dl && !p.g_dl_dense && p.gds.x == 0;
    for (int nc = 0; nc < N; nc += CH) {
        float2 glo[CH], mk[CH], qp[CH], sgv[CH], gsu[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int n = min(nc + i, N - 1);
            const int64_t o = base + (int64_t)n * p.hw;
            glo[i] = (active && p.g_logits) ? ldg2(p.g_logits + o) : make_float2(0.f, 0.f);
            mk[i] = active ? tt_mask2(p, mrow, mmode, b, n, y, x) : make_float2(0.f, 0.f);
            qp[i] = (active && has_gp) ? ldg2(p.g_prob + o) : make_float2(0.f, 0.f);
            if (MIX) {
                sgv[i] = active ? ldg2(p.sigma + o) : make_float2(1.f, 1.f);
                gsu[i] = (active && p.g_sigma) ? ldg2(p.g_sigma + o) : make_float2(0.f, 0.f);
            }
        }
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int n = nc + i;
            if (n >= N) break;
            const int64_t o = base + (int64_t)n * p.hw;
            float gdl = 0.0f;
            if (active) {
                const float2 pi = *reinterpret_cast<const float2*>(tile + (size_t)n * tw + px);
                const float m0 = mk[i].x, m1 = mk[i].y;
                float dl0 = drow[n], dl1 = dl0;
                if (disp_px) {
                    dl0 = __ldg(p.disp_layered + soff(p.ds, b, n, y, x));
                    dl1 = __ldg(p.disp_layered + soff(p.ds, b, n, y, x + 1));
                }
                const float gp0 = gd.x * dl0 + qp[i].x, gp1 = gd.y * dl1 + qp[i].y;
                float pr0 = pi.x, pr1 = pi.y, gl0, gl1;
                if (MIX) {
                    const float2 sg = sgv[i];
                    const float a0 = 1.0f / sg.x, a1 = 1.0f / sg.y;
                    pr0 = pi.x * a0 * m0 * iZ.x, pr1 = pi.y * a1 * m1 * iZ.y;
                    const float gw0 = (gp0 - dot0) * iZ.x, gw1 = (gp1 - dot1) * iZ.y;  // d / d w_n with probability = w / sum w
                    gl0 = pi.x * (gw0 * m0 * a0) + glo[i].x, gl1 = pi.y * (gw1 * m1 * a1) + glo[i].y;
                    const float gs0 = -gw0 * pi.x * m0 * a0 * a0 + gsu[i].x, gs1 = -gw1 * pi.y * m1 * a1 * a1 + gsu[i].y;
                    // clamp passes the gradient inside [0.01, 1]; sigmoid' = s (1 - s) with s = sigma there
                    if (p.g_sraw) stg2(p.g_sraw + o, (sg.x > 0.01f) ? gs0 * sg.x * (1.0f - sg.x) : 0.0f, (sg.y > 0.01f) ? gs1 * sg.y * (1.0f - sg.y) : 0.0f);
                } else {
                    gl0 = pi.x * (gp0 - dot0) + glo[i].x, gl1 = pi.y * (gp1 - dot1) + glo[i].y;
                }
                if (p.g_raw) stg2(p.g_raw + o, gl0 * m0, gl1 * m1);  // logits = raw * mask
                if (p.g_dl) {
                    const float v0 = gd.x * pr0, v1 = gd.y * pr1;
                    if (p.g_dl_dense) {
                        float* dst = p.g_dl + soff(p.gds, b, n, y, x);
                        dst[0] = v0, dst[p.gds.x] = v1;
                    } else if (p.gds.x == 0) {
                        gdl = v0 + v1;
                    } else {  // reduced over another dimension only: per-pixel atomics
                        if (v0 != 0.0f) atomicAdd(p.g_dl + soff(p.gds, b, n, y, x), v0);
                        if (v1 != 0.0f) atomicAdd(p.g_dl + soff(p.gds, b, n, y, x + 1), v1);
                    }
                }
            }
            if (xred) {  // reduced over x (and possibly y): warp sum into the warp's own row
                const float s_ = warp_sum(gdl);
                if (lane == 0) gacc[n] += s_;
            }
        }
    }
    if (p.g_dl && !p.g_dl_dense && p.gds.x == 0) {
        __syncthreads();
        for (int n = threadIdx.x; n < N; n += blockDim.x) {
            float v = 0.0f;
#pragma unroll
            for (int w = 0; w < TT_BWD_WARPS; ++w) v += gacc0[w * N + n];
            if (v != 0.0f) atomicAdd(p.g_dl + soff(p.gds, b, n, y, 0), v);  // a zero y stride folds the rows as well
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
inline bool tile_cfg(const TailParams& p, bool mix_fwd, int hdr_floats_per_plane, TileCfg& c) {
    const int W = p.W, N = p.N;
    if (W % 4 != 0 || N > TT_GROUP * TT_MAXG) return false;
    // widest tile that divides W, keeps 64-bit accesses aligned (even width) and lets at least two CTAs share an SM
    const size_t budget = 100 * 1024;
    int best = 0;
    for (int tw = 320; tw >= 32; tw -= 4) {
        if (W % tw) continue;
        const size_t smem = TT_MAXG * 8 + (((size_t)hdr_floats_per_plane * N * 4 + 15) & ~(size_t)15) + (size_t)N * tw * 4 * (mix_fwd ? 2 : 1);
        if (smem <= budget) {
            best = tw;
            c.smem = smem;
            break;
        }
    }
    if (!best) return false;
    // the mixture forward stages two rows per plane: at N = 63 the tile shrinks to 160 columns and the direct kernel
    // measured 8-10% faster (profiles/r2_tail_timings.md), so narrow partial-width tiles are left to it
    if (mix_fwd && best < 256 && best < W) return false;
    c.tw = best;
    c.tiles = W / best;
    c.threads = ((best / TT_PX + 31) / 32) * 32;
    return c.threads <= 160;
}

inline bool tile_ptrs_ok(const TailParams& p) {
    const void* ptrs[] = {p.raw, p.sraw, p.logits, p.sigma, p.prob, p.pi, p.disp, p.depth, p.stats, p.g_logits, p.g_sigma, p.g_prob,
                          p.g_disp, p.g_depth, p.g_raw, p.g_sraw};
    for (const void* q : ptrs)
        if (q && (reinterpret_cast<uintptr_t>(q) & 15)) return false;
    // x-varying strided operands are read pixel by pixel; a dense gradient of disp_layered is written pairwise
    return true;
}

}  // namespace tl
}  // namespace pd
