// Occlusion masks and post-processed disparity of Trainer.generate_post_process_disp (trainer.py:421-466), the part
// after the flipped forward pass.  Per image of the left half (b < B; the decoder ran on the
// 2B-image batch cat([img, img.flip(-1)])):
//     plr  = softmax_n( warp(logits[b],          x + D[b]) )              :441-443
//     o_l  = min(1, sum_n warp(plr_n,            x - D[B+b]) )            :444-447
//     pfrl = softmax_n( warp(logits[B+b].flip,   x - D[B+b]) )            :449-451
//     o_fr = min(1, sum_n warp(pfrl_n,           x + D[b]) )              :452-454
//     mask_novel = min(1, sum_n warp(probability[b]_n, x + D[b]) )        :461-463
//     disp_pp = blend of disp[b], disp[B+b].flip by o_fr, o_l             :456-459
// Two kernels, one thread per target pixel, planes in a loop: "warp + softmax over planes" (samples are recomputed
// in the second pass instead of being parked in HBM) and "warp + sum over planes + clip".  The warps are horizontal
// (v = y): the default path samples at the exact positions u = x +- D, two taps in row y (the same, documented,
// deviation as the streamed stereo kernels); PD_FLAG_EXACT_COORDS reproduces the reference's fp32 normalise /
// un-normalise round trip in x and y with the four ATen taps.
#pragma once
#include "pd_device.cuh"

namespace pd {
namespace oc {

struct OcclParams {
    int B, N, H, W;
    int64_t hw;
    pd_strides4 ds;        // strides of disp_layered
    float wm1, hm1;
};

// value of plane row-major `plane` at integer column xi of row yi, zero outside, optionally mirrored in x
template <bool FLIP>
__device__ __forceinline__ float px(const float* __restrict__ plane, int yi, int xi, int W, int H) {
    if ((unsigned)xi >= (unsigned)W || (unsigned)yi >= (unsigned)H) return 0.0f;
    return __ldg(plane + (int64_t)yi * W + (FLIP ? W - 1 - xi : xi));
}

template <bool FAITHFUL, bool FLIP>
__device__ __forceinline__ float sample_shift(const float* __restrict__ plane, float u, int y, const OcclParams& p) {
    if (FAITHFUL) {
        const Taps t = make_taps(roundtrip(u, p.wm1), roundtrip((float)y, p.hm1), p.W, p.H);
        TapVals v;
        v.nw = px<FLIP>(plane, t.y0, t.x0, p.W, p.H), v.ne = px<FLIP>(plane, t.y0, t.x0 + 1, p.W, p.H);
        v.sw = px<FLIP>(plane, t.y0 + 1, t.x0, p.W, p.H), v.se = px<FLIP>(plane, t.y0 + 1, t.x0 + 1, p.W, p.H);
        return blend(v, t);
    }
    u = fminf(fmaxf(u, -2.0f), (float)(p.W + 1));
    const int x0 = __float2int_rd(u);  // one conversion; back to float on the ALU pipe (exact: |x0| <= W + 1)
    const float f0 = (float)x0;
    const float w1 = u - f0;
    return fmaf(px<FLIP>(plane, y, x0 + 1, p.W, p.H), w1, px<FLIP>(plane, y, x0, p.W, p.H) * (1.0f - w1));
}

// Q[b,n] = softmax_n( warp(P[pb + b, n] (mirrored if FLIP), x + sign * D[db + b, n]) ), b < B
template <bool FAITHFUL, bool FLIP>
__global__ void __launch_bounds__(256) warp_softmax_kernel(const OcclParams p, const float* __restrict__ P, int pb, const float* __restrict__ D, int db,
                                                           float sign, float* __restrict__ Q) {
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= (int64_t)p.B * p.hw) return;
    const int b = (int)(pix / p.hw);
    const int rem = (int)(pix - (int64_t)b * p.hw);
    const int y = rem / p.W, x = rem - y * p.W;
    const float* src = P + (int64_t)(pb + b) * p.N * p.hw;
    float M = -INFINITY, S = 0.0f;
    for (int n = 0; n < p.N; ++n) {
        const float d = __ldg(D + soff(p.ds, db + b, n, y, x));
        const float l2 = sample_shift<FAITHFUL, FLIP>(src + (int64_t)n * p.hw, (float)x + sign * d, y, p) * kLog2e;
        const float mn = fmaxf(M, l2);
        S = fmaf(S, fast_exp2(M - mn), fast_exp2(l2 - mn));
        M = mn;
    }
    const float invS = 1.0f / S;
    float* q = Q + (int64_t)b * p.N * p.hw + rem;
    for (int n = 0; n < p.N; ++n) {
        const float d = __ldg(D + soff(p.ds, db + b, n, y, x));
        const float l2 = sample_shift<FAITHFUL, FLIP>(src + (int64_t)n * p.hw, (float)x + sign * d, y, p) * kLog2e;
        q[(int64_t)n * p.hw] = fast_exp2(l2 - M) * invS;
    }
}

// out[b] = min(1, sum_n warp(P[pb + b, n], x + sign * D[db + b, n]))
template <bool FAITHFUL>
__global__ void __launch_bounds__(256) warp_sum_kernel(const OcclParams p, const float* __restrict__ P, int pb, const float* __restrict__ D, int db,
                                                       float sign, float* __restrict__ out) {
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= (int64_t)p.B * p.hw) return;
    const int b = (int)(pix / p.hw);
    const int rem = (int)(pix - (int64_t)b * p.hw);
    const int y = rem / p.W, x = rem - y * p.W;
    const float* src = P + (int64_t)(pb + b) * p.N * p.hw;
    float acc = 0.0f;
    for (int n = 0; n < p.N; ++n) {
        const float d = __ldg(D + soff(p.ds, db + b, n, y, x));
        acc += sample_shift<FAITHFUL, false>(src + (int64_t)n * p.hw, (float)x + sign * d, y, p);
    }
    out[pix] = fminf(acc, 1.0f);  // o[o > 1] = 1
}

// Both stages of one occlusion map fused per image row (the warps are horizontal, so a row only needs itself):
//     out[b, y, x] = min(1, sum_n warp2( softmax_n( warp1(P[pb + b, n]) ) ))      warp_i: u = x + sign_i * D[db_i + b, n]
// One CTA per row, PPT pixels per thread.  Pass 1 runs the softmax statistics (max, sum) of the once-warped logits over the
// planes; pass 2 recomputes each plane's sample (second read of the row: L1 / L2), turns it into its probability, parks
// the probability row in shared memory (double-buffered, zero pads = padding_mode "zeros") and gathers the second warp
// from there.  Nothing but the [B,1,H,W] result goes back to HBM.  Exact sample positions (see the file header).
constexpr int OC_PAD = 4;

template <bool FLIP, int PPT>
__global__ void __launch_bounds__(1024) occlusion_row_kernel(const OcclParams p, const float* __restrict__ P, int pb, const float* __restrict__ D,
                                                             int db1, float sign1, int db2, float sign2, float* __restrict__ out) {
    extern __shared__ float erow[];  // [2][W + 2 * OC_PAD]
    const int W = p.W, N = p.N;
    const int pitch = W + 2 * OC_PAD;
    const float Wp1 = (float)(W + 1);
    const int row = blockIdx.x, b = row / p.H, y = row - b * p.H;
    for (int i = threadIdx.x; i < 2 * pitch; i += blockDim.x) erow[i] = 0.0f;  // pads stay zero; interiors are rewritten per plane
    const float* src = P + ((int64_t)(pb + b) * N * p.H + y) * W;  // row y of plane 0; planes are hw apart
    int x[PPT];
    bool live[PPT];
    float M[PPT], S[PPT], acc[PPT];
    const float* d1[PPT];
    const float* d2[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        x[k] = threadIdx.x + k * blockDim.x;
        live[k] = x[k] < W;
        const int xc = live[k] ? x[k] : W - 1;
        d1[k] = D + soff(p.ds, db1 + b, 0, y, xc);
        d2[k] = D + soff(p.ds, db2 + b, 0, y, xc);
        M[k] = -INFINITY, S[k] = 0.0f, acc[k] = 0.0f;
    }
    // sample of the row of plane n at u (two taps, zero padding, optional mirror), in log2 units
    auto sample = [&](const float* r, float u) -> float {
        u = fminf(fmaxf(u, -2.0f), Wp1);
        const int x0 = __float2int_rd(u);
        const float f0 = (float)x0;
        const float w1 = u - f0;
        const float a = ((unsigned)x0 < (unsigned)W) ? __ldg(r + (FLIP ? W - 1 - x0 : x0)) : 0.0f;
        const float c = ((unsigned)(x0 + 1) < (unsigned)W) ? __ldg(r + (FLIP ? W - 2 - x0 : x0 + 1)) : 0.0f;
        return fmaf(c, w1, a * (1.0f - w1)) * kLog2e;
    };
    for (int n = 0; n < N; ++n) {
        const float* r = src + (int64_t)n * p.hw;
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const float l2 = sample(r, (float)x[k] + sign1 * __ldg(d1[k] + (int64_t)n * p.ds.n));
            const float mn = fmaxf(M[k], l2);
            S[k] = fmaf(S[k], fast_exp2(M[k] - mn), fast_exp2(l2 - mn));
            M[k] = mn;
        }
    }
    float invS[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) invS[k] = 1.0f / S[k];
    __syncthreads();  // the zero fill above
    float* e = erow + OC_PAD;
    for (int n = 0; n < N; ++n, e = (e == erow + OC_PAD) ? erow + pitch + OC_PAD : erow + OC_PAD) {
        const float* r = src + (int64_t)n * p.hw;
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const float l2 = sample(r, (float)x[k] + sign1 * __ldg(d1[k] + (int64_t)n * p.ds.n));
            if (live[k]) e[x[k]] = fast_exp2(l2 - M[k]) * invS[k];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            float u = (float)x[k] + sign2 * __ldg(d2[k] + (int64_t)n * p.ds.n);
            u = fminf(fmaxf(u, -2.0f), Wp1);
            const int x0 = __float2int_rd(u);  // in [-2, W+1]: both taps inside the padded row
            const float f0 = (float)x0;
            const float w1 = u - f0;
            acc[k] = fmaf(e[x0 + 1], w1, fmaf(e[x0], 1.0f - w1, acc[k]));
        }
    }
#pragma unroll
    for (int k = 0; k < PPT; ++k)
        if (live[k]) out[(int64_t)row * W + x[k]] = fminf(acc[k], 1.0f);  // o[o > 1] = 1
}

// trainer.py:456-459
__global__ void __launch_bounds__(256) disp_pp_kernel(const OcclParams p, const float* __restrict__ disp, const float* __restrict__ o_l,
                                                      const float* __restrict__ o_fr, float* __restrict__ disp_pp) {
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= (int64_t)p.B * p.hw) return;
    const int b = (int)(pix / p.hw);
    const int rem = (int)(pix - (int64_t)b * p.hw);
    const int y = rem / p.W, x = rem - y * p.W;
    const float dl = __ldg(disp + pix);
    const float df = __ldg(disp + (int64_t)(p.B + b) * p.hw + (int64_t)y * p.W + (p.W - 1 - x));
    const float ofr = o_fr[pix], ol = o_l[pix];
    const float mean = dl * 0.5f + df * 0.5f;
    float v = mean * ofr + dl * (1.0f - ofr);
    v = v * ol + df * (1.0f - ol);
    disp_pp[pix] = v;
}

}  // namespace oc
}  // namespace pd
