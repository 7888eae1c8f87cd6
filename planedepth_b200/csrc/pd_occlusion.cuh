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
    const float f0 = floorf(u);
    const int x0 = (int)f0;
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
