// Decoder tail (networks/depth_decoder.py:258-291, render_probability off): everything the
// DepthDecoder does after its dispconv / sigmaconv convolutions,
//     logits = raw * mask;  pi = softmax_n(logits);
//     mixture: sigma = clamp(sigmoid(sraw), 0.01, 1);  w = pi / sigma * mask;  probability = w / sum_n w   (else probability = pi)
//     disp = sum_n probability * disp_layered;  depth = 0.1 * 0.58 * W / disp
// as one forward and one backward kernel (one thread per pixel, planes in loops; softmax statistics saved) instead of
// ~10 (plain) / ~20 (mixture) elementwise passes over [B,N,H,W] tensors each way.
#pragma once
#include "pd_device.cuh"

namespace pd {
namespace tl {

struct TailParams {
    int B, N, H, W, mask_dtype;
    int64_t hw;
    pd_strides4 ds, ms, gds;
    float depth_c;  // 0.1 * 0.58 * W
    // forward
    const float* raw;
    const float* sraw;
    const float* disp_layered;
    const void* mask;
    float* logits;
    float* sigma;
    float* prob;
    float* pi;     // optional
    float* disp;
    float* depth;  // optional
    float* stats;  // [B,3,H,W]: max logit * log2(e), sum exp, sum of mixture weights
    // backward
    const float* g_logits;
    const float* g_sigma;
    const float* g_prob;
    const float* g_disp;
    const float* g_depth;
    float* g_raw;
    float* g_sraw;
    float* g_dl;
    int g_dl_dense, warp_rows;
    int smem_acc;  // backward: sum the fully compact ([B,N,1,1]) disparity gradient per CTA in shared memory
};

__device__ __forceinline__ float sigmoid_clamped(float x) {
    const float s = 1.0f / (1.0f + fast_exp(-x));
    return fminf(fmaxf(s, 0.01f), 1.0f);  // depth_decoder.py:279-280
}

// Both kernels: one thread per pixel, T = blockDim.x pixels per CTA, the planes in loops.  A [N][T] column cache in
// dynamic shared memory keeps the per-plane value the later passes need (forward: logit -> mixture weight; backward:
// softmax probability), so every [B,N,H,W] input is read from global memory once.
template <bool MIX>
__global__ void __launch_bounds__(256) tail_fwd_kernel(const TailParams p) {
    extern __shared__ float col[];  // [N][T]
    const int T = blockDim.x;
    const int64_t pix = (int64_t)blockIdx.x * T + threadIdx.x;
    if (pix >= (int64_t)p.B * p.hw) return;
    const int b = (int)(pix / p.hw);
    const int rem = (int)(pix - (int64_t)b * p.hw);
    const int y = rem / p.W, x = rem - y * p.W;
    const int64_t base = (int64_t)b * p.N * p.hw + rem;
    const int64_t m0 = soff(p.ms, b, 0, y, x), d0 = soff(p.ds, b, 0, y, x);
    float* c = col + threadIdx.x;
    float M = -INFINITY, S = 0.0f;
#pragma unroll 8
    for (int n = 0; n < p.N; ++n) {
        const float m = load_mask(p.mask, p.mask_dtype, m0 + (int64_t)n * p.ms.n);
        const float l = __ldg(p.raw + base + (int64_t)n * p.hw) * m;
        p.logits[base + (int64_t)n * p.hw] = l;
        const float l2 = l * kLog2e, mn = fmaxf(M, l2);
        c[n * T] = l2;
        S = fmaf(S, fast_exp2(M - mn), fast_exp2(l2 - mn));
        M = mn;
    }
    const float invS = 1.0f / S;
    float Z = 1.0f;
    if (MIX) {
        Z = 0.0f;
    #pragma unroll 8
    for (int n = 0; n < p.N; ++n) {
            const float m = load_mask(p.mask, p.mask_dtype, m0 + (int64_t)n * p.ms.n);
            const float pi = fast_exp2(c[n * T] - M) * invS;
            const float sg = sigmoid_clamped(__ldg(p.sraw + base + (int64_t)n * p.hw));
            p.sigma[base + (int64_t)n * p.hw] = sg;
            if (p.pi) p.pi[base + (int64_t)n * p.hw] = pi;
            const float w = pi / sg * m;
            c[n * T] = w;
            Z += w;
        }
    }
    const float invZ = 1.0f / Z;
    float dsum = 0.0f;
#pragma unroll 8
    for (int n = 0; n < p.N; ++n) {
        const float pr = MIX ? c[n * T] * invZ : fast_exp2(c[n * T] - M) * invS;
        p.prob[base + (int64_t)n * p.hw] = pr;
        dsum = fmaf(pr, __ldg(p.disp_layered + d0 + (int64_t)n * p.ds.n), dsum);
    }
    p.disp[pix] = dsum;
    if (p.depth) p.depth[pix] = p.depth_c / dsum;
    float* st = p.stats + (int64_t)b * 3 * p.hw + rem;
    st[0] = M, st[p.hw] = S, st[2 * p.hw] = Z;
}

// Inputs of the recomputation: the saved logits (= raw * mask), sigma (clamped), statistics.
template <bool MIX>
__global__ void __launch_bounds__(256) tail_bwd_kernel(const TailParams p) {
    // d L / d disp_layered of the decoder's [B,N,1,1] disparities reduces over all pixels of an image: summed per CTA in
    // shared memory and flushed with N atomics per CTA (a global atomic per warp and plane serialises on a handful of L2
    // lines: 1.4 ms instead of 0.7 ms at cfg 2).  The host enables it only when hw % T == 0 (a CTA inside one image).
    extern __shared__ float col[];  // [N][T] softmax probabilities, then [N] accumulators
    const int T = blockDim.x;
    float* gacc = col + (size_t)p.N * T;
    const bool compact = p.smem_acc != 0;
    if (compact) {
        for (int i = threadIdx.x; i < p.N; i += T) gacc[i] = 0.0f;
        __syncthreads();
    }
    const int64_t total = (int64_t)p.B * p.hw;
    const int64_t pixr = (int64_t)blockIdx.x * T + threadIdx.x;
    const bool live = pixr < total;
    const int64_t pix = live ? pixr : total - 1;  // whole warps stay for the shuffles below
    const int b = (int)(pix / p.hw);
    const int rem = (int)(pix - (int64_t)b * p.hw);
    const int y = rem / p.W, x = rem - y * p.W;
    const int lane = threadIdx.x & 31;
    const int64_t base = (int64_t)b * p.N * p.hw + rem;
    const int64_t m0 = soff(p.ms, b, 0, y, x), d0 = soff(p.ds, b, 0, y, x);
    float* c = col + threadIdx.x;
    const float* st = p.stats + (int64_t)b * 3 * p.hw + rem;
    const float M = __ldg(st), invS = 1.0f / __ldg(st + p.hw), invZ = 1.0f / __ldg(st + 2 * p.hw);
    float gd = p.g_disp ? __ldg(p.g_disp + pix) : 0.0f;
    if (p.g_depth) {
        const float dv = __ldg(p.disp + pix);
        gd -= __ldg(p.g_depth + pix) * p.depth_c / (dv * dv);  // depth = c / disp
    }
    if (!live) gd = 0.0f;
    const bool has_gp = p.g_prob != nullptr && live;
    // pass 1: dotp = sum_k probability_k * gp_k,  gp_k = g_prob_k + gd * disp_layered_k
    float dotp = 0.0f;
#pragma unroll 8
    for (int n = 0; n < p.N; ++n) {
        const int64_t o = base + (int64_t)n * p.hw;
        const float pi = fast_exp2(fmaf(__ldg(p.logits + o), kLog2e, -M)) * invS;
        c[n * T] = pi;
        float pr = pi;
        if (MIX) pr = pi / __ldg(p.sigma + o) * load_mask(p.mask, p.mask_dtype, m0 + (int64_t)n * p.ms.n) * invZ;
        const float gp = (has_gp ? __ldg(p.g_prob + o) : 0.0f) + gd * __ldg(p.disp_layered + d0 + (int64_t)n * p.ds.n);
        dotp = fmaf(pr, gp, dotp);
    }
    // (mixture) the softmax-normaliser term sum_k pi_k g_pi_k, g_pi_k = (gp_k - dotp) m_k / (sigma_k Z), vanishes identically:
    // it equals sum_k probability_k (gp_k - dotp) = dotp - dotp — probability = w / sum w does not change when pi is rescaled
    // pass 2: gradients
#pragma unroll 8
    for (int n = 0; n < p.N; ++n) {
        const int64_t o = base + (int64_t)n * p.hw;
        const float m = load_mask(p.mask, p.mask_dtype, m0 + (int64_t)n * p.ms.n);
        const float pi = c[n * T];
        const float gp = (has_gp ? __ldg(p.g_prob + o) : 0.0f) + gd * __ldg(p.disp_layered + d0 + (int64_t)n * p.ds.n);
        const float glo = (p.g_logits && live) ? __ldg(p.g_logits + o) : 0.0f;
        float pr = pi, gl;
        if (MIX) {
            const float sg = __ldg(p.sigma + o), a = 1.0f / sg;
            pr = pi * a * m * invZ;
            const float gw = (gp - dotp) * invZ;  // d / d w_n with probability = w / sum w
            const float gpi = gw * m * a;
            gl = pi * gpi + glo;
            const float gsg = -gw * pi * m * a * a + ((p.g_sigma && live) ? __ldg(p.g_sigma + o) : 0.0f);
            // clamp passes the gradient inside [0.01, 1]; sigmoid' = s (1 - s) with s = sigma there
            if (live && p.g_sraw) p.g_sraw[o] = (sg > 0.01f) ? gsg * sg * (1.0f - sg) : 0.0f;
        } else {
            gl = pi * (gp - dotp) + glo;
        }
        if (live && p.g_raw) p.g_raw[o] = gl * m;  // logits = raw * mask
        if (p.g_dl) {
            const float gdl = live ? gd * pr : 0.0f;
            float* dst = p.g_dl + soff(p.gds, b, n, y, x);
            if (p.g_dl_dense) {
                if (live) *dst = gdl;
            } else if (compact) {
                const float s = warp_sum(gdl);
                if (lane == 0 && s != 0.0f) atomicAdd(gacc + n, s);
            } else if (p.gds.x == 0 && p.warp_rows) {
                const float s = warp_sum(gdl);
                if (lane == 0 && s != 0.0f) atomicAdd(dst, s);
            } else if (gdl != 0.0f) {
                atomicAdd(dst, gdl);
            }
        }
    }
    if (compact) {
        __syncthreads();
        for (int i = threadIdx.x; i < p.N; i += T) {
            const float v = gacc[i];
            if (v != 0.0f) atomicAdd(p.g_dl + soff(p.gds, b, i, 0, 0), v);
        }
    }
}

}  // namespace tl
}  // namespace pd
