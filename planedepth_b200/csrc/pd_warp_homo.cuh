// Fast path of the homography warp (layers.py:206-234 + trainer.py:556-603): same arithmetic as the
// reference-faithful kernels of pd_warp_general.cuh, re-organised for instruction count:
//   * the per-plane parameters (H_t2s, R·n) sit in shared memory (three 128-bit broadcast loads per
//     plane instead of 21 global loads per sample); the pixel's ray inv_K·(x,y,1) is formed once;
//   * the source colour is packed to one rgbx float4 per pixel by a small pre-pass, so a sample's
//     12 colour taps are four 128-bit loads;
//   * zero padding by clamped addresses and zeroed weights: no predicated loads;
//   * q_xy / max(q_z, 1e-7) through one Newton-refined reciprocal and a residual correction (correctly
//     rounded quotients without the IEEE-division sequence); the normalise / un-normalise round trip
//     keeps the reference's fp32 rounding (division-free form where the size allows it);
//   * backward: the 9 sums of dL/dH per (plane, warp) shrink to 6 (a warp works inside one row, so the
//     y-weighted sums are y times the plain ones) and are reduced with a 9-shuffle reduce-scatter
//     instead of 45 butterfly shuffles, accumulated per WARP in shared memory (plain read-modify-writes:
//     a shared float atomicAdd is a compare-and-swap loop), summed and flushed once per CTA;
//   * per-sample overhead: the image index is CTA-uniform (blockIdx), the round-trip variant is a launch-time
//     template parameter, floor() is one float->int conversion, and both passes instantiate the sample body
//     twice: all four taps inside the image (raw weights, no per-tap tests) and the general clamped form.
// PD_FLAG_EXACT_COORDS keeps the general kernels (IEEE divisions).
#pragma once
#include <type_traits>

#include "pd_warp_general.cuh"

namespace pd {
namespace hm {

constexpr int HT = 256;  // threads per CTA = consecutive pixels of one image

__global__ void __launch_bounds__(256) pack_rgbx_kernel(const float* __restrict__ src, float4* __restrict__ dst, int64_t hw, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t b = i / hw, o = i - b * hw;
    const float* s = src + b * 3 * hw + o;
    dst[i] = make_float4(__ldg(s), __ldg(s + hw), __ldg(s + 2 * hw), 0.0f);
}

struct HTaps {
    int o00, o01, o10, o11;      // pixel offsets of the four (clamped) taps inside a plane
    float w00, w01, w10, w11;    // bilinear weights, zero for taps outside the image (padding_mode="zeros")
    float rx0, rx1, ry0, ry1;    // raw per-axis weights as ATen forms them: rx0 = (x0+1) - x, rx1 = x - x0
    bool ix0, ix1, iy0, iy1;     // column x0 / x0+1, row y0 / y0+1 inside the image
};

// FAST: both image sizes admit the division-free round trip (rows_rcp() != 0), decided once per launch
template <bool FAST>
__device__ __forceinline__ float rt(float p, float size_m1, float rcp) {
    if constexpr (FAST) return roundtrip_fast(p, size_m1, rcp);
    else return (rcp != 0.0f) ? roundtrip_fast(p, size_m1, rcp) : roundtrip(p, size_m1);
}

// ATen grid_sampler_2d (bilinear, zeros padding) taps of the sample position (x, y), cf. make_taps()
__device__ __forceinline__ HTaps make_htaps(float x, float y, int W, int H, float Wp1, float Hp1) {
    x = fminf(fmaxf(x, -2.0f), Wp1);  // Wp1 = W + 1, Hp1 = H + 1 as floats (hoisted by the caller)
    y = fminf(fmaxf(y, -2.0f), Hp1);
    // one float -> int conversion per axis; the integer goes back to float on the ALU pipe (exact: |x0| <= W + 1)
    const int x0 = __float2int_rd(x), y0 = __float2int_rd(y);
    const float fx0 = (float)x0, fy0 = (float)y0;
    HTaps t;
    t.rx1 = x - fx0, t.rx0 = (fx0 + 1.0f) - x;
    t.ry1 = y - fy0, t.ry0 = (fy0 + 1.0f) - y;
    if ((unsigned)x0 < (unsigned)(W - 1) && (unsigned)y0 < (unsigned)(H - 1)) {  // all four taps inside (the common case)
        t.ix0 = t.ix1 = t.iy0 = t.iy1 = true;
        t.o00 = y0 * W + x0, t.o01 = t.o00 + 1, t.o10 = t.o00 + W, t.o11 = t.o10 + 1;
        t.w00 = t.rx0 * t.ry0, t.w01 = t.rx1 * t.ry0, t.w10 = t.rx0 * t.ry1, t.w11 = t.rx1 * t.ry1;
        return t;
    }
    t.ix0 = (unsigned)x0 < (unsigned)W, t.ix1 = (unsigned)(x0 + 1) < (unsigned)W;
    t.iy0 = (unsigned)y0 < (unsigned)H, t.iy1 = (unsigned)(y0 + 1) < (unsigned)H;
    const int xa = min(max(x0, 0), W - 1), xb = min(max(x0 + 1, 0), W - 1);
    const int ya = min(max(y0, 0), H - 1) * W, yb = min(max(y0 + 1, 0), H - 1) * W;
    t.o00 = ya + xa, t.o01 = ya + xb, t.o10 = yb + xa, t.o11 = yb + xb;
    t.w00 = (t.ix0 && t.iy0) ? t.rx0 * t.ry0 : 0.0f;
    t.w01 = (t.ix1 && t.iy0) ? t.rx1 * t.ry0 : 0.0f;
    t.w10 = (t.ix0 && t.iy1) ? t.rx0 * t.ry1 : 0.0f;
    t.w11 = (t.ix1 && t.iy1) ? t.rx1 * t.ry1 : 0.0f;
    return t;
}

// ATen accumulates nw, ne, sw, se in that order (blend() of pd_device.cuh)
__device__ __forceinline__ float hblend(float nw, float ne, float sw, float se, const HTaps& t) {
    float acc = nw * t.w00;
    acc = fmaf(ne, t.w01, acc);
    acc = fmaf(sw, t.w10, acc);
    acc = fmaf(se, t.w11, acc);
    return acc;
}

struct HCoord {
    float u, v, m, zinv, dz;  // sample position (before the round trip), validity mask, 1/zc, d zc / d qz
};

// layers.py:221-228 for one pixel (fx, fy) with ray (rx, ry, rz) = inv_K (x, y, 1) and the plane's 12 parameters
__device__ __forceinline__ HCoord homo_coords(const float4 h0, const float4 h1, const float4 h2, float fx, float fy, float rx, float ry, float rz) {
    const float qx = fmaf(h0.y, fy, h0.x * fx) + h0.z;
    const float qy = fmaf(h1.x, fy, h0.w * fx) + h1.y;
    const float qz = fmaf(h1.w, fy, h1.z * fx) + h2.x;
    const float facing = rx * h2.y + ry * h2.z + rz * h2.w;
    HCoord c;
    c.m = ((facing > 0.0f) && (qz > 1e-7f)) ? 1.0f : 0.0f;
    const float zc = (qz < 1e-7f) ? 1e-7f : qz;
    c.dz = (qz < 1e-7f) ? 0.0f : 1.0f;
    const float r0 = fast_rcp(zc);
    c.zinv = fmaf(r0, fmaf(-zc, r0, 1.0f), r0);  // one Newton step: ~0.5 ulp
    // quotients with a residual correction (Markstein): the correctly rounded q / zc of layers.py:227-228 in all but
    // pathological cases, so that floor() of the sample position falls as it does in the reference
    const float u0 = qx * c.zinv, v0 = qy * c.zinv;
    c.u = fmaf(fmaf(-u0, zc, qx), c.zinv, u0);
    c.v = fmaf(fmaf(-v0, zc, qy), c.zinv, v0);
    return c;
}

__device__ __forceinline__ void load_plane_params(const float* sh, int n, float4& h0, float4& h1, float4& h2) {
    const float4* q = reinterpret_cast<const float4*>(sh) + 3 * n;
    h0 = q[0], h1 = q[1], h2 = q[2];
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <bool MIX, bool FASTRT>
__global__ void __launch_bounds__(HT) homo_fwd_kernel(const WarpParams p, const float4* __restrict__ rgbx, float rcp_w, float rcp_h) {
    extern __shared__ __align__(16) float sh[];  // [N][12]
    const int N = p.d.N, W = p.d.W, H = p.d.H;
    // hw % HT == 0: a CTA stays inside one image, so the image index (and every per-image base pointer) is CTA-uniform
    const int b = (int)(((int64_t)blockIdx.x * HT) / p.hw);
    const int rem = (int)((int64_t)blockIdx.x * HT - (int64_t)b * p.hw) + (int)threadIdx.x;
    const int64_t pix = (int64_t)b * p.hw + rem;
    const int y = rem / W, x = rem - y * W;
    const float Wp1 = (float)(W + 1), Hp1 = (float)(H + 1);
    const unsigned Wm1 = (unsigned)(W - 1), Hm1 = (unsigned)(H - 1);
    for (int i = threadIdx.x; i < N * 12; i += HT) sh[i] = __ldg(p.in.hmat + (int64_t)b * N * 12 + i);
    __syncthreads();
    const float fx = (float)x, fy = (float)y;
    const float* ik = p.in.cam + (int64_t)b * 9;
    const float rx = fmaf(__ldg(ik + 1), fy, __ldg(ik + 0) * fx) + __ldg(ik + 2);
    const float ry = fmaf(__ldg(ik + 4), fy, __ldg(ik + 3) * fx) + __ldg(ik + 5);
    const float rz = fmaf(__ldg(ik + 7), fy, __ldg(ik + 6) * fx) + __ldg(ik + 8);

    float tr = 0, tg = 0, tb = 0, err_auto = 0;
    if (MIX) {
        const float* tp = p.in.tgt + (int64_t)b * p.chw3 + rem;
        tr = __ldg(tp), tg = __ldg(tp + p.hw), tb = __ldg(tp + 2 * p.hw);
        if (p.d.automask) {
            const float4 s = __ldg(rgbx + pix);
            err_auto = (fabsf(s.x - tr) + fabsf(s.y - tg) + fabsf(s.z - tb)) * (1.0f / 3.0f);
        }
    }
    const float4* src = rgbx + (int64_t)b * p.hw;
    const float* lg = p.in.logits + (int64_t)b * N * p.hw;
    const float* sgp = MIX ? p.in.sigma + (int64_t)b * N * p.hw : nullptr;

    // online softmax over planes in base 2: every accumulator is a sum of exp2(l2_n - Mx) * something
    float Mx = -INFINITY, S = 0, A = 0, R0 = 0, R1 = 0, R2 = 0, Q = 0, Qa = 0;
    for (int n = 0; n < N; ++n, lg += p.hw) {
        float4 h0, h1, h2;
        load_plane_params(sh, n, h0, h1, h2);
        const HCoord c = homo_coords(h0, h1, h2, fx, fy, rx, ry, rz);
        float cr = 0.0f, cg = 0.0f, cb = 0.0f, l2 = 0.0f, sraw = 0.0f;  // a masked plane enters with logit 0, colour 0 (:580-583)
        if (c.m != 0.0f) {
            float su = rt<FASTRT>(c.u, p.wm1, rcp_w), sv = rt<FASTRT>(c.v, p.hm1, rcp_h);
            su = fminf(fmaxf(su, -2.0f), Wp1);
            sv = fminf(fmaxf(sv, -2.0f), Hp1);
            // one float -> int conversion per axis; the integer goes back to float on the ALU pipe (exact: |x0| <= W + 1)
            const int x0 = __float2int_rd(su), y0 = __float2int_rd(sv);
            const float fx0 = (float)x0, fy0 = (float)y0;
            if ((unsigned)x0 < Wm1 && (unsigned)y0 < Hm1) {
                // all four taps inside the image (the common case): one base offset, immediate / row offsets, raw weights
                const float wx1 = su - fx0, wx0 = (fx0 + 1.0f) - su, wy1 = sv - fy0, wy0 = (fy0 + 1.0f) - sv;
                const float w00 = wx0 * wy0, w01 = wx1 * wy0, w10 = wx0 * wy1, w11 = wx1 * wy1;
                const int o = y0 * W + x0;
                const float4* s0 = src + o;
                const float* g0 = lg + o;
                const float4 a = __ldg(s0), bq = __ldg(s0 + 1), cq = __ldg(s0 + W), d = __ldg(s0 + W + 1);
                const float l00 = __ldg(g0), l01 = __ldg(g0 + 1), l10 = __ldg(g0 + W), l11 = __ldg(g0 + W + 1);
                cr = fmaf(d.x, w11, fmaf(cq.x, w10, fmaf(bq.x, w01, a.x * w00)));
                cg = fmaf(d.y, w11, fmaf(cq.y, w10, fmaf(bq.y, w01, a.y * w00)));
                cb = fmaf(d.z, w11, fmaf(cq.z, w10, fmaf(bq.z, w01, a.z * w00)));
                l2 = fmaf(l11, w11, fmaf(l10, w10, fmaf(l01, w01, l00 * w00))) * kLog2e;
                if (MIX) {
                    const float* q0 = sgp + (int64_t)n * p.hw + o;
                    sraw = fmaf(__ldg(q0 + W + 1), w11, fmaf(__ldg(q0 + W), w10, fmaf(__ldg(q0 + 1), w01, __ldg(q0) * w00)));
                }
            } else {
                const HTaps t = make_htaps(su, sv, W, H, Wp1, Hp1);
                const float4 a = __ldg(src + t.o00), bq = __ldg(src + t.o01), cq = __ldg(src + t.o10), d = __ldg(src + t.o11);
                const float l00 = __ldg(lg + t.o00), l01 = __ldg(lg + t.o01), l10 = __ldg(lg + t.o10), l11 = __ldg(lg + t.o11);
                cr = hblend(a.x, bq.x, cq.x, d.x, t);
                cg = hblend(a.y, bq.y, cq.y, d.y, t);
                cb = hblend(a.z, bq.z, cq.z, d.z, t);
                l2 = hblend(l00, l01, l10, l11, t) * kLog2e;
                if (MIX) {
                    const float* sp = sgp + (int64_t)n * p.hw;
                    sraw = hblend(__ldg(sp + t.o00), __ldg(sp + t.o01), __ldg(sp + t.o10), __ldg(sp + t.o11), t);
                }
            }
        }
        const float mnew = fmaxf(Mx, l2);
        const float sc = fast_exp2(Mx - mnew);
        const float e = fast_exp2(l2 - mnew);
        Mx = mnew;
        S = fmaf(S, sc, e);
        if (MIX) {
            const float sg = fminf(fmaxf(sraw, 0.01f), 1.0f);  // trainer.py:597
            const float inv = 1.0f / sg;
            const float es = e * inv;
            A = fmaf(A, sc, es);
            R0 = fmaf(R0, sc, es * cr);
            R1 = fmaf(R1, sc, es * cg);
            R2 = fmaf(R2, sc, es * cb);
            const float err = (fabsf(cr - tr) + fabsf(cg - tg) + fabsf(cb - tb)) * (1.0f / 3.0f);
            Q = fmaf(Q, sc, e * (0.5f * fast_exp(-err * inv) * inv));  // layers.py:454-455
            Qa = fmaf(Qa, sc, e * (0.5f * fast_exp(-err_auto * inv) * inv));
        } else {
            R0 = fmaf(R0, sc, e * cr);
            R1 = fmaf(R1, sc, e * cg);
            R2 = fmaf(R2, sc, e * cb);
        }
    }
    const float invS = 1.0f / S;
    const float invD = MIX ? 1.0f / A : invS;
    float* rr = p.out.rgb_rec + (int64_t)b * p.chw3 + rem;
    rr[0] = R0 * invD;
    rr[p.hw] = R1 * invD;
    rr[2 * p.hw] = R2 * invD;
    float* st = p.out.stats + (int64_t)b * (MIX ? PD_STATS_MIXTURE : PD_STATS_PLAIN) * p.hw + rem;
    st[0] = Mx;  // reference logit in log2 units (shared convention of all kernels)
    st[p.hw] = S;
    if (MIX) {
        const float D = Q * invS + 1e-7f;  // layers.py:466
        st[2 * p.hw] = A;
        st[3 * p.hw] = D;
        p.out.nll[pix] = -logf(D);
        if (p.d.automask) p.out.nll_auto[pix] = -logf(Qa * invS + 1e-7f);
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void hscatter(float* __restrict__ plane, const HTaps& t, float g) {
    if (t.w00 != 0.0f) atomicAdd(plane + t.o00, g * t.w00);
    if (t.w01 != 0.0f) atomicAdd(plane + t.o01, g * t.w01);
    if (t.w10 != 0.0f) atomicAdd(plane + t.o10, g * t.w10);
    if (t.w11 != 0.0f) atomicAdd(plane + t.o11, g * t.w11);
}

// Sum of 8 per-lane values over the warp with a reduce-scatter: after the call lane L with (L & 3) == 0 holds in the
// return value the warp total of v[(L >> 2) & 7].  9 shuffles instead of 40.
__device__ __forceinline__ float warp_reduce_scatter8(const float (&v)[8]) {
    const int lane = threadIdx.x & 31;
    float a[4];
    {
        const bool up = lane & 16;  // upper half keeps 4..7
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float send = up ? v[i] : v[4 + i];
            const float keep = up ? v[4 + i] : v[i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    float c[2];
    {
        const bool up = lane & 8;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float send = up ? a[i] : a[2 + i];
            const float keep = up ? a[2 + i] : a[i];
            c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    float d;
    {
        const bool up = lane & 4;
        const float send = up ? c[0] : c[1];
        const float keep = up ? c[1] : c[0];
        d = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;  // value index = 4*bit4 + 2*bit3 + bit2 of the lane
}

// FUSED: the upstream gradients arrive in pd_warp_grad_out's fused form (formed here from the photometric forward's unit
// gradient); otherwise the prologue is the plain load of g_rgb_rec / g_nll.  Two instantiations: the generic prologue
// (pointer tests, 64-bit offsets kept live) changed the register allocation of the plane loop and cost 10 % at cfg 4.
template <bool MIX, bool FUSED, bool FASTRT>
__global__ void __launch_bounds__(HT) homo_bwd_kernel(const WarpParams p, const float4* __restrict__ rgbx, float rcp_w, float rcp_h) {
    extern __shared__ __align__(16) float sh[];  // [N][12] parameters, then [HT / 32][N][9] dL/dH accumulators, one set per warp
    const int N = p.d.N, W = p.d.W, H = p.d.H;
    float* gacc0 = sh + N * 12;
    float* gacc = gacc0 + (threadIdx.x >> 5) * N * 9;  // private to the warp: plain read-modify-write, no shared atomics (CAS loops)
    const int b = (int)(((int64_t)blockIdx.x * HT) / p.hw);  // CTA-uniform, see homo_fwd_kernel
    const int rem = (int)((int64_t)blockIdx.x * HT - (int64_t)b * p.hw) + (int)threadIdx.x;
    const int64_t pix = (int64_t)b * p.hw + rem;
    const int y = rem / W, x = rem - y * W;  // W % 32 == 0: a warp stays inside one row
    const float Wp1 = (float)(W + 1), Hp1 = (float)(H + 1);
    const unsigned Wm1 = (unsigned)(W - 1), Hm1 = (unsigned)(H - 1);
    const int lane = threadIdx.x & 31;
    const bool want_h = p.gin.g_hmat != nullptr;
    for (int i = threadIdx.x; i < N * 12; i += HT) sh[i] = __ldg(p.in.hmat + (int64_t)b * N * 12 + i);
    for (int i = threadIdx.x; i < (HT / 32) * N * 9; i += HT) gacc0[i] = 0.0f;
    __syncthreads();
    const float fx = (float)x, fy = (float)y;
    const float* ik = p.in.cam + (int64_t)b * 9;
    const float rx = fmaf(__ldg(ik + 1), fy, __ldg(ik + 0) * fx) + __ldg(ik + 2);
    const float ry = fmaf(__ldg(ik + 4), fy, __ldg(ik + 3) * fx) + __ldg(ik + 5);
    const float rz = fmaf(__ldg(ik + 7), fy, __ldg(ik + 6) * fx) + __ldg(ik + 8);

    float g0, g1, g2, gn_up = 0.0f;
    if (FUSED) {
        // common case first, shaped like the plain prologue: three loads of the unit gradient times one scalar; the optional
        // extra terms (perceptual gradient, a genuine g_rgb_rec / g_nll on top) sit behind one uniform branch
        const float gph = upstream_scale(p);
        g0 = g1 = g2 = 0.0f;
        if (p.gout.g_unit) {
            const float* gp = p.gout.g_unit + (int64_t)b * p.chw3 + rem;
            g0 = gph * __ldg(gp), g1 = gph * __ldg(gp + p.hw), g2 = gph * __ldg(gp + 2 * p.hw);
        }
        if (MIX && p.gout.g_unit_nll) gn_up = gph * __ldg(p.gout.g_unit_nll + pix);
        if (p.gout.g_pred || p.gout.g_rgb_rec || p.gout.g_nll) {
            const int64_t gi = (int64_t)b * p.chw3 + rem;
            if (p.gout.g_pred) {
                const float m = p.gout.mask_novel ? __ldg(p.gout.mask_novel + pix) : 1.0f;
                g0 = fmaf(__ldg(p.gout.g_pred + gi), m, g0), g1 = fmaf(__ldg(p.gout.g_pred + gi + p.hw), m, g1);
                g2 = fmaf(__ldg(p.gout.g_pred + gi + 2 * p.hw), m, g2);
            }
            if (p.gout.g_rgb_rec) {
                g0 += __ldg(p.gout.g_rgb_rec + gi), g1 += __ldg(p.gout.g_rgb_rec + gi + p.hw), g2 += __ldg(p.gout.g_rgb_rec + gi + 2 * p.hw);
            }
            if (MIX && p.gout.g_nll) gn_up += __ldg(p.gout.g_nll + pix);
        }
    } else {
        const float* gp = p.gout.g_rgb_rec + (int64_t)b * p.chw3 + rem;
        g0 = __ldg(gp), g1 = __ldg(gp + p.hw), g2 = __ldg(gp + 2 * p.hw);
        if (MIX) gn_up = p.gout.g_nll ? __ldg(p.gout.g_nll + pix) : 0.0f;
    }
    const float* rp = p.out.rgb_rec + (int64_t)b * p.chw3 + rem;
    const float Gbar = g0 * __ldg(rp) + g1 * __ldg(rp + p.hw) + g2 * __ldg(rp + 2 * p.hw);
    const float* st = p.out.stats + (int64_t)b * (MIX ? PD_STATS_MIXTURE : PD_STATS_PLAIN) * p.hw + rem;
    const float Ml2 = __ldg(st), Sv = __ldg(st + p.hw), invS = 1.0f / Sv;
    float tr = 0, tg = 0, tb = 0, Zinv = 0, gD = 0, gDD = 0;
    if (MIX) {
        const float* tp = p.in.tgt + (int64_t)b * p.chw3 + rem;
        tr = __ldg(tp), tg = __ldg(tp + p.hw), tb = __ldg(tp + 2 * p.hw);
        Zinv = Sv / __ldg(st + 2 * p.hw);  // 1/Z, Z = sum pi/sigma = A/S
        const float D = __ldg(st + 3 * p.hw);
        const float gn = gn_up;
        gD = -gn / D;            // d loss / d D,  nll = -log D
        gDD = gD * (D - 1e-7f);  // = sum_k pi_k P_k
    }
    const float4* src = rgbx + (int64_t)b * p.hw;
    const float* lg = p.in.logits + (int64_t)b * N * p.hw;
    const float* sgp = MIX ? p.in.sigma + (int64_t)b * N * p.hw : nullptr;
    float* glg = p.gin.g_logits ? p.gin.g_logits + (int64_t)b * N * p.hw : nullptr;
    float* gsg = (MIX && p.gin.g_sigma) ? p.gin.g_sigma + (int64_t)b * N * p.hw : nullptr;

    // per-plane base pointers advance by one plane per iteration (no 64-bit multiply per access)
    for (int n = 0; n < N; ++n, lg += p.hw, sgp += MIX ? p.hw : 0, glg += glg ? p.hw : 0, gsg += gsg ? p.hw : 0) {
        float4 h0, h1, h2;
        load_plane_params(sh, n, h0, h1, h2);
        const HCoord c = homo_coords(h0, h1, h2, fx, fy, rx, ry, rz);
        float gqx = 0.0f, gqy = 0.0f, gqz = 0.0f;
        // One sample: loads, blends, the gradients of the logit / sigma / colour taps, the scatter and (want_h) the coordinate
        // gradient.  Instantiated twice: ALL_IN = every tap inside the image (the common case: raw weights, no per-tap selects,
        // unconditional reductions) and the general form with clamped addresses and zeroed weights.
        auto sample = [&](const HTaps& t, auto all_in) {
            constexpr bool ALL_IN = decltype(all_in)::value;
            float dl = 0.0f, dsg = 0.0f;
            const float4 a = __ldg(src + t.o00), bq = __ldg(src + t.o01), cq = __ldg(src + t.o10), d = __ldg(src + t.o11);
            const float l00 = __ldg(lg + t.o00), l01 = __ldg(lg + t.o01), l10 = __ldg(lg + t.o10), l11 = __ldg(lg + t.o11);
            const float cr = hblend(a.x, bq.x, cq.x, d.x, t);
            const float cg = hblend(a.y, bq.y, cq.y, d.y, t);
            const float cb = hblend(a.z, bq.z, cq.z, d.z, t);
            const float l = hblend(l00, l01, l10, l11, t);
            const float pi = fast_exp2(fmaf(l, kLog2e, -Ml2)) * invS;
            const float Gn = g0 * cr + g1 * cg + g2 * cb;
            float dcr, dcg, dcb;
            float s00 = 0, s01 = 0, s10 = 0, s11 = 0;
            if (MIX) {
                const float* sp = sgp;
                s00 = __ldg(sp + t.o00), s01 = __ldg(sp + t.o01), s10 = __ldg(sp + t.o10), s11 = __ldg(sp + t.o11);
                const float sraw = hblend(s00, s01, s10, s11, t);
                const float sg = fminf(fmaxf(sraw, 0.01f), 1.0f);
                const float inv = 1.0f / sg;
                const float w = pi * inv * Zinv;  // compositing weight
                const float err = (fabsf(cr - tr) + fabsf(cg - tg) + fabsf(cb - tb)) * (1.0f / 3.0f);
                const float lap = 0.5f * fast_exp(-err * inv) * inv;
                const float P = (Gn - Gbar) * inv * Zinv + gD * lap;
                dl = pi * (P - gDD);
                const float dsgt = -(Gn - Gbar) * w * inv + gD * pi * lap * (err - sg) * inv * inv;
                dsg = (sraw >= 0.01f && sraw <= 1.0f) ? dsgt : 0.0f;  // clamp backward
                const float ce = -gD * pi * lap * inv * (1.0f / 3.0f);
                dcr = w * g0 + ce * ((cr > tr) ? 1.0f : ((cr < tr) ? -1.0f : 0.0f));
                dcg = w * g1 + ce * ((cg > tg) ? 1.0f : ((cg < tg) ? -1.0f : 0.0f));
                dcb = w * g2 + ce * ((cb > tb) ? 1.0f : ((cb < tb) ? -1.0f : 0.0f));
            } else {
                dl = pi * (Gn - Gbar);
                dcr = pi * g0, dcg = pi * g1, dcb = pi * g2;
            }
            if (want_h) {
                // The sample is linear in the tap values, so the coordinate gradient (blend_grad() of pd_device.cuh, i.e. ATen's
                // gix / giy) is taken once on the combined tap T = dcr*r + dcg*g + dcb*b + dl*logit (+ dsg*sigma); taps outside
                // the image read as zero (their loads were clamped onto a neighbour)
                float tnw = fmaf(dcr, a.x, fmaf(dcg, a.y, fmaf(dcb, a.z, dl * l00)));
                float tne = fmaf(dcr, bq.x, fmaf(dcg, bq.y, fmaf(dcb, bq.z, dl * l01)));
                float tsw = fmaf(dcr, cq.x, fmaf(dcg, cq.y, fmaf(dcb, cq.z, dl * l10)));
                float tse = fmaf(dcr, d.x, fmaf(dcg, d.y, fmaf(dcb, d.z, dl * l11)));
                if (MIX) tnw = fmaf(dsg, s00, tnw), tne = fmaf(dsg, s01, tne), tsw = fmaf(dsg, s10, tsw), tse = fmaf(dsg, s11, tse);
                if constexpr (!ALL_IN) {
                    tnw = (t.ix0 && t.iy0) ? tnw : 0.0f, tne = (t.ix1 && t.iy0) ? tne : 0.0f;
                    tsw = (t.ix0 && t.iy1) ? tsw : 0.0f, tse = (t.ix1 && t.iy1) ? tse : 0.0f;
                }
                const float gx = (tne - tnw) * t.ry0 + (tse - tsw) * t.ry1;
                const float gy = (tsw - tnw) * t.rx0 + (tse - tne) * t.rx1;
                // u = qx / zc, v = qy / zc, zc = max(qz, 1e-7)   (layers.py:227-228)
                gqx = gx * c.zinv, gqy = gy * c.zinv;
                gqz = -(gx * c.u + gy * c.v) * c.zinv * c.dz;
            }
            if constexpr (ALL_IN) {
                // four in-image taps: reductions without the per-tap zero tests (a zero weight adds an exact zero)
                if (glg && dl != 0.0f) {
                    atomicAdd(glg + t.o00, dl * t.w00), atomicAdd(glg + t.o01, dl * t.w01);
                    atomicAdd(glg + t.o10, dl * t.w10), atomicAdd(glg + t.o11, dl * t.w11);
                }
                if (MIX && gsg && dsg != 0.0f) {
                    atomicAdd(gsg + t.o00, dsg * t.w00), atomicAdd(gsg + t.o01, dsg * t.w01);
                    atomicAdd(gsg + t.o10, dsg * t.w10), atomicAdd(gsg + t.o11, dsg * t.w11);
                }
            } else {
                if (glg && dl != 0.0f) hscatter(glg, t, dl);
                if (MIX && gsg && dsg != 0.0f) hscatter(gsg, t, dsg);
            }
        };
        if (c.m != 0.0f) {  // every gradient of a masked plane carries the factor m = 0 (:580)
            float su = rt<FASTRT>(c.u, p.wm1, rcp_w), sv = rt<FASTRT>(c.v, p.hm1, rcp_h);
            su = fminf(fmaxf(su, -2.0f), Wp1);
            sv = fminf(fmaxf(sv, -2.0f), Hp1);
            const int x0 = __float2int_rd(su), y0 = __float2int_rd(sv);
            if ((unsigned)x0 < Wm1 && (unsigned)y0 < Hm1) {
                const float fx0 = (float)x0, fy0 = (float)y0;
                HTaps t;
                t.rx1 = su - fx0, t.rx0 = (fx0 + 1.0f) - su, t.ry1 = sv - fy0, t.ry0 = (fy0 + 1.0f) - sv;
                t.ix0 = t.ix1 = t.iy0 = t.iy1 = true;
                t.o00 = y0 * W + x0, t.o01 = t.o00 + 1, t.o10 = t.o00 + W, t.o11 = t.o10 + 1;
                t.w00 = t.rx0 * t.ry0, t.w01 = t.rx1 * t.ry0, t.w10 = t.rx0 * t.ry1, t.w11 = t.rx1 * t.ry1;
                sample(t, std::true_type{});
            } else {
                sample(make_htaps(su, sv, W, H, Wp1, Hp1), std::false_type{});
            }
        }
        if (want_h) {
            // dL/dH[i][j] = sum_pixels gq_i * (x, y, 1)_j; y is the same for the whole warp
            const float v8[8] = {gqx, gqy, gqz, gqx * fx, gqy * fx, gqz * fx, 0.0f, 0.0f};
            const float tot = warp_reduce_scatter8(v8);
            const int k = lane >> 2;
            if ((lane & 3) == 0 && k < 6 && tot != 0.0f) {
                float* dst = gacc + n * 9;  // the six writing lanes of a warp touch nine distinct elements
                if (k < 3) {
                    dst[3 * k + 2] += tot;
                    dst[3 * k + 1] += tot * fy;
                } else {
                    dst[3 * (k - 3)] += tot;
                }
            }
        }
    }
    if (want_h) {
        __syncthreads();
        float* dst = p.gin.g_hmat + (int64_t)b * N * 9;
        for (int i = threadIdx.x; i < N * 9; i += HT) {
            float v = 0.0f;
#pragma unroll
            for (int w = 0; w < HT / 32; ++w) v += gacc0[w * N * 9 + i];
            if (v != 0.0f) atomicAdd(dst + i, v);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
inline bool homo_path_supported(const WarpParams& p) {
    // a CTA of HT consecutive pixels must stay inside one image and a warp inside one row
    return p.d.warp_type == PD_WARP_HOMOGRAPHY && p.hw % HT == 0 && p.d.W % 32 == 0 && p.hw * (int64_t)p.d.N < (1ll << 31);
}

inline size_t homo_workspace_bytes(const pd_warp_desc* d) { return (size_t)d->B * d->H * d->W * sizeof(float4); }

inline void homo_pack(const WarpParams& p, float4* rgbx, cudaStream_t st) {
    const int64_t total = (int64_t)p.d.B * p.hw;
    pack_rgbx_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p.in.src, rgbx, p.hw, total);
}

inline void launch_homo_fwd(const WarpParams& p, const float4* rgbx, cudaStream_t st) {
    const unsigned grid = (unsigned)((int64_t)p.d.B * p.hw / HT);
    const size_t smem = (size_t)p.d.N * 12 * sizeof(float);
    const float rw = rows_rcp(p.d.W), rh = rows_rcp(p.d.H);
    const bool fast = rw != 0.0f && rh != 0.0f;
    if (p.d.mixture) {
        if (fast) homo_fwd_kernel<true, true><<<grid, HT, smem, st>>>(p, rgbx, rw, rh);
        else homo_fwd_kernel<true, false><<<grid, HT, smem, st>>>(p, rgbx, rw, rh);
    } else {
        if (fast) homo_fwd_kernel<false, true><<<grid, HT, smem, st>>>(p, rgbx, rw, rh);
        else homo_fwd_kernel<false, false><<<grid, HT, smem, st>>>(p, rgbx, rw, rh);
    }
}

inline void launch_homo_bwd(const WarpParams& p, const float4* rgbx, cudaStream_t st) {
    const unsigned grid = (unsigned)((int64_t)p.d.B * p.hw / HT);
    const size_t smem = (size_t)p.d.N * (12 + (HT / 32) * 9) * sizeof(float);
    const float rw = rows_rcp(p.d.W), rh = rows_rcp(p.d.H);
    const bool fused = p.gout.g_ph_sum != nullptr, fast = rw != 0.0f && rh != 0.0f;
    auto launch = [&](auto kern) {
        smem_optin((const void*)kern, smem);  // N = 256 planes need 86 KB; cached per (kernel, size), no attribute call in steady state
        kern<<<grid, HT, smem, st>>>(p, rgbx, rw, rh);
    };
    if (p.d.mixture) {
        if (fused) fast ? launch(homo_bwd_kernel<true, true, true>) : launch(homo_bwd_kernel<true, true, false>);
        else fast ? launch(homo_bwd_kernel<true, false, true>) : launch(homo_bwd_kernel<true, false, false>);
    } else {
        if (fused) fast ? launch(homo_bwd_kernel<false, true, true>) : launch(homo_bwd_kernel<false, true, false>);
        else fast ? launch(homo_bwd_kernel<false, false, true>) : launch(homo_bwd_kernel<false, false, false>);
    }
}

}  // namespace hm
}  // namespace pd
