// Device-side building blocks shared by the warp/composite kernels.
//
// Everything here restates arithmetic the reference performs through PyTorch ops; the citations say
// which.  fp32 throughout; rounding-sensitive steps use explicit round-to-nearest intrinsics so that
// nvcc cannot contract them into FMAs the reference does not perform.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/planedepth_b200.h"
#include "pd_common.h"

namespace pd {

constexpr float kLog2e = 1.4426950408889634f;

// exp() used for softmax / Laplacian terms.  ex2.approx after an exact-rounded scale: relative error
// ~2 ulp + 6e-8*|x|, far inside the 1e-4 parity budget (|x| <= ~100 on this path).
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_exp(float x) { return fast_exp2(x * kLog2e); }
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// The reference normalises pixel coordinates to [-1,1] (trainer.py:549-551, layers.py:179-181,
// layers.py:231-233: p/(size-1), (p-0.5)*2) and ATen un-normalises them again
// (GridSampler.h grid_sampler_unnormalize, align_corners=True: ((g+1)/2)*(size-1)).  The fp32 round
// trip moves integer coordinates by up to 6e-5 px, which changes taps and weights; reproduce it.
__device__ __forceinline__ float roundtrip(float p, float size_m1) {
    float q = __fdiv_rn(p, size_m1);
    float g = __fmul_rn(__fsub_rn(q, 0.5f), 2.0f);
    return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), size_m1);
}

// Same value as roundtrip() for 2 <= size-1 < 2^23 whose significand is not all ones, without the
// IEEE division: q0 = p*r, residual by FMA, one Newton correction (Markstein) gives the correctly
// rounded quotient when r = RN(1/size_m1).  Host code decides which variant is legal for a size.
__device__ __forceinline__ float roundtrip_fast(float p, float size_m1, float rcp) {
    float q0 = __fmul_rn(p, rcp);
    float rem = __fmaf_rn(-q0, size_m1, p);
    float q = __fmaf_rn(rem, rcp, q0);
    float g = __fmaf_rn(q, 2.0f, -1.0f);  // (q-0.5)*2: the doubling is exact, so one rounding either way
    // ((g+1)*0.5)*(size-1) == (g+1)*((size-1)/2): the halving is exact and (size-1)/2 is representable
    return __fmul_rn(__fadd_rn(g, 1.0f), 0.5f * size_m1);
}

// reciprocal handed to roundtrip_fast(); 0 = "use the IEEE-division form" (size-1 = 2^k - 1: the Newton step is not
// guaranteed to round correctly)
inline float rows_rcp(int size) { return ((size & (size - 1)) == 0) ? 0.0f : 1.0f / (float)(size - 1); }

struct Taps {
    // integer corner (x0,y0), fractional weights exactly as ATen forms them
    int x0, y0;
    float wx0, wx1, wy0, wy1;  // wx0 = (x0+1) - x, wx1 = x - x0
    bool in_x0, in_x1, in_y0, in_y1;
};

// ATen grid_sampler_2d, bilinear, padding_mode=zeros: floor, the four weights, per-tap bounds.
// Coordinates are clamped to [-2, size+1] first: anything outside has all taps out of range, and the
// clamp keeps float->int conversion defined for inf/NaN (NaN -> -2 -> all taps out of range).
__device__ __forceinline__ Taps make_taps(float x, float y, int W, int H) {
    Taps t;
    x = fminf(fmaxf(x, -2.0f), (float)(W + 1));
    y = fminf(fmaxf(y, -2.0f), (float)(H + 1));
    // floor as ONE float -> int conversion per axis; the integer goes back to float on the ALU pipe (exact: |x0| <= size + 1)
    t.x0 = __float2int_rd(x);
    t.y0 = __float2int_rd(y);
    const float fx0 = (float)t.x0, fy0 = (float)t.y0;
    t.wx1 = x - fx0;
    t.wx0 = (fx0 + 1.0f) - x;
    t.wy1 = y - fy0;
    t.wy0 = (fy0 + 1.0f) - y;
    t.in_x0 = (unsigned)t.x0 < (unsigned)W;
    t.in_x1 = (unsigned)(t.x0 + 1) < (unsigned)W;
    t.in_y0 = (unsigned)t.y0 < (unsigned)H;
    t.in_y1 = (unsigned)(t.y0 + 1) < (unsigned)H;
    return t;
}

struct TapVals {
    float nw, ne, sw, se;
};

__device__ __forceinline__ TapVals load_taps(const float* __restrict__ plane, const Taps& t, int W) {
    TapVals v;
    const float* r0 = plane + (int64_t)t.y0 * W + t.x0;
    const float* r1 = r0 + W;
    v.nw = (t.in_y0 && t.in_x0) ? __ldg(r0) : 0.0f;
    v.ne = (t.in_y0 && t.in_x1) ? __ldg(r0 + 1) : 0.0f;
    v.sw = (t.in_y1 && t.in_x0) ? __ldg(r1) : 0.0f;
    v.se = (t.in_y1 && t.in_x1) ? __ldg(r1 + 1) : 0.0f;
    return v;
}

// ATen accumulates nw, ne, sw, se in that order.
__device__ __forceinline__ float blend(const TapVals& v, const Taps& t) {
    float acc = v.nw * (t.wx0 * t.wy0);
    acc = fmaf(v.ne, t.wx1 * t.wy0, acc);
    acc = fmaf(v.sw, t.wx0 * t.wy1, acc);
    acc = fmaf(v.se, t.wx1 * t.wy1, acc);
    return acc;
}

// d(sample)/dx and d(sample)/dy in pixel units (ATen grid_sampler_2d_backward's gix/giy before its
// (size-1)/2 factor, which the reference's 2/(size-1) normalisation cancels).
__device__ __forceinline__ void blend_grad(const TapVals& v, const Taps& t, float& dx, float& dy) {
    dx = (v.ne - v.nw) * t.wy0 + (v.se - v.sw) * t.wy1;
    dy = (v.sw - v.nw) * t.wx0 + (v.se - v.ne) * t.wx1;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ int64_t soff(const pd_strides4& s, int b, int n, int y, int x) {
    return (int64_t)b * s.b + (int64_t)n * s.n + (int64_t)y * s.y + (int64_t)x * s.x;
}

__device__ __forceinline__ float load_mask(const void* mask, int dtype, int64_t off) {
    if (dtype == PD_MASK_F32) return __ldg(reinterpret_cast<const float*>(mask) + off);
    if (dtype == PD_MASK_U8) return (float)__ldg(reinterpret_cast<const unsigned char*>(mask) + off);
    return 1.0f;
}

}  // namespace pd
