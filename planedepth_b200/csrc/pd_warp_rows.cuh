// Row-tiled fast path for stereo disparity warps (trainer.py:540-554) whose disparity does not vary
// along x (vertical planes: one scalar per (image, plane); xz ground planes: one scalar per row —
// depth_decoder.py:153-156, 163-181).  For such a plane the warp of a row is a shift + lerp of the
// source row, so
//   * one CTA owns one target row (b, y) at a time (persistent CTAs stride over rows),
//   * the (row-blended) source row sits in shared memory once and serves all N planes,
//   * every thread owns 4 consecutive pixels: taps come in as aligned 128-bit windows (shared memory
//     for rgb, global/L2 for the plane's logit row, prefetched one plane ahead) instead of scalar
//     gathers,
//   * the softmax over planes runs against a lazily updated reference logit (one exp per plane),
//   * the backward pass is a GATHER: per-target contributions are exchanged through shared memory
//     and every gradient row is written exactly once with coalesced 128-bit stores — no atomics and
//     no zero-fill pass (ATen's grid_sampler_2d_backward scatters with atomicAdd).
// Arithmetic follows the general kernels bit-for-bit where it matters (the fp32 normalise /
// un-normalise round trip of the coordinates); see DESIGN.md for the two documented deviations
// (row pre-blend order, dropped <=6.2e-5-weighted cross-row taps in the backward).
#pragma once
#include "pd_warp_general.cuh"

namespace pd {

constexpr int RP = 4;        // pixels per thread
constexpr int ROW_PAD = 8;   // zero floats on both sides of every shared-memory row

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
// 128-bit shared load kept opaque so that the compiler does not split it into conflicting narrow loads
__device__ __forceinline__ float4 lds4(const float* p) {
    float4 v;
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}

struct RowGeom {  // vertical taps of target row y (same for every plane and pixel of a disparity warp)
    int r0, r1;      // source rows, -1 = out of range / unused
    float w0, w1;    // their weights (ATen's (y0+1)-y and y-y0)
};

__device__ __forceinline__ RowGeom row_geom(int y, int H, float hm1) {
    float yh = roundtrip((float)y, hm1);
    yh = fminf(fmaxf(yh, -2.0f), (float)(H + 1));
    float f0 = floorf(yh);
    int y0 = (int)f0;
    RowGeom g;
    g.w1 = yh - f0;
    g.w0 = (f0 + 1.0f) - yh;
    g.r0 = ((unsigned)y0 < (unsigned)H) ? y0 : -1;
    g.r1 = ((unsigned)(y0 + 1) < (unsigned)H && g.w1 != 0.0f) ? y0 + 1 : -1;
    return g;
}

// per-(row, plane) scalars staged in shared memory
struct PlaneRow {
    float sd;    // sign * disparity
    float k0f;   // floor(sd) as float
    int k0;      // floor(sd)
    int regular; // 1: ix(x) == x + k0 for every x of the row (frac(sd) is clear of the round-trip wobble)
    float m;     // row mask value (MASK_ROW) or 1
};

enum { MASK_ROW = 0, MASK_DENSE_F32 = 1 };

struct Win8 {
    float4 lo, hi;
};

__device__ __forceinline__ void unpack(const Win8& w, float v[8]) {
    v[0] = w.lo.x, v[1] = w.lo.y, v[2] = w.lo.z, v[3] = w.lo.w, v[4] = w.hi.x, v[5] = w.hi.y, v[6] = w.hi.z, v[7] = w.hi.w;
}

// 8 consecutive floats [a, a+8) of a padded shared row (a is a multiple of 4, clamped into the zero pads)
__device__ __forceinline__ void win_smem(const float* row, int a, int W, float v[8]) {
    int ac = min(max(a, -ROW_PAD), W);
    Win8 w;
    w.lo = lds4(row + ac);
    w.hi = lds4(row + ac + 4);
    unpack(w, v);
}

// same window from a global row with zero padding outside [0, W)
__device__ __forceinline__ Win8 win_gmem(const float* row, int a, int W) {
    Win8 w;
    w.lo = ((unsigned)a < (unsigned)W) ? ldg4(row + a) : zero4();
    w.hi = ((unsigned)(a + 4) < (unsigned)W) ? ldg4(row + a + 4) : zero4();
    return w;
}

// raw windows of the (at most two) source rows of a plane; blended when consumed so that the loads
// can be issued a whole plane ahead of their first use
struct PlaneWin {
    Win8 a, b;
};

__device__ __forceinline__ void load_plane_win(const float* plane, const RowGeom& g, int a, int W, PlaneWin& w) {
    if (g.r0 >= 0) w.a = win_gmem(plane + (int64_t)g.r0 * W, a, W);
    else w.a.lo = w.a.hi = zero4();
    if (g.r1 >= 0) w.b = win_gmem(plane + (int64_t)g.r1 * W, a, W);
}

__device__ __forceinline__ void blend_plane_win(const PlaneWin& w, const RowGeom& g, float v[8]) {
    unpack(w.a, v);
    if (g.w0 != 1.0f) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= g.w0;
    }
    if (g.r1 >= 0) {
        float t[8];
        unpack(w.b, t);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaf(t[i], g.w1, v[i]);
    }
}

struct RowCtx {
    int b, y, x0, W, H, N;
    bool active;
    float xf[RP];
    RowGeom g;
};

// window base (multiple of 4) of a regular plane for this thread
__device__ __forceinline__ int win_base(const RowCtx& c, const PlaneRow& pr) { return c.x0 + (pr.k0 & ~3); }

// stage the vertically blended source row (3 channels) and the per-plane scalars of row (b, y)
template <int MASKMODE>
__device__ __forceinline__ void stage_row(const WarpParams& p, const RowCtx& c, float* srow, PlaneRow* pr) {
    const int W = c.W, pitch = W + 2 * ROW_PAD;
    for (int i = threadIdx.x; i < 3 * (W / 4); i += blockDim.x) {
        int ch = i / (W / 4), xq = (i - ch * (W / 4)) * 4;
        const float* sp = p.in.src + ((int64_t)c.b * 3 + ch) * p.hw;
        float4 v = zero4();
        if (c.g.r0 >= 0) {
            float4 a = ldg4(sp + (int64_t)c.g.r0 * W + xq);
            v = make_float4(a.x * c.g.w0, a.y * c.g.w0, a.z * c.g.w0, a.w * c.g.w0);
        }
        if (c.g.r1 >= 0) {
            float4 a = ldg4(sp + (int64_t)c.g.r1 * W + xq);
            v = make_float4(fmaf(a.x, c.g.w1, v.x), fmaf(a.y, c.g.w1, v.y), fmaf(a.z, c.g.w1, v.z), fmaf(a.w, c.g.w1, v.w));
        }
        *reinterpret_cast<float4*>(srow + ch * pitch + ROW_PAD + xq) = v;
    }
    for (int n = threadIdx.x; n < c.N; n += blockDim.x) {
        PlaneRow r;
        float d = __ldg(p.in.disp + soff(p.d.disp_stride, c.b, n, c.y, 0));
        r.sd = __fmul_rn(p.d.disp_sign, d);
        float kf = floorf(r.sd);
        // tap index x + k0 holds for every x when frac(sd) stays clear of the accumulated fp32 wobble
        // of (x + sd) and of the normalise / un-normalise round trip (<= 2 ulp of the coordinate)
        float margin = 1e-3f * fmaxf(1.0f, ((float)W + fabsf(r.sd)) * (1.0f / 2048.0f));
        float fr = r.sd - kf;
        bool sane = fabsf(r.sd) < 1.0e6f;
        r.k0f = kf;
        r.k0 = sane ? (int)kf : 0;
        r.regular = (sane && fr > margin && fr < 1.0f - margin) ? 1 : 0;
        r.m = (MASKMODE == MASK_ROW) ? load_mask(p.in.mask, p.d.mask_dtype, soff(p.d.mask_stride, c.b, n, c.y, 0)) : 1.0f;
        pr[n] = r;
    }
}

// Samples of one plane for the 4 pixels of a thread: cr/cg/cb/l (/s) BEFORE the mask multiply, plus
// the horizontal weights and (GRAD) the finite differences tap[ix+1]-tap[ix] for the coordinate gradient.
template <bool MIX>
struct PlaneSamples {
    float w0[RP], w1[RP];
    float cr[RP], cg[RP], cb[RP], l[RP], s[RP];
    float dr[RP], dg[RP], db[RP], dl[RP], ds[RP];
};

// horizontal weights on a regular plane: the tap index is x + k0 exactly, the weight carries the
// fp32 round-trip wobble (division-free form, bit-identical to roundtrip(); see pd_debug_roundtrip)
__device__ __forceinline__ void hweights(const float xf[RP], const PlaneRow& pr, float wm1, float rcp, float w0[RP], float w1[RP]) {
#pragma unroll
    for (int i = 0; i < RP; ++i) {
        float uh = roundtrip_fast(__fadd_rn(xf[i], pr.sd), wm1, rcp);
        w1[i] = uh - (xf[i] + pr.k0f);
        w0[i] = 1.0f - w1[i];
    }
}

template <bool GRAD, int R>
__device__ __forceinline__ void lerp_win(const float v[8], const float w0[RP], const float w1[RP], float out[RP], float diff[RP]) {
#pragma unroll
    for (int i = 0; i < RP; ++i) {
        out[i] = fmaf(w1[i], v[R + i + 1], w0[i] * v[R + i]);
        if (GRAD) diff[i] = v[R + i + 1] - v[R + i];
    }
}

template <bool MIX, bool GRAD, int R>
__device__ __forceinline__ void sample_regular(const WarpParams& p, const RowCtx& c, const float* srow, const PlaneRow& pr, float rcp,
                                               const PlaneWin& lw, const PlaneWin& sw, PlaneSamples<MIX>& o) {
    const int W = c.W, pitch = W + 2 * ROW_PAD;
    hweights(c.xf, pr, p.wm1, rcp, o.w0, o.w1);
    const int a = win_base(c, pr);
    float v[8];
    win_smem(srow + ROW_PAD, a, W, v);
    lerp_win<GRAD, R>(v, o.w0, o.w1, o.cr, o.dr);
    win_smem(srow + pitch + ROW_PAD, a, W, v);
    lerp_win<GRAD, R>(v, o.w0, o.w1, o.cg, o.dg);
    win_smem(srow + 2 * pitch + ROW_PAD, a, W, v);
    lerp_win<GRAD, R>(v, o.w0, o.w1, o.cb, o.db);
    blend_plane_win(lw, c.g, v);
    lerp_win<GRAD, R>(v, o.w0, o.w1, o.l, o.dl);
    if (MIX) {
        blend_plane_win(sw, c.g, v);
        lerp_win<GRAD, R>(v, o.w0, o.w1, o.s, o.ds);
    }
}

__device__ __forceinline__ float plane_tap(const float* plane, const RowGeom& g, int x, int W) {
    if ((unsigned)x >= (unsigned)W) return 0.0f;
    float v = 0.0f;
    if (g.r0 >= 0) v = __ldg(plane + (int64_t)g.r0 * W + x) * g.w0;
    if (g.r1 >= 0) v = fmaf(__ldg(plane + (int64_t)g.r1 * W + x), g.w1, v);
    return v;
}

// irregular plane (frac of the disparity within the round-trip wobble of an integer): scalar floor path
template <bool MIX, bool GRAD>
__device__ __forceinline__ void sample_irregular(const WarpParams& p, const RowCtx& c, const float* srow, const PlaneRow& pr, int n,
                                              PlaneSamples<MIX>& o) {
    const int W = c.W, pitch = W + 2 * ROW_PAD;
    const int64_t pl = ((int64_t)c.b * c.N + n) * p.hw;
#pragma unroll
    for (int i = 0; i < RP; ++i) {
        float uh = roundtrip(__fadd_rn(c.xf[i], pr.sd), p.wm1);
        uh = fminf(fmaxf(uh, -2.0f), (float)(W + 1));
        int x0 = __float2int_rd(uh);
        float f0 = (float)x0;
        o.w1[i] = uh - f0;
        o.w0[i] = (f0 + 1.0f) - uh;
        int xs = min(max(x0, -ROW_PAD), W + ROW_PAD - 2);  // pads are zero
        const float* s0 = srow + ROW_PAD + xs;
        float a0 = s0[0], a1 = s0[1], b0 = s0[pitch], b1 = s0[pitch + 1], c0 = s0[2 * pitch], c1 = s0[2 * pitch + 1];
        o.cr[i] = fmaf(o.w1[i], a1, o.w0[i] * a0);
        o.cg[i] = fmaf(o.w1[i], b1, o.w0[i] * b0);
        o.cb[i] = fmaf(o.w1[i], c1, o.w0[i] * c0);
        float l0 = plane_tap(p.in.logits + pl, c.g, x0, W), l1 = plane_tap(p.in.logits + pl, c.g, x0 + 1, W);
        o.l[i] = fmaf(o.w1[i], l1, o.w0[i] * l0);
        if (GRAD) { o.dr[i] = a1 - a0, o.dg[i] = b1 - b0, o.db[i] = c1 - c0, o.dl[i] = l1 - l0; }
        if (MIX) {
            float s0v = plane_tap(p.in.sigma + pl, c.g, x0, W), s1v = plane_tap(p.in.sigma + pl, c.g, x0 + 1, W);
            o.s[i] = fmaf(o.w1[i], s1v, o.w0[i] * s0v);
            if (GRAD) o.ds[i] = s1v - s0v;
        }
    }
}

template <bool MIX, bool GRAD>
__device__ __forceinline__ void sample_any(const WarpParams& p, const RowCtx& c, const float* srow, const PlaneRow& pr, int n, float rcp,
                                           const PlaneWin& lw, const PlaneWin& sw, PlaneSamples<MIX>& o) {
    if (pr.regular) {
        switch (pr.k0 & 3) {
            case 0: sample_regular<MIX, GRAD, 0>(p, c, srow, pr, rcp, lw, sw, o); break;
            case 1: sample_regular<MIX, GRAD, 1>(p, c, srow, pr, rcp, lw, sw, o); break;
            case 2: sample_regular<MIX, GRAD, 2>(p, c, srow, pr, rcp, lw, sw, o); break;
            default: sample_regular<MIX, GRAD, 3>(p, c, srow, pr, rcp, lw, sw, o); break;
        }
    } else {
        sample_irregular<MIX, GRAD>(p, c, srow, pr, n, o);
    }
}

// issue the global loads of plane n's logit (and sigma) windows; consumed one iteration later
template <bool MIX>
__device__ __forceinline__ void prefetch_plane(const WarpParams& p, const RowCtx& c, const PlaneRow& pr, int n, PlaneWin& lw, PlaneWin& sw) {
    if (!pr.regular) return;
    const int64_t pl = ((int64_t)c.b * c.N + n) * p.hw;
    const int a = win_base(c, pr);
    load_plane_win(p.in.logits + pl, c.g, a, c.W, lw);
    if (MIX) load_plane_win(p.in.sigma + pl, c.g, a, c.W, sw);
}

template <int MASKMODE>
__device__ __forceinline__ bool plane_mask(const WarpParams& p, const RowCtx& c, int n, float mrow, float m[RP]) {
    // returns false when the mask is identically 1 for this thread's pixels (no multiplies needed)
    if constexpr (MASKMODE == MASK_DENSE_F32) {
        float4 v = ldg4(reinterpret_cast<const float*>(p.in.mask) + soff(p.d.mask_stride, c.b, n, c.y, c.x0));
        m[0] = v.x, m[1] = v.y, m[2] = v.z, m[3] = v.w;
        return true;
    } else {
#pragma unroll
        for (int i = 0; i < RP; ++i) m[i] = mrow;
        return mrow != 1.0f;
    }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <bool MIX, int MASKMODE, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) warp_composite_fwd_rows(const WarpParams p, const float rcp_wm1) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int W = p.d.W, H = p.d.H, N = p.d.N, pitch = W + 2 * ROW_PAD;
    float* srow = reinterpret_cast<float*>(smem_raw);
    PlaneRow* prs = reinterpret_cast<PlaneRow*>(srow + 3 * pitch);
    for (int i = threadIdx.x; i < 3 * pitch; i += blockDim.x) srow[i] = 0.0f;  // pads stay zero for the whole kernel

    RowCtx c;
    c.W = W, c.H = H, c.N = N;
    c.x0 = threadIdx.x * RP;
    c.active = c.x0 < W;
#pragma unroll
    for (int i = 0; i < RP; ++i) c.xf[i] = (float)(c.x0 + i);
    const int rows_total = p.d.B * H;
    const float l2e = kLog2e;

    for (int row = blockIdx.x; row < rows_total; row += gridDim.x) {
        c.b = row / H;
        c.y = row - c.b * H;
        c.g = row_geom(c.y, H, p.hm1);
        __syncthreads();  // previous row fully consumed (and the zero fill above)
        stage_row<MASKMODE>(p, c, srow, prs);
        __syncthreads();
        if (!c.active) continue;
        const int64_t rem = (int64_t)c.y * W + c.x0;
        float tr[RP], tg[RP], tb[RP], ea[RP];
        if (MIX) {
            const float* tp = p.in.tgt + (int64_t)c.b * p.chw3 + rem;
            float4 a = ldg4(tp), bq = ldg4(tp + p.hw), cq = ldg4(tp + 2 * p.hw);
            tr[0] = a.x, tr[1] = a.y, tr[2] = a.z, tr[3] = a.w;
            tg[0] = bq.x, tg[1] = bq.y, tg[2] = bq.z, tg[3] = bq.w;
            tb[0] = cq.x, tb[1] = cq.y, tb[2] = cq.z, tb[3] = cq.w;
            if (p.d.automask) {
                const float* sp = p.in.src + (int64_t)c.b * p.chw3 + rem;
                float4 sa = ldg4(sp), sb = ldg4(sp + p.hw), sc = ldg4(sp + 2 * p.hw);
                ea[0] = (fabsf(sa.x - tr[0]) + fabsf(sb.x - tg[0]) + fabsf(sc.x - tb[0])) * (1.0f / 3.0f);
                ea[1] = (fabsf(sa.y - tr[1]) + fabsf(sb.y - tg[1]) + fabsf(sc.y - tb[1])) * (1.0f / 3.0f);
                ea[2] = (fabsf(sa.z - tr[2]) + fabsf(sb.z - tg[2]) + fabsf(sc.z - tb[2])) * (1.0f / 3.0f);
                ea[3] = (fabsf(sa.w - tr[3]) + fabsf(sb.w - tg[3]) + fabsf(sc.w - tb[3])) * (1.0f / 3.0f);
            } else {
#pragma unroll
                for (int i = 0; i < RP; ++i) ea[i] = 0.0f;
            }
        }
        // softmax over planes against a lazily updated reference logit: exp(l - ref) cannot overflow
        // while (l - ref)*log2(e) <= 64, so the reference only moves when a logit exceeds it by that much
        // (always on the first plane: ref starts at -inf).  Ml2 = ref * log2(e).
        float Ml2[RP], S[RP], A[RP], R0[RP], R1[RP], R2[RP], Q[RP], Qa[RP];
#pragma unroll
        for (int i = 0; i < RP; ++i) { Ml2[i] = -INFINITY, S[i] = A[i] = R0[i] = R1[i] = R2[i] = Q[i] = Qa[i] = 0.0f; }

        PlaneWin lw_cur, sw_cur, lw_nxt, sw_nxt;
        prefetch_plane<MIX>(p, c, prs[0], 0, lw_cur, sw_cur);
        for (int n = 0; n < N; ++n) {
            const PlaneRow pr = prs[n];
            if (n + 1 < N) prefetch_plane<MIX>(p, c, prs[n + 1], n + 1, lw_nxt, sw_nxt);
            PlaneSamples<MIX> sm;
            sample_any<MIX, false>(p, c, srow, pr, n, rcp_wm1, lw_cur, sw_cur, sm);
            float m[RP];
            if (plane_mask<MASKMODE>(p, c, n, pr.m, m)) {
#pragma unroll
                for (int i = 0; i < RP; ++i) {
                    sm.l[i] *= m[i], sm.cr[i] *= m[i], sm.cg[i] *= m[i], sm.cb[i] *= m[i];
                    if (MIX) sm.s[i] *= m[i];
                }
            }
            float t[RP];
#pragma unroll
            for (int i = 0; i < RP; ++i) t[i] = fmaf(sm.l[i], l2e, -Ml2[i]);
            if (fmaxf(fmaxf(t[0], t[1]), fmaxf(t[2], t[3])) > 64.0f) {
#pragma unroll
                for (int i = 0; i < RP; ++i) {
                    if (t[i] > 64.0f) {
                        float nl2 = sm.l[i] * l2e;
                        float sc = fast_exp2(Ml2[i] - nl2);
                        S[i] *= sc, R0[i] *= sc, R1[i] *= sc, R2[i] *= sc;
                        if (MIX) { A[i] *= sc, Q[i] *= sc, Qa[i] *= sc; }
                        Ml2[i] = nl2;
                        t[i] = 0.0f;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < RP; ++i) {
                float e = fast_exp2(t[i]);
                S[i] += e;
                if (MIX) {
                    float sg = fminf(fmaxf(sm.s[i], 0.01f), 1.0f);
                    float inv = fast_rcp(sg);
                    float es = e * inv;
                    A[i] += es;
                    R0[i] = fmaf(es, sm.cr[i], R0[i]), R1[i] = fmaf(es, sm.cg[i], R1[i]), R2[i] = fmaf(es, sm.cb[i], R2[i]);
                    float err = (fabsf(sm.cr[i] - tr[i]) + fabsf(sm.cg[i] - tg[i]) + fabsf(sm.cb[i] - tb[i])) * (1.0f / 3.0f);
                    float il2 = inv * l2e;
                    Q[i] = fmaf(e, 0.5f * fast_exp2(-err * il2) * inv, Q[i]);
                    Qa[i] = fmaf(e, 0.5f * fast_exp2(-ea[i] * il2) * inv, Qa[i]);
                } else {
                    R0[i] = fmaf(e, sm.cr[i], R0[i]), R1[i] = fmaf(e, sm.cg[i], R1[i]), R2[i] = fmaf(e, sm.cb[i], R2[i]);
                }
            }
            lw_cur = lw_nxt;
            if (MIX) sw_cur = sw_nxt;
        }
        float o0[RP], o1[RP], o2[RP], dq[RP], nl[RP], na[RP];
#pragma unroll
        for (int i = 0; i < RP; ++i) {
            float invS = 1.0f / S[i];
            float invD = MIX ? 1.0f / A[i] : invS;
            o0[i] = R0[i] * invD, o1[i] = R1[i] * invD, o2[i] = R2[i] * invD;
            if (MIX) {
                dq[i] = Q[i] * invS + 1e-7f;
                nl[i] = -logf(dq[i]);
                na[i] = -logf(Qa[i] * invS + 1e-7f);
            }
        }
        float* rr = p.out.rgb_rec + (int64_t)c.b * p.chw3 + rem;
        *reinterpret_cast<float4*>(rr) = make_float4(o0[0], o0[1], o0[2], o0[3]);
        *reinterpret_cast<float4*>(rr + p.hw) = make_float4(o1[0], o1[1], o1[2], o1[3]);
        *reinterpret_cast<float4*>(rr + 2 * p.hw) = make_float4(o2[0], o2[1], o2[2], o2[3]);
        float* st = p.out.stats + (int64_t)c.b * (MIX ? PD_STATS_MIXTURE : PD_STATS_PLAIN) * p.hw + rem;
        *reinterpret_cast<float4*>(st) = make_float4(Ml2[0], Ml2[1], Ml2[2], Ml2[3]);
        *reinterpret_cast<float4*>(st + p.hw) = make_float4(S[0], S[1], S[2], S[3]);
        if (MIX) {
            *reinterpret_cast<float4*>(st + 2 * p.hw) = make_float4(A[0], A[1], A[2], A[3]);
            *reinterpret_cast<float4*>(st + 3 * p.hw) = make_float4(dq[0], dq[1], dq[2], dq[3]);
            *reinterpret_cast<float4*>(p.out.nll + (int64_t)c.b * p.hw + rem) = make_float4(nl[0], nl[1], nl[2], nl[3]);
            if (p.d.automask) *reinterpret_cast<float4*>(p.out.nll_auto + (int64_t)c.b * p.hw + rem) = make_float4(na[0], na[1], na[2], na[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward (gather formulation)
// ------------------------------------------------------------------------------------------------
// Exchange buffers per plane of a chunk: E0[x] = dl(x)*w0(x), E1[x] = dl(x)*w1(x) indexed by TARGET x
// (and the same pair for sigma with MIX).  The gradient of source column j of the plane's row is then
//     g[j] = E0[j - k0] + E1[j - k0 - 1]
// because target x reads taps x+k0 (weight w0) and x+k0+1 (weight w1) on a regular plane.
template <int RB>
__device__ __forceinline__ void gather_regular(const float* e0, const float* e1, int x0, int k0, int W, float g[RP]) {
    // s = x0 - k0 = 4q + RB ; E0 needs [s, s+4) ; E1 needs [s-1, s+3)
    const int a0 = x0 - k0 - RB;
    float v[8];
    win_smem(e0, a0, W, v);
#pragma unroll
    for (int i = 0; i < RP; ++i) g[i] = v[RB + i];
    if (RB >= 1) {
        win_smem(e1, a0, W, v);
#pragma unroll
        for (int i = 0; i < RP; ++i) g[i] += v[RB - 1 + i];
    } else {
        win_smem(e1, a0 - 4, W, v);
#pragma unroll
        for (int i = 0; i < RP; ++i) g[i] += v[3 + i];
    }
}

__device__ __forceinline__ void gather_irregular(float wm1, const PlaneRow& pr, const float* e0, const float* e1, int x0, int W, float g[RP]) {
    // re-evaluate the forward coordinate of the candidate targets and test which of their two taps land
    // on column j (bit-faithful to the forward pass)
#pragma unroll
    for (int i = 0; i < RP; ++i) {
        const int j = x0 + i;
        float acc = 0.0f;
        for (int dx = -2; dx <= 1; ++dx) {
            const int x = j - pr.k0 + dx;
            if ((unsigned)x >= (unsigned)W) continue;
            float uh = roundtrip(__fadd_rn((float)x, pr.sd), wm1);
            uh = fminf(fmaxf(uh, -2.0f), (float)(W + 1));
            const int ix = __float2int_rd(uh);
            if (ix == j) acc += e0[x];
            if (ix + 1 == j) acc += e1[x];
        }
        g[i] = acc;
    }
}

__device__ __forceinline__ void gather_any(const WarpParams& p, const PlaneRow& pr, const float* e0, const float* e1, int x0, int W, float g[RP]) {
    if (pr.regular) {
        switch ((-pr.k0) & 3) {
            case 0: gather_regular<0>(e0, e1, x0, pr.k0, W, g); break;
            case 1: gather_regular<1>(e0, e1, x0, pr.k0, W, g); break;
            case 2: gather_regular<2>(e0, e1, x0, pr.k0, W, g); break;
            default: gather_regular<3>(e0, e1, x0, pr.k0, W, g); break;
        }
    } else {
        gather_irregular(p.wm1, pr, e0, e1, x0, W, g);
    }
}

template <bool MIX, int MASKMODE, bool WANT_DISP, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) warp_composite_bwd_rows(const WarpParams p, const float rcp_wm1, const int chunk) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int W = p.d.W, H = p.d.H, N = p.d.N, pitch = W + 2 * ROW_PAD;
    constexpr int NE = MIX ? 4 : 2;
    float* srow = reinterpret_cast<float*>(smem_raw);
    float* ex = srow + 3 * pitch;                       // [chunk][NE][pitch]
    float* gacc = ex + (size_t)chunk * NE * pitch;      // [N]
    PlaneRow* prs = reinterpret_cast<PlaneRow*>(gacc + N);
    for (int i = threadIdx.x; i < (3 + chunk * NE) * pitch; i += blockDim.x) srow[i] = 0.0f;  // incl. all pads

    RowCtx c;
    c.W = W, c.H = H, c.N = N;
    c.x0 = threadIdx.x * RP;
    c.active = c.x0 < W;
#pragma unroll
    for (int i = 0; i < RP; ++i) c.xf[i] = (float)(c.x0 + i);
    const int rows_total = p.d.B * H;
    const float l2e = kLog2e;
    const int lane = threadIdx.x & 31;

    for (int row = blockIdx.x; row < rows_total; row += gridDim.x) {
        c.b = row / H;
        c.y = row - c.b * H;
        c.g = row_geom(c.y, H, p.hm1);
        // this CTA owns gradient row y: weight of source row y among the (at most two) vertical taps;
        // the other tap carries a weight <= 6.2e-5 and is dropped (deviations)
        const float wyp = (c.g.r0 == c.y) ? c.g.w0 : ((c.g.r1 == c.y) ? c.g.w1 : 0.0f);
        __syncthreads();
        stage_row<MASKMODE>(p, c, srow, prs);
        if (WANT_DISP)
            for (int n = threadIdx.x; n < N; n += blockDim.x) gacc[n] = 0.0f;
        __syncthreads();
        const int64_t rem = (int64_t)c.y * W + c.x0;
        float g0[RP], g1[RP], g2[RP], Gbar[RP], Ml2[RP], invS[RP];
        float tr[RP], tg[RP], tb[RP], Zinv[RP], gD[RP], gDD[RP];
        if (c.active) {
            const float gph = upstream_scale(p);
            const int64_t gi = (int64_t)c.b * p.chw3 + rem, gpix = (int64_t)c.b * p.hw + rem;
            const float* rp = p.out.rgb_rec + (int64_t)c.b * p.chw3 + rem;
            float4 a = upstream_rgb4(p, gph, gi, gpix), bq = upstream_rgb4(p, gph, gi + p.hw, gpix), cq = upstream_rgb4(p, gph, gi + 2 * p.hw, gpix);
            float4 ra = ldg4(rp), rb = ldg4(rp + p.hw), rc = ldg4(rp + 2 * p.hw);
            g0[0] = a.x, g0[1] = a.y, g0[2] = a.z, g0[3] = a.w;
            g1[0] = bq.x, g1[1] = bq.y, g1[2] = bq.z, g1[3] = bq.w;
            g2[0] = cq.x, g2[1] = cq.y, g2[2] = cq.z, g2[3] = cq.w;
            Gbar[0] = a.x * ra.x + bq.x * rb.x + cq.x * rc.x;
            Gbar[1] = a.y * ra.y + bq.y * rb.y + cq.y * rc.y;
            Gbar[2] = a.z * ra.z + bq.z * rb.z + cq.z * rc.z;
            Gbar[3] = a.w * ra.w + bq.w * rb.w + cq.w * rc.w;
            const float* st = p.out.stats + (int64_t)c.b * (MIX ? PD_STATS_MIXTURE : PD_STATS_PLAIN) * p.hw + rem;
            float4 m4 = ldg4(st), s4 = ldg4(st + p.hw);
            Ml2[0] = m4.x, Ml2[1] = m4.y, Ml2[2] = m4.z, Ml2[3] = m4.w;
            float Sv[RP] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
            for (int i = 0; i < RP; ++i) invS[i] = 1.0f / Sv[i];
            if (MIX) {
                const float* tp = p.in.tgt + (int64_t)c.b * p.chw3 + rem;
                float4 ta = ldg4(tp), tbq = ldg4(tp + p.hw), tcq = ldg4(tp + 2 * p.hw);
                tr[0] = ta.x, tr[1] = ta.y, tr[2] = ta.z, tr[3] = ta.w;
                tg[0] = tbq.x, tg[1] = tbq.y, tg[2] = tbq.z, tg[3] = tbq.w;
                tb[0] = tcq.x, tb[1] = tcq.y, tb[2] = tcq.z, tb[3] = tcq.w;
                float4 a4 = ldg4(st + 2 * p.hw), d4 = ldg4(st + 3 * p.hw);
                float4 gn = upstream_nll4(p, gph, gpix);
                float Av[RP] = {a4.x, a4.y, a4.z, a4.w}, Dv[RP] = {d4.x, d4.y, d4.z, d4.w}, gv[RP] = {gn.x, gn.y, gn.z, gn.w};
#pragma unroll
                for (int i = 0; i < RP; ++i) {
                    Zinv[i] = Sv[i] / Av[i];          // 1/Z, Z = sum pi/sigma = A/S
                    gD[i] = -gv[i] / Dv[i];           // d loss / d D, nll = -log D
                    gDD[i] = gD[i] * (Dv[i] - 1e-7f); // = sum_k pi_k P_k
                }
            }
        }
        for (int n0 = 0; n0 < N; n0 += chunk) {
            const int n1 = min(n0 + chunk, N);
            // ---------------- phase A: per-target contributions ----------------
            PlaneWin lw_cur, sw_cur, lw_nxt, sw_nxt;
            if (c.active) prefetch_plane<MIX>(p, c, prs[n0], n0, lw_cur, sw_cur);
            for (int n = n0; n < n1; ++n) {
                const PlaneRow pr = prs[n];
                float* e = ex + (size_t)(n - n0) * NE * pitch + ROW_PAD;
                float gdsum = 0.0f;
                if (c.active) {
                    if (n + 1 < n1) prefetch_plane<MIX>(p, c, prs[n + 1], n + 1, lw_nxt, sw_nxt);
                    PlaneSamples<MIX> sm;
                    sample_any<MIX, WANT_DISP>(p, c, srow, pr, n, rcp_wm1, lw_cur, sw_cur, sm);
                    float m[RP];
                    const bool masked = plane_mask<MASKMODE>(p, c, n, pr.m, m);
                    float e0[RP], e1[RP], f0[RP], f1[RP];
#pragma unroll
                    for (int i = 0; i < RP; ++i) {
                        float l = sm.l[i], cr = sm.cr[i], cg = sm.cg[i], cb = sm.cb[i];
                        if (masked) { l *= m[i], cr *= m[i], cg *= m[i], cb *= m[i]; }
                        float pi = fast_exp2(fmaf(l, l2e, -Ml2[i])) * invS[i];
                        float Gn = g0[i] * cr + g1[i] * cg + g2[i] * cb;
                        float dl, dsg = 0.0f, dcr = 0.0f, dcg = 0.0f, dcb = 0.0f;
                        if (MIX) {
                            float sraw = masked ? sm.s[i] * m[i] : sm.s[i];
                            float sg = fminf(fmaxf(sraw, 0.01f), 1.0f);
                            float inv = fast_rcp(sg);
                            float w = pi * inv * Zinv[i];
                            float err = (fabsf(cr - tr[i]) + fabsf(cg - tg[i]) + fabsf(cb - tb[i])) * (1.0f / 3.0f);
                            float lap = 0.5f * fast_exp2(-err * inv * l2e) * inv;
                            float P = (Gn - Gbar[i]) * inv * Zinv[i] + gD[i] * lap;
                            dl = pi * (P - gDD[i]);
                            float dsgt = -(Gn - Gbar[i]) * w * inv + gD[i] * pi * lap * (err - sg) * inv * inv;
                            dsg = (sraw >= 0.01f && sraw <= 1.0f) ? dsgt : 0.0f;  // clamp backward
                            if (WANT_DISP) {
                                float ce = -gD[i] * pi * lap * inv * (1.0f / 3.0f);
                                dcr = w * g0[i] + ce * ((cr > tr[i]) ? 1.0f : ((cr < tr[i]) ? -1.0f : 0.0f));
                                dcg = w * g1[i] + ce * ((cg > tg[i]) ? 1.0f : ((cg < tg[i]) ? -1.0f : 0.0f));
                                dcb = w * g2[i] + ce * ((cb > tb[i]) ? 1.0f : ((cb < tb[i]) ? -1.0f : 0.0f));
                            }
                        } else {
                            dl = pi * (Gn - Gbar[i]);
                            if (WANT_DISP) { dcr = pi * g0[i], dcg = pi * g1[i], dcb = pi * g2[i]; }
                        }
                        if (WANT_DISP) {
                            float gx = dcr * sm.dr[i] + dcg * sm.dg[i] + dcb * sm.db[i] + dl * sm.dl[i];
                            if (MIX) gx = fmaf(dsg, sm.ds[i], gx);
                            gdsum = masked ? fmaf(gx, m[i], gdsum) : gdsum + gx;
                        }
                        float dy = dl * wyp;
                        if (masked) dy *= m[i];
                        e0[i] = dy * sm.w0[i], e1[i] = dy * sm.w1[i];
                        if (MIX) {
                            float sy = dsg * wyp;
                            if (masked) sy *= m[i];
                            f0[i] = sy * sm.w0[i], f1[i] = sy * sm.w1[i];
                        }
                    }
                    *reinterpret_cast<float4*>(e + c.x0) = make_float4(e0[0], e0[1], e0[2], e0[3]);
                    *reinterpret_cast<float4*>(e + pitch + c.x0) = make_float4(e1[0], e1[1], e1[2], e1[3]);
                    if (MIX) {
                        *reinterpret_cast<float4*>(e + 2 * pitch + c.x0) = make_float4(f0[0], f0[1], f0[2], f0[3]);
                        *reinterpret_cast<float4*>(e + 3 * pitch + c.x0) = make_float4(f1[0], f1[1], f1[2], f1[3]);
                    }
                    lw_cur = lw_nxt;
                    if (MIX) sw_cur = sw_nxt;
                }
                if (WANT_DISP) {
                    float s = warp_sum(gdsum);
                    if (lane == 0 && s != 0.0f) atomicAdd(gacc + n, s * p.d.disp_sign);
                }
            }
            __syncthreads();
            // ---------------- phase B: gather per source column, one coalesced store per row ----------------
            if (c.active) {
                for (int n = n0; n < n1; ++n) {
                    const PlaneRow pr = prs[n];
                    const float* e = ex + (size_t)(n - n0) * NE * pitch + ROW_PAD;
                    const int64_t o = (((int64_t)c.b * N + n) * H + c.y) * W + c.x0;
                    float g[RP];
                    if (p.gin.g_logits) {
                        gather_any(p, pr, e, e + pitch, c.x0, W, g);
                        __stcs(reinterpret_cast<float4*>(p.gin.g_logits + o), make_float4(g[0], g[1], g[2], g[3]));
                    }
                    if (MIX && p.gin.g_sigma) {
                        gather_any(p, pr, e + 2 * pitch, e + 3 * pitch, c.x0, W, g);
                        __stcs(reinterpret_cast<float4*>(p.gin.g_sigma + o), make_float4(g[0], g[1], g[2], g[3]));
                    }
                }
            }
            __syncthreads();
        }
        if (WANT_DISP) {
            for (int n = threadIdx.x; n < N; n += blockDim.x) {
                float v = gacc[n];
                if (v != 0.0f) atomicAdd(p.gin.g_disp + soff(p.gin.g_disp_stride, c.b, n, c.y, 0), v);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
inline bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

inline int rows_mask_mode(const WarpParams& p) {
    return (p.d.mask_dtype == PD_MASK_NONE || p.d.mask_stride.x == 0) ? MASK_ROW : MASK_DENSE_F32;
}

inline bool rows_path_supported(const WarpParams& p) {
    if (p.d.warp_type != PD_WARP_DISP) return false;
    const int W = p.d.W;
    if (W % 4 != 0 || W < 8 || W > 1280) return false;
    if ((W & (W - 1)) == 0) return false;  // W-1 all ones: the division-free round trip is not guaranteed exact
    if (p.d.disp_stride.x != 0) return false;  // disparity must be constant along x
    if (p.d.mask_dtype != PD_MASK_NONE && p.d.mask_stride.x != 0) {
        // dense mask: fp32, 4 consecutive elements per thread must be one aligned vector
        if (p.d.mask_dtype != PD_MASK_F32 || p.d.mask_stride.x != 1) return false;
        if (p.d.mask_stride.y % 4 || p.d.mask_stride.n % 4 || p.d.mask_stride.b % 4 || !aligned16(p.in.mask)) return false;
    }
    // w.r.t. disp_layered the fast backward supports no gradient or a gradient reduced over x
    if (p.gin.g_disp && p.gin.g_disp_stride.x != 0) return false;
    const void* ptrs[] = {p.in.src, p.in.tgt, p.in.logits, p.in.sigma, p.out.rgb_rec, p.out.stats, p.out.nll, p.out.nll_auto,
                          p.gout.g_rgb_rec, p.gout.g_nll, p.gout.g_unit, p.gout.g_unit_nll, p.gout.g_pred, p.gout.mask_novel,
                          p.gin.g_logits, p.gin.g_sigma};
    for (const void* q : ptrs)
        if (q && !aligned16(q)) return false;
    return true;
}

inline int rows_grid(int rows_total, int threads, size_t smem, const void* kernel) {
    const long long g = (long long)sm_count() * resident_ctas(kernel, threads, smem);
    return (int)(g < rows_total ? g : rows_total);
}

inline int rows_threads(int W) { return ((W / RP + 31) / 32) * 32; }
inline size_t rows_smem_fwd(int N, int W) { return (size_t)3 * (W + 2 * ROW_PAD) * sizeof(float) + (size_t)N * sizeof(PlaneRow); }

// Opt a kernel in to > 48 KB of dynamic shared memory.  Done once per kernel and size (never again for a
// size already granted) so that steady-state launches issue no attribute call — those are not capturable
// into a CUDA graph.
template <typename K>
inline void rows_launch_cfg(K kern, size_t smem) {
    smem_optin((const void*)kern, smem);
}

template <bool MIX, int MASKMODE>
inline void launch_fwd_rows_t(const WarpParams& p, cudaStream_t st) {
    const int threads = rows_threads(p.d.W);
    const size_t smem = rows_smem_fwd(p.d.N, p.d.W);
    if (threads <= 160) {
        auto kern = warp_composite_fwd_rows<MIX, MASKMODE, 160, MIX ? 2 : 4>;
        rows_launch_cfg(kern, smem);
        kern<<<rows_grid(p.d.B * p.d.H, threads, smem, (const void*)kern), threads, smem, st>>>(p, rows_rcp(p.d.W));
    } else {
        auto kern = warp_composite_fwd_rows<MIX, MASKMODE, 320, MIX ? 1 : 2>;
        rows_launch_cfg(kern, smem);
        kern<<<rows_grid(p.d.B * p.d.H, threads, smem, (const void*)kern), threads, smem, st>>>(p, rows_rcp(p.d.W));
    }
}

inline void launch_fwd_rows(const WarpParams& p, cudaStream_t st) {
    const int mm = rows_mask_mode(p);
    if (p.d.mixture) {
        if (mm == MASK_ROW) launch_fwd_rows_t<true, MASK_ROW>(p, st);
        else launch_fwd_rows_t<true, MASK_DENSE_F32>(p, st);
    } else {
        if (mm == MASK_ROW) launch_fwd_rows_t<false, MASK_ROW>(p, st);
        else launch_fwd_rows_t<false, MASK_DENSE_F32>(p, st);
    }
}

inline int rows_bwd_chunk(int N, int W, bool mix) {
    const size_t per_plane = (size_t)(mix ? 4 : 2) * (W + 2 * ROW_PAD) * sizeof(float);
    const size_t budget = (W > 640) ? 64 * 1024 : 36 * 1024;
    int c = (int)(budget / per_plane);
    return c < 1 ? 1 : (c > N ? N : c);
}

inline size_t rows_smem_bwd(int N, int W, bool mix, int chunk) {
    const size_t pitch = W + 2 * ROW_PAD;
    return (3 + (size_t)chunk * (mix ? 4 : 2)) * pitch * sizeof(float) + (size_t)N * sizeof(float) + (size_t)N * sizeof(PlaneRow);
}

template <bool MIX, int MASKMODE, bool WANT_DISP>
inline void launch_bwd_rows_t(const WarpParams& p, cudaStream_t st) {
    const int threads = rows_threads(p.d.W);
    const int chunk = rows_bwd_chunk(p.d.N, p.d.W, MIX);
    const size_t smem = rows_smem_bwd(p.d.N, p.d.W, MIX, chunk);
    if (threads <= 160) {
        auto kern = warp_composite_bwd_rows<MIX, MASKMODE, WANT_DISP, 160, MIX ? 2 : 3>;
        rows_launch_cfg(kern, smem);
        kern<<<rows_grid(p.d.B * p.d.H, threads, smem, (const void*)kern), threads, smem, st>>>(p, rows_rcp(p.d.W), chunk);
    } else {
        auto kern = warp_composite_bwd_rows<MIX, MASKMODE, WANT_DISP, 320, 1>;
        rows_launch_cfg(kern, smem);
        kern<<<rows_grid(p.d.B * p.d.H, threads, smem, (const void*)kern), threads, smem, st>>>(p, rows_rcp(p.d.W), chunk);
    }
}

template <bool MIX, int MASKMODE>
inline void launch_bwd_rows_m(const WarpParams& p, cudaStream_t st) {
    if (p.gin.g_disp) launch_bwd_rows_t<MIX, MASKMODE, true>(p, st);
    else launch_bwd_rows_t<MIX, MASKMODE, false>(p, st);
}

inline void launch_bwd_rows(const WarpParams& p, cudaStream_t st) {
    const int mm = rows_mask_mode(p);
    if (p.d.mixture) {
        if (mm == MASK_ROW) launch_bwd_rows_m<true, MASK_ROW>(p, st);
        else launch_bwd_rows_m<true, MASK_DENSE_F32>(p, st);
    } else {
        if (mm == MASK_ROW) launch_bwd_rows_m<false, MASK_ROW>(p, st);
        else launch_bwd_rows_m<false, MASK_DENSE_F32>(p, st);
    }
}

}  // namespace pd
