// General (any warp type, any strides) warp + composite kernels: one thread per target pixel, loop
// over planes, source taps through the read-only cache.  This is the reference-faithful baseline path
// that every specialised kernel is tested against; the row-tiled fast path for stereo disparity warps
// lives in pd_warp_rows.cuh.
//
// Replaces, per target side, trainer.py:533-603 (forward) and what autograd replays through it
// (backward).  See include/planedepth_b200.h for the buffer contract.
#pragma once
#include "pd_device.cuh"

namespace pd {

struct WarpParams {
    pd_warp_desc d;
    pd_warp_in in;
    pd_warp_out out;
    pd_warp_grad_out gout;
    pd_warp_grad_in gin;
    float wm1, hm1;        // W-1, H-1 as fp32
    float depth_c;         // 0.1*0.58*W  (trainer.py:535)
    int64_t hw, chw3;      // H*W, 3*H*W
    int g_disp_dense;      // g_disp has no zero stride -> plain stores
    int warp_aligned_rows; // W % 32 == 0: a warp never straddles image rows
    // Dense-mask row summary kept behind the saved statistics (pd_warp_composite_stats_bytes): per image row a 64-bit set
    // over planes, bit n = "plane n's mask row is not all 1.0".  Written by the streamed forward (its consumers see every
    // mask value anyway), lets the streamed backward leave all-ones mask rows in HBM.  NULL = not kept.
    unsigned long long* mask_rows;
};

// Upstream gradients in pd_warp_grad_out's (optionally fused) form: what pd_photometric_bwd would have written,
//   g_rgb_rec_eff = [g_rgb_rec] + gph * g_unit + [g_pred * mask_novel],   g_nll_eff = [g_nll] + gph * g_unit_nll,
// with gph = ph_scale * g_ph_sum[0] (same operation order as photometric_bwd_kernel: product, then fused add).
__device__ __forceinline__ float upstream_scale(const WarpParams& p) {
    return p.gout.g_ph_sum ? __ldg(p.gout.g_ph_sum) * (p.gout.ph_scale != 0.0f ? p.gout.ph_scale : 1.0f) : 0.0f;
}
// i = offset into [B,3,H,W], pix = offset into [B,1,H,W]
__device__ __forceinline__ float upstream_rgb(const WarpParams& p, float gph, int64_t i, int64_t pix) {
    float g = 0.0f;
    if (p.gout.g_ph_sum) {
        if (p.gout.g_unit) g = gph * __ldg(p.gout.g_unit + i);
        if (p.gout.g_pred) g = fmaf(__ldg(p.gout.g_pred + i), p.gout.mask_novel ? __ldg(p.gout.mask_novel + pix) : 1.0f, g);
    }
    if (p.gout.g_rgb_rec) g += __ldg(p.gout.g_rgb_rec + i);
    return g;
}
__device__ __forceinline__ float upstream_nll(const WarpParams& p, float gph, int64_t pix) {
    float g = (p.gout.g_ph_sum && p.gout.g_unit_nll) ? gph * __ldg(p.gout.g_unit_nll + pix) : 0.0f;
    if (p.gout.g_nll) g += __ldg(p.gout.g_nll + pix);
    return g;
}
__device__ __forceinline__ float4 upstream_rgb4(const WarpParams& p, float gph, int64_t i, int64_t pix) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.gout.g_ph_sum) {
        if (p.gout.g_unit) {
            const float4 u = __ldg(reinterpret_cast<const float4*>(p.gout.g_unit + i));
            g = make_float4(gph * u.x, gph * u.y, gph * u.z, gph * u.w);
        }
        if (p.gout.g_pred) {
            const float4 e = __ldg(reinterpret_cast<const float4*>(p.gout.g_pred + i));
            float4 m = make_float4(1.f, 1.f, 1.f, 1.f);
            if (p.gout.mask_novel) m = __ldg(reinterpret_cast<const float4*>(p.gout.mask_novel + pix));
            g.x = fmaf(e.x, m.x, g.x), g.y = fmaf(e.y, m.y, g.y), g.z = fmaf(e.z, m.z, g.z), g.w = fmaf(e.w, m.w, g.w);
        }
    }
    if (p.gout.g_rgb_rec) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(p.gout.g_rgb_rec + i));
        g.x += r.x, g.y += r.y, g.z += r.z, g.w += r.w;
    }
    return g;
}
__device__ __forceinline__ float4 upstream_nll4(const WarpParams& p, float gph, int64_t pix) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.gout.g_ph_sum && p.gout.g_unit_nll) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(p.gout.g_unit_nll + pix));
        g = make_float4(gph * u.x, gph * u.y, gph * u.z, gph * u.w);
    }
    if (p.gout.g_nll) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(p.gout.g_nll + pix));
        g.x += r.x, g.y += r.y, g.z += r.z, g.w += r.w;
    }
    return g;
}

// Source coordinates (u,v) of target pixel (x,y) on plane n, plus the multiplicative validity mask.
// aux[] receives what the backward needs to chain the coordinate gradient to the warp parameters.
template <int WARP>
__device__ __forceinline__ void plane_coords(const WarpParams& p, int b, int n, int y, int x, float& u, float& v,
                                             float& m, float aux[4]) {
    const float fx = (float)x, fy = (float)y;
    if (WARP == PD_WARP_DISP) {
        // trainer.py:540-554
        float dsp = __ldg(p.in.disp + soff(p.d.disp_stride, b, n, y, x));
        u = __fadd_rn(fx, __fmul_rn(p.d.disp_sign, dsp));
        v = fy;
        m = load_mask(p.in.mask, p.d.mask_dtype, soff(p.d.mask_stride, b, n, y, x));
    } else if (WARP == PD_WARP_HOMOGRAPHY) {
        // layers.py:221-228
        const float* hm = p.in.hmat + ((int64_t)b * p.d.N + n) * 12;
        const float* ik = p.in.cam + (int64_t)b * 9;
        float qx = fmaf(__ldg(hm + 1), fy, __ldg(hm + 0) * fx) + __ldg(hm + 2);
        float qy = fmaf(__ldg(hm + 4), fy, __ldg(hm + 3) * fx) + __ldg(hm + 5);
        float qz = fmaf(__ldg(hm + 7), fy, __ldg(hm + 6) * fx) + __ldg(hm + 8);
        float rx = fmaf(__ldg(ik + 1), fy, __ldg(ik + 0) * fx) + __ldg(ik + 2);
        float ry = fmaf(__ldg(ik + 4), fy, __ldg(ik + 3) * fx) + __ldg(ik + 5);
        float rz = fmaf(__ldg(ik + 7), fy, __ldg(ik + 6) * fx) + __ldg(ik + 8);
        float facing = rx * __ldg(hm + 9) + ry * __ldg(hm + 10) + rz * __ldg(hm + 11);
        bool ok = (facing > 0.0f) && (qz > 1e-7f);
        float zc = (qz < 1e-7f) ? 1e-7f : qz;
        u = __fdiv_rn(qx, zc);
        v = __fdiv_rn(qy, zc);
        m = ok ? 1.0f : 0.0f;
        aux[0] = zc;
        aux[1] = (qz < 1e-7f) ? 0.0f : 1.0f;  // d zc / d qz
    } else {
        // trainer.py:534-538, layers.py:150-156, 169-182
        const float* cm = p.in.cam + (int64_t)b * 21;
        float dsp = __ldg(p.in.disp + soff(p.d.disp_stride, b, n, y, x));
        float Z = __fdiv_rn(p.depth_c, dsp);
        float rx = fmaf(__ldg(cm + 1), fy, __ldg(cm + 0) * fx) + __ldg(cm + 2);
        float ry = fmaf(__ldg(cm + 4), fy, __ldg(cm + 3) * fx) + __ldg(cm + 5);
        float rz = fmaf(__ldg(cm + 7), fy, __ldg(cm + 6) * fx) + __ldg(cm + 8);
        float X = Z * rx, Y = Z * ry, Zc = Z * rz;
        const float* P = cm + 9;
        float cx = fmaf(__ldg(P + 2), Zc, fmaf(__ldg(P + 1), Y, __ldg(P + 0) * X)) + __ldg(P + 3);
        float cy = fmaf(__ldg(P + 6), Zc, fmaf(__ldg(P + 5), Y, __ldg(P + 4) * X)) + __ldg(P + 7);
        float cz = fmaf(__ldg(P + 10), Zc, fmaf(__ldg(P + 9), Y, __ldg(P + 8) * X)) + __ldg(P + 11);
        float den = cz + 1e-7f;
        u = __fdiv_rn(cx, den);
        v = __fdiv_rn(cy, den);
        m = load_mask(p.in.mask, p.d.mask_dtype, soff(p.d.mask_stride, b, n, y, x));
        aux[0] = den;
        // d(cx,cy,cz)/dZ = P[:, :3] . ray
        aux[1] = __ldg(P + 0) * rx + __ldg(P + 1) * ry + __ldg(P + 2) * rz;
        aux[2] = __ldg(P + 4) * rx + __ldg(P + 5) * ry + __ldg(P + 6) * rz;
        aux[3] = __ldg(P + 8) * rx + __ldg(P + 9) * ry + __ldg(P + 10) * rz;
        // Z and dsp are recomputed by the caller when needed
    }
}

struct Sampled {
    float r, g, b, l, s;  // masked samples of rgb, logit, sigma (trainer.py:573-583)
};

template <bool MIX>
__device__ __forceinline__ Sampled sample_plane(const WarpParams& p, int b, int n, const Taps& t, float m) {
    const int W = p.d.W;
    const float* src = p.in.src + (int64_t)b * p.chw3;
    const int64_t pl = ((int64_t)b * p.d.N + n) * p.hw;
    Sampled s;
    s.r = blend(load_taps(src, t, W), t) * m;
    s.g = blend(load_taps(src + p.hw, t, W), t) * m;
    s.b = blend(load_taps(src + 2 * p.hw, t, W), t) * m;
    s.l = blend(load_taps(p.in.logits + pl, t, W), t) * m;
    s.s = MIX ? blend(load_taps(p.in.sigma + pl, t, W), t) * m : 1.0f;
    return s;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int WARP, bool MIX, bool DEBUG>
__global__ void __launch_bounds__(256) warp_composite_fwd_general(const WarpParams p) {
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)p.d.B * p.hw;
    if (pix >= total) return;
    const int b = (int)(pix / p.hw);
    const int rem = (int)(pix - (int64_t)b * p.hw);
    const int y = rem / p.d.W, x = rem - y * p.d.W;
    const int N = p.d.N, W = p.d.W, H = p.d.H;

    float tr = 0, tg = 0, tb = 0, err_auto = 0;
    if (MIX) {
        const float* tp = p.in.tgt + (int64_t)b * p.chw3 + rem;
        tr = __ldg(tp), tg = __ldg(tp + p.hw), tb = __ldg(tp + 2 * p.hw);
        if (p.d.automask) {
            const float* sp = p.in.src + (int64_t)b * p.chw3 + rem;
            err_auto = (fabsf(__ldg(sp) - tr) + fabsf(__ldg(sp + p.hw) - tg) + fabsf(__ldg(sp + 2 * p.hw) - tb)) * (1.0f / 3.0f);
        }
    }

    // online softmax over planes: every accumulator is a sum of exp(l_n - Mx) * something
    float Mx = -INFINITY, S = 0, A = 0, R0 = 0, R1 = 0, R2 = 0, Q = 0, Qa = 0;
    for (int n = 0; n < N; ++n) {
        float u, v, m, aux[4];
        plane_coords<WARP>(p, b, n, y, x, u, v, m, aux);
        Taps t = make_taps(roundtrip(u, p.wm1), roundtrip(v, p.hm1), W, H);
        Sampled s = sample_plane<MIX>(p, b, n, t, m);
        float mnew = fmaxf(Mx, s.l);
        float sc = fast_exp(Mx - mnew);
        float e = fast_exp(s.l - mnew);
        Mx = mnew;
        S = fmaf(S, sc, e);
        if (MIX) {
            float sg = fminf(fmaxf(s.s, 0.01f), 1.0f);  // trainer.py:597
            float inv = 1.0f / sg;
            float es = e * inv;
            A = fmaf(A, sc, es);
            R0 = fmaf(R0, sc, es * s.r);
            R1 = fmaf(R1, sc, es * s.g);
            R2 = fmaf(R2, sc, es * s.b);
            float err = (fabsf(s.r - tr) + fabsf(s.g - tg) + fabsf(s.b - tb)) * (1.0f / 3.0f);
            Q = fmaf(Q, sc, e * (0.5f * fast_exp(-err * inv) * inv));  // layers.py:454-455
            Qa = fmaf(Qa, sc, e * (0.5f * fast_exp(-err_auto * inv) * inv));
        } else {
            R0 = fmaf(R0, sc, e * s.r);
            R1 = fmaf(R1, sc, e * s.g);
            R2 = fmaf(R2, sc, e * s.b);
        }
    }
    const float invS = 1.0f / S;
    const float invD = MIX ? 1.0f / A : invS;
    float* rr = p.out.rgb_rec + (int64_t)b * p.chw3 + rem;
    rr[0] = R0 * invD;
    rr[p.hw] = R1 * invD;
    rr[2 * p.hw] = R2 * invD;
    float* st = p.out.stats + (int64_t)b * (MIX ? PD_STATS_MIXTURE : PD_STATS_PLAIN) * p.hw + rem;
    st[0] = Mx * kLog2e;  // reference logit in log2 units (shared convention with the row-tiled kernels)
    st[p.hw] = S;
    if (MIX) {
        float D = Q * invS + 1e-7f;  // layers.py:466
        st[2 * p.hw] = A;
        st[3 * p.hw] = D;
        p.out.nll[pix] = -logf(D);
        if (p.d.automask) p.out.nll_auto[pix] = -logf(Qa * invS + 1e-7f);
    }
    if (DEBUG) {
        for (int n = 0; n < N; ++n) {
            float u, v, m, aux[4];
            plane_coords<WARP>(p, b, n, y, x, u, v, m, aux);
            Taps t = make_taps(roundtrip(u, p.wm1), roundtrip(v, p.hm1), W, H);
            Sampled s = sample_plane<MIX>(p, b, n, t, m);
            float e = fast_exp2(fmaf(s.l, kLog2e, -Mx * kLog2e));
            const int64_t o = ((int64_t)b * N + n) * p.hw + rem;
            if (p.out.rgb_rec_layered) {
                float* q = p.out.rgb_rec_layered + ((int64_t)b * N + n) * p.chw3 + rem;
                q[0] = s.r, q[p.hw] = s.g, q[2 * p.hw] = s.b;
            }
            if (p.out.logit_rec) p.out.logit_rec[o] = s.l;
            float pi = e * invS, prob = pi;
            if (MIX) {
                float sg = fminf(fmaxf(s.s, 0.01f), 1.0f);
                prob = (e / sg) * invD;
                if (p.out.sigma_rec) p.out.sigma_rec[o] = sg;
                if (p.out.pi_rec) p.out.pi_rec[o] = pi;
            }
            if (p.out.probability_rec) p.out.probability_rec[o] = prob;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward (scatter formulation: atomics into zero-filled gradient planes)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void scatter_taps(float* __restrict__ plane, const Taps& t, int W, float g) {
    float* r0 = plane + (int64_t)t.y0 * W + t.x0;
    float* r1 = r0 + W;
    float w;
    if (t.in_y0 && t.in_x0 && (w = t.wx0 * t.wy0) != 0.0f) atomicAdd(r0, g * w);
    if (t.in_y0 && t.in_x1 && (w = t.wx1 * t.wy0) != 0.0f) atomicAdd(r0 + 1, g * w);
    if (t.in_y1 && t.in_x0 && (w = t.wx0 * t.wy1) != 0.0f) atomicAdd(r1, g * w);
    if (t.in_y1 && t.in_x1 && (w = t.wx1 * t.wy1) != 0.0f) atomicAdd(r1 + 1, g * w);
}

template <int WARP, bool MIX>
__global__ void __launch_bounds__(256) warp_composite_bwd_general(const WarpParams p) {
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)p.d.B * p.hw;
    const bool live = pix < total;
    const int64_t pc = live ? pix : total - 1;  // keep whole warps alive for the shuffles below
    const int b = (int)(pc / p.hw);
    const int rem = (int)(pc - (int64_t)b * p.hw);
    const int y = rem / p.d.W, x = rem - y * p.d.W;
    const int N = p.d.N, W = p.d.W, H = p.d.H;
    const int lane = threadIdx.x & 31;

    const float gph = upstream_scale(p);
    const int64_t gi = (int64_t)b * p.chw3 + rem;
    float g0 = live ? upstream_rgb(p, gph, gi, pc) : 0.0f, g1 = live ? upstream_rgb(p, gph, gi + p.hw, pc) : 0.0f,
          g2 = live ? upstream_rgb(p, gph, gi + 2 * p.hw, pc) : 0.0f;
    const float* rp = p.out.rgb_rec + (int64_t)b * p.chw3 + rem;
    const float Gbar = g0 * __ldg(rp) + g1 * __ldg(rp + p.hw) + g2 * __ldg(rp + 2 * p.hw);
    const float* st = p.out.stats + (int64_t)b * (MIX ? PD_STATS_MIXTURE : PD_STATS_PLAIN) * p.hw + rem;
    const float Ml2 = __ldg(st), invS = 1.0f / __ldg(st + p.hw);
    float tr = 0, tg = 0, tb = 0, invA = 0, gD = 0, gDD = 0;
    if (MIX) {
        const float* tp = p.in.tgt + (int64_t)b * p.chw3 + rem;
        tr = __ldg(tp), tg = __ldg(tp + p.hw), tb = __ldg(tp + 2 * p.hw);
        invA = 1.0f / __ldg(st + 2 * p.hw);
        float D = __ldg(st + 3 * p.hw);
        float gn = live ? upstream_nll(p, gph, pc) : 0.0f;
        gD = -gn / D;            // d loss / d D,  nll = -log D
        gDD = gD * (D - 1e-7f);  // = sum_k pi_k P_k (backward algebra)
    }
    const bool want_coord = (WARP == PD_WARP_HOMOGRAPHY) ? (p.gin.g_hmat != nullptr) : (p.gin.g_disp != nullptr);
    const float* src = p.in.src + (int64_t)b * p.chw3;

    for (int n = 0; n < N; ++n) {
        float u, v, m, aux[4];
        plane_coords<WARP>(p, b, n, y, x, u, v, m, aux);
        Taps t = make_taps(roundtrip(u, p.wm1), roundtrip(v, p.hm1), W, H);
        const int64_t pl = ((int64_t)b * N + n) * p.hw;
        TapVals vr = load_taps(src, t, W), vg = load_taps(src + p.hw, t, W), vb = load_taps(src + 2 * p.hw, t, W);
        TapVals vl = load_taps(p.in.logits + pl, t, W);
        float cr = blend(vr, t) * m, cg = blend(vg, t) * m, cb = blend(vb, t) * m;
        float l = blend(vl, t) * m;
        float pi = fast_exp2(fmaf(l, kLog2e, -Ml2)) * invS;
        float Gn = g0 * cr + g1 * cg + g2 * cb;
        float dl, dcr, dcg, dcb, dsg = 0.0f;
        TapVals vs;
        if (MIX) {
            vs = load_taps(p.in.sigma + pl, t, W);
            float sraw = blend(vs, t) * m;
            float sg = fminf(fmaxf(sraw, 0.01f), 1.0f);
            float inv = 1.0f / sg;
            float w = pi * inv * (1.0f / invS) * invA;  // (e/sg)/A with e = pi*S
            float err = (fabsf(cr - tr) + fabsf(cg - tg) + fabsf(cb - tb)) * (1.0f / 3.0f);
            float lap = 0.5f * fast_exp(-err * inv) * inv;
            float Zinv = (1.0f / invS) * invA;  // 1/Z, Z = sum pi/sg = A/S
            float P = (Gn - Gbar) * inv * Zinv + gD * lap;
            dl = pi * (P - gDD);
            float dsgt = -(Gn - Gbar) * w * inv + gD * pi * lap * (err - sg) * inv * inv;
            dsg = (sraw >= 0.01f && sraw <= 1.0f) ? dsgt : 0.0f;  // clamp backward
            float ce = -gD * pi * lap * inv * (1.0f / 3.0f);
            dcr = w * g0 + ce * ((cr > tr) ? 1.0f : ((cr < tr) ? -1.0f : 0.0f));
            dcg = w * g1 + ce * ((cg > tg) ? 1.0f : ((cg < tg) ? -1.0f : 0.0f));
            dcb = w * g2 + ce * ((cb > tb) ? 1.0f : ((cb < tb) ? -1.0f : 0.0f));
        } else {
            dl = pi * (Gn - Gbar);
            dcr = pi * g0, dcg = pi * g1, dcb = pi * g2;
        }
        if (!live) dl = dsg = dcr = dcg = dcb = 0.0f;
        // through the mask multiply
        dl *= m, dsg *= m, dcr *= m, dcg *= m, dcb *= m;
        if (p.gin.g_logits && dl != 0.0f) scatter_taps(p.gin.g_logits + pl, t, W, dl);
        if (MIX && p.gin.g_sigma && dsg != 0.0f) scatter_taps(p.gin.g_sigma + pl, t, W, dsg);
        if (want_coord) {
            float gx = 0, gy = 0, dx, dy;
            blend_grad(vr, t, dx, dy); gx = fmaf(dcr, dx, gx); gy = fmaf(dcr, dy, gy);
            blend_grad(vg, t, dx, dy); gx = fmaf(dcg, dx, gx); gy = fmaf(dcg, dy, gy);
            blend_grad(vb, t, dx, dy); gx = fmaf(dcb, dx, gx); gy = fmaf(dcb, dy, gy);
            blend_grad(vl, t, dx, dy); gx = fmaf(dl, dx, gx); gy = fmaf(dl, dy, gy);
            if (MIX) { blend_grad(vs, t, dx, dy); gx = fmaf(dsg, dx, gx); gy = fmaf(dsg, dy, gy); }
            if (WARP == PD_WARP_HOMOGRAPHY) {
                float zi = 1.0f / aux[0];
                float gqx = gx * zi, gqy = gy * zi;
                float gqz = -(gx * u + gy * v) * zi * aux[1];
                float fx = (float)x, fy = (float)y;
                float acc[9] = {gqx * fx, gqx * fy, gqx, gqy * fx, gqy * fy, gqy, gqz * fx, gqz * fy, gqz};
                float* dst = p.gin.g_hmat + ((int64_t)b * N + n) * 9;
                if (p.warp_aligned_rows) {
#pragma unroll
                    for (int k = 0; k < 9; ++k) {
                        float s = warp_sum(acc[k]);
                        if (lane == 0 && s != 0.0f) atomicAdd(dst + k, s);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 9; ++k)
                        if (acc[k] != 0.0f) atomicAdd(dst + k, acc[k]);
                }
            } else {
                float gd;
                if (WARP == PD_WARP_DISP) {
                    gd = gx * p.d.disp_sign;
                } else {
                    float di = 1.0f / aux[0];
                    float gcx = gx * di, gcy = gy * di, gcz = -(gx * u + gy * v) * di;
                    float gZ = gcx * aux[1] + gcy * aux[2] + gcz * aux[3];
                    float dsp = __ldg(p.in.disp + soff(p.d.disp_stride, b, n, y, x));
                    gd = -gZ * (p.depth_c / dsp) / dsp;
                }
                float* dst = p.gin.g_disp + soff(p.gin.g_disp_stride, b, n, y, x);
                if (p.g_disp_dense) {
                    if (live) *dst = gd;
                } else if (p.gin.g_disp_stride.x == 0 && p.warp_aligned_rows) {
                    float s = warp_sum(gd);
                    if (lane == 0 && s != 0.0f) atomicAdd(dst, s);
                } else if (gd != 0.0f) {
                    atomicAdd(dst, gd);
                }
            }
        }
    }
}

}  // namespace pd
