// Photometric term of Trainer.compute_losses for one target side (trainer.py:720-742), plus the
// 0.85*SSIM + 0.15*L1 mode of compute_reprojection_loss (trainer.py:687-699, layers.py:276-306).
// One 32x8 pixel tile per CTA; SSIM windows are evaluated from a reflect-padded shared-memory tile.
#pragma once
#include "pd_device.cuh"

namespace pd {

constexpr int LT_W = 32, LT_H = 8, LT_THREADS = LT_W * LT_H;
constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;  // layers.py:289-290

struct LossParams {
    pd_loss_desc d;
    pd_loss_in in;
    pd_loss_out out;
    pd_loss_grad_out gout;
    pd_loss_grad_in gin;
    float* partials;  // [gridDim.x*gridDim.y*gridDim.z]
    int64_t hw;
};

__device__ __forceinline__ int reflect(int i, int n) {  // nn.ReflectionPad2d(1) index map
    return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i);
}

// blended prediction, trainer.py:724-726
__device__ __forceinline__ float blend_pred(float rec, float tgt, float m) { return rec * m + tgt * (1.0f - m); }

struct WinStats {
    float mx, my, sx, sy, sxy;
};

// 3x3 box statistics around (cy,cx) of a padded tile with row pitch `pitch` (layers.py:296-301)
__device__ __forceinline__ WinStats win_stats(const float* X, const float* Y, int cy, int cx, int pitch) {
    float sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            float a = X[(cy + dy) * pitch + cx + dx], b = Y[(cy + dy) * pitch + cx + dx];
            sx += a, sy += b;
            sxx = fmaf(a, a, sxx), syy = fmaf(b, b, syy), sxy = fmaf(a, b, sxy);
        }
    const float k = 1.0f / 9.0f;
    WinStats w;
    w.mx = sx * k, w.my = sy * k;
    w.sx = sxx * k - w.mx * w.mx;
    w.sy = syy * k - w.my * w.my;
    w.sxy = sxy * k - w.mx * w.my;
    return w;
}

__device__ __forceinline__ float ssim_val(const WinStats& w, float& n, float& d) {
    n = (2.0f * w.mx * w.my + kC1) * (2.0f * w.sxy + kC2);
    d = (w.mx * w.mx + w.my * w.my + kC1) * (w.sx + w.sy + kC2);
    return (1.0f - n / d) * 0.5f;  // layers.py:303-306 before the clamp
}

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float t = 0.0f;
    if (threadIdx.x < 32) {
        t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0f;
        t = warp_sum(t);
    }
    return t;  // valid in thread 0
}

// Loads the (LT_H+2*HALO) x (LT_W+2*HALO) reflect-padded tiles of pred / tgt / src for one channel set.
template <int HALO, bool NEED_SRC, bool HASMASK>
__device__ __forceinline__ void load_tiles(const LossParams& p, int b, int ty0, int tx0, float* sP, float* sT, float* sS) {
    constexpr int PW = LT_W + 2 * HALO, PH = LT_H + 2 * HALO;
    const int H = p.d.H, W = p.d.W;
    for (int i = threadIdx.x; i < PW * PH; i += LT_THREADS) {
        int ly = i / PW, lx = i - ly * PW;
        // padded coordinate -> image coordinate; clamp first so far-outside halo cells stay addressable
        int gy = reflect(min(max(ty0 + ly - HALO, -1), H), H);
        int gx = reflect(min(max(tx0 + lx - HALO, -1), W), W);
        gy = min(max(gy, 0), H - 1);
        gx = min(max(gx, 0), W - 1);
        int64_t o = (int64_t)gy * W + gx;
        float m = HASMASK ? __ldg(p.in.mask_novel + (int64_t)b * p.hw + o) : 1.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int64_t oc = ((int64_t)b * 3 + c) * p.hw + o;
            float t = __ldg(p.in.tgt + oc), r = __ldg(p.in.rgb_rec + oc);
            sT[c * PW * PH + i] = t;
            sP[c * PW * PH + i] = HASMASK ? blend_pred(r, t, m) : r;
            if (NEED_SRC) sS[c * PW * PH + i] = __ldg(p.in.src + oc);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int MODE, bool AUTO, bool HASMASK>
__global__ void __launch_bounds__(LT_THREADS) photometric_fwd_kernel(const LossParams p) {
    constexpr bool SSIM = (MODE == PD_LOSS_SSIM_L1);
    constexpr int PW = LT_W + 2, PH = LT_H + 2;
    __shared__ float sP[SSIM ? 3 * PW * PH : 1], sT[SSIM ? 3 * PW * PH : 1], sS[(SSIM && AUTO) ? 3 * PW * PH : 1];
    __shared__ float red[LT_THREADS / 32];
    const int H = p.d.H, W = p.d.W;
    const int b = blockIdx.z, ty0 = blockIdx.y * LT_H, tx0 = blockIdx.x * LT_W;
    const int ly = threadIdx.x / LT_W, lx = threadIdx.x % LT_W;
    const int y = ty0 + ly, x = tx0 + lx;
    const bool live = (y < H) && (x < W);
    if (SSIM) {
        load_tiles<1, AUTO, HASMASK>(p, b, ty0, tx0, sP, sT, sS);
        __syncthreads();
    }
    float ph = 0.0f;
    if (live) {
        const int64_t o = (int64_t)y * W + x;
        const float m = HASMASK ? __ldg(p.in.mask_novel + (int64_t)b * p.hw + o) : 1.0f;
        if (MODE == PD_LOSS_MIXTURE) {
            ph = __ldg(p.in.nll + (int64_t)b * p.hw + o);
            if (AUTO) ph = fminf(ph, __ldg(p.in.nll_auto + (int64_t)b * p.hw + o));
            if (HASMASK) ph *= m;
            if (HASMASK) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    int64_t oc = ((int64_t)b * 3 + c) * p.hw + o;
                    p.out.pred[oc] = blend_pred(__ldg(p.in.rgb_rec + oc), __ldg(p.in.tgt + oc), m);
                }
            }
        } else {
            float l1 = 0, l1a = 0, ss = 0, ssa = 0;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                int64_t oc = ((int64_t)b * 3 + c) * p.hw + o;
                float t = __ldg(p.in.tgt + oc), r = __ldg(p.in.rgb_rec + oc);
                float pr = HASMASK ? blend_pred(r, t, m) : r;
                if (HASMASK) p.out.pred[oc] = pr;
                l1 += fabsf(pr - t);
                float s = 0.0f;
                if (AUTO) {
                    s = __ldg(p.in.src + oc);
                    l1a += fabsf(s - t);
                }
                if (SSIM) {
                    float n, d;
                    const int cy = ly + 1, cx = lx + 1;
                    float v = ssim_val(win_stats(sP + c * PW * PH, sT + c * PW * PH, cy, cx, PW), n, d);
                    ss += fminf(fmaxf(v, 0.0f), 1.0f);
                    if (AUTO) {
                        float va = ssim_val(win_stats(sS + c * PW * PH, sT + c * PW * PH, cy, cx, PW), n, d);
                        ssa += fminf(fmaxf(va, 0.0f), 1.0f);
                    }
                }
            }
            const float k3 = 1.0f / 3.0f;
            ph = SSIM ? (0.85f * (ss * k3) + 0.15f * (l1 * k3)) : l1 * k3;
            if (AUTO) {
                float pa = SSIM ? (0.85f * (ssa * k3) + 0.15f * (l1a * k3)) : l1a * k3;
                ph = fminf(ph, pa);
            }
        }
        if (p.out.ph_map) p.out.ph_map[(int64_t)b * p.hw + o] = ph;
    }
    float tot = block_sum(ph, red);
    if (threadIdx.x == 0) p.partials[((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = tot;
}

// deterministic second stage: one CTA sums the per-tile partials in a fixed order
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const float* __restrict__ partials, int64_t n, float* __restrict__ out) {
    __shared__ float red[32];
    float acc = 0.0f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += partials[i];
    float t = block_sum(acc, red);
    if (threadIdx.x == 0) *out = t;
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
template <int MODE, bool AUTO, bool HASMASK>
__global__ void __launch_bounds__(LT_THREADS) photometric_bwd_kernel(const LossParams p) {
    constexpr bool SSIM = (MODE == PD_LOSS_SSIM_L1);
    constexpr int PW = LT_W + 4, PH = LT_H + 4;  // padded values: tile + 2
    constexpr int CW = LT_W + 2, CH = LT_H + 2;  // window centres: tile + 1
    __shared__ float sP[SSIM ? 3 * PW * PH : 1], sT[SSIM ? 3 * PW * PH : 1], sS[(SSIM && AUTO) ? 3 * PW * PH : 1];
    __shared__ float cA[SSIM ? 3 * CW * CH : 1], cB[SSIM ? 3 * CW * CH : 1], cC[SSIM ? 3 * CW * CH : 1];
    __shared__ float gate[SSIM ? CW * CH : 1];
    const int H = p.d.H, W = p.d.W;
    const int b = blockIdx.z, ty0 = blockIdx.y * LT_H, tx0 = blockIdx.x * LT_W;
    const int ly = threadIdx.x / LT_W, lx = threadIdx.x % LT_W;
    const int y = ty0 + ly, x = tx0 + lx;
    const bool live = (y < H) && (x < W);
    const float gph = __ldg(p.gout.g_ph_sum);
    const float k3 = 1.0f / 3.0f;

    if (SSIM) {
        load_tiles<2, AUTO, HASMASK>(p, b, ty0, tx0, sP, sT, sS);
        __syncthreads();
        // per-centre coefficients: d ssim_term / d pred_i = A + B*pred_i + C*tgt_i for the 9 window cells
        for (int i = threadIdx.x; i < CW * CH; i += LT_THREADS) {
            int cyl = i / CW, cxl = i - cyl * CW;
            int qy = ty0 + cyl - 1, qx = tx0 + cxl - 1;
            bool inside = (qy >= 0) && (qy < H) && (qx >= 0) && (qx < W);
            float a[3] = {0, 0, 0}, bb[3] = {0, 0, 0}, cc[3] = {0, 0, 0};
            float g = 0.0f;
            if (inside) {
                const int cy = cyl + 1, cx = cxl + 1;  // position in the padded tile
                float l1 = 0, l1a = 0, ss = 0, ssa = 0;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float* X = sP + c * PW * PH;
                    const float* Y = sT + c * PW * PH;
                    WinStats w = win_stats(X, Y, cy, cx, PW);
                    float n, d;
                    float v = ssim_val(w, n, d);
                    ss += fminf(fmaxf(v, 0.0f), 1.0f);
                    l1 += fabsf(X[cy * PW + cx] - Y[cy * PW + cx]);
                    if (v >= 0.0f && v <= 1.0f) {  // clamp backward
                        float a1 = 2.0f * w.mx * w.my + kC1, a2 = 2.0f * w.sxy + kC2;
                        float b1 = w.mx * w.mx + w.my * w.my + kC1, b2 = w.sx + w.sy + kC2;
                        float id = 1.0f / d, nd2 = n * id * id;
                        const float k9 = 1.0f / 9.0f;
                        bb[c] = nd2 * b1 * k9;
                        cc[c] = -a1 * id * k9;
                        a[c] = k9 * ((a1 - a2) * w.my * id + nd2 * (b2 - b1) * w.mx);
                    }
                    if (AUTO) {
                        const float* S = sS + c * PW * PH;
                        float na, da;
                        float va = ssim_val(win_stats(S, Y, cy, cx, PW), na, da);
                        ssa += fminf(fmaxf(va, 0.0f), 1.0f);
                        l1a += fabsf(S[cy * PW + cx] - Y[cy * PW + cx]);
                    }
                }
                g = gph;
                if (AUTO) {
                    float ph = 0.85f * (ss * k3) + 0.15f * (l1 * k3);
                    float pa = 0.85f * (ssa * k3) + 0.15f * (l1a * k3);
                    if (!(ph <= pa)) g = 0.0f;  // min() routes the gradient to the first minimum
                }
            }
            gate[i] = g;
            const float ks = g * 0.85f * k3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                cA[c * CW * CH + i] = a[c] * ks;
                cB[c * CW * CH + i] = bb[c] * ks;
                cC[c * CW * CH + i] = cc[c] * ks;
            }
        }
        __syncthreads();
    }
    if (!live) return;
    const int64_t o = (int64_t)y * W + x;
    const float m = HASMASK ? __ldg(p.in.mask_novel + (int64_t)b * p.hw + o) : 1.0f;
    if (MODE == PD_LOSS_MIXTURE) {
        float nl = __ldg(p.in.nll + (int64_t)b * p.hw + o);
        bool sel = true;
        if (AUTO) sel = nl <= __ldg(p.in.nll_auto + (int64_t)b * p.hw + o);
        p.gin.g_nll[(int64_t)b * p.hw + o] = sel ? gph * m : 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int64_t oc = ((int64_t)b * 3 + c) * p.hw + o;
            float ge = p.gout.g_pred ? __ldg(p.gout.g_pred + oc) : 0.0f;
            p.gin.g_rgb_rec[oc] = ge * m;
        }
        return;
    }
    float gsel = gph;
    float pr[3], tg[3];
    {
        float l1 = 0, l1a = 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int64_t oc = ((int64_t)b * 3 + c) * p.hw + o;
            tg[c] = __ldg(p.in.tgt + oc);
            float r = __ldg(p.in.rgb_rec + oc);
            pr[c] = HASMASK ? blend_pred(r, tg[c], m) : r;
            l1 += fabsf(pr[c] - tg[c]);
            if (AUTO && !SSIM) l1a += fabsf(__ldg(p.in.src + oc) - tg[c]);
        }
        if (SSIM) gsel = gate[(ly + 1) * CW + lx + 1];
        else if (AUTO && !(l1 * k3 <= l1a * k3)) gsel = 0.0f;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int64_t oc = ((int64_t)b * 3 + c) * p.hw + o;
        float diff = pr[c] - tg[c];
        float sg = (diff > 0.0f) ? 1.0f : ((diff < 0.0f) ? -1.0f : 0.0f);
        float g = gsel * (SSIM ? 0.15f : 1.0f) * k3 * sg;
        if (SSIM) {
            float sa = 0, sb = 0, sc = 0;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                int qy = y + dy;
                if (qy < 0 || qy >= H) continue;
                // how many padded rows that mirror onto row y lie inside centre qy's window
                float wy = 1.0f + ((y == 1 && qy == 0) ? 1.0f : 0.0f) + ((y == H - 2 && qy == H - 1) ? 1.0f : 0.0f);
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    int qx = x + dx;
                    if (qx < 0 || qx >= W) continue;
                    float wx = 1.0f + ((x == 1 && qx == 0) ? 1.0f : 0.0f) + ((x == W - 2 && qx == W - 1) ? 1.0f : 0.0f);
                    int ci = c * CW * CH + (ly + 1 + dy) * CW + (lx + 1 + dx);
                    float wgt = wy * wx;
                    sa = fmaf(wgt, cA[ci], sa), sb = fmaf(wgt, cB[ci], sb), sc = fmaf(wgt, cC[ci], sc);
                }
            }
            g += sa + sb * pr[c] + sc * tg[c];
        }
        if (p.gout.g_pred) g += __ldg(p.gout.g_pred + oc);
        p.gin.g_rgb_rec[oc] = HASMASK ? g * m : g;
    }
}

}  // namespace pd
