// Photometric term of Trainer.compute_losses for one target side (trainer.py:720-742), plus the
// 0.85*SSIM + 0.15*L1 mode of compute_reprojection_loss (trainer.py:687-699, layers.py:276-306).
//
// The forward kernels produce, next to the loss sum, the UNIT gradient d(ph_sum)/d(rgb_rec) (and
// d(ph_sum)/d(nll) in mixture mode) while the tiles they need are still in shared memory; the backward
// is then a pure streaming pass  g_rgb_rec = g_ph_sum * unit + g_pred * m  (autograd replays the
// pads / pools / clamps of layers.py:292-306 instead).
//
// SSIM mode: one 8x64-pixel tile per CTA; stage 1 loads the reflect-padded (tile+2) values of pred /
// target (/ source for the automask) into shared memory, stage 2 evaluates the 3x3 window statistics of
// the (tile+1) window centres and turns each into the three coefficients of its derivative
//   d ssim_term(centre) / d pred(cell) = A + B * pred(cell) + C * tgt(cell)
// (already gated by the automask min and scaled by 0.85/3), stage 3 gathers the 9 centres around every
// tile pixel with the reflect-padding multiplicities.
#pragma once
#include <type_traits>

#include "pd_device.cuh"

namespace pd {

constexpr int LT_W = 64, LT_H = 8, LT_THREADS = 256;
constexpr int LT_PW = LT_W + 4, LT_PH = LT_H + 4;  // padded values: tile + 2 on every side
constexpr int LT_CW = LT_W + 2, LT_CH = LT_H + 2;  // window centres: tile + 1 on every side
constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;  // layers.py:289-290

struct LossParams {
    pd_loss_desc d;
    pd_loss_in in;
    pd_loss_out out;
    pd_loss_grad_out gout;
    pd_loss_grad_in gin;
    float* partials;  // [number of CTAs]
    int64_t hw;
    int64_t total4;   // elementwise kernels: number of 4-pixel groups (0 = scalar tail handling only)
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

// n / d for d in [C1*C2, ~10] (no denormals, no overflow): reciprocal estimate + one Newton step, quotient with a residual
// correction (Markstein) -- the correctly rounded quotient of layers.py:306 without the IEEE-division sequence and its
// slow-path branch (cf. homo_coords() of pd_warp_homo.cuh).  id returns the refined 1 / d.
__device__ __forceinline__ float quotient_rn(float n, float d, float& id) {
    const float r0 = fast_rcp(d);
    id = fmaf(r0, fmaf(-d, r0, 1.0f), r0);
    const float q0 = n * id;
    return fmaf(fmaf(-q0, d, n), id, q0);
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

__device__ __forceinline__ float sgnf(float d) { return (d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : 0.0f); }

// ------------------------------------------------------------------------------------------------
// SSIM + L1 forward (with unit gradient)
// ------------------------------------------------------------------------------------------------
inline size_t ssim_smem_bytes(bool automask, bool want_g) {
    return sizeof(float) * ((size_t)(automask ? 9 : 6) * LT_PW * LT_PH + (want_g ? (size_t)10 * LT_CW * LT_CH : 0));
}

template <bool AUTO, bool HASMASK, bool WANT_G>
__global__ void __launch_bounds__(LT_THREADS) ssim_l1_fwd_kernel(const LossParams p) {
    constexpr int PW = LT_PW, PH = LT_PH, CW = LT_CW, CH = LT_CH;
    extern __shared__ __align__(16) float lsm[];  // ssim_smem_bytes(AUTO, WANT_G)
    float* sP = lsm;
    float* sT = sP + 3 * PW * PH;
    float* sS = sT + 3 * PW * PH;
    float* cA = sS + (AUTO ? 3 * PW * PH : 0);
    float* cB = cA + (WANT_G ? 3 * CW * CH : 0);
    float* cC = cB + (WANT_G ? 3 * CW * CH : 0);
    float* gate = cC + (WANT_G ? 3 * CW * CH : 0);
    __shared__ float red[LT_THREADS / 32];
    const int H = p.d.H, W = p.d.W;
    const int b = blockIdx.z, ty0 = blockIdx.y * LT_H, tx0 = blockIdx.x * LT_W;
    const float k3 = 1.0f / 3.0f;

    // ---- stage 1: reflect-padded tiles (padded coordinate -> image coordinate; cells further than one pixel outside
    // the image are never read by a centre inside it, they only need to stay addressable)
    {
        constexpr int NL = (PW * PH + LT_THREADS - 1) / LT_THREADS;  // positions per thread: all loads are issued before
        float vt[NL][3], vr[NL][3], vs[NL][3], vm[NL];                // the first store so that they overlap
#pragma unroll
        for (int q = 0; q < NL; ++q) {
            const int i = threadIdx.x + q * LT_THREADS;
            const int ic = min(i, PW * PH - 1);
            const int ly = ic / PW, lx = ic - ly * PW;
            int gy = reflect(min(max(ty0 + ly - 2, -1), H), H);
            int gx = reflect(min(max(tx0 + lx - 2, -1), W), W);
            gy = min(max(gy, 0), H - 1);
            gx = min(max(gx, 0), W - 1);
            const int64_t o = (int64_t)gy * W + gx;
            vm[q] = HASMASK ? __ldg(p.in.mask_novel + (int64_t)b * p.hw + o) : 1.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int64_t oc = ((int64_t)b * 3 + c) * p.hw + o;
                vt[q][c] = __ldg(p.in.tgt + oc);
                vr[q][c] = __ldg(p.in.rgb_rec + oc);
                vs[q][c] = AUTO ? __ldg(p.in.src + oc) : 0.0f;
            }
        }
#pragma unroll
        for (int q = 0; q < NL; ++q) {
            const int i = threadIdx.x + q * LT_THREADS;
            if (i < PW * PH) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    sT[c * PW * PH + i] = vt[q][c];
                    sP[c * PW * PH + i] = HASMASK ? blend_pred(vr[q][c], vt[q][c], vm[q]) : vr[q][c];
                    if (AUTO) sS[c * PW * PH + i] = vs[q][c];
                }
            }
        }
    }
    __syncthreads();

    // ---- stage 2: window centres of the tile and of the ring around it.  One thread walks a column of 5 centres
    // (half of the 10 centre rows) with rolling row sums: 6 shared loads per window row instead of 18 per window.
    float ph_acc = 0.0f;
    if (threadIdx.x < 2 * CW) {
        constexpr int NJ = CH / 2;  // centres per thread
        const int half = threadIdx.x / CW, cxl = threadIdx.x - half * CW;
        const int idx0 = half * NJ;  // first centre row of this thread; its window starts at padded row idx0
        const int qx = tx0 + cxl - 1;
        const bool in_x = (qx >= 0) && (qx < W);
        float ss[NJ], l1[NJ], ssa[NJ], l1a[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) ss[j] = l1[j] = ssa[j] = l1a[j] = 0.0f;
        const float k9 = 1.0f / 9.0f, ks = 0.85f * k3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* X = sP + c * PW * PH + cxl;
            const float* Y = sT + c * PW * PH + cxl;
            const float* S = sS + (AUTO ? c * PW * PH + cxl : 0);
            // row sums of the two previous padded rows: {sum x, sum y, sum xx, sum yy, sum xy, sum s, sum ss, sum sy}
            float r0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, r1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            float pa = 0, pb = 0, ps = 0;  // centre-column values of the previous row
#pragma unroll
            for (int k = 0; k < NJ + 2; ++k) {
                const int ly = idx0 + k;
                const float a0 = X[ly * PW], a1 = X[ly * PW + 1], a2 = X[ly * PW + 2];
                const float b0 = Y[ly * PW], b1 = Y[ly * PW + 1], b2 = Y[ly * PW + 2];
                float rs[8];
                rs[0] = a0 + a1 + a2, rs[1] = b0 + b1 + b2;
                rs[2] = fmaf(a0, a0, fmaf(a1, a1, a2 * a2)), rs[3] = fmaf(b0, b0, fmaf(b1, b1, b2 * b2));
                rs[4] = fmaf(a0, b0, fmaf(a1, b1, a2 * b2));
                float s1 = 0;
                if (AUTO) {
                    const float s0 = S[ly * PW], s2 = S[ly * PW + 2];
                    s1 = S[ly * PW + 1];
                    rs[5] = s0 + s1 + s2, rs[6] = fmaf(s0, s0, fmaf(s1, s1, s2 * s2)), rs[7] = fmaf(s0, b0, fmaf(s1, b1, s2 * b2));
                }
                if (k >= 2) {
                    const int j = k - 2, ci = (idx0 + j) * CW + cxl;
                    const int qy = ty0 + idx0 + j - 1;
                    const bool inside = in_x && (qy >= 0) && (qy < H);
                    WinStats w;
                    w.mx = (r0[0] + r1[0] + rs[0]) * k9, w.my = (r0[1] + r1[1] + rs[1]) * k9;
                    w.sx = (r0[2] + r1[2] + rs[2]) * k9 - w.mx * w.mx;
                    w.sy = (r0[3] + r1[3] + rs[3]) * k9 - w.my * w.my;
                    w.sxy = (r0[4] + r1[4] + rs[4]) * k9 - w.mx * w.my;
                    float n, d;
                    const float v = ssim_val(w, n, d);
                    ss[j] += fminf(fmaxf(v, 0.0f), 1.0f);
                    l1[j] += fabsf(pa - pb);
                    if (WANT_G) {
                        float ca = 0, cb = 0, cc = 0;
                        if (inside && v >= 0.0f && v <= 1.0f) {  // clamp backward
                            const float a1c = 2.0f * w.mx * w.my + kC1, a2c = 2.0f * w.sxy + kC2;
                            const float b1c = w.mx * w.mx + w.my * w.my + kC1, b2c = w.sx + w.sy + kC2;
                            const float id = 1.0f / d, nd2 = n * id * id;
                            cb = nd2 * b1c * (k9 * ks);
                            cc = -a1c * id * (k9 * ks);
                            ca = (k9 * ks) * ((a1c - a2c) * w.my * id + nd2 * (b2c - b1c) * w.mx);
                        }
                        cA[c * CW * CH + ci] = ca, cB[c * CW * CH + ci] = cb, cC[c * CW * CH + ci] = cc;
                    }
                    if (AUTO) {
                        WinStats u;
                        u.mx = (r0[5] + r1[5] + rs[5]) * k9, u.my = w.my;
                        u.sx = (r0[6] + r1[6] + rs[6]) * k9 - u.mx * u.mx;
                        u.sy = w.sy;
                        u.sxy = (r0[7] + r1[7] + rs[7]) * k9 - u.mx * u.my;
                        float na, da;
                        ssa[j] += fminf(fmaxf(ssim_val(u, na, da), 0.0f), 1.0f);
                        l1a[j] += fabsf(ps - pb);
                    }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) r0[q] = r1[q], r1[q] = rs[q];
                pa = a1, pb = b1, ps = s1;
            }
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int cyl = idx0 + j, ci = cyl * CW + cxl;
            const int qy = ty0 + cyl - 1;
            const bool inside = in_x && (qy >= 0) && (qy < H);
            float ph = 0.85f * (ss[j] * k3) + 0.15f * (l1[j] * k3);
            float g = inside ? 1.0f : 0.0f;
            if (AUTO) {
                const float pa2 = 0.85f * (ssa[j] * k3) + 0.15f * (l1a[j] * k3);
                if (!(ph <= pa2)) g = 0.0f;  // min() routes the gradient to the first minimum
                ph = fminf(ph, pa2);
            }
            if (inside && cyl >= 1 && cyl <= LT_H && cxl >= 1 && cxl <= LT_W) {  // a pixel of this tile
                ph_acc += ph;
                if (p.out.ph_map) p.out.ph_map[(int64_t)b * p.hw + (int64_t)qy * W + qx] = ph;
            }
            if (WANT_G) {
                gate[ci] = g;
                if (AUTO && g == 0.0f) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) cA[c * CW * CH + ci] = cB[c * CW * CH + ci] = cC[c * CW * CH + ci] = 0.0f;
                }
            }
        }
    }
    if (WANT_G) {
        __syncthreads();
        // ---- stage 3: unit gradient of every tile pixel
        for (int i = threadIdx.x; i < LT_W * LT_H; i += LT_THREADS) {
            const int ly = i / LT_W, lx = i - ly * LT_W;
            const int y = ty0 + ly, x = tx0 + lx;
            if (y >= H || x >= W) continue;
            const int64_t o = (int64_t)y * W + x;
            const float m = HASMASK ? __ldg(p.in.mask_novel + (int64_t)b * p.hw + o) : 1.0f;
            const float gsel = gate[(ly + 1) * CW + lx + 1];
            // multiplicities of the reflected copies of row y / column x inside the neighbouring windows
            float wy[3], wx[3];
#pragma unroll
            for (int d = -1; d <= 1; ++d) {
                const int qy = y + d, qx = x + d;
                wy[d + 1] = (qy < 0 || qy >= H) ? 0.0f : 1.0f + ((y == 1 && qy == 0) ? 1.0f : 0.0f) + ((y == H - 2 && qy == H - 1) ? 1.0f : 0.0f);
                wx[d + 1] = (qx < 0 || qx >= W) ? 0.0f : 1.0f + ((x == 1 && qx == 0) ? 1.0f : 0.0f) + ((x == W - 2 && qx == W - 1) ? 1.0f : 0.0f);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float pr = sP[c * PW * PH + (ly + 2) * PW + lx + 2], tg = sT[c * PW * PH + (ly + 2) * PW + lx + 2];
                float sa = 0, sb = 0, sc = 0;
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                    for (int dx = -1; dx <= 1; ++dx) {
                        const int ci = c * CW * CH + (ly + 1 + dy) * CW + (lx + 1 + dx);
                        const float wgt = wy[dy + 1] * wx[dx + 1];
                        sa = fmaf(wgt, cA[ci], sa), sb = fmaf(wgt, cB[ci], sb), sc = fmaf(wgt, cC[ci], sc);
                    }
                float g = gsel * (0.15f * k3) * sgnf(pr - tg) + sa + sb * pr + sc * tg;
                if (HASMASK) g *= m;
                p.out.g_unit[((int64_t)b * 3 + c) * p.hw + o] = g;
            }
        }
    }
    if (HASMASK) {
        // blended prediction for the consumers outside the path (perceptual term, logging)
        for (int i = threadIdx.x; i < LT_W * LT_H; i += LT_THREADS) {
            const int ly = i / LT_W, lx = i - ly * LT_W;
            const int y = ty0 + ly, x = tx0 + lx;
            if (y >= H || x >= W) continue;
#pragma unroll
            for (int c = 0; c < 3; ++c) p.out.pred[((int64_t)b * 3 + c) * p.hw + (int64_t)y * W + x] = sP[c * PW * PH + (ly + 2) * PW + lx + 2];
        }
    }
    const float tot = block_sum(ph_acc, red);
    if (threadIdx.x == 0) p.partials[((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = tot;
}

// ------------------------------------------------------------------------------------------------
// SSIM + L1 forward (with unit gradient), streamed: no shared memory, no CTA barrier.
//
// A warp owns a strip of 28 output columns (lanes 2..29; lanes 0,1,30,31 carry the halo) and walks a segment of
// rows top to bottom.  Per row every lane loads one value per tensor and channel (coalesced), the horizontal 3-sums
// of {x, y, xx, yy, xy} come from two warp shuffles per value, the vertical 3-sums from two rows of registers.  The
// centre finished at row y is (y-1); its derivative coefficients (A, B, C) are summed 3-wide with the reflect-padding
// multiplicities the same way (shuffles across, registers down), which completes the unit gradient of pixel row
// (y-2):  g = gate * 0.05 * sgn(pred - tgt) + SA + SB * pred + SC * tgt.
// ------------------------------------------------------------------------------------------------
constexpr int SW_COLS = 28, SW_WARPS = 8, SW_THREADS = SW_WARPS * 32;

struct SsimRow {  // one row of loaded values of a lane (prediction already blended with mask_novel)
    float a[3], b[3], s[3], m;
};

template <bool AUTO, bool HASMASK>
__device__ __forceinline__ void ssim_load_row(const LossParams& p, int b, int y, int xv, SsimRow& r) {
    const int H = p.d.H;
    const int yv = reflect(min(max(y, -1), H), H);
    const int64_t o = (int64_t)yv * p.d.W + xv;
    r.m = HASMASK ? __ldg(p.in.mask_novel + (int64_t)b * p.hw + o) : 1.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int64_t oc = ((int64_t)b * 3 + c) * p.hw + o;
        r.b[c] = __ldg(p.in.tgt + oc);
        const float rec = __ldg(p.in.rgb_rec + oc);
        r.a[c] = HASMASK ? blend_pred(rec, r.b[c], r.m) : rec;
        r.s[c] = AUTO ? __ldg(p.in.src + oc) : 0.0f;
    }
}

template <bool AUTO, bool HASMASK, bool WANT_G>
__global__ void __launch_bounds__(SW_THREADS, 2) ssim_l1_stream_kernel(const LossParams p, int strips, int segs, int rs) {
    __shared__ float red[SW_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int H = p.d.H, W = p.d.W;
    const int64_t task = (int64_t)blockIdx.x * SW_WARPS + wid;
    const int64_t per_img = (int64_t)segs * strips;
    const float k3 = 1.0f / 3.0f, k9 = 1.0f / 9.0f, ks = 0.85f * k3, kg = k9 * ks;
    float ph_acc = 0.0f;
    {
        // Control flow is kept warp-uniform by construction (and visibly so: the trip count is a kernel parameter), so the
        // shuffles need no reconvergence protocol: a warp past the last task repeats it with every store masked off.
        const int64_t ntask = (int64_t)p.d.B * per_img;
        const bool live = task < ntask;
        const int64_t tk = live ? task : ntask - 1;
        const int b = (int)(tk / per_img);
        const int rem = (int)(tk - (int64_t)b * per_img);
        const int seg = rem / strips, strip = rem - seg * strips;
        const int ys = seg * rs, ye = min(H, ys + rs);
        const int x = strip * SW_COLS - 2 + lane;
        const int xv = reflect(min(max(x, -1), W), W);
        const bool cx_ok = (x >= 0) && (x < W);
        const bool out_lane = live && (lane >= 2) && (lane < 2 + SW_COLS) && (x < W);
        // multiplicities of column x inside the windows of the centres x-1, x, x+1 (reflect padding)
        float wxm = (x - 1 < 0) ? 0.0f : 1.0f + ((x == W - 2 && x - 1 == W - 1) ? 1.0f : 0.0f);
        float wxp = (x + 1 >= W) ? 0.0f : 1.0f + ((x == W - 2) ? 1.0f : 0.0f);
        if (x == 1) wxm += 1.0f;  // centre 0 sees column 1 a second time as the reflected column -1

        // Two-row histories live in ping-pong slots: at row i slot [i & 1] holds the values of two rows back (it is consumed
        // and then overwritten by the current row), slot [(i + 1) & 1] the previous row.  The loop body is instantiated for
        // both parities, so the slot indices are compile-time constants and no register is moved between rows.
        float Hh[2][3][5], Sh[2][3][3];     // horizontal sums of the two previous rows (prediction / target; source)
        float av[2][3], bv[2][3], sv[2][3]; // values of the two previous rows
        float Ch[2][3][3];                  // column-summed derivative coefficients of the two previous centre rows
        float gatev[2] = {0.0f, 0.0f}, mv[2] = {1.0f, 1.0f};
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
#pragma unroll
                for (int k = 0; k < 5; ++k) Hh[q][c][k] = 0.0f;
#pragma unroll
                for (int k = 0; k < 3; ++k) Sh[q][c][k] = Ch[q][c][k] = 0.0f;
                av[q][c] = bv[q][c] = sv[q][c] = 0.0f;
            }
        SsimRow nxt;
        ssim_load_row<AUTO, HASMASK>(p, b, ys - 2, xv, nxt);
        const int iters = rs + 4;  // rows past ye (last segment of an image) are walked with clamped loads and masked stores

        auto row_step = [&](const int i, auto parity) {
            constexpr int O = decltype(parity)::value;  // slot of two rows back (overwritten at the end), 1 - O: previous row
            constexpr int P = 1 - O;
            const int y = ys - 2 + i;  // value row of this iteration; centre row y-1; gradient row y-2
            const SsimRow cur = nxt;
            if (i + 1 < iters) ssim_load_row<AUTO, HASMASK>(p, b, y + 1, xv, nxt);
            if (HASMASK && out_lane && y >= ys && y < ye) {
#pragma unroll
                for (int c = 0; c < 3; ++c) p.out.pred[((int64_t)b * 3 + c) * p.hw + (int64_t)y * W + x] = cur.a[c];
            }
            const int cy = y - 1;
            const bool inside = cx_ok && (cy >= 0) && (cy < H);
            float ss = 0.0f, l1 = 0.0f, ssa = 0.0f, l1a = 0.0f;
            float ca[3], cb[3], cc[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float a = cur.a[c], t = cur.b[c];
                // neighbours' values come from 4 shuffles; their squares / products are recomputed locally
                const float al = __shfl_up_sync(0xffffffffu, a, 1), ar = __shfl_down_sync(0xffffffffu, a, 1);
                const float tl = __shfl_up_sync(0xffffffffu, t, 1), tr = __shfl_down_sync(0xffffffffu, t, 1);
                float hc[5];
                hc[0] = al + a + ar, hc[1] = tl + t + tr;
                hc[2] = fmaf(al, al, fmaf(ar, ar, a * a)), hc[3] = fmaf(tl, tl, fmaf(tr, tr, t * t));
                hc[4] = fmaf(al, tl, fmaf(ar, tr, a * t));
                WinStats w;
                w.mx = (Hh[O][c][0] + Hh[P][c][0] + hc[0]) * k9, w.my = (Hh[O][c][1] + Hh[P][c][1] + hc[1]) * k9;
                w.sx = (Hh[O][c][2] + Hh[P][c][2] + hc[2]) * k9 - w.mx * w.mx;
                w.sy = (Hh[O][c][3] + Hh[P][c][3] + hc[3]) * k9 - w.my * w.my;
                w.sxy = (Hh[O][c][4] + Hh[P][c][4] + hc[4]) * k9 - w.mx * w.my;
                const float a1c = 2.0f * w.mx * w.my + kC1, a2c = 2.0f * w.sxy + kC2;
                const float b1c = w.mx * w.mx + w.my * w.my + kC1, b2c = w.sx + w.sy + kC2;
                const float n = a1c * a2c, d = b1c * b2c;
                // the value feeds comparisons (clamp, automask min): IEEE division as in layers.py:306, so that knife-edge
                // decisions fall the way the reference's do; the smooth derivative terms use the fast reciprocal
                float id;
                const float v = (1.0f - quotient_rn(n, d, id)) * 0.5f;
                ss += fminf(fmaxf(v, 0.0f), 1.0f);
                l1 += fabsf(av[P][c] - bv[P][c]);
                ca[c] = cb[c] = cc[c] = 0.0f;
                if (WANT_G && v >= 0.0f && v <= 1.0f) {  // clamp backward
                    const float nd2 = n * id * id;
                    cb[c] = nd2 * b1c * kg;
                    cc[c] = -a1c * id * kg;
                    ca[c] = kg * ((a1c - a2c) * w.my * id + nd2 * (b2c - b1c) * w.mx);
                }
                if (AUTO) {
                    const float s = cur.s[c];
                    const float sl = __shfl_up_sync(0xffffffffu, s, 1), sr = __shfl_down_sync(0xffffffffu, s, 1);
                    float hs[3];
                    hs[0] = sl + s + sr, hs[1] = fmaf(sl, sl, fmaf(sr, sr, s * s)), hs[2] = fmaf(sl, tl, fmaf(sr, tr, s * t));
                    WinStats u;
                    u.mx = (Sh[O][c][0] + Sh[P][c][0] + hs[0]) * k9, u.my = w.my;
                    u.sx = (Sh[O][c][1] + Sh[P][c][1] + hs[1]) * k9 - u.mx * u.mx;
                    u.sy = w.sy;
                    u.sxy = (Sh[O][c][2] + Sh[P][c][2] + hs[2]) * k9 - u.mx * u.my;
                    const float na = (2.0f * u.mx * u.my + kC1) * (2.0f * u.sxy + kC2);
                    const float da = (u.mx * u.mx + u.my * u.my + kC1) * (u.sx + u.sy + kC2);
                    float ida;
                    ssa += fminf(fmaxf((1.0f - quotient_rn(na, da, ida)) * 0.5f, 0.0f), 1.0f);
                    l1a += fabsf(sv[P][c] - bv[P][c]);
#pragma unroll
                    for (int k = 0; k < 3; ++k) Sh[O][c][k] = hs[k];
                }
#pragma unroll
                for (int k = 0; k < 5; ++k) Hh[O][c][k] = hc[k];
            }
            float ph = 0.85f * (ss * k3) + 0.15f * (l1 * k3);
            float gate = inside ? 1.0f : 0.0f;
            if (AUTO) {
                const float pa2 = 0.85f * (ssa * k3) + 0.15f * (l1a * k3);
                if (!(ph <= pa2)) gate = 0.0f;  // min() routes the gradient to the first minimum
                ph = fminf(ph, pa2);
            }
            if (inside && out_lane && cy >= ys && cy < ye) {
                ph_acc += ph;
                if (p.out.ph_map) p.out.ph_map[(int64_t)b * p.hw + (int64_t)cy * W + x] = ph;
            }
            if (WANT_G) {
                const int py = y - 2;
                // multiplicities of row py inside the windows of the centre rows py-1, py, py+1
                float wym = (py - 1 < 0) ? 0.0f : 1.0f;
                float wyp = (py + 1 >= H) ? 0.0f : 1.0f + ((py == H - 2) ? 1.0f : 0.0f);
                if (py == 1) wym += 1.0f;
                const bool store = out_lane && py >= ys && py < ye;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float h[3];
                    const float za = ca[c] * gate, zb = cb[c] * gate, zc = cc[c] * gate;  // gate is 0 / 1
                    h[0] = fmaf(wxm, __shfl_up_sync(0xffffffffu, za, 1), fmaf(wxp, __shfl_down_sync(0xffffffffu, za, 1), za));
                    h[1] = fmaf(wxm, __shfl_up_sync(0xffffffffu, zb, 1), fmaf(wxp, __shfl_down_sync(0xffffffffu, zb, 1), zb));
                    h[2] = fmaf(wxm, __shfl_up_sync(0xffffffffu, zc, 1), fmaf(wxp, __shfl_down_sync(0xffffffffu, zc, 1), zc));
                    if (store) {
                        // Ch[O]: centre row py-1 (two back), Ch[P]: centre row py, h: centre row py+1
                        const float sa = fmaf(wym, Ch[O][c][0], fmaf(wyp, h[0], Ch[P][c][0]));
                        const float sb = fmaf(wym, Ch[O][c][1], fmaf(wyp, h[1], Ch[P][c][1]));
                        const float sc = fmaf(wym, Ch[O][c][2], fmaf(wyp, h[2], Ch[P][c][2]));
                        const float pr = av[O][c], tg = bv[O][c];  // values of row py (two back)
                        float g = gatev[P] * (0.15f * k3) * sgnf(pr - tg) + sa + sb * pr + sc * tg;
                        if (HASMASK) g *= mv[O];
                        p.out.g_unit[((int64_t)b * 3 + c) * p.hw + (int64_t)py * W + x] = g;
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) Ch[O][c][k] = h[k];
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) av[O][c] = cur.a[c], bv[O][c] = cur.b[c], sv[O][c] = cur.s[c];
            mv[O] = cur.m;
            gatev[O] = gate;
        };
        for (int i = 0; i < iters; i += 2) {
            row_step(i, std::integral_constant<int, 0>{});
            if (i + 1 < iters) row_step(i + 1, std::integral_constant<int, 1>{});
        }
    }
    ph_acc = warp_sum(ph_acc);
    if (lane == 0) red[wid] = ph_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
#pragma unroll
        for (int i = 0; i < SW_WARPS; ++i) t += red[i];
        p.partials[blockIdx.x] = t;
    }
}

// rows per task: the fewest (waves x rows walked per task) over the whole grid of resident warps
inline int ssim_stream_rows(int B, int H, int W) {
    const int strips = (W + SW_COLS - 1) / SW_COLS;
    const long long resident = 148ll * 2 * SW_WARPS;
    int best = 16;
    long long best_cost = -1;
    const int cand[] = {8, 12, 16, 24, 32, 48, 64, 96, 128};
    for (int rs : cand) {
        const long long tasks = (long long)B * strips * ((H + rs - 1) / rs);
        const long long waves = (tasks + resident - 1) / resident;
        const long long cost = waves * (rs + 4);
        if (best_cost < 0 || cost < best_cost) best_cost = cost, best = rs;
    }
    return best;
}

// ------------------------------------------------------------------------------------------------
// L1 / mixture forward (elementwise), with unit gradients
// ------------------------------------------------------------------------------------------------
constexpr int EW_THREADS = 256;

template <int MODE, bool AUTO, bool HASMASK, bool WANT_G>
__global__ void __launch_bounds__(EW_THREADS) elementwise_fwd_kernel(const LossParams p) {
    __shared__ float red[EW_THREADS / 32];
    const int64_t total = (int64_t)p.d.B * p.hw;
    float acc = 0.0f;
    for (int64_t pix = (int64_t)blockIdx.x * EW_THREADS + threadIdx.x; pix < total; pix += (int64_t)gridDim.x * EW_THREADS) {
        const int b = (int)(pix / p.hw);
        const int64_t o = pix - (int64_t)b * p.hw;
        const float m = HASMASK ? __ldg(p.in.mask_novel + pix) : 1.0f;
        float ph;
        if (MODE == PD_LOSS_MIXTURE) {
            ph = __ldg(p.in.nll + pix);
            float g = 1.0f;
            if (AUTO) {
                const float pa = __ldg(p.in.nll_auto + pix);
                if (!(ph <= pa)) g = 0.0f;
                ph = fminf(ph, pa);
            }
            if (HASMASK) ph *= m, g *= m;
            if (WANT_G) p.out.g_unit_nll[pix] = g;
            if (HASMASK) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int64_t oc = ((int64_t)b * 3 + c) * p.hw + o;
                    p.out.pred[oc] = blend_pred(__ldg(p.in.rgb_rec + oc), __ldg(p.in.tgt + oc), m);
                }
            }
        } else {
            float l1 = 0, l1a = 0, sg[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int64_t oc = ((int64_t)b * 3 + c) * p.hw + o;
                const float t = __ldg(p.in.tgt + oc), r = __ldg(p.in.rgb_rec + oc);
                const float pr = HASMASK ? blend_pred(r, t, m) : r;
                if (HASMASK) p.out.pred[oc] = pr;
                l1 += fabsf(pr - t);
                sg[c] = sgnf(pr - t);
                if (AUTO) l1a += fabsf(__ldg(p.in.src + oc) - t);
            }
            const float k3 = 1.0f / 3.0f;
            ph = l1 * k3;
            float g = k3;
            if (AUTO) {
                const float pa = l1a * k3;
                if (!(ph <= pa)) g = 0.0f;
                ph = fminf(ph, pa);
            }
            if (WANT_G) {
                if (HASMASK) g *= m;
#pragma unroll
                for (int c = 0; c < 3; ++c) p.out.g_unit[((int64_t)b * 3 + c) * p.hw + o] = g * sg[c];
            }
        }
        if (p.out.ph_map) p.out.ph_map[pix] = ph;
        acc += ph;
    }
    const float tot = block_sum(acc, red);
    if (threadIdx.x == 0) p.partials[blockIdx.x] = tot;
}

// deterministic second stage: one CTA sums the per-CTA partials in a fixed order
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const float* __restrict__ partials, int64_t n, float* __restrict__ out, float scale) {
    __shared__ float red[32];
    float acc = 0.0f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += partials[i];
    float t = block_sum(acc, red);
    if (threadIdx.x == 0) *out = t * scale;
}

// ------------------------------------------------------------------------------------------------
// backward: g_rgb_rec = g_ph_sum * unit + g_pred * m        (streaming, 4 pixels per thread when aligned)
//           g_nll     = g_ph_sum * unit_nll                 (mixture)
// ------------------------------------------------------------------------------------------------
template <bool MIXTURE, bool HASMASK>
__global__ void __launch_bounds__(EW_THREADS) photometric_bwd_kernel(const LossParams p) {
    const float gph = __ldg(p.gout.g_ph_sum) * (p.d.out_scale != 0.0f ? p.d.out_scale : 1.0f);
    const int64_t total = (int64_t)p.d.B * p.hw;
    for (int64_t pix = (int64_t)blockIdx.x * EW_THREADS + threadIdx.x; pix < total; pix += (int64_t)gridDim.x * EW_THREADS) {
        const int b = (int)(pix / p.hw);
        const int64_t o = pix - (int64_t)b * p.hw;
        const float m = HASMASK ? __ldg(p.in.mask_novel + pix) : 1.0f;
        if (MIXTURE) p.gin.g_nll[pix] = gph * __ldg(p.out.g_unit_nll + pix);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int64_t oc = ((int64_t)b * 3 + c) * p.hw + o;
            float g = MIXTURE ? 0.0f : gph * __ldg(p.out.g_unit + oc);
            if (p.gout.g_pred) {
                const float ge = __ldg(p.gout.g_pred + oc);
                g += HASMASK ? ge * m : ge;
            }
            p.gin.g_rgb_rec[oc] = g;
        }
    }
}

// same, 4 consecutive pixels per thread (H*W % 4 == 0 and 16-byte aligned pointers)
template <bool MIXTURE, bool HASMASK>
__global__ void __launch_bounds__(EW_THREADS) photometric_bwd_kernel_v4(const LossParams p) {
    const float gph = __ldg(p.gout.g_ph_sum) * (p.d.out_scale != 0.0f ? p.d.out_scale : 1.0f);
    const int64_t hw4 = p.hw / 4;
    for (int64_t q = (int64_t)blockIdx.x * EW_THREADS + threadIdx.x; q < p.total4; q += (int64_t)gridDim.x * EW_THREADS) {
        const int b = (int)(q / hw4);
        const int64_t o4 = q - (int64_t)b * hw4;
        float4 m = make_float4(1.f, 1.f, 1.f, 1.f);
        if (HASMASK) m = __ldg(reinterpret_cast<const float4*>(p.in.mask_novel) + q);
        if (MIXTURE) {
            const float4 u = __ldg(reinterpret_cast<const float4*>(p.out.g_unit_nll) + q);
            reinterpret_cast<float4*>(p.gin.g_nll)[q] = make_float4(gph * u.x, gph * u.y, gph * u.z, gph * u.w);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int64_t oc4 = ((int64_t)b * 3 + c) * hw4 + o4;
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!MIXTURE) {
                const float4 u = __ldg(reinterpret_cast<const float4*>(p.out.g_unit) + oc4);
                g = make_float4(gph * u.x, gph * u.y, gph * u.z, gph * u.w);
            }
            if (p.gout.g_pred) {
                const float4 ge = __ldg(reinterpret_cast<const float4*>(p.gout.g_pred) + oc4);
                g.x = fmaf(ge.x, m.x, g.x), g.y = fmaf(ge.y, m.y, g.y), g.z = fmaf(ge.z, m.z, g.z), g.w = fmaf(ge.w, m.w, g.w);
            }
            reinterpret_cast<float4*>(p.gin.g_rgb_rec)[oc4] = g;
        }
    }
}

}  // namespace pd

// ------------------------------------------------------------------------------------------------
// Edge-aware first-order smoothness of the composited disparity (layers.py:243-256), on the crop
// [..., x0:] that compute_losses hands over (trainer.py:768-771):
//   loss = mean_{x pairs} |d[x]-d[x+1]| exp(-gamma mean_c|I[x]-I[x+1]|) + mean_{y pairs} (same along y)
// One pass forward (two deterministic sums), one pass backward (gradient w.r.t. the disparity only; the
// image is data).  Replaces ~12 elementwise launches forward and ~20 backward on [B,1,H,W] tensors.
// ------------------------------------------------------------------------------------------------
namespace pd {

struct SmoothParams {
    int B, H, W, x0;
    float gamma;
    const float* disp;  // [B,1,H,W]
    const float* img;   // [B,3,H,W]
    float* partials;    // [2][gridDim.x]
    float* out;         // [1]
    const float* g_loss;
    float* g_disp;      // [B,1,H,W]
    int64_t hw;
};

// weight of the pixel pair (o, o + step): exp(-gamma * mean_c |I[o] - I[o + step]|)
__device__ __forceinline__ float smooth_weight(const float* __restrict__ im, int64_t hw, int64_t o, int64_t step, float gamma) {
    const float g = (fabsf(__ldg(im + o) - __ldg(im + o + step)) + fabsf(__ldg(im + hw + o) - __ldg(im + hw + o + step)) +
                     fabsf(__ldg(im + 2 * hw + o) - __ldg(im + 2 * hw + o + step))) * (1.0f / 3.0f);
    return fast_exp(-gamma * g);
}

__global__ void __launch_bounds__(EW_THREADS) smooth_fwd_kernel(const SmoothParams p) {
    __shared__ float red[EW_THREADS / 32];
    const int Wc = p.W - p.x0;
    const int64_t total = (int64_t)p.B * p.H * Wc;
    float ax = 0.0f, ay = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * EW_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * EW_THREADS) {
        const int b = (int)(i / ((int64_t)p.H * Wc));
        const int r = (int)(i - (int64_t)b * p.H * Wc);
        const int y = r / Wc, x = p.x0 + (r - y * Wc);
        const int64_t o = (int64_t)y * p.W + x;
        const float* d = p.disp + (int64_t)b * p.hw;
        const float* im = p.img + (int64_t)b * 3 * p.hw;
        const float dc = __ldg(d + o);
        if (x + 1 < p.W) ax += fabsf(dc - __ldg(d + o + 1)) * smooth_weight(im, p.hw, o, 1, p.gamma);
        if (y + 1 < p.H) ay += fabsf(dc - __ldg(d + o + p.W)) * smooth_weight(im, p.hw, o, p.W, p.gamma);
    }
    float t = block_sum(ax, red);
    if (threadIdx.x == 0) p.partials[blockIdx.x] = t;
    __syncthreads();
    t = block_sum(ay, red);
    if (threadIdx.x == 0) p.partials[gridDim.x + blockIdx.x] = t;
}

__global__ void __launch_bounds__(1024) smooth_reduce_kernel(const float* __restrict__ partials, int n, float inv_nx, float inv_ny, float* __restrict__ out) {
    __shared__ float red[32];
    float ax = 0.0f, ay = 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) ax += partials[i], ay += partials[n + i];
    const float tx = block_sum(ax, red);
    __syncthreads();
    const float ty = block_sum(ay, red);
    if (threadIdx.x == 0) *out = tx * inv_nx + ty * inv_ny;  // grad_disp_x.mean() + grad_disp_y.mean()
}

__global__ void __launch_bounds__(EW_THREADS) smooth_bwd_kernel(const SmoothParams p, float inv_nx, float inv_ny) {
    const float gl = __ldg(p.g_loss);
    const float kx = gl * inv_nx, ky = gl * inv_ny;
    const int64_t total = (int64_t)p.B * p.hw;
    for (int64_t i = (int64_t)blockIdx.x * EW_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * EW_THREADS) {
        const int b = (int)(i / p.hw);
        const int64_t o = i - (int64_t)b * p.hw;
        const int y = (int)(o / p.W), x = (int)(o - (int64_t)y * p.W);
        float g = 0.0f;
        if (x >= p.x0) {
            const float* d = p.disp + (int64_t)b * p.hw;
            const float* im = p.img + (int64_t)b * 3 * p.hw;
            const float dc = __ldg(d + o);
            if (x + 1 < p.W) g += kx * sgnf(dc - __ldg(d + o + 1)) * smooth_weight(im, p.hw, o, 1, p.gamma);
            if (x - 1 >= p.x0) g -= kx * sgnf(__ldg(d + o - 1) - dc) * smooth_weight(im, p.hw, o - 1, 1, p.gamma);
            if (y + 1 < p.H) g += ky * sgnf(dc - __ldg(d + o + p.W)) * smooth_weight(im, p.hw, o, p.W, p.gamma);
            if (y >= 1) g -= ky * sgnf(__ldg(d + o - p.W) - dc) * smooth_weight(im, p.hw, o - p.W, p.W, p.gamma);
        }
        p.g_disp[i] = g;
    }
}

}  // namespace pd
