// Streamed row kernels for stereo disparity warps (trainer.py:540-554) whose disparity does not vary
// along x (vertical planes: one scalar per (image, plane); xz ground planes: one scalar per row —
// depth_decoder.py:153-156, 163-181).  Same job as pd_warp_rows.cuh, re-organised around the B200
// memory system:
//   * warp-specialised CTAs: one producer warp brings every logit / sigma / mask row into a multi-stage
//     shared-memory ring with the TMA engine (cp.async.bulk, 1-D), full / empty mbarriers per stage;
//     consumer warps issue no global loads in the plane loop and never meet a CTA-wide barrier in the
//     forward pass, so they drift apart by up to the ring depth;
//   * the source rgb rows are TMA-staged the same way, double-buffered across row groups;
//   * a plane's warp is "shift by k0 = floor(d) and lerp with frac(d)": the per-(row, plane)
//     coefficients (k0, the two weights, pre-multiplied by the row mask and by log2(e) for the logit)
//     are computed once per row group into shared memory, so a sample costs 2 FMAs per channel;
//   * taps are read as aligned 128-bit shared-memory windows; the softmax over planes is online with a
//     lazily moved reference (one ex2 per sample);
//   * the backward is a gather: per-target dL/dlogit (dL/dsigma) rows are exchanged through a
//     double-buffered shared-memory block and each gradient row is written once with 128-bit streaming
//     stores (no atomics, no zero-fill); one consumer-only named barrier per block of planes.
//
// Coordinates: u = x + sign*d and v = y are used exactly.  The reference evaluates the same numbers
// through an fp32 normalise / un-normalise round trip (trainer.py:549-551 + ATen
// grid_sampler_unnormalize) that perturbs them by a few ulp of the coordinate (<= 6e-5 px at W=640,
// 1.2e-4 px at W=1280); pd_warp_rows.cuh reproduces that perturbation bit for bit and is selected with
// PD_FLAG_EXACT_COORDS.  DESIGN.md (deviations) quantifies the difference.
#pragma once
#include "pd_warp_general.cuh"

namespace pd {
namespace ts {

constexpr int PAD = 12;  // zero floats on both sides of every shared-memory row (>= window size)
constexpr int PADB = 16;  // zero bf16 values on both sides of a raw bf16 row (32 bytes: the TMA destination stays 16-byte aligned)

struct __align__(16) PlaneCoef {
    // first 16 bytes: everything the backward's gather phase needs (one 128-bit broadcast load)
    int k4;          // 4 * k0, k0 = floor(sign * disparity) clamped so that x + k0 cannot overflow: the shift in bytes
    float wc0, wc1;  // (1 - frac) * m, frac * m           (colour / sigma taps, gradient gather)
    float wl0;       // wc0 * log2(e)                       (logit taps, softmax in base 2)
    float wl1;       // wc1 * log2(e)
    float m;         // row mask value (1 when the mask is dense or absent)
    float skip;      // SMASK_SUMMARY (backward): != 0 when the row summary says the mask row is all ones, so it is not read
    int pad;
};

struct StreamCfg {
    int rpc;      // rows per CTA iteration ("row group")
    int tpr;      // threads per row = W / PX
    int pitch;    // floats per shared row = W + 2 * PAD
    int hs;       // planes per pipeline block (= ring stage)
    int nst;      // ring stages
    int nblk;     // blocks per row group = ceil(N / hs)
    int ngroups;  // ceil(B * H / rpc)
    int early;    // 1: the producer stages the NEXT group's coefficients / source rows while this one streams (producer_loop)
    int nc;       // consumer threads (multiple of 32); the CTA has nc + 32 threads, the last warp produces
    int l2_hint;  // streamed rows carry the L2 evict-first policy
    int bf16;     // logits / sigma / their gradients are stored as bf16 (pd_warp_desc.dtype)
};

constexpr int MAX_STAGES = 8;
constexpr int BAR_FULL = 0, BAR_EMPTY = MAX_STAGES, BAR_SRC = 2 * MAX_STAGES, BAR_BYTES = 256;

// SMASK_ROW: no mask or one value per (row, plane), folded into the coefficients.  SMASK_DENSE: fp32 mask rows travel
// through the TMA ring next to the logits.  SMASK_SUMMARY (backward only): dense mask whose rows the forward pass has
// summarised (WarpParams::mask_rows): all-ones / all-zero rows are folded into the coefficients like SMASK_ROW, the
// remaining rows are read from global memory by the consumers; no mask ring.
enum { SMASK_ROW = 0, SMASK_DENSE = 1, SMASK_SUMMARY = 2 };

// ------------------------------------------------------------------------------------------------
// mbarrier / TMA primitives (PTX; SASS: SYNCS / UBLKCP)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* q) { return (uint32_t)__cvta_generic_to_shared(q); }
// suspend-time hint of mbarrier.try_wait: a waiting thread is parked by the hardware until the phase completes (or this many
// ns pass) instead of spinning through try_wait / branch pairs that compete with the working warps for issue slots
constexpr uint32_t kSuspendHintNs = 20000u;

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// consumer-only CTA barrier (hardware barrier 1); the producer warp never joins it
__device__ __forceinline__ void consumer_sync(int nc) { asm volatile("bar.sync 1, %0;" ::"r"(nc) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(kSuspendHintNs)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(kSuspendHintNs)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned)
__device__ __forceinline__ void tma_row(float* dst, const float* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// The [B,N,H,W] streams (logits, sigma, dense mask: 289 MB each at cfg 2) pass through the chip exactly once per kernel while
// the per-pixel context around them (source colour, rgb_rec, statistics, upstream / unit gradients: ~100 MB) is produced by
// one kernel of the step and consumed by the next.  The streamed rows therefore travel with an L2 evict-first policy, which
// leaves the 126 MB L2 to the context.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_row_hint(float* dst, const float* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}

__device__ __forceinline__ float4 lds128(const float* q) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(q)));
    return v;
}

__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

// WF consecutive floats from the 16-byte aligned shared address a
template <int WF>
__device__ __forceinline__ void load_window(uint32_t a, float (&v)[WF]) {
#pragma unroll
    for (int i = 0; i < WF / 4; ++i) {
        float4 t = lds128(a + 16 * i);
        v[4 * i] = t.x, v[4 * i + 1] = t.y, v[4 * i + 2] = t.z, v[4 * i + 3] = t.w;
    }
}

template <int WF>
__device__ __forceinline__ void load_window(const float* q, float (&v)[WF]) {
#pragma unroll
    for (int i = 0; i < WF / 4; ++i) {
        float4 t = lds128(q + 4 * i);
        v[4 * i] = t.x, v[4 * i + 1] = t.y, v[4 * i + 2] = t.z, v[4 * i + 3] = t.w;
    }
}

template <int PX>
__device__ __forceinline__ void load_px_global(const float* q, float (&v)[PX]) {
#pragma unroll
    for (int i = 0; i < PX / 4; ++i) {
        float4 t = __ldg(reinterpret_cast<const float4*>(q) + i);
        v[4 * i] = t.x, v[4 * i + 1] = t.y, v[4 * i + 2] = t.z, v[4 * i + 3] = t.w;
    }
}

template <int PX>
__device__ __forceinline__ void load_upstream(const WarpParams& p, float gph, int64_t i, int64_t pix, float (&v)[PX]) {
#pragma unroll
    for (int k = 0; k < PX / 4; ++k) {
        const float4 t = upstream_rgb4(p, gph, i + 4 * k, pix + 4 * k);
        v[4 * k] = t.x, v[4 * k + 1] = t.y, v[4 * k + 2] = t.z, v[4 * k + 3] = t.w;
    }
}

template <int PX>
__device__ __forceinline__ void store_px_global(float* q, const float (&v)[PX]) {
#pragma unroll
    for (int i = 0; i < PX / 4; ++i) reinterpret_cast<float4*>(q)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

template <int PX>
__device__ __forceinline__ void store_px_stream(float* q, const float (&v)[PX]) {
#pragma unroll
    for (int i = 0; i < PX / 4; ++i) __stcs(reinterpret_cast<float4*>(q) + i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
}

// ------------------------------------------------------------------------------------------------
// shared-memory carve-up, common to both kernels
// ------------------------------------------------------------------------------------------------
struct Smem {
    uint64_t* bars;   // BAR_FULL + stage, BAR_EMPTY + stage, BAR_SRC + {0,1}
    PlaneCoef* coef;  // [2][rpc][N]
    float* src;       // [2][rpc][3][pitch]
    float* lring;     // [nst*hs][rpc][pitch]
    float* sring;     // mixture: [nst*hs][rpc][pitch]
    float* mring;     // dense mask: [nst*hs][rpc][pitch]
    float* dbuf;      // backward: [2][hs][NE][rpc][pitch] exchange rows (NE = 1, mixture 2)
    float* gacc;      // backward with d/d disp: [max(rpc, consumer warps)][N] (one private row per warp when warps do not straddle rows)
    float* fend;      // one past the last float
    unsigned char* bring;  // bf16 storage: TMA ring of raw bf16 rows [nst*hs][1 + mix][rpc][(W + 2 * PADB) * 2 bytes] (behind the float region)
};

__host__ __device__ inline size_t stream_smem_floats(const StreamCfg& c, int N, bool mix, bool dense, int ne_bwd, bool want_disp) {
    size_t rowf = (size_t)c.rpc * c.pitch;
    const size_t streams = c.bf16 ? (dense ? 1 : 0) : (1 + (mix ? 1 : 0) + (dense ? 1 : 0));  // fp32 ring rows
    size_t f = 2 * 3 * rowf + (size_t)c.nst * c.hs * rowf * streams;
    f += (size_t)2 * c.hs * ne_bwd * rowf;
    if (want_disp) f += (size_t)(c.rpc > c.nc / 32 ? c.rpc : c.nc / 32) * N;
    return f;
}

// bytes of one raw bf16 row in the TMA ring, zero pads included (W % 8 == 0: a multiple of 16)
__host__ __device__ inline size_t stream_bf16_row_bytes(const StreamCfg& c) { return (size_t)(c.pitch - 2 * PAD + 2 * PADB) * 2; }

__host__ __device__ inline size_t stream_smem_bytes(const StreamCfg& c, int N, bool mix, bool dense, int ne_bwd, bool want_disp) {
    size_t b = BAR_BYTES + (size_t)2 * c.rpc * N * sizeof(PlaneCoef) + stream_smem_floats(c, N, mix, dense, ne_bwd, want_disp) * sizeof(float) + 16;
    if (c.bf16) b += (size_t)c.nst * c.hs * (1 + (mix ? 1 : 0)) * c.rpc * stream_bf16_row_bytes(c) + 16;
    return b;
}

__device__ __forceinline__ Smem carve(unsigned char* raw, const StreamCfg& c, int N, bool mix, bool dense, int ne_bwd, bool want_disp) {
    Smem s;
    const size_t rowf = (size_t)c.rpc * c.pitch;
    s.bars = reinterpret_cast<uint64_t*>(raw);
    s.coef = reinterpret_cast<PlaneCoef*>(raw + BAR_BYTES);
    s.src = reinterpret_cast<float*>(s.coef + (size_t)2 * c.rpc * N);
    s.lring = s.src + 2 * 3 * rowf;
    float* q = s.lring;
    if (!c.bf16) q += (size_t)c.nst * c.hs * rowf;
    s.sring = q;
    if (mix && !c.bf16) q += (size_t)c.nst * c.hs * rowf;
    s.mring = q;
    if (dense) q += (size_t)c.nst * c.hs * rowf;
    s.dbuf = q;
    q += (size_t)2 * c.hs * ne_bwd * rowf;
    s.gacc = q;
    if (want_disp) q += (size_t)(c.rpc > c.nc / 32 ? c.rpc : c.nc / 32) * N;
    s.fend = q;
    s.bring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(q) + 15) & ~(uintptr_t)15);
    return s;
}

// The zero pads on both sides of every shared row are written once and never touched again (TMA and the
// exchange stores only write the W interior floats); row interiors need no initialisation.
__device__ __forceinline__ void zero_pads(const Smem& s, const StreamCfg& c, int W) {
    const int nrows = (int)((s.gacc - s.src) / c.pitch);
    for (int i = threadIdx.x; i < nrows * 2 * PAD; i += blockDim.x) {
        const int rw = i / (2 * PAD), q = i - rw * 2 * PAD;
        s.src[(size_t)rw * c.pitch + (q < PAD ? q : W + q)] = 0.0f;
    }
}

// same for the raw bf16 ring rows (PADB values on both sides, written as 32-bit words)
__device__ __forceinline__ void zero_pads_bf16(const Smem& s, const StreamCfg& c, int W, int nrows) {
    const size_t rbp = stream_bf16_row_bytes(c);
    constexpr int WORDS = PADB / 2;  // 32-bit words per pad
    for (int i = threadIdx.x; i < nrows * 2 * WORDS; i += blockDim.x) {
        const int rw = i / (2 * WORDS), q = i - rw * 2 * WORDS;
        uint32_t* row = reinterpret_cast<uint32_t*>(s.bring + (size_t)rw * rbp);
        row[q < WORDS ? q : W / 2 + q] = 0u;
    }
}

// per-(row, plane) coefficients of row group g (executed by the 32 lanes of the producer warp)
template <int MASKMODE>
__device__ __forceinline__ void stage_coef(const WarpParams& p, const StreamCfg& c, PlaneCoef* coef, int g, int rows_total) {
    const int N = p.d.N, H = p.d.H, W = p.d.W;
    for (int idx = threadIdx.x & 31; idx < c.rpc * N; idx += 32) {
        const int r = idx / N, n = idx - r * N;
        const int row = g * c.rpc + r;
        PlaneCoef k;
        int k0 = W + 16;
        k.wc0 = k.wc1 = k.wl0 = k.wl1 = 0.0f, k.m = 0.0f, k.skip = 0.0f, k.pad = 0;
        if (row < rows_total) {
            const int b = row / H, y = row - b * H;
            const float d = __ldg(p.in.disp + soff(p.d.disp_stride, b, n, y, 0));
            const float sd = p.d.disp_sign * d;
            const float kf = floorf(sd);
            const bool sane = fabsf(sd) < (float)(W + 8);  // otherwise every tap is out of range
            const float w1 = sane ? sd - kf : 0.0f;
            const float m = (MASKMODE == SMASK_ROW) ? load_mask(p.in.mask, p.d.mask_dtype, soff(p.d.mask_stride, b, n, y, 0)) : 1.0f;
            if (MASKMODE == SMASK_SUMMARY) {
                // set bit n = some pixel of the plane's mask row differs from 1.0
                if (!((__ldg(p.mask_rows + row) >> n) & 1ull)) k.skip = 1.0f;
            }
            k0 = sane ? (int)kf : W + 16;
            k.m = m;
            k.wc1 = w1 * m;
            k.wc0 = (1.0f - w1) * m;
            k.wl0 = k.wc0 * kLog2e;
            k.wl1 = k.wc1 * kLog2e;
        }
        k.k4 = k0 * 4;
        coef[idx] = k;
    }
}

// Producer warp: walks the CTA's (row group, block) sequence; per block it waits for the ring stage to be
// released by every consumer warp, (at a group start) stages the group's coefficients and source rows, then
// arms the stage's full barrier with the byte count and issues one bulk copy per (plane, row, stream).
// Releases that make reuse safe: the stage's empty barrier is armed by the consumers after their last read
// of block jb - nst; with nst <= nblk that also covers the coefficient / source-row buffers of group it - 2.
template <bool MIX, int MASKMODE, bool BF16 = false>
__device__ __forceinline__ void producer_loop(const WarpParams& p, const StreamCfg& c, const Smem& s, int nit) {
    constexpr bool DENSE = (MASKMODE == SMASK_DENSE);
    static_assert(!(BF16 && DENSE), "bf16 storage is built for row masks only");
    const int lane = threadIdx.x & 31;
    const int W = p.d.W, H = p.d.H, N = p.d.N, rows_total = p.d.B * H;
    const uint32_t rowbytes = (uint32_t)(W * sizeof(float));
    const int streams = 1 + (MIX ? 1 : 0) + (DENSE ? 1 : 0);
    const uint64_t pol = l2_evict_first_policy();
    const bool hint = c.l2_hint != 0;
    int stage = 0, use = 0;  // use = how many times the ring wrapped
    // coefficients and source rows of one row group into its (group & 1) buffers
    auto stage_group = [&](int it2) {
        const int g2 = blockIdx.x + it2 * gridDim.x;
        const int r0 = g2 * c.rpc;
        const int nr = min(c.rpc, rows_total - r0);
        stage_coef<MASKMODE>(p, c, s.coef + (size_t)(it2 & 1) * c.rpc * N, g2, rows_total);
        __syncwarp();
        if (lane == 0) {
            uint64_t* bar = s.bars + BAR_SRC + (it2 & 1);
            mbar_expect_tx(bar, (uint32_t)(nr * 3) * rowbytes);
            for (int r = 0; r < nr; ++r) {
                const int row = r0 + r, b = row / H, y = row - b * H;
                const float* gp = p.in.src + ((int64_t)b * 3 * H + y) * W;
                float* sp = s.src + ((size_t)((it2 & 1) * c.rpc + r) * 3) * c.pitch + PAD;
                for (int ch = 0; ch < 3; ++ch) tma_row(sp + (size_t)ch * c.pitch, gp + (int64_t)ch * p.hw, rowbytes, bar);
            }
        }
    };
    // With more blocks than ring stages the NEXT group is staged while this one streams: at block nst the awaited release
    // is of block 0 of this group by every consumer warp, so the previous group (the other user of those buffers and of
    // that barrier) is finished.  Otherwise a group is staged at its own first block.
    const bool early = c.nblk > c.nst && c.early;
    for (int it = 0; it < nit; ++it) {
        const int g = blockIdx.x + it * gridDim.x;
        const int row0 = g * c.rpc;
        const int nrows = min(c.rpc, rows_total - row0);
        for (int j = 0; j < c.nblk; ++j) {
            if (use > 0) mbar_wait(s.bars + BAR_EMPTY + stage, (use - 1) & 1);
            if (j == 0 && (it == 0 || !early)) stage_group(it);
            const int n0 = j * c.hs, n1 = min(N, n0 + c.hs);
            if (BF16) {
                // raw bf16 rows: [stage][plane][logit | sigma][row][PADB | W | PADB]; the consumers' tap windows convert on load
                if (lane == 0) {
                    uint64_t* bar = s.bars + BAR_FULL + stage;
                    const uint32_t rb = rowbytes / 2;
                    constexpr int NEc = MIX ? 2 : 1;
                    mbar_expect_tx(bar, (uint32_t)((n1 - n0) * nrows * NEc) * rb);
                    const unsigned char* lg = reinterpret_cast<const unsigned char*>(p.in.logits);
                    const unsigned char* sg = reinterpret_cast<const unsigned char*>(p.in.sigma);
                    for (int r = 0; r < nrows; ++r) {
                        const int row = row0 + r, b = row / H, y = row - b * H;
                        int64_t off = ((((int64_t)b * N + n0) * H + y) * W) * 2;
                        const size_t rbp = stream_bf16_row_bytes(c);
                        unsigned char* dst = s.bring + ((size_t)(stage * c.hs) * NEc * c.rpc + r) * rbp + PADB * 2;
                        for (int n = n0; n < n1; ++n) {
                            if (hint) tma_row_hint(reinterpret_cast<float*>(dst), reinterpret_cast<const float*>(lg + off), rb, bar, pol);
                            else tma_row(reinterpret_cast<float*>(dst), reinterpret_cast<const float*>(lg + off), rb, bar);
                            if (MIX) {
                                if (hint) tma_row_hint(reinterpret_cast<float*>(dst + (size_t)c.rpc * rbp), reinterpret_cast<const float*>(sg + off), rb, bar, pol);
                                else tma_row(reinterpret_cast<float*>(dst + (size_t)c.rpc * rbp), reinterpret_cast<const float*>(sg + off), rb, bar);
                            }
                            off += p.hw * 2;
                            dst += (size_t)NEc * c.rpc * rbp;
                        }
                    }
                }
            } else if (lane == 0) {
                uint64_t* bar = s.bars + BAR_FULL + stage;
                mbar_expect_tx(bar, (uint32_t)((n1 - n0) * nrows * streams) * rowbytes);
                for (int r = 0; r < nrows; ++r) {
                    const int row = row0 + r, b = row / H, y = row - b * H;
                    int64_t off = (((int64_t)b * N + n0) * H + y) * W;
                    size_t slot = ((size_t)(stage * c.hs) * c.rpc + r) * c.pitch + PAD;
                    for (int n = n0; n < n1; ++n) {
                        if (hint) {
                            tma_row_hint(s.lring + slot, p.in.logits + off, rowbytes, bar, pol);
                            if (MIX) tma_row_hint(s.sring + slot, p.in.sigma + off, rowbytes, bar, pol);
                            if (DENSE) tma_row_hint(s.mring + slot, reinterpret_cast<const float*>(p.in.mask) + soff(p.d.mask_stride, b, n, y, 0), rowbytes, bar, pol);
                        } else {
                            tma_row(s.lring + slot, p.in.logits + off, rowbytes, bar);
                            if (MIX) tma_row(s.sring + slot, p.in.sigma + off, rowbytes, bar);
                            if (DENSE) tma_row(s.mring + slot, reinterpret_cast<const float*>(p.in.mask) + soff(p.d.mask_stride, b, n, y, 0), rowbytes, bar);
                        }
                        off += p.hw;
                        slot += (size_t)c.rpc * c.pitch;
                    }
                }
            }
            // after this block's copies are in flight: the staging costs a dependent global load of the disparities
            if (early && j == c.nst && it + 1 < nit) stage_group(it + 1);
            if (++stage == c.nst) stage = 0, ++use;
        }
    }
}

__device__ __forceinline__ PlaneCoef load_coef(uint32_t a32) {
    const float4 a = lds128(a32);
    const float4 b = lds128(a32 + 16);
    PlaneCoef k;
    k.k4 = __float_as_int(a.x), k.wc0 = a.y, k.wc1 = a.z, k.wl0 = a.w, k.wl1 = b.x, k.m = b.y, k.skip = b.z, k.pad = 0;
    return k;
}
// the gather phase of the backward: shift and the two colour-domain weights only
__device__ __forceinline__ void load_coef_gather(uint32_t a32, int& k4, float& wc0, float& wc1) {
    const float4 a = lds128(a32);
    k4 = __float_as_int(a.x), wc0 = a.y, wc1 = a.z;
}

// byte offset (multiple of 16, clamped into the zero pads) of the window holding the taps that start at byte
// offset at4 of a row
__device__ __forceinline__ int window_off(int at4, int W4) { return min(max(at4 & ~15, -4 * PAD), W4); }

template <int PX>
__device__ __forceinline__ bool all_ones(const float (&m)[PX]) {
    unsigned acc = 0xffffffffu, orr = 0u;
#pragma unroll
    for (int i = 0; i < PX; ++i) acc &= __float_as_uint(m[i]), orr |= __float_as_uint(m[i]);
    return acc == 0x3f800000u && orr == 0x3f800000u;
}

// bf16 storage: a tap window of 8 consecutive bf16 values from the 8-byte aligned shared address a, widened to fp32 in
// registers (bf16 -> fp32 is a 16-bit shift / mask: one ALU instruction per value the window body actually uses)
__device__ __forceinline__ void load_window_bf16(uint32_t a, float (&v)[8]) {
    uint32_t w0, w1, w2, w3;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(w0), "=r"(w1) : "r"(a));
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(w2), "=r"(w3) : "r"(a + 8));
    v[0] = __uint_as_float(w0 << 16), v[1] = __uint_as_float(w0 & 0xffff0000u);
    v[2] = __uint_as_float(w1 << 16), v[3] = __uint_as_float(w1 & 0xffff0000u);
    v[4] = __uint_as_float(w2 << 16), v[5] = __uint_as_float(w2 & 0xffff0000u);
    v[6] = __uint_as_float(w3 << 16), v[7] = __uint_as_float(w3 & 0xffff0000u);
}

// logit / sigma window of the plane body: fp32 rows or raw bf16 rows
template <int WF, bool BF16>
__device__ __forceinline__ void load_plane_window(uint32_t a, float (&v)[WF]) {
    if constexpr (BF16) {
        static_assert(WF == 8, "bf16 storage: 4 pixels per thread");
        load_window_bf16(a, v);
    } else {
        load_window<WF>(a, v);
    }
}

// round-to-nearest-even bf16 pair (x -> low half, y -> high half)
__device__ __forceinline__ uint32_t pack_bf16x2(float x, float y) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(y), "f"(x));
    return r;
}

template <int PX>
__device__ __forceinline__ void store_px_stream_bf16(void* base, int64_t o, const float (&v)[PX]) {
    uint2* q = reinterpret_cast<uint2*>(reinterpret_cast<unsigned short*>(base) + o);
#pragma unroll
    for (int i = 0; i < PX / 4; ++i) __stcs(q + i, make_uint2(pack_bf16x2(v[4 * i], v[4 * i + 1]), pack_bf16x2(v[4 * i + 2], v[4 * i + 3])));
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <bool MIX, int PX>
struct FwdAcc {
    float Ml2[PX], S[PX], R0[PX], R1[PX], R2[PX];
    float A[MIX ? PX : 1], Q[MIX ? PX : 1], Qa[MIX ? PX : 1];
    float tr[MIX ? PX : 1], tg[MIX ? PX : 1], tb[MIX ? PX : 1], ea[MIX ? PX : 1];
};

template <bool MIX, int PX, int R, bool PERPIX, bool BF16 = false>
__device__ __forceinline__ void fwd_plane(uint32_t srow, uint32_t lrow, uint32_t sgrow, uint32_t pitch4, const PlaneCoef& k,
                                          const float (&mm)[PX], FwdAcc<MIX, PX>& a) {
    // srow / lrow / sgrow: shared addresses of the (already window-aligned) first value of the tap windows
    constexpr int WF = PX + 4;
    float v[WF], t[PX];
    load_plane_window<WF, BF16>(lrow, v);
#pragma unroll
    for (int i = 0; i < PX; ++i) {
        if (PERPIX) t[i] = fmaf(fmaf(k.wl0, v[R + i], k.wl1 * v[R + i + 1]), mm[i], -a.Ml2[i]);
        else t[i] = fmaf(k.wl0, v[R + i], fmaf(k.wl1, v[R + i + 1], -a.Ml2[i]));
    }
    float tmx = t[0];
#pragma unroll
    for (int i = 1; i < PX; ++i) tmx = fmaxf(tmx, t[i]);
    if (tmx > 64.0f) {
        // move the softmax reference of the pixels whose logit ran away from it (always on the first plane:
        // the reference starts at -inf); exp2(t) cannot overflow below that threshold
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            if (t[i] > 64.0f) {
                float nl2 = fmaf(k.wl0, v[R + i], k.wl1 * v[R + i + 1]);
                if (PERPIX) nl2 *= mm[i];
                const float sc = fast_exp2(a.Ml2[i] - nl2);
                a.S[i] *= sc, a.R0[i] *= sc, a.R1[i] *= sc, a.R2[i] *= sc;
                if constexpr (MIX) { a.A[i] *= sc, a.Q[i] *= sc, a.Qa[i] *= sc; }
                a.Ml2[i] = nl2;
                t[i] = 0.0f;
            }
        }
    }
    float e[PX];
#pragma unroll
    for (int i = 0; i < PX; ++i) {
        e[i] = fast_exp2(t[i]);
        a.S[i] += e[i];
    }
    if constexpr (!MIX) {
        float a0[PX], a1[PX];
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            const float em = PERPIX ? e[i] * mm[i] : e[i];
            a0[i] = em * k.wc0, a1[i] = em * k.wc1;
        }
        load_window<WF>(srow, v);
#pragma unroll
        for (int i = 0; i < PX; ++i) a.R0[i] = fmaf(a0[i], v[R + i], fmaf(a1[i], v[R + i + 1], a.R0[i]));
        load_window<WF>(srow + pitch4, v);
#pragma unroll
        for (int i = 0; i < PX; ++i) a.R1[i] = fmaf(a0[i], v[R + i], fmaf(a1[i], v[R + i + 1], a.R1[i]));
        load_window<WF>(srow + 2 * pitch4, v);
#pragma unroll
        for (int i = 0; i < PX; ++i) a.R2[i] = fmaf(a0[i], v[R + i], fmaf(a1[i], v[R + i + 1], a.R2[i]));
    } else {
        float es[PX], inv[PX], err[PX];
        load_plane_window<WF, BF16>(sgrow, v);
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            float s = fmaf(k.wc0, v[R + i], k.wc1 * v[R + i + 1]);
            if (PERPIX) s *= mm[i];
            const float sg = fminf(fmaxf(s, 0.01f), 1.0f);  // trainer.py:597
            inv[i] = fast_rcp(sg);
            es[i] = e[i] * inv[i];
            a.A[i] += es[i];
        }
        load_window<WF>(srow, v);
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            float c = fmaf(k.wc0, v[R + i], k.wc1 * v[R + i + 1]);
            if (PERPIX) c *= mm[i];
            a.R0[i] = fmaf(es[i], c, a.R0[i]);
            err[i] = fabsf(c - a.tr[i]);
        }
        load_window<WF>(srow + pitch4, v);
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            float c = fmaf(k.wc0, v[R + i], k.wc1 * v[R + i + 1]);
            if (PERPIX) c *= mm[i];
            a.R1[i] = fmaf(es[i], c, a.R1[i]);
            err[i] += fabsf(c - a.tg[i]);
        }
        load_window<WF>(srow + 2 * pitch4, v);
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            float c = fmaf(k.wc0, v[R + i], k.wc1 * v[R + i + 1]);
            if (PERPIX) c *= mm[i];
            a.R2[i] = fmaf(es[i], c, a.R2[i]);
            err[i] += fabsf(c - a.tb[i]);
        }
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            const float il2 = inv[i] * (kLog2e * (1.0f / 3.0f));  // err holds the channel SUM
            const float hes = 0.5f * es[i];                      // e * 0.5 / sigma
            a.Q[i] = fmaf(hes, fast_exp2(-err[i] * il2), a.Q[i]);  // layers.py:454-455
            a.Qa[i] = fmaf(hes, fast_exp2(-a.ea[i] * il2), a.Qa[i]);
        }
    }
}

// at4 = byte offset of the first tap inside the row; the windows start at the aligned offset below it
template <bool MIX, int PX, bool PERPIX, bool BF16 = false>
__device__ __forceinline__ void fwd_plane_any(uint32_t srow, uint32_t lrow, uint32_t sgrow, uint32_t pitch4, int at4, int W4, const PlaneCoef& k,
                                              const float (&mm)[PX], FwdAcc<MIX, PX>& a) {
    const int wo = window_off(at4, W4);
    const int wl = BF16 ? wo / 2 : wo;  // the logit / sigma rows hold 2-byte values with bf16 storage
    srow += wo, lrow += wl, sgrow += wl;
    if (at4 & 8) {
        if (at4 & 4) fwd_plane<MIX, PX, 3, PERPIX, BF16>(srow, lrow, sgrow, pitch4, k, mm, a);
        else fwd_plane<MIX, PX, 2, PERPIX, BF16>(srow, lrow, sgrow, pitch4, k, mm, a);
    } else {
        if (at4 & 4) fwd_plane<MIX, PX, 1, PERPIX, BF16>(srow, lrow, sgrow, pitch4, k, mm, a);
        else fwd_plane<MIX, PX, 0, PERPIX, BF16>(srow, lrow, sgrow, pitch4, k, mm, a);
    }
}

template <bool MIX, int MASKMODE, int PX, int THREADS, int MINB, bool BF16 = false>
__global__ void __launch_bounds__(THREADS, MINB) rows_fwd_stream(const WarpParams p, const StreamCfg cfg) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr bool DENSE = (MASKMODE == SMASK_DENSE);
    constexpr int NEc = MIX ? 2 : 1;
    const int W = p.d.W, H = p.d.H, N = p.d.N, pitch = cfg.pitch, rpc = cfg.rpc, hs = cfg.hs, NB = cfg.nblk;
    const int rows_total = p.d.B * H;
    const Smem s = carve(smem_raw, cfg, N, MIX, DENSE, 0, false);
    zero_pads(s, cfg, W);
    if constexpr (BF16) zero_pads_bf16(s, cfg, W, cfg.nst * cfg.hs * NEc * cfg.rpc);
    if (threadIdx.x == 0) {
        const uint32_t readers = (uint32_t)(cfg.nc / 32);
        for (int i = 0; i < cfg.nst; ++i) {
            mbar_init(s.bars + BAR_FULL + i, 1);
            mbar_init(s.bars + BAR_EMPTY + i, readers);
        }
        mbar_init(s.bars + BAR_SRC, 1);
        mbar_init(s.bars + BAR_SRC + 1, 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int nit = (cfg.ngroups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if ((int)threadIdx.x >= cfg.nc) {
        producer_loop<MIX, MASKMODE, BF16>(p, cfg, s, nit);
        return;
    }
    const int lane = threadIdx.x & 31;
    const int r = threadIdx.x / cfg.tpr;
    const int x0 = (threadIdx.x - r * cfg.tpr) * PX;
    // 32-bit shared-window addresses of the row interiors (byte units from here on)
    const uint32_t bars = smem_u32(s.bars), coef0 = smem_u32(s.coef), src0 = smem_u32(s.src + PAD), lring0 = smem_u32(s.lring + PAD);
    const uint32_t pitch4 = (uint32_t)pitch * 4u, rowpitch4 = (uint32_t)rpc * pitch4;
    // fp32 storage: logit rows of a stage are rpc rows apart, the sigma ring sits sdelta behind the logit ring.  bf16 storage:
    // the tap windows read the raw rows [stage][plane][logit | sigma][row] (byte pitch rbp, interiors PADB values in)
    const uint32_t rbp = (uint32_t)stream_bf16_row_bytes(cfg);
    const uint32_t sdelta = BF16 ? (uint32_t)rpc * rbp : (uint32_t)((s.sring - s.lring) * sizeof(float));
    const uint32_t mdelta = (uint32_t)((s.mring - s.lring) * sizeof(float));
    const uint32_t plane4 = BF16 ? NEc * (uint32_t)rpc * rbp : rowpitch4;
    const uint32_t bring0 = smem_u32(s.bring) + PADB * 2;
    const int x04 = x0 * 4, W4 = W * 4;
    int stage = 0, jb = 0;
    uint32_t fphase = 0;

    for (int it = 0; it < nit; ++it) {
        const int g = blockIdx.x + it * gridDim.x;
        const int row = g * rpc + r;
        const bool active = (r < rpc) && (row < rows_total);
        const int b = active ? row / H : 0, y = active ? row - b * H : 0;
        const int64_t rem = (int64_t)y * W + x0;

        FwdAcc<MIX, PX> acc;
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            acc.Ml2[i] = -INFINITY, acc.S[i] = acc.R0[i] = acc.R1[i] = acc.R2[i] = 0.0f;
            if constexpr (MIX) acc.A[i] = acc.Q[i] = acc.Qa[i] = acc.ea[i] = 0.0f;
        }
        if constexpr (MIX) if (active) {
            const float* tp = p.in.tgt + (int64_t)b * p.chw3 + rem;
            load_px_global<PX>(tp, acc.tr);
            load_px_global<PX>(tp + p.hw, acc.tg);
            load_px_global<PX>(tp + 2 * p.hw, acc.tb);
            if (p.d.automask) {
                const float* sp = p.in.src + (int64_t)b * p.chw3 + rem;
                float sr[PX], sg[PX], sb[PX];
                load_px_global<PX>(sp, sr);
                load_px_global<PX>(sp + p.hw, sg);
                load_px_global<PX>(sp + 2 * p.hw, sb);
#pragma unroll
                for (int i = 0; i < PX; ++i) acc.ea[i] = fabsf(sr[i] - acc.tr[i]) + fabsf(sg[i] - acc.tg[i]) + fabsf(sb[i] - acc.tb[i]);  // channel SUM
            }
        }
        uint32_t coef_a = coef0 + (uint32_t)(((it & 1) * rpc + r) * N) * (uint32_t)sizeof(PlaneCoef);
        const uint32_t srow = src0 + (uint32_t)(((it & 1) * rpc + r) * 3) * pitch4;
        mbar_wait(bars + 8 * (BAR_SRC + (it & 1)), (it >> 1) & 1);
        unsigned long long not_ones = 0ull;  // dense mask: planes whose mask this thread saw differ from 1.0

        for (int j = 0; j < NB; ++j, ++jb) {
            // every consumer thread waits (also idle ones: a warp must not run ahead of the ring and arrive twice
            // in one phase of an empty barrier); the acquire also publishes the group's coefficients
            mbar_wait(bars + 8 * (BAR_FULL + stage), fphase);
            const int np = min(hs, N - j * hs);
            if (active) {
                uint32_t lrow = BF16 ? bring0 + (uint32_t)(stage * hs * NEc * rpc + r) * rbp : lring0 + (uint32_t)(stage * hs * rpc + r) * pitch4;
                for (int q = 0; q < np; ++q, lrow += plane4, coef_a += (uint32_t)sizeof(PlaneCoef)) {
                    const PlaneCoef k = load_coef(coef_a);
                    float mm[PX] = {};
                    bool perpix = false;
                    if (DENSE) {
                        load_window<PX>(lrow + mdelta + x04, mm);
                        perpix = !all_ones<PX>(mm);
                        if (perpix) not_ones |= 1ull << (j * hs + q);
                    }
                    if (DENSE && perpix) fwd_plane_any<MIX, PX, true>(srow, lrow, lrow + sdelta, pitch4, x04 + k.k4, W4, k, mm, acc);
                    else fwd_plane_any<MIX, PX, false, BF16>(srow, lrow, lrow + sdelta, pitch4, x04 + k.k4, W4, k, mm, acc);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + 8 * (BAR_EMPTY + stage));
            if (++stage == cfg.nst) stage = 0, fphase ^= 1;
        }
        if (DENSE && p.mask_rows) {
            // row summary for the backward pass: OR over the threads of the row (rows are zeroed by the host before launch)
            const unsigned lo = __reduce_or_sync(0xffffffffu, active ? (unsigned)not_ones : 0u);
            const unsigned hi = __reduce_or_sync(0xffffffffu, active ? (unsigned)(not_ones >> 32) : 0u);
            const int r_first = (int)((threadIdx.x & ~31u) / cfg.tpr), r_last = (int)((threadIdx.x | 31u) / cfg.tpr);
            if (r_first == r_last) {
                if (lane == 0 && active && (lo | hi)) atomicOr(p.mask_rows + row, ((unsigned long long)hi << 32) | lo);
            } else if (active && not_ones) {
                atomicOr(p.mask_rows + row, not_ones);  // a warp that straddles rows: per thread
            }
        }
        if (!active) continue;
        float o0[PX], o1[PX], o2[PX];
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            float invD;
            if constexpr (MIX) invD = 1.0f / acc.A[i];
            else invD = 1.0f / acc.S[i];
            o0[i] = acc.R0[i] * invD, o1[i] = acc.R1[i] * invD, o2[i] = acc.R2[i] * invD;
        }
        float* rr = p.out.rgb_rec + (int64_t)b * p.chw3 + rem;
        store_px_global<PX>(rr, o0);
        store_px_global<PX>(rr + p.hw, o1);
        store_px_global<PX>(rr + 2 * p.hw, o2);
        float* st = p.out.stats + (int64_t)b * (MIX ? PD_STATS_MIXTURE : PD_STATS_PLAIN) * p.hw + rem;
        store_px_global<PX>(st, acc.Ml2);
        store_px_global<PX>(st + p.hw, acc.S);
        if constexpr (MIX) {
            float dq[PX], nl[PX], na[PX];
#pragma unroll
            for (int i = 0; i < PX; ++i) {
                const float invS = 1.0f / acc.S[i];
                dq[i] = acc.Q[i] * invS + 1e-7f;  // layers.py:466
                nl[i] = -logf(dq[i]);
                na[i] = -logf(acc.Qa[i] * invS + 1e-7f);
            }
            store_px_global<PX>(st + 2 * p.hw, acc.A);
            store_px_global<PX>(st + 3 * p.hw, dq);
            store_px_global<PX>(p.out.nll + (int64_t)b * p.hw + rem, nl);
            if (p.d.automask) store_px_global<PX>(p.out.nll_auto + (int64_t)b * p.hw + rem, na);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
template <bool MIX, int PX>
struct BwdCtx {
    float g0[PX], g1[PX], g2[PX], nGbar[PX], Ml2[PX];  // nGbar = -sum_c g_c rgb_rec_c; Ml2 = reference logit * log2(e) + log2(sum exp)
    float tr[MIX ? PX : 1], tg[MIX ? PX : 1], tb[MIX ? PX : 1], Zinv[MIX ? PX : 1], gD[MIX ? PX : 1], gDD[MIX ? PX : 1];
};

__device__ __forceinline__ float sgn(float a, float b) { return (a > b) ? 1.0f : ((a < b) ? -1.0f : 0.0f); }

// phase A of one plane: dL/d(masked logit) [and dL/d(clamped-through sigma)] per target pixel into the exchange
// rows; returns this thread's contribution to dL/d(sign*disparity) of the plane row
template <bool MIX, bool WANT_DISP, int PX, int R, bool PERPIX, bool BF16 = false>
__device__ __forceinline__ float bwd_plane(uint32_t srow, uint32_t lrow, uint32_t sgrow, uint32_t pitch4, const PlaneCoef& k,
                                           const float (&mm)[PX], const BwdCtx<MIX, PX>& c, uint32_t dst, uint32_t fdelta) {
    // srow / lrow / sgrow: window-aligned shared addresses; dst: this thread's slot in the exchange row of dL/dlogit,
    // dst + fdelta the one of dL/dsigma
    constexpr int WF = PX + 4;
    float v[WF], pi[PX], Gn[PX], dlu[PX], gx[PX];
    load_plane_window<WF, BF16>(lrow, v);
#pragma unroll
    for (int i = 0; i < PX; ++i) {
        float t;
        if (PERPIX) t = fmaf(fmaf(k.wl0, v[R + i], k.wl1 * v[R + i + 1]), mm[i], -c.Ml2[i]);
        else t = fmaf(k.wl0, v[R + i], fmaf(k.wl1, v[R + i + 1], -c.Ml2[i]));
        pi[i] = fast_exp2(t);  // softmax probability: 1 / S sits in the reference (Ml2 + log2 S)
        if (WANT_DISP) dlu[i] = v[R + i + 1] - v[R + i];
    }
    float cr[PX], cg[PX], cb[PX], dr[WANT_DISP ? PX : 1], dg[WANT_DISP ? PX : 1], db[WANT_DISP ? PX : 1];
    load_window<WF>(srow, v);
#pragma unroll
    for (int i = 0; i < PX; ++i) {
        cr[i] = fmaf(k.wc0, v[R + i], k.wc1 * v[R + i + 1]);
        if (PERPIX) cr[i] *= mm[i];
        if (WANT_DISP) dr[i] = v[R + i + 1] - v[R + i];
    }
    load_window<WF>(srow + pitch4, v);
#pragma unroll
    for (int i = 0; i < PX; ++i) {
        cg[i] = fmaf(k.wc0, v[R + i], k.wc1 * v[R + i + 1]);
        if (PERPIX) cg[i] *= mm[i];
        if (WANT_DISP) dg[i] = v[R + i + 1] - v[R + i];
    }
    load_window<WF>(srow + 2 * pitch4, v);
#pragma unroll
    for (int i = 0; i < PX; ++i) {
        cb[i] = fmaf(k.wc0, v[R + i], k.wc1 * v[R + i + 1]);
        if (PERPIX) cb[i] *= mm[i];
        if (WANT_DISP) db[i] = v[R + i + 1] - v[R + i];
        Gn[i] = fmaf(c.g0[i], cr[i], fmaf(c.g1[i], cg[i], fmaf(c.g2[i], cb[i], c.nGbar[i])));  // sum_c g_c colour_c - Gbar
    }
    float dl[PX], ds[MIX ? PX : 1];
    if constexpr (!MIX) {
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            dl[i] = pi[i] * Gn[i];
            if (WANT_DISP) gx[i] = fmaf(dl[i], dlu[i], pi[i] * (c.g0[i] * dr[i] + c.g1[i] * dg[i] + c.g2[i] * db[i]));
        }
    } else {
        load_plane_window<WF, BF16>(sgrow, v);
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            float sraw = fmaf(k.wc0, v[R + i], k.wc1 * v[R + i + 1]);
            if (PERPIX) sraw *= mm[i];
            const float sg = fminf(fmaxf(sraw, 0.01f), 1.0f);
            const float inv = fast_rcp(sg);
            // with a = 1/sigma, w = pi a / Z (compositing weight), q = gD pi lap (the plane's share of dL/dD):
            //   dL/dlogit = w dG + q - pi gDD,  dL/dsigma = a (q a (err - sigma) - w dG),  dL/dcolour = w g -+ q a / 3
            const float w = pi[i] * inv * c.Zinv[i];
            const float dG = Gn[i];
            const float err = (fabsf(cr[i] - c.tr[i]) + fabsf(cg[i] - c.tg[i]) + fabsf(cb[i] - c.tb[i])) * (1.0f / 3.0f);
            const float ea = err * inv;
            const float q = c.gD[i] * pi[i] * (0.5f * inv) * fast_exp2(-ea * kLog2e);
            const float wdG = w * dG;
            dl[i] = fmaf(-pi[i], c.gDD[i], wdG + q);
            const float dsgt = inv * fmaf(q, ea - 1.0f, -wdG);  // q a (err - sigma) = q (err a - 1)
            ds[i] = (sraw >= 0.01f && sraw <= 1.0f) ? dsgt : 0.0f;  // clamp backward
            if (WANT_DISP) {
                const float ce = -q * inv * (1.0f / 3.0f);
                const float dcr = fmaf(w, c.g0[i], ce * sgn(cr[i], c.tr[i]));
                const float dcg = fmaf(w, c.g1[i], ce * sgn(cg[i], c.tg[i]));
                const float dcb = fmaf(w, c.g2[i], ce * sgn(cb[i], c.tb[i]));
                gx[i] = dcr * dr[i] + dcg * dg[i] + dcb * db[i] + dl[i] * dlu[i] + ds[i] * (v[R + i + 1] - v[R + i]);
            }
        }
    }
    float gsum = 0.0f;
    if (PERPIX) {
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            if (WANT_DISP) gsum = fmaf(gx[i], mm[i], gsum);
            dl[i] *= mm[i];
            if constexpr (MIX) ds[i] *= mm[i];
        }
    } else if (WANT_DISP) {
#pragma unroll
        for (int i = 0; i < PX; ++i) gsum += gx[i];
        gsum *= k.m;
    }
#pragma unroll
    for (int i = 0; i < PX / 4; ++i) {
        sts128(dst + 16 * i, dl[4 * i], dl[4 * i + 1], dl[4 * i + 2], dl[4 * i + 3]);
        if constexpr (MIX) sts128(dst + fdelta + 16 * i, ds[4 * i], ds[4 * i + 1], ds[4 * i + 2], ds[4 * i + 3]);
    }
    return gsum;
}

template <bool MIX, bool WANT_DISP, int PX, bool PERPIX, bool BF16 = false>
__device__ __forceinline__ float bwd_plane_any(uint32_t srow, uint32_t lrow, uint32_t sgrow, uint32_t pitch4, int at4, int W4, const PlaneCoef& k,
                                               const float (&mm)[PX], const BwdCtx<MIX, PX>& c, uint32_t dst, uint32_t fdelta) {
    const int wo = window_off(at4, W4);
    const int wl = BF16 ? wo / 2 : wo;  // see fwd_plane_any
    srow += wo, lrow += wl, sgrow += wl;
    if (at4 & 8) {
        if (at4 & 4) return bwd_plane<MIX, WANT_DISP, PX, 3, PERPIX, BF16>(srow, lrow, sgrow, pitch4, k, mm, c, dst, fdelta);
        return bwd_plane<MIX, WANT_DISP, PX, 2, PERPIX, BF16>(srow, lrow, sgrow, pitch4, k, mm, c, dst, fdelta);
    }
    if (at4 & 4) return bwd_plane<MIX, WANT_DISP, PX, 1, PERPIX, BF16>(srow, lrow, sgrow, pitch4, k, mm, c, dst, fdelta);
    return bwd_plane<MIX, WANT_DISP, PX, 0, PERPIX, BF16>(srow, lrow, sgrow, pitch4, k, mm, c, dst, fdelta);
}

// phase B of one plane: gradient of source column j = wc0 * D[j - k0] + wc1 * D[j - k0 - 1]
template <int PX, int R>
__device__ __forceinline__ void gather_row(uint32_t dwin, float w0, float w1, float (&g)[PX]) {
    constexpr int WF = PX + 4;
    float v[WF];
    load_window<WF>(dwin, v);
#pragma unroll
    for (int i = 0; i < PX; ++i) g[i] = fmaf(w0, v[R + i + 1], w1 * v[R + i]);
}

// drow: shared address of the exchange row interior; at4 = 4 * (x0 - k0 - 1)
template <int PX>
__device__ __forceinline__ void gather_any(uint32_t drow, int at4, int W4, float w0, float w1, float (&g)[PX]) {
    const uint32_t dwin = drow + window_off(at4, W4);
    if (at4 & 8) {
        if (at4 & 4) gather_row<PX, 3>(dwin, w0, w1, g);
        else gather_row<PX, 2>(dwin, w0, w1, g);
    } else {
        if (at4 & 4) gather_row<PX, 1>(dwin, w0, w1, g);
        else gather_row<PX, 0>(dwin, w0, w1, g);
    }
}

template <bool MIX, int MASKMODE, bool WANT_DISP, int PX, int THREADS, int MINB, bool BF16 = false>
__global__ void __launch_bounds__(THREADS, MINB) rows_bwd_stream(const WarpParams p, const StreamCfg cfg) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr bool DENSE = (MASKMODE == SMASK_DENSE), SUMM = (MASKMODE == SMASK_SUMMARY);
    constexpr int NE = MIX ? 2 : 1;
    const int W = p.d.W, H = p.d.H, N = p.d.N, pitch = cfg.pitch, rpc = cfg.rpc, hs = cfg.hs, NB = cfg.nblk;
    const int rows_total = p.d.B * H;
    const Smem s = carve(smem_raw, cfg, N, MIX, DENSE, NE, WANT_DISP);
    zero_pads(s, cfg, W);
    if constexpr (BF16) zero_pads_bf16(s, cfg, W, cfg.nst * cfg.hs * NE * cfg.rpc);
    if (threadIdx.x == 0) {
        for (int i = 0; i < cfg.nst; ++i) {
            mbar_init(s.bars + BAR_FULL + i, 1);
            mbar_init(s.bars + BAR_EMPTY + i, (uint32_t)(cfg.nc / 32));
        }
        mbar_init(s.bars + BAR_SRC, 1);
        mbar_init(s.bars + BAR_SRC + 1, 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int nit = (cfg.ngroups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if ((int)threadIdx.x >= cfg.nc) {
        producer_loop<MIX, MASKMODE, BF16>(p, cfg, s, nit);
        return;
    }
    const int lane = threadIdx.x & 31;
    const int r = threadIdx.x / cfg.tpr;
    const int x0 = (threadIdx.x - r * cfg.tpr) * PX;
    const uint32_t bars = smem_u32(s.bars), coef0 = smem_u32(s.coef), src0 = smem_u32(s.src + PAD), lring0 = smem_u32(s.lring + PAD);
    const uint32_t dbuf0 = smem_u32(s.dbuf + PAD);
    const uint32_t pitch4 = (uint32_t)pitch * 4u, rowpitch4 = (uint32_t)rpc * pitch4;
    const uint32_t rbp = (uint32_t)stream_bf16_row_bytes(cfg);  // see rows_fwd_stream
    const uint32_t sdelta = BF16 ? (uint32_t)rpc * rbp : (uint32_t)((s.sring - s.lring) * sizeof(float));
    const uint32_t mdelta = (uint32_t)((s.mring - s.lring) * sizeof(float));
    const uint32_t plane4 = BF16 ? NE * (uint32_t)rpc * rbp : rowpitch4;
    const uint32_t bring0 = smem_u32(s.bring) + PADB * 2;
    const int x04 = x0 * 4, W4 = W * 4;
    int stage = 0, jb = 0;
    uint32_t fphase = 0;
    const float gph = upstream_scale(p);

    for (int it = 0; it < nit; ++it) {
        const int g = blockIdx.x + it * gridDim.x;
        const int row = g * rpc + r;
        const bool active = (r < rpc) && (row < rows_total);
        const int b = active ? row / H : 0, y = active ? row - b * H : 0;
        const int64_t rem = (int64_t)y * W + x0;
        // d/d disparity sums of the group.  When every warp works inside one row (tpr % 32 == 0) each warp owns a private row of
        // sums and adds to it with plain read-modify-writes (a shared float atomicAdd compiles to a compare-and-swap loop that
        // all warps of a row would contend for); otherwise rows are shared and the adds are atomic.
        const bool wrows = WANT_DISP && (cfg.tpr & 31) == 0;
        const int wpr = cfg.tpr >> 5;  // warps per row (wrows)
        if (WANT_DISP) {
            const int nacc = (wrows ? cfg.nc / 32 : rpc) * N;
            for (int i = threadIdx.x; i < nacc; i += cfg.nc) s.gacc[i] = 0.0f;
            consumer_sync(cfg.nc);
        }

        BwdCtx<MIX, PX> c;
        if (active) {
            const float* rp = p.out.rgb_rec + (int64_t)b * p.chw3 + rem;
            float ra[PX], rb[PX], rc[PX], Sv[PX];
            // upstream gradient, optionally formed from the photometric forward's unit gradient (pd_warp_grad_out)
            const int64_t gi = (int64_t)b * p.chw3 + rem, gpix = (int64_t)b * p.hw + rem;
            load_upstream<PX>(p, gph, gi, gpix, c.g0);
            load_upstream<PX>(p, gph, gi + p.hw, gpix, c.g1);
            load_upstream<PX>(p, gph, gi + 2 * p.hw, gpix, c.g2);
            load_px_global<PX>(rp, ra);
            load_px_global<PX>(rp + p.hw, rb);
            load_px_global<PX>(rp + 2 * p.hw, rc);
            const float* st = p.out.stats + (int64_t)b * (MIX ? PD_STATS_MIXTURE : PD_STATS_PLAIN) * p.hw + rem;
            load_px_global<PX>(st, c.Ml2);
            load_px_global<PX>(st + p.hw, Sv);
#pragma unroll
            for (int i = 0; i < PX; ++i) {
                c.nGbar[i] = -(c.g0[i] * ra[i] + c.g1[i] * rb[i] + c.g2[i] * rc[i]);
                c.Ml2[i] += log2f(Sv[i]);
            }
            if constexpr (MIX) {
                const float* tp = p.in.tgt + (int64_t)b * p.chw3 + rem;
                float Av[PX], Dv[PX], gv[PX];
                load_px_global<PX>(tp, c.tr);
                load_px_global<PX>(tp + p.hw, c.tg);
                load_px_global<PX>(tp + 2 * p.hw, c.tb);
                load_px_global<PX>(st + 2 * p.hw, Av);
                load_px_global<PX>(st + 3 * p.hw, Dv);
#pragma unroll
                for (int i = 0; i < PX / 4; ++i) {
                    const float4 t4 = upstream_nll4(p, gph, gpix + 4 * i);
                    gv[4 * i] = t4.x, gv[4 * i + 1] = t4.y, gv[4 * i + 2] = t4.z, gv[4 * i + 3] = t4.w;
                }
#pragma unroll
                for (int i = 0; i < PX; ++i) {
                    c.Zinv[i] = Sv[i] / Av[i];              // 1/Z, Z = sum pi/sigma = A/S
                    c.gD[i] = -gv[i] / Dv[i];               // d loss / d D, nll = -log D
                    c.gDD[i] = c.gD[i] * (Dv[i] - 1e-7f);   // = sum_k pi_k P_k
                }
            }
        }
        const uint32_t coef_g = coef0 + (uint32_t)(((it & 1) * rpc + r) * N) * (uint32_t)sizeof(PlaneCoef);
        const uint32_t srow = src0 + (uint32_t)(((it & 1) * rpc + r) * 3) * pitch4;
        mbar_wait(bars + 8 * (BAR_SRC + (it & 1)), (it >> 1) & 1);

        for (int j = 0; j < NB; ++j, ++jb) {
            const int n0 = j * hs, np = min(hs, N - n0);
            // exchange rows of this block: [plane][NE][rpc][pitch], double-buffered over blocks
            const uint32_t dblk = dbuf0 + (uint32_t)((jb & 1) * hs * NE * rpc + r) * pitch4;
            mbar_wait(bars + 8 * (BAR_FULL + stage), fphase);
            // ---------------- phase A: per-target gradients into the exchange rows ----------------
            {
                uint32_t lrow = BF16 ? bring0 + (uint32_t)(stage * hs * NE * rpc + r) * rbp : lring0 + (uint32_t)(stage * hs * rpc + r) * pitch4;
                uint32_t coef_a = coef_g + (uint32_t)n0 * (uint32_t)sizeof(PlaneCoef);
                uint32_t drow = dblk;
                for (int q = 0; q < np; ++q, lrow += plane4, coef_a += (uint32_t)sizeof(PlaneCoef), drow += NE * rowpitch4) {
                    float gsum = 0.0f;
                    if (active) {
                        const PlaneCoef k = load_coef(coef_a);
                        float mm[PX] = {};
                        bool perpix = false;
                        if (DENSE) {
                            load_window<PX>(lrow + mdelta + x04, mm);
                            perpix = !all_ones<PX>(mm);
                        } else if (SUMM && k.skip == 0.0f) {  // a row the summary could not fold away: straight from global memory
                            load_px_global<PX>(reinterpret_cast<const float*>(p.in.mask) + soff(p.d.mask_stride, b, n0 + q, y, x0), mm);
                            perpix = !all_ones<PX>(mm);
                        }
                        if ((DENSE || SUMM) && perpix)
                            gsum = bwd_plane_any<MIX, WANT_DISP, PX, true>(srow, lrow, lrow + sdelta, pitch4, x04 + k.k4, W4, k, mm, c, drow + x04, rowpitch4);
                        else
                            gsum = bwd_plane_any<MIX, WANT_DISP, PX, false, BF16>(srow, lrow, lrow + sdelta, pitch4, x04 + k.k4, W4, k, mm, c, drow + x04, rowpitch4);
                    }
                    if (WANT_DISP) {
                        // warp sum when the whole warp works on one row, per-thread shared atomics otherwise
                        const int n = n0 + q;
                        const int r_first = (int)((threadIdx.x & ~31u) / cfg.tpr), r_last = (int)((threadIdx.x | 31u) / cfg.tpr);
                        if (wrows) {
                            const float sum = warp_sum(gsum);
                            if (lane == 0) s.gacc[(threadIdx.x >> 5) * N + n] += sum * p.d.disp_sign;
                        } else if (r_first == r_last) {
                            const float sum = warp_sum(gsum);
                            if (lane == 0 && r < rpc && sum != 0.0f) atomicAdd(s.gacc + r * N + n, sum * p.d.disp_sign);
                        } else if (active && gsum != 0.0f) {
                            atomicAdd(s.gacc + r * N + n, gsum * p.d.disp_sign);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + 8 * (BAR_EMPTY + stage));  // the ring stage is no longer needed
            if (++stage == cfg.nst) stage = 0, fphase ^= 1;
            // exchange rows of block jb complete.  They are double-buffered: block jb+1 writes the other buffer, and
            // block jb+2 is only written after the next barrier, which every thread reaches after this gather.
            consumer_sync(cfg.nc);
            // ---------------- phase B: gather per source column, one streaming store per row ----------------
            if (active) {
                uint32_t coef_a = coef_g + (uint32_t)n0 * (uint32_t)sizeof(PlaneCoef);
                uint32_t drow = dblk;
                int64_t o = (((int64_t)b * N + n0) * H + y) * W + x0;
                for (int q = 0; q < np; ++q, coef_a += (uint32_t)sizeof(PlaneCoef), drow += NE * rowpitch4, o += p.hw) {
                    int k4;
                    float wc0, wc1;
                    load_coef_gather(coef_a, k4, wc0, wc1);
                    // with a dense mask the per-pixel mask is already folded into the exchange rows (k.m == 1)
                    const int at4 = x04 - k4 - 4;
                    float gg[PX];
                    if (p.gin.g_logits) {
                        gather_any<PX>(drow, at4, W4, wc0, wc1, gg);
                        if constexpr (BF16) store_px_stream_bf16<PX>(p.gin.g_logits, o, gg);
                        else store_px_stream<PX>(p.gin.g_logits + o, gg);
                    }
                    if constexpr (MIX) if (p.gin.g_sigma) {
                        gather_any<PX>(drow + rowpitch4, at4, W4, wc0, wc1, gg);
                        if constexpr (BF16) store_px_stream_bf16<PX>(p.gin.g_sigma, o, gg);
                        else store_px_stream<PX>(p.gin.g_sigma + o, gg);
                    }
                }
            }
        }
        if (WANT_DISP) {
            consumer_sync(cfg.nc);  // every shared atomic of this group has landed
            for (int i = threadIdx.x; i < rpc * N; i += cfg.nc) {
                const int rr = i / N, n = i - rr * N;
                const int rw = g * rpc + rr;
                float v = 0.0f;
                if (wrows) {
                    for (int w = rr * wpr; w < (rr + 1) * wpr; ++w) v += s.gacc[w * N + n];
                } else {
                    v = s.gacc[i];
                }
                if (rw < rows_total && v != 0.0f) {
                    const int bb = rw / H, yy = rw - bb * H;
                    atomicAdd(p.gin.g_disp + soff(p.gin.g_disp_stride, bb, n, yy, 0), v);
                }
            }
            consumer_sync(cfg.nc);  // before the next group zeroes gacc
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
inline bool ts_aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

inline int stream_mask_mode(const WarpParams& p) {
    return (p.d.mask_dtype == PD_MASK_NONE || p.d.mask_stride.x == 0) ? SMASK_ROW : SMASK_DENSE;
}

inline bool stream_path_supported(const WarpParams& p) {
    if (p.d.warp_type != PD_WARP_DISP) return false;
    const int W = p.d.W;
    if (W % 4 != 0 || W < 8 || W > 2048) return false;
    if (p.d.disp_stride.x != 0) return false;  // disparity must be constant along x
    if (p.d.mask_dtype != PD_MASK_NONE && p.d.mask_stride.x != 0) {
        // dense mask rows travel through TMA: fp32, unit x stride, 16-byte aligned rows
        if (p.d.mask_dtype != PD_MASK_F32 || p.d.mask_stride.x != 1) return false;
        if (p.d.mask_stride.y % 4 || p.d.mask_stride.n % 4 || p.d.mask_stride.b % 4 || !ts_aligned16(p.in.mask)) return false;
    }
    if (p.gin.g_disp && p.gin.g_disp_stride.x != 0) return false;  // d/d disp only reduced over x
    const void* ptrs[] = {p.in.src, p.in.tgt, p.in.logits, p.in.sigma, p.out.rgb_rec, p.out.stats, p.out.nll, p.out.nll_auto,
                          p.gout.g_rgb_rec, p.gout.g_nll, p.gout.g_unit, p.gout.g_unit_nll, p.gout.g_pred, p.gout.mask_novel,
                          p.gin.g_logits, p.gin.g_sigma};
    for (const void* q : ptrs)
        if (q && !ts_aligned16(q)) return false;
    return true;
}

template <int PX, int THREADS>
inline StreamCfg stream_cfg(const WarpParams& p, bool mix, bool dense, int ne_bwd, bool want_disp, int budget_kb = 0) {
    StreamCfg c;
    c.bf16 = p.d.dtype == PD_DTYPE_BF16 ? 1 : 0;
    c.tpr = p.d.W / PX;
    c.rpc = THREADS / c.tpr;
    if (c.rpc < 1) c.rpc = 1;
    if (c.rpc > 8) c.rpc = 8;
    c.pitch = p.d.W + 2 * PAD;
    c.nc = ((c.rpc * c.tpr + 31) / 32) * 32;
    const pd_tuning& tn = tuning();  // clamped when it was set: hs >= 1, 1 <= nst <= MAX_STAGES
    // measured (profiles/r2c_*): the narrow plain backward prefers fewer, longer blocks (a consumer barrier per block)
    const bool narrow_plain_bwd = ne_bwd > 0 && !mix && c.nc <= 160;
    // bf16 forward: rows are half as long, so a stage of 8 planes costs what 4 cost in fp32 (measured, profiles/r2b2_*: 0.1026 -> 0.0989 ms)
    c.hs = tn.stream_hs > 0 ? tn.stream_hs : (narrow_plain_bwd ? 5 : ((c.bf16 && ne_bwd == 0) ? 8 : 4));
    c.nst = tn.stream_nst > 0 ? tn.stream_nst : (narrow_plain_bwd ? 2 : 3);
    // measured (profiles/r2v3_*): the narrow plain forward prefers blocks that divide the plane count evenly -- N = 49 as 7 blocks
    // of 7 planes in a two-stage ring (three resident CTAs) beats 13 blocks of 4 in three stages (four CTAs) by 3 %; 8, 9 or 10
    // planes per block (ragged last block) are slower than either
    if (tn.stream_hs == 0 && tn.stream_nst == 0 && ne_bwd == 0 && !mix && c.nc <= 160 && p.d.N % 7 == 0) c.hs = 7, c.nst = c.bf16 ? 3 : 2;
    if (c.hs > p.d.N) c.hs = p.d.N;
    if (c.nst > MAX_STAGES) c.nst = MAX_STAGES;
    // shrink the pipeline until the CTA fits the shared-memory budget (default: three CTAs per SM)
    // (three CTAs of <= 192 threads per SM, two of the 352-thread CTAs that wide rows need)
    // (the wide mixture backward runs one CTA per SM on registers anyway: it gets a deep ring instead, cfg3 0.82 -> 0.69 ms)
    const int dflt_kb = c.nc > 160 ? ((mix && ne_bwd > 0) ? 200 : 110) : 72;
    const size_t budget = (size_t)(budget_kb > 0 ? budget_kb : (tn.stream_smem_kb > 0 ? tn.stream_smem_kb : dflt_kb)) * 1024;
    while (stream_smem_bytes(c, p.d.N, mix, dense, ne_bwd, want_disp) > budget) {
        if (c.nst > 2) --c.nst;
        else if (c.hs > 1) --c.hs;
        else break;
    }
    c.nblk = (p.d.N + c.hs - 1) / c.hs;
    if (c.nst > c.nblk) c.nst = c.nblk;  // reuse of the per-group buffers relies on nst <= nblk (see producer_loop)
    c.ngroups = (p.d.B * p.d.H + c.rpc - 1) / c.rpc;
    c.l2_hint = tn.stream_no_l2_hint ? 0 : 1;
    // measured (profiles/r2e2_*): staging the next row group early gains 0.6 - 1 % except in the narrow plain backward,
    // whose two-stage ring loses 1 % to the producer's detour
    c.early = narrow_plain_bwd ? 0 : 1;
    return c;
}

template <typename K>
inline void stream_smem_optin(K kern, size_t smem) {
    smem_optin((const void*)kern, smem);
}

inline int stream_grid(int ngroups, int threads, size_t smem, const void* kernel) {
    int per_sm = resident_ctas(kernel, threads, smem);
    const int cap = tuning().stream_ctas_per_sm;
    if (cap > 0 && per_sm > cap) per_sm = cap;
    const long long g = (long long)sm_count() * per_sm;
    return (int)(g < ngroups ? g : ngroups);
}

// The two translation units that include this header define PD_TS_FWD_ONLY / PD_TS_BWD_ONLY so that each instantiates one
// kernel family only (non-template inline launchers would instantiate every kernel they mention in both).
#ifndef PD_TS_BWD_ONLY
// THREADS = consumer threads the row groups are packed into + the producer warp
template <bool MIX, int MASKMODE, int PX, int THREADS, int MINB, bool BF16 = false>
inline bool launch_fwd_stream_t(const WarpParams& p, cudaStream_t st, bool dry) {
    // more than four resident CTAs only fit with a shallower ring: 220 KB / MINB each
    const StreamCfg c = stream_cfg<PX, THREADS - 32>(p, MIX, MASKMODE == SMASK_DENSE, 0, false, MINB > 4 ? 220 / MINB : 0);
    const int threads = c.nc + 32;
    if (threads > THREADS) return false;
    const size_t smem = stream_smem_bytes(c, p.d.N, MIX, MASKMODE == SMASK_DENSE, 0, false);
    if (smem > 220 * 1024) return false;
    if (dry) return true;
    auto kern = rows_fwd_stream<MIX, MASKMODE, PX, THREADS, MINB, BF16>;
    stream_smem_optin(kern, smem);
    kern<<<stream_grid(c.ngroups, threads, smem, (const void*)kern), threads, smem, st>>>(p, c);
    return true;
}

#endif  // PD_TS_BWD_ONLY

#ifndef PD_TS_FWD_ONLY
template <bool MIX, int MASKMODE, bool WANT_DISP, int PX, int THREADS, int MINB, bool BF16 = false>
inline bool launch_bwd_stream_t(const WarpParams& p, cudaStream_t st, bool dry) {
    const StreamCfg c = stream_cfg<PX, THREADS - 32>(p, MIX, MASKMODE == SMASK_DENSE, MIX ? 2 : 1, WANT_DISP, MINB > 3 ? 220 / MINB : 0);
    const int threads = c.nc + 32;
    if (threads > THREADS) return false;
    const size_t smem = stream_smem_bytes(c, p.d.N, MIX, MASKMODE == SMASK_DENSE, MIX ? 2 : 1, WANT_DISP);
    if (smem > 220 * 1024) return false;
    if (dry) return true;
    auto kern = rows_bwd_stream<MIX, MASKMODE, WANT_DISP, PX, THREADS, MINB, BF16>;
    stream_smem_optin(kern, smem);
    kern<<<stream_grid(c.ngroups, threads, smem, (const void*)kern), threads, smem, st>>>(p, c);
    return true;
}

#endif  // PD_TS_FWD_ONLY

#ifndef PD_TS_BWD_ONLY
template <bool MIX, int MASKMODE>
inline bool launch_fwd_stream_m(const WarpParams& p, cudaStream_t st, bool dry) {
    const int W = p.d.W;
    if (p.d.dtype == PD_DTYPE_BF16) {  // bf16 storage: row masks, 4 pixels per thread, rows of whole 16-byte units
        if constexpr (MASKMODE == SMASK_ROW) {
            if (W % 8 != 0) return false;
            if (W / 4 <= 160) return launch_fwd_stream_t<MIX, MASKMODE, 4, 192, MIX ? 2 : 4, true>(p, st, dry);
            if (W / 4 <= 320) return launch_fwd_stream_t<MIX, MASKMODE, 4, 352, MIX ? 1 : 2, true>(p, st, dry);
        }
        return false;
    }
    if (W % 8 == 0 && W / 8 <= 160 && tuning().stream_px8) return launch_fwd_stream_t<MIX, MASKMODE, 8, 192, 2>(p, st, dry);
    if constexpr (!MIX && MASKMODE == SMASK_ROW) {  // occupancy experiments (pd_tuning.stream_fwd_minb): fewer registers, shallower ring
        if (W / 4 <= 160 && tuning().stream_fwd_minb == 5) return launch_fwd_stream_t<MIX, MASKMODE, 4, 192, 5>(p, st, dry);
        if (W / 4 <= 160 && tuning().stream_fwd_minb == 6) return launch_fwd_stream_t<MIX, MASKMODE, 4, 192, 6>(p, st, dry);
    }
    if (W / 4 <= 160) return launch_fwd_stream_t<MIX, MASKMODE, 4, 192, MIX ? 2 : 4>(p, st, dry);
    if (W / 4 <= 320) return launch_fwd_stream_t<MIX, MASKMODE, 4, 352, MIX ? 2 : 2>(p, st, dry);
    return false;
}

inline bool launch_fwd_stream(const WarpParams& p, cudaStream_t st, bool dry = false) {
    const int mm = stream_mask_mode(p);
    if (p.d.mixture) return mm == SMASK_ROW ? launch_fwd_stream_m<true, SMASK_ROW>(p, st, dry) : launch_fwd_stream_m<true, SMASK_DENSE>(p, st, dry);
    return mm == SMASK_ROW ? launch_fwd_stream_m<false, SMASK_ROW>(p, st, dry) : launch_fwd_stream_m<false, SMASK_DENSE>(p, st, dry);
}

#endif  // PD_TS_BWD_ONLY

#ifndef PD_TS_FWD_ONLY
template <bool MIX, int MASKMODE, bool WANT_DISP>
inline bool launch_bwd_stream_w(const WarpParams& p, cudaStream_t st, bool dry) {
    const int W = p.d.W;
    if (p.d.dtype == PD_DTYPE_BF16) {
        if constexpr (MASKMODE == SMASK_ROW) {
            if (W % 8 != 0) return false;
            if (W / 4 <= 160) return launch_bwd_stream_t<MIX, MASKMODE, WANT_DISP, 4, 192, MIX ? 2 : 4, true>(p, st, dry);
            if (W / 4 <= 320) return launch_bwd_stream_t<MIX, MASKMODE, WANT_DISP, 4, 352, 1, true>(p, st, dry);
        }
        return false;
    }
    if (W % 8 == 0 && W / 8 <= 160 && tuning().stream_px8) return launch_bwd_stream_t<MIX, MASKMODE, WANT_DISP, 8, 192, 1>(p, st, dry);
    if constexpr (!MIX && MASKMODE == SMASK_ROW) {
        // plain narrow backward: four resident CTAs at 80 registers with a 2 x 3 ring beat three at 96 registers with 2 x 5
        // (0.1619 vs 0.1640 ms at cfg 2, profiles/r2t_*); pd_tuning.stream_bwd_minb = 3 keeps the three-CTA build for A/B runs
        if (W / 4 <= 160 && tuning().stream_bwd_minb != 3) return launch_bwd_stream_t<MIX, MASKMODE, WANT_DISP, 4, 192, 4>(p, st, dry);
    }
    if (W / 4 <= 160) return launch_bwd_stream_t<MIX, MASKMODE, WANT_DISP, 4, 192, MIX ? 2 : 3>(p, st, dry);
    if (W / 4 <= 320) return launch_bwd_stream_t<MIX, MASKMODE, WANT_DISP, 4, 352, 1>(p, st, dry);
    return false;
}


template <bool MIX, int MASKMODE>
inline bool launch_bwd_stream_m(const WarpParams& p, cudaStream_t st, bool dry) {
    return p.gin.g_disp ? launch_bwd_stream_w<MIX, MASKMODE, true>(p, st, dry) : launch_bwd_stream_w<MIX, MASKMODE, false>(p, st, dry);
}

inline bool launch_bwd_stream(const WarpParams& p, cudaStream_t st, bool dry = false) {
    int mm = stream_mask_mode(p);
    if (mm == SMASK_DENSE && p.mask_rows) mm = SMASK_SUMMARY;  // the forward pass left a row summary of the mask
    if (p.d.mixture) {
        if (mm == SMASK_SUMMARY) return launch_bwd_stream_m<true, SMASK_SUMMARY>(p, st, dry);
        return mm == SMASK_ROW ? launch_bwd_stream_m<true, SMASK_ROW>(p, st, dry) : launch_bwd_stream_m<true, SMASK_DENSE>(p, st, dry);
    }
    if (mm == SMASK_SUMMARY) return launch_bwd_stream_m<false, SMASK_SUMMARY>(p, st, dry);
    return mm == SMASK_ROW ? launch_bwd_stream_m<false, SMASK_ROW>(p, st, dry) : launch_bwd_stream_m<false, SMASK_DENSE>(p, st, dry);
}
#endif  // PD_TS_FWD_ONLY

}  // namespace ts
}  // namespace pd
