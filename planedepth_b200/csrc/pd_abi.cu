// C ABI of planedepth_b200 (see include/planedepth_b200.h): argument validation, kernel selection
// and launches.  No torch / ATen dependency; everything is enqueued on the caller's stream.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>

#include "pd_loss.cuh"
#include "pd_occlusion.cuh"
#include "pd_tail.cuh"
#include "pd_warp_general.cuh"
#include "pd_warp_homo.cuh"
#include "pd_warp_rows.cuh"
#include "pd_warp_stream.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};  // process-wide: autograd runs backward on its own threads

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PD_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    ++g_launches;
    return PD_OK;
}

// bit-faithful coordinate arithmetic requested by the caller (or forced for a whole process, for tests)
bool exact_coords(const pd_warp_desc* d) { return (d->flags & PD_FLAG_EXACT_COORDS) || getenv("PD_EXACT_COORDS"); }

int check_device() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(PD_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return fail(PD_ERR_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
    if (major != 10) return fail(PD_ERR_ARCH, "planedepth_b200 is built for sm_100a only (device is sm_%d*)", major);
    return PD_OK;
}

int validate_warp(const pd_warp_desc* d, const pd_warp_in* in) {
    if (!d || !in) return fail(PD_ERR_ARG, "NULL descriptor");
    if (d->B < 1 || d->N < 1 || d->H < 2 || d->W < 2) return fail(PD_ERR_SHAPE, "B,N >= 1 and H,W >= 2 required (got %d,%d,%d,%d)", d->B, d->N, d->H, d->W);
    if (d->N > PD_MAX_PLANES) return fail(PD_ERR_SHAPE, "N=%d exceeds PD_MAX_PLANES", d->N);
    if ((int64_t)d->H * d->W >= (1ll << 31)) return fail(PD_ERR_SHAPE, "H*W too large");
    if (d->warp_type < PD_WARP_DISP || d->warp_type > PD_WARP_DEPTH) return fail(PD_ERR_ARG, "bad warp_type %d", d->warp_type);
    if (!in->src || !in->logits) return fail(PD_ERR_ARG, "src / logits must not be NULL");
    if (d->mixture && (!in->sigma || !in->tgt)) return fail(PD_ERR_ARG, "mixture needs sigma and tgt");
    if (d->warp_type == PD_WARP_HOMOGRAPHY) {
        if (!in->hmat || !in->cam) return fail(PD_ERR_ARG, "homography_warp needs hmat and cam");
    } else {
        if (!in->disp) return fail(PD_ERR_ARG, "disp_warp / depth_warp need disp");
        if (d->warp_type == PD_WARP_DEPTH && !in->cam) return fail(PD_ERR_ARG, "depth_warp needs cam");
        if (d->mask_dtype != PD_MASK_NONE && !in->mask) return fail(PD_ERR_ARG, "mask_dtype set but mask is NULL");
        if (d->mask_dtype < PD_MASK_NONE || d->mask_dtype > PD_MASK_U8) return fail(PD_ERR_ARG, "bad mask_dtype");
    }
    return PD_OK;
}

pd::WarpParams make_params(const pd_warp_desc* d, const pd_warp_in* in) {
    pd::WarpParams p;
    memset(&p, 0, sizeof(p));
    p.d = *d;
    p.in = *in;
    if (!in->mask) p.d.mask_dtype = PD_MASK_NONE;
    p.wm1 = (float)(d->W - 1);
    p.hm1 = (float)(d->H - 1);
    p.depth_c = 0.1f * 0.58f * (float)d->W;
    p.hw = (int64_t)d->H * d->W;
    p.chw3 = 3 * p.hw;
    p.warp_aligned_rows = (d->W % 32 == 0);
    return p;
}

// Dense-mask row summary behind the saved statistics (WarpParams::mask_rows): per image row one 64-bit set over planes.
size_t stats_floats(const pd_warp_desc* d) { return (size_t)d->B * (d->mixture ? PD_STATS_MIXTURE : PD_STATS_PLAIN) * d->H * d->W; }
size_t mask_summary_bytes(const pd_warp_desc* d) { return (size_t)d->B * d->H * sizeof(unsigned long long); }

void attach_mask_summary(pd::WarpParams& p, const float* stats) {
    const pd_warp_desc& d = p.d;
    const bool keep = d.warp_type == PD_WARP_DISP && p.in.mask && d.mask_dtype == PD_MASK_F32 && d.mask_stride.x == 1 && d.N <= 64 &&
                      (stats_floats(&d) % 2 == 0) && !getenv("PD_NO_MASK_SUMMARY");
    p.mask_rows = keep ? reinterpret_cast<unsigned long long*>(const_cast<float*>(stats) + stats_floats(&d)) : nullptr;
}

int64_t strided_extent(const pd_strides4& s, int B, int N, int H, int W) {
    return (int64_t)(B - 1) * s.b + (int64_t)(N - 1) * s.n + (int64_t)(H - 1) * s.y + (int64_t)(W - 1) * s.x + 1;
}

template <int WARP, bool MIX>
void launch_fwd_general(const pd::WarpParams& p, bool debug, cudaStream_t st) {
    const int64_t total = (int64_t)p.d.B * p.hw;
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (debug)
        pd::warp_composite_fwd_general<WARP, MIX, true><<<grid, 256, 0, st>>>(p);
    else
        pd::warp_composite_fwd_general<WARP, MIX, false><<<grid, 256, 0, st>>>(p);
}

template <int WARP, bool MIX>
void launch_bwd_general(const pd::WarpParams& p, cudaStream_t st) {
    const int64_t total = (int64_t)p.d.B * p.hw;
    const unsigned grid = (unsigned)((total + 255) / 256);
    pd::warp_composite_bwd_general<WARP, MIX><<<grid, 256, 0, st>>>(p);
}

dim3 loss_grid(const pd_loss_desc* d) {
    return dim3((d->W + pd::LT_W - 1) / pd::LT_W, (d->H + pd::LT_H - 1) / pd::LT_H, d->B);
}

// persistent grid of the elementwise kernels
unsigned ew_grid(int64_t work_items) {
    const int64_t want = (work_items + pd::EW_THREADS - 1) / pd::EW_THREADS;
    const int64_t cap = 148 * 8;
    return (unsigned)(want < cap ? (want < 1 ? 1 : want) : cap);
}

int validate_loss(const pd_loss_desc* d, const pd_loss_in* in) {
    if (!d || !in) return fail(PD_ERR_ARG, "NULL descriptor");
    if (d->B < 1 || d->H < 2 || d->W < 2 || d->B > 65535) return fail(PD_ERR_SHAPE, "1 <= B <= 65535 and H,W >= 2 required");
    if (d->loss_mode < PD_LOSS_L1 || d->loss_mode > PD_LOSS_SSIM_L1) return fail(PD_ERR_ARG, "bad loss_mode %d", d->loss_mode);
    if (d->has_mask_novel && !in->mask_novel) return fail(PD_ERR_ARG, "has_mask_novel set but mask_novel is NULL");
    return PD_OK;
}

int validate_loss_fwd(const pd_loss_desc* d, const pd_loss_in* in) {
    int rc = validate_loss(d, in);
    if (rc) return rc;
    if (!in->rgb_rec || !in->tgt) return fail(PD_ERR_ARG, "rgb_rec / tgt must not be NULL");
    if (d->loss_mode == PD_LOSS_MIXTURE) {
        if (!in->nll || (d->automask && !in->nll_auto)) return fail(PD_ERR_ARG, "mixture loss needs nll (and nll_auto with automask)");
    } else if (d->automask && !in->src) {
        return fail(PD_ERR_ARG, "automask needs src");
    }
    return PD_OK;
}

template <typename K>
void loss_smem_optin(K kern, size_t smem) {
    static std::mutex mu;
    static std::map<const void*, size_t> granted;
    if (smem <= 48 * 1024) return;
    std::lock_guard<std::mutex> lock(mu);
    size_t& g = granted[(const void*)kern];
    if (smem > g) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        g = smem;
    }
}

template <bool AUTO, bool HASMASK, bool WANT_G>
void launch_ssim(const pd::LossParams& p, dim3 g, cudaStream_t st) {
    const size_t smem = pd::ssim_smem_bytes(AUTO, WANT_G);
    auto kern = pd::ssim_l1_fwd_kernel<AUTO, HASMASK, WANT_G>;
    loss_smem_optin(kern, smem);
    kern<<<g, pd::LT_THREADS, smem, st>>>(p);
}

template <bool AUTO, bool HASMASK, bool WANT_G>
void launch_ssim_stream(const pd::LossParams& p, int strips, int segs, int rs, unsigned grid, cudaStream_t st) {
    pd::ssim_l1_stream_kernel<AUTO, HASMASK, WANT_G><<<grid, pd::SW_THREADS, 0, st>>>(p, strips, segs, rs);
}

template <int MODE, bool AUTO, bool HASMASK>
void launch_ew(const pd::LossParams& p, unsigned g, bool want_g, cudaStream_t st) {
    if (want_g) pd::elementwise_fwd_kernel<MODE, AUTO, HASMASK, true><<<g, pd::EW_THREADS, 0, st>>>(p);
    else pd::elementwise_fwd_kernel<MODE, AUTO, HASMASK, false><<<g, pd::EW_THREADS, 0, st>>>(p);
}

__global__ void debug_roundtrip_kernel(const float* __restrict__ u, int64_t n, float size_m1, float rcp, float* __restrict__ exact, float* __restrict__ fast) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    exact[i] = pd::roundtrip(u[i], size_m1);
    fast[i] = (rcp != 0.0f) ? pd::roundtrip_fast(u[i], size_m1, rcp) : pd::roundtrip(u[i], size_m1);
}

}  // namespace

extern "C" {

int pd_version(void) { return PD_ABI_VERSION; }
const char* pd_last_error(void) { return g_err; }
int64_t pd_launch_count(void) { return g_launches.load(); }
void pd_reset_launch_count(void) { g_launches.store(0); }

size_t pd_warp_composite_workspace_bytes(const pd_warp_desc* d) {
    // homography fast path: the source colour packed to one rgbx float4 per pixel (pd_warp_homo.cuh)
    if (!d || d->B < 1 || d->H < 1 || d->W < 1) return 0;
    if (d->warp_type == PD_WARP_HOMOGRAPHY && ((int64_t)d->H * d->W) % pd::hm::HT == 0 && d->W % 32 == 0) return pd::hm::homo_workspace_bytes(d);
    return 0;
}

size_t pd_warp_composite_stats_bytes(const pd_warp_desc* d) {
    if (!d || d->B < 1 || d->H < 1 || d->W < 1) return 0;
    return stats_floats(d) * sizeof(float) + mask_summary_bytes(d);
}

int pd_warp_composite_fwd(const pd_warp_desc* d, const pd_warp_in* in, pd_warp_out* out, void* workspace, pd_stream_t stream) {
    int rc = validate_warp(d, in);
    if (rc) return rc;
    if (!out || !out->rgb_rec || !out->stats) return fail(PD_ERR_ARG, "rgb_rec / stats outputs must not be NULL");
    if (d->mixture && (!out->nll || (d->automask && !out->nll_auto))) return fail(PD_ERR_ARG, "mixture needs nll (and nll_auto with automask)");
    if ((rc = check_device())) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    pd::WarpParams p = make_params(d, in);
    p.out = *out;
    attach_mask_summary(p, out->stats);
    const bool debug = out->rgb_rec_layered || out->logit_rec || out->probability_rec || out->sigma_rec || out->pi_rec;
    const bool streamed = !debug && !exact_coords(d) && pd::ts::stream_path_supported(p);
    if (p.mask_rows) {
        // bit n of a row = "plane n's mask row is not all ones".  The streamed forward ORs bits into a cleared summary;
        // the other forward kernels keep none: all bits set means "read the mask"
        cudaError_t e = cudaMemsetAsync(p.mask_rows, streamed ? 0 : 0xff, mask_summary_bytes(d), st);
        if (e != cudaSuccess) return fail(PD_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    }
    if (streamed && pd::ts::launch_fwd_stream(p, st)) return check_launch("rows_fwd_stream");
    if (streamed && p.mask_rows) {  // no launch configuration after all
        cudaError_t e = cudaMemsetAsync(p.mask_rows, 0xff, mask_summary_bytes(d), st);
        if (e != cudaSuccess) return fail(PD_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    }
    if (!debug && pd::rows_path_supported(p)) {
        pd::launch_fwd_rows(p, st);
        return check_launch("warp_composite_fwd_rows");
    }
    if (!debug && !exact_coords(d) && pd::hm::homo_path_supported(p)) {
        if (!workspace) return fail(PD_ERR_WORKSPACE, "homography warp needs the workspace of pd_warp_composite_workspace_bytes()");
        pd::hm::homo_pack(p, (float4*)workspace, st);
        if ((rc = check_launch("pack_rgbx"))) return rc;
        pd::hm::launch_homo_fwd(p, (const float4*)workspace, st);
        return check_launch("homo_fwd");
    }
    const bool mix = d->mixture != 0;
    switch (d->warp_type) {
        case PD_WARP_DISP: mix ? launch_fwd_general<PD_WARP_DISP, true>(p, debug, st) : launch_fwd_general<PD_WARP_DISP, false>(p, debug, st); break;
        case PD_WARP_HOMOGRAPHY: mix ? launch_fwd_general<PD_WARP_HOMOGRAPHY, true>(p, debug, st) : launch_fwd_general<PD_WARP_HOMOGRAPHY, false>(p, debug, st); break;
        default: mix ? launch_fwd_general<PD_WARP_DEPTH, true>(p, debug, st) : launch_fwd_general<PD_WARP_DEPTH, false>(p, debug, st); break;
    }
    return check_launch("warp_composite_fwd_general");
}

int pd_warp_composite_bwd(const pd_warp_desc* d, const pd_warp_in* in, const pd_warp_out* saved, const pd_warp_grad_out* gout,
                          pd_warp_grad_in* gin, void* workspace, pd_stream_t stream) {
    int rc = validate_warp(d, in);
    if (rc) return rc;
    if (!saved || !saved->rgb_rec || !saved->stats) return fail(PD_ERR_ARG, "saved rgb_rec / stats must not be NULL");
    if (!gout || !gout->g_rgb_rec) return fail(PD_ERR_ARG, "g_rgb_rec must not be NULL");
    if (!gin) return fail(PD_ERR_ARG, "NULL grad_in");
    if ((rc = check_device())) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    pd::WarpParams p = make_params(d, in);
    p.out = *saved;
    attach_mask_summary(p, saved->stats);
    p.gout = *gout;
    p.gin = *gin;
    if (!d->mixture) p.gin.g_sigma = nullptr;
    if (d->warp_type == PD_WARP_HOMOGRAPHY) p.gin.g_disp = nullptr; else p.gin.g_hmat = nullptr;
    const pd_strides4& gs = p.gin.g_disp_stride;
    p.g_disp_dense = p.gin.g_disp && gs.b != 0 && gs.n != 0 && gs.y != 0 && gs.x != 0;

    const size_t plane_bytes = (size_t)d->B * d->N * p.hw * sizeof(float);
    cudaError_t e = cudaSuccess;
    const bool streamed = !exact_coords(d) && pd::ts::stream_path_supported(p) && pd::ts::stream_bwd_fits(p);
    const bool rows = streamed || pd::rows_path_supported(p);
    // scatter targets are accumulated with atomics in the general path: zero them first
    if (!rows) {
        if (p.gin.g_logits) e = cudaMemsetAsync(p.gin.g_logits, 0, plane_bytes, st);
        if (e == cudaSuccess && p.gin.g_sigma) e = cudaMemsetAsync(p.gin.g_sigma, 0, plane_bytes, st);
    }
    if (e == cudaSuccess && p.gin.g_disp && !p.g_disp_dense)
        e = cudaMemsetAsync(p.gin.g_disp, 0, (size_t)strided_extent(gs, d->B, d->N, d->H, d->W) * sizeof(float), st);
    if (e == cudaSuccess && p.gin.g_hmat) e = cudaMemsetAsync(p.gin.g_hmat, 0, (size_t)d->B * d->N * 9 * sizeof(float), st);
    if (e != cudaSuccess) return fail(PD_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    if (streamed) {
        if (!pd::ts::launch_bwd_stream(p, st)) return fail(PD_ERR_SHAPE, "rows_bwd_stream: no launch configuration");
        return check_launch("rows_bwd_stream");
    }
    if (rows) {
        pd::launch_bwd_rows(p, st);
        return check_launch("warp_composite_bwd_rows");
    }
    if (!exact_coords(d) && pd::hm::homo_path_supported(p)) {
        if (!workspace) return fail(PD_ERR_WORKSPACE, "homography warp needs the workspace of pd_warp_composite_workspace_bytes()");
        pd::hm::homo_pack(p, (float4*)workspace, st);
        if ((rc = check_launch("pack_rgbx"))) return rc;
        pd::hm::launch_homo_bwd(p, (const float4*)workspace, st);
        return check_launch("homo_bwd");
    }
    const bool mix = d->mixture != 0;
    switch (d->warp_type) {
        case PD_WARP_DISP: mix ? launch_bwd_general<PD_WARP_DISP, true>(p, st) : launch_bwd_general<PD_WARP_DISP, false>(p, st); break;
        case PD_WARP_HOMOGRAPHY: mix ? launch_bwd_general<PD_WARP_HOMOGRAPHY, true>(p, st) : launch_bwd_general<PD_WARP_HOMOGRAPHY, false>(p, st); break;
        default: mix ? launch_bwd_general<PD_WARP_DEPTH, true>(p, st) : launch_bwd_general<PD_WARP_DEPTH, false>(p, st); break;
    }
    return check_launch("warp_composite_bwd_general");
}

// ---------------------------------------------------------------------------------------------
// decoder tail (networks/depth_decoder.py:258-291)
// ---------------------------------------------------------------------------------------------
namespace {
// pixels per CTA: the [N][T] column cache stays within 64 KB
int tail_threads(int N) { return N <= 64 ? 256 : (N <= 128 ? 128 : (N <= 256 ? 64 : 32)); }

int tail_params(const pd_tail_desc* d, const pd_tail_in* in, pd::tl::TailParams& p) {
    if (!d || !in) return fail(PD_ERR_ARG, "NULL descriptor");
    if (d->B < 1 || d->N < 1 || d->H < 1 || d->W < 1 || d->N > PD_MAX_PLANES) return fail(PD_ERR_SHAPE, "bad B,N,H,W");
    if ((int64_t)d->H * d->W >= (1ll << 31)) return fail(PD_ERR_SHAPE, "H*W too large");
    if (!in->disp_layered) return fail(PD_ERR_ARG, "disp_layered must not be NULL");
    if (d->mask_dtype < PD_MASK_NONE || d->mask_dtype > PD_MASK_U8) return fail(PD_ERR_ARG, "bad mask_dtype");
    memset(&p, 0, sizeof(p));
    p.B = d->B, p.N = d->N, p.H = d->H, p.W = d->W;
    p.mask_dtype = in->mask ? d->mask_dtype : PD_MASK_NONE;
    p.hw = (int64_t)d->H * d->W;
    p.ds = d->disp_stride, p.ms = d->mask_stride;
    p.depth_c = 0.1f * 0.58f * (float)d->W;
    p.raw = in->logits_raw, p.sraw = in->sigma_raw, p.disp_layered = in->disp_layered, p.mask = in->mask;
    p.warp_rows = (d->W % 32 == 0);
    return PD_OK;
}
}  // namespace

int pd_plane_tail_fwd(const pd_tail_desc* d, const pd_tail_in* in, pd_tail_out* out, pd_stream_t stream) {
    pd::tl::TailParams p;
    int rc = tail_params(d, in, p);
    if (rc) return rc;
    if (!in->logits_raw || (d->mixture && !in->sigma_raw)) return fail(PD_ERR_ARG, "logits_raw (and sigma_raw with mixture) must not be NULL");
    if (!out || !out->logits || !out->probability || !out->disp || !out->stats || (d->mixture && !out->sigma))
        return fail(PD_ERR_ARG, "logits / probability / disp / stats (and sigma with mixture) outputs must not be NULL");
    if ((rc = check_device())) return rc;
    p.logits = out->logits, p.sigma = out->sigma, p.prob = out->probability, p.pi = out->pi, p.disp = out->disp, p.depth = out->depth, p.stats = out->stats;
    const int T = tail_threads(d->N);
    const size_t smem = (size_t)d->N * T * sizeof(float);
    const unsigned grid = (unsigned)(((int64_t)d->B * p.hw + T - 1) / T);
    if (d->mixture) {
        loss_smem_optin(pd::tl::tail_fwd_kernel<true>, smem);
        pd::tl::tail_fwd_kernel<true><<<grid, T, smem, (cudaStream_t)stream>>>(p);
    } else {
        loss_smem_optin(pd::tl::tail_fwd_kernel<false>, smem);
        pd::tl::tail_fwd_kernel<false><<<grid, T, smem, (cudaStream_t)stream>>>(p);
    }
    return check_launch("tail_fwd");
}

int pd_plane_tail_bwd(const pd_tail_desc* d, const pd_tail_in* in, const pd_tail_out* saved, const pd_tail_grad_out* gout,
                      pd_tail_grad_in* gin, pd_stream_t stream) {
    pd::tl::TailParams p;
    int rc = tail_params(d, in, p);
    if (rc) return rc;
    if (!saved || !saved->logits || !saved->stats || !saved->disp || (d->mixture && !saved->sigma))
        return fail(PD_ERR_ARG, "saved logits / disp / stats (and sigma with mixture) must not be NULL");
    if (!gout || !gin) return fail(PD_ERR_ARG, "NULL gradient structs");
    if ((rc = check_device())) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    p.logits = saved->logits, p.sigma = saved->sigma, p.disp = saved->disp, p.stats = saved->stats;
    p.g_logits = gout->g_logits, p.g_sigma = gout->g_sigma, p.g_prob = gout->g_probability, p.g_disp = gout->g_disp, p.g_depth = gout->g_depth;
    p.g_raw = gin->g_logits_raw, p.g_sraw = d->mixture ? gin->g_sigma_raw : nullptr, p.g_dl = gin->g_disp_layered, p.gds = gin->g_disp_stride;
    const pd_strides4& gs = p.gds;
    p.g_dl_dense = p.g_dl && gs.b != 0 && gs.n != 0 && gs.y != 0 && gs.x != 0;
    if (p.g_dl && !p.g_dl_dense) {
        cudaError_t e = cudaMemsetAsync(p.g_dl, 0, (size_t)strided_extent(gs, d->B, d->N, d->H, d->W) * sizeof(float), st);
        if (e != cudaSuccess) return fail(PD_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    }
    const int T = tail_threads(d->N);
    const size_t smem = ((size_t)d->N * T + d->N) * sizeof(float);
    // fully compact disparity gradient ([B,N,1,1]): summed per CTA in shared memory, one flush of N atomics per CTA
    // (needs CTAs that do not straddle images)
    if (p.g_dl && !p.g_dl_dense && gs.y == 0 && gs.x == 0 && p.hw % T == 0) p.smem_acc = 1;
    const unsigned grid = (unsigned)(((int64_t)d->B * p.hw + T - 1) / T);
    if (d->mixture) {
        loss_smem_optin(pd::tl::tail_bwd_kernel<true>, smem);
        pd::tl::tail_bwd_kernel<true><<<grid, T, smem, st>>>(p);
    } else {
        loss_smem_optin(pd::tl::tail_bwd_kernel<false>, smem);
        pd::tl::tail_bwd_kernel<false><<<grid, T, smem, st>>>(p);
    }
    return check_launch("tail_bwd");
}

// ---------------------------------------------------------------------------------------------
// smoothness term (layers.py:243-256)
// ---------------------------------------------------------------------------------------------
namespace {
int validate_smooth(const pd_smooth_desc* d, const float* disp, const float* img) {
    if (!d || !disp || !img) return fail(PD_ERR_ARG, "NULL argument");
    if (d->B < 1 || d->H < 2 || d->x0 < 0 || d->W - d->x0 < 2) return fail(PD_ERR_SHAPE, "smoothness needs H >= 2 and W - x0 >= 2");
    return PD_OK;
}
unsigned smooth_grid(int64_t items) {
    const int64_t want = (items + pd::EW_THREADS - 1) / pd::EW_THREADS;
    return (unsigned)(want < 148 * 8 ? (want < 1 ? 1 : want) : 148 * 8);
}
}  // namespace

size_t pd_smooth_loss_workspace_bytes(const pd_smooth_desc* d) {
    (void)d;
    return (size_t)2 * 148 * 8 * sizeof(float);
}

int pd_smooth_loss_fwd(const pd_smooth_desc* d, const float* disp, const float* img, float* loss, void* workspace, pd_stream_t stream) {
    int rc = validate_smooth(d, disp, img);
    if (rc) return rc;
    if (!loss) return fail(PD_ERR_ARG, "loss must not be NULL");
    if (!workspace) return fail(PD_ERR_WORKSPACE, "workspace of pd_smooth_loss_workspace_bytes() required");
    if ((rc = check_device())) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    pd::SmoothParams p;
    memset(&p, 0, sizeof(p));
    p.B = d->B, p.H = d->H, p.W = d->W, p.x0 = d->x0, p.gamma = d->gamma;
    p.disp = disp, p.img = img, p.partials = (float*)workspace, p.out = loss, p.hw = (int64_t)d->H * d->W;
    const int Wc = d->W - d->x0;
    const unsigned g = smooth_grid((int64_t)d->B * d->H * Wc);
    pd::smooth_fwd_kernel<<<g, pd::EW_THREADS, 0, st>>>(p);
    if ((rc = check_launch("smooth_fwd"))) return rc;
    const float inv_nx = 1.0f / ((float)d->B * d->H * (Wc - 1)), inv_ny = 1.0f / ((float)d->B * (d->H - 1) * Wc);
    pd::smooth_reduce_kernel<<<1, 1024, 0, st>>>(p.partials, (int)g, inv_nx, inv_ny, loss);
    return check_launch("smooth_reduce");
}

int pd_smooth_loss_bwd(const pd_smooth_desc* d, const float* disp, const float* img, const float* g_loss, float* g_disp, pd_stream_t stream) {
    int rc = validate_smooth(d, disp, img);
    if (rc) return rc;
    if (!g_loss || !g_disp) return fail(PD_ERR_ARG, "g_loss / g_disp must not be NULL");
    if ((rc = check_device())) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    pd::SmoothParams p;
    memset(&p, 0, sizeof(p));
    p.B = d->B, p.H = d->H, p.W = d->W, p.x0 = d->x0, p.gamma = d->gamma;
    p.disp = disp, p.img = img, p.g_loss = g_loss, p.g_disp = g_disp, p.hw = (int64_t)d->H * d->W;
    const int Wc = d->W - d->x0;
    const float inv_nx = 1.0f / ((float)d->B * d->H * (Wc - 1)), inv_ny = 1.0f / ((float)d->B * (d->H - 1) * Wc);
    pd::smooth_bwd_kernel<<<smooth_grid((int64_t)d->B * p.hw), pd::EW_THREADS, 0, st>>>(p, inv_nx, inv_ny);
    return check_launch("smooth_bwd");
}

// ---------------------------------------------------------------------------------------------
// occlusion masks / post-processed disparity (trainer.py:421-466)
// ---------------------------------------------------------------------------------------------
size_t pd_occlusion_masks_workspace_bytes(const pd_occl_desc* d) {
    if (!d || d->B < 1 || d->N < 1 || d->H < 1 || d->W < 1) return 0;
    return (size_t)d->B * d->N * d->H * d->W * sizeof(float);
}

int pd_occlusion_masks_fwd(const pd_occl_desc* d, const pd_occl_in* in, pd_occl_out* out, void* workspace, pd_stream_t stream) {
    if (!d || !in || !out) return fail(PD_ERR_ARG, "NULL descriptor");
    if (d->B < 1 || d->N < 1 || d->H < 2 || d->W < 2) return fail(PD_ERR_SHAPE, "B,N >= 1 and H,W >= 2 required");
    if ((int64_t)d->H * d->W >= (1ll << 31)) return fail(PD_ERR_SHAPE, "H*W too large");
    if (!in->logits || !in->disp_layered) return fail(PD_ERR_ARG, "logits / disp_layered must not be NULL");
    if (!out->o_l || !out->o_fr) return fail(PD_ERR_ARG, "o_l / o_fr outputs must not be NULL");
    if (out->mask_novel && !in->probability) return fail(PD_ERR_ARG, "mask_novel needs probability");
    if (out->disp_pp && !in->disp) return fail(PD_ERR_ARG, "disp_pp needs disp");
    if (!workspace) return fail(PD_ERR_WORKSPACE, "workspace of pd_occlusion_masks_workspace_bytes() required");
    int rc;
    if ((rc = check_device())) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    pd::oc::OcclParams p;
    p.B = d->B, p.N = d->N, p.H = d->H, p.W = d->W;
    p.hw = (int64_t)d->H * d->W;
    p.ds = d->disp_stride;
    p.wm1 = (float)(d->W - 1), p.hm1 = (float)(d->H - 1);
    const bool exact = (d->flags & PD_FLAG_EXACT_COORDS) || getenv("PD_EXACT_COORDS");
    const unsigned grid = (unsigned)(((int64_t)d->B * p.hw + 255) / 256);
    float* Q = (float*)workspace;
    const float* D = in->disp_layered;
    const int B = d->B;
    if (!exact && d->W <= 2048) {
        // fused per row: warp -> softmax over planes -> warp back -> sum -> clip, nothing parked in HBM
        const size_t smem = (size_t)2 * (d->W + 2 * pd::oc::OC_PAD) * sizeof(float);
        const unsigned rows = (unsigned)(d->B * d->H);
        if (d->W <= 1024) {
            const int threads = ((d->W + 31) / 32) * 32;
            pd::oc::occlusion_row_kernel<false, 1><<<rows, threads, smem, st>>>(p, in->logits, 0, D, 0, +1.0f, B, -1.0f, out->o_l);
            if ((rc = check_launch("occlusion_row"))) return rc;
            pd::oc::occlusion_row_kernel<true, 1><<<rows, threads, smem, st>>>(p, in->logits, B, D, B, -1.0f, 0, +1.0f, out->o_fr);
        } else {
            const int threads = (((d->W + 1) / 2 + 31) / 32) * 32;
            pd::oc::occlusion_row_kernel<false, 2><<<rows, threads, smem, st>>>(p, in->logits, 0, D, 0, +1.0f, B, -1.0f, out->o_l);
            if ((rc = check_launch("occlusion_row"))) return rc;
            pd::oc::occlusion_row_kernel<true, 2><<<rows, threads, smem, st>>>(p, in->logits, B, D, B, -1.0f, 0, +1.0f, out->o_fr);
        }
        if ((rc = check_launch("occlusion_row"))) return rc;
    } else {
        // left logits -> right view -> softmax -> back to the left view
        if (exact) pd::oc::warp_softmax_kernel<true, false><<<grid, 256, 0, st>>>(p, in->logits, 0, D, 0, +1.0f, Q);
        else pd::oc::warp_softmax_kernel<false, false><<<grid, 256, 0, st>>>(p, in->logits, 0, D, 0, +1.0f, Q);
        if ((rc = check_launch("warp_softmax"))) return rc;
        if (exact) pd::oc::warp_sum_kernel<true><<<grid, 256, 0, st>>>(p, Q, 0, D, B, -1.0f, out->o_l);
        else pd::oc::warp_sum_kernel<false><<<grid, 256, 0, st>>>(p, Q, 0, D, B, -1.0f, out->o_l);
        if ((rc = check_launch("warp_sum"))) return rc;
        // flipped half, mirrored back, the other way round
        if (exact) pd::oc::warp_softmax_kernel<true, true><<<grid, 256, 0, st>>>(p, in->logits, B, D, B, -1.0f, Q);
        else pd::oc::warp_softmax_kernel<false, true><<<grid, 256, 0, st>>>(p, in->logits, B, D, B, -1.0f, Q);
        if ((rc = check_launch("warp_softmax"))) return rc;
        if (exact) pd::oc::warp_sum_kernel<true><<<grid, 256, 0, st>>>(p, Q, 0, D, 0, +1.0f, out->o_fr);
        else pd::oc::warp_sum_kernel<false><<<grid, 256, 0, st>>>(p, Q, 0, D, 0, +1.0f, out->o_fr);
        if ((rc = check_launch("warp_sum"))) return rc;
    }
    if (out->mask_novel) {
        if (exact) pd::oc::warp_sum_kernel<true><<<grid, 256, 0, st>>>(p, in->probability, 0, D, 0, +1.0f, out->mask_novel);
        else pd::oc::warp_sum_kernel<false><<<grid, 256, 0, st>>>(p, in->probability, 0, D, 0, +1.0f, out->mask_novel);
        if ((rc = check_launch("warp_sum"))) return rc;
    }
    if (out->disp_pp) {
        pd::oc::disp_pp_kernel<<<grid, 256, 0, st>>>(p, in->disp, out->o_l, out->o_fr, out->disp_pp);
        if ((rc = check_launch("disp_pp"))) return rc;
    }
    return PD_OK;
}

int pd_debug_roundtrip(const float* u, int64_t n, int32_t size, float* out_exact, float* out_fast, pd_stream_t stream) {
    if (!u || !out_exact || !out_fast || n < 1 || size < 2) return fail(PD_ERR_ARG, "bad arguments");
    debug_roundtrip_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(u, n, (float)(size - 1), pd::rows_rcp(size), out_exact, out_fast);
    return check_launch("debug_roundtrip");
}

// ---------------------------------------------------------------------------------------------
// photometric term
// ---------------------------------------------------------------------------------------------
size_t pd_photometric_workspace_bytes(const pd_loss_desc* d) {
    if (!d) return 0;
    dim3 g = loss_grid(d);
    // one partial per CTA of whichever forward kernel runs: SSIM tiles / streamed SSIM warps (<= one CTA per 8 warps
    // of at least 8 rows x 28 columns, fewer than the 8x64 tiles) / persistent elementwise grid
    const size_t tiles = (size_t)g.x * g.y * g.z, ew = 148 * 8;
    const size_t stream = ((size_t)d->B * ((d->W + pd::SW_COLS - 1) / pd::SW_COLS) * ((d->H + 7) / 8) + pd::SW_WARPS - 1) / pd::SW_WARPS;
    size_t n = tiles > ew ? tiles : ew;
    if (stream > n) n = stream;
    return n * sizeof(float);
}

int pd_photometric_fwd(const pd_loss_desc* d, const pd_loss_in* in, pd_loss_out* out, void* workspace, pd_stream_t stream) {
    int rc = validate_loss_fwd(d, in);
    if (rc) return rc;
    if (!out || !out->ph_sum) return fail(PD_ERR_ARG, "ph_sum must not be NULL");
    if (d->has_mask_novel && !out->pred) return fail(PD_ERR_ARG, "has_mask_novel needs the pred output");
    if (!workspace) return fail(PD_ERR_WORKSPACE, "workspace of pd_photometric_workspace_bytes() required");
    if ((rc = check_device())) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    pd::LossParams p;
    memset(&p, 0, sizeof(p));
    p.d = *d; p.in = *in; p.out = *out; p.partials = (float*)workspace; p.hw = (int64_t)d->H * d->W;
    const bool a = d->automask != 0, m = d->has_mask_novel != 0;
    int64_t nparts;
    if (d->loss_mode == PD_LOSS_SSIM_L1) {
        const bool wg = out->g_unit != nullptr;
        if (!getenv("PD_SSIM_TILES")) {
            const int rs = pd::ssim_stream_rows(d->B, d->H, d->W);
            const int strips = (d->W + pd::SW_COLS - 1) / pd::SW_COLS, segs = (d->H + rs - 1) / rs;
            const int64_t tasks = (int64_t)d->B * strips * segs;
            const unsigned g = (unsigned)((tasks + pd::SW_WARPS - 1) / pd::SW_WARPS);
            nparts = g;
            if (a) { if (m) { wg ? launch_ssim_stream<true, true, true>(p, strips, segs, rs, g, st) : launch_ssim_stream<true, true, false>(p, strips, segs, rs, g, st); }
                     else   { wg ? launch_ssim_stream<true, false, true>(p, strips, segs, rs, g, st) : launch_ssim_stream<true, false, false>(p, strips, segs, rs, g, st); } }
            else   { if (m) { wg ? launch_ssim_stream<false, true, true>(p, strips, segs, rs, g, st) : launch_ssim_stream<false, true, false>(p, strips, segs, rs, g, st); }
                     else   { wg ? launch_ssim_stream<false, false, true>(p, strips, segs, rs, g, st) : launch_ssim_stream<false, false, false>(p, strips, segs, rs, g, st); } }
            if ((rc = check_launch("ssim_l1_stream"))) return rc;
            pd::reduce_partials_kernel<<<1, 1024, 0, st>>>(p.partials, nparts, out->ph_sum, d->out_scale != 0.0f ? d->out_scale : 1.0f);
            return check_launch("reduce_partials");
        }
        const dim3 g = loss_grid(d);
        nparts = (int64_t)g.x * g.y * g.z;
        if (a) { if (m) { wg ? launch_ssim<true, true, true>(p, g, st) : launch_ssim<true, true, false>(p, g, st); }
                 else   { wg ? launch_ssim<true, false, true>(p, g, st) : launch_ssim<true, false, false>(p, g, st); } }
        else   { if (m) { wg ? launch_ssim<false, true, true>(p, g, st) : launch_ssim<false, true, false>(p, g, st); }
                 else   { wg ? launch_ssim<false, false, true>(p, g, st) : launch_ssim<false, false, false>(p, g, st); } }
    } else {
        const unsigned g = ew_grid((int64_t)d->B * p.hw);
        nparts = g;
        if (d->loss_mode == PD_LOSS_MIXTURE) {
            const bool wg = out->g_unit_nll != nullptr;
            if (a) { m ? launch_ew<PD_LOSS_MIXTURE, true, true>(p, g, wg, st) : launch_ew<PD_LOSS_MIXTURE, true, false>(p, g, wg, st); }
            else   { m ? launch_ew<PD_LOSS_MIXTURE, false, true>(p, g, wg, st) : launch_ew<PD_LOSS_MIXTURE, false, false>(p, g, wg, st); }
        } else {
            const bool wg = out->g_unit != nullptr;
            if (a) { m ? launch_ew<PD_LOSS_L1, true, true>(p, g, wg, st) : launch_ew<PD_LOSS_L1, true, false>(p, g, wg, st); }
            else   { m ? launch_ew<PD_LOSS_L1, false, true>(p, g, wg, st) : launch_ew<PD_LOSS_L1, false, false>(p, g, wg, st); }
        }
    }
    if ((rc = check_launch("photometric_fwd"))) return rc;
    pd::reduce_partials_kernel<<<1, 1024, 0, st>>>(p.partials, nparts, out->ph_sum, d->out_scale != 0.0f ? d->out_scale : 1.0f);
    return check_launch("reduce_partials");
}

int pd_photometric_bwd(const pd_loss_desc* d, const pd_loss_in* in, const pd_loss_out* saved, const pd_loss_grad_out* gout,
                       pd_loss_grad_in* gin, void* workspace, pd_stream_t stream) {
    (void)workspace;
    int rc = validate_loss(d, in);
    if (rc) return rc;
    if (!gout || !gout->g_ph_sum) return fail(PD_ERR_ARG, "g_ph_sum must not be NULL");
    if (!gin || !gin->g_rgb_rec) return fail(PD_ERR_ARG, "g_rgb_rec must not be NULL");
    const bool mix = d->loss_mode == PD_LOSS_MIXTURE;
    if (!saved || (mix ? !saved->g_unit_nll : !saved->g_unit)) return fail(PD_ERR_ARG, "the unit gradient saved by pd_photometric_fwd is required");
    if (mix && !gin->g_nll) return fail(PD_ERR_ARG, "mixture loss needs g_nll");
    if ((rc = check_device())) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    pd::LossParams p;
    memset(&p, 0, sizeof(p));
    p.d = *d; p.in = *in; p.out = *saved; p.gout = *gout; p.gin = *gin; p.hw = (int64_t)d->H * d->W;
    const bool m = d->has_mask_novel != 0;
    const void* ptrs[] = {in->mask_novel, saved->g_unit, saved->g_unit_nll, gout->g_pred, gin->g_rgb_rec, gin->g_nll};
    bool v4 = (p.hw % 4 == 0);
    for (const void* q : ptrs) v4 = v4 && (!q || (reinterpret_cast<uintptr_t>(q) & 15) == 0);
    if (v4) {
        p.total4 = (int64_t)d->B * p.hw / 4;
        const unsigned g = ew_grid(p.total4);
        if (mix) { m ? pd::photometric_bwd_kernel_v4<true, true><<<g, pd::EW_THREADS, 0, st>>>(p) : pd::photometric_bwd_kernel_v4<true, false><<<g, pd::EW_THREADS, 0, st>>>(p); }
        else     { m ? pd::photometric_bwd_kernel_v4<false, true><<<g, pd::EW_THREADS, 0, st>>>(p) : pd::photometric_bwd_kernel_v4<false, false><<<g, pd::EW_THREADS, 0, st>>>(p); }
    } else {
        const unsigned g = ew_grid((int64_t)d->B * p.hw);
        if (mix) { m ? pd::photometric_bwd_kernel<true, true><<<g, pd::EW_THREADS, 0, st>>>(p) : pd::photometric_bwd_kernel<true, false><<<g, pd::EW_THREADS, 0, st>>>(p); }
        else     { m ? pd::photometric_bwd_kernel<false, true><<<g, pd::EW_THREADS, 0, st>>>(p) : pd::photometric_bwd_kernel<false, false><<<g, pd::EW_THREADS, 0, st>>>(p); }
    }
    return check_launch("photometric_bwd");
}

}  // extern "C"
