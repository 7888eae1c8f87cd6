// C ABI of planedepth_b200 (see include/planedepth_b200.h): shared host state (error string, launch counter, tuning
// block, attribute caches), argument validation and kernel selection of the warp + composite entry points.  The kernel
// families live in their own translation units (pd_tu_*.cu) behind pd_common.h.  No torch / ATen dependency;
// everything is enqueued on the caller's stream.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <tuple>

#include "pd_warp_general.cuh"  // WarpParams (its kernel templates are instantiated in pd_tu_general.cu)

namespace pd {

namespace {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};  // process-wide: autograd runs backward on its own threads

constexpr int kMaxDevices = 64;
std::atomic<int> g_dev_major[kMaxDevices];  // 0 = not queried yet
std::atomic<int> g_dev_sms[kMaxDevices];

std::mutex g_cache_mu;
std::map<const void*, size_t> g_smem_granted;
std::map<std::tuple<const void*, int, size_t>, int> g_resident;

int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

pd_tuning sanitize(pd_tuning t) {
    t.stream_ctas_per_sm = clampi(t.stream_ctas_per_sm, 0, 32);
    t.stream_hs = clampi(t.stream_hs, 0, PD_MAX_PLANES);
    t.stream_nst = clampi(t.stream_nst, 0, 8);
    t.stream_smem_kb = clampi(t.stream_smem_kb, 0, 220);
    t.stream_px8 = t.stream_px8 != 0;
    t.ssim_tiles = t.ssim_tiles != 0;
    t.homo_tiles = clampi(t.homo_tiles, -1, 1);
    t.stream_fwd_minb = (t.stream_fwd_minb == 5 || t.stream_fwd_minb == 6) ? t.stream_fwd_minb : 0;
    t.stream_no_l2_hint = t.stream_no_l2_hint != 0;
    t.tail_direct = t.tail_direct != 0;
    t.stream_bwd_minb = (t.stream_bwd_minb == 3 || t.stream_bwd_minb == 4) ? t.stream_bwd_minb : 0;
    return t;
}

int env_int(const char* name) {
    const char* v = getenv(name);
    return v ? atoi(v) : 0;
}

// read ONCE, at library load
pd_tuning tuning_from_env() {
    pd_tuning t;
    memset(&t, 0, sizeof(t));
    t.stream_ctas_per_sm = env_int("PD_STREAM_CTAS");
    t.stream_hs = env_int("PD_STREAM_HS");
    t.stream_nst = env_int("PD_STREAM_NST");
    t.stream_smem_kb = env_int("PD_STREAM_SMEM_KB");
    t.stream_px8 = env_int("PD_STREAM_PX8");
    t.ssim_tiles = env_int("PD_SSIM_TILES");
    t.homo_tiles = env_int("PD_HOMO_TILES");
    t.stream_fwd_minb = env_int("PD_STREAM_FWD_MINB");
    t.stream_no_l2_hint = env_int("PD_STREAM_NO_L2_HINT");
    t.tail_direct = env_int("PD_TAIL_DIRECT");
    t.stream_bwd_minb = env_int("PD_STREAM_BWD_MINB");
    return sanitize(t);
}

const pd_tuning g_env_tuning = tuning_from_env();
pd_tuning g_tuning = g_env_tuning;

int current_device(int* dev) {
    cudaError_t e = cudaGetDevice(dev);
    if (e != cudaSuccess) return fail(PD_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
    if (*dev < 0 || *dev >= kMaxDevices) return fail(PD_ERR_CUDA, "device ordinal %d out of range", *dev);
    return PD_OK;
}
}  // namespace

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

int check_device() {
    int dev = 0, rc = current_device(&dev);
    if (rc) return rc;
    int major = g_dev_major[dev].load(std::memory_order_relaxed);
    if (major == 0) {
        cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        if (e != cudaSuccess) return fail(PD_ERR_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
        g_dev_major[dev].store(major, std::memory_order_relaxed);
    }
    if (major != 10) return fail(PD_ERR_ARCH, "planedepth_b200 is built for sm_100a only (device is sm_%d*)", major);
    return PD_OK;
}

int sm_count() {
    int dev = 0;
    if (current_device(&dev)) return 148;
    int sms = g_dev_sms[dev].load(std::memory_order_relaxed);
    if (sms == 0) {
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
        g_dev_sms[dev].store(sms, std::memory_order_relaxed);
    }
    return sms;
}

const pd_tuning& tuning() { return g_tuning; }

void smem_optin(const void* kernel, size_t smem) {
    if (smem <= 48 * 1024) return;
    std::lock_guard<std::mutex> lock(g_cache_mu);
    size_t& g = g_smem_granted[kernel];
    if (smem > g) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        g = smem;
    }
}

int resident_ctas(const void* kernel, int threads, size_t smem) {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    auto key = std::make_tuple(kernel, threads, smem);
    auto it = g_resident.find(key);
    if (it != g_resident.end()) return it->second;
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    g_resident[key] = per_sm;
    return per_sm;
}

}  // namespace pd

namespace {

using pd::check_device;
using pd::check_launch;
using pd::fail;
namespace api = pd::api;

bool exact_coords(const pd_warp_desc* d) { return (d->flags & PD_FLAG_EXACT_COORDS) != 0; }

int validate_warp(const pd_warp_desc* d, const pd_warp_in* in) {
    if (!d || !in) return fail(PD_ERR_ARG, "NULL descriptor");
    if (d->B < 1 || d->N < 1 || d->H < 2 || d->W < 2) return fail(PD_ERR_SHAPE, "B,N >= 1 and H,W >= 2 required (got %d,%d,%d,%d)", d->B, d->N, d->H, d->W);
    if (d->N > PD_MAX_PLANES) return fail(PD_ERR_SHAPE, "N=%d exceeds PD_MAX_PLANES", d->N);
    if ((int64_t)d->H * d->W >= (1ll << 31)) return fail(PD_ERR_SHAPE, "H*W too large");
    if (d->warp_type < PD_WARP_DISP || d->warp_type > PD_WARP_DEPTH) return fail(PD_ERR_ARG, "bad warp_type %d", d->warp_type);
    if (d->dtype != PD_DTYPE_F32 && d->dtype != PD_DTYPE_BF16) return fail(PD_ERR_ARG, "bad dtype %d", d->dtype);
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

// The summary SLOT exists whenever the descriptor's shape allows one; both passes derive that from the descriptor alone.
// Whether its CONTENT may fold mask rows away is written into the slot by the forward pass (all bits set = "read every
// mask row"), so a backward call never trusts bits that no forward kernel produced.
unsigned long long* mask_summary_slot(const pd::WarpParams& p, const float* stats) {
    const pd_warp_desc& d = p.d;
    const bool has = d.warp_type == PD_WARP_DISP && p.in.mask && d.mask_dtype == PD_MASK_F32 && d.mask_stride.x == 1 && d.N <= 64 &&
                     (stats_floats(&d) % 2 == 0);
    return has ? reinterpret_cast<unsigned long long*>(const_cast<float*>(stats) + stats_floats(&d)) : nullptr;
}

int64_t strided_extent(const pd_strides4& s, int B, int N, int H, int W) {
    return (int64_t)(B - 1) * s.b + (int64_t)(N - 1) * s.n + (int64_t)(H - 1) * s.y + (int64_t)(W - 1) * s.x + 1;
}

__global__ void debug_roundtrip_kernel(const float* __restrict__ u, int64_t n, float size_m1, float rcp, float* __restrict__ exact, float* __restrict__ fast) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    exact[i] = pd::roundtrip(u[i], size_m1);
    fast[i] = (rcp != 0.0f) ? pd::roundtrip_fast(u[i], size_m1, rcp) : pd::roundtrip(u[i], size_m1);
}

// pd_x_constant_check: one warp per (b, n, y) row; a row whose elements are not all equal to its column 0 counts once
template <typename T>
__global__ void __launch_bounds__(256) x_constant_kernel(const T* __restrict__ data, pd_strides4 s, int B, int N, int H, int W, int32_t* __restrict__ violations) {
    const int64_t rows = (int64_t)B * N * H;
    const int lane = threadIdx.x & 31;
    int bad_rows = 0;
    for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int y = (int)(row % H);
        const int64_t bn = row / H;
        const int n = (int)(bn % N), b = (int)(bn / N);
        const T* q = data + (int64_t)b * s.b + (int64_t)n * s.n + (int64_t)y * s.y;
        const T first = q[0];
        bool bad = false;
        for (int x = lane; x < W; x += 32) bad |= !(q[(int64_t)x * s.x] == first);  // NaN != NaN counts as varying
        if (__any_sync(0xffffffffu, bad)) ++bad_rows;
    }
    if (lane == 0 && bad_rows) atomicAdd(violations, bad_rows);
}

}  // namespace

extern "C" {

int pd_version(void) { return PD_ABI_VERSION; }
const char* pd_last_error(void) { return pd::g_err; }
int64_t pd_launch_count(void) { return pd::g_launches.load(); }
void pd_reset_launch_count(void) { pd::g_launches.store(0); }

void pd_get_tuning(pd_tuning* out) {
    if (out) *out = pd::g_tuning;
}
void pd_set_tuning(const pd_tuning* in) { pd::g_tuning = in ? pd::sanitize(*in) : pd::g_env_tuning; }

size_t pd_warp_composite_workspace_bytes(const pd_warp_desc* d) {
    // homography fast paths: the source colour packed to one rgbx float4 per pixel (pd_warp_homo.cuh)
    if (!d || d->B < 1 || d->H < 1 || d->W < 1) return 0;
    return api::homo_workspace_bytes(d);
}

size_t pd_warp_composite_stats_bytes(const pd_warp_desc* d) {
    if (!d || d->B < 1 || d->H < 1 || d->W < 1) return 0;
    return stats_floats(d) * sizeof(float) + mask_summary_bytes(d);
}

int pd_warp_composite_supports(const pd_warp_desc* d, const pd_warp_in* in) {
    if (validate_warp(d, in)) return 0;
    if (d->dtype == PD_DTYPE_F32) return 1;
    // bf16: the streamed forward AND backward must find a configuration (dry runs; mask / stride / alignment rules included)
    pd::WarpParams p = make_params(d, in);
    if (exact_coords(d) || !api::stream_supported(p)) return 0;
    return (api::stream_fwd_fits(p) && api::stream_bwd_fits(p)) ? 1 : 0;
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
    unsigned long long* slot = mask_summary_slot(p, out->stats);
    const bool debug = out->rgb_rec_layered || out->logit_rec || out->probability_rec || out->sigma_rec || out->pi_rec;
    const bool streamed = !debug && !exact_coords(d) && api::stream_supported(p);
    if (d->dtype == PD_DTYPE_BF16) {
        // bf16 storage exists in the streamed stereo kernels only; nothing else may reinterpret the bf16 arrays as fp32
        if (streamed && api::stream_fwd(p, st)) return check_launch("rows_fwd_stream_bf16");
        return fail(PD_ERR_UNSUPPORTED, "bf16 storage is served by the streamed stereo kernels only (see pd_warp_composite_supports)");
    }
    // bit n of a row = "plane n's mask row is not all ones".  Only the streamed forward produces the summary (it ORs bits
    // into a cleared slot); in every other case the slot says "read every mask row"
    const bool summarise = streamed && slot && !(d->flags & PD_FLAG_NO_MASK_SUMMARY);
    p.mask_rows = summarise ? slot : nullptr;
    if (slot) {
        cudaError_t e = cudaMemsetAsync(slot, summarise ? 0 : 0xff, mask_summary_bytes(d), st);
        if (e != cudaSuccess) return fail(PD_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    }
    if (streamed && api::stream_fwd(p, st)) return check_launch("rows_fwd_stream");
    if (summarise) {  // no launch configuration after all: nobody writes the summary
        p.mask_rows = nullptr;
        cudaError_t e = cudaMemsetAsync(slot, 0xff, mask_summary_bytes(d), st);
        if (e != cudaSuccess) return fail(PD_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    }
    if (!debug && api::rows_supported(p)) {
        api::rows_fwd(p, st);
        return check_launch("warp_composite_fwd_rows");
    }
    if (!debug && !exact_coords(d) && api::homo_supported(p)) return api::homo_fwd(p, workspace, st);
    api::general_fwd(p, debug, st);
    return check_launch("warp_composite_fwd_general");
}

int pd_warp_composite_bwd(const pd_warp_desc* d, const pd_warp_in* in, const pd_warp_out* saved, const pd_warp_grad_out* gout,
                          pd_warp_grad_in* gin, void* workspace, pd_stream_t stream) {
    int rc = validate_warp(d, in);
    if (rc) return rc;
    if (!saved || !saved->rgb_rec || !saved->stats) return fail(PD_ERR_ARG, "saved rgb_rec / stats must not be NULL");
    if (!gout || (!gout->g_rgb_rec && !gout->g_ph_sum)) return fail(PD_ERR_ARG, "g_rgb_rec must not be NULL (unless the fused form g_ph_sum is given)");
    if (!gin) return fail(PD_ERR_ARG, "NULL grad_in");
    if ((rc = check_device())) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    pd::WarpParams p = make_params(d, in);
    p.out = *saved;
    p.mask_rows = mask_summary_slot(p, saved->stats);  // content written by the forward pass (all ones = read everything)
    p.gout = *gout;
    if (!p.gout.g_ph_sum) p.gout.g_unit = p.gout.g_unit_nll = p.gout.g_pred = p.gout.mask_novel = nullptr;
    if (!d->mixture) p.gout.g_nll = p.gout.g_unit_nll = nullptr;
    p.gin = *gin;
    if (!d->mixture) p.gin.g_sigma = nullptr;
    if (d->warp_type == PD_WARP_HOMOGRAPHY) p.gin.g_disp = nullptr; else p.gin.g_hmat = nullptr;
    const pd_strides4& gs = p.gin.g_disp_stride;
    p.g_disp_dense = p.gin.g_disp && gs.b != 0 && gs.n != 0 && gs.y != 0 && gs.x != 0;

    const size_t plane_bytes = (size_t)d->B * d->N * p.hw * sizeof(float);
    cudaError_t e = cudaSuccess;
    const bool streamed = !exact_coords(d) && api::stream_supported(p) && api::stream_bwd_fits(p);
    if (d->dtype == PD_DTYPE_BF16 && !streamed) return fail(PD_ERR_UNSUPPORTED, "bf16 storage is served by the streamed stereo kernels only (see pd_warp_composite_supports)");
    const bool rows = !streamed && api::rows_supported(p);
    const bool accumulate = (d->flags & PD_FLAG_ACCUMULATE) != 0;
    if (accumulate && (streamed || rows))
        return fail(PD_ERR_UNSUPPORTED, "PD_FLAG_ACCUMULATE: this descriptor is served by kernels that write their gradient rows");
    // scatter targets are accumulated with atomics in the general / homography paths: zero them first (unless the caller
    // accumulates several target sides into buffers it zeroed itself)
    if (!streamed && !rows && !accumulate) {
        if (p.gin.g_logits) e = cudaMemsetAsync(p.gin.g_logits, 0, plane_bytes, st);
        if (e == cudaSuccess && p.gin.g_sigma) e = cudaMemsetAsync(p.gin.g_sigma, 0, plane_bytes, st);
    }
    if (e == cudaSuccess && p.gin.g_disp && !p.g_disp_dense)
        e = cudaMemsetAsync(p.gin.g_disp, 0, (size_t)strided_extent(gs, d->B, d->N, d->H, d->W) * sizeof(float), st);
    if (e == cudaSuccess && p.gin.g_hmat) e = cudaMemsetAsync(p.gin.g_hmat, 0, (size_t)d->B * d->N * 9 * sizeof(float), st);
    if (e != cudaSuccess) return fail(PD_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    if (streamed) {
        // stream_bwd_fits() is a dry run of the same launcher, so this cannot come back empty-handed
        if (!api::stream_bwd(p, st)) return fail(PD_ERR_CUDA, "rows_bwd_stream: launch configuration vanished between the dry run and the launch");
        return check_launch("rows_bwd_stream");
    }
    if (rows) {
        api::rows_bwd(p, st);
        return check_launch("warp_composite_bwd_rows");
    }
    if (!exact_coords(d) && api::homo_supported(p)) return api::homo_bwd(p, workspace, st);
    api::general_bwd(p, st);
    return check_launch("warp_composite_bwd_general");
}

int pd_x_constant_check(const void* data, int32_t dtype, const pd_strides4* s, int32_t B, int32_t N, int32_t H, int32_t W, int32_t* violations,
                        pd_stream_t stream) {
    if (!data || !s || !violations) return fail(PD_ERR_ARG, "NULL argument");
    if (B < 1 || N < 1 || H < 1 || W < 1) return fail(PD_ERR_SHAPE, "B,N,H,W >= 1 required");
    if (dtype != PD_MASK_F32 && dtype != PD_MASK_U8) return fail(PD_ERR_ARG, "dtype must be PD_MASK_F32 or PD_MASK_U8");
    int rc;
    if ((rc = check_device())) return rc;
    const int64_t rows = (int64_t)B * N * H;
    const int64_t want = (rows + 7) / 8, cap = (int64_t)pd::sm_count() * 8;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (dtype == PD_MASK_F32) x_constant_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)data, *s, B, N, H, W, violations);
    else x_constant_kernel<unsigned char><<<grid, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)data, *s, B, N, H, W, violations);
    return check_launch("x_constant_check");
}

int pd_debug_roundtrip(const float* u, int64_t n, int32_t size, float* out_exact, float* out_fast, pd_stream_t stream) {
    if (!u || !out_exact || !out_fast || n < 1 || size < 2) return fail(PD_ERR_ARG, "bad arguments");
    debug_roundtrip_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(u, n, (float)(size - 1), pd::rows_rcp(size), out_exact, out_fast);
    return check_launch("debug_roundtrip");
}

}  // extern "C"
