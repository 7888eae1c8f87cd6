// Translation unit: photometric term (pd_photometric_*) and smoothness term (pd_smooth_loss_*) of compute_losses.
#include <string.h>

#include "pd_loss.cuh"

using pd::check_device;
using pd::check_launch;
using pd::fail;

namespace {

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
    pd::smem_optin((const void*)kern, smem);
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

}  // namespace

extern "C" {

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
        if (!pd::tuning().ssim_tiles) {
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
