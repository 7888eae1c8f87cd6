// Translation unit: decoder tail (pd_plane_tail_*) and occlusion masks / post-processed disparity (pd_occlusion_masks_fwd).
#include <string.h>

#include "pd_occlusion.cuh"
#include "pd_tail_tile.cuh"

using pd::check_device;
using pd::check_launch;
using pd::fail;

namespace {
int64_t strided_extent(const pd_strides4& s, int B, int N, int H, int W) {
    return (int64_t)(B - 1) * s.b + (int64_t)(N - 1) * s.n + (int64_t)(H - 1) * s.y + (int64_t)(W - 1) * s.x + 1;
}
template <typename K>
void loss_smem_optin(K kern, size_t smem) {
    pd::smem_optin((const void*)kern, smem);
}
}  // namespace

extern "C" {

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
    pd::tl::TileCfg tc;
    if (!pd::tuning().tail_direct && pd::tl::tile_cfg(p, d->mixture != 0, 2, tc) && pd::tl::tile_ptrs_ok(p)) {
        // TMA-tile kernel: one CTA per (row, column tile), every plane row of the tile in flight at once
        const unsigned tgrid = (unsigned)((int64_t)d->B * d->H * tc.tiles);
        if (d->mixture) {
            loss_smem_optin(pd::tl::tail_fwd_tile_kernel<true>, tc.smem);
            pd::tl::tail_fwd_tile_kernel<true><<<tgrid, tc.threads, tc.smem, (cudaStream_t)stream>>>(p, tc.tw, tc.tiles);
        } else {
            loss_smem_optin(pd::tl::tail_fwd_tile_kernel<false>, tc.smem);
            pd::tl::tail_fwd_tile_kernel<false><<<tgrid, tc.threads, tc.smem, (cudaStream_t)stream>>>(p, tc.tw, tc.tiles);
        }
        return check_launch("tail_fwd_tile");
    }
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
    pd::tl::TileCfg tc;
    if (!pd::tuning().tail_direct && pd::tl::tile_cfg(p, false, pd::tl::TT_BWD_HDR, tc) && pd::tl::tile_ptrs_ok(p)) {
        const unsigned tgrid = (unsigned)((int64_t)d->B * d->H * tc.tiles);
        if (d->mixture) {
            loss_smem_optin(pd::tl::tail_bwd_tile_kernel<true>, tc.smem);
            pd::tl::tail_bwd_tile_kernel<true><<<tgrid, tc.threads, tc.smem, st>>>(p, tc.tw, tc.tiles);
        } else {
            loss_smem_optin(pd::tl::tail_bwd_tile_kernel<false>, tc.smem);
            pd::tl::tail_bwd_tile_kernel<false><<<tgrid, tc.threads, tc.smem, st>>>(p, tc.tw, tc.tiles);
        }
        return check_launch("tail_bwd_tile");
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
    const bool exact = (d->flags & PD_FLAG_EXACT_COORDS) != 0;
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

}  // extern "C"
