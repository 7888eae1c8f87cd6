/*
 * planedepth_b200 — C ABI of the B200-native photometric-reconstruction path.
 *
 * The reference (svip-lab/PlaneDepth) is pure Python on PyTorch and has NO plugin / FFI surface; its
 * seam for this path is two Python methods plus a dict contract:
 *     Trainer.pred_novel_images(inputs, outputs)   /root/reference/trainer.py:523-603
 *     Trainer.compute_losses(inputs, outputs)      /root/reference/trainer.py:701-773
 * This header is the boundary a binding for that seam talks to (INTEGRATION.md shows the ctypes stub
 * and the two-line patch to the reference's trainer.py).  Plain pointers and sizes only: no torch /
 * ATen / pybind types cross this line.  Each entry point cites the reference code it replaces.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to fp32 unless stated otherwise; tensors are NCHW-contiguous
 *    unless a pd_strides4 says otherwise (element strides, 0 = broadcast along that dimension);
 *  - the caller owns every buffer (incl. workspace); the library allocates nothing persistent and
 *    keeps no global state apart from a thread-local error string, small caches of device / kernel
 *    attributes and the explicit tuning block below (pd_set_tuning; tests and benchmarks only);
 *  - all work is enqueued on the given stream; no entry point synchronises the host;
 *  - return value: PD_OK (0) or a pd_status error code; never aborts, never throws;
 *    pd_last_error() returns the message of the last failing call on this thread;
 *  - outputs are fully overwritten by the call (the library zero-fills scatter targets itself).
 */
#ifndef PLANEDEPTH_B200_H
#define PLANEDEPTH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PD_ABI_VERSION 9

typedef void* pd_stream_t; /* a cudaStream_t */

typedef enum pd_status {
    PD_OK = 0,
    PD_ERR_ARG = 1,       /* NULL / inconsistent argument */
    PD_ERR_SHAPE = 2,     /* unsupported shape (B,N,H,W must be >=1; N <= PD_MAX_PLANES) */
    PD_ERR_ALIGN = 3,     /* pointer / stride alignment */
    PD_ERR_ARCH = 4,      /* device is not sm_100 */
    PD_ERR_CUDA = 5,      /* a CUDA runtime call failed (message in pd_last_error) */
    PD_ERR_WORKSPACE = 6, /* workspace NULL while pd_*_workspace_bytes() > 0 */
    PD_ERR_UNSUPPORTED = 7 /* valid request no kernel is built for (bf16 storage outside the streamed stereo path) */
} pd_status;

#define PD_MAX_PLANES 256

/* trainer.py:533 / 540 / 556 — options.py --warp_type */
typedef enum pd_warp_type { PD_WARP_DISP = 0, PD_WARP_HOMOGRAPHY = 1, PD_WARP_DEPTH = 2 } pd_warp_type;

/* photometric term: trainer.py:738-742 (L1), 729-736 (Laplacian mixture NLL, layers.py:454-466),
 * 687-699 + layers.py:276-306 (0.85 SSIM + 0.15 L1 — BASELINE.json north_star) */
typedef enum pd_loss_mode { PD_LOSS_L1 = 0, PD_LOSS_MIXTURE = 1, PD_LOSS_SSIM_L1 = 2 } pd_loss_mode;

typedef enum pd_mask_dtype { PD_MASK_NONE = 0, PD_MASK_F32 = 1, PD_MASK_U8 = 2 } pd_mask_dtype;

/* pd_warp_desc.flags.
 * PD_FLAG_EXACT_COORDS: reproduce, bit for bit, the fp32 rounding of the reference's coordinate
 * normalise / un-normalise round trip (trainer.py:549-551 + ATen grid_sampler_unnormalize), which moves
 * sample positions by a few ulp of the coordinate (<= 1.2e-4 px at W = 1280).  Without the flag the
 * stereo disparity fast path samples at the exact positions u = x + sign*d, v = y (deviations);
 * homography / depth warps and dense disparities always use the reference's arithmetic. */
/* PD_FLAG_NO_MASK_SUMMARY: the streamed forward does not summarise a dense padding_mask per row (the saved
 * statistics then say "read every mask row" and the backward streams the mask again); a measurement knob. */
/* PD_FLAG_ACCUMULATE (pd_warp_composite_bwd only): g_logits / g_sigma are ADDED to instead of being zero-filled first, so that
 * the target sides of one step (trainer.py:528: for side in self.target_sides) scatter into one caller-zeroed buffer -- what
 * autograd otherwise does with one [B,N,H,W] add per extra side.  Only the scatter kernels can do that (homography / depth
 * warps, x-varying disparity); a descriptor served by kernels that WRITE their gradient rows (streamed / bit-faithful stereo
 * kernels) answers PD_ERR_UNSUPPORTED. */
/* PD_FLAG_WORKSPACE_READY: the workspace already holds what an earlier call with the same `src` and shape left there (the
 * homography fast path's packed rgbx copy of the source image), so the call does not rebuild it: the target sides of a step
 * share the source frame (trainer.py:567). */
typedef enum pd_warp_flags { PD_FLAG_EXACT_COORDS = 1, PD_FLAG_NO_MASK_SUMMARY = 2, PD_FLAG_ACCUMULATE = 4, PD_FLAG_WORKSPACE_READY = 8 } pd_warp_flags;

/* Storage type of the [B,N,H,W] network outputs and of their gradients (pd_warp_desc.dtype): logits, sigma, g_logits,
 * g_sigma.  Everything else (colours, rgb_rec, statistics, nll, plane geometry, masks) is fp32 whatever this says, and all
 * arithmetic is fp32.  PD_DTYPE_BF16 is implemented by the streamed stereo kernels (disp_warp, x-constant disparity, no
 * mask or a row mask, W % 8 == 0, no debug outputs, no PD_FLAG_EXACT_COORDS): pd_warp_composite_supports() says whether a
 * descriptor is served; other requests return PD_ERR_UNSUPPORTED (the Python boundary then upcasts with torch). */
typedef enum pd_dtype { PD_DTYPE_F32 = 0, PD_DTYPE_BF16 = 1 } pd_dtype;

typedef struct pd_strides4 {
    int64_t b, n, y, x;
} pd_strides4;

/* ------------------------------------------------------------------------------------------------
 * Warp + composite: replaces trainer.py:533-603 for ONE target side (grid build, N-plane bilinear
 * warp of [rgb | logit | sigma], validity mask, softmax / Laplacian-mixture weights, compositing)
 * without materialising the N warped tensors.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pd_warp_desc {
    int32_t B, N, H, W;
    int32_t warp_type;      /* pd_warp_type */
    int32_t mixture;        /* opt.use_mixture_loss: sigma channel, w = (pi/sigma)/sum, NLL map */
    int32_t automask;       /* mixture only: also write nll_auto (trainer.py:731-734) */
    int32_t mask_dtype;     /* pd_mask_dtype of pd_warp_in.mask */
    float disp_sign;        /* PD_WARP_DISP: +1 target 'r', -1 target 'l', 0 otherwise (trainer.py:545-548) */
    int32_t flags;          /* bit set of pd_warp_flags */
    int32_t dtype;          /* pd_dtype of logits / sigma / g_logits / g_sigma (the pointers below are then bf16 arrays) */
    int32_t reserved0;
    pd_strides4 disp_stride; /* PD_WARP_DISP / PD_WARP_DEPTH: strides of disp_layered (0 allowed:
                                depth_decoder.py:156 hands out a stride-0 expand) */
    pd_strides4 mask_stride; /* strides of padding_mask */
} pd_warp_desc;

typedef struct pd_warp_in {
    const float* src;    /* [B,3,H,W] source colour  inputs[(color,"l")]           trainer.py:567 */
    const float* tgt;    /* [B,3,H,W] target colour; required iff mixture (NLL)    trainer.py:729 */
    const float* logits; /* [B,N,H,W] outputs["logits"]                           trainer.py:568 */
    const float* sigma;  /* [B,N,H,W] outputs["sigma"]; required iff mixture      trainer.py:571 */
    const float* disp;   /* outputs["disp_layered"], strided (DISP / DEPTH)        trainer.py:534,541 */
    const void* mask;    /* outputs["padding_mask"], strided, mask_dtype; may be NULL (all valid).
                            Ignored for HOMOGRAPHY (mask computed in-kernel, layers.py:223-226) */
    const float* hmat;   /* HOMOGRAPHY: [B*N,12] = H_t2s row-major (9, layers.py:220) then R·n (3,
                            layers.py:223) */
    const float* cam;    /* HOMOGRAPHY: [B,9] inv_K 3x3 row-major.  DEPTH: [B,21] = inv_K 3x3 (9) then
                            (K·T)[:3,:] row-major (12)  (layers.py:152,172) */
} pd_warp_in;

/* Statistics saved for the backward pass: an opaque buffer of pd_warp_composite_stats_bytes(desc) bytes that only
 * pd_warp_composite_bwd reads.  It holds [B,PD_STATS(mixture),H,W] fp32 per-pixel values — non-mixture: {reference
 * logit * log2(e), sum exp(l - ref)}; mixture: additionally {sum exp/sigma, mixture density sum pi*lap + 1e-7} —
 * followed by a row summary of a dense padding_mask (which rows of which planes are all ones) that lets
 * the backward pass leave those mask rows in HBM. */
#define PD_STATS_PLAIN 2
#define PD_STATS_MIXTURE 4

typedef struct pd_warp_out {
    float* rgb_rec;  /* [B,3,H,W]   outputs[("rgb_rec", s)]                       trainer.py:603 */
    float* stats;    /* pd_warp_composite_stats_bytes(desc) bytes, 16-byte aligned, saved for pd_warp_composite_bwd */
    float* nll;      /* [B,1,H,W] mixture: -log(sum pi*lap + 1e-7)                trainer.py:730 */
    float* nll_auto; /* [B,1,H,W] mixture+automask: same with err = |src - tgt|   trainer.py:732-733 */
    /* optional (NULL = not materialised; tests and tensorboard-style consumers only) */
    float* rgb_rec_layered; /* [B,N,3,H,W]                                        trainer.py:582 */
    float* logit_rec;       /* [B,N,H,W]                                          trainer.py:583 */
    float* probability_rec; /* [B,N,H,W]                                          trainer.py:593,602 */
    float* sigma_rec;       /* [B,N,H,W]                                          trainer.py:598 */
    float* pi_rec;          /* [B,N,H,W]                                          trainer.py:599 */
} pd_warp_out;

typedef struct pd_warp_grad_out { /* upstream gradients */
    const float* g_rgb_rec; /* [B,3,H,W] d loss / d rgb_rec (photometric + perceptual); may be NULL (= 0) iff g_ph_sum is set */
    const float* g_nll;     /* [B,1,H,W] d loss / d nll map; mixture only, may be NULL (= 0) */
    /* Fused photometric backward (optional): with g_ph_sum != NULL the kernels form pd_photometric_bwd's result in
     * their prologue instead of reading it from HBM (one launch and one [B,3,H,W] round trip less per step):
     *     g_rgb_rec_eff = [g_rgb_rec] + ph_scale * g_ph_sum[0] * g_unit + [g_pred * mask_novel]
     *     g_nll_eff     = [g_nll]     + ph_scale * g_ph_sum[0] * g_unit_nll
     * g_unit / g_unit_nll are what pd_photometric_fwd left in pd_loss_out (trainer.py:720-742 differentiated). */
    const float* g_ph_sum;   /* [1] device scalar d loss / d ph_sum */
    float ph_scale;          /* pd_loss_desc.out_scale of the forward call (0 is read as 1) */
    const float* g_unit;     /* [B,3,H,W] d ph_sum / d rgb_rec; may be NULL (mixture: the photometric term acts through nll) */
    const float* g_unit_nll; /* [B,1,H,W] d ph_sum / d nll; mixture only, may be NULL */
    const float* g_pred;     /* [B,3,H,W] d loss / d pred from outside (perceptual term); may be NULL */
    const float* mask_novel; /* [B,1,H,W] multiplies g_pred (trainer.py:724-726); NULL = ones */
} pd_warp_grad_out;

typedef struct pd_warp_grad_in { /* produced gradients; any pointer may be NULL (= not needed) */
    float* g_logits; /* [B,N,H,W] */
    float* g_sigma;  /* [B,N,H,W] mixture only */
    float* g_disp;   /* DISP / DEPTH: gradient w.r.t. disp_layered, laid out with g_disp_stride; a 0
                        stride means "reduce over that dimension" (the transpose of the expand) */
    pd_strides4 g_disp_stride;
    float* g_hmat;   /* HOMOGRAPHY: [B*N,9] gradient w.r.t. H_t2s (autograd carries it through
                        inverse / pose algebra on the host side, layers.py:217-220) */
} pd_warp_grad_in;

int pd_version(void);
const char* pd_last_error(void);

/* Launch tuning of the streamed row kernels: a process-wide block for tests (forcing the persistent loop to iterate),
 * benchmarks and kernel experiments.  Initialised ONCE at library load from the environment (PD_STREAM_CTAS,
 * PD_STREAM_HS, PD_STREAM_NST, PD_STREAM_SMEM_KB, PD_STREAM_PX8, PD_STREAM_FWD_MINB, PD_STREAM_BWD_MINB, PD_STREAM_NO_L2_HINT, PD_SSIM_TILES, PD_TAIL_DIRECT), values clamped to legal ranges;
 * 0 = the library's default.  Numerics never depend on it (PD_FLAG_EXACT_COORDS is a per-call descriptor flag). */
typedef struct pd_tuning {
    int32_t stream_ctas_per_sm; /* cap on resident CTAs per SM of the persistent grids (0 = occupancy limit) */
    int32_t stream_hs;          /* planes per ring stage (default 4) */
    int32_t stream_nst;         /* ring stages (default 3, <= 8) */
    int32_t stream_smem_kb;     /* shared-memory budget per CTA the ring is shrunk to (0 = per-shape default) */
    int32_t stream_px8;         /* 8 pixels per thread where the width allows */
    int32_t ssim_tiles;         /* SSIM+L1 forward: shared-memory tile kernel instead of the streamed warp-column kernel */
    int32_t homo_tiles;         /* reserved for the homography kernels (unused) */
    int32_t stream_fwd_minb;    /* resident CTAs per SM the plain narrow forward is compiled for: 0 / 4 = default, 5, 6 (experiments) */
    int32_t stream_no_l2_hint;  /* 1: the logit / sigma / mask rows travel without the L2 evict-first policy (A/B measurements) */
    int32_t stream_bwd_minb;    /* resident CTAs per SM of the plain narrow backward: 0 / 4 = default, 3 = the three-CTA build (A/B runs) */
    int32_t tail_direct;        /* 1: pd_plane_tail_* use the thread-per-pixel kernels instead of the TMA-tile kernels (A/B runs, tests) */
    int32_t reserved[5];
} pd_tuning;
void pd_get_tuning(pd_tuning* out);
void pd_set_tuning(const pd_tuning* in); /* NULL restores the values read from the environment at load */

/* ------------------------------------------------------------------------------------------------
 * Verification of the integrator's "plane geometry does not vary along x" promise (HotPathMixin.disp_rowwise,
 * INTEGRATION.md): compares every element of a [B,N,H,W] tensor with column 0 of its row and adds the number of
 * rows that differ to *violations (an int32 in device memory; the Python boundary copies it to pinned host memory on the
 * same stream and polls it on its next call, so nothing synchronises).  One streaming pass over the tensor.
 * ---------------------------------------------------------------------------------------------- */
int pd_x_constant_check(const void* data, int32_t dtype /* pd_mask_dtype: PD_MASK_F32 | PD_MASK_U8 */, const pd_strides4* strides,
                        int32_t B, int32_t N, int32_t H, int32_t W, int32_t* violations, pd_stream_t stream);

/* Scratch the caller passes as `workspace` to both calls below (contents need not survive between them).  Non-zero
 * for the homography fast path only: the source colour re-packed to one rgbx float4 per pixel. */
size_t pd_warp_composite_workspace_bytes(const pd_warp_desc* desc);
size_t pd_warp_composite_stats_bytes(const pd_warp_desc* desc);

/* 1 if pd_warp_composite_fwd / _bwd serve this descriptor + pointer set (strides, alignment, dtype), else 0.  Only the
 * dtype question can come back 0 for otherwise valid arguments. */
int pd_warp_composite_supports(const pd_warp_desc* desc, const pd_warp_in* in);

int pd_warp_composite_fwd(const pd_warp_desc* desc, const pd_warp_in* in, pd_warp_out* out,
                          void* workspace, pd_stream_t stream);

/* Transpose of the above (what autograd replays through trainer.py:573-603: softmax / mixture
 * weights, mask, ATen grid_sampler_2d_backward incl. the grid gradient, the grid arithmetic). */
int pd_warp_composite_bwd(const pd_warp_desc* desc, const pd_warp_in* in, const pd_warp_out* saved,
                          const pd_warp_grad_out* gout, pd_warp_grad_in* gin, void* workspace,
                          pd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Photometric term of compute_losses for ONE target side: trainer.py:720-742 (+ 687-699 for SSIM).
 *   pred = rgb_rec*m + tgt*(1-m)            if mask_novel                 trainer.py:724-726
 *   L1:       ph = mean_c|pred-tgt|  [min with mean_c|src-tgt| if automask]          :738-741
 *   MIXTURE:  ph = nll [min with nll_auto if automask] [* m]                          :730-736
 *   SSIM_L1:  ph = 0.85*mean_c SSIM(pred,tgt) + 0.15*mean_c|pred-tgt| [min with the same on src]
 *   ph_sum = sum_{b,y,x} ph   (the caller divides by B*H*W:  ph_loss.mean(), :742)
 * ---------------------------------------------------------------------------------------------- */
typedef struct pd_loss_desc {
    int32_t B, H, W;
    int32_t loss_mode; /* pd_loss_mode */
    int32_t automask;
    int32_t has_mask_novel;
    float out_scale; /* ph_sum = out_scale * sum ph (0 is read as 1): folds ph_loss.mean()'s 1/(B*H*W) (:742) into the
                        reduction; the backward pass applies the same factor */
} pd_loss_desc;

typedef struct pd_loss_in {
    const float* rgb_rec;    /* [B,3,H,W] */
    const float* tgt;        /* [B,3,H,W] */
    const float* src;        /* [B,3,H,W] iff automask and loss_mode != MIXTURE */
    const float* mask_novel; /* [B,1,H,W] iff has_mask_novel */
    const float* nll;        /* [B,1,H,W] iff MIXTURE */
    const float* nll_auto;   /* [B,1,H,W] iff MIXTURE && automask */
} pd_loss_in;

typedef struct pd_loss_out {
    float* pred;   /* [B,3,H,W] blended prediction; required iff has_mask_novel (else pred == rgb_rec) */
    float* ph_map; /* [B,1,H,W] optional per-pixel photometric term (NULL = skip) */
    float* ph_sum; /* [1] */
    /* Unit gradients, produced by the forward pass while its tiles are on chip and consumed by
     * pd_photometric_bwd (NULL = the caller will not differentiate): */
    float* g_unit;     /* [B,3,H,W] d ph_sum / d rgb_rec        (L1, SSIM_L1) */
    float* g_unit_nll; /* [B,1,H,W] d ph_sum / d nll            (MIXTURE)     */
} pd_loss_out;

typedef struct pd_loss_grad_out {
    const float* g_ph_sum; /* [1] device scalar: d loss / d ph_sum */
    const float* g_pred;   /* [B,3,H,W] d loss / d pred from outside (perceptual term); may be NULL */
} pd_loss_grad_out;

typedef struct pd_loss_grad_in {
    float* g_rgb_rec; /* [B,3,H,W] */
    float* g_nll;     /* [B,1,H,W] MIXTURE only */
} pd_loss_grad_in;

size_t pd_photometric_workspace_bytes(const pd_loss_desc* desc);
int pd_photometric_fwd(const pd_loss_desc* desc, const pd_loss_in* in, pd_loss_out* out, void* workspace,
                       pd_stream_t stream);
/* g_rgb_rec = g_ph_sum * saved->g_unit + g_pred * mask_novel ; g_nll = g_ph_sum * saved->g_unit_nll.
 * Only in->mask_novel is read from `in` (iff has_mask_novel). */
int pd_photometric_bwd(const pd_loss_desc* desc, const pd_loss_in* in, const pd_loss_out* saved,
                       const pd_loss_grad_out* gout, pd_loss_grad_in* gin, void* workspace, pd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Edge-aware smoothness term of compute_losses: get_smooth_loss_disp (layers.py:243-256) on the crop
 * disp[..., x0:], color[..., x0:] with x0 = int(0.2 W) (trainer.py:768-771):
 *   loss = mean_x |d[x]-d[x+1]| exp(-gamma mean_c|I[x]-I[x+1]|) + mean_y (the same along y)
 * Gradient w.r.t. disp only (the image is data).  Needs W - x0 >= 2 and H >= 2.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pd_smooth_desc {
    int32_t B, H, W;
    int32_t x0;   /* first column of the crop */
    float gamma;  /* opt.gamma_smooth */
} pd_smooth_desc;

size_t pd_smooth_loss_workspace_bytes(const pd_smooth_desc* desc);
/* disp [B,1,H,W], img [B,3,H,W] contiguous; loss [1] */
int pd_smooth_loss_fwd(const pd_smooth_desc* desc, const float* disp, const float* img, float* loss, void* workspace,
                       pd_stream_t stream);
/* g_loss [1] device scalar; g_disp [B,1,H,W], fully written (zero left of the crop) */
int pd_smooth_loss_bwd(const pd_smooth_desc* desc, const float* disp, const float* img, const float* g_loss,
                       float* g_disp, pd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Decoder tail: replaces networks/depth_decoder.py:258-291 (render_probability off) — what DepthDecoder.forward
 * does after its dispconv / sigmaconv convolutions:
 *   logits = raw * padding_mask (:259);  pi = softmax_n(logits) (:276)
 *   mixture: sigma = clamp(sigmoid(sigma_raw), 0.01, 1) (:279-280);  w = pi / sigma * padding_mask;
 *            probability = w / sum_n w (:282-285)                      (else probability = pi)
 *   disp = sum_n probability * disp_layered (:288);  depth = 0.1 * 0.58 * W / disp (:290)
 * ---------------------------------------------------------------------------------------------- */
typedef struct pd_tail_desc {
    int32_t B, N, H, W;
    int32_t mixture;
    int32_t mask_dtype;      /* pd_mask_dtype of padding_mask */
    pd_strides4 disp_stride; /* element strides of disp_layered (0 = broadcast) */
    pd_strides4 mask_stride; /* element strides of padding_mask */
} pd_tail_desc;

typedef struct pd_tail_in {
    const float* logits_raw;   /* [B,N,H,W] convs["dispconv"](x) */
    const float* sigma_raw;    /* [B,N,H,W] convs["sigmaconv"](x); required iff mixture */
    const float* disp_layered; /* strided */
    const void* mask;          /* padding_mask, strided; NULL = all ones */
} pd_tail_in;

typedef struct pd_tail_out {
    float* logits;      /* [B,N,H,W] outputs["logits"] */
    float* sigma;       /* [B,N,H,W] outputs["sigma"]; required iff mixture */
    float* probability; /* [B,N,H,W] outputs["probability"] */
    float* pi;          /* [B,N,H,W] outputs["pi"] (mixture); may be NULL: nothing reads it */
    float* disp;        /* [B,1,H,W] outputs["disp"] */
    float* depth;       /* [B,1,H,W] outputs["depth"]; may be NULL */
    float* stats;       /* [B,3,H,W] saved for pd_plane_tail_bwd */
} pd_tail_out;

typedef struct pd_tail_grad_out { /* upstream gradients, any may be NULL (= 0) */
    const float* g_logits;      /* [B,N,H,W] */
    const float* g_sigma;       /* [B,N,H,W] */
    const float* g_probability; /* [B,N,H,W] */
    const float* g_disp;        /* [B,1,H,W] */
    const float* g_depth;       /* [B,1,H,W] */
} pd_tail_grad_out;

typedef struct pd_tail_grad_in { /* any may be NULL (= not needed) */
    float* g_logits_raw;   /* [B,N,H,W] */
    float* g_sigma_raw;    /* [B,N,H,W] mixture only */
    float* g_disp_layered; /* laid out with g_disp_stride; a 0 stride means "reduce over that dimension" */
    pd_strides4 g_disp_stride;
} pd_tail_grad_in;

int pd_plane_tail_fwd(const pd_tail_desc* desc, const pd_tail_in* in, pd_tail_out* out, pd_stream_t stream);
int pd_plane_tail_bwd(const pd_tail_desc* desc, const pd_tail_in* in, const pd_tail_out* saved,
                      const pd_tail_grad_out* gout, pd_tail_grad_in* gin, pd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Occlusion masks + post-processed disparity: replaces trainer.py:421-466 (the part of
 * Trainer.generate_post_process_disp after the flipped forward pass; no gradients flow).  The decoder
 * ran on the 2B-image batch cat([img, img.flip(-1)]); B below is the size of one half.
 *   o_l        = min(1, sum_n warp_l(softmax_n(warp_r(logits[:B]))))                  :441-447
 *   o_fr       = min(1, sum_n warp_r(softmax_n(warp_l(logits[B:].flip(-1)))))         :449-454
 *   mask_novel = min(1, sum_n warp_r(probability[:B]))                                 :461-463
 *   disp_pp    = (mean*o_fr + disp_l*(1-o_fr))*o_l + disp_f*(1-o_l)                    :456-459
 * with warp_r / warp_l = bilinear sampling at x + disp_layered[:B] / x - disp_layered[B:].
 * ---------------------------------------------------------------------------------------------- */
typedef struct pd_occl_desc {
    int32_t B, N, H, W;      /* B = images of one half */
    int32_t flags;           /* PD_FLAG_EXACT_COORDS: four-tap sampling through the reference's coordinate round trip */
    pd_strides4 disp_stride; /* element strides of disp_layered [2B,N,H,W] (0 = broadcast) */
} pd_occl_desc;

typedef struct pd_occl_in {
    const float* logits;       /* [2B,N,H,W] outputs["logits"] */
    const float* probability;  /* [2B,N,H,W] outputs["probability"] (first half read) */
    const float* disp_layered; /* [2B,N,H,W] strided */
    const float* disp;         /* [2B,1,H,W] outputs["disp"]; required iff out->disp_pp */
} pd_occl_in;

typedef struct pd_occl_out { /* [B,1,H,W] each */
    float* o_l;        /* required */
    float* o_fr;       /* required */
    float* mask_novel; /* may be NULL */
    float* disp_pp;    /* may be NULL */
} pd_occl_out;

size_t pd_occlusion_masks_workspace_bytes(const pd_occl_desc* desc); /* one [B,N,H,W] fp32 plane set */
int pd_occlusion_masks_fwd(const pd_occl_desc* desc, const pd_occl_in* in, pd_occl_out* out, void* workspace,
                           pd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Device-side input staging: replaces the per-sample CPU float resize of the reference's loader
 * (datasets/pair_transforms.py:63-78 Resize, :28-48 RandomResizeCrop) + the fp32 H2D copy of trainer.py:328-329.  The raw
 * uint8 frame crosses PCIe; one kernel does ToTensor (/255), F.interpolate(mode="bicubic", align_corners=True) to
 * [Hf, Wf], the crop [y0 : y0+H, x0 : x0+W] and .clamp(0, 1) into the fp32 [B,3,H,W] tensor the path reads.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pd_resize_desc {
    int32_t B, Hs, Ws;   /* frames, source size */
    int32_t Hf, Wf;      /* size of the full resized frame (0 = H, W: plain Resize) */
    int32_t y0, x0;      /* crop offset inside the resized frame */
    int32_t H, W;        /* output size */
    int32_t src_layout;  /* 0 = [B,Hs,Ws,3] interleaved (a decoder's output), 1 = [B,3,Hs,Ws] planar */
} pd_resize_desc;
int pd_resize_bicubic_u8(const pd_resize_desc* desc, const unsigned char* src, float* dst, pd_stream_t stream);

/* Introspection for tests / bench: number of kernels the library has launched in this process since
 * the last pd_reset_launch_count() (bench.py reports it as gpu_launches). */
int64_t pd_launch_count(void);
void pd_reset_launch_count(void);

/* Test hook: evaluates the coordinate normalise / un-normalise round trip of `size` (trainer.py:549-551
 * + ATen grid_sampler_unnormalize) for n coordinates with the IEEE-division form (out_exact) and with the
 * division-free form the row-tiled kernels use (out_fast); tests require them to be bit-identical. */
int pd_debug_roundtrip(const float* u, int64_t n, int32_t size, float* out_exact, float* out_fast, pd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PLANEDEPTH_B200_H */
