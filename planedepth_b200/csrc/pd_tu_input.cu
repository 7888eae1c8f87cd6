// Translation unit: device-side input staging.
//
// The reference's loader decodes a frame to uint8, converts it to float on the CPU (ToTensor: /255) and resizes it with
// F.interpolate(mode="bicubic", align_corners=True).clamp(0, 1) per sample (/root/reference/datasets/pair_transforms.py:
// 28-48 RandomResizeCrop, 63-78 Resize), then ships fp32 [3,H,W] tensors over PCIe (trainer.py:328-329).  Here the raw
// uint8 frame crosses PCIe (1.4 MB per KITTI frame whatever the training resolution; 4x fewer bytes than fp32 at 1280x384)
// and one kernel does /255 + bicubic resize (+ crop) + clamp into the fp32 NCHW tensor the path reads.
//
// One thread per output pixel, 48 byte taps through L1.  A shared-memory tile variant (64 x 4 output pixels per CTA, source
// window staged as floats, interleaved and channel-planar layouts) was built and measured in round 2: 31.5 / 43 us against
// 33 us for this kernel at [12,3,192,640] <- 375 x 1242 — at a 2x reduction a 4-row tile needs 12 source rows, so staging
// converts almost as many elements as the taps read — and was dropped.
//
// Arithmetic = ATen upsample_bicubic2d with align_corners=True: source coordinate s = o * (in - 1) / (out - 1) (scale formed
// in fp32), cubic-convolution weights with A = -0.75, taps at floor(s) - 1 .. floor(s) + 2 with indices clamped to the image,
// rows interpolated along x first, then along y.
#include <string.h>

#include "pd_device.cuh"

namespace {

using pd::check_device;
using pd::check_launch;
using pd::fail;

struct ResizeParams {
    pd_resize_desc d;
    const unsigned char* src;
    float* dst;
    float sy, sx;  // (in - 1) / (out_full - 1)
};

__device__ __forceinline__ void cubic_weights(float t, float (&w)[4]) {
    const float A = -0.75f;
    const float x0 = t + 1.0f, x3 = 2.0f - t, x2 = 1.0f - t;
    w[0] = ((A * x0 - 5.0f * A) * x0 + 8.0f * A) * x0 - 4.0f * A;
    w[1] = ((A + 2.0f) * t - (A + 3.0f)) * t * t + 1.0f;
    w[2] = ((A + 2.0f) * x2 - (A + 3.0f)) * x2 * x2 + 1.0f;
    w[3] = ((A * x3 - 5.0f * A) * x3 + 8.0f * A) * x3 - 4.0f * A;
}

template <bool HWC>
__global__ void __launch_bounds__(256) resize_bicubic_u8_kernel(const ResizeParams p) {
    const int W = p.d.W, H = p.d.H;
    const int64_t hw = (int64_t)H * W;
    const int64_t total = (int64_t)p.d.B * hw;
    const float k255 = 1.0f / 255.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / hw);
        const int rem = (int)(i - (int64_t)b * hw);
        const int y = rem / W, x = rem - y * W;
        const float fy = p.sy * (float)(y + p.d.y0), fx = p.sx * (float)(x + p.d.x0);
        const int iy = __float2int_rd(fy), ix = __float2int_rd(fx);
        const float fy0 = (float)iy, fx0 = (float)ix;
        float wy[4], wx[4];
        cubic_weights(fy - fy0, wy);
        cubic_weights(fx - fx0, wx);
        int ys[4], xs[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ys[k] = min(max(iy - 1 + k, 0), p.d.Hs - 1);
            xs[k] = min(max(ix - 1 + k, 0), p.d.Ws - 1);
        }
        const int64_t img = (int64_t)b * 3 * p.d.Hs * p.d.Ws;
        float acc[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float row[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int64_t o = HWC ? img + ((int64_t)ys[j] * p.d.Ws + xs[k]) * 3 + c : img + ((int64_t)c * p.d.Hs + ys[j]) * p.d.Ws + xs[k];
                    const float v = (float)__ldg(p.src + o) * k255;  // ToTensor: uint8 / 255
                    row[c] = (k == 0) ? v * wx[0] : fmaf(v, wx[k], row[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) acc[c] = (j == 0) ? row[c] * wy[0] : fmaf(row[c], wy[j], acc[c]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) p.dst[((int64_t)b * 3 + c) * hw + rem] = fminf(fmaxf(acc[c], 0.0f), 1.0f);  // .clamp(0, 1)
    }
}

}  // namespace

extern "C" {

int pd_resize_bicubic_u8(const pd_resize_desc* d, const unsigned char* src, float* dst, pd_stream_t stream) {
    if (!d || !src || !dst) return fail(PD_ERR_ARG, "NULL argument");
    if (d->B < 1 || d->Hs < 1 || d->Ws < 1 || d->H < 1 || d->W < 1) return fail(PD_ERR_SHAPE, "B, Hs, Ws, H, W >= 1 required");
    const int Hf = d->Hf > 0 ? d->Hf : d->H, Wf = d->Wf > 0 ? d->Wf : d->W;
    if (d->y0 < 0 || d->x0 < 0 || d->y0 + d->H > Hf || d->x0 + d->W > Wf) return fail(PD_ERR_SHAPE, "crop window outside the resized frame");
    if (d->src_layout != 0 && d->src_layout != 1) return fail(PD_ERR_ARG, "src_layout must be 0 (HWC) or 1 (CHW)");
    int rc;
    if ((rc = check_device())) return rc;
    ResizeParams p;
    memset(&p, 0, sizeof(p));
    p.d = *d;
    p.d.Hf = Hf, p.d.Wf = Wf;
    p.src = src, p.dst = dst;
    // ATen area_pixel_compute_scale(align_corners=True): (in - 1) / (out - 1) in fp32, 0 when out == 1
    p.sy = Hf > 1 ? (float)(d->Hs - 1) / (float)(Hf - 1) : 0.0f;
    p.sx = Wf > 1 ? (float)(d->Ws - 1) / (float)(Wf - 1) : 0.0f;
    const int64_t total = (int64_t)d->B * d->H * d->W;
    const int64_t want = (total + 255) / 256, cap = (int64_t)pd::sm_count() * 16;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (d->src_layout == 0) resize_bicubic_u8_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    else resize_bicubic_u8_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("resize_bicubic_u8");
}

}  // extern "C"
