// Frame sink: decoded RGB frames -> 8-bit interleaved images, on the device.
//
// What the reference does per frame and output key after forward_flow
// (/root/reference/test_animating/test_v1_4eval_rawsize.py:240-242, 284-286): bilinear resize to the raw
// size (F.interpolate, align_corners=False), permute to HWC, .cpu().numpy(), * 0.5 + 0.5, * 255, RGB -> BGR,
// cv2.imwrite (which rounds to nearest-even and saturates to 0..255).  The float frame crosses PCIe (12 bytes
// per pixel) and the rest runs on one host core.  Here one kernel does resize + scale + round + saturate +
// channel swap + interleave and writes 3 bytes per pixel, so that only those leave the device.
#include "slr_common.cuh"
#include "slr_host.h"

namespace slr {

// torch's area_pixel_compute_source_index for bilinear, align_corners=False
__device__ __forceinline__ float source_index(float scale, int dst, int in_size)
{
    const float s = scale * ((float)dst + 0.5f) - 0.5f;
    return fminf(fmaxf(s, 0.0f), (float)(in_size - 1));       // lower clamp as torch; the upper one only guards the index
}

__global__ void __launch_bounds__(256)
frame_sink_kernel(const float* __restrict__ in, uint8_t* __restrict__ out, int n, int H, int W, int Ho, int Wo,
                  float scale_h, float scale_w, float mul, float add, int swap_rb)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t total = (int64_t)n * Ho * Wo;
    if (i >= total) return;
    const int xo = (int)(i % Wo), yo = (int)(i / Wo % Ho), f = (int)(i / ((int64_t)Wo * Ho));
    const float* frame = in + (int64_t)f * 3 * H * W;
    float v[3];
    if (Ho == H && Wo == W) {
        #pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = __ldg(frame + ((int64_t)c * H + yo) * W + xo);
    } else {
        const float sy = source_index(scale_h, yo, H), sx = source_index(scale_w, xo, W);
        const int y0 = (int)sy, x0 = (int)sx;
        const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
        const float ly = sy - (float)y0, lx = sx - (float)x0, hy = 1.0f - ly, hx = 1.0f - lx;
        #pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* pl = frame + (int64_t)c * H * W;
            // same association as torch's upsample_bilinear2d: h0 * (w0 * a + w1 * b) + h1 * (w0 * c + w1 * d)
            const float top = __fadd_rn(__fmul_rn(hx, __ldg(pl + (int64_t)y0 * W + x0)), __fmul_rn(lx, __ldg(pl + (int64_t)y0 * W + x1)));
            const float bot = __fadd_rn(__fmul_rn(hx, __ldg(pl + (int64_t)y1 * W + x0)), __fmul_rn(lx, __ldg(pl + (int64_t)y1 * W + x1)));
            v[c] = __fadd_rn(__fmul_rn(hy, top), __fmul_rn(ly, bot));
        }
    }
    uint8_t* o = out + i * 3;
    #pragma unroll
    for (int c = 0; c < 3; ++c) {
        // (x * 0.5 + 0.5) * 255 with the reference's two roundings, then cv2's saturate_cast<uchar>(cvRound(.))
        const float s = __fmul_rn(__fadd_rn(__fmul_rn(v[c], mul), add), 255.0f);
        const float r = fminf(fmaxf(rintf(s), 0.0f), 255.0f);
        o[swap_rb ? 2 - c : c] = (uint8_t)(s == s ? r : 0.0f);      // NaN -> 0
    }
}

}  // namespace slr

extern "C" int slr_frame_sink_u8(const float* frames, uint8_t* out, int64_t n, int64_t H, int64_t W,
                                 int64_t out_H, int64_t out_W, float mul, float add, int swap_rb, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(frames && out && n > 0 && H > 0 && W > 0 && out_H > 0 && out_W > 0 && n * out_H * out_W < (1ll << 40),
                   "slr_frame_sink_u8: bad arguments");
    const int64_t total = n * out_H * out_W;
    slr::frame_sink_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
        frames, out, (int)n, (int)H, (int)W, (int)out_H, (int)out_W, (float)H / (float)out_H, (float)W / (float)out_W,
        mul, add, swap_rb);
    return SLR_LAUNCH_STATUS();
}
