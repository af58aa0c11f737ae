// The splat-input producer of the TRAINING forward, fused into the splat and differentiable
// (SURVEY 8 f2).  The reference builds, per direction,
//     tenInput = cat([fs * Z_norm.exp() * alpha, Z_norm.exp() * alpha], 1)          B x (C + 1) x H x W
// (/root/reference/models/animating_softmax_splating.py:606 forward, :651 backward direction: another
// feature set `end_fs` / `Z_p` and 1 - alpha), splats it (ModuleSoftsplat('summation'), :629-632 / :672-676 ->
// kernel_Softsplat_updateOutput, models/softsplat.py:157-202) and lets autograd walk back through the
// splat's two backward kernels (softsplat.py:204-326, :427-477) and the cat / exp / mul nodes.
//
//   producer_splat_fwd   one pass: e^Zn * alpha and the C products are made in registers and scattered;
//                        tenInput never exists.  `accumulate` adds the second direction into the same
//                        accumulator (the reference adds the two splat outputs, :684-686).
//   producer_splat_bwd   one pass per source pixel over the C + 1 channels: the four gradient corners are
//                        read ONCE and feed d(fs), d(Zn) and d(flow) together (the reference reads them in
//                        two kernels and then runs five elementwise backward nodes over B x 65 x H x W).
//
// Arithmetic follows the reference's association: value = (fs * e^Zn) * alpha, splat adds value * weight;
// grad wrt tenInput sums g * w over NW, NE, SW, SE (softsplat.py:230-252); grad wrt flow accumulates
// (value * g) * d(weight) channel by channel (:304-322).
#include "slr_common.cuh"
#include "slr_host.h"
#include <algorithm>

namespace slr {

constexpr int kTrainBlock = 256;

// grid: (ceil(P / 256), channel chunks, B); chunk 0 also does the normaliser channel C
__global__ void __launch_bounds__(kTrainBlock)
producer_splat_fwd_kernel(const float* __restrict__ fs, const float* __restrict__ zn, const float* __restrict__ flow,
                          const float* __restrict__ alpha, float* __restrict__ acc, int C, int H, int W, int c_per_block)
{
    const int64_t P = (int64_t)H * W;
    const int64_t p = (int64_t)blockIdx.x * kTrainBlock + threadIdx.x;
    if (p >= P) return;
    const int b = blockIdx.z;
    const int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
    const float* fl = flow + (int64_t)b * 2 * P;
    const Footprint f = landing(x, y, fl[p], fl[P + p], H, W);
    if (f.ok == 0u) return;
    const float ez = expf(zn[(int64_t)b * P + p]);
    const float a = alpha[b];
    const int c0 = blockIdx.y * c_per_block;
    const int c1 = min(C, c0 + c_per_block);
    const int64_t q = (int64_t)f.y0 * W + f.x0;       // NW cell (may be outside; only used with the ok bits)
    const float* src = fs + ((int64_t)b * C + c0) * P + p;
    float* dst = acc + ((int64_t)b * (C + 1) + c0) * P + q;
    auto scatter = [&](float* d, float v) {
        if (f.ok & 1u) red_add(d, v * f.w[0]);
        if (f.ok & 2u) red_add(d + 1, v * f.w[1]);
        if (f.ok & 4u) red_add(d + W, v * f.w[2]);
        if (f.ok & 8u) red_add(d + W + 1, v * f.w[3]);
    };
    #pragma unroll 4
    for (int c = c0; c < c1; ++c, src += P, dst += P) scatter(dst, __fmul_rn(__fmul_rn(*src, ez), a));
    if (blockIdx.y == 0) scatter(acc + ((int64_t)b * (C + 1) + C) * P + q, __fmul_rn(ez, a));
}

// grid: (ceil(P / 256), 1, B); any of d_fs / d_zn / d_flow may be NULL (not needed)
__global__ void __launch_bounds__(kTrainBlock)
producer_splat_bwd_kernel(const float* __restrict__ fs, const float* __restrict__ zn, const float* __restrict__ flow,
                          const float* __restrict__ alpha, const float* __restrict__ gacc,
                          float* __restrict__ d_fs, float* __restrict__ d_zn, float* __restrict__ d_flow,
                          int C, int H, int W)
{
    const int64_t P = (int64_t)H * W;
    const int64_t p = (int64_t)blockIdx.x * kTrainBlock + threadIdx.x;
    if (p >= P) return;
    const int b = blockIdx.z;
    const int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
    const float* fl = flow + (int64_t)b * 2 * P;
    const float fx = fl[p], fy = fl[P + p];
    const Footprint f = landing(x, y, fx, fy, H, W);
    const float ox = (float)x + fx, oy = (float)y + fy;
    const float bx = (float)f.x0, by = (float)f.y0, ex = (float)(f.x0 + 1), ey = (float)(f.y0 + 1);
    // d(weight)/d(flow_x) and d(weight)/d(flow_y) for NW, NE, SW, SE (softsplat.py:290-300)
    const float dxw[4] = { -1.0f * (ey - oy), +1.0f * (ey - oy), -1.0f * (oy - by), +1.0f * (oy - by) };
    const float dyw[4] = { (ex - ox) * -1.0f, (ox - bx) * -1.0f, (ex - ox) * +1.0f, (ox - bx) * +1.0f };
    const float ez = expf(zn[(int64_t)b * P + p]);
    const float a = alpha[b];
    const int64_t q = (int64_t)f.y0 * W + f.x0;
    const float* src = fs + (int64_t)b * C * P + p;
    const float* g = gacc + (int64_t)b * (C + 1) * P + q;
    float* o = d_fs ? d_fs + (int64_t)b * C * P + p : nullptr;
    float gx = 0.0f, gy = 0.0f, d_ez = 0.0f;
    for (int c = 0; c <= C; ++c, src += P, g += P) {
        const bool norm = c == C;
        const float s = norm ? 1.0f : *src;
        const float v = norm ? __fmul_rn(ez, a) : __fmul_rn(__fmul_rn(s, ez), a);     // the tenInput value of this channel
        const float g0 = (f.ok & 1u) ? g[0] : 0.0f, g1 = (f.ok & 2u) ? g[1] : 0.0f;
        const float g2 = (f.ok & 4u) ? g[W] : 0.0f, g3 = (f.ok & 8u) ? g[W + 1] : 0.0f;
        // grad wrt tenInput (softsplat.py:230-252)
        float G = 0.0f;
        if (f.ok & 1u) G += g0 * f.w[0];
        if (f.ok & 2u) G += g1 * f.w[1];
        if (f.ok & 4u) G += g2 * f.w[2];
        if (f.ok & 8u) G += g3 * f.w[3];
        const float Ga = G * a;                        // ... wrt fs * e^Zn (or e^Zn for the normaliser channel)
        if (!norm && o) { *o = Ga * ez; o += P; }
        d_ez += norm ? Ga : Ga * s;
        // grad wrt flow (softsplat.py:304-322)
        if (f.ok & 1u) { const float t = v * g0; gx += t * dxw[0]; gy += t * dyw[0]; }
        if (f.ok & 2u) { const float t = v * g1; gx += t * dxw[1]; gy += t * dyw[1]; }
        if (f.ok & 4u) { const float t = v * g2; gx += t * dxw[2]; gy += t * dyw[2]; }
        if (f.ok & 8u) { const float t = v * g3; gx += t * dxw[3]; gy += t * dyw[3]; }
    }
    if (d_zn) d_zn[(int64_t)b * P + p] = d_ez * ez;
    if (d_flow) {
        d_flow[(int64_t)b * 2 * P + p] = gx;
        d_flow[(int64_t)b * 2 * P + P + p] = gy;
    }
}

}  // namespace slr

using namespace slr;

extern "C" int slr_producer_splat_fwd(const float* fs, const float* zn, const float* flow, const float* alpha,
                                      float* acc, int64_t B, int64_t C, int64_t H, int64_t W, int accumulate,
                                      slr_stream_t stream_)
{
    SLR_CHECK_ARGS(fs && zn && flow && alpha && acc && B > 0 && B <= 65535 && C > 0 && H > 0 && W > 0 && H * W < (1ll << 31),
                   "slr_producer_splat_fwd: bad arguments");
    cudaStream_t s = (cudaStream_t)stream_;
    const int64_t P = H * W;
    if (!accumulate) SLR_CUDA(cudaMemsetAsync(acc, 0, sizeof(float) * (size_t)(B * (C + 1) * P), s));
    const int c_per_block = 16;
    dim3 grid((unsigned)((P + kTrainBlock - 1) / kTrainBlock), (unsigned)((C + c_per_block - 1) / c_per_block), (unsigned)B);
    producer_splat_fwd_kernel<<<grid, kTrainBlock, 0, s>>>(fs, zn, flow, alpha, acc, (int)C, (int)H, (int)W, c_per_block);
    return SLR_LAUNCH_STATUS();
}

extern "C" int slr_producer_splat_bwd(const float* fs, const float* zn, const float* flow, const float* alpha,
                                      const float* grad_acc, float* d_fs, float* d_zn, float* d_flow,
                                      int64_t B, int64_t C, int64_t H, int64_t W, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(fs && zn && flow && alpha && grad_acc && (d_fs || d_zn || d_flow) && B > 0 && B <= 65535 && C > 0 &&
                   H > 0 && W > 0 && H * W < (1ll << 31), "slr_producer_splat_bwd: bad arguments");
    const int64_t P = H * W;
    dim3 grid((unsigned)((P + kTrainBlock - 1) / kTrainBlock), 1, (unsigned)B);
    producer_splat_bwd_kernel<<<grid, kTrainBlock, 0, (cudaStream_t)stream_>>>(fs, zn, flow, alpha, grad_acc, d_fs, d_zn, d_flow,
                                                                             (int)C, (int)H, (int)W);
    return SLR_LAUNCH_STATUS();
}
