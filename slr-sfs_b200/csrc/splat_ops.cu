// Operator-level kernels: one per reference kernel launch, plus the scatter
// variant of the joint block and the normalise kernel.  sm_100a.
//
// These are HBM / L2-atomic bound integer-and-fp32 kernels (no tensor-core
// work).  Unlike the reference (one thread per ELEMENT, /root/reference/
// models/softsplat.py:162-166) every kernel here runs one thread per PIXEL and
// loops over channels, so the flow is read and the footprint computed once per
// pixel instead of C times, and loads along x stay coalesced.
#include "slr_common.cuh"
#include "slr_host.h"
#include <algorithm>

namespace slr {

constexpr int kBlock = 256;

// ---------------------------------------------------------------------------
// summation splat (softsplat.py:157-202)
// grid: (ceil(P/kBlock), channel chunks, B)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
softsplat_sum_fwd_kernel(const float* __restrict__ in, const float* __restrict__ flow,
                         float* __restrict__ out, int C, int H, int W, int c_per_block)
{
    const int64_t P = (int64_t)H * W;
    const int64_t p = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (p >= P) return;
    const int b = blockIdx.z;
    const int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
    const float* fl = flow + (int64_t)b * 2 * P;
    const Footprint f = landing(x, y, fl[p], fl[P + p], H, W);
    if (f.ok == 0u) return;
    const int c0 = blockIdx.y * c_per_block;
    const int c1 = min(C, c0 + c_per_block);
    const int64_t q = (int64_t)f.y0 * W + f.x0;   // NW cell (may be outside; only used with ok bits)
    const float* src = in + ((int64_t)b * C + c0) * P + p;
    float* dst = out + ((int64_t)b * C + c0) * P + q;
    #pragma unroll 4
    for (int c = c0; c < c1; ++c, src += P, dst += P) {
        const float v = *src;
        if (f.ok & 1u) red_add(dst, v * f.w[0]);
        if (f.ok & 2u) red_add(dst + 1, v * f.w[1]);
        if (f.ok & 4u) red_add(dst + W, v * f.w[2]);
        if (f.ok & 8u) red_add(dst + W + 1, v * f.w[3]);
    }
}

// ---------------------------------------------------------------------------
// grad wrt input (softsplat.py:204-255): gather, deterministic
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
softsplat_grad_input_kernel(const float* __restrict__ flow, const float* __restrict__ gout,
                            float* __restrict__ gin, int C, int H, int W, int c_per_block)
{
    const int64_t P = (int64_t)H * W;
    const int64_t p = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (p >= P) return;
    const int b = blockIdx.z;
    const int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
    const float* fl = flow + (int64_t)b * 2 * P;
    const Footprint f = landing(x, y, fl[p], fl[P + p], H, W);
    const int c0 = blockIdx.y * c_per_block;
    const int c1 = min(C, c0 + c_per_block);
    const int64_t q = (int64_t)f.y0 * W + f.x0;
    const float* g = gout + ((int64_t)b * C + c0) * P + q;
    float* o = gin + ((int64_t)b * C + c0) * P + p;
    #pragma unroll 4
    for (int c = c0; c < c1; ++c, g += P, o += P) {
        float acc = 0.0f;
        if (f.ok & 1u) acc += g[0] * f.w[0];
        if (f.ok & 2u) acc += g[1] * f.w[1];
        if (f.ok & 4u) acc += g[W] * f.w[2];
        if (f.ok & 8u) acc += g[W + 1] * f.w[3];
        *o = acc;
    }
}

// ---------------------------------------------------------------------------
// grad wrt flow (softsplat.py:257-326): one thread per pixel produces both
// components; channel loop order and the (in * gout) * dweight association
// follow :304-322.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
softsplat_grad_flow_kernel(const float* __restrict__ in, const float* __restrict__ flow,
                           const float* __restrict__ gout, float* __restrict__ gflow,
                           int C, int H, int W)
{
    const int64_t P = (int64_t)H * W;
    const int64_t p = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (p >= P) return;
    const int b = blockIdx.z;
    const int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
    const float* fl = flow + (int64_t)b * 2 * P;
    const float fx = fl[p], fy = fl[P + p];
    const Footprint f = landing(x, y, fx, fy, H, W);
    const float ox = (float)x + fx, oy = (float)y + fy;
    const float bx = (float)f.x0, by = (float)f.y0, ex = (float)(f.x0 + 1), ey = (float)(f.y0 + 1);
    // d(weight)/d(flow_x) and d(weight)/d(flow_y) for NW, NE, SW, SE (:290-300)
    const float dxw[4] = { -1.0f * (ey - oy), +1.0f * (ey - oy), -1.0f * (oy - by), +1.0f * (oy - by) };
    const float dyw[4] = { (ex - ox) * -1.0f, (ox - bx) * -1.0f, (ex - ox) * +1.0f, (ox - bx) * +1.0f };
    const int64_t q = (int64_t)f.y0 * W + f.x0;
    const float* src = in + (int64_t)b * C * P + p;
    const float* g = gout + (int64_t)b * C * P + q;
    float gx = 0.0f, gy = 0.0f;
    for (int c = 0; c < C; ++c, src += P, g += P) {
        const float v = *src;
        if (f.ok & 1u) { const float t = v * g[0];     gx += t * dxw[0]; gy += t * dyw[0]; }
        if (f.ok & 2u) { const float t = v * g[1];     gx += t * dxw[1]; gy += t * dyw[1]; }
        if (f.ok & 4u) { const float t = v * g[W];     gx += t * dxw[2]; gy += t * dyw[2]; }
        if (f.ok & 8u) { const float t = v * g[W + 1]; gx += t * dxw[3]; gy += t * dyw[3]; }
    }
    gflow[(int64_t)b * 2 * P + p] = gx;
    gflow[(int64_t)b * 2 * P + P + p] = gy;
}

// ---------------------------------------------------------------------------
// max splat (softsplat.py:12-82) and the inverse gather (:84-155)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
fill_kernel(float* __restrict__ p, float v, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) p[i] = v;
}

__global__ void __launch_bounds__(kBlock)
maxsplat_fwd_kernel(const float* __restrict__ in, const float* __restrict__ flow,
                    float* __restrict__ out, int C, int H, int W)
{
    const int64_t P = (int64_t)H * W;
    const int64_t p = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (p >= P) return;
    const int b = blockIdx.z;
    const int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
    const float* fl = flow + (int64_t)b * 2 * P;
    const Footprint f = landing(x, y, fl[p], fl[P + p], H, W);
    if (f.ok == 0u) return;
    const int64_t q = (int64_t)f.y0 * W + f.x0;
    const float* src = in + (int64_t)b * C * P + p;
    float* dst = out + (int64_t)b * C * P + q;
    for (int c = 0; c < C; ++c, src += P, dst += P) {
        const float v = *src;
        if (f.ok & 1u) red_max(dst, v * f.w[0]);
        if (f.ok & 2u) red_max(dst + 1, v * f.w[1]);
        if (f.ok & 4u) red_max(dst + W, v * f.w[2]);
        if (f.ok & 8u) red_max(dst + W + 1, v * f.w[3]);
    }
}

__global__ void __launch_bounds__(kBlock)
inversesplat_kernel(const float* __restrict__ own, const float* __restrict__ warped,
                    const float* __restrict__ flow, float* __restrict__ out, int C, int H, int W)
{
    const int64_t P = (int64_t)H * W;
    const int64_t p = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (p >= P) return;
    const int b = blockIdx.z;
    const int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
    const float* fl = flow + (int64_t)b * 2 * P;
    const Footprint f = landing(x, y, fl[p], fl[P + p], H, W);
    const int64_t q = (int64_t)f.y0 * W + f.x0;
    for (int c = 0; c < C; ++c) {
        const int64_t base = ((int64_t)b * C + c) * P;
        const float* g = warped + base + q;
        float m = own[base + p];
        if (f.ok & 1u) m = fmaxf(g[0], m);
        if (f.ok & 2u) m = fmaxf(g[1], m);
        if (f.ok & 4u) m = fmaxf(g[W], m);
        if (f.ok & 8u) m = fmaxf(g[W + 1], m);
        out[base + p] = m;
    }
}

// ---------------------------------------------------------------------------
// Euler integration (euler_integration_manipulator.py:18-56).  Each pixel's
// chain is independent (the field is constant), so the T dependent steps run in
// registers: no per-step launches, no boolean-mask indexing, no host sync.
// rintf = round-half-to-even like torch.round; adds are plain fp32 in the
// reference's order, so the result is bit-identical.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
euler_kernel(const float* __restrict__ motion, float sign, int T, float* __restrict__ disp,
             float* __restrict__ visible, int H, int W)
{
    const int64_t P = (int64_t)H * W;
    const int64_t p = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (p >= P) return;
    const int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
    const float cx = (float)x, cy = (float)y;
    const float xmax = (float)(W - 1), ymax = (float)(H - 1);
    float dx = cx, dy = cy;
    bool invalid = false;
    for (int k = 0; k < T; ++k) {
        const int64_t at = (int64_t)rintf(dy) * W + (int64_t)rintf(dx);
        const float mx = __fmul_rn(sign, __ldg(motion + at));
        const float my = __fmul_rn(sign, __ldg(motion + P + at));
        dx = __fadd_rn(dx, mx);
        dy = __fadd_rn(dy, my);
        // written so that a NaN coordinate (NaN / inf motion) FAILS the test: the chain turns invalid
        // instead of indexing the field with (int64_t)rintf(NaN)
        invalid = invalid || !(dx <= xmax && dx >= 0.0f && dy <= ymax && dy >= 0.0f);
        if (invalid) { dx = cx; dy = cy; }
    }
    const float sentinel = (float)(max(H, W) + 1);
    disp[p] = invalid ? sentinel : __fsub_rn(dx, cx);
    disp[P + p] = invalid ? sentinel : __fsub_rn(dy, cy);
    if (visible) visible[p] = invalid ? 0.0f : 1.0f;
}

// ---------------------------------------------------------------------------
// euler backward: gradient of the displacements with respect to the motion field.
// In the reference the chain is differentiable through the VALUES it samples
// (destination_coords + motion[0][:, iy, ix], euler_integration_manipulator.py:37-38;
// the rounded indices carry no gradient), and a pixel that ends invalid has its
// displacement overwritten by the sentinel constant (:53-55), i.e. no gradient.  So
//   grad_motion[:, q] += sign * grad_disp[:, p]   for every step of a VALID chain p that sampled q.
// One thread per pixel: the chain is walked once to learn its validity, then again to
// add the gradient at every visited sample (fp32 reductions at L2).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
euler_grad_kernel(const float* __restrict__ motion, float sign, int T, const float* __restrict__ gdisp,
                  float* __restrict__ gmotion, int H, int W)
{
    const int64_t P = (int64_t)H * W;
    const int64_t p = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (p >= P) return;
    const int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
    const float cx = (float)x, cy = (float)y;
    const float xmax = (float)(W - 1), ymax = (float)(H - 1);
    float dx = cx, dy = cy;
    bool invalid = false;
    for (int k = 0; k < T && !invalid; ++k) {
        const int64_t at = (int64_t)rintf(dy) * W + (int64_t)rintf(dx);
        dx = __fadd_rn(dx, __fmul_rn(sign, __ldg(motion + at)));
        dy = __fadd_rn(dy, __fmul_rn(sign, __ldg(motion + P + at)));
        invalid = !(dx <= xmax && dx >= 0.0f && dy <= ymax && dy >= 0.0f);
    }
    if (invalid) return;
    const float gx = __fmul_rn(sign, gdisp[p]), gy = __fmul_rn(sign, gdisp[P + p]);
    dx = cx; dy = cy;
    for (int k = 0; k < T; ++k) {
        const int64_t at = (int64_t)rintf(dy) * W + (int64_t)rintf(dx);
        red_add(gmotion + at, gx);
        red_add(gmotion + P + at, gy);
        dx = __fadd_rn(dx, __fmul_rn(sign, __ldg(motion + at)));
        dy = __fadd_rn(dy, __fmul_rn(sign, __ldg(motion + P + at)));
    }
}

// ---------------------------------------------------------------------------
// max reduction into a device scalar (Z.max())
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
reduce_max_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out)
{
    __shared__ float part[kBlock / 32];
    float m = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
        m = fmaxf(m, x[i]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < kBlock / 32 ? part[threadIdx.x] : -INFINITY;
        m = warp_max(m);
        if (threadIdx.x == 0) red_max(out, m);
    }
}

// ---------------------------------------------------------------------------
// joint block, scatter variant ("algorithm A"): both directions in one pass over
// the features, accumulating into a [C+n_tail+1,H,W] buffer with fp32 REDs.
// Value association follows the reference: ((feat * e^Z) * a_d) * w.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
joint_scatter_kernel(const float* __restrict__ feat, const float* __restrict__ z,
                     const float* __restrict__ zsub, const float* __restrict__ tail, int n_tail,
                     const float* __restrict__ disp_f, const float* __restrict__ disp_b,
                     float a_f, float a_b, float* __restrict__ acc, int C, int H, int W, int c_per_block)
{
    const int64_t P = (int64_t)H * W;
    const int64_t p = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (p >= P) return;
    const int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
    const Footprint ff = landing(x, y, disp_f[p], disp_f[P + p], H, W);
    const Footprint fb = landing(x, y, disp_b[p], disp_b[P + p], H, W);
    if ((ff.ok | fb.ok) == 0u) return;
    const float ez = expf(z[p] - (zsub ? *zsub : 0.0f));
    const int64_t qf = (int64_t)ff.y0 * W + ff.x0, qb = (int64_t)fb.y0 * W + fb.x0;

    auto splat = [&](float v, float* plane) {
        const float vf = v * a_f, vb = v * a_b;
        float* d = plane + qf;
        if (ff.ok & 1u) red_add(d, vf * ff.w[0]);
        if (ff.ok & 2u) red_add(d + 1, vf * ff.w[1]);
        if (ff.ok & 4u) red_add(d + W, vf * ff.w[2]);
        if (ff.ok & 8u) red_add(d + W + 1, vf * ff.w[3]);
        d = plane + qb;
        if (fb.ok & 1u) red_add(d, vb * fb.w[0]);
        if (fb.ok & 2u) red_add(d + 1, vb * fb.w[1]);
        if (fb.ok & 4u) red_add(d + W, vb * fb.w[2]);
        if (fb.ok & 8u) red_add(d + W + 1, vb * fb.w[3]);
    };

    const int c0 = blockIdx.y * c_per_block;
    const int c1 = min(C, c0 + c_per_block);
    #pragma unroll 2
    for (int c = c0; c < c1; ++c) splat(feat[(int64_t)c * P + p] * ez, acc + (int64_t)c * P);
    if (blockIdx.y == 0) {
        for (int j = 0; j < n_tail; ++j) splat(tail[(int64_t)j * P + p], acc + (int64_t)(C + j) * P);
        splat(ez, acc + (int64_t)(C + n_tail) * P);
    }
}

// out[c] = acc[c] / max(acc[norm_ch], eps); float4 along x when the row length allows.
__global__ void __launch_bounds__(kBlock)
normalize_kernel(const float* __restrict__ acc, float* __restrict__ out, float* __restrict__ mask,
                 int n_out, int norm_ch, float eps, int64_t P)
{
    const int64_t p = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (p >= P) return;
    const float nrm = acc[(int64_t)norm_ch * P + p];
    const float den = fmaxf(nrm, eps);
    if (mask && blockIdx.y == 0) mask[p] = nrm > eps ? 1.0f : 0.0f;
    const int per = (n_out + gridDim.y - 1) / gridDim.y;
    const int c0 = blockIdx.y * per, c1 = min(n_out, c0 + per);
    #pragma unroll 4
    for (int c = c0; c < c1; ++c) out[(int64_t)c * P + p] = acc[(int64_t)c * P + p] / den;
}

}  // namespace slr

// ===========================================================================
// C ABI (include/slr_splat.h)
// ===========================================================================
using namespace slr;

static inline unsigned blocks_for(int64_t n) { return (unsigned)((n + kBlock - 1) / kBlock); }

// Enough channel chunks to put >= ~4 CTAs of work on every SM without making the
// per-chunk footprint recomputation dominant.
static int channel_chunks(int64_t pixel_blocks, int64_t C)
{
    const int sms = slr_host::sm_count();
    int chunks = 1;
    while (pixel_blocks * chunks < (int64_t)sms * 8 && chunks * 8 < C) chunks *= 2;
    return chunks;
}

extern "C" int slr_softsplat_sum_fwd(const float* in, const float* flow, float* out,
                                     int64_t B, int64_t C, int64_t H, int64_t W,
                                     int zero_out, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(in && flow && out && B > 0 && C > 0 && H > 0 && W > 0 && H * W < (1ll << 31) && B < 65536,
                   "slr_softsplat_sum_fwd: bad arguments");
    cudaStream_t s = (cudaStream_t)stream_;
    if (zero_out) SLR_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * B * C * H * W, s));
    const int chunks = channel_chunks(blocks_for(H * W) * B, C);
    const int per = (int)((C + chunks - 1) / chunks);
    dim3 grid(blocks_for(H * W), (unsigned)((C + per - 1) / per), (unsigned)B);
    softsplat_sum_fwd_kernel<<<grid, kBlock, 0, s>>>(in, flow, out, (int)C, (int)H, (int)W, per);
    return SLR_LAUNCH_STATUS();
}

extern "C" int slr_softsplat_grad_input(const float* flow, const float* gout, float* gin,
                                        int64_t B, int64_t C, int64_t H, int64_t W, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(flow && gout && gin && B > 0 && C > 0 && H > 0 && W > 0 && H * W < (1ll << 31) && B < 65536,
                   "slr_softsplat_grad_input: bad arguments");
    cudaStream_t s = (cudaStream_t)stream_;
    const int chunks = channel_chunks(blocks_for(H * W) * B, C);
    const int per = (int)((C + chunks - 1) / chunks);
    dim3 grid(blocks_for(H * W), (unsigned)((C + per - 1) / per), (unsigned)B);
    softsplat_grad_input_kernel<<<grid, kBlock, 0, s>>>(flow, gout, gin, (int)C, (int)H, (int)W, per);
    return SLR_LAUNCH_STATUS();
}

extern "C" int slr_softsplat_grad_flow(const float* in, const float* flow, const float* gout,
                                       float* gflow, int64_t B, int64_t C, int64_t H, int64_t W,
                                       slr_stream_t stream_)
{
    SLR_CHECK_ARGS(in && flow && gout && gflow && B > 0 && C > 0 && H > 0 && W > 0 && H * W < (1ll << 31) && B < 65536,
                   "slr_softsplat_grad_flow: bad arguments");
    dim3 grid(blocks_for(H * W), 1, (unsigned)B);
    softsplat_grad_flow_kernel<<<grid, kBlock, 0, (cudaStream_t)stream_>>>(in, flow, gout, gflow, (int)C, (int)H, (int)W);
    return SLR_LAUNCH_STATUS();
}

extern "C" int slr_maxsplat_fwd(const float* in, const float* flow, float* out, float init,
                                int64_t B, int64_t C, int64_t H, int64_t W, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(in && flow && out && B > 0 && C > 0 && H > 0 && W > 0 && H * W < (1ll << 31) && B < 65536,
                   "slr_maxsplat_fwd: bad arguments");
    cudaStream_t s = (cudaStream_t)stream_;
    const int64_t n = B * C * H * W;
    fill_kernel<<<(unsigned)std::min<int64_t>(blocks_for(n), (int64_t)slr_host::sm_count() * 16), kBlock, 0, s>>>(out, init, n);
    dim3 grid(blocks_for(H * W), 1, (unsigned)B);
    maxsplat_fwd_kernel<<<grid, kBlock, 0, s>>>(in, flow, out, (int)C, (int)H, (int)W);
    return SLR_LAUNCH_STATUS();
}

extern "C" int slr_maxwarpnorm(const float* in, const float* flow, float* scratch, float* out,
                               int64_t B, int64_t C, int64_t H, int64_t W, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(scratch && out, "slr_maxwarpnorm: bad arguments");
    int rc = slr_maxsplat_fwd(in, flow, scratch, -1000.0f, B, C, H, W, stream_);
    if (rc) return rc;
    dim3 grid(blocks_for(H * W), 1, (unsigned)B);
    inversesplat_kernel<<<grid, kBlock, 0, (cudaStream_t)stream_>>>(in, scratch, flow, out, (int)C, (int)H, (int)W);
    return SLR_LAUNCH_STATUS();
}

extern "C" int slr_euler(const float* motion, float sign, int T, float* disp, float* visible,
                         int64_t H, int64_t W, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(motion && disp && T >= 0 && H > 0 && W > 0 && H * W < (1ll << 31) && (sign == 1.0f || sign == -1.0f),
                   "slr_euler: bad arguments");
    euler_kernel<<<blocks_for(H * W), kBlock, 0, (cudaStream_t)stream_>>>(motion, sign, T, disp, visible, (int)H, (int)W);
    return SLR_LAUNCH_STATUS();
}

extern "C" int slr_euler_grad_motion(const float* motion, float sign, int T, const float* grad_disp,
                                     float* grad_motion, int64_t H, int64_t W, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(motion && grad_disp && grad_motion && T >= 0 && H > 0 && W > 0 && H * W < (1ll << 31) &&
                   (sign == 1.0f || sign == -1.0f), "slr_euler_grad_motion: bad arguments");
    cudaStream_t s = (cudaStream_t)stream_;
    SLR_CUDA(cudaMemsetAsync(grad_motion, 0, sizeof(float) * 2 * (size_t)(H * W), s));
    if (T > 0)
        euler_grad_kernel<<<blocks_for(H * W), kBlock, 0, s>>>(motion, sign, T, grad_disp, grad_motion, (int)H, (int)W);
    return SLR_LAUNCH_STATUS();
}

extern "C" int slr_reduce_max(const float* x, int64_t n, float* out_scalar, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(x && out_scalar && n > 0, "slr_reduce_max: bad arguments");
    cudaStream_t s = (cudaStream_t)stream_;
    fill_kernel<<<1, kBlock, 0, s>>>(out_scalar, -INFINITY, 1);
    const unsigned grid = (unsigned)std::min<int64_t>(blocks_for(n), (int64_t)slr_host::sm_count() * 4);
    reduce_max_kernel<<<grid, kBlock, 0, s>>>(x, n, out_scalar);
    return SLR_LAUNCH_STATUS();
}

extern "C" int slr_joint_scatter_weights(const float* feat, const float* z, const float* zsub,
                                         const float* tail, int n_tail,
                                         const float* disp_f, const float* disp_b, float w_fwd, float w_bwd,
                                         float* acc, int64_t C, int64_t H, int64_t W, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(feat && z && disp_f && disp_b && acc && C > 0 && H > 0 && W > 0 && H * W < (1ll << 31) &&
                   n_tail >= 0 && (n_tail == 0 || tail), "slr_joint_scatter: bad arguments");
    cudaStream_t s = (cudaStream_t)stream_;
    SLR_CUDA(cudaMemsetAsync(acc, 0, sizeof(float) * (C + n_tail + 1) * H * W, s));
    const int chunks = channel_chunks(blocks_for(H * W), C);
    const int per = (int)((C + chunks - 1) / chunks);
    dim3 grid(blocks_for(H * W), (unsigned)((C + per - 1) / per), 1);
    joint_scatter_kernel<<<grid, kBlock, 0, s>>>(feat, z, zsub, tail, n_tail, disp_f, disp_b, w_fwd, w_bwd, acc,
                                                 (int)C, (int)H, (int)W, per);
    return SLR_LAUNCH_STATUS();
}

extern "C" int slr_joint_scatter(const float* feat, const float* z, const float* zsub,
                                 const float* tail, int n_tail,
                                 const float* disp_f, const float* disp_b, float alpha,
                                 float* acc, int64_t C, int64_t H, int64_t W, slr_stream_t stream_)
{
    return slr_joint_scatter_weights(feat, z, zsub, tail, n_tail, disp_f, disp_b, alpha, 1.0f - alpha, acc, C, H, W, stream_);
}

extern "C" int slr_normalize(const float* acc, float* out, float* mask, int64_t n_out, int64_t norm_ch,
                             int64_t n_acc, float eps, int64_t H, int64_t W, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(acc && out && n_out > 0 && n_out <= n_acc && norm_ch >= 0 && norm_ch < n_acc && H > 0 && W > 0,
                   "slr_normalize: bad arguments");
    const int64_t P = H * W;
    int chunks = channel_chunks(blocks_for(P), n_out);
    dim3 grid(blocks_for(P), (unsigned)chunks, 1);
    normalize_kernel<<<grid, kBlock, 0, (cudaStream_t)stream_>>>(acc, out, mask, (int)n_out, (int)norm_ch, eps, P);
    return SLR_LAUNCH_STATUS();
}
