/*
 * slr_oracle.c -- CPU restatement of the SLR-SFS frame-synthesis hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (slr-sfs_b200/) may
 * link, import or call this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker.
 *
 * Every function restates arithmetic of the reference (paths relative to
 * /root/reference):
 *   splat scatter      models/softsplat.py:164-200  (kernel_Softsplat_updateOutput)
 *   grad wrt input     models/softsplat.py:213-253  (kernel_Softsplat_updateGradInput)
 *   grad wrt flow      models/softsplat.py:268-324  (kernel_Softsplat_updateGradFlow)
 *   max splat          models/softsplat.py:19-80    (kernel_Maximumsplat_updateOutput)
 *   inverse max gather models/softsplat.py:91-153   (kernel_Inversesplat_updateOutput)
 *   Euler integration  models/projection/euler_integration_manipulator.py:18-56
 *
 * Parity pinning: oracle/build_ref.py compiles the reference's own kernel text
 * (extracted at build time from models/softsplat.py, never copied into this
 * repo) for the CPU into oracle/_ref/, and tests/golden/ holds vectors produced
 * by that build and by the reference's own euler_integration() run on CPU
 * (tests/golden/make_golden.py).  tests/test_oracle_*.py check this file
 * against both.
 *
 * All tensors are fp32, contiguous NCHW; flow channel 0 = x, 1 = y, in pixels.
 * Compile with -ffp-contract=off so products are rounded before they are added,
 * which is what the reference scatter does (the product is an atomicAdd operand).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef int64_t i64;

/* The 2x2 landing footprint of one source pixel. */
typedef struct {
    int x0, y0;          /* north-west cell */
    float w[4];          /* NW, NE, SW, SE bilinear weights */
    int ok[4];           /* corner inside the frame? */
} footprint_t;

/* softsplat.py:166-200: landing position = pixel + flow, NW = floor, weights are
 * products of the distances to the opposite corner, each corner bounds-tested. */
static footprint_t landing(float x, float y, float fx, float fy, i64 H, i64 W)
{
    footprint_t f;
    float ox = x + fx, oy = y + fy;
    float flx = floorf(ox), fly = floorf(oy);
    /* keep the int conversion defined for huge / non-finite flow: such pixels
     * are outside every frame either way */
    int far = !(flx > -2.0f && flx < (float)W + 1.0f && fly > -2.0f && fly < (float)H + 1.0f);
    int ix = far ? -4 : (int)flx, iy = far ? -4 : (int)fly;
    float ex = (float)(ix + 1), ey = (float)(iy + 1), bx = (float)ix, by = (float)iy;
    f.x0 = ix; f.y0 = iy;
    f.w[0] = (ex - ox) * (ey - oy);
    f.w[1] = (ox - bx) * (ey - oy);
    f.w[2] = (ex - ox) * (oy - by);
    f.w[3] = (ox - bx) * (oy - by);
    for (int k = 0; k < 4; ++k) {
        int cx = ix + (k & 1), cy = iy + (k >> 1);
        f.ok[k] = !far && cx >= 0 && cx < W && cy >= 0 && cy < H;
    }
    return f;
}

/* out must be zeroed (or hold a previous accumulation) by the caller. */
void orc_softsplat_sum_fwd(const float *in, const float *flow, float *out,
                           i64 B, i64 C, i64 H, i64 W)
{
    const i64 P = H * W;
    for (i64 b = 0; b < B; ++b)
        for (i64 y = 0; y < H; ++y)
            for (i64 x = 0; x < W; ++x) {
                const i64 p = y * W + x;
                footprint_t f = landing((float)x, (float)y, flow[(b * 2 + 0) * P + p],
                                        flow[(b * 2 + 1) * P + p], H, W);
                for (int k = 0; k < 4; ++k) {
                    if (!f.ok[k]) continue;
                    const i64 q = (i64)(f.y0 + (k >> 1)) * W + (f.x0 + (k & 1));
                    for (i64 c = 0; c < C; ++c)
                        out[(b * C + c) * P + q] += in[(b * C + c) * P + p] * f.w[k];
                }
            }
}

/* Same scatter with double accumulators (products still rounded to fp32 as in the
 * reference).  Used by the tests to measure how far any fp32 summation order can
 * drift, i.e. to justify the tolerance. */
void orc_softsplat_sum_fwd_f64acc(const float *in, const float *flow, double *out,
                                  i64 B, i64 C, i64 H, i64 W)
{
    const i64 P = H * W;
    for (i64 b = 0; b < B; ++b)
        for (i64 y = 0; y < H; ++y)
            for (i64 x = 0; x < W; ++x) {
                const i64 p = y * W + x;
                footprint_t f = landing((float)x, (float)y, flow[(b * 2 + 0) * P + p],
                                        flow[(b * 2 + 1) * P + p], H, W);
                for (int k = 0; k < 4; ++k) {
                    if (!f.ok[k]) continue;
                    const i64 q = (i64)(f.y0 + (k >> 1)) * W + (f.x0 + (k & 1));
                    for (i64 c = 0; c < C; ++c)
                        out[(b * C + c) * P + q] += (double)(in[(b * C + c) * P + p] * f.w[k]);
                }
            }
}

/* softsplat.py:213-253: gin = sum over in-bounds corners of gout[corner] * w. */
void orc_softsplat_grad_input(const float *flow, const float *gout, float *gin,
                              i64 B, i64 C, i64 H, i64 W)
{
    const i64 P = H * W;
    for (i64 b = 0; b < B; ++b)
        for (i64 y = 0; y < H; ++y)
            for (i64 x = 0; x < W; ++x) {
                const i64 p = y * W + x;
                footprint_t f = landing((float)x, (float)y, flow[(b * 2 + 0) * P + p],
                                        flow[(b * 2 + 1) * P + p], H, W);
                for (i64 c = 0; c < C; ++c) {
                    float g = 0.0f;
                    for (int k = 0; k < 4; ++k) {
                        if (!f.ok[k]) continue;
                        const i64 q = (i64)(f.y0 + (k >> 1)) * W + (f.x0 + (k & 1));
                        g += gout[(b * C + c) * P + q] * f.w[k];
                    }
                    gin[(b * C + c) * P + p] = g;
                }
            }
}

/* softsplat.py:268-324: d(weight)/d(flow_x) and /d(flow_y) per corner, then
 * sum_c in[c] * gout[c, corner] * dweight, channels outermost-in-time as in the
 * reference loop (:304-322). */
void orc_softsplat_grad_flow(const float *in, const float *flow, const float *gout,
                             float *gflow, i64 B, i64 C, i64 H, i64 W)
{
    const i64 P = H * W;
    for (i64 b = 0; b < B; ++b)
        for (i64 y = 0; y < H; ++y)
            for (i64 x = 0; x < W; ++x) {
                const i64 p = y * W + x;
                const float fx = flow[(b * 2 + 0) * P + p], fy = flow[(b * 2 + 1) * P + p];
                footprint_t f = landing((float)x, (float)y, fx, fy, H, W);
                const float ox = (float)x + fx, oy = (float)y + fy;
                const float bx = (float)f.x0, by = (float)f.y0;
                const float ex = (float)(f.x0 + 1), ey = (float)(f.y0 + 1);
                float dw[2][4];
                dw[0][0] = -1.0f * (ey - oy); dw[0][1] = +1.0f * (ey - oy);
                dw[0][2] = -1.0f * (oy - by); dw[0][3] = +1.0f * (oy - by);
                dw[1][0] = (ex - ox) * -1.0f; dw[1][1] = (ox - bx) * -1.0f;
                dw[1][2] = (ex - ox) * +1.0f; dw[1][3] = (ox - bx) * +1.0f;
                for (int d = 0; d < 2; ++d) {
                    float g = 0.0f;
                    for (i64 c = 0; c < C; ++c) {
                        const float v = in[(b * C + c) * P + p];
                        for (int k = 0; k < 4; ++k) {
                            if (!f.ok[k]) continue;
                            const i64 q = (i64)(f.y0 + (k >> 1)) * W + (f.x0 + (k & 1));
                            g += v * gout[(b * C + c) * P + q] * dw[d][k];
                        }
                    }
                    gflow[(b * 2 + d) * P + p] = g;
                }
            }
}

/* softsplat.py:19-80: out[corner] = max(out[corner], in * w).  out is
 * pre-initialised by the caller (-1000 in _FunctionMaximumWarpNormsplat, :590). */
void orc_maxsplat_fwd(const float *in, const float *flow, float *out,
                      i64 B, i64 C, i64 H, i64 W)
{
    const i64 P = H * W;
    for (i64 b = 0; b < B; ++b)
        for (i64 y = 0; y < H; ++y)
            for (i64 x = 0; x < W; ++x) {
                const i64 p = y * W + x;
                footprint_t f = landing((float)x, (float)y, flow[(b * 2 + 0) * P + p],
                                        flow[(b * 2 + 1) * P + p], H, W);
                for (int k = 0; k < 4; ++k) {
                    if (!f.ok[k]) continue;
                    const i64 q = (i64)(f.y0 + (k >> 1)) * W + (f.x0 + (k & 1));
                    for (i64 c = 0; c < C; ++c) {
                        float *o = &out[(b * C + c) * P + q];
                        *o = fmaxf(in[(b * C + c) * P + p] * f.w[k], *o);
                    }
                }
            }
}

/* softsplat.py:91-153: out[p] = max(out[p], maxwarped[corner]) over the
 * in-bounds corners of p's own footprint.  out is pre-initialised by the caller
 * (a clone of the splatted tensor, :607). */
void orc_inversesplat(const float *maxwarped, const float *flow, float *out,
                      i64 B, i64 C, i64 H, i64 W)
{
    const i64 P = H * W;
    for (i64 b = 0; b < B; ++b)
        for (i64 y = 0; y < H; ++y)
            for (i64 x = 0; x < W; ++x) {
                const i64 p = y * W + x;
                footprint_t f = landing((float)x, (float)y, flow[(b * 2 + 0) * P + p],
                                        flow[(b * 2 + 1) * P + p], H, W);
                for (i64 c = 0; c < C; ++c) {
                    float m = out[(b * C + c) * P + p];
                    for (int k = 0; k < 4; ++k) {
                        if (!f.ok[k]) continue;
                        const i64 q = (i64)(f.y0 + (k >> 1)) * W + (f.x0 + (k & 1));
                        m = fmaxf(maxwarped[(b * C + c) * P + q], m);
                    }
                    out[(b * C + c) * P + p] = m;
                }
            }
}

/* euler_integration_manipulator.py:18-56 for one [1,2,H,W] motion field.
 * Every pixel's chain is independent because the motion field is constant:
 *   dest <- dest + motion[:, rne(dest_y), rne(dest_x)]            (:37-38)
 *   out-of-bounds test is strict (> W-1, < 0)                       (:39-40)
 *   invalidity is sticky                                            (:41-42)
 *   invalid pixels restart from their own coordinate every step     (:45-46)
 *   result = dest - coord, invalid -> max(H,W)+1 in both channels   (:53-55)
 * T == 0 returns zeros / all visible (:33-34).  rintf() rounds half to even
 * under the default rounding mode, like torch.round. */
void orc_euler(const float *motion, int T, float *disp, float *visible, i64 H, i64 W)
{
    const i64 P = H * W;
    const float sentinel = (float)((H > W ? H : W) + 1);
    for (i64 y = 0; y < H; ++y)
        for (i64 x = 0; x < W; ++x) {
            const float cx = (float)x, cy = (float)y;
            float dx = cx, dy = cy;
            int invalid = 0;
            for (int k = 1; k <= T; ++k) {
                const i64 ix = (i64)rintf(dx), iy = (i64)rintf(dy);
                const float mx = motion[0 * P + iy * W + ix];
                const float my = motion[1 * P + iy * W + ix];
                dx = dx + mx;
                dy = dy + my;
                if (dx > (float)(W - 1) || dx < 0.0f || dy > (float)(H - 1) || dy < 0.0f)
                    invalid = 1;
                if (invalid) { dx = cx; dy = cy; }
            }
            const i64 p = y * W + x;
            disp[0 * P + p] = invalid ? sentinel : dx - cx;
            disp[1 * P + p] = invalid ? sentinel : dy - cy;
            visible[p] = invalid ? 0.0f : 1.0f;
        }
}

/* Gradient of orc_euler's displacements with respect to the motion field, as torch autograd
 * derives it from the reference's eager ops (euler_integration_manipulator.py:36-55):
 *   destination_coords + motion[0][:, iy, ix]     differentiable in the sampled VALUES only (:37-38)
 *   destination_coords[invalid] = coord[invalid]  cuts the chain of a pixel once it is invalid (:45-46)
 *   displacements[invalid] = max(H,W)+1           a constant: no gradient for pixels that end invalid (:53-55)
 * Invalidity is sticky, so a pixel valid at the end was valid throughout and every one of its T
 * samples receives grad_disp[:, p].  grad_motion is overwritten. */
void orc_euler_grad_motion(const float *motion, int T, const float *gdisp, float *gmotion, i64 H, i64 W)
{
    const i64 P = H * W;
    for (i64 i = 0; i < 2 * P; ++i) gmotion[i] = 0.0f;
    for (i64 y = 0; y < H; ++y)
        for (i64 x = 0; x < W; ++x) {
            const float cx = (float)x, cy = (float)y;
            float dx = cx, dy = cy;
            int invalid = 0;
            for (int k = 1; k <= T && !invalid; ++k) {
                const i64 ix = (i64)rintf(dx), iy = (i64)rintf(dy);
                dx = dx + motion[0 * P + iy * W + ix];
                dy = dy + motion[1 * P + iy * W + ix];
                if (dx > (float)(W - 1) || dx < 0.0f || dy > (float)(H - 1) || dy < 0.0f) invalid = 1;
            }
            if (invalid) continue;
            const i64 p = y * W + x;
            dx = cx; dy = cy;
            for (int k = 1; k <= T; ++k) {
                const i64 ix = (i64)rintf(dx), iy = (i64)rintf(dy);
                gmotion[0 * P + iy * W + ix] += gdisp[0 * P + p];
                gmotion[1 * P + iy * W + ix] += gdisp[1 * P + p];
                dx = dx + motion[0 * P + iy * W + ix];
                dy = dy + motion[1 * P + iy * W + ix];
            }
        }
}
