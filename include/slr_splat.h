/*
 * slr_splat.h -- C ABI of libslr_splat.so, the B200 (sm_100a) replacement for the
 * cupy/NVRTC kernels of SLR-SFS's frame-synthesis hot path.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; device pointers unless the name ends in
 *     `_host`; fp32, contiguous NCHW; flow / motion channel 0 = x, 1 = y, pixels;
 *   - the caller owns every buffer; nothing is allocated, freed or synchronised
 *     behind the caller's back except where a function says so (`_host` calls);
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = the
 *     legacy default stream) of the CURRENT device;
 *   - return value 0 = success, otherwise a cudaError_t value (or -1 for an
 *     argument error); slr_last_error_string() describes the last failure on the
 *     calling thread.
 *
 * Paths cited below are relative to the reference tree (simon3dv/SLR-SFS).
 */
#ifndef SLR_SPLAT_H
#define SLR_SPLAT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* slr_stream_t; /* cudaStream_t */

/* Library version (major*10000 + minor*100 + patch) and last error text. */
int slr_version(void);
const char* slr_last_error_string(void);

/* ---------------------------------------------------------------------------
 * Operator level: one entry point per reference kernel launch.
 * ------------------------------------------------------------------------- */

/* Summation splat.  Replaces kernel_Softsplat_updateOutput as launched by
 * _FunctionSoftsplat.forward (models/softsplat.py:157-202, 390-423).
 * out[b,c,floor(y+fy)+dy,floor(x+fx)+dx] += in[b,c,y,x] * w(dy,dx) for the
 * in-bounds corners.  zero_out != 0: `out` is cleared first (the reference's
 * new_zeros, :404); zero_out == 0: accumulate into `out`. */
int slr_softsplat_sum_fwd(const float* in, const float* flow, float* out,
                          int64_t B, int64_t C, int64_t H, int64_t W,
                          int zero_out, slr_stream_t stream);

/* Gradient wrt the splatted tensor.  Replaces kernel_Softsplat_updateGradInput
 * (models/softsplat.py:204-255, launched :445-456). */
int slr_softsplat_grad_input(const float* flow, const float* grad_out, float* grad_in,
                             int64_t B, int64_t C, int64_t H, int64_t W, slr_stream_t stream);

/* Gradient wrt the flow.  Replaces kernel_Softsplat_updateGradFlow
 * (models/softsplat.py:257-326, launched :460-471). */
int slr_softsplat_grad_flow(const float* in, const float* flow, const float* grad_out,
                            float* grad_flow, int64_t B, int64_t C, int64_t H, int64_t W,
                            slr_stream_t stream);

/* Max splat: out = max over contributions of in*w, cells nobody reaches keep
 * `init`.  Replaces kernel_Maximumsplat_updateOutput (models/softsplat.py:12-82)
 * as launched by _FunctionMaximumsplat.forward (:482-518, init 0) and by
 * _FunctionMaximumWarpNormsplat (:590, init -1000). */
int slr_maxsplat_fwd(const float* in, const float* flow, float* out, float init,
                     int64_t B, int64_t C, int64_t H, int64_t W, slr_stream_t stream);

/* _FunctionMaximumWarpNormsplat (models/softsplat.py:576-624): max-splat `in`
 * into `scratch` (filled with -1000 here), then out[p] = max(in[p], scratch at
 * p's in-bounds corners) (kernel_Inversesplat_updateOutput, :84-155).
 * scratch and out are caller-provided [B,C,H,W]. */
int slr_maxwarpnorm(const float* in, const float* flow, float* scratch, float* out,
                    int64_t B, int64_t C, int64_t H, int64_t W, slr_stream_t stream);

/* euler_integration(motion, T) for one [1,2,H,W] field
 * (models/projection/euler_integration_manipulator.py:7-56).  `sign` (+1 / -1)
 * multiplies the motion (the models integrate -flow for the backward direction,
 * models/animating_softmax_splating.py:848).  Writes displacements [2,H,W] and,
 * when non-NULL, visible [H,W] (1 = still inside the frame).  Invalid pixels get
 * max(H,W)+1 in both displacement channels (:53-55).  No host synchronisation. */
int slr_euler(const float* motion, float sign, int T, float* disp, float* visible,
              int64_t H, int64_t W, slr_stream_t stream);

/* ---------------------------------------------------------------------------
 * Joint block level: what forward_flow() does between the encoder and the
 * decoder (models/animating_softmax_splating.py:847-924 and
 * models/animating_softmax_splating_2layers_alpha_seperate.py:921-1045).
 * ------------------------------------------------------------------------- */

/* max over a tensor into a device scalar (Z.max(), animating_softmax_splating.py:855). */
int slr_reduce_max(const float* x, int64_t n, float* out_scalar, slr_stream_t stream);

/* Scatter variant of the joint block ("algorithm A", kept as the measured
 * baseline for the gather pipeline below).  For every source pixel and both
 * directions d in {forward, backward}:
 *   acc[c]      += ((feat[c] * e^{Z - zsub}) * a_d) * w    c < C
 *   acc[C + j]  += ((extra_val[j] * e_j) * a_d) * w         j < n_extra   (2-layer alpha channels)
 *   acc[last]   += (e^{Z - zsub} * a_d) * w
 * with a_fwd = alpha, a_bwd = 1 - alpha, and displacement fields disp_f / disp_b
 * ([2,H,W] each, e.g. from slr_euler).  zsub points at a device scalar (Z.max())
 * or is NULL (no subtraction, use_softmax_splatter_v1).  `splat_in` optional
 * channels: `tail` is [n_tail,H,W] of already-weighted extra channels that are
 * multiplied by a_d and splatted as they are (the 2-layer model's
 * a_f*e^A, e^A planes, 2layers...py:967-972); they land in acc[C .. C+n_tail).
 * The e^Z weight plane lands in acc[C+n_tail].  acc is [C+n_tail+1,H,W] and is
 * cleared first. */
int slr_joint_scatter(const float* feat, const float* z, const float* zsub,
                      const float* tail, int n_tail,
                      const float* disp_f, const float* disp_b, float alpha,
                      float* acc, int64_t C, int64_t H, int64_t W, slr_stream_t stream);

/* Normalise-and-blend: out[c] = acc[c] / max(acc[norm_ch], eps) for c < n_out
 * (animating_softmax_splating.py:923-924; eps = 1e-8).  mask (optional, [H,W]) =
 * acc[norm_ch] > eps (2layers...py:1039).  Exact-zero holes stay exactly 0. */
int slr_normalize(const float* acc, float* out, float* mask, int64_t n_out, int64_t norm_ch,
                  int64_t n_acc, float eps, int64_t H, int64_t W, slr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SLR_SPLAT_H */
