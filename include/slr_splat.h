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

/* Backward of slr_euler with respect to the motion field (the reference's chain is
 * differentiable through the sampled values, euler_integration_manipulator.py:37-38, which is
 * what --train_motion relies on: models/animating_softmax_splating.py:514-580).  grad_motion
 * [2,H,W] is overwritten with
 *   sum over VALID chains p and their steps k of sign * grad_disp[:, p] at the sample visited in step k;
 * chains that end invalid carry the sentinel constant (:53-55) and contribute nothing. */
int slr_euler_grad_motion(const float* motion, float sign, int T, const float* grad_disp,
                          float* grad_motion, int64_t H, int64_t W, slr_stream_t stream);

/* The same summation splat (softsplat.py:157-202; _FunctionSoftsplat.forward, :407-416) through the GATHER
 * pipeline of the joint block instead of fp32 atomics: per batch element the input is interleaved, the flow
 * becomes a one-frame table, every source writes its list cells and every destination pixel pulls its
 * contributions and writes the un-normalised sum once (output fully overwritten, exact zeros in holes).
 * Same results up to fp32 summation order.  Pays off for large frames (1 x 65 x 768 x 1024 with a smooth flow:
 * 0.29 ms against 0.40 ms, the reference's own kernel 0.41 ms on the same B200); the Python operator picks it
 * above 2^25 elements per batch item.  scratch: slr_softsplat_gather_scratch_bytes(C, H, W) bytes,
 * 256-byte aligned, reused for every batch element.  Needs the default SLR_GATHER_MODE (returns -1 / 0 bytes
 * otherwise) and H * W < 2^27. */
size_t slr_softsplat_gather_scratch_bytes(int64_t C, int64_t H, int64_t W);
int slr_softsplat_sum_fwd_gather(const float* input, const float* flow, float* output,
                                 int64_t B, int64_t C, int64_t H, int64_t W,
                                 void* scratch, size_t scratch_bytes, slr_stream_t stream);

/* ---------------------------------------------------------------------------
 * Training forward: the splat-input producer fused into the splat, and its backward.
 * Replaces, per direction, models/animating_softmax_splating.py:606 (tenInput_f = cat([start_fs *
 * Z_f_norm.exp() * alpha, Z_f_norm.exp() * alpha], 1); :651 for the backward direction with end_fs, Z_p and
 * 1 - alpha), the ModuleSoftsplat('summation') call on it (:629-632 / :672-676 -> softsplat.py:157-202) and,
 * in the backward pass, autograd through both (softsplat.py:204-326, :427-477 and the cat / exp / mul nodes).
 *   fs [B,C,H,W], zn [B,1,H,W] (the normalised, clamped importance: :596-605), flow [B,2,H,W] (the
 *   integrated displacement), alpha [B] on the device (:584-585).
 * slr_producer_splat_fwd: acc [B,C+1,H,W] (+)= splat of (fs * e^zn * alpha, e^zn * alpha); accumulate != 0
 *   adds into acc (the second direction: the reference adds the two splat outputs, :684-686), else acc is
 *   overwritten.
 * slr_producer_splat_bwd: from grad_acc [B,C+1,H,W] the gradients d_fs [B,C,H,W], d_zn [B,1,H,W], d_flow
 *   [B,2,H,W]; any of the three may be NULL. */
int slr_producer_splat_fwd(const float* fs, const float* zn, const float* flow, const float* alpha,
                           float* acc, int64_t B, int64_t C, int64_t H, int64_t W, int accumulate,
                           slr_stream_t stream);
int slr_producer_splat_bwd(const float* fs, const float* zn, const float* flow, const float* alpha,
                           const float* grad_acc, float* d_fs, float* d_zn, float* d_flow,
                           int64_t B, int64_t C, int64_t H, int64_t W, slr_stream_t stream);

/* ---------------------------------------------------------------------------
 * Joint block level: what forward_flow() does between the encoder and the
 * decoder (models/animating_softmax_splating.py:847-924 and
 * models/animating_softmax_splating_2layers_alpha_seperate.py:921-1045).
 * ------------------------------------------------------------------------- */

/* max over a tensor into a device scalar (Z.max(), animating_softmax_splating.py:855). */
int slr_reduce_max(const float* x, int64_t n, float* out_scalar, slr_stream_t stream);

/* Scatter variant of the joint block ("algorithm A": the design BASELINE.json's
 * north_star sketches; kept as the measured baseline of the gather pipeline
 * below).  acc is [C + n_tail + 1, H, W] and is cleared first; for every source
 * pixel and both directions d in {forward, backward}, with a_fwd = alpha,
 * a_bwd = 1 - alpha and displacement fields disp_f / disp_b ([2,H,W], e.g. from
 * slr_euler):
 *   acc[c]          += ((feat[c] * e^(Z - *zsub)) * a_d) * w     c < C
 *   acc[C + j]      += (tail[j] * a_d) * w                        j < n_tail
 *   acc[C + n_tail] += (e^(Z - *zsub) * a_d) * w                  (the normaliser)
 * zsub points at a device scalar (Z.max(), animating_softmax_splating.py:855) or
 * is NULL (use_softmax_splatter_v1, :853).  tail ([n_tail,H,W], may be NULL) holds
 * the 2-layer model's extra, already weighted channels (a_f*e^A and e^A,
 * 2layers...py:967-972, or a_f*e^Z, :974-976). */
int slr_joint_scatter(const float* feat, const float* z, const float* zsub,
                      const float* tail, int n_tail,
                      const float* disp_f, const float* disp_b, float alpha,
                      float* acc, int64_t C, int64_t H, int64_t W, slr_stream_t stream);

/* The same with the two directions' blend weights given separately (a_fwd = w_fwd, a_bwd = w_bwd)
 * instead of (alpha, 1 - alpha): AnimatingSoftmaxSplating.warp_flow
 * (models/animating_softmax_splating.py:1064-1138) splats an RGB image with Z = 1, weights the forward
 * direction by e^(Z - max) * alpha and the backward one by e^Z * (1 - alpha), alpha without the "+ 1"
 * of forward_flow (:1064), and takes precomputed displacement fields. */
int slr_joint_scatter_weights(const float* feat, const float* z, const float* zsub,
                              const float* tail, int n_tail,
                              const float* disp_f, const float* disp_b, float w_fwd, float w_bwd,
                              float* acc, int64_t C, int64_t H, int64_t W, slr_stream_t stream);

/* Normalise-and-blend: out[c] = acc[c] / max(acc[norm_ch], eps) for c < n_out
 * (animating_softmax_splating.py:923-924; eps = 1e-8).  mask (optional, [H,W]) =
 * acc[norm_ch] > eps (2layers...py:1039).  Exact-zero holes stay exactly 0. */
int slr_normalize(const float* acc, float* out, float* mask, int64_t n_out, int64_t norm_ch,
                  int64_t n_acc, float eps, int64_t H, int64_t W, slr_stream_t stream);

/* ---------------------------------------------------------------------------
 * Clip level ("algorithm G", the fast path): bin source pixels by destination
 * tile, then gather with register accumulators and write normalised frames
 * directly.  Same results as slr_euler x2 + slr_joint_scatter + slr_normalize
 * up to fp32 summation order.  The unit of work is the reference's frame loop
 * (test_animating/test_v1_4eval_rawsize.py:233-239): for t in range(N):
 * index = [start, t, end]; forward_flow(batch).
 * ------------------------------------------------------------------------- */

/* Bytes of the per-scene buffer slr_scene_prep fills (16-byte aligned storage). */
size_t slr_scene_bytes(int64_t C, int n_tail, int64_t H, int64_t W);

/* Bytes of the leading part of that buffer from which slr_scene_quilt derives the rest: what has to
 * travel when a prepared scene is broadcast to other GPUs (BASELINE.json configs[3]). */
size_t slr_scene_core_bytes(int64_t C, int n_tail, int64_t H, int64_t W);

/* Rebuilds the trailing (TMA-friendly, pixel-major, bank-swizzled) copy of the features inside a
 * scene buffer from its leading part.  slr_scene_prep does this itself; a rank that received only
 * slr_scene_core_bytes of a scene calls it once before synthesising frames. */
int slr_scene_quilt(void* scene, int64_t C, int n_tail, int64_t H, int64_t W, slr_stream_t stream);

/* Once per scene (features, Z and motion are constant over the clip): writes the
 * pre-weighted, channel-interleaved features feat[c]*e^(Z - *zsub) and the scalar
 * planes (tail..., e^(Z - *zsub)) into `scene` (16-byte aligned,
 * slr_scene_bytes).  zsub / tail as in slr_joint_scatter; n_tail <= 2. */
int slr_scene_prep(const float* feat, const float* z, const float* zsub,
                   const float* tail, int n_tail, void* scene,
                   int64_t C, int64_t H, int64_t W, slr_stream_t stream);

/* Bytes of scratch slr_clip_frames needs for a batch of n_frames. */
size_t slr_clip_workspace_bytes(int64_t H, int64_t W, int n_frames);

/* Synthesise frames t0 .. t0+n_frames-1 (n_frames <= 64) of the clip whose index
 * triplets are [start, t, end]: forward displacement = t - start Euler steps of
 * +motion, backward = end - t + 1 steps of -motion
 * (animating_softmax_splating.py:847-848), alpha = 1 - (t-start)/(end-start+1)
 * clamped to [alpha_lo, alpha_hi] (:860; 2layers...py:952 clamps to
 * [1/600, 599/600], the baseline model does not clamp: pass 0 and 1).
 *   out  [n_frames, C, H, W]  features / max(norm, 1e-8)           (:923-924)
 *   aux  [n_frames, n_tail + 1, H, W] or NULL: raw sums of the tail channels and
 *        the norm (the 2-layer model divides them itself, 2layers...py:1038-1045)
 *   mask [n_frames, H, W] or NULL: norm > 1e-8                      (2layers...py:1039)
 *   nnz  [n_frames, H, W] or NULL: number of channels of `out` that are != 0 at the pixel.  The
 *        decoder starts with mask = (x != 0) over all channels (networks/architectures.py:369) and its
 *        first partial convolution only ever uses that mask summed over the channels and the window
 *        (layers/partialconv2d.py:61-66): nnz is exactly the per-pixel channel sum, so neither the
 *        per-element mask nor the all-ones mask convolution over C channels has to be materialised
 * workspace: slr_clip_workspace_bytes(H, W, n_frames) bytes, 16-byte aligned.
 * slr_clip_frames = slr_clip_plan (Euler chains, landing table; with the bin pipeline also the
 * destination-tile bins; depends on the motion only), slr_clip_expand (per-lane source lists of every
 * destination row pair; depends on the motion and the blend weights.  Default: every moving source
 * pixel writes its list cells straight into the lists of the destination lanes -- the "direct index";
 * environment variable SLR_GATHER_MODE=bins|staged: from per-tile bins, tile by tile),
 * slr_clip_gather (the gather itself: every tile whose lists fit) and slr_clip_heavy
 * (the few tiles in convergence zones of the flow whose lists do not fit, by fp32
 * reductions at L2; touches only those tiles) on the same workspace and the same motion
 * (pixels whose motion is exactly zero are not binned; their self-contribution is
 * added implicitly).  The three steps may be issued on different streams with the
 * obvious dependencies, e.g. plan + expand of the next batch beside the gather of
 * the current one (each batch needs its own workspace then). */
int slr_clip_plan(const float* motion, int64_t H, int64_t W, int start, int end, int t0,
                  int n_frames, void* workspace, size_t workspace_bytes, slr_stream_t stream);
int slr_clip_expand(const void* scene, const float* motion, int64_t C, int n_tail, int64_t H, int64_t W,
                    int start, int end, int t0, int n_frames, float alpha_lo, float alpha_hi,
                    void* workspace, size_t workspace_bytes, slr_stream_t stream);
int slr_clip_gather(const void* scene, const float* motion, int64_t C, int n_tail, int64_t H, int64_t W,
                    int start, int end, int t0, int n_frames, float alpha_lo, float alpha_hi,
                    float* out, float* aux, float* mask, float* nnz,
                    const void* workspace, size_t workspace_bytes, slr_stream_t stream);
int slr_clip_heavy(const void* scene, const float* motion, int64_t C, int n_tail, int64_t H, int64_t W,
                   int start, int end, int t0, int n_frames, float alpha_lo, float alpha_hi,
                   float* out, float* aux, float* mask, float* nnz,
                   const void* workspace, size_t workspace_bytes, slr_stream_t stream);
int slr_clip_frames(const void* scene, const float* motion, int64_t C, int n_tail,
                    int64_t H, int64_t W, int start, int end, int t0, int n_frames,
                    float alpha_lo, float alpha_hi, float* out, float* aux, float* mask, float* nnz,
                    void* workspace, size_t workspace_bytes, slr_stream_t stream);

/* Clip table: slr_clip_plan split in two, so that the part that depends on the motion only --
 * both Euler chains, the landing coordinates and the bin sizes of every frame -- is computed ONCE
 * for a whole run of frames (one forward and one backward chain: t0+n-1-start and end-t0+1 steps
 * for the run, where per-batch plans re-integrate from zero for every batch, as the reference
 * does for every frame, euler_integration_manipulator.py:36) and the batches are cut from it:
 *   slr_clip_table(motion, ..., t0, n_frames, table)              n_frames <= 4096, table of
 *                                                                 slr_clip_table_bytes(H, W, n_frames)
 *   slr_clip_bin(table, ..., table_frames, f0, n, workspace)      bins of frames f0 .. f0+n-1 of the
 *                                                                 table (n <= 64) into a batch workspace
 * after which slr_clip_expand / slr_clip_gather / slr_clip_heavy run on that workspace with
 * t0 = (the table's t0) + f0 exactly as after slr_clip_plan.  With the direct index (default)
 * slr_clip_bin does not copy anything out of the table: the workspace REFERS to the table's landing
 * coordinates, so the table must stay valid and unchanged until the batch's slr_clip_heavy has run
 * (the bin pipeline only needs it until slr_clip_bin has run).  A table and the workspaces cut from
 * it must be used under the same SLR_GATHER_MODE they were built under. */
size_t slr_clip_table_bytes(int64_t H, int64_t W, int n_frames);
int slr_clip_table(const float* motion, int64_t H, int64_t W, int start, int end, int t0,
                   int n_frames, void* table, size_t table_bytes, slr_stream_t stream);
int slr_clip_bin(const void* table, size_t table_bytes, int64_t H, int64_t W, int table_frames,
                 int f0, int n_frames, void* workspace, size_t workspace_bytes, slr_stream_t stream);

/* Diagnostics (a `_host` call: copies counters and the tile flags to the host and synchronises
 * `stream`).  After slr_clip_expand on `workspace`: stats[0] = flagged destination tiles of the batch
 * (some lane's source list was cut at the list depth), stats[1] = of those, tiles done entirely by
 * per-pair reductions from their bins (bin pipeline only), stats[2] = (destination, source) pairs beyond
 * the list depth, stats[3] = capacity of that excess list (direct index: stats[2] > stats[3] means the
 * whole batch was redone by scatter + divide inside slr_clip_heavy), stats[4] (valid after slr_clip_gather) = tiles whose
 * sources did not fit the shared-memory staging area and were gathered through L1 instead,
 * stats[5] = tiles x frames of the batch. */
int slr_clip_stats_host(const void* workspace, size_t workspace_bytes, int64_t H, int64_t W, int n_frames,
                        uint32_t stats[6], slr_stream_t stream);

/* ---------------------------------------------------------------------------
 * Frame sink (after the decoder): what test_animating/test_v1_4eval_rawsize.py:240-242,284-286 does per
 * frame on the host -- bilinear resize to the raw size (align_corners=False), (x * mul + add) * 255
 * (mul = add = 0.5 for images in [-1, 1]; mul = 1, add = 0 for alpha maps, :245), round to nearest even,
 * saturate to 0..255, RGB -> BGR (swap_rb) and interleave -- as one kernel: frames [n,3,H,W] fp32 ->
 * out [n,out_H,out_W,3] uint8.  3 bytes per pixel leave the device instead of 12.
 * ------------------------------------------------------------------------- */
int slr_frame_sink_u8(const float* frames, uint8_t* out, int64_t n, int64_t H, int64_t W,
                      int64_t out_H, int64_t out_W, float mul, float add, int swap_rb, slr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SLR_SPLAT_H */
