// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/slr_splat.h"

namespace slr_host {
// Records `msg` as the calling thread's last error and returns `code`.
int fail(int code, const char* msg);
// SM count of the current device (cached per device).
int sm_count();
// Which gather runs (environment variable SLR_GATHER_MODE, read at every call): "ldg" (default) = rowgather_kernel
// for every tile; "staged" = stagegather_kernel (sources staged in shared memory by the TMA unit) with
// rowgather_kernel for the tiles that do not fit.  Measured on B200 at 768x1024x64, motion A (profiles/r02): the
// staged gather's main loop is ~30 % faster, but the staging plan (in expand_kernel), the copy issue and the
// per-chunk synchronisation cost more than that saves; see DESIGN.md 4.2.
bool gather_staged();
// How the per-lane source lists of a batch are built (same variable): by default ("ldg") every moving source
// pixel writes its pairs straight into the lists of the destination lanes (insert_kernel, csrc/clip_gather.cu);
// "bins" and "staged" sort the sources into per-tile bins first and expand them tile by tile (bin_fill_kernel +
// expand_kernel: the round-1 pipeline, which the staging plan of "staged" is built on).
bool index_direct();
// slr_scene_prep without requiring an importance plane (z == NULL: e^Z = 1), and the one-frame table of a given
// flow: the pieces slr_softsplat_sum_fwd_gather puts in front of the clip pipeline (csrc/clip_plan.cu).
int scene_prep(const float* feat, const float* z, const float* zsub, const float* tail, int n_tail, void* scene,
               int64_t C, int64_t H, int64_t W, slr_stream_t stream);
int flow_table(const float* flow, int64_t H, int64_t W, void* table, size_t table_bytes, slr_stream_t stream);
}  // namespace slr_host

#define SLR_CHECK_ARGS(cond, msg) \
    do { if (!(cond)) return slr_host::fail(-1, msg); } while (0)

#define SLR_CUDA(expr) \
    do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return slr_host::fail((int)e_, cudaGetErrorString(e_)); } while (0)

// After a kernel launch: report launch-configuration errors without synchronising.
#define SLR_LAUNCH_STATUS() \
    ([]() -> int { cudaError_t e_ = cudaGetLastError(); \
                   return e_ == cudaSuccess ? 0 : slr_host::fail((int)e_, cudaGetErrorString(e_)); })()
