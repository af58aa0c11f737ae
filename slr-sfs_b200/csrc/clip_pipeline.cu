// Clip pipeline ("algorithm G"): the joint block of the reference models
// (/root/reference/models/animating_softmax_splating.py:847-924) restructured
// for sm_100a as  bin-by-destination-tile  +  gather with register accumulators.
//
// Why not scatter: a frame at 768x1024x65 is 2 directions x 4 corners x 51 M
// elements = 409 M fp32 atomics.  The measured L2 reduction rate on B200
// (profiles/r01: 0.65 ms per frame) is 10x the HBM time of the same bytes.  The
// splat weights do not depend on the channel, so the scatter is a sparse
// (destination x source) matrix applied to all channels: we build that sparse
// structure once per frame from the displacement alone (2 planes, cheap) and then
// every destination pixel PULLS its contributions, accumulates them in registers
// and writes the normalised result once.  No feature atomics, no accumulator
// round trip through HBM, no separate normalise pass.
//
//   scene_prep      once per scene: G4[g][p] = feat[4g..4g+3][p] * e^(Z[p]-zsub) as
//                   float4 (one 16-byte load later fetches 4 channels of a source
//                   pixel), S[j][p] = scalar planes (2-layer tail channels, e^Z).
//   euler_table     once per batch of frames: both Euler chains in registers,
//                   landing coordinates of every (frame, direction, pixel), and
//                   per-(frame, destination tile) entry counts.
//   bin_scan        exclusive scan of the counts -> bin offsets.
//   bin_fill        every (pixel, direction) is appended to the bins of the
//                   destination tiles its 2x2 footprint touches.
//   gather          one CTA per (destination tile 32x8, frame): expands its bin
//                   into per-destination-pixel (source, weight) lists in shared
//                   memory, each thread loads its list into registers and loops
//                   over the channel groups: LDG.128 + 4 FMA per (pair, group).
#include "slr_common.cuh"
#include "slr_host.h"
#include <algorithm>

namespace slr {

constexpr int TW = 32;                 // destination tile width  (one warp per tile row)
constexpr int TH = 8;                  // destination tile height
constexpr int TILE = TW * TH;          // threads per gather CTA, one destination pixel each
#ifndef SLR_GATHER_DEPTH
#define SLR_GATHER_DEPTH 16            // 12 or 16
#endif
#ifndef SLR_GATHER_EXPERIMENT
#define SLR_GATHER_EXPERIMENT 0        // 1 / 2: timing experiments (no stores / no feature loads), results invalid
#endif
#ifndef SLR_GATHER_MINBLOCKS
#define SLR_GATHER_MINBLOCKS 2         // resident CTAs per SM the register budget is sized for
#endif
constexpr int kDepth = SLR_GATHER_DEPTH;   // (source, weight) pairs a thread holds in registers
#ifndef SLR_GATHER_SMEM_DEPTH
#define SLR_GATHER_SMEM_DEPTH 12
#endif
#ifndef SLR_EXPAND_MINBLOCKS
#define SLR_EXPAND_MINBLOCKS 8
#endif
constexpr int kListDepth = 64;          // slots per pixel in the global row lists (>= kSmemDepth; deeper = heavy tile)
constexpr int kHeavyGroups = 4;         // channel groups per work item of heavy_tile_kernel
constexpr int kSmemDepth = SLR_GATHER_SMEM_DEPTH;   // pairs per destination pixel the shared list table holds
constexpr int kChunk = 8192;           // bin entries expanded per pass
constexpr int kGatherSmem = kSmemDepth * TILE * 8 + TILE * 4;
constexpr int kMaxFrames = 64;         // frames per launch (alpha table lives in the parameters)
constexpr unsigned kDirBit = 0x80000000u;
constexpr float kStaticLand = -1.0e30f;  // landing marker of pixels that are not binned

struct FrameAlphas { float a[kMaxFrames]; };

#ifndef SLR_GATHER_PATCH_LOG2H
#define SLR_GATHER_PATCH_LOG2H 0       // warp patch height = 1 << this (0: 32x1, 1: 16x2, 2: 8x4)
#endif
constexpr int kPH2 = SLR_GATHER_PATCH_LOG2H;      // log2 patch height
constexpr int kPW2 = 5 - kPH2;                    // log2 patch width
constexpr int kPatchesX = TW >> kPW2;             // patches per tile row
// thread index <-> pixel inside the 32x8 tile
__device__ __forceinline__ int tile_lx(int tid) { return (((tid >> 5) % kPatchesX) << kPW2) + (tid & ((1 << kPW2) - 1)); }
__device__ __forceinline__ int tile_ly(int tid) { return (((tid >> 5) / kPatchesX) << kPH2) + ((tid & 31) >> kPW2); }
__device__ __forceinline__ int tile_thread(int lx, int ly)
{
    const int warp = (ly >> kPH2) * kPatchesX + (lx >> kPW2);
    const int lane = ((ly & ((1 << kPH2) - 1)) << kPW2) + (lx & ((1 << kPW2) - 1));
    return warp * 32 + lane;
}

// Destination tiles touched by a footprint, in a fixed order shared by the count
// and the fill pass.  East / south columns only count when their weight is
// non-zero (landing exactly on a cell -- static pixels -- touches one cell).
__device__ __forceinline__ void touched_tiles(const Footprint& f, float ox, float oy, int H, int W,
                                              int tiles_x, int out[4])
{
    out[0] = out[1] = out[2] = out[3] = -1;
    if (f.ok == 0u) return;
    const bool c0 = f.x0 >= 0 && f.x0 < W;
    const bool c1 = f.x0 + 1 >= 0 && f.x0 + 1 < W && ox > (float)f.x0;
    const bool r0 = f.y0 >= 0 && f.y0 < H;
    const bool r1 = f.y0 + 1 >= 0 && f.y0 + 1 < H && oy > (float)f.y0;
    const int tc0 = c0 ? f.x0 / TW : -1;
    int tc1 = c1 ? (f.x0 + 1) / TW : -1;
    const int tr0 = r0 ? f.y0 / TH : -1;
    int tr1 = r1 ? (f.y0 + 1) / TH : -1;
    if (tc1 == tc0) tc1 = -1;
    if (tr1 == tr0) tr1 = -1;
    if (tr0 >= 0 && tc0 >= 0) out[0] = tr0 * tiles_x + tc0;
    if (tr0 >= 0 && tc1 >= 0) out[1] = tr0 * tiles_x + tc1;
    if (tr1 >= 0 && tc0 >= 0) out[2] = tr1 * tiles_x + tc0;
    if (tr1 >= 0 && tc1 >= 0) out[3] = tr1 * tiles_x + tc1;
}

// Warp-aggregated "reserve one slot in counters[key]" for the lanes whose key >= 0.
// All 32 lanes must call.  Returns the lane's slot (undefined for key < 0).
__device__ __forceinline__ unsigned warp_reserve(unsigned* counters, int key)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(peers) - 1;
    unsigned base = 0;
    if (key >= 0 && (int)lane == leader) base = atomicAdd(counters + key, (unsigned)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    return base + (unsigned)__popc(peers & ((1u << lane) - 1u));
}

// ---------------------------------------------------------------------------
// scene_prep: pre-weighted, channel-interleaved features
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
scene_prep_kernel(const float* __restrict__ feat, const float* __restrict__ z, const float* __restrict__ zsub,
                  const float* __restrict__ tail, int n_tail, float4* __restrict__ G4, float* __restrict__ S,
                  int C, int64_t P)
{
    // planes have stride P + 1: pixel P is the all-zero pixel unused gather slots read
    const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (p > P) return;
    if (p == P) {
        const int groups = (C + 3) >> 2;
        if (blockIdx.y == 0) {
            for (int g = 0; g < groups; ++g) G4[(int64_t)g * (P + 1) + P] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            for (int j = 0; j <= n_tail; ++j) S[(int64_t)j * (P + 1) + P] = 0.0f;
        }
        return;
    }
    const float ez = expf(z[p] - (zsub ? *zsub : 0.0f));
    const int groups = (C + 3) >> 2;
    const int g0 = blockIdx.y * ((groups + gridDim.y - 1) / gridDim.y);
    const int g1 = min(groups, g0 + (groups + (int)gridDim.y - 1) / (int)gridDim.y);
    for (int g = g0; g < g1; ++g) {
        float v[4];
        #pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = 4 * g + j;
            v[j] = c < C ? feat[(int64_t)c * P + p] * ez : 0.0f;
        }
        G4[(int64_t)g * (P + 1) + p] = make_float4(v[0], v[1], v[2], v[3]);
    }
    if (blockIdx.y == 0) {
        for (int j = 0; j < n_tail; ++j) S[(int64_t)j * (P + 1) + p] = tail[(int64_t)j * P + p];
        S[(int64_t)n_tail * (P + 1) + p] = ez;
    }
}

// ---------------------------------------------------------------------------
// euler_table: both chains for frames f = 0..n-1 of the batch.
//   forward  steps of frame f: steps_f0 + f      (t - start)
//   backward steps of frame f: steps_b0 - f      (end - t + 1)
// Same arithmetic as euler_kernel (bit-identical displacements); the landing
// coordinate stored is x + (dest - x), i.e. exactly what the reference splat
// computes from the displacement (softsplat.py:169-170).
// ---------------------------------------------------------------------------
struct EulerState { float dx, dy; bool invalid; };

__device__ __forceinline__ void euler_step(EulerState& s, const float* __restrict__ motion, float sign,
                                           float cx, float cy, float xmax, float ymax, int W, int64_t P)
{
    const int64_t at = (int64_t)rintf(s.dy) * W + (int64_t)rintf(s.dx);
    const float mx = __fmul_rn(sign, __ldg(motion + at));
    const float my = __fmul_rn(sign, __ldg(motion + P + at));
    s.dx = __fadd_rn(s.dx, mx);
    s.dy = __fadd_rn(s.dy, my);
    s.invalid = s.invalid || s.dx > xmax || s.dx < 0.0f || s.dy > ymax || s.dy < 0.0f;
    if (s.invalid) { s.dx = cx; s.dy = cy; }
}

__global__ void __launch_bounds__(256)
euler_table_kernel(const float* __restrict__ motion, int H, int W, int steps_f0, int steps_b0, int n,
                   float* __restrict__ land, unsigned* __restrict__ counts, int tiles_x, int n_tiles)
{
    const int64_t P = (int64_t)H * W;
    const int64_t p_raw = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const bool active = p_raw < P;
    const int64_t p = active ? p_raw : 0;
    const int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
    const float cx = (float)x, cy = (float)y;
    const float xmax = (float)(W - 1), ymax = (float)(H - 1);
    const float sentinel = (float)(max(H, W) + 1);

    // Static pixels (motion exactly 0 at the pixel: it never moves, in either direction) are
    // not binned at all: the gather adds their self-contribution implicitly.  Their landing
    // entry is a far-away marker, which the count and fill passes skip like any off-frame pixel.
    const bool is_static = __ldg(motion + p) == 0.0f && __ldg(motion + P + p) == 0.0f;
    if (__all_sync(0xffffffffu, is_static || !active)) {
        if (active)
            for (int i = 0; i < 4 * n; ++i) land[(int64_t)i * P + p] = kStaticLand;
        return;
    }

    auto emit = [&](const EulerState& s, int f, int dir) {
        const float ddx = s.invalid ? sentinel : __fsub_rn(s.dx, cx);
        const float ddy = s.invalid ? sentinel : __fsub_rn(s.dy, cy);
        const float ox = is_static ? kStaticLand : __fadd_rn(cx, ddx);
        const float oy = is_static ? kStaticLand : __fadd_rn(cy, ddy);
        int tiles[4];
        if (active) {
            float* l = land + ((int64_t)(f * 2 + dir) * 2) * P + p;
            l[0] = ox;
            l[P] = oy;
            const Footprint fp = footprint_at(ox, oy, H, W);
            touched_tiles(fp, ox, oy, H, W, tiles_x, tiles);
        } else {
            tiles[0] = tiles[1] = tiles[2] = tiles[3] = -1;
        }
        unsigned* cnt = counts + (int64_t)f * n_tiles;
        #pragma unroll
        for (int k = 0; k < 4; ++k)
            if (__any_sync(0xffffffffu, tiles[k] >= 0)) warp_reserve(cnt, tiles[k]);
    };

    // forward chain
    EulerState s = {cx, cy, false};
    if (steps_f0 == 0) emit(s, 0, 0);
    const int last_f = steps_f0 + n - 1;
    for (int k = 1; k <= last_f; ++k) {
        euler_step(s, motion, 1.0f, cx, cy, xmax, ymax, W, P);
        if (k >= steps_f0) emit(s, k - steps_f0, 0);
    }
    // backward chain (-motion); frame f needs steps_b0 - f steps
    s = {cx, cy, false};
    const int first_b = steps_b0 - (n - 1);          // >= 0, checked by the host
    if (first_b == 0) emit(s, n - 1, 1);
    for (int k = 1; k <= steps_b0; ++k) {
        euler_step(s, motion, -1.0f, cx, cy, xmax, ymax, W, P);
        if (k >= first_b) emit(s, steps_b0 - k, 1);
    }
}

// ---------------------------------------------------------------------------
// bin_scan: per frame, exclusive scan of tile counts -> offsets; counts are zeroed
// so that the fill pass can reuse them as cursors.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
bin_scan_kernel(unsigned* __restrict__ counts, unsigned* __restrict__ offsets, int n_tiles)
{
    __shared__ unsigned warp_sums[32];
    __shared__ unsigned carry_s;
    unsigned* cnt = counts + (int64_t)blockIdx.x * n_tiles;
    unsigned* off = offsets + (int64_t)blockIdx.x * (n_tiles + 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned v = i < n_tiles ? cnt[i] : 0u;
        unsigned incl = v;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned w = warp_sums[lane];
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_sums[lane] = w;        // inclusive over warps
        }
        __syncthreads();
        const unsigned carry = carry_s;
        const unsigned before = carry + (warp ? warp_sums[warp - 1] : 0u) + incl - v;
        if (i < n_tiles) { off[i] = before; cnt[i] = 0u; }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_sums[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) off[n_tiles] = carry_s;
}

// ---------------------------------------------------------------------------
// bin_fill: append (pixel | direction, landing x, landing y) to every touched tile.
// grid: (ceil(P/256), frames)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bin_fill_kernel(const float* __restrict__ land, const unsigned* __restrict__ offsets,
                unsigned* __restrict__ cursors, float4* __restrict__ ent,
                int H, int W, int tiles_x, int n_tiles, int64_t cap)
{
    const int64_t P = (int64_t)H * W;
    const int f = blockIdx.y;
    const int64_t p_raw = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const bool active = p_raw < P;
    const int64_t p = active ? p_raw : 0;
    const unsigned* off = offsets + (int64_t)f * (n_tiles + 1);
    unsigned* cur = cursors + (int64_t)f * n_tiles;
    float4* e = ent + (int64_t)f * cap;
    #pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
        const float* l = land + ((int64_t)(f * 2 + dir) * 2) * P + p;
        const float ox = __ldcs(l), oy = __ldcs(l + P);
        int tiles[4];
        if (active) {
            const Footprint fp = footprint_at(ox, oy, H, W);
            touched_tiles(fp, ox, oy, H, W, tiles_x, tiles);
        } else {
            tiles[0] = tiles[1] = tiles[2] = tiles[3] = -1;
        }
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!__any_sync(0xffffffffu, tiles[k] >= 0)) continue;
            const unsigned slot = warp_reserve(cur, tiles[k]);
            if (tiles[k] >= 0) {
                const int64_t at = (int64_t)off[tiles[k]] + slot;
                e[at] = make_float4(__uint_as_float((unsigned)p | (dir ? kDirBit : 0u)), ox, oy, 0.0f);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// gather
// ---------------------------------------------------------------------------
// Scene buffer layout: every plane carries one extra, all-zero pixel at index P
// (plane stride P + 1).  Unused list slots point at it, so the hot loop needs no
// predicates and no zero-initialisation: an unused slot loads zeros with weight 0.
struct GatherParams {
    const char* G;             // [groups] planes of (P + 1) float4
    const float* S;            // [n_tail + 1] planes of (P + 1) float   (last plane = e^Z)
    const float4* ent;         // [frames][cap]  (pixel | dir << 31, landing x, landing y, -)
    const float* motion;       // [2][P]: a destination pixel with zero motion contributes to itself
    const unsigned* offsets;   // [frames][n_tiles + 1]
    uint2* lists;              // [frames][n_rows][kListDepth][32]: per 32-pixel row, slot-major (source, weight)
    unsigned* row_k;           // [frames][n_rows]: slots in use in that row (warp-uniform list length)
    unsigned* tile_flag;       // [frames][n_tiles]: 1 = lists overflowed, tile is done by the multi-pass kernel
    unsigned* flag_list;       // [frames * n_tiles]: compacted (tile * n_frames + f) of the flagged tiles
    unsigned* flag_count;      // [1], zeroed by slr_clip_plan
    float* out;                // [frames][C][P]
    float* aux;                // [frames][n_tail + 1][P] raw sums (tail..., norm) or NULL
    float* mask;               // [frames][P] norm > eps, or NULL
    int C, groups, H, W, tiles_x, n_tiles, n_frames;
    int64_t P, cap;
    float eps;
    FrameAlphas alphas;
};

struct GatherCtx {
    const char* G;       // group plane 0
    const float* S;      // scalar plane 0
    const uint2* ell;    // this thread's column of the list table (slot stride ell_stride), for slots >= kDepth
    int ell_stride;
    float* out;          // this thread's pixel in plane 0 of the frame
    int64_t P;
    int groups, C, my_cnt, kmax;
    float eps;
    bool inframe, whole_bin, wrote;
};

// address of pixel `p` in a float4 plane: one IMAD.WIDE
__device__ __forceinline__ const float4* px16(const char* plane, unsigned p)
{
    unsigned long long a;
    asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(a) : "r"(p), "l"(plane));
    return reinterpret_cast<const float4*>(a);
}

// One pass of a destination pixel over its (source, weight) list.
//   K    compile-time number of register-resident list slots (the warp's longest list
//        rounded up, at most kDepth); longer lists continue from shared memory.
//   GI   channel groups per iteration.  A warp pays one full HBM latency per iteration
//        (some lane always misses L1), and the 16 resident warps per SM are too few to hide
//        it, so every iteration puts K * GI <= 16 independent LDG.128 in flight.
//   FAST the whole bin fits one pass and C % 4 == 0: plain normalised streaming stores.
template <int NT, int K, int GI, bool FAST>
__device__ __forceinline__ void gather_lists(const GatherCtx& c, const unsigned (&pk)[kDepth],
                                             const float (&wk)[kDepth], float& nrm, float (&tl)[NT > 0 ? NT : 1])
{
    const int64_t sstride = c.P + 1;
    // scalar planes: tail channels, then the e^Z weight (the normaliser)
    {
        float sv[NT + 1][K];
        #pragma unroll
        for (int k = 0; k < K; ++k) {
            #pragma unroll
            for (int t = 0; t <= NT; ++t) sv[t][k] = __ldg(c.S + (int64_t)t * sstride + pk[k]);
        }
        #pragma unroll
        for (int k = 0; k < K; ++k) {
            #pragma unroll
            for (int t = 0; t < NT; ++t) tl[t] = fmaf(sv[t][k], wk[k], tl[t]);
            nrm = fmaf(sv[NT][k], wk[k], nrm);
        }
    }
    if (K == kDepth) {
        for (int k = kDepth; k < c.kmax; ++k) {
            const uint2 e = c.ell[k * c.ell_stride];         // unused slots hold (zero pixel, 0)
            const unsigned p = e.x;
            const float w = __uint_as_float(e.y);
            #pragma unroll
            for (int t = 0; t < NT; ++t) tl[t] = fmaf(__ldg(c.S + (int64_t)t * sstride + p), w, tl[t]);
            nrm = fmaf(__ldg(c.S + (int64_t)NT * sstride + p), w, nrm);
        }
    }
    const float inv = c.whole_bin ? 1.0f / fmaxf(nrm, c.eps) : 1.0f;

    const char* Gg = c.G;
    const size_t gstride = (size_t)(c.P + 1) * 16;
    const size_t ostride = (size_t)c.P;
    float* o = c.out;
    for (int g = 0; g < c.groups; g += GI, Gg += GI * gstride, o += 4 * GI * ostride) {
        // K <= 8: all K * GI loads of the iteration are issued before the first FMA;
        // longer lists (GI == 1) go in two halves to stay inside the register budget
        constexpr int B = K <= 8 ? K : K / 2;
        float4 accs[GI];
        #pragma unroll
        for (int gi = 0; gi < GI; ++gi) accs[gi] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        #pragma unroll
        for (int kb = 0; kb < K; kb += B) {
            float4 v[GI][B];
            #pragma unroll
            for (int gi = 0; gi < GI; ++gi) {
                #pragma unroll
#if SLR_GATHER_EXPERIMENT == 2 || SLR_GATHER_EXPERIMENT == 3     // timing experiment only: no feature loads
                for (int k = 0; k < B; ++k) v[gi][k] = make_float4(__uint_as_float(pk[kb + k]), 1.0f, 2.0f, 3.0f);
#else
                for (int k = 0; k < B; ++k) v[gi][k] = __ldg(px16(Gg + gi * gstride, pk[kb + k]));
#endif
            }
            #pragma unroll
            for (int gi = 0; gi < GI; ++gi) {
                #pragma unroll
                for (int k = 0; k < B; ++k) {
                    accs[gi].x = fmaf(v[gi][k].x, wk[kb + k], accs[gi].x);
                    accs[gi].y = fmaf(v[gi][k].y, wk[kb + k], accs[gi].y);
                    accs[gi].z = fmaf(v[gi][k].z, wk[kb + k], accs[gi].z);
                    accs[gi].w = fmaf(v[gi][k].w, wk[kb + k], accs[gi].w);
                }
            }
        }
        #pragma unroll
        for (int gi = 0; gi < GI; ++gi) {
            float4 acc = accs[gi];
            if (K == kDepth) {
                for (int k = kDepth; k < c.kmax; ++k) {       // rare: lists longer than the register file holds
                    const uint2 e = c.ell[k * c.ell_stride];
                    const float4 t = __ldg(px16(Gg + gi * gstride, e.x));
                    const float w = __uint_as_float(e.y);
                    acc.x = fmaf(t.x, w, acc.x);
                    acc.y = fmaf(t.y, w, acc.y);
                    acc.z = fmaf(t.z, w, acc.z);
                    acc.w = fmaf(t.w, w, acc.w);
                }
            }
            if (c.inframe) {
                float* og = o + 4 * gi * ostride;
#if SLR_GATHER_EXPERIMENT == 1 || SLR_GATHER_EXPERIMENT == 3     // timing experiment only: no output stores
                if (acc.x + acc.y + acc.z + acc.w == 1.2345e30f) *og = inv;
#else
                if (FAST) {
                    __stcs(og, acc.x * inv);
                    __stcs(og + ostride, acc.y * inv);
                    __stcs(og + 2 * ostride, acc.z * inv);
                    __stcs(og + 3 * ostride, acc.w * inv);
                } else {
                    const float r[4] = {acc.x, acc.y, acc.z, acc.w};
                    #pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (4 * (g + gi) + j < c.C) {
                            float* oj = og + j * ostride;
                            if (c.whole_bin) __stcs(oj, r[j] * inv);
                            else *oj = c.wrote ? *oj + r[j] : r[j];
                        }
                    }
                }
#endif
            }
        }
    }
}

template <int NT, int K>
__device__ __forceinline__ void gather_dispatch(const GatherCtx& c, const unsigned (&pk)[kDepth],
                                                const float (&wk)[kDepth], float& nrm, float (&tl)[NT > 0 ? NT : 1])
{
    constexpr int GI = K == 1 ? 16 : K == 2 ? 8 : K <= 4 ? 4 : K <= 8 ? 2 : 1;
    if (c.whole_bin && (c.C & 3) == 0) {
        if (c.groups % GI == 0) gather_lists<NT, K, GI, true>(c, pk, wk, nrm, tl);
        else gather_lists<NT, K, 1, true>(c, pk, wk, nrm, tl);
    } else {
        gather_lists<NT, K, 1, false>(c, pk, wk, nrm, tl);
    }
}

// ---------------------------------------------------------------------------
// heavy_tile_kernel: destination tiles in which some pixel receives more pairs than the list
// table holds (convergence zones of the flow: the synthetic 60-step fields pile up to ~100
// sources on single pixels and 10x the average number of pairs on single tiles).  Work is
// assigned per PAIR instead of per destination pixel: every thread walks bin entries and adds
// into a per-tile accumulator in shared memory with shared-memory atomics, one channel group
// at a time, so the cost is linear in the number of pairs whatever their distribution.  A small
// fixed grid walks the compacted list of flagged tiles that expand_kernel produced.
// ---------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(TILE)
heavy_tile_kernel(const GatherParams prm)
{
    __shared__ float acc[kHeavyGroups * TILE * 4];      // [group in chunk][pixel][4 channels]
    __shared__ float sn[(NT + 1) * TILE];

    const int tid = threadIdx.x;
    const int64_t P = prm.P;
    const int64_t sstride = P + 1;
    const size_t gstride = (size_t)(P + 1) * 16;
    // work item = (flagged tile, chunk of kHeavyGroups channel groups): the few heavy tiles are
    // spread over many CTAs so that the slowest one does not serialise the launch
    const unsigned n_chunks = (unsigned)((prm.groups + kHeavyGroups - 1) / kHeavyGroups);
    const unsigned n_work = *prm.flag_count * n_chunks;
    for (unsigned wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
        const unsigned item = prm.flag_list[wi / n_chunks];
        const int g_lo = (int)(wi % n_chunks) * kHeavyGroups, g_hi = min(prm.groups, g_lo + kHeavyGroups);
        const int f = (int)(item % (unsigned)prm.n_frames), tile = (int)(item / (unsigned)prm.n_frames);
        const int tx = tile % prm.tiles_x, ty = tile / prm.tiles_x;
        const int X = tx * TW + tile_lx(tid), Y = ty * TH + tile_ly(tid);
        const bool inframe = X < prm.W && Y < prm.H;
        const int64_t pix = (int64_t)Y * prm.W + X;
        const float a_f = prm.alphas.a[f], a_b = 1.0f - a_f;
        const unsigned* off = prm.offsets + (int64_t)f * (prm.n_tiles + 1);
        const unsigned beg = off[tile], end = off[tile + 1];
        const float4* ent = prm.ent + (int64_t)f * prm.cap;
        const bool self_static = inframe && __ldg(prm.motion + pix) == 0.0f && __ldg(prm.motion + P + pix) == 0.0f;
        const float w_self = a_f + a_b;
        float* out = prm.out + (int64_t)f * prm.C * P + pix;

        // visits every (destination thread, source pixel, weight) pair of the bin that lands in this tile
        auto for_each_pair = [&](auto&& fn) {
            if (self_static) fn(tid, (unsigned)pix, w_self);
            for (unsigned e = beg + tid; e < end; e += TILE) {
                const float4 en = __ldg(ent + e);
                const unsigned pd = __float_as_uint(en.x);
                const Footprint fp = footprint_at(en.y, en.z, prm.H, prm.W);
                const float a = (pd >> 31) ? a_b : a_f;
                #pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int lx = fp.x0 + (k & 1) - tx * TW, ly = fp.y0 + (k >> 1) - ty * TH;
                    const float wa = fp.w[k] * a;
                    if ((fp.ok >> k & 1u) && lx >= 0 && lx < TW && ly >= 0 && ly < TH && wa != 0.0f)
                        fn(tile_thread(lx, ly), pd & ~kDirBit, wa);
                }
            }
        };

        #pragma unroll
        for (int t = 0; t <= NT; ++t) sn[t * TILE + tid] = 0.0f;
        #pragma unroll
        for (int j = 0; j < 4 * kHeavyGroups; ++j) acc[j * TILE + tid] = 0.0f;
        __syncthreads();
        for_each_pair([&](int d, unsigned p, float w) {
            #pragma unroll
            for (int t = 0; t <= NT; ++t) atomicAdd(&sn[t * TILE + d], __ldg(prm.S + (int64_t)t * sstride + p) * w);
        });
        // all channel groups of the chunk in one pass over the pairs: kHeavyGroups loads in flight
        const char* Gg = prm.G + (size_t)g_lo * gstride;
        for_each_pair([&](int d, unsigned p, float w) {
            float4 v[kHeavyGroups];
            #pragma unroll
            for (int gi = 0; gi < kHeavyGroups; ++gi)
                v[gi] = g_lo + gi < g_hi ? __ldg(px16(Gg + gi * gstride, p)) : make_float4(0.f, 0.f, 0.f, 0.f);
            #pragma unroll
            for (int gi = 0; gi < kHeavyGroups; ++gi) {
                float* a = acc + (gi * TILE + d) * 4;
                atomicAdd(a + 0, v[gi].x * w);
                atomicAdd(a + 1, v[gi].y * w);
                atomicAdd(a + 2, v[gi].z * w);
                atomicAdd(a + 3, v[gi].w * w);
            }
        });
        __syncthreads();
        const float nrm = sn[NT * TILE + tid];
        const float inv = 1.0f / fmaxf(nrm, prm.eps);
        if (inframe) {
            for (int g = g_lo; g < g_hi; ++g) {
                #pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = 4 * g + j;
                    if (c < prm.C) __stcs(out + (int64_t)c * P, acc[((g - g_lo) * TILE + tid) * 4 + j] * inv);
                }
            }
        }
        if (inframe && g_lo == 0) {
            if (prm.aux) {
                float* a = prm.aux + (int64_t)f * (NT + 1) * P + pix;
                #pragma unroll
                for (int j = 0; j <= NT; ++j) a[(int64_t)j * P] = sn[j * TILE + tid];
            }
            if (prm.mask) prm.mask[(int64_t)f * P + pix] = nrm > prm.eps ? 1.0f : 0.0f;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// expand_kernel: one CTA per (destination tile, frame).  Expands the tile's bin into
// per-destination-pixel (source, weight) lists in shared memory (canonical slots, see
// multipass_gather_kernel) and writes them out per 32-pixel row, slot-major, so that the
// gather kernel needs no shared memory and no barriers.  Small register footprint: several
// CTAs per SM hide the latency of this pointer-chasing part.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE, SLR_EXPAND_MINBLOCKS)
expand_kernel(const GatherParams prm)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint2* ell = reinterpret_cast<uint2*>(smem_raw);
    unsigned* cnt = reinterpret_cast<unsigned*>(ell + kSmemDepth * TILE);

    const int tid = threadIdx.x;
    const int f = blockIdx.x % prm.n_frames, tile = blockIdx.x / prm.n_frames;
    const int tx = tile % prm.tiles_x, ty = tile / prm.tiles_x;
    const int X = tx * TW + tile_lx(tid), Y = ty * TH + tile_ly(tid);
    const bool inframe = X < prm.W && Y < prm.H;
    const int64_t P = prm.P;
    const int64_t pix = (int64_t)Y * prm.W + X;
    const float a_f = prm.alphas.a[f], a_b = 1.0f - a_f;
    const unsigned* off = prm.offsets + (int64_t)f * (prm.n_tiles + 1);
    const unsigned beg = __ldg(off + tile), end = __ldg(off + tile + 1);
    const float4* ent = prm.ent + (int64_t)f * prm.cap;
    // a destination pixel with exactly zero motion receives itself with weight alpha + (1 - alpha)
    // (its forward and backward splat both land exactly on it); it was not binned
    const bool self_static = inframe && __ldg(prm.motion + pix) == 0.0f && __ldg(prm.motion + P + pix) == 0.0f;

    const int n_rows = prm.n_tiles * TH;
    uint2* lists_tile = prm.lists + ((int64_t)f * n_rows + (int64_t)tile * TH) * (kListDepth * 32);
    cnt[tid] = self_static ? 1u : 0u;
    if (self_static) ell[tid] = make_uint2((unsigned)pix, __float_as_uint(a_f + a_b));
    __syncthreads();
    for (unsigned e = beg + tid; e < end; e += TILE) {
        const float4 en = __ldcs(ent + e);
        const unsigned pd = __float_as_uint(en.x);
        const Footprint fp = footprint_at(en.y, en.z, prm.H, prm.W);
        const unsigned dir = pd >> 31;
        const float a = dir ? a_b : a_f;
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int lx = fp.x0 + (k & 1) - tx * TW, ly = fp.y0 + (k >> 1) - ty * TH;
            const float wa = fp.w[k] * a;
            if ((fp.ok >> k & 1u) && lx >= 0 && lx < TW && ly >= 0 && ly < TH && wa != 0.0f) {
                const int d = tile_thread(lx, ly);
                const unsigned pref = 2u * k + dir;
                const unsigned old = atomicOr(&cnt[d], 1u << pref);
                int slot = pref;
                if (old >> pref & 1u) slot = 8 + (int)(atomicAdd(&cnt[d], 256u) >> 8);
                const uint2 pair = make_uint2(pd & ~kDirBit, __float_as_uint(wa));
                if (slot < kSmemDepth) {
                    ell[slot * TILE + d] = pair;
                } else if (slot < kListDepth) {
                    // deeper than the shared table: straight to its place in the global row list
                    __stcg(lists_tile + ((int64_t)(d >> 5) * kListDepth + slot) * 32 + (d & 31), pair);
                }
            }
        }
    }
    __syncthreads();
    const int over = __syncthreads_or(8 + (int)(cnt[tid] >> 8) > kListDepth);
    if (tid == 0) {
        prm.tile_flag[(int64_t)f * prm.n_tiles + tile] = over ? 1u : 0u;
        if (over) prm.flag_list[atomicAdd(prm.flag_count, 1u)] = blockIdx.x;
    }
    if (over) return;

    const unsigned occ = cnt[tid] & 0xffu;
    const int n_ovf = (int)(cnt[tid] >> 8);
    const int my_cnt = n_ovf > 0 ? 8 + n_ovf : 32 - __clz(occ);
    uint2 e0 = ell[tid], e1 = ell[TILE + tid];
    // same source in slots 0 and 1 (a static pixel's forward and backward self-splat): one slot
    const bool merged = my_cnt == 2 && occ == 3u && e0.x == e1.x;
    if (merged) e0.y = __float_as_uint(__uint_as_float(e0.y) + __uint_as_float(e1.y));
    const int kmax = __reduce_max_sync(0xffffffffu, merged ? 1 : my_cnt);
    const int64_t row = (int64_t)f * n_rows + (int64_t)tile * TH + (tid >> 5);
    if ((tid & 31) == 0) prm.row_k[row] = (unsigned)kmax;
    uint2* dst = prm.lists + row * (kListDepth * 32) + (tid & 31);
    const uint2 none = make_uint2((unsigned)P, 0u);       // the zero pixel, weight 0
    for (int k = 0; k < min(kmax, kSmemDepth); ++k) {
        uint2 e = k == 0 ? e0 : ell[k * TILE + tid];
        bool used = k < 8 ? (occ >> k & 1u) : (k - 8 < n_ovf);
        if (merged && k == 1) used = false;
        __stcg(dst + k * 32, used ? e : none);
    }
    // slots past the shared table were written in place; pad this lane's unused ones
    for (int k = max(my_cnt, kSmemDepth); k < kmax; ++k) __stcg(dst + k * 32, none);
}

// ---------------------------------------------------------------------------
// rowgather_kernel: the hot kernel.  One warp per 32-pixel destination row (the 8 warps of a
// CTA are the 8 rows of one tile, so vertically shared sources stay in this SM's L1).  No
// shared memory, no barriers: every warp is an independent stream of
//   read its list -> [K*GI LDG.128 -> 4*K*GI FMA -> 4*GI STG] per iteration,
// so the loads of some warps overlap the FMAs and stores of others.
// CTA order is frame-fastest: the CTAs resident at any moment work on the SAME destination
// tiles of all frames of the batch; their source regions differ only by the frame-to-frame
// displacement, so a source line is fetched from HBM once per batch and the other frames hit
// it in L2 (one frame's features alone, 204 MB, exceed the 126 MB L2).
// ---------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(TILE, SLR_GATHER_MINBLOCKS)
rowgather_kernel(const GatherParams prm)
{
    const int tid = threadIdx.x;
    const int f = blockIdx.x % prm.n_frames, tile = blockIdx.x / prm.n_frames;
    if (__ldg(prm.tile_flag + (int64_t)f * prm.n_tiles + tile) != 0u) return;    // multi-pass kernel's tile
    const int tx = tile % prm.tiles_x, ty = tile / prm.tiles_x;
    const int X = tx * TW + (tid & 31), Y = ty * TH + (tid >> 5);
    const bool inframe = X < prm.W && Y < prm.H;
    const int64_t P = prm.P;
    const int64_t pix = (int64_t)Y * prm.W + X;
    const int n_rows = prm.n_tiles * TH;
    const int64_t row = (int64_t)f * n_rows + (int64_t)tile * TH + (tid >> 5);
    const int kmax = (int)__ldg(prm.row_k + row);
    const uint2* src = prm.lists + row * (kListDepth * 32) + (tid & 31);

    unsigned pk[kDepth];
    float wk[kDepth];
    #pragma unroll
    for (int k = 0; k < kDepth; ++k) {
        uint2 e = make_uint2((unsigned)P, 0u);
        if (k < kmax) e = __ldcg(src + k * 32);
        pk[k] = e.x;
        wk[k] = __uint_as_float(e.y);
    }
    float nrm = 0.0f;
    float tl[NT > 0 ? NT : 1] = {0.0f};
    GatherCtx ctx;
    ctx.G = prm.G; ctx.S = prm.S; ctx.ell = src; ctx.ell_stride = 32; ctx.P = P;
    ctx.groups = prm.groups; ctx.C = prm.C; ctx.eps = prm.eps;
    ctx.out = prm.out + (int64_t)f * prm.C * P + pix;
    ctx.inframe = inframe; ctx.whole_bin = true; ctx.wrote = false;
    ctx.my_cnt = kmax; ctx.kmax = kmax;
    switch ((kmax + 1) >> 1) {
        case 0: gather_dispatch<NT, 1>(ctx, pk, wk, nrm, tl); break;
        case 1: if (kmax == 1) gather_dispatch<NT, 1>(ctx, pk, wk, nrm, tl);
                else gather_dispatch<NT, 2>(ctx, pk, wk, nrm, tl);
                break;
        case 2: gather_dispatch<NT, 4>(ctx, pk, wk, nrm, tl); break;
        case 3: gather_dispatch<NT, 6>(ctx, pk, wk, nrm, tl); break;
        case 4: gather_dispatch<NT, 8>(ctx, pk, wk, nrm, tl); break;
        case 5: gather_dispatch<NT, 10>(ctx, pk, wk, nrm, tl); break;
#if SLR_GATHER_DEPTH == 16
        case 6: gather_dispatch<NT, 12>(ctx, pk, wk, nrm, tl); break;
        case 7: gather_dispatch<NT, 14>(ctx, pk, wk, nrm, tl); break;
        default: gather_dispatch<NT, 16>(ctx, pk, wk, nrm, tl); break;
#else
        default: gather_dispatch<NT, 12>(ctx, pk, wk, nrm, tl); break;
#endif
    }
    if (!inframe) return;
    if (prm.aux) {
        float* a = prm.aux + (int64_t)f * (NT + 1) * P + pix;
        #pragma unroll
        for (int j = 0; j < NT; ++j) a[(int64_t)j * P] = tl[j];
        a[(int64_t)NT * P] = nrm;
    }
    if (prm.mask) prm.mask[(int64_t)f * P + pix] = nrm > prm.eps ? 1.0f : 0.0f;
}

}  // namespace slr

// ===========================================================================
// C ABI
// ===========================================================================
using namespace slr;

namespace {

struct Workspace {
    float* land;          // [n][2 dirs][2][P]
    unsigned* counts;     // [n][n_tiles]      (counts, then cursors)
    unsigned* offsets;    // [n][n_tiles + 1]
    float4* ent;          // [n][cap]
    uint2* lists;         // [n][n_rows][kListDepth][32]
    unsigned* row_k;      // [n][n_rows]
    unsigned* tile_flag;  // [n][n_tiles]
    unsigned* flag_list;  // [n * n_tiles]
    unsigned* flag_count; // [1]
    size_t bytes;
};

inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

Workspace carve(void* base, int64_t H, int64_t W, int n)
{
    const int64_t P = H * W;
    const int64_t tiles = ((W + TW - 1) / TW) * ((H + TH - 1) / TH);
    const int64_t cap = 8 * P;
    char* p = (char*)base;
    size_t o = 0;
    Workspace w;
    w.land = (float*)(p + o);        o += align_up(sizeof(float) * 4 * P * n);
    w.counts = (unsigned*)(p + o);   o += align_up(sizeof(unsigned) * tiles * n);
    w.offsets = (unsigned*)(p + o);  o += align_up(sizeof(unsigned) * (tiles + 1) * n);
    w.ent = (float4*)(p + o);        o += align_up(sizeof(float4) * cap * n);
    w.lists = (uint2*)(p + o);       o += align_up(sizeof(uint2) * 32 * kListDepth * (size_t)(tiles * TH) * n);
    w.row_k = (unsigned*)(p + o);    o += align_up(sizeof(unsigned) * tiles * TH * n);
    w.tile_flag = (unsigned*)(p + o); o += align_up(sizeof(unsigned) * tiles * n);
    w.flag_list = (unsigned*)(p + o); o += align_up(sizeof(unsigned) * tiles * n);
    w.flag_count = (unsigned*)(p + o); o += align_up(sizeof(unsigned));
    w.bytes = o;
    return w;
}

}  // namespace

extern "C" size_t slr_clip_workspace_bytes(int64_t H, int64_t W, int n_frames)
{
    if (H <= 0 || W <= 0 || n_frames <= 0) return 0;
    return carve(nullptr, H, W, n_frames).bytes;
}

extern "C" size_t slr_scene_bytes(int64_t C, int n_tail, int64_t H, int64_t W)
{
    if (C <= 0 || n_tail < 0 || H <= 0 || W <= 0) return 0;
    return sizeof(float) * (size_t)(((C + 3) / 4) * 4 + n_tail + 1) * (size_t)(H * W + 1);
}

extern "C" int slr_scene_prep(const float* feat, const float* z, const float* zsub,
                              const float* tail, int n_tail, void* scene,
                              int64_t C, int64_t H, int64_t W, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(feat && z && scene && C > 0 && H > 0 && W > 0 && H * W < (1ll << 27) &&
                   n_tail >= 0 && n_tail <= 2 && (n_tail == 0 || tail) && ((uintptr_t)scene & 15) == 0,
                   "slr_scene_prep: bad arguments");
    const int64_t P = H * W;
    const int groups = (int)((C + 3) / 4);
    float4* G4 = (float4*)scene;
    float* S = (float*)scene + (int64_t)groups * 4 * (P + 1);
    dim3 grid((unsigned)((P + 1 + 255) / 256), (unsigned)std::min(groups, 4), 1);
    scene_prep_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(feat, z, zsub, tail, n_tail, G4, S, (int)C, P);
    return SLR_LAUNCH_STATUS();
}

extern "C" int slr_clip_plan(const float* motion, int64_t H, int64_t W, int start, int end, int t0,
                             int n_frames, void* workspace, size_t workspace_bytes, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(motion && workspace && H > 0 && W > 0 && H * W < (1ll << 27) &&
                   n_frames > 0 && n_frames <= kMaxFrames &&
                   t0 >= start && t0 + n_frames - 1 <= end + 1 && ((uintptr_t)workspace & 15) == 0,
                   "slr_clip_plan: bad arguments");
    const int64_t P = H * W;
    const int tiles_x = (int)((W + TW - 1) / TW), tiles_y = (int)((H + TH - 1) / TH);
    const int n_tiles = tiles_x * tiles_y;
    const Workspace ws = carve(workspace, H, W, n_frames);
    SLR_CHECK_ARGS(ws.bytes <= workspace_bytes, "slr_clip_plan: workspace too small (see slr_clip_workspace_bytes)");
    cudaStream_t s = (cudaStream_t)stream_;

    SLR_CUDA(cudaMemsetAsync(ws.counts, 0, sizeof(unsigned) * (size_t)n_tiles * n_frames, s));
    SLR_CUDA(cudaMemsetAsync(ws.flag_count, 0, sizeof(unsigned), s));
    const unsigned pblocks = (unsigned)((P + 255) / 256);
    euler_table_kernel<<<pblocks, 256, 0, s>>>(motion, (int)H, (int)W, t0 - start, end - t0 + 1, n_frames,
                                               ws.land, ws.counts, tiles_x, n_tiles);
    bin_scan_kernel<<<n_frames, 1024, 0, s>>>(ws.counts, ws.offsets, n_tiles);
    bin_fill_kernel<<<dim3(pblocks, n_frames), 256, 0, s>>>(ws.land, ws.offsets, ws.counts, ws.ent,
                                                            (int)H, (int)W, tiles_x, n_tiles, 8 * P);
    return SLR_LAUNCH_STATUS();
}

extern "C" int slr_clip_gather(const void* scene, const float* motion, int64_t C, int n_tail, int64_t H, int64_t W,
                               int start, int end, int t0, int n_frames, float alpha_lo, float alpha_hi,
                               float* out, float* aux, float* mask,
                               const void* workspace, size_t workspace_bytes, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(scene && motion && out && workspace && C > 0 && H > 0 && W > 0 && H * W < (1ll << 27) &&
                   n_tail >= 0 && n_tail <= 2 && n_frames > 0 && n_frames <= kMaxFrames &&
                   t0 >= start && t0 + n_frames - 1 <= end + 1 && ((uintptr_t)workspace & 15) == 0,
                   "slr_clip_gather: bad arguments");
    const int64_t P = H * W;
    const int tiles_x = (int)((W + TW - 1) / TW), tiles_y = (int)((H + TH - 1) / TH);
    const int n_tiles = tiles_x * tiles_y;
    const Workspace ws = carve(const_cast<void*>(workspace), H, W, n_frames);
    SLR_CHECK_ARGS(ws.bytes <= workspace_bytes, "slr_clip_gather: workspace too small (see slr_clip_workspace_bytes)");

    GatherParams prm;
    const int groups = (int)((C + 3) / 4);
    prm.G = (const char*)scene;
    prm.S = (const float*)scene + (int64_t)groups * 4 * (P + 1);
    prm.ent = ws.ent; prm.motion = motion; prm.offsets = ws.offsets;
    prm.lists = ws.lists; prm.row_k = ws.row_k; prm.tile_flag = ws.tile_flag;
    prm.flag_list = ws.flag_list; prm.flag_count = ws.flag_count;
    prm.out = out; prm.aux = aux; prm.mask = mask;
    prm.C = (int)C; prm.groups = groups; prm.H = (int)H; prm.W = (int)W;
    prm.tiles_x = tiles_x; prm.n_tiles = n_tiles; prm.P = P; prm.cap = 8 * P; prm.eps = 1e-8f;
    for (int f = 0; f < n_frames; ++f) {
        // alpha = 1 - (t - start) / (end - start + 1) in fp32 (animating_softmax_splating.py:860),
        // optionally clamped (2layers...py:952)
        float a = 1.0f - (float)(t0 + f - start) / (float)(end - start + 1);
        a = fminf(fmaxf(a, alpha_lo), alpha_hi);
        prm.alphas.a[f] = a;
    }
    prm.n_frames = n_frames;
    dim3 grid((unsigned)n_tiles * (unsigned)n_frames, 1, 1);
    cudaStream_t s = (cudaStream_t)stream_;
    static bool attr_set = false;      // opt in to > 48 KB dynamic shared memory once
    if (!attr_set) {
        SLR_CUDA(cudaFuncSetAttribute(expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGatherSmem));
        attr_set = true;
    }
    expand_kernel<<<grid, TILE, kGatherSmem, s>>>(prm);
    const unsigned mp_grid = std::min<unsigned>(grid.x, 8u * (unsigned)slr_host::sm_count());
    if (n_tail == 0) {
        rowgather_kernel<0><<<grid, TILE, 0, s>>>(prm);
        heavy_tile_kernel<0><<<mp_grid, TILE, 0, s>>>(prm);
    } else if (n_tail == 1) {
        rowgather_kernel<1><<<grid, TILE, 0, s>>>(prm);
        heavy_tile_kernel<1><<<mp_grid, TILE, 0, s>>>(prm);
    } else {
        rowgather_kernel<2><<<grid, TILE, 0, s>>>(prm);
        heavy_tile_kernel<2><<<mp_grid, TILE, 0, s>>>(prm);
    }
    return SLR_LAUNCH_STATUS();
}

extern "C" int slr_clip_frames(const void* scene, const float* motion, int64_t C, int n_tail,
                               int64_t H, int64_t W, int start, int end, int t0, int n_frames,
                               float alpha_lo, float alpha_hi,
                               float* out, float* aux, float* mask,
                               void* workspace, size_t workspace_bytes, slr_stream_t stream_)
{
    int rc = slr_clip_plan(motion, H, W, start, end, t0, n_frames, workspace, workspace_bytes, stream_);
    if (rc) return rc;
    return slr_clip_gather(scene, motion, C, n_tail, H, W, start, end, t0, n_frames, alpha_lo, alpha_hi,
                           out, aux, mask, workspace, workspace_bytes, stream_);
}
