// Clip pipeline, part 1: everything that depends on the scene or on the motion only.
//
//   scene_prep      once per scene: G8[g][p] = feat[8g..8g+7][p] * e^(Z[p]-zsub), 32 bytes
//                   (one 256-bit load later fetches 8 channels of a source pixel),
//                   S[j][p] = scalar planes (2-layer tail channels, e^Z).
//   euler_table     once per run of frames: both Euler chains in registers, landing
//                   coordinates of every (frame, direction, pixel); pixels with exactly zero motion never
//                   move: they carry a marker (clip_gather.cu adds their self-contribution).  Direct index
//                   (default): also the 256-pixel blocks in which something moves; bin pipeline: also the
//                   per-(frame, destination tile) entry counts.
//   static_lanes    direct index: the lanes' initial slot masks (static pixels reserve their own slot),
//   slot_fill       copied into every frame of a batch; bind_batch: what a batch workspace refers to.
//   bin_scan        bin pipeline: exclusive scan of the counts -> bin offsets.
//   bin_fill        bin pipeline: every moving (pixel, direction) is appended to the bins of the
//                   destination tiles its 2x2 footprint touches.
//   flow_table      a one-frame table from a given flow (slr_softsplat_sum_fwd_gather).
//
// Reference: the Euler + splat-input part of forward_flow,
// /root/reference/models/animating_softmax_splating.py:847-862,895 and
// /root/reference/models/projection/euler_integration_manipulator.py:18-56.
#include "clip_common.cuh"

namespace slr {

// ---------------------------------------------------------------------------
// scene_prep: pre-weighted, channel-interleaved features
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
scene_prep_kernel(const float* __restrict__ feat, const float* __restrict__ z, const float* __restrict__ zsub,
                  const float* __restrict__ tail, int n_tail, float8* __restrict__ G8, float* __restrict__ S,
                  int C, int64_t P)
{
    // planes have stride P + 1: pixel P is the all-zero pixel unused gather slots read
    const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (p > P) return;
    const int groups = (int)scene_groups8(C);
    const float4 zero4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (p == P) {
        if (blockIdx.y == 0) {
            for (int g = 0; g < groups; ++g) { float8 v; v.lo = zero4; v.hi = zero4; G8[(int64_t)g * (P + 1) + P] = v; }
            for (int j = 0; j <= n_tail; ++j) S[(int64_t)j * (P + 1) + P] = 0.0f;
        }
        return;
    }
    const float ez = z ? expf(z[p] - (zsub ? *zsub : 0.0f)) : 1.0f;      // no importance plane: plain features (weight 1)
    const int g0 = blockIdx.y * ((groups + gridDim.y - 1) / gridDim.y);
    const int g1 = min(groups, g0 + (groups + (int)gridDim.y - 1) / (int)gridDim.y);
    for (int g = g0; g < g1; ++g) {
        float v[8];
        #pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = 8 * g + j;
            v[j] = c < C ? feat[(int64_t)c * P + p] * ez : 0.0f;
        }
        float8 o;
        o.lo = make_float4(v[0], v[1], v[2], v[3]);
        o.hi = make_float4(v[4], v[5], v[6], v[7]);
        G8[(int64_t)g * (P + 1) + p] = o;
    }
    if (blockIdx.y == 0) {
        for (int j = 0; j < n_tail; ++j) S[(int64_t)j * (P + 1) + p] = tail[(int64_t)j * P + p];
        S[(int64_t)n_tail * (P + 1) + p] = ez;
    }
}

// ---------------------------------------------------------------------------
// scene_quilt: the Q region of the scene buffer from its G4 region (clip_common.cuh): chunks of 16
// channels, 128-byte blocks of two horizontally adjacent pixels, float4 units permuted by the column.  Pure data movement;
// its own kernel so that a rank that RECEIVED the G4 / S regions (sharding.SceneExchange broadcasts
// only those) can derive Q locally.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
scene_quilt_kernel(const char* __restrict__ G, float4* __restrict__ Q, int groups4, int H, int W, int64_t P)
{
    // one thread per (pixel column slot of a block row, chunk): x runs over 2 * Wb columns (the last may be padding)
    const int Wb = (int)quilt_row_blocks(W);
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t n = (int64_t)H * Wb * 2;
    const int q = blockIdx.y;
    float4* plane = Q + (int64_t)q * quilt_plane_blocks(H, W) * (kBlockBytes / 16);
    if (i >= n) {
        if (i < n + 8) plane[n * 4 + (i - n)] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);      // the trailing all-zero block
        return;
    }
    const int y = (int)(i / (2 * Wb)), x = (int)(i - (int64_t)y * 2 * Wb);
    const bool real = x < W;
    const int64_t p = real ? (int64_t)y * W + x : P;              // padding column of an odd-width image: the zero pixel
    float4* block = plane + ((int64_t)y * Wb + (x >> 1)) * (kBlockBytes / 16);
    const unsigned slot = quilt_slot((unsigned)x);
    #pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int g = 4 * q + u;
        block[slot ^ (unsigned)u] = g < groups4 ? ldg_group4(G, P, g, (unsigned)p) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
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
    // a NaN coordinate fails the test (see euler_kernel): no out-of-range index on the next step
    s.invalid = s.invalid || !(s.dx <= xmax && s.dx >= 0.0f && s.dy <= ymax && s.dy >= 0.0f);
    if (s.invalid) { s.dx = cx; s.dy = cy; }
}

// COUNT: also count the entries of every (frame, destination tile) bin (the bin pipeline); the direct
// index (insert_kernel) has no bins.
template <bool COUNT>
__global__ void __launch_bounds__(256)
euler_table_kernel(const float* __restrict__ motion, int H, int W, int steps_f0, int steps_b0, int n,
                   float* __restrict__ land, unsigned* __restrict__ counts, int tiles_x, int n_tiles, unsigned* __restrict__ moving)
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
    if (!COUNT) {        // direct index: insert_kernel only visits the blocks in which something moves
        const int any = __syncthreads_or(active && !is_static);
        if (any && threadIdx.x == 0) moving[1u + atomicAdd(moving, 1u)] = blockIdx.x;
    }
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
        int tiles[4] = {-1, -1, -1, -1};
        if (active) {
            float* l = land + ((int64_t)(f * 2 + dir) * 2) * P + p;
            l[0] = ox;
            l[P] = oy;
            if (COUNT) {
                const Footprint fp = footprint_at(ox, oy, H, W);
                touched_tiles(fp, ox, oy, H, W, tiles_x, tiles);
            }
        }
        if (COUNT) {
            unsigned* cnt = counts + (int64_t)f * n_tiles;
            #pragma unroll
            for (int k = 0; k < 4; ++k)
                if (__any_sync(0xffffffffu, tiles[k] >= 0)) warp_reserve(cnt, tiles[k]);
        }
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
// bin_fill: append (pixel | direction, landing x, landing y, pixel's row << 16 | column) to every touched tile.
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
    const unsigned xy = pack_xy((int)(p % W), (int)(p / W));      // the source's own coordinates, for the staging plan
    #pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
        const float* l = land + ((int64_t)(f * 2 + dir) * 2) * P + p;
        const float ox = __ldcs(l);
        // static pixels carry the marker in every frame and direction: a warp of them (about half of the
        // warps of a scene with a still background) has nothing to bin
        if (__all_sync(0xffffffffu, !active || ox == kStaticLand)) break;
        const float oy = __ldcs(l + P);
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
                e[at] = make_float4(__uint_as_float((unsigned)p | (dir ? kDirBit : 0u)), ox, oy, __uint_as_float(xy));
            }
        }
    }
}

// Slot masks of one frame's lanes before any source is inserted (a function of the motion field only: built once
// per clip table, copied into every frame of a batch by slot_fill_kernel).  A lane whose top (bottom) pixel has
// exactly zero motion receives that pixel itself with weight a + (1 - a) in canonical slot 0 (1): the slot is
// marked taken (a moving source that maps to it goes to the overflow slots, as it always did) and flagged as a
// self entry, which rowgather_kernel makes up instead of loading it.  One thread per PAIR of lanes (one word).
__global__ void __launch_bounds__(256)
static_lanes_kernel(const float* __restrict__ motion, unsigned* __restrict__ mask0, int H, int W, int tiles_x, int n_tiles)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;           // (row pair of the frame) * 16 + lane / 2
    if (i >= (int64_t)n_tiles * kPairsPerTile * 16) return;
    const int64_t P = (int64_t)H * W;
    const int pair = (int)(i >> 4), tile = pair / kPairsPerTile;
    const int Y = (tile / tiles_x) * TH + 2 * (pair % kPairsPerTile);
    unsigned word = 0u;
    #pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int X = (tile % tiles_x) * TW + 2 * (int)(i & 15) + h;
        unsigned m = 0u;
        if (X < W && Y < H) {
            const int64_t px = (int64_t)Y * W + X;
            if (__ldg(motion + px) == 0.0f && __ldg(motion + P + px) == 0.0f) m |= 1u | kSelfTop;
            if (Y + 1 < H && __ldg(motion + px + W) == 0.0f && __ldg(motion + P + px + W) == 0.0f) m |= 2u | kSelfBottom;
        }
        word |= m << (16 * h);
    }
    mask0[i] = word;
}

// masks <- the initial masks, overflow counts <- 0, for every frame of the batch
__global__ void __launch_bounds__(256)
slot_fill_kernel(const unsigned* __restrict__ mask0, unsigned* __restrict__ slot_mask, unsigned* __restrict__ slot_over,
                 int64_t words_per_frame, int n_frames)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= words_per_frame) return;
    const unsigned v = __ldg(mask0 + i);
    for (int f = 0; f < n_frames; ++f) {
        __stcg(slot_mask + (int64_t)f * words_per_frame + i, v);
        __stcg(reinterpret_cast<uint2*>(slot_over) + (int64_t)f * words_per_frame + i, make_uint2(0u, 0u));
    }
}

// A one-frame, one-direction "clip table" from a GIVEN flow (the operator-level summation splat through the
// gather, slr_softsplat_sum_fwd_gather): landing = pixel + flow exactly as the reference scatter computes it
// (softsplat.py:169-170); pixels with zero flow carry the static marker (they receive themselves through the
// lanes' initial masks); the second direction does not exist (every pixel static: its threads leave at once).
__global__ void __launch_bounds__(256)
flow_table_kernel(const float* __restrict__ flow, int H, int W, float* __restrict__ land, unsigned* __restrict__ moving)
{
    const int64_t P = (int64_t)H * W;
    const int64_t p_raw = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const bool active = p_raw < P;
    const int64_t p = active ? p_raw : 0;
    const float fx = __ldg(flow + p), fy = __ldg(flow + P + p);
    const bool is_static = fx == 0.0f && fy == 0.0f;
    const int any = __syncthreads_or(active && !is_static);
    if (any && threadIdx.x == 0) moving[1u + atomicAdd(moving, 1u)] = blockIdx.x;
    if (!active) return;
    const int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
    land[p] = is_static ? kStaticLand : __fadd_rn((float)x, fx);
    land[P + p] = is_static ? kStaticLand : __fadd_rn((float)y, fy);
    land[2 * P + p] = kStaticLand;
    land[3 * P + p] = kStaticLand;
}

}  // namespace slr

// ===========================================================================
// C ABI
// ===========================================================================
using namespace slr;
using slr_host::carve;
using slr_host::Workspace;

extern "C" size_t slr_clip_workspace_bytes(int64_t H, int64_t W, int n_frames)
{
    if (H <= 0 || W <= 0 || n_frames <= 0) return 0;
    return carve(nullptr, H, W, n_frames).bytes;
}

extern "C" size_t slr_scene_core_bytes(int64_t C, int n_tail, int64_t H, int64_t W)
{
    if (C <= 0 || n_tail < 0 || H <= 0 || W <= 0) return 0;
    return sizeof(float) * (size_t)scene_core_floats(C, n_tail, H * W);
}

extern "C" size_t slr_scene_bytes(int64_t C, int n_tail, int64_t H, int64_t W)
{
    if (C <= 0 || n_tail < 0 || H <= 0 || W <= 0) return 0;
    const int64_t P = H * W;
    return sizeof(float) * (size_t)(scene_quilt_offset_floats(C, n_tail, P) + scene_chunks(C) * quilt_plane_blocks(H, W) * (kBlockBytes / 4));
}

extern "C" int slr_scene_quilt(void* scene, int64_t C, int n_tail, int64_t H, int64_t W, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(scene && C > 0 && H > 0 && W > 0 && H * W < (1ll << 27) && n_tail >= 0 && n_tail <= 2 &&
                   ((uintptr_t)scene & 31) == 0, "slr_scene_quilt: bad arguments");
    const int64_t P = H * W;
    const int groups = (int)(scene_groups8(C) * 2);          // groups of four channels held by the G8 region
    float4* Q = (float4*)((float*)scene + scene_quilt_offset_floats(C, n_tail, P));
    const int64_t n = H * quilt_row_blocks(W) * 2 + 8;
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)scene_chunks(C), 1);
    scene_quilt_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>((const char*)scene, Q, groups, (int)H, (int)W, P);
    return SLR_LAUNCH_STATUS();
}

extern "C" int slr_scene_prep(const float* feat, const float* z, const float* zsub,
                              const float* tail, int n_tail, void* scene,
                              int64_t C, int64_t H, int64_t W, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(z, "slr_scene_prep: bad arguments");
    return slr_host::scene_prep(feat, z, zsub, tail, n_tail, scene, C, H, W, stream_);
}

// z == NULL: plain features (e^Z = 1)
int slr_host::scene_prep(const float* feat, const float* z, const float* zsub, const float* tail, int n_tail, void* scene,
                         int64_t C, int64_t H, int64_t W, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(feat && scene && C > 0 && H > 0 && W > 0 && H * W < (1ll << 27) &&
                   n_tail >= 0 && n_tail <= 2 && (n_tail == 0 || tail) && ((uintptr_t)scene & 31) == 0,
                   "slr_scene_prep: bad arguments");
    const int64_t P = H * W;
    const int groups = (int)scene_groups8(C);
    float8* G8 = (float8*)scene;
    float* S = (float*)scene + (int64_t)groups * kGroupChannels * (P + 1);
    dim3 grid((unsigned)((P + 1 + 255) / 256), (unsigned)std::min(groups, 4), 1);
    scene_prep_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(feat, z, zsub, tail, n_tail, G8, S, (int)C, P);
    const int rc = SLR_LAUNCH_STATUS();
    if (rc || !slr_host::gather_staged()) return rc;      // only the staged gather reads the quilted copy
    return slr_scene_quilt(scene, C, n_tail, H, W, stream_);
}

namespace {

// Euler chains, landing table, per-tile counts and bin offsets of frames t0 .. t0+n-1 into `tab`.
int build_table(const float* motion, int64_t H, int64_t W, int start, int end, int t0, int n_frames,
                const slr_host::ClipTable& tab, cudaStream_t s)
{
    const int64_t P = H * W;
    const int tiles_x = (int)((W + TW - 1) / TW), tiles_y = (int)((H + TH - 1) / TH);
    const int n_tiles = tiles_x * tiles_y;
    const unsigned pblocks = (unsigned)((P + 255) / 256);
    if (slr_host::index_direct()) {        // landing coordinates and the lanes' initial slot words: insert_kernel needs no bins
        SLR_CUDA(cudaMemsetAsync(tab.moving, 0, sizeof(unsigned), s));
        euler_table_kernel<false><<<pblocks, 256, 0, s>>>(motion, (int)H, (int)W, t0 - start, end - t0 + 1, n_frames,
                                                          tab.land, tab.counts, tiles_x, n_tiles, tab.moving);
        const int64_t words = (int64_t)n_tiles * kPairsPerTile * 16;
        static_lanes_kernel<<<(unsigned)((words + 255) / 256), 256, 0, s>>>(motion, tab.mask0, (int)H, (int)W, tiles_x, n_tiles);
        return SLR_LAUNCH_STATUS();
    }
    SLR_CUDA(cudaMemsetAsync(tab.counts, 0, sizeof(unsigned) * (size_t)n_tiles * n_frames, s));
    euler_table_kernel<true><<<pblocks, 256, 0, s>>>(motion, (int)H, (int)W, t0 - start, end - t0 + 1, n_frames,
                                                     tab.land, tab.counts, tiles_x, n_tiles, nullptr);
    bin_scan_kernel<<<n_frames, 1024, 0, s>>>(tab.counts, tab.offsets, n_tiles);
    return SLR_LAUNCH_STATUS();
}

// Direct index: a batch only REFERS to its landing coordinates (insert_kernel reads them where they are: in the
// clip table, which must stay valid until the batch's slr_clip_heavy has run); its per-lane slot words, tile
// flags and counters start at zero.
__global__ void bind_batch_kernel(BatchRefs* refs, const float* land, const unsigned* moving, unsigned* flag_count, unsigned* excess_count)
{
    if (threadIdx.x == 0) { refs->land = land; refs->moving = moving; *flag_count = 0u; *excess_count = 0u; }
}

int bind_batch(const float* land, const unsigned* mask0, const unsigned* moving, int64_t H, int64_t W, int n_frames, const Workspace& ws, cudaStream_t s)
{
    const int n_tiles = (int)(((W + TW - 1) / TW) * ((H + TH - 1) / TH));
    const int64_t words = (int64_t)n_tiles * kPairsPerTile * 16;
    slot_fill_kernel<<<(unsigned)((words + 255) / 256), 256, 0, s>>>(mask0, ws.slot_mask, ws.slot_over, words, n_frames);
    SLR_CUDA(cudaMemsetAsync(ws.tile_flag, 0, sizeof(unsigned) * (size_t)n_tiles * n_frames, s));
    bind_batch_kernel<<<1, 32, 0, s>>>(ws.refs, land, moving, ws.flag_count, ws.excess_count);
    return SLR_LAUNCH_STATUS();
}

// Bins of frames f0 .. f0+n-1 of a table into the batch workspace `ws` (whose offsets are already in place).
int fill_bins(const float* land, int64_t H, int64_t W, int n_frames, const Workspace& ws, cudaStream_t s)
{
    const int64_t P = H * W;
    const int tiles_x = (int)((W + TW - 1) / TW), tiles_y = (int)((H + TH - 1) / TH);
    const int n_tiles = tiles_x * tiles_y;
    SLR_CUDA(cudaMemsetAsync(ws.flag_count, 0, sizeof(unsigned), s));
    SLR_CUDA(cudaMemsetAsync(ws.excess_count, 0, sizeof(unsigned), s));
    const unsigned pblocks = (unsigned)((P + 255) / 256);
    bin_fill_kernel<<<dim3(pblocks, n_frames), 256, 0, s>>>(land, ws.offsets, ws.counts, ws.ent,
                                                            (int)H, (int)W, tiles_x, n_tiles, 8 * P);
    return SLR_LAUNCH_STATUS();
}

}  // namespace

// One-frame table of a given flow (direct index only): landing coordinates, initial slot masks, moving blocks.
int slr_host::flow_table(const float* flow, int64_t H, int64_t W, void* table, size_t table_bytes, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(flow && table && H > 0 && W > 0 && H * W < (1ll << 27) && ((uintptr_t)table & 15) == 0 && slr_host::index_direct(),
                   "flow_table: bad arguments");
    const slr_host::ClipTable tab = slr_host::carve_table(table, H, W, 1);
    SLR_CHECK_ARGS(tab.bytes <= table_bytes, "flow_table: table too small");
    cudaStream_t s = (cudaStream_t)stream_;
    const int64_t P = H * W;
    const int tiles_x = (int)((W + TW - 1) / TW), tiles_y = (int)((H + TH - 1) / TH);
    const int n_tiles = tiles_x * tiles_y;
    SLR_CUDA(cudaMemsetAsync(tab.moving, 0, sizeof(unsigned), s));
    flow_table_kernel<<<(unsigned)((P + 255) / 256), 256, 0, s>>>(flow, (int)H, (int)W, tab.land, tab.moving);
    const int64_t words = (int64_t)n_tiles * kPairsPerTile * 16;
    static_lanes_kernel<<<(unsigned)((words + 255) / 256), 256, 0, s>>>(flow, tab.mask0, (int)H, (int)W, tiles_x, n_tiles);
    return SLR_LAUNCH_STATUS();
}

extern "C" size_t slr_clip_table_bytes(int64_t H, int64_t W, int n_frames)
{
    if (H <= 0 || W <= 0 || n_frames <= 0 || n_frames > slr_host::kMaxTableFrames) return 0;
    return slr_host::carve_table(nullptr, H, W, n_frames).bytes;
}

extern "C" int slr_clip_table(const float* motion, int64_t H, int64_t W, int start, int end, int t0,
                              int n_frames, void* table, size_t table_bytes, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(motion && table && H > 0 && W > 0 && H * W < (1ll << 27) &&
                   n_frames > 0 && n_frames <= slr_host::kMaxTableFrames &&
                   t0 >= start && t0 + n_frames - 1 <= end + 1 && ((uintptr_t)table & 15) == 0,
                   "slr_clip_table: bad arguments");
    const slr_host::ClipTable tab = slr_host::carve_table(table, H, W, n_frames);
    SLR_CHECK_ARGS(tab.bytes <= table_bytes, "slr_clip_table: table too small (see slr_clip_table_bytes)");
    return build_table(motion, H, W, start, end, t0, n_frames, tab, (cudaStream_t)stream_);
}

extern "C" int slr_clip_bin(const void* table, size_t table_bytes, int64_t H, int64_t W, int table_frames,
                            int f0, int n_frames, void* workspace, size_t workspace_bytes, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(table && workspace && H > 0 && W > 0 && H * W < (1ll << 27) &&
                   table_frames > 0 && table_frames <= slr_host::kMaxTableFrames &&
                   n_frames > 0 && n_frames <= kMaxFrames && f0 >= 0 && f0 + n_frames <= table_frames &&
                   ((uintptr_t)table & 15) == 0 && ((uintptr_t)workspace & 15) == 0,
                   "slr_clip_bin: bad arguments");
    const int64_t P = H * W;
    const int n_tiles = (int)(((W + TW - 1) / TW) * ((H + TH - 1) / TH));
    const slr_host::ClipTable tab = slr_host::carve_table(const_cast<void*>(table), H, W, table_frames);
    SLR_CHECK_ARGS(tab.bytes <= table_bytes, "slr_clip_bin: table too small (see slr_clip_table_bytes)");
    const Workspace ws = carve(workspace, H, W, n_frames);
    SLR_CHECK_ARGS(ws.bytes <= workspace_bytes, "slr_clip_bin: workspace too small (see slr_clip_workspace_bytes)");
    cudaStream_t s = (cudaStream_t)stream_;
    if (slr_host::index_direct()) return bind_batch(tab.land + (size_t)f0 * 4 * P, tab.mask0, tab.moving, H, W, n_frames, ws, s);
    // the batch's own copy of its bin offsets (expand / gather / heavy read them from the workspace);
    // the fill cursors start at zero
    SLR_CUDA(cudaMemcpyAsync(ws.offsets, tab.offsets + (size_t)f0 * (n_tiles + 1),
                             sizeof(unsigned) * (size_t)(n_tiles + 1) * n_frames, cudaMemcpyDeviceToDevice, s));
    SLR_CUDA(cudaMemsetAsync(ws.counts, 0, sizeof(unsigned) * (size_t)n_tiles * n_frames, s));
    return fill_bins(tab.land + (size_t)f0 * 4 * P, H, W, n_frames, ws, s);
}

extern "C" int slr_clip_plan(const float* motion, int64_t H, int64_t W, int start, int end, int t0,
                             int n_frames, void* workspace, size_t workspace_bytes, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(motion && workspace && H > 0 && W > 0 && H * W < (1ll << 27) &&
                   n_frames > 0 && n_frames <= kMaxFrames &&
                   t0 >= start && t0 + n_frames - 1 <= end + 1 && ((uintptr_t)workspace & 15) == 0,
                   "slr_clip_plan: bad arguments");
    const Workspace ws = carve(workspace, H, W, n_frames);
    SLR_CHECK_ARGS(ws.bytes <= workspace_bytes, "slr_clip_plan: workspace too small (see slr_clip_workspace_bytes)");
    cudaStream_t s = (cudaStream_t)stream_;
    // the table of exactly this batch lives in the workspace itself
    slr_host::ClipTable tab;
    tab.land = ws.land; tab.counts = ws.counts; tab.offsets = ws.offsets; tab.mask0 = ws.mask0; tab.moving = ws.moving; tab.bytes = 0;
    const int rc = build_table(motion, H, W, start, end, t0, n_frames, tab, s);
    if (rc) return rc;
    if (slr_host::index_direct()) return bind_batch(ws.land, ws.mask0, ws.moving, H, W, n_frames, ws, s);
    return fill_bins(ws.land, H, W, n_frames, ws, s);       // bin_scan left the counts at zero: they are the cursors
}
