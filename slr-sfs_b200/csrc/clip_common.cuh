// Shared definitions of the clip pipeline (csrc/clip_plan.cu, csrc/clip_gather.cu).
#pragma once
#include "slr_common.cuh"
#include "slr_host.h"
#include <algorithm>
#include <stdlib.h>

namespace slr {

constexpr int TW = 32;                 // destination tile width  (one warp per tile row)
constexpr int TH = 8;                  // destination tile height
constexpr int TILE = TW * TH;          // destination pixels per tile
constexpr int kMaxFrames = 64;         // frames per launch (alpha table lives in the kernel parameters)
constexpr unsigned kDirBit = 0x80000000u;   // bin entry: source pixel | direction << 31
constexpr float kStaticLand = -1.0e30f;     // landing marker of pixels that are not binned

// Row-pair lists (the interface between expand_kernel and rowgather_kernel): a destination
// tile is 4 row pairs; a lane of a row pair owns the pixels (x, y) and (x, y + 1).  Per row
// pair the lists are slot-major: slot k holds, for each of the 32 lanes, one source pixel and
// its weights for the top and the bottom pixel.
constexpr int kPairsPerTile = TH / 2;
constexpr int kCanon = 12;             // canonical slots (direction x source-row offset x east/west)
#ifndef SLR_LIST_DEPTH
#define SLR_LIST_DEPTH 96
#endif
constexpr int kListDepth = SLR_LIST_DEPTH;   // slots per lane in the global lists; deeper = heavy tile

struct FrameAlphas { float a[kMaxFrames]; };

// Direct index: what a batch workspace refers to instead of copying it (both live in the clip table)
struct BatchRefs {
    const float* land;         // landing coordinates [frames][2 dirs][2][P] of the batch's first frame
    const unsigned* moving;    // [0] = number of 256-pixel blocks in which something moves, [1 ...] = those blocks
};

// Staging plan of one (destination tile, frame pair), written by expand_kernel and executed by
// stagegather_kernel: which rows of blocks of the Q region to copy where in shared memory.
constexpr int kStageFrames = 2;                      // frames that share one staged source region
constexpr int kPlanRows = 128;                       // source rows per set a plan can describe
constexpr int kPlanSets = 3;
constexpr int kMaxCopies = kPlanSets * kPlanRows;
constexpr int kCopyLenBits = 12;
struct StageRecord {
    unsigned n_copies;
    unsigned stages;                                 // 0: does not fit -> rowgather_kernel; 1 / 2: single / double buffered
    unsigned warp_bytes[8];                          // bytes per chunk of copies i with i % 8 == w (warp w of the gather issues those)
    unsigned copy_src[kMaxCopies];                   // first block (within a chunk plane of Q) of a row segment
    unsigned copy_dst[kMaxCopies];                   // staged block index << kCopyLenBits | blocks
};

// Row-pair list entry (uint4): x = source pixel | set << kSetShift, y / z = weights for the lane's top /
// bottom pixel, w = the source's (row << 16 | column).  The set says which staged region of the
// destination tile the source belongs to (stagegather_kernel stages the three separately: the
// forward and the backward sources of a tile lie a whole displacement apart):
constexpr int kSetShift = 28;                       // H * W < 2^27: the bits above are free
constexpr unsigned kPixelMask = (1u << kSetShift) - 1u;
// Direct index: the slot mask of a lane (16 bits; two lanes share a 32-bit word of Workspace::slot_mask, so the
// atomic claims of a warp touch 64 bytes: on B200 the claims are bound by the sectors the L2 atomic units see) --
// bits 0 .. kCanon-1: canonical slot in use; bits 14 / 15: the lane's top / bottom pixel has zero motion and
// receives itself (slot 0 / 1 is then reserved and holds no entry)
constexpr unsigned kSelfTop = 1u << 14, kSelfBottom = 1u << 15;
__host__ __device__ __forceinline__ unsigned lane_mask_shift(unsigned at) { return (at & 1u) << 4; }     // of lane `at` in word at >> 1
enum SourceSet { kSetForward = 0, kSetBackward = 1, kSetSelf = 2, kSets = 3 };   // self: static pixels receive themselves

__host__ __device__ __forceinline__ unsigned pack_xy(int x, int y) { return (unsigned)y << 16 | (unsigned)x; }

// Scene buffer (slr_scene_prep), three regions:
//   G8  [groups8][P + 1] x 32 B  channel groups of 8, pre-weighted by e^(Z - zsub); pixel P is all-zero.  One 256-bit
//                                load (LDG.E.256, sm_100) fetches 8 channels of a source pixel: for runs of 32 pixels
//                                starting anywhere, L1 delivers 95 B/clk/SM that way against 63 with 128-bit loads of a
//                                16-byte layout (profiles/microbench/ldg_width.cu)
//   S   [n_tail + 1][P + 1] float  scalar planes (2-layer tail channels, then e^(Z - zsub))
//   Q   [chunks][H * Wb + 1] x 128 B   the same features in chunks of 16 channels, pixel-PAIR-major: a block
//                                holds the 2 x 4 float4 units of the horizontally adjacent pixels (2b, y) and
//                                (2b + 1, y), b < Wb = ceil(W / 2); unit u of pixel x sits in slot quilt_slot(x) ^ u.
//                                Rows of blocks are what the TMA unit copies into shared memory for
//                                stagegather_kernel; the slot permutation makes a warp that reads one unit of 8
//                                pixels at column stride 1 OR 2 hit 8 different 16-byte bank groups.
constexpr int kChunkChannels = 16;
constexpr int kBlockBytes = 128;                     // per pixel pair and chunk
__host__ __device__ __forceinline__ unsigned quilt_slot(unsigned x) { return ((x & 1u) << 2) ^ ((x >> 1) & 7u); }
// byte offset of unit 0 of pixel column x inside a row of blocks that starts at block `row_block` (units u: ^ (u << 4))
__host__ __device__ __forceinline__ unsigned quilt_offset(unsigned row_block, unsigned x)
{
    return (row_block + (x >> 1)) * kBlockBytes + (quilt_slot(x) << 4);
}
__host__ __device__ __forceinline__ int64_t quilt_row_blocks(int64_t W) { return (W + 1) / 2; }
__host__ __device__ __forceinline__ int64_t quilt_plane_blocks(int64_t H, int64_t W) { return H * quilt_row_blocks(W) + 1; }
constexpr int kGroupChannels = 8;
constexpr int kGroupBytes = 32;                      // per pixel and group of G8
__host__ __device__ __forceinline__ int64_t scene_groups8(int64_t C) { return (C + kGroupChannels - 1) / kGroupChannels; }
__host__ __device__ __forceinline__ int64_t scene_core_floats(int64_t C, int n_tail, int64_t P)
{
    return (scene_groups8(C) * kGroupChannels + n_tail + 1) * (P + 1);
}
__host__ __device__ __forceinline__ int64_t scene_quilt_offset_floats(int64_t C, int n_tail, int64_t P)
{
    return (scene_core_floats(C, n_tail, P) + 31) / 32 * 32;          // 128-byte aligned
}
__host__ __device__ __forceinline__ int64_t scene_chunks(int64_t C) { return (C + kChunkChannels - 1) / kChunkChannels; }

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


// 8 channels of a source pixel
struct alignas(32) float8 { float4 lo, hi; };

// address of pixel `p` in a plane of G8: one IMAD.WIDE
__device__ __forceinline__ const float8* px32(const char* plane, unsigned p)
{
#if defined(SLR_CPU_EMULATION)      // tests/emu: the same sources compiled for the CPU
    return reinterpret_cast<const float8*>(plane + (size_t)p * kGroupBytes);
#else
    unsigned long long a;
    asm("mad.wide.u32 %0, %1, 32, %2;" : "=l"(a) : "r"(p), "l"(plane));
    return reinterpret_cast<const float8*>(a);
#endif
}

// one 256-bit read-only load (LDG.E.256.CONSTANT)
__device__ __forceinline__ float8 ldg256(const float8* p)
{
#if defined(SLR_CPU_EMULATION)
    emu::check(p);
    return *p;
#else
    float8 r;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w) : "l"(p));
    return r;
#endif
}

// the 4 channels 4 * g4 .. 4 * g4 + 3 of pixel p (heavy kernels, scene_quilt: per group of FOUR)
__device__ __forceinline__ float4 ldg_group4(const char* G, int64_t P, int g4, unsigned p)
{
    const char* a = G + ((size_t)(g4 >> 1) * (size_t)(P + 1) + p) * kGroupBytes + (g4 & 1) * 16;
    return __ldg(reinterpret_cast<const float4*>(a));
}

}  // namespace slr

// ---------------------------------------------------------------------------
// host side: layout of the caller-provided workspace
// ---------------------------------------------------------------------------
namespace slr_host {

struct Workspace {
    float* land;          // [n][2 dirs][2][P]   landing coordinates
    unsigned* counts;     // [n][n_tiles]        entries per destination tile (then fill cursors)
    unsigned* offsets;    // [n][n_tiles + 1]    bin offsets
    float4* ent;          // [n][cap]            bin entries (pixel | dir << 31, landing x, landing y, -)
    uint4* lists;         // [n][n_tiles * 4][kListDepth][32]  row-pair lists (source, w_top, w_bottom, -)
    unsigned* row_k;      // [n][n_tiles * 4]    slots in use per row pair
    unsigned* tile_flag;  // [n][n_tiles]        1 = heavy tile
    unsigned* fallback;   // [n][n_tiles]        1 = the tile's sources do not fit the staging area: rowgather_kernel does it
    slr::StageRecord* records;   // [ceil(n / 2)][n_tiles]   staging plans (expand_kernel -> stagegather_kernel)
    unsigned* flag_list;  // [n * n_tiles]       compacted heavy tiles
    unsigned* flag_count; // [1]
    uint4* excess;        // [excess_cap]        pairs beyond kListDepth (dest pixel, source, weight, frame)
    unsigned* excess_count; // [1]
    unsigned excess_cap;
    float* heavy_sums;    // [n][3][P]           (tail..., norm) sums of flagged tiles
    unsigned* slot_mask;  // [n][n_tiles * 4][16] direct index: per lane 16 bits (canonical slots in use, self flags)
    unsigned* slot_over;  // [n][n_tiles * 4][32] direct index: per lane, overflow slots claimed
    unsigned* mask0;      // [n_tiles * 4][16]   direct index, slr_clip_plan only: the initial slot masks (else in the clip table)
    unsigned* moving;     // [1 + ceil(P / 256)] direct index, slr_clip_plan only: the moving blocks (else in the clip table)
    slr::BatchRefs* refs; // [1]                 direct index: where the batch's landing coordinates / moving blocks are
    size_t bytes;
};

inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

// Clip table: everything that depends on the motion only, for a run of consecutive frames
// (slr_clip_table).  The batch workspace below embeds one for its own frames (slr_clip_plan).
constexpr int kMaxTableFrames = 4096;
struct ClipTable {
    float* land;          // [n][2 dirs][2][P]   landing coordinates
    unsigned* counts;     // [n][n_tiles]        entries per destination tile (zero again after the scan)
    unsigned* offsets;    // [n][n_tiles + 1]    bin offsets
    unsigned* mask0;      // [n_tiles * 4][16]   direct index: the lanes' slot masks before any source is inserted
    unsigned* moving;     // [1 + ceil(P / 256)] direct index: count, then the 256-pixel blocks in which something moves
    size_t bytes;
};

inline ClipTable carve_table(void* base, int64_t H, int64_t W, int n)
{
    using namespace slr;
    const int64_t P = H * W;
    const int64_t tiles = ((W + TW - 1) / TW) * ((H + TH - 1) / TH);
    char* p = (char*)base;
    size_t o = 0;
    ClipTable t;
    t.land = (float*)(p + o);        o += align_up(sizeof(float) * 4 * P * n);
    t.counts = (unsigned*)(p + o);   o += align_up(sizeof(unsigned) * tiles * n);
    t.offsets = (unsigned*)(p + o);  o += align_up(sizeof(unsigned) * (tiles + 1) * n);
    t.mask0 = (unsigned*)(p + o);    o += align_up(sizeof(unsigned) * 16 * (size_t)(tiles * kPairsPerTile));
    t.moving = (unsigned*)(p + o);   o += align_up(sizeof(unsigned) * (size_t)(1 + (P + 255) / 256));
    t.bytes = o;
    return t;
}

inline Workspace carve(void* base, int64_t H, int64_t W, int n)
{
    using namespace slr;
    const int64_t P = H * W;
    const int64_t tiles = ((W + TW - 1) / TW) * ((H + TH - 1) / TH);
    const bool direct = index_direct();      // no bins, no staging plans
    const int64_t cap = direct ? 0 : 8 * P;  // every (pixel, direction) touches at most 4 tiles
    char* p = (char*)base;
    size_t o = 0;
    Workspace w;
    w.land = (float*)(p + o);        o += align_up(sizeof(float) * 4 * P * n);
    w.counts = (unsigned*)(p + o);   o += align_up(sizeof(unsigned) * tiles * n);
    w.offsets = (unsigned*)(p + o);  o += align_up(sizeof(unsigned) * (tiles + 1) * n);
    w.ent = (float4*)(p + o);        o += align_up(sizeof(float4) * cap * n);
    w.lists = (uint4*)(p + o);       o += align_up(sizeof(uint4) * 32 * kListDepth * (size_t)(tiles * kPairsPerTile) * n);
    w.row_k = (unsigned*)(p + o);    o += align_up(sizeof(unsigned) * tiles * kPairsPerTile * n);
    w.tile_flag = (unsigned*)(p + o); o += align_up(sizeof(unsigned) * tiles * n);
    w.fallback = (unsigned*)(p + o); o += align_up(sizeof(unsigned) * tiles * n);
    w.records = (slr::StageRecord*)(p + o); o += direct ? 0 : align_up(sizeof(slr::StageRecord) * tiles * ((n + kStageFrames - 1) / kStageFrames));
    w.flag_list = (unsigned*)(p + o); o += align_up(sizeof(unsigned) * tiles * n);
    w.flag_count = (unsigned*)(p + o); o += align_up(sizeof(unsigned));
    w.excess_cap = (unsigned)std::min<int64_t>(2 * P * n, 1ll << 30);
    w.excess = (uint4*)(p + o);      o += align_up(sizeof(uint4) * (size_t)w.excess_cap);
    w.excess_count = (unsigned*)(p + o); o += align_up(sizeof(unsigned));
    w.heavy_sums = (float*)(p + o);  o += align_up(sizeof(float) * 3 * P * n);
    w.slot_mask = (unsigned*)(p + o); o += direct ? align_up(sizeof(unsigned) * 16 * (size_t)(tiles * kPairsPerTile) * n) : 0;
    w.slot_over = (unsigned*)(p + o); o += direct ? align_up(sizeof(unsigned) * 32 * (size_t)(tiles * kPairsPerTile) * n) : 0;
    w.mask0 = (unsigned*)(p + o);    o += direct ? align_up(sizeof(unsigned) * 16 * (size_t)(tiles * kPairsPerTile)) : 0;
    w.moving = (unsigned*)(p + o);   o += direct ? align_up(sizeof(unsigned) * (size_t)(1 + (P + 255) / 256)) : 0;
    w.refs = (slr::BatchRefs*)(p + o); o += align_up(sizeof(slr::BatchRefs));
    w.bytes = o;
    return w;
}

}  // namespace slr_host
