// Clip pipeline, part 2 ("algorithm G"): the joint block of the reference models
// (/root/reference/models/animating_softmax_splating.py:862-924: two summation splats of
// [fs*e^Z*a, e^Z*a], their sum, clamp and divide) as bin-by-destination + GATHER.
//
// Why not scatter: a frame at 768x1024x65 is 2 directions x 4 corners x 51 M elements =
// 409 M fp32 atomics; the measured L2 reduction rate on B200 (profiles/r01: 0.65 ms per
// frame) is 10x the HBM time of the same bytes.  The splat weights do not depend on the
// channel, so the scatter is a sparse (destination x source) operator applied to all
// channels: its structure is built once per frame from the displacement alone, and then
// every destination pixel PULLS its contributions, accumulates them in registers and
// writes the normalised result once.  No feature atomics, no accumulator round trip
// through HBM, no separate normalise pass.
//
//   insert_kernel     (default: the "direct index") one thread per (moving source pixel, direction, frame): the
//                     source's 2-4 list cells (source, w_top, w_bottom) written straight into the lists of the
//                     destination lanes; a slot is claimed by one atomicOr on the lane's 16-bit slot mask.
//   expand_kernel     (SLR_GATHER_MODE=bins / staged) one CTA per (destination tile 32x8, frame): turns the tile's
//                     bin (bin_fill_kernel, clip_plan.cu) into the same lists through a shared-memory table; staged:
//                     also the staging plan of stagegather_kernel.
//   rowgather_kernel  the hot kernel.  One warp per row pair (2 x 32 destination pixels), no
//                     shared memory, no barriers.  A lane owns the pixels (x, y) and (x, y+1):
//                     a source that feeds both (its north corners land on y, its south
//                     corners on y+1) is loaded ONCE -- 12 loads instead of 16 per pixel pair
//                     in regular flow.  Per group of 8 channels: up to 12 independent 256-bit loads in flight,
//                     then the FMAs, then streaming stores.  A CTA is 4 such warps: 2 row pairs of
//                     one tile in 2 consecutive frames (their sources differ by one frame's
//                     displacement, so they share lines in L1).
//   heavy_*_kernel    tiles with a lane deeper than the lists hold (convergence zones of the
//                     flow): per-pair fp32 reductions at L2, cost linear in pairs.
#include "clip_common.cuh"
#include "tma.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace slr {

#ifndef SLR_GATHER_MINBLOCKS
#define SLR_GATHER_MINBLOCKS 4         // resident 128-thread CTAs per SM the register budget is sized for
#endif
#ifndef SLR_GATHER_FRAMES
#define SLR_GATHER_FRAMES 2            // default CTA shape of rowgather_kernel: frames x row pairs
#define SLR_GATHER_PAIRS 2             // (measured: 2x2 = 4x1 > 1x4 > 2x4 > 4x4, profiles/r01/sweep_variants.jsonl; 1x4 kept for A/B)
#endif
#ifndef SLR_EXPAND_MINBLOCKS
#define SLR_EXPAND_MINBLOCKS 6
#endif
constexpr int kRegSlots = 16;          // list slots a lane keeps in registers; deeper ones are re-read per group
constexpr int kSmemSlots = 16;         // list slots expand_kernel stages in shared memory (>= kCanon)
static_assert(kSmemSlots >= kCanon && kSmemSlots <= kRegSlots, "shared list table: kCanon <= slots <= kRegSlots");
constexpr int kCols = TW * kPairsPerTile;   // 128 lanes (columns of row pairs) per tile
#ifndef SLR_HEAVY_GROUPS
#define SLR_HEAVY_GROUPS 4
#endif
constexpr int kHeavyGroups = SLR_HEAVY_GROUPS;   // channel groups per work item of heavy_scatter_kernel
constexpr unsigned kEmpty = 0xffffffffu;
#ifndef SLR_STAGE_BYTES
#define SLR_STAGE_BYTES (104 * 1024)
#endif
constexpr int kStageBytes = SLR_STAGE_BYTES;                   // staging area per CTA of stagegather_kernel (two CTAs per SM)
constexpr int kStageBlocks1 = kStageBytes / kBlockBytes - 1;           // one stage  (block 0 = the all-zero block)
constexpr int kStageBlocks2 = kStageBytes / 2 / kBlockBytes - 1;       // two stages (double buffered)

struct GatherParams {
    const char* G;             // [groups] planes of (P + 1) x 32 bytes (8 channels per pixel); pixel P of every plane is all-zero
    const float* S;            // [n_tail + 1] planes of (P + 1) float (last plane = e^Z)
    const float4* ent;         // [frames][cap] bin entries
    const float* motion;       // [2][P]: a destination pixel with zero motion contributes to itself
    const unsigned* offsets;   // [frames][n_tiles + 1]
    uint4* lists;              // [frames][n_tiles * 4][kListDepth][32]
    unsigned* row_k;           // [frames][n_tiles * 4]    slots in use per row pair
    unsigned* tile_flag;       // [frames][n_tiles]: 0 normal, 1 = some lane's list was cut at kListDepth (the
                               // rest is in `excess`), 2 = excess list full: the whole tile goes the heavy way
    unsigned* flag_list;       // [frames * n_tiles]: compacted (tile * n_frames + f) of the flagged tiles
    unsigned* flag_count;      // [1], zeroed by slr_clip_plan
    uint4* excess;             // [excess_cap]: (destination pixel, source pixel, weight, frame) beyond kListDepth
    unsigned* excess_count;    // [1], zeroed by slr_clip_plan
    unsigned excess_cap;
    float* heavy_sums;         // [frames][3][P] un-normalised (tail..., norm) sums of every destination pixel, written by
                               // expand_kernel; flagged tiles get their excess pairs added by heavy_excess_kernel
    const char* Q;             // [chunks] planes of (P + 1) x 64 B: the staged copy of G (clip_common.cuh)
    unsigned* fallback;        // [frames][n_tiles]: 1 = stagegather_kernel left the tile to rowgather_kernel
    int only_fallback;         // rowgather_kernel: skip the tiles stagegather_kernel has done
    int n_tail;                // scalar planes of S besides the normaliser
    int staged;                // expand_kernel: plan the staging (stagegather_kernel follows) or not (rowgather_kernel for all)
    StageRecord* records;      // [frame pairs][n_tiles] staging plans
    unsigned* slot_mask;       // direct index: [frames][n_tiles * 4][16] per lane 16 bits (bit k: canonical slot k in use; self flags)
    unsigned* slot_over;       // direct index: [frames][n_tiles * 4][32] per lane, overflow slots claimed
    const BatchRefs* refs;     // direct index: where the batch's landing coordinates and moving-block list are (in the clip table)
    int direct;                // lists built by insert_kernel (no bins, no row_k: the gather reads the slot masks)
    int plain_sum;             // write the un-normalised sums (the operator-level summation splat, slr_softsplat_sum_fwd_gather)
    float* out;                // [frames][C][P]
    float* aux;                // [frames][n_tail + 1][P] raw sums (tail..., norm) or NULL
    float* mask;               // [frames][P] norm > eps, or NULL
    float* nnz;                // [frames][P] number of non-zero output channels of the pixel, or NULL: all the decoder's
                               // per-element hole mask (x != 0, networks/architectures.py:369) contributes to its first
                               // partial convolution's mask update (layers/partialconv2d.py:61)
    int C, groups, H, W, tiles_x, n_tiles, n_frames;
    int64_t P, cap;
    float eps;
    FrameAlphas alphas;
};

// Canonical slot of a pair: direction, row offset of the source's north-west cell relative
// to the lane's TOP pixel (-1, 0, +1) and whether the pair is a west (dx = 0) or east corner.
//   slots 0..7  : dx*4 + dir*2 + cy      cy = 0: feeds top (north corner) AND bottom (south corner)
//                                         cy = 1: feeds the bottom pixel only (north corner)
//   slots 8..11 : 8 + dir*2 + dx         cy = -1: feeds the top pixel only (south corner)
// Slot 0 / 1 are where a static pixel's own contribution goes, so static rows need 2 slots.
__device__ __forceinline__ int canon_slot(unsigned dir, int cy, int dx)
{
    return cy < 0 ? 8 + 2 * (int)dir + dx : (dx << 2) | ((int)dir << 1) | cy;
}
enum SlotRole { kBoth, kTopOnly, kBottomOnly };
__host__ __device__ constexpr SlotRole slot_role(int k)
{
    return k < 8 ? ((k & 1) ? kBottomOnly : kBoth) : (k < kCanon ? kTopOnly : kBoth);
}


// ---------------------------------------------------------------------------
// expand_kernel
// One CTA per (destination tile, pair of consecutive frames).
//
// Phase A/B (staged gather only): the STAGING PLAN of the tile for the frame pair.  Every bin entry of the
// tile in either frame names a source pixel; per source set (forward / backward / static self) and source
// row the kernel takes the column range of the sources (one pair of warp reductions per 32 consecutive
// entries, which mostly share their row, then shared-memory min / max), lays the row segments out in the
// staging area and writes the list of copies to the tile's StageRecord.  Both frames share the region:
// their sources differ by one frame's displacement.  If the region does not fit, the tile's frames are
// flagged for rowgather_kernel and keep global pixel indices in their lists.
//
// Phase C, per frame: bin -> per-lane (source, w_top, w_bottom) lists of the tile's 4 row pairs.  A
// canonical slot of a lane is claimed (atomicCAS on its source field) by the first source that asks for
// it; a second source with the same (direction, row offset, east/west) -- the flow compresses there --
// goes to the lane's overflow slots.  For a staged tile the source field is the byte offset of the
// source in the staging area (final: stagegather_kernel uses it as it is).
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned plan_key(unsigned set, unsigned xy)      // (set, row, column), ordered like the plan's rows
{
    return set << 30 | (xy >> 16 & 0x3fffu) << 16 | (xy & 0xffffu);
}

// STAGED: with the staging plan (stagegather_kernel follows); otherwise lists only, one frame per CTA (rowgather_kernel follows).
template <bool STAGED>
__global__ void __launch_bounds__(TILE, SLR_EXPAND_MINBLOCKS)
expand_kernel(const GatherParams prm)
{
    __shared__ uint4 tab[kSmemSlots * kCols];      // tab[slot * kCols + col] = (source, w_top, w_bottom, row << 16 | column)
    __shared__ unsigned occ[kCols];                // used canonical slots (bit mask)
    __shared__ unsigned ovf[kCols];                // overflow slots in use (kCanon, kCanon + 1, ...)
    __shared__ unsigned excess_full;               // the global excess list ran out of room
    __shared__ int row_block[STAGED ? kSets : 1][STAGED ? kPlanRows : 1];    // staged block of column 0 of a source row (hashed by row % kPlanRows)
    __shared__ int ylo[kSets], yhi[kSets];
    __shared__ unsigned stages_s;
    __shared__ unsigned warp_bytes[8];             // bytes per chunk of the copies warp w of stagegather_kernel will issue
    static_assert(sizeof(tab) >= 2 * kSets * kPlanRows * sizeof(int), "the plan's column ranges alias the list table");
    int* xlo = reinterpret_cast<int*>(tab);        // [kSets][kPlanRows], phases A and B only
    int* xhi = xlo + kSets * kPlanRows;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // staged: a CTA plans and expands the kStageFrames frames that share a staged region; otherwise one frame per CTA
    constexpr int per_cta = STAGED ? kStageFrames : 1;
    const int n_fg = (prm.n_frames + per_cta - 1) / per_cta;
    const int fg = (int)(blockIdx.x % (unsigned)n_fg), tile = (int)(blockIdx.x / (unsigned)n_fg);
    const int tx = tile % prm.tiles_x, ty = tile / prm.tiles_x;
    const int64_t P = prm.P;
    const int X = tx * TW + lane, Y = ty * TH + warp;         // this thread's destination pixel
    const bool inside = X < prm.W && Y < prm.H;
    const int64_t pix = inside ? (int64_t)Y * prm.W + X : 0;
    const bool still = inside && __ldg(prm.motion + pix) == 0.0f && __ldg(prm.motion + P + pix) == 0.0f;

    unsigned stages = 0u;
    if (STAGED) {
        for (int i = tid; i < kSets * kPlanRows; i += TILE) { xlo[i] = 0x7fffffff; xhi[i] = -1; }
        if (tid < kSets) { ylo[tid] = 0x7fffffff; yhi[tid] = -1; }
        if (tid < 8) warp_bytes[tid] = 0u;
        __syncthreads();
        auto note = [&](unsigned set, int y, int x0, int x1) {
            atomicMin(&ylo[set], y); atomicMax(&yhi[set], y);
            atomicMin(&xlo[set * kPlanRows + (y & (kPlanRows - 1))], x0);
            atomicMax(&xhi[set * kPlanRows + (y & (kPlanRows - 1))], x1);
        };
        for (int fi = 0; fi < kStageFrames; ++fi) {
            const int f = fg * kStageFrames + fi;
            if (f >= prm.n_frames) break;
            const unsigned* off = prm.offsets + (int64_t)f * (prm.n_tiles + 1);
            const unsigned beg = __ldg(off + tile), end = __ldg(off + tile + 1);
            const float4* ent = prm.ent + (int64_t)f * prm.cap;
            for (unsigned e0 = beg + 32u * warp; e0 < end; e0 += TILE) {       // warp-uniform trip count
                const unsigned e = e0 + lane;
                const bool valid = e < end;
                unsigned key = 0u;
                if (valid) {
                    const float4 en = __ldg(ent + e);
                    key = plan_key(__float_as_uint(en.x) >> 31, __float_as_uint(en.w));
                }
                const unsigned k0 = __reduce_min_sync(0xffffffffu, valid ? key : 0xffffffffu);
                const unsigned k1 = __reduce_max_sync(0xffffffffu, valid ? key : 0u);
                if ((k0 >> 16) == (k1 >> 16)) {          // one (set, row) for the whole warp: the usual case
                    if (lane == 0) note(k0 >> 30, (int)(k0 >> 16 & 0x3fffu), (int)(k0 & 0xffffu), (int)(k1 & 0xffffu));
                } else if (valid) {
                    note(key >> 30, (int)(key >> 16 & 0x3fffu), (int)(key & 0xffffu), (int)(key & 0xffffu));
                }
            }
        }
        {   // static pixels receive themselves: the tile's own rows (a warp is one tile row)
            const unsigned m = __ballot_sync(0xffffffffu, still);
            if (m != 0u && lane == 0) note((unsigned)kSetSelf, Y, tx * TW + __ffs((int)m) - 1, tx * TW + 31 - __clz((int)m));
        }
        __syncthreads();
        if (warp == 0) {          // lay the row segments out: whole blocks (pixel pairs), block 0 is the all-zero block
            StageRecord* rec = prm.records + (int64_t)fg * prm.n_tiles + tile;
            const int Wb = (int)quilt_row_blocks(prm.W);
            int total = 1, n_copies = 0;
            bool ok = true;
            #pragma unroll
            for (int set = 0; set < kSets; ++set) {
                const int y0 = ylo[set], y1 = yhi[set];
                if (y1 < y0) continue;
                if (y1 - y0 >= kPlanRows) { ok = false; continue; }
                for (int r0 = 0; r0 < kPlanRows; r0 += 32) {
                    const int r = r0 + lane;
                    const int lo = xlo[set * kPlanRows + r], hi = xhi[set * kPlanRows + r];
                    const int len = hi >= lo ? (hi >> 1) - (lo >> 1) + 1 : 0;       // blocks
                    int incl = len;
                    #pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += t;
                    }
                    const int at = total + incl - len;
                    const unsigned m = __ballot_sync(0xffffffffu, len > 0);
                    if (len >= (1 << kCopyLenBits)) ok = false;
                    if (len > 0 && at + len <= kStageBlocks1 + 1) {
                        const int y = y0 + ((r - y0) & (kPlanRows - 1));       // the row of [y0, y1] that hashes to r
                        const int i = n_copies + __popc(m & ((1u << lane) - 1u));
                        row_block[set][r] = at - (lo >> 1);
                        rec->copy_src[i] = (unsigned)(y * Wb + (lo >> 1));
                        rec->copy_dst[i] = (unsigned)at << kCopyLenBits | (unsigned)len;
                        atomicAdd(&warp_bytes[i & 7], (unsigned)len * kBlockBytes);
                    }
                    total += __shfl_sync(0xffffffffu, incl, 31);
                    n_copies += __popc(m);
                }
            }
            ok = __all_sync(0xffffffffu, ok) && total <= kStageBlocks1 + 1;
            if (lane == 0) {
                stages_s = !ok ? 0u : (total <= kStageBlocks2 + 1 ? 2u : 1u);
                rec->n_copies = (unsigned)n_copies;
                rec->stages = stages_s;
            }
            __syncwarp();
            if (lane < 8) rec->warp_bytes[lane] = warp_bytes[lane];
        }
        __syncthreads();
        stages = stages_s;
    }
    // the source field of a list entry: staged byte offset, or global pixel | set
    auto source_field = [&](unsigned set, unsigned p, unsigned xy) -> unsigned {
        if (!STAGED || stages == 0u) return p | set << kSetShift;
        return quilt_offset((unsigned)row_block[set][xy >> 16 & (kPlanRows - 1)], xy & 0xffffu);
    };
    // unused slots: the all-zero pixel, weights 0 (its (row, column) = (H, 0) is pixel P of the scalar planes)
    const uint4 none = make_uint4(stages == 0u ? (unsigned)P : 0u, 0u, 0u, pack_xy(0, prm.H));

    for (int fi = 0; fi < per_cta; ++fi) {
        const int f = fg * per_cta + fi;
        if (f >= prm.n_frames) break;
        const float a_f = prm.alphas.a[f], a_b = 1.0f - a_f;
        const unsigned* off = prm.offsets + (int64_t)f * (prm.n_tiles + 1);
        const unsigned beg = __ldg(off + tile), end = __ldg(off + tile + 1);
        const float4* ent = prm.ent + (int64_t)f * prm.cap;
        const int64_t pair0 = (int64_t)f * prm.n_tiles * kPairsPerTile + (int64_t)tile * kPairsPerTile;
        uint4* lists_tile = prm.lists + pair0 * (kListDepth * 32);

        __syncthreads();            // the table is free: phase B, or the previous frame's write-out, is over
        for (int i = tid; i < kCanon * kCols; i += TILE) tab[i] = make_uint4(kEmpty, 0u, 0u, 0u);
        if (tid < kCols) { occ[tid] = 0u; ovf[tid] = 0u; }
        if (tid == 0) excess_full = 0u;
        __syncthreads();

        // a pair whose canonical slot belongs to another source: the lane's next overflow slot
        // (`src`: source field of the entry; `p`: the source's pixel, for the excess list)
        auto spill = [&](int lx, int ly, unsigned src, unsigned p, float w, unsigned xy) {
            const int col = (ly >> 1) * TW + lx, r = ly & 1;
            const int so = kCanon + (int)atomicAdd(&ovf[col], 1u);
            const uint4 e = make_uint4(src, r ? 0u : __float_as_uint(w), r ? __float_as_uint(w) : 0u, xy);
            if (so < kSmemSlots) {
                tab[so * kCols + col] = e;
            } else if (so < kListDepth) {     // deeper than the shared table: straight to its place in the global list
                __stcg(lists_tile + ((int64_t)(ly >> 1) * kListDepth + so) * 32 + lx, e);
            } else {
                // deeper than the lists (a convergence point): this one pair is added by an fp32
                // reduction at L2 after the gather (heavy_excess_kernel)
                const unsigned i = atomicAdd(prm.excess_count, 1u);
                const unsigned dpix = (unsigned)((ty * TH + ly) * prm.W + tx * TW + lx);
                if (i < prm.excess_cap) __stcg(prm.excess + i, make_uint4(dpix, p, __float_as_uint(w), (unsigned)f));
                else excess_full = 1u;
            }
        };
        // one (destination pixel, source, weight) pair -> its lane's list
        auto insert = [&](int lx, int ly, unsigned src, unsigned p, float w, unsigned dir, int dx, int dy, unsigned xy) {
            uint4* cell = tab + canon_slot(dir, (ly & 1) - dy, dx) * kCols + (ly >> 1) * TW + lx;
            const unsigned old = atomicCAS(&cell->x, kEmpty, src);
            if (old == kEmpty) {
                atomicOr(&occ[(ly >> 1) * TW + lx], 1u << canon_slot(dir, (ly & 1) - dy, dx));
                cell->w = xy;
            }
            // the slot is this source's: the other row's corner of the same source shares it
            if (old == kEmpty || old == src) ((ly & 1) ? cell->z : cell->y) = __float_as_uint(w);
            else spill(lx, ly, src, p, w, xy);
        };

        // a destination pixel with exactly zero motion receives itself with weight a + (1 - a)
        // (its forward and backward splat both land exactly on it); it was not binned
        if (still) insert(lane, warp, source_field((unsigned)kSetSelf, (unsigned)pix, pack_xy(X, Y)), (unsigned)pix, a_f + a_b, 0u, 0, 0, pack_xy(X, Y));
        for (unsigned e = beg + tid; e < end; e += TILE) {
            const float4 en = __ldcs(ent + e);
            const unsigned pd = __float_as_uint(en.x), xy = __float_as_uint(en.w);
            const Footprint fp = footprint_at(en.y, en.z, prm.H, prm.W);
            const unsigned dir = pd >> 31;
            const float a = dir ? a_b : a_f;
            const unsigned src = source_field(dir, pd & ~kDirBit, xy);
            #pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int lx = fp.x0 + (k & 1) - tx * TW, ly = fp.y0 + (k >> 1) - ty * TH;
                const float wa = fp.w[k] * a;
                if ((fp.ok >> k & 1u) && lx >= 0 && lx < TW && ly >= 0 && ly < TH && wa != 0.0f)
                    insert(lx, ly, src, pd & ~kDirBit, wa, dir, k & 1, k >> 1, xy);
            }
        }
        __syncthreads();
        const bool deep = tid < kCols && kCanon + (int)ovf[tid] > kListDepth;
        const int any_deep = __syncthreads_or(deep);
        const unsigned flag = excess_full ? 2u : (any_deep ? 1u : 0u);
        if (tid == 0) {
            prm.tile_flag[(int64_t)f * prm.n_tiles + tile] = flag;
            prm.fallback[(int64_t)f * prm.n_tiles + tile] = STAGED && stages == 0u ? 1u : 0u;
            if (flag) prm.flag_list[atomicAdd(prm.flag_count, 1u)] = (unsigned)(tile * prm.n_frames + f);
        }
        if (flag != 2u && tid < kCols) {
            // write the lists out, slot-major per row pair
            const unsigned my_occ = occ[tid];
            const int n_ovf = min((int)ovf[tid], kListDepth - kCanon);     // the rest is in the excess list
            const int my_hi = n_ovf > 0 ? kCanon + n_ovf : 32 - __clz(my_occ);    // slots [0, my_hi) may be used
            const int kmax = __reduce_max_sync(0xffffffffu, my_hi);
            const int64_t pair = pair0 + (tid >> 5);
            if ((tid & 31) == 0) prm.row_k[pair] = (unsigned)kmax;
            uint4* dst = prm.lists + pair * (kListDepth * 32) + (tid & 31);
            // ... and, on the way, the un-normalised sums of the scalar planes (tail channels, then the
            // normaliser e^Z) of this lane's two pixels: same products in the same order as rowgather_kernel
            float st[3] = {0.0f, 0.0f, 0.0f}, sb[3] = {0.0f, 0.0f, 0.0f};
            const int64_t sstride = P + 1;
            const bool want_sums = STAGED && stages != 0u;        // stagegather_kernel reads them; rowgather_kernel sums for itself
            auto scalars = [&](const uint4& e, int k) {
                if (!want_sums) return;
                const unsigned p = (e.w >> 16) * (unsigned)prm.W + (e.w & 0xffffu);
                #pragma unroll
                for (int j = 0; j < 3; ++j) {
                    if (j <= prm.n_tail) {
                        const float sv = __ldg(prm.S + (int64_t)j * sstride + p);
                        if (slot_role(k) != kBottomOnly) st[j] = fmaf(sv, __uint_as_float(e.y), st[j]);
                        if (slot_role(k) != kTopOnly) sb[j] = fmaf(sv, __uint_as_float(e.z), sb[j]);
                    }
                }
            };
            for (int k = 0; k < min(kmax, kSmemSlots); ++k) {
                const bool used = k < kCanon ? (my_occ >> k & 1u) : (k - kCanon < n_ovf);
                const uint4 e = used ? tab[k * kCols + tid] : none;
                __stcg(dst + k * 32, e);
                if (used) scalars(e, k);
            }
            // slots past the shared table were written in place (by whichever thread met the pair)
            if (want_sums)
                for (int k = kSmemSlots; k < my_hi; ++k) scalars(__ldcg(dst + k * 32), k);
            // pad this lane's unused ones
            for (int k = max(my_hi, kSmemSlots); k < kmax; ++k) __stcg(dst + k * 32, none);
            const int Xc = tx * TW + (tid & 31), Yc = ty * TH + 2 * (tid >> 5);
            if (want_sums && Xc < prm.W && Yc < prm.H) {
                float* hs = prm.heavy_sums + (int64_t)f * 3 * P + (int64_t)Yc * prm.W + Xc;
                #pragma unroll
                for (int j = 0; j < 3; ++j) {
                    if (j <= prm.n_tail) {
                        hs[(int64_t)j * P] = st[j];
                        if (Yc + 1 < prm.H) hs[(int64_t)j * P + prm.W] = sb[j];
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// insert_kernel: the DIRECT index (default).  No bins, no per-tile pass: one thread per (source pixel, frame)
// turns the pixel's two landing positions into list cells and writes them where the gather reads them.
//
// A source with north-west cell (x0, y0) feeds, per column cx = x0 + dx, the pixels (cx, y0) and (cx, y0 + 1).
// y0 even: both belong to the lane cx of ONE row pair -> one cell (source, w_north, w_south) in the canonical
// slot (direction, dx, "both").  y0 odd: the bottom pixel of one row pair and the top pixel of the next -> two
// cells with one weight each (slots "bottom only" / "top only").  So a thread always writes WHOLE cells, and a
// slot is claimed with one atomicOr on the lane's slot word: nobody ever reads a cell another thread is still
// writing.  A second source that maps to the same canonical slot (the flow compresses there) takes the lane's
// next overflow slot (atomicAdd); beyond kListDepth its pairs go to the excess list and the tile is flagged,
// exactly like expand_kernel does it.  Cells of the 32 consecutive sources a warp holds are consecutive
// 16-byte entries of one slot row in regular flow: 512-byte stores, one atomic request per 128-byte line.
// ---------------------------------------------------------------------------
#ifndef SLR_INSERT_STORE
#define SLR_INSERT_STORE __stcs        // measured against __stcg: profiles/r02 (direct4_*)
#endif
constexpr int kCellsPerSource = 4;      // per direction: 2 columns x (north-or-both cell, south cell)

// The list cells one source pixel writes for one direction of one frame.  Cell i is unused when slot[i] < 0.
struct SourceCells {
    int slot[kCellsPerSource];          // canonical slot
    unsigned at[kCellsPerSource];       // destination lane: (row pair within the batch) * 32 + lane   (< 2^32: n * P / 2)
    float wt[kCellsPerSource];          // weight for the lane's top pixel
    float wb[kCellsPerSource];          // ... bottom pixel
};

__device__ __forceinline__ unsigned lane_index(const GatherParams& prm, int f, int cx, int ytop)
{
    const int tile = (ytop / TH) * prm.tiles_x + cx / TW;
    return (unsigned)((((int64_t)f * prm.n_tiles + tile) * kPairsPerTile + (ytop % TH) / 2) * 32 + cx % TW);
}

// top pixel (x, y) of a destination lane
__device__ __forceinline__ void lane_pixel(const GatherParams& prm, unsigned at, int& cx, int& ytop)
{
    const unsigned pair = (at >> 5) % (unsigned)(prm.n_tiles * kPairsPerTile);
    const int tile = (int)(pair / kPairsPerTile);
    cx = (tile % prm.tiles_x) * TW + (int)(at & 31u);
    ytop = (tile / prm.tiles_x) * TH + 2 * (int)(pair % kPairsPerTile);
}

// Cells of a source that lands on (ox, oy) in direction `dir` of frame f.
__device__ __forceinline__ void source_cells(const GatherParams& prm, float ox, float oy, int f, int dir, SourceCells& c)
{
    const float a_f = prm.alphas.a[f];
    const float a = dir ? 1.0f - a_f : a_f;
    const Footprint fp = footprint_at(ox, oy, prm.H, prm.W);
    const bool even = (fp.y0 & 1) == 0;
    #pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
        const int i = dx * 2;
        const int cx = fp.x0 + dx;
        const float wn = (fp.ok >> dx & 1u) ? fp.w[dx] * a : 0.0f;               // north corner (cx, y0)
        const float ws = (fp.ok >> (2 + dx) & 1u) ? fp.w[2 + dx] * a : 0.0f;     // south corner (cx, y0 + 1)
        c.slot[i] = c.slot[i + 1] = -1;
        c.at[i] = c.at[i + 1] = 0u;
        c.wt[i] = c.wb[i] = c.wt[i + 1] = c.wb[i + 1] = 0.0f;
        if (even) {          // both corners belong to one lane: one cell
            if (wn != 0.0f || ws != 0.0f) {
                c.slot[i] = canon_slot((unsigned)dir, 0, dx);
                c.at[i] = lane_index(prm, f, cx, fp.y0);
                c.wt[i] = wn; c.wb[i] = ws;
            }
        } else {             // bottom pixel of one row pair, top pixel of the next
            if (wn != 0.0f) {
                c.slot[i] = canon_slot((unsigned)dir, 1, dx);
                c.at[i] = lane_index(prm, f, cx, fp.y0 - 1);
                c.wb[i] = wn;
            }
            if (ws != 0.0f) {
                c.slot[i + 1] = canon_slot((unsigned)dir, -1, dx);
                c.at[i + 1] = lane_index(prm, f, cx, fp.y0 + 1);
                c.wt[i + 1] = ws;
            }
        }
    }
}

// A cell beyond the list depth (a convergence point): its pairs go to the excess list and are added by fp32
// reductions at L2 after the gather (heavy_excess_kernel), which leaves the flagged tile un-normalised for them.
__device__ __forceinline__ void excess_cell(const GatherParams& prm, int f, unsigned at, unsigned src, float wt, float wb)
{
    int cx, ytop;
    lane_pixel(prm, at, cx, ytop);
    const int tile = (ytop / TH) * prm.tiles_x + cx / TW;
    if (atomicExch(prm.tile_flag + (int64_t)f * prm.n_tiles + tile, 1u) == 0u)
        prm.flag_list[atomicAdd(prm.flag_count, 1u)] = (unsigned)(tile * prm.n_frames + f);
    const unsigned n = (wt != 0.0f ? 1u : 0u) + (wb != 0.0f ? 1u : 0u);
    unsigned i = atomicAdd(prm.excess_count, n);       // a count beyond excess_cap = "overflowed": see overflow_*_kernel
    const unsigned dpix = (unsigned)(ytop * prm.W + cx);
    if (wt != 0.0f) { if (i < prm.excess_cap) __stcg(prm.excess + i, make_uint4(dpix, src, __float_as_uint(wt), (unsigned)f)); ++i; }
    if (wb != 0.0f) { if (i < prm.excess_cap) __stcg(prm.excess + i, make_uint4(dpix + (unsigned)prm.W, src, __float_as_uint(wb), (unsigned)f)); }
}

// grid: (ceil(P / 256), 2 * frames): blockIdx.y = frame * 2 + direction.  Pixels with exactly zero motion carry
// the marker in every frame and direction and leave after one load: what they receive from themselves is not a
// list entry at all (rowgather_kernel makes it up from the lane's slot word, see static_lanes_kernel).
__global__ void __launch_bounds__(256)
insert_kernel(const GatherParams prm)
{
    // only the blocks of 256 pixels in which something moves (euler_table_kernel lists them): the others leave at once
    const unsigned* moving = prm.refs->moving;
    if (blockIdx.x >= moving[0]) return;
    const int f = (int)(blockIdx.y >> 1), dir = (int)(blockIdx.y & 1u);
    const int64_t p = (int64_t)moving[1 + blockIdx.x] * 256 + threadIdx.x;
    if (p >= prm.P) return;
    const float* land = prm.refs->land + ((int64_t)f * 4 + 2 * dir) * prm.P + p;
    const float ox = __ldcs(land), oy = __ldcs(land + prm.P);     // both before the test: one memory latency, not two
    if (ox == kStaticLand) return;
    SourceCells c;
    source_cells(prm, ox, oy, f, dir, c);
    // Three rounds, each issued for all the thread's cells before the first answer is looked at (a warp pays one
    // memory latency per round, not one per cell): claim the canonical slots; the cells that lost theirs (the flow
    // compresses there: in the benchmark scene some lane of two warps in three) take an overflow slot; store.
    unsigned old[kCellsPerSource], ovf[kCellsPerSource];
    #pragma unroll
    for (int i = 0; i < kCellsPerSource; ++i) {
        old[i] = 0u;
        if (c.slot[i] >= 0) old[i] = atomicOr(&prm.slot_mask[c.at[i] >> 1], (1u << c.slot[i]) << lane_mask_shift(c.at[i]));
    }
    #pragma unroll
    for (int i = 0; i < kCellsPerSource; ++i) {
        ovf[i] = 0xffffffffu;          // "kept the canonical slot"
        if (c.slot[i] >= 0 && (old[i] >> lane_mask_shift(c.at[i]) >> c.slot[i] & 1u)) ovf[i] = atomicAdd(&prm.slot_over[c.at[i]], 1u);
    }
    const int y = (int)(p / prm.W), x = (int)(p - (int64_t)y * prm.W);
    const unsigned xy = pack_xy(x, y);
    #pragma unroll
    for (int i = 0; i < kCellsPerSource; ++i) {
        if (c.slot[i] < 0) continue;
        const int so = ovf[i] == 0xffffffffu ? c.slot[i] : (int)min(ovf[i], (unsigned)kListDepth) + kCanon;
        // streaming stores: the lists (40 MB per frame, next read by the gather) need not stay in L2
        if (so < kListDepth)
            SLR_INSERT_STORE(prm.lists + ((int64_t)(c.at[i] >> 5) * kListDepth + so) * 32 + (c.at[i] & 31u),
                             make_uint4((unsigned)p, __float_as_uint(c.wt[i]), __float_as_uint(c.wb[i]), xy));
        else excess_cell(prm, f, c.at[i], (unsigned)p, c.wt[i], c.wb[i]);
    }
}

// ---------------------------------------------------------------------------
// rowgather_kernel
// One warp per row pair; a CTA is F frames x R row pairs of one destination tile (template
// parameters, default 2 x 2: vertically shared sources and the sources of the same rows in the
// next frame stay in this SM's L1).  CTA order is frame-fastest: the CTAs resident at any moment
// work on the SAME destination tiles of all frames of the batch; their source regions differ
// only by the frame-to-frame displacement, so a source line is fetched from HBM once per batch
// and the other frames hit it in L2 (one frame's features alone, 204 MB, exceed the 126 MB L2).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void fma4(float4& a, const float4& v, float w)
{
    a.x = fmaf(v.x, w, a.x); a.y = fmaf(v.y, w, a.y); a.z = fmaf(v.z, w, a.z); a.w = fmaf(v.w, w, a.w);
}

#ifndef SLR_GATHER_LOADS
#define SLR_GATHER_LOADS 12            // 256-bit loads a warp has in flight per batch (12 KB); measured 4 ... 16, with the bin pipeline (8 was best) and again with the direct index (6 / 8 / 10 / 12: 12 is best): profiles/r02/tune_gather.jsonl
#endif
constexpr int kGatherLoads = SLR_GATHER_LOADS;
// largest divisor B of K with B * GI <= kGatherLoads
__host__ __device__ constexpr int batch_slots(int K, int GI)
{
    int best = 1;
    for (int b = 1; b <= K; ++b)
        if (K % b == 0 && b * GI <= kGatherLoads) best = b;
    return best;
}

struct RowCtx {
    const char* G;        // group plane 0
    const float* S;       // scalar plane 0
    const uint4* list;    // this lane's column of the row-pair list (slot stride 32)
    float* out_top;       // this lane's top pixel in plane 0 of the frame (bottom = + W)
    int64_t P;
    int W, groups, C;
    int my_hi;            // slots [0, my_hi) of THIS lane's column hold entries (direct index: the rest was never
                          // written); the warp's list length is the maximum over its lanes
    float eps;
    bool in_top, in_bot;
    bool raw;             // flagged tile: write un-normalised sums, heavy_finish_kernel divides
};

// entry k >= kRegSlots of the lane's column, or the all-zero pixel with zero weights
__device__ __forceinline__ uint4 deep_entry(const RowCtx& c, int k)
{
    return k < c.my_hi ? __ldcg(c.list + k * 32) : make_uint4((unsigned)c.P, 0u, 0u, 0u);
}

// K  = compile-time number of register-resident slots (the warp's list length rounded up);
// GI = groups of 8 channels per iteration.  Per batch at most 6 independent 256-bit loads (6 KB per warp) are
//      issued before the first FMA that consumes them (a warp pays one full memory latency per batch).
// NZ = count the non-zero outputs (GatherParams::nnz).
template <int NT, int K, int GI, bool NZ>
__device__ __forceinline__ void gather_rows(const RowCtx& c, const unsigned (&pk)[kRegSlots],
                                            const float (&wt)[kRegSlots], const float (&wb)[kRegSlots],
                                            float (&sum_t)[NT + 1], float (&sum_b)[NT + 1], int (&nz)[2])
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
            for (int t = 0; t <= NT; ++t) {
                if (slot_role(k) != kBottomOnly) sum_t[t] = fmaf(sv[t][k], wt[k], sum_t[t]);
                if (slot_role(k) != kTopOnly) sum_b[t] = fmaf(sv[t][k], wb[k], sum_b[t]);
            }
        }
    }
    const int kmax = K == kRegSlots ? __reduce_max_sync(0xffffffffu, c.my_hi) : K;
    if (K == kRegSlots) {
        for (int k = kRegSlots; k < kmax; ++k) {          // rare: lists deeper than the registers hold
            const uint4 e = deep_entry(c, k);
            #pragma unroll
            for (int t = 0; t <= NT; ++t) {
                const float s = __ldg(c.S + (int64_t)t * sstride + (e.x & kPixelMask));
                sum_t[t] = fmaf(s, __uint_as_float(e.y), sum_t[t]);
                sum_b[t] = fmaf(s, __uint_as_float(e.z), sum_b[t]);
            }
        }
    }
    const float inv_t = c.raw ? 1.0f : 1.0f / fmaxf(sum_t[NT], c.eps);
    const float inv_b = c.raw ? 1.0f : 1.0f / fmaxf(sum_b[NT], c.eps);

    const char* Gg = c.G;
    const size_t gstride = (size_t)(c.P + 1) * kGroupBytes;
    const size_t ostride = (size_t)c.P;
    float* o = c.out_top;
    for (int g = 0; g < c.groups; g += GI, Gg += GI * gstride, o += kGroupChannels * GI * ostride) {
        constexpr int B = batch_slots(K, GI);        // slots per batch: B * GI loads in flight
        float8 at[GI], ab[GI];
        #pragma unroll
        for (int gi = 0; gi < GI; ++gi) {
            at[gi].lo = at[gi].hi = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            ab[gi] = at[gi];
        }
        #pragma unroll
        for (int kb = 0; kb < K; kb += B) {
            float8 v[GI][B];
            #pragma unroll
            for (int gi = 0; gi < GI; ++gi) {
                #pragma unroll
                for (int k = 0; k < B; ++k) v[gi][k] = ldg256(px32(Gg + gi * gstride, pk[kb + k]));
            }
            #pragma unroll
            for (int gi = 0; gi < GI; ++gi) {
                #pragma unroll
                for (int k = 0; k < B; ++k) {
                    if (slot_role(kb + k) != kBottomOnly) { fma4(at[gi].lo, v[gi][k].lo, wt[kb + k]); fma4(at[gi].hi, v[gi][k].hi, wt[kb + k]); }
                    if (slot_role(kb + k) != kTopOnly) { fma4(ab[gi].lo, v[gi][k].lo, wb[kb + k]); fma4(ab[gi].hi, v[gi][k].hi, wb[kb + k]); }
                }
            }
        }
        #pragma unroll
        for (int gi = 0; gi < GI; ++gi) {
            if (K == kRegSlots) {
                for (int k = kRegSlots; k < kmax; ++k) {
                    const uint4 e = deep_entry(c, k);
                    const float8 t = ldg256(px32(Gg + gi * gstride, e.x & kPixelMask));
                    const float w0 = __uint_as_float(e.y), w1 = __uint_as_float(e.z);
                    fma4(at[gi].lo, t.lo, w0); fma4(at[gi].hi, t.hi, w0);
                    fma4(ab[gi].lo, t.lo, w1); fma4(ab[gi].hi, t.hi, w1);
                }
            }
            float* og = o + kGroupChannels * gi * ostride;
            const float rt[8] = {at[gi].lo.x * inv_t, at[gi].lo.y * inv_t, at[gi].lo.z * inv_t, at[gi].lo.w * inv_t,
                                 at[gi].hi.x * inv_t, at[gi].hi.y * inv_t, at[gi].hi.z * inv_t, at[gi].hi.w * inv_t};
            const float rb[8] = {ab[gi].lo.x * inv_b, ab[gi].lo.y * inv_b, ab[gi].lo.z * inv_b, ab[gi].lo.w * inv_b,
                                 ab[gi].hi.x * inv_b, ab[gi].hi.y * inv_b, ab[gi].hi.z * inv_b, ab[gi].hi.w * inv_b};
            #pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (kGroupChannels * (g + gi) + j < c.C) {
                    if (c.in_top) __stcs(og + j * ostride, rt[j]);
                    if (c.in_bot) __stcs(og + j * ostride + c.W, rb[j]);
                    if (NZ) { nz[0] += rt[j] != 0.0f; nz[1] += rb[j] != 0.0f; }
                }
            }
        }
    }
}

template <int NT, int K, bool NZ>
__device__ __forceinline__ void gather_rows_dispatch(const RowCtx& c, const unsigned (&pk)[kRegSlots],
                                                     const float (&wt)[kRegSlots], const float (&wb)[kRegSlots],
                                                     float (&sum_t)[NT + 1], float (&sum_b)[NT + 1], int (&nz)[2])
{
    constexpr int GI = K * 2 <= kGatherLoads ? 2 : 1;
    if (c.groups % GI == 0) gather_rows<NT, K, GI, NZ>(c, pk, wt, wb, sum_t, sum_b, nz);
    else gather_rows<NT, K, 1, NZ>(c, pk, wt, wb, sum_t, sum_b, nz);
}


// CTA shape: F frames x R row pairs (warps) of one destination tile, F * R * 32 threads.  Warps of
// the SAME tile in consecutive frames read source regions that differ only by one frame's
// displacement, so with F > 1 they share most of their lines in this SM's L1 as well (the
// frame-fastest CTA order already shares them in L2).  R < 4 splits a tile's row pairs over CTAs.
// DIRECT: the lists were written by insert_kernel: which slots of a lane hold an entry is in its slot word
// (GatherParams::slot_mask), the others were never written and read as the all-zero pixel.
template <int NT, int F, int R, bool NZ, bool DIRECT>
__device__ __forceinline__ void rowgather_item(const GatherParams& prm, unsigned item)
{
    constexpr int kParts = kPairsPerTile / R;           // CTAs per (tile, frame group)
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int n_fg = (prm.n_frames + F - 1) / F;
    const int f = (int)(item % (unsigned)n_fg) * F + warp / R;
    const int part = (int)(item / (unsigned)n_fg);
    const int tile = part / kParts, pr = (part % kParts) * R + warp % R;      // row pair within the tile
    if (f >= prm.n_frames) return;
    const unsigned flag = __ldg(prm.tile_flag + (int64_t)f * prm.n_tiles + tile);
    if (flag == 2u) return;                                   // done entirely by the heavy kernels
    if (prm.only_fallback && __ldcg(prm.fallback + (int64_t)f * prm.n_tiles + tile) == 0u) return;
    const int tx = tile % prm.tiles_x, ty = tile / prm.tiles_x;
    const int X = tx * TW + (tid & 31), Y = ty * TH + 2 * pr;
    const int64_t P = prm.P;
    const int64_t pix = (int64_t)Y * prm.W + X;
    const int64_t pair = (int64_t)f * prm.n_tiles * kPairsPerTile + (int64_t)tile * kPairsPerTile + pr;
    int kmax, my_hi;
    unsigned used = 0xffffffffu;          // register-resident slots of this lane that hold an entry
    unsigned self = 0u;                   // direct index: the lane's top / bottom pixel receives itself (static_lanes_kernel)
    if (DIRECT) {
        const unsigned at = (unsigned)(pair * 32 + (tid & 31));
        const unsigned mask = __ldcg(prm.slot_mask + (at >> 1)) >> lane_mask_shift(at);
        const int n_ovf = (int)min(__ldcg(prm.slot_over + at), (unsigned)(kListDepth - kCanon));      // the rest is in the excess list
        const unsigned canon = mask & ((1u << kCanon) - 1u);
        my_hi = n_ovf > 0 ? kCanon + n_ovf : 32 - __clz((int)canon);
        kmax = __reduce_max_sync(0xffffffffu, my_hi);
        used = canon | ((1u << min(n_ovf, kRegSlots - kCanon)) - 1u) << kCanon;
        self = mask & (kSelfTop | kSelfBottom);
        if (self & kSelfTop) used &= ~1u;          // slots 0 / 1 of a static pixel hold no entry: made up below
        if (self & kSelfBottom) used &= ~2u;
    } else {
        kmax = my_hi = (int)__ldg(prm.row_k + pair);
    }

    RowCtx c;
    c.G = prm.G; c.S = prm.S; c.P = P; c.W = prm.W; c.groups = prm.groups; c.C = prm.C; c.eps = prm.eps;
    c.list = prm.lists + pair * (kListDepth * 32) + (tid & 31);
    c.out_top = prm.out + (int64_t)f * prm.C * P + pix;
    c.in_top = X < prm.W && Y < prm.H;
    c.in_bot = X < prm.W && Y + 1 < prm.H;
    c.my_hi = my_hi;
    c.raw = flag == 1u || prm.plain_sum != 0;

    unsigned pk[kRegSlots];
    float wt[kRegSlots], wb[kRegSlots];
    #pragma unroll
    for (int k = 0; k < kRegSlots; ++k) {
        uint4 e = make_uint4((unsigned)P, 0u, 0u, 0u);
        if (DIRECT ? (used >> k & 1u) != 0u : k < kmax) e = __ldcg(c.list + k * 32);
        pk[k] = e.x & kPixelMask;
        wt[k] = __uint_as_float(e.y);
        wb[k] = __uint_as_float(e.z);
    }
    if (DIRECT) {
        // a pixel with exactly zero motion receives itself with weight a + (1 - a): its forward and backward splat
        // both land exactly on it (slot 0: the top pixel, slot 1: the bottom pixel; never inserted, see insert_kernel)
        const float a_f = prm.alphas.a[f];
        const float w_self = a_f + (1.0f - a_f);
        if (self & kSelfTop) { pk[0] = (unsigned)pix; wt[0] = w_self; wb[0] = 0.0f; }
        if (self & kSelfBottom) { pk[1] = (unsigned)pix + (unsigned)prm.W; wb[1] = w_self; }
    }
    float sum_t[NT + 1] = {0.0f}, sum_b[NT + 1] = {0.0f};
    int nz[2] = {0, 0};
    // the list length is warp-uniform: pick the unroll that fits
    if (kmax <= 2) gather_rows_dispatch<NT, 2, NZ>(c, pk, wt, wb, sum_t, sum_b, nz);
    else if (kmax <= 4) gather_rows_dispatch<NT, 4, NZ>(c, pk, wt, wb, sum_t, sum_b, nz);
    else if (kmax <= 6) gather_rows_dispatch<NT, 6, NZ>(c, pk, wt, wb, sum_t, sum_b, nz);
    else if (kmax <= 8) gather_rows_dispatch<NT, 8, NZ>(c, pk, wt, wb, sum_t, sum_b, nz);
    else if (kmax <= 12) gather_rows_dispatch<NT, 12, NZ>(c, pk, wt, wb, sum_t, sum_b, nz);
    // 12 canonical + 1-2 overflow slots: most rows of compressing flow (+0.5 %; without tail planes only: the 2-layer
    // instantiations spill with one more variant)
    else if (NT == 0 && kmax <= 14) gather_rows_dispatch<NT, 14, NZ>(c, pk, wt, wb, sum_t, sum_b, nz);
    else gather_rows_dispatch<NT, 16, NZ>(c, pk, wt, wb, sum_t, sum_b, nz);

    #pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (!(r ? c.in_bot : c.in_top)) continue;
        const float* sum = r ? sum_b : sum_t;
        const int64_t px = pix + (r ? prm.W : 0);
        if (c.raw) {        // the excess pairs are still to come: leave the sums for heavy_finish_kernel
            float* hs = prm.heavy_sums + (int64_t)f * 3 * P + px;
            #pragma unroll
            for (int j = 0; j <= NT; ++j) hs[(int64_t)j * P] = sum[j];
            continue;
        }
        if (prm.aux) {
            float* a = prm.aux + (int64_t)f * (NT + 1) * P + px;
            #pragma unroll
            for (int j = 0; j <= NT; ++j) a[(int64_t)j * P] = sum[j];
        }
        if (prm.mask) prm.mask[(int64_t)f * P + px] = sum[NT] > prm.eps ? 1.0f : 0.0f;
        if (NZ) prm.nnz[(int64_t)f * P + px] = (float)nz[r];
    }
}

template <int NT, int F, int R, bool NZ, bool DIRECT>
__global__ void __launch_bounds__(32 * F * R, (SLR_GATHER_MINBLOCKS * kCols) / (32 * F * R))
rowgather_kernel(const GatherParams prm)
{
    rowgather_item<NT, F, R, NZ, DIRECT>(prm, blockIdx.x);
}

// ---------------------------------------------------------------------------
// stagegather_kernel: the gather with its sources STAGED IN SHARED MEMORY BY THE TMA UNIT.
//
// rowgather_kernel above pulls every (slot, channel group) with an LDG.128 through L1: on B200 it is
// bound by the L1 data pipe (73 % of peak wavefronts, 43 % of the sectors missing L1; profiles/r01),
// not by HBM.  Here a CTA owns one destination tile in kStageFrames consecutive frames and executes
// the staging plan expand_kernel made for it (StageRecord):
//   * for every chunk of 16 channels the row segments the tile's sources lie in are copied from the Q
//     region of the scene buffer (128-byte blocks of two pixels) into shared memory with cp.async.bulk
//     -- no registers, no L1, no LSU wavefronts; each warp issues a share of the copies (one elected
//     lane), all of them complete on one transaction barrier per stage, and the copies of chunk q+1
//     fly while chunk q is being consumed (two stages; one when the region is large);
//   * the lanes gather with LDS.128 from the offsets their list entries already hold (128 B/clk/SM
//     whatever the alignment; the slot permutation of the blocks keeps 8 neighbouring pixels on 8
//     different bank groups), same register accumulators and normalise-and-store epilogue as
//     rowgather_kernel.
// The two frames of a CTA share one staged region (their sources differ by one frame's displacement), so a
// source pixel crosses the L2 -> SM link once per two frames.  Tiles whose sources do not fit the staging area
// (convergence zones; incoherent flow) were flagged by expand_kernel and are left to rowgather_kernel.
// ---------------------------------------------------------------------------
constexpr int kStageWarps = kStageFrames * kPairsPerTile;      // one warp per (frame, row pair)
constexpr int kStageThreads = 32 * kStageWarps;

struct StageShared {
    unsigned copy_src[kMaxCopies];
    unsigned copy_dst[kMaxCopies];
    tma::Barrier full[2];
};

struct StageCtx {
    const char* Q;            // chunk plane 0
    size_t plane_bytes;       // bytes of a chunk plane
    unsigned char* stage;     // staging area
    tma::SharedAddr stage_addr, full_addr;     // the same and the two barriers, as the copy instruction wants them
    StageShared* sh;
    const uint4* list;        // this lane's column of the row-pair list
    float* out_top;
    int64_t P;
    int W, C, kmax, chunks, warp, n_copies, n_stage;
    unsigned my_bytes;        // bytes per chunk of the copies this warp issues
    float inv_t, inv_b;
    bool in_top, in_bot;
    bool count_nz;            // count the non-zero outputs of the lane's two pixels (GatherParams::nnz)
};

// Issues this warp's share of the copies of chunk q into its stage and announces their bytes.
__device__ __forceinline__ void stage_issue(const StageCtx& c, int q)
{
    const int s = c.n_stage == 2 ? (q & 1) : 0;
    if (tma::elect_one()) {
        const tma::SharedAddr base = c.stage_addr + (tma::SharedAddr)s * (kStageBytes / 2);
        const tma::SharedAddr bar = c.full_addr + (tma::SharedAddr)s * sizeof(tma::Barrier);
        tma::arrive_expect_tx_at(bar, c.my_bytes);
        const char* src = c.Q + (size_t)q * c.plane_bytes;
        for (int i = c.warp; i < c.n_copies; i += kStageWarps) {
            const unsigned d = c.sh->copy_dst[i];
            tma::load_at(base + (d >> kCopyLenBits) * kBlockBytes, src + (size_t)c.sh->copy_src[i] * kBlockBytes,
                         (d & ((1u << kCopyLenBits) - 1u)) * kBlockBytes, bar);
        }
    }
    __syncwarp();
}

// normalise and store 4 channels of the lane's two pixels
__device__ __forceinline__ void stage_store(const StageCtx& c, float* og, int ch0, const float4& at, const float4& ab, int (&nz)[2])
{
    const size_t ostride = (size_t)c.P;
    const float rt[4] = {at.x * c.inv_t, at.y * c.inv_t, at.z * c.inv_t, at.w * c.inv_t};
    const float rb[4] = {ab.x * c.inv_b, ab.y * c.inv_b, ab.z * c.inv_b, ab.w * c.inv_b};
    #pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (ch0 + j < c.C) {
            if (c.in_top) __stcs(og + j * ostride, rt[j]);
            if (c.in_bot) __stcs(og + j * ostride + c.W, rb[j]);
            if (c.count_nz) { nz[0] += rt[j] != 0.0f; nz[1] += rb[j] != 0.0f; }
        }
    }
}

// The chunk loop of one warp.  K = compile-time number of register-resident slots (0: a warp without
// work -- no frame, or a whole-tile heavy tile -- that only takes part in the copies and barriers).
// The copies of chunk 0 were issued by the caller.
template <int NT, int K>
__device__ __forceinline__ void stage_rows(const StageCtx& c, const unsigned (&pk)[kRegSlots],
                                           const float (&wt)[kRegSlots], const float (&wb)[kRegSlots], int (&nz)[2])
{
    const int n_stage = c.n_stage;
    const size_t ostride = (size_t)c.P;
    for (int q = 0; q < c.chunks; ++q) {
        if (n_stage == 2 && q + 1 < c.chunks) {
            // the stage chunk q+1 goes to was read for chunk q-1: every warp must be done with it
            if (q >= 1) __syncthreads();
            stage_issue(c, q + 1);
        }
        const int s = n_stage == 2 ? (q & 1) : 0;
        tma::wait(&c.sh->full[s], (unsigned)(n_stage == 2 ? (q >> 1) : q) & 1u);
        if (K > 0) {
            const unsigned char* base = c.stage + (size_t)s * (kStageBytes / 2);
            float* o = c.out_top + (size_t)q * kChunkChannels * ostride;
            constexpr int B = K == 0 ? 1 : (K <= 8 ? K : K / 2);      // loads in flight (shared-memory latency is short)
            if (K == kRegSlots && c.kmax > kRegSlots) {
                // Lists deeper than the registers hold (convergence zones): all four channel groups of the chunk are
                // accumulated together, so that a deep slot's entry is fetched once per chunk (not once per group)
                // and its four shared-memory loads overlap.  Same FMAs in the same order per accumulator.
                float4 at[4], ab[4];
                #pragma unroll
                for (int u = 0; u < 4; ++u) {
                    at[u] = ab[u] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    #pragma unroll
                    for (int kb = 0; kb < K; kb += 4) {
                        float4 v[4];
                        #pragma unroll
                        for (int k = 0; k < 4; ++k) v[k] = *reinterpret_cast<const float4*>(base + (pk[kb + k] ^ ((unsigned)u << 4)));
                        #pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (slot_role(kb + k) != kBottomOnly) fma4(at[u], v[k], wt[kb + k]);
                            if (slot_role(kb + k) != kTopOnly) fma4(ab[u], v[k], wb[kb + k]);
                        }
                    }
                }
                uint4 e = __ldca(c.list + kRegSlots * 32);
                for (int k = kRegSlots; k < c.kmax; ++k) {
                    float4 t[4];
                    #pragma unroll
                    for (int u = 0; u < 4; ++u) t[u] = *reinterpret_cast<const float4*>(base + (e.x ^ ((unsigned)u << 4)));
                    const float w0 = __uint_as_float(e.y), w1 = __uint_as_float(e.z);
                    if (k + 1 < c.kmax) e = __ldca(c.list + (k + 1) * 32);        // the next entry flies during the FMAs
                    #pragma unroll
                    for (int u = 0; u < 4; ++u) { fma4(at[u], t[u], w0); fma4(ab[u], t[u], w1); }
                }
                #pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (q * kChunkChannels + 4 * u < c.C) stage_store(c, o + (size_t)(4 * u) * ostride, q * kChunkChannels + 4 * u, at[u], ab[u], nz);
            } else {
                #pragma unroll
                for (int u = 0; u < 4; ++u) {                  // the chunk's four channel groups
                    const int ch0 = q * kChunkChannels + 4 * u;
                    if (ch0 < c.C) {
                        float4 at = make_float4(0.0f, 0.0f, 0.0f, 0.0f), ab = at;
                        #pragma unroll
                        for (int kb = 0; kb < K; kb += B) {
                            float4 v[B];
                            #pragma unroll
                            for (int k = 0; k < B; ++k)
                                v[k] = *reinterpret_cast<const float4*>(base + (pk[kb + k] ^ ((unsigned)u << 4)));
                            #pragma unroll
                            for (int k = 0; k < B; ++k) {
                                if (slot_role(kb + k) != kBottomOnly) fma4(at, v[k], wt[kb + k]);
                                if (slot_role(kb + k) != kTopOnly) fma4(ab, v[k], wb[kb + k]);
                            }
                        }
                        stage_store(c, o + (size_t)(4 * u) * ostride, ch0, at, ab, nz);
                    }
                }
            }
        }
        if (n_stage == 1 && q + 1 < c.chunks) { __syncthreads(); stage_issue(c, q + 1); }
    }
}

template <int NT>
__global__ void __launch_bounds__(kStageThreads, 2)
stagegather_kernel(const GatherParams prm)
{
    SLR_DYNAMIC_SMEM(stage_mem);
    __shared__ StageShared sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_fg = (prm.n_frames + kStageFrames - 1) / kStageFrames;
    const int fg = (int)(blockIdx.x % (unsigned)n_fg), tile = (int)(blockIdx.x / (unsigned)n_fg);
    const StageRecord* rec = prm.records + (int64_t)fg * prm.n_tiles + tile;
    const int n_stage = (int)__ldcg(&rec->stages);
    if (n_stage == 0) return;               // the sources do not fit: rowgather_kernel does this tile
    const int n_copies = (int)__ldcg(&rec->n_copies);
    const int f = fg * kStageFrames + warp / kPairsPerTile, pr = warp % kPairsPerTile;
    const bool have_frame = f < prm.n_frames;
    const unsigned flag = have_frame ? __ldg(prm.tile_flag + (int64_t)f * prm.n_tiles + tile) : 2u;
    const bool active = flag != 2u;                      // warp-uniform; flag 2 = done entirely by the heavy kernels
    const int64_t P = prm.P;

    for (int i = tid; i < n_copies; i += kStageThreads) {
        sh.copy_src[i] = __ldcg(rec->copy_src + i);
        sh.copy_dst[i] = __ldcg(rec->copy_dst + i);
    }
    if (tid < 16)     // staged block 0 of either stage: the all-zero block unused list slots read
        reinterpret_cast<float4*>(stage_mem + (size_t)(tid >> 3) * (kStageBytes / 2))[tid & 7] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (tid == 0) {
        tma::init(&sh.full[0], kStageWarps);
        tma::init(&sh.full[1], kStageWarps);
        tma::fence_init();
    }

    const int tx = tile % prm.tiles_x, ty = tile / prm.tiles_x;
    const int X = tx * TW + lane, Y = ty * TH + 2 * pr;
    const int64_t pix = (int64_t)Y * prm.W + X;
    const int64_t pair = (int64_t)f * prm.n_tiles * kPairsPerTile + (int64_t)tile * kPairsPerTile + pr;
    const int kmax = active ? (int)__ldg(prm.row_k + pair) : 0;

    StageCtx c;
    c.Q = prm.Q; c.plane_bytes = (size_t)quilt_plane_blocks(prm.H, prm.W) * kBlockBytes;
    c.stage = stage_mem; c.sh = &sh; c.P = P; c.W = prm.W; c.C = prm.C; c.kmax = kmax;
    c.stage_addr = tma::shared_addr(stage_mem); c.full_addr = tma::shared_addr(&sh.full[0]);
    c.chunks = (int)scene_chunks(prm.C); c.warp = warp; c.n_copies = n_copies; c.n_stage = n_stage;
    c.my_bytes = __ldcg(rec->warp_bytes + warp);
    c.list = prm.lists + pair * (kListDepth * 32) + lane;
    c.out_top = prm.out + (int64_t)f * prm.C * P + pix;
    c.in_top = active && X < prm.W && Y < prm.H;
    c.in_bot = active && X < prm.W && Y + 1 < prm.H;

    __syncthreads();                      // copy list, zero blocks and barriers are in place
    stage_issue(c, 0);                    // the first chunk's copies fly while the lists are read

    // ---- the lists (staged offsets and weights) and the sums of the scalar planes expand_kernel left
    unsigned pk[kRegSlots];
    float wt[kRegSlots], wb[kRegSlots];
    #pragma unroll
    for (int k = 0; k < kRegSlots; ++k) {
        uint4 e = make_uint4(0u, 0u, 0u, 0u);
        if (k < kmax) e = __ldcg(c.list + k * 32);
        pk[k] = e.x;
        wt[k] = __uint_as_float(e.y);
        wb[k] = __uint_as_float(e.z);
    }
    float sum_t[NT + 1] = {0.0f}, sum_b[NT + 1] = {0.0f};
    {
        const float* hs = prm.heavy_sums + (int64_t)f * 3 * P + pix;
        #pragma unroll
        for (int j = 0; j <= NT; ++j) {
            if (c.in_top) sum_t[j] = __ldcg(hs + (int64_t)j * P);
            if (c.in_bot) sum_b[j] = __ldcg(hs + (int64_t)j * P + prm.W);
        }
    }
    const bool raw = flag == 1u;          // flagged tile: un-normalised outputs, heavy_excess / heavy_finish complete them
    c.inv_t = raw ? 1.0f : 1.0f / fmaxf(sum_t[NT], prm.eps);
    c.inv_b = raw ? 1.0f : 1.0f / fmaxf(sum_b[NT], prm.eps);

    // the list length is warp-uniform: pick the unroll that fits
    int nz[2] = {0, 0};
    c.count_nz = prm.nnz != nullptr && !raw;
    if (!active) stage_rows<NT, 0>(c, pk, wt, wb, nz);
    else if (kmax <= 2) stage_rows<NT, 2>(c, pk, wt, wb, nz);
    else if (kmax <= 4) stage_rows<NT, 4>(c, pk, wt, wb, nz);
    else if (kmax <= 6) stage_rows<NT, 6>(c, pk, wt, wb, nz);
    else if (kmax <= 8) stage_rows<NT, 8>(c, pk, wt, wb, nz);
    else if (kmax <= 12) stage_rows<NT, 12>(c, pk, wt, wb, nz);
    else stage_rows<NT, 16>(c, pk, wt, wb, nz);

    if (raw) return;                      // sums stay where they are for heavy_excess_kernel / heavy_finish_kernel
    #pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (!(r ? c.in_bot : c.in_top)) continue;
        const float* sum = r ? sum_b : sum_t;
        const int64_t px = pix + (r ? prm.W : 0);
        if (prm.aux) {
            float* a = prm.aux + (int64_t)f * (NT + 1) * P + px;
            #pragma unroll
            for (int j = 0; j <= NT; ++j) a[(int64_t)j * P] = sum[j];
        }
        if (prm.mask) prm.mask[(int64_t)f * P + px] = sum[NT] > prm.eps ? 1.0f : 0.0f;
        if (prm.nnz) prm.nnz[(int64_t)f * P + px] = (float)nz[r];
    }
}

// ---------------------------------------------------------------------------
// Heavy tiles: destination tiles in which some lane's list is deeper than kListDepth
// (convergence zones of the flow: the synthetic 60-step fields pile up to ~150 sources on
// single pixels and 10x the average number of pairs on single tiles).  They are done the way
// the reference does everything: scatter with fp32 reductions at L2 (native RED.ADD.F32;
// shared-memory float atomics are CAS loops and collapse under this contention).  Work is
// per PAIR, so the cost is linear in the number of pairs whatever their distribution:
//   heavy_prepare  zeroes the tile's outputs and its (tail..., norm) sums,
//   heavy_scatter  one CTA per (heavy tile, chunk of channel groups): RED every pair in,
//   heavy_finish   divides by the norm.
// Fixed grids walk the compacted list of heavy tiles that expand_kernel produced.
// ---------------------------------------------------------------------------
struct HeavyTile {
    int f, tx, ty;
    int64_t pix;
    bool inframe;
};

__device__ __forceinline__ HeavyTile heavy_tile(const GatherParams& prm, unsigned item, int tid)
{
    HeavyTile t;
    t.f = (int)(item % (unsigned)prm.n_frames);
    const int tile = (int)(item / (unsigned)prm.n_frames);
    t.tx = tile % prm.tiles_x;
    t.ty = tile / prm.tiles_x;
    const int X = t.tx * TW + (tid & 31), Y = t.ty * TH + (tid >> 5);
    t.inframe = X < prm.W && Y < prm.H;
    t.pix = (int64_t)Y * prm.W + X;
    return t;
}

__global__ void __launch_bounds__(TILE)
heavy_prepare_kernel(const GatherParams prm, int n_sums)
{
    const unsigned n = *prm.flag_count;
    for (unsigned i = blockIdx.x; i < n; i += gridDim.x) {
        const unsigned item = prm.flag_list[i];
        const HeavyTile t = heavy_tile(prm, item, threadIdx.x);
        const int tile = t.ty * prm.tiles_x + t.tx;
        if (prm.tile_flag[(int64_t)t.f * prm.n_tiles + tile] != 2u) continue;    // flag 1: the gather wrote raw sums
        if (!t.inframe) continue;
        float* out = prm.out + (int64_t)t.f * prm.C * prm.P + t.pix;
        for (int c = 0; c < prm.C; ++c) out[(int64_t)c * prm.P] = 0.0f;
        float* sums = prm.heavy_sums + (int64_t)t.f * 3 * prm.P + t.pix;
        for (int j = 0; j < n_sums; ++j) sums[(int64_t)j * prm.P] = 0.0f;
    }
}

template <int NT>
__global__ void __launch_bounds__(TILE, 4)
heavy_scatter_kernel(const GatherParams prm)
{
    const int tid = threadIdx.x;
    const int64_t P = prm.P;
    const int64_t sstride = P + 1;
    // chunk 0 = the scalar planes (tail..., e^Z); chunk c > 0 = channel groups [(c-1)*kHeavyGroups, ...)
    const unsigned n_chunks = 1u + (unsigned)((2 * prm.groups + kHeavyGroups - 1) / kHeavyGroups);
    const unsigned n_work = *prm.flag_count * n_chunks;
    for (unsigned wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
        const unsigned item = prm.flag_list[wi / n_chunks];
        const int chunk = (int)(wi % n_chunks);
        const HeavyTile t = heavy_tile(prm, item, tid);
        const int tile = t.ty * prm.tiles_x + t.tx;
        if (prm.tile_flag[(int64_t)t.f * prm.n_tiles + tile] != 2u) continue;
        const float a_f = prm.alphas.a[t.f], a_b = 1.0f - a_f;
        const unsigned* off = prm.offsets + (int64_t)t.f * (prm.n_tiles + 1);
        const unsigned beg = off[tile], end = off[tile + 1];
        const float4* ent = prm.ent + (int64_t)t.f * prm.cap;
        const bool self_static = t.inframe && __ldg(prm.motion + t.pix) == 0.0f && __ldg(prm.motion + P + t.pix) == 0.0f;
        float* out = prm.out + (int64_t)t.f * prm.C * P;
        float* sums = prm.heavy_sums + (int64_t)t.f * 3 * P;
        const int groups4 = 2 * prm.groups;           // groups of four channels
        const int g_lo = (chunk - 1) * kHeavyGroups, g_hi = min(groups4, g_lo + kHeavyGroups);

        // one (destination pixel, source pixel, weight) pair
        auto add_pair = [&](int64_t dpix, unsigned p, float w) {
            if (chunk == 0) {
                #pragma unroll
                for (int j = 0; j <= NT; ++j)
                    red_add(sums + (int64_t)j * P + dpix, __ldg(prm.S + (int64_t)j * sstride + p) * w);
            } else {
                float4 v[kHeavyGroups];
                #pragma unroll
                for (int gi = 0; gi < kHeavyGroups; ++gi)
                    v[gi] = g_lo + gi < g_hi ? ldg_group4(prm.G, P, g_lo + gi, p) : make_float4(0.f, 0.f, 0.f, 0.f);
                #pragma unroll
                for (int gi = 0; gi < kHeavyGroups; ++gi) {
                    const float r[4] = {v[gi].x * w, v[gi].y * w, v[gi].z * w, v[gi].w * w};
                    #pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = 4 * (g_lo + gi) + j;
                        if (c < prm.C) red_add(out + (int64_t)c * P + dpix, r[j]);
                    }
                }
            }
        };

        if (self_static) add_pair(t.pix, (unsigned)t.pix, a_f + a_b);
        for (unsigned e = beg + tid; e < end; e += TILE) {
            const float4 en = __ldg(ent + e);
            const unsigned pd = __float_as_uint(en.x);
            const Footprint fp = footprint_at(en.y, en.z, prm.H, prm.W);
            const float a = (pd >> 31) ? a_b : a_f;
            #pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int cx = fp.x0 + (k & 1), cy = fp.y0 + (k >> 1);
                const int lx = cx - t.tx * TW, ly = cy - t.ty * TH;
                const float wa = fp.w[k] * a;
                if ((fp.ok >> k & 1u) && lx >= 0 && lx < TW && ly >= 0 && ly < TH && wa != 0.0f)
                    add_pair((int64_t)cy * prm.W + cx, pd & ~kDirBit, wa);
            }
        }
    }
}

// One (destination pixel, source pixel, weight) pair of frame f by fp32 reductions at L2: onto the un-normalised
// sums in `out` / `heavy_sums`.
template <int NT>
__device__ __forceinline__ void red_pair(const GatherParams& prm, int f, int64_t dpix, unsigned src, float w)
{
    const int64_t P = prm.P;
    const int64_t sstride = P + 1;
    float* sums = prm.heavy_sums + (int64_t)f * 3 * P + dpix;
    #pragma unroll
    for (int j = 0; j <= NT; ++j) red_add(sums + (int64_t)j * P, __ldg(prm.S + (int64_t)j * sstride + src) * w);
    float* out = prm.out + (int64_t)f * prm.C * P + dpix;
    for (int g0 = 0; g0 < 2 * prm.groups; g0 += 4) {     // groups of four channels, four loads in flight per thread
        float4 v[4];
        #pragma unroll
        for (int gi = 0; gi < 4; ++gi)
            v[gi] = g0 + gi < 2 * prm.groups ? ldg_group4(prm.G, P, g0 + gi, src) : make_float4(0.f, 0.f, 0.f, 0.f);
        #pragma unroll
        for (int gi = 0; gi < 4; ++gi) {
            const float r[4] = {v[gi].x * w, v[gi].y * w, v[gi].z * w, v[gi].w * w};
            #pragma unroll
            for (int j = 0; j < 4; ++j)
                if (4 * (g0 + gi) + j < prm.C) red_add(out + (int64_t)(4 * (g0 + gi) + j) * P, r[j]);
        }
    }
}

// The pairs that did not fit the lists of flag-1 tiles: one thread per pair, fp32 reductions at
// L2 onto the un-normalised sums the gather left in `out` / `heavy_sums`.
template <int NT>
__global__ void __launch_bounds__(TILE)
heavy_excess_kernel(const GatherParams prm)
{
    if (prm.direct && *prm.excess_count > prm.excess_cap) return;      // the batch is redone by overflow_*_kernel
    const unsigned n = min(*prm.excess_count, prm.excess_cap);
    for (unsigned i = blockIdx.x * TILE + threadIdx.x; i < n; i += gridDim.x * TILE) {
        const uint4 e = __ldcg(prm.excess + i);
        const int f = (int)e.w;
        const int64_t dpix = e.x;
        const int tile = (int)(dpix / prm.W) / TH * prm.tiles_x + (int)(dpix % prm.W) / TW;
        if (prm.tile_flag[(int64_t)f * prm.n_tiles + tile] != 1u) continue;     // flag 2: done from the bin
        red_pair<NT>(prm, f, dpix, e.y, __uint_as_float(e.z));
    }
}

// Divides the un-normalised sums of one pixel by its norm (and writes the optional planes).
template <int NT>
__device__ __forceinline__ void finish_pixel(const GatherParams& prm, int f, int64_t pix)
{
    const int64_t P = prm.P;
    const float* sums = prm.heavy_sums + (int64_t)f * 3 * P + pix;
    const float nrm = __ldcg(sums + (int64_t)NT * P);
    const float inv = 1.0f / fmaxf(nrm, prm.eps);
    float* out = prm.out + (int64_t)f * prm.C * P + pix;
    int nz = 0;
    // eight independent loads in flight per thread (one load -> store chain per channel is a
    // full memory latency each: 64 of them made this kernel 33 us whatever the tile count)
    for (int c0 = 0; c0 < prm.C; c0 += 8) {
        float v[8];
        #pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = c0 + j < prm.C ? __ldcg(out + (int64_t)(c0 + j) * P) : 0.0f;
        #pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (c0 + j < prm.C) {
                const float r = v[j] * inv;
                out[(int64_t)(c0 + j) * P] = r;
                nz += r != 0.0f;
            }
        }
    }
    if (prm.nnz) prm.nnz[(int64_t)f * P + pix] = (float)nz;
    if (prm.aux) {
        float* a = prm.aux + (int64_t)f * (NT + 1) * P + pix;
        #pragma unroll
        for (int j = 0; j <= NT; ++j) a[(int64_t)j * P] = __ldcg(sums + (int64_t)j * P);
    }
    if (prm.mask) prm.mask[(int64_t)f * P + pix] = nrm > prm.eps ? 1.0f : 0.0f;
}

template <int NT>
__global__ void __launch_bounds__(TILE)
heavy_finish_kernel(const GatherParams prm)
{
    if (prm.plain_sum) return;                                         // un-normalised sums are the result
    if (prm.direct && *prm.excess_count > prm.excess_cap) return;      // the batch is redone by overflow_*_kernel
    const unsigned n = *prm.flag_count;
    for (unsigned i = blockIdx.x; i < n; i += gridDim.x) {
        const HeavyTile t = heavy_tile(prm, prm.flag_list[i], threadIdx.x);
        if (t.inframe) finish_pixel<NT>(prm, t.f, t.pix);
    }
}

// ---------------------------------------------------------------------------
// Direct index, excess list full (a flow that piles more than excess_cap pairs beyond the list depth onto
// single lanes -- adversarial input; never seen on a motion field of the benchmark): pairs were dropped, so the
// whole batch is redone the way the reference does every frame: scatter with fp32 reductions, then divide.
// Always launched, each kernel leaves at once unless the counter says "overflowed": correctness for any
// input at the price of three empty launches per batch.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
overflow_zero_kernel(const GatherParams prm)
{
    if (*prm.excess_count <= prm.excess_cap) return;
    const int64_t n_out = (int64_t)prm.n_frames * prm.C * prm.P, n_sum = (int64_t)prm.n_frames * 3 * prm.P;
    const int64_t step = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_out; i += step) prm.out[i] = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_sum; i += step) prm.heavy_sums[i] = 0.0f;
}

template <int NT>
__global__ void __launch_bounds__(256)
overflow_scatter_kernel(const GatherParams prm)
{
    if (*prm.excess_count <= prm.excess_cap) return;
    const int64_t total = prm.P * prm.n_frames;
    for (int64_t i0 = (int64_t)blockIdx.x * 256 + threadIdx.x; i0 < total; i0 += (int64_t)gridDim.x * 256) {
        const int f = (int)(i0 / prm.P);
        const int64_t p = i0 - (int64_t)f * prm.P;
        const float* land = prm.refs->land + (int64_t)f * 4 * prm.P + p;
        if (__ldcs(land) == kStaticLand) {        // a pixel with zero motion receives itself with weight a + (1 - a)
            const float a_f = prm.alphas.a[f];
            red_pair<NT>(prm, f, p, (unsigned)p, a_f + (1.0f - a_f));
            continue;
        }
        #pragma unroll 1
        for (int dir = 0; dir < 2; ++dir) {
            SourceCells c;
            source_cells(prm, __ldcs(land + 2 * dir * prm.P), __ldcs(land + (2 * dir + 1) * prm.P), f, dir, c);
            #pragma unroll 1
            for (int j = 0; j < 2 * kCellsPerSource; ++j) {       // (cell, row): not unrolled -- this path is never hot
                const int i = j >> 1;
                const float w = (j & 1) ? c.wb[i] : c.wt[i];
                if (c.slot[i] < 0 || w == 0.0f) continue;
                int cx, ytop;
                lane_pixel(prm, c.at[i], cx, ytop);
                red_pair<NT>(prm, f, (int64_t)(ytop + (j & 1)) * prm.W + cx, (unsigned)p, w);
            }
        }
    }
}

template <int NT>
__global__ void __launch_bounds__(256)
overflow_finish_kernel(const GatherParams prm)
{
    if (*prm.excess_count <= prm.excess_cap || prm.plain_sum) return;
    const int64_t total = prm.P * prm.n_frames;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256)
        finish_pixel<NT>(prm, (int)(i / prm.P), i % prm.P);
}

}  // namespace slr

// ===========================================================================
// C ABI
// ===========================================================================
using namespace slr;
using slr_host::carve;
using slr_host::Workspace;

namespace {

// Fills the kernel parameters shared by slr_clip_expand and slr_clip_gather.
int make_params(GatherParams& prm, const void* scene, const float* motion, int64_t C, int n_tail,
                int64_t H, int64_t W, int start, int end, int t0, int n_frames, float alpha_lo, float alpha_hi,
                float* out, float* aux, float* mask, float* nnz, const void* workspace, size_t workspace_bytes)
{
    SLR_CHECK_ARGS(scene && motion && workspace && C > 0 && H > 0 && W > 0 && H * W < (1ll << 27) &&
                   n_tail >= 0 && n_tail <= 2 && n_frames > 0 && n_frames <= kMaxFrames &&
                   t0 >= start && t0 + n_frames - 1 <= end + 1 && ((uintptr_t)workspace & 15) == 0,
                   "slr_clip_expand / slr_clip_gather: bad arguments");
    const int64_t P = H * W;
    const int tiles_x = (int)((W + TW - 1) / TW), tiles_y = (int)((H + TH - 1) / TH);
    const Workspace ws = carve(const_cast<void*>(workspace), H, W, n_frames);
    SLR_CHECK_ARGS(ws.bytes <= workspace_bytes, "workspace too small (see slr_clip_workspace_bytes)");
    const int groups = (int)scene_groups8(C);
    prm.G = (const char*)scene;
    prm.S = (const float*)scene + (int64_t)groups * kGroupChannels * (P + 1);
    prm.Q = (const char*)((const float*)scene + scene_quilt_offset_floats(C, n_tail, P));
    prm.fallback = ws.fallback;
    prm.only_fallback = 0;
    prm.records = ws.records;
    prm.n_tail = n_tail;
    // the plan packs a source row into 14 bits and a column into 16 (plan_key)
    prm.staged = slr_host::gather_staged() && H <= 16384 && W <= 65535 ? 1 : 0;
    prm.direct = slr_host::index_direct() ? 1 : 0;
    prm.plain_sum = 0;
    prm.slot_mask = ws.slot_mask; prm.slot_over = ws.slot_over; prm.refs = ws.refs;
    prm.ent = ws.ent; prm.motion = motion; prm.offsets = ws.offsets;
    prm.lists = ws.lists; prm.row_k = ws.row_k; prm.tile_flag = ws.tile_flag;
    prm.flag_list = ws.flag_list; prm.flag_count = ws.flag_count; prm.heavy_sums = ws.heavy_sums;
    prm.excess = ws.excess; prm.excess_count = ws.excess_count; prm.excess_cap = ws.excess_cap;
    prm.out = out; prm.aux = aux; prm.mask = mask; prm.nnz = nnz;
    prm.C = (int)C; prm.groups = groups; prm.H = (int)H; prm.W = (int)W;
    prm.tiles_x = tiles_x; prm.n_tiles = tiles_x * tiles_y; prm.n_frames = n_frames;
    prm.P = P; prm.cap = prm.direct ? 0 : 8 * P; prm.eps = 1e-8f;
    for (int f = 0; f < n_frames; ++f) {
        // alpha = 1 - (t - start) / (end - start + 1) in fp32 (animating_softmax_splating.py:860),
        // optionally clamped (2layers...py:952)
        float a = 1.0f - (float)(t0 + f - start) / (float)(end - start + 1);
        a = fminf(fmaxf(a, alpha_lo), alpha_hi);
        prm.alphas.a[f] = a;
    }
    return 0;
}

// CTA shape of rowgather_kernel: "<frames>x<row pairs>", 2x2 (default, chosen by measurement:
// profiles/README.md) or 1x4; the environment variable SLR_GATHER_SHAPE overrides the default for A/B runs.
struct GatherShape { int frames, pairs; };

GatherShape gather_shape()
{
    GatherShape g = {SLR_GATHER_FRAMES, SLR_GATHER_PAIRS};
    const char* e = getenv("SLR_GATHER_SHAPE");
    int f = 0, r = 0;
    if (e && sscanf(e, "%dx%d", &f, &r) == 2) { g.frames = f; g.pairs = r; }
    return g;
}

template <int NT>
void launch_stagegather(const GatherParams& prm, unsigned grid, cudaStream_t s)
{
    static bool configured = false;      // per process and instantiation; setting it twice is harmless
    if (!configured) {
        cudaFuncSetAttribute((const void*)stagegather_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageBytes);
        configured = true;
    }
    stagegather_kernel<NT><<<grid, kStageThreads, kStageBytes, s>>>(prm);
}

template <int NT>
void launch_heavy(const GatherParams& prm, unsigned heavy_grid, cudaStream_t s)
{
    if (!prm.direct) {           // whole-tile heavy tiles (flag 2) exist with bins only
        heavy_prepare_kernel<<<heavy_grid, TILE, 0, s>>>(prm, NT + 1);
        heavy_scatter_kernel<NT><<<heavy_grid, TILE, 0, s>>>(prm);
    }
    heavy_excess_kernel<NT><<<heavy_grid, TILE, 0, s>>>(prm);
    heavy_finish_kernel<NT><<<heavy_grid, TILE, 0, s>>>(prm);
    if (prm.direct) {
        // fixed small grids (grid-stride loops): an empty launch costs its CTAs
        overflow_zero_kernel<<<heavy_grid, 256, 0, s>>>(prm);
        overflow_scatter_kernel<NT><<<heavy_grid, 256, 0, s>>>(prm);
        overflow_finish_kernel<NT><<<heavy_grid, 256, 0, s>>>(prm);
    }
}

template <int F, int R, bool NZ, bool DIRECT>
void launch_rowgather_nz(const GatherParams& prm, int n_tail, cudaStream_t s)
{
    const unsigned grid = (unsigned)prm.n_tiles * (unsigned)(kPairsPerTile / R) * (unsigned)((prm.n_frames + F - 1) / F);
    if (n_tail == 0) rowgather_kernel<0, F, R, NZ, DIRECT><<<grid, 32 * F * R, 0, s>>>(prm);
    else if (n_tail == 1) rowgather_kernel<1, F, R, NZ, DIRECT><<<grid, 32 * F * R, 0, s>>>(prm);
    else rowgather_kernel<2, F, R, NZ, DIRECT><<<grid, 32 * F * R, 0, s>>>(prm);
}

template <int F, int R>
void launch_rowgather(const GatherParams& prm, int n_tail, cudaStream_t s)
{
    if (prm.direct) {
        if (prm.nnz) launch_rowgather_nz<F, R, true, true>(prm, n_tail, s);
        else launch_rowgather_nz<F, R, false, true>(prm, n_tail, s);
    } else {
        if (prm.nnz) launch_rowgather_nz<F, R, true, false>(prm, n_tail, s);
        else launch_rowgather_nz<F, R, false, false>(prm, n_tail, s);
    }
}

}  // namespace

extern "C" int slr_clip_expand(const void* scene, const float* motion, int64_t C, int n_tail, int64_t H, int64_t W,
                               int start, int end, int t0, int n_frames, float alpha_lo, float alpha_hi,
                               void* workspace, size_t workspace_bytes, slr_stream_t stream_)
{
    GatherParams prm;
    const int rc = make_params(prm, scene, motion, C, n_tail, H, W, start, end, t0, n_frames, alpha_lo, alpha_hi,
                               nullptr, nullptr, nullptr, nullptr, workspace, workspace_bytes);
    if (rc) return rc;
    if (prm.direct) {
        insert_kernel<<<dim3((unsigned)((prm.P + 255) / 256), 2u * (unsigned)n_frames), 256, 0, (cudaStream_t)stream_>>>(prm);
        return SLR_LAUNCH_STATUS();
    }
    const int per_cta = prm.staged ? kStageFrames : 1;
    const unsigned grid = (unsigned)prm.n_tiles * (unsigned)((n_frames + per_cta - 1) / per_cta);
    if (prm.staged) expand_kernel<true><<<grid, TILE, 0, (cudaStream_t)stream_>>>(prm);
    else expand_kernel<false><<<grid, TILE, 0, (cudaStream_t)stream_>>>(prm);
    return SLR_LAUNCH_STATUS();
}

extern "C" int slr_clip_gather(const void* scene, const float* motion, int64_t C, int n_tail, int64_t H, int64_t W,
                               int start, int end, int t0, int n_frames, float alpha_lo, float alpha_hi,
                               float* out, float* aux, float* mask, float* nnz,
                               const void* workspace, size_t workspace_bytes, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(out, "slr_clip_gather: bad arguments");
    GatherParams prm;
    const int rc = make_params(prm, scene, motion, C, n_tail, H, W, start, end, t0, n_frames, alpha_lo, alpha_hi,
                               out, aux, mask, nnz, workspace, workspace_bytes);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream_;
    if (prm.staged) {
        // sources staged in shared memory by the TMA unit; the tiles whose sources do not fit are flagged ...
        const unsigned grid = (unsigned)prm.n_tiles * (unsigned)((n_frames + kStageFrames - 1) / kStageFrames);
        if (n_tail == 0) launch_stagegather<0>(prm, grid, s);
        else if (n_tail == 1) launch_stagegather<1>(prm, grid, s);
        else launch_stagegather<2>(prm, grid, s);
        const int rc2 = SLR_LAUNCH_STATUS();
        if (rc2) return rc2;
        prm.only_fallback = 1;          // ... and done by the L1 gather below
    }
    const GatherShape shape = gather_shape();
    if (shape.frames == 1 && shape.pairs == 4) launch_rowgather<1, 4>(prm, n_tail, s);
    else launch_rowgather<2, 2>(prm, n_tail, s);
    return SLR_LAUNCH_STATUS();
}

extern "C" int slr_clip_heavy(const void* scene, const float* motion, int64_t C, int n_tail, int64_t H, int64_t W,
                              int start, int end, int t0, int n_frames, float alpha_lo, float alpha_hi,
                              float* out, float* aux, float* mask, float* nnz,
                              const void* workspace, size_t workspace_bytes, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(out, "slr_clip_heavy: bad arguments");
    GatherParams prm;
    const int rc = make_params(prm, scene, motion, C, n_tail, H, W, start, end, t0, n_frames, alpha_lo, alpha_hi,
                               out, aux, mask, nnz, workspace, workspace_bytes);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream_;
    const unsigned grid = (unsigned)prm.n_tiles * (unsigned)n_frames;
    const unsigned heavy_grid = std::min<unsigned>(grid, 8u * (unsigned)slr_host::sm_count());
    if (n_tail == 0) launch_heavy<0>(prm, heavy_grid, s);
    else if (n_tail == 1) launch_heavy<1>(prm, heavy_grid, s);
    else launch_heavy<2>(prm, heavy_grid, s);
    return SLR_LAUNCH_STATUS();
}

// ---------------------------------------------------------------------------
// The operator-level summation splat (kernel_Softsplat_updateOutput, softsplat.py:157-202, as _FunctionSoftsplat.forward
// launches it, :407-416) through the gather: per batch element the input is interleaved like a scene (no importance:
// weight 1), the flow becomes a one-frame, one-direction table, the sources write their list cells, and every
// destination pixel pulls its contributions and writes the UN-NORMALISED sum once -- no fp32 atomics on the
// output (a 65-channel frame at 768x1024 is 204 M atomics for the scatter: 0.40 ms here and 0.41 ms for the
// reference's own kernel on the same B200 with a smooth flow; this path 0.29 ms: profiles/r02/forward_op.json).
// Direct index only.
// ---------------------------------------------------------------------------
namespace {
struct SplatScratch { void* scene; void* table; void* ws; size_t scene_bytes, table_bytes, ws_bytes, bytes; };
SplatScratch carve_splat_scratch(void* base, int64_t C, int64_t H, int64_t W)
{
    SplatScratch s;
    s.scene_bytes = slr_scene_core_bytes(C, 0, H, W);      // the quilted copy is for the staged gather only
    s.table_bytes = slr_clip_table_bytes(H, W, 1);
    s.ws_bytes = slr_clip_workspace_bytes(H, W, 1);
    char* p = (char*)base;
    size_t o = 0;
    s.scene = p + o; o += slr_host::align_up(s.scene_bytes);
    s.table = p + o; o += slr_host::align_up(s.table_bytes);
    s.ws = p + o;    o += slr_host::align_up(s.ws_bytes);
    s.bytes = o;
    return s;
}
}  // namespace

extern "C" size_t slr_softsplat_gather_scratch_bytes(int64_t C, int64_t H, int64_t W)
{
    if (C <= 0 || H <= 0 || W <= 0 || !slr_host::index_direct()) return 0;
    return carve_splat_scratch(nullptr, C, H, W).bytes;
}

extern "C" int slr_softsplat_sum_fwd_gather(const float* in, const float* flow, float* out,
                                            int64_t B, int64_t C, int64_t H, int64_t W,
                                            void* scratch, size_t scratch_bytes, slr_stream_t stream_)
{
    SLR_CHECK_ARGS(in && flow && out && scratch && B > 0 && C > 0 && H > 0 && W > 0 && H * W < (1ll << 27) &&
                   ((uintptr_t)scratch & 255) == 0 && slr_host::index_direct(), "slr_softsplat_sum_fwd_gather: bad arguments");
    const SplatScratch sc = carve_splat_scratch(scratch, C, H, W);
    SLR_CHECK_ARGS(sc.bytes <= scratch_bytes, "slr_softsplat_sum_fwd_gather: scratch too small (see slr_softsplat_gather_scratch_bytes)");
    cudaStream_t s = (cudaStream_t)stream_;
    const int64_t P = H * W;
    for (int64_t b = 0; b < B; ++b) {
        const float* in_b = in + b * C * P;
        const float* flow_b = flow + b * 2 * P;
        int rc = slr_host::scene_prep(in_b, nullptr, nullptr, nullptr, 0, sc.scene, C, H, W, stream_);
        if (rc) return rc;
        rc = slr_host::flow_table(flow_b, H, W, sc.table, sc.table_bytes, stream_);
        if (rc) return rc;
        rc = slr_clip_bin(sc.table, sc.table_bytes, H, W, 1, 0, 1, sc.ws, sc.ws_bytes, stream_);
        if (rc) return rc;
        GatherParams prm;
        // clip [0, 0], frame 0: alpha = 1 -> the (absent) second direction has weight 0, a static pixel weight 1
        rc = make_params(prm, sc.scene, flow_b, C, 0, H, W, 0, 0, 0, 1, 0.0f, 1.0f, out + b * C * P, nullptr, nullptr, nullptr,
                         sc.ws, sc.ws_bytes);
        if (rc) return rc;
        prm.plain_sum = 1;
        insert_kernel<<<dim3((unsigned)((P + 255) / 256), 1u), 256, 0, s>>>(prm);          // the first direction only
        launch_rowgather<2, 2>(prm, 0, s);
        const unsigned heavy_grid = std::min<unsigned>((unsigned)prm.n_tiles, 8u * (unsigned)slr_host::sm_count());
        launch_heavy<0>(prm, heavy_grid, s);
        rc = SLR_LAUNCH_STATUS();
        if (rc) return rc;
    }
    return 0;
}

extern "C" int slr_clip_frames(const void* scene, const float* motion, int64_t C, int n_tail,
                               int64_t H, int64_t W, int start, int end, int t0, int n_frames,
                               float alpha_lo, float alpha_hi,
                               float* out, float* aux, float* mask, float* nnz,
                               void* workspace, size_t workspace_bytes, slr_stream_t stream_)
{
    int rc = slr_clip_plan(motion, H, W, start, end, t0, n_frames, workspace, workspace_bytes, stream_);
    if (rc) return rc;
    rc = slr_clip_expand(scene, motion, C, n_tail, H, W, start, end, t0, n_frames, alpha_lo, alpha_hi,
                         workspace, workspace_bytes, stream_);
    if (rc) return rc;
    rc = slr_clip_gather(scene, motion, C, n_tail, H, W, start, end, t0, n_frames, alpha_lo, alpha_hi,
                         out, aux, mask, nnz, workspace, workspace_bytes, stream_);
    if (rc) return rc;
    return slr_clip_heavy(scene, motion, C, n_tail, H, W, start, end, t0, n_frames, alpha_lo, alpha_hi,
                          out, aux, mask, nnz, workspace, workspace_bytes, stream_);
}

extern "C" int slr_clip_stats_host(const void* workspace, size_t workspace_bytes, int64_t H, int64_t W, int n_frames,
                                   uint32_t stats[6], slr_stream_t stream_)
{
    SLR_CHECK_ARGS(workspace && stats && H > 0 && W > 0 && n_frames > 0 && n_frames <= kMaxFrames,
                   "slr_clip_stats_host: bad arguments");
    const Workspace ws = carve(const_cast<void*>(workspace), H, W, n_frames);
    SLR_CHECK_ARGS(ws.bytes <= workspace_bytes, "workspace too small (see slr_clip_workspace_bytes)");
    cudaStream_t s = (cudaStream_t)stream_;
    const int64_t tiles = ((W + TW - 1) / TW) * ((H + TH - 1) / TH) * n_frames;
    uint32_t n_flag = 0, n_excess = 0;
    std::vector<uint32_t> flags((size_t)tiles), fallback((size_t)tiles);
    SLR_CUDA(cudaMemcpyAsync(&n_flag, ws.flag_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    SLR_CUDA(cudaMemcpyAsync(&n_excess, ws.excess_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    SLR_CUDA(cudaMemcpyAsync(flags.data(), ws.tile_flag, sizeof(uint32_t) * (size_t)tiles, cudaMemcpyDeviceToHost, s));
    SLR_CUDA(cudaMemcpyAsync(fallback.data(), ws.fallback, sizeof(uint32_t) * (size_t)tiles, cudaMemcpyDeviceToHost, s));
    SLR_CUDA(cudaStreamSynchronize(s));
    uint32_t n_full = 0, n_fallback = 0;
    for (size_t i = 0; i < (size_t)tiles; ++i) {
        n_full += flags[i] == 2u;
        n_fallback += flags[i] != 2u && fallback[i] == 1u;
    }
    stats[0] = n_flag; stats[1] = n_full; stats[2] = n_excess; stats[3] = ws.excess_cap;
    stats[4] = n_fallback; stats[5] = (uint32_t)tiles;
    return 0;
}
