// Bulk asynchronous copies global -> shared memory through the TMA unit (sm_90+; SASS: UBLKCP.S.G)
// and the transaction barriers (mbarrier; SASS: SYNCS.*) that signal their completion.
//
// Thin wrappers over the PTX so that the kernels read as what they do.  Under SLR_CPU_EMULATION
// (tests/emu: the CUDA sources compiled for the CPU) the same names copy synchronously and keep the
// barrier's arithmetic (pending arrivals, pending bytes, phase bit), so the staged gather's index
// logic and phase bookkeeping are executed by the CPU test suite; what the emulation cannot show is
// asynchrony itself (a copy that is still in flight when shared memory is read).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace slr {
namespace tma {

#if defined(SLR_CPU_EMULATION)

// barrier word: [63] phase | [48..62] arrival count per phase | [32..47] pending arrivals | [0..31] pending bytes (signed)
typedef uint64_t Barrier;

inline void emu_settle(Barrier* b)
{
    const uint64_t v = *b;
    const unsigned pending = (unsigned)(v >> 32) & 0xffffu;
    const int32_t tx = (int32_t)(uint32_t)v;
    if (pending == 0 && tx == 0) {
        const uint64_t count = (v >> 48) & 0x7fffu;
        *b = ((v >> 63) ^ 1ull) << 63 | count << 48 | count << 32;
    }
}
inline void init(Barrier* b, unsigned count) { *b = (uint64_t)count << 48 | (uint64_t)count << 32; }
inline void fence_init() {}
inline void emu_update(Barrier* b, int d_pending, int32_t d_tx)
{
    const uint64_t v = *b;
    const unsigned pending = ((unsigned)(v >> 32) & 0xffffu) - (unsigned)d_pending;
    const int32_t tx = (int32_t)(uint32_t)v + d_tx;
    *b = (v & 0xffff000000000000ull) | (uint64_t)(pending & 0xffffu) << 32 | (uint32_t)tx;
    emu_settle(b);
}
inline void arrive(Barrier* b) { emu_update(b, 1, 0); }
inline void arrive_expect_tx(Barrier* b, unsigned bytes) { emu_update(b, 1, (int32_t)bytes); }
inline void wait(Barrier* b, unsigned parity)
{
    unsigned long spins = 0;
    while ((unsigned)(*b >> 63) == (parity & 1u)) {
        if (++spins > 1000000ul) { fprintf(stderr, "cuda-emu: an mbarrier wait never completes\n"); abort(); }
        emu::spin();
    }
}
inline void load(void* dst_shared, const void* src_global, unsigned bytes, Barrier* b)
{
    if (((uintptr_t)dst_shared | (uintptr_t)src_global | bytes) & 15u) emu::misaligned(src_global, 16);
    memcpy(dst_shared, src_global, bytes);
    emu_update(b, 0, -(int32_t)bytes);
}
inline bool elect_one() { __syncwarp(); return (threadIdx.x & 31u) == 0u; }
// shared-memory addresses in the form the copy instruction takes them (computed once, then offset)
typedef uintptr_t SharedAddr;
inline SharedAddr shared_addr(const void* p) { return (SharedAddr)p; }
inline void load_at(SharedAddr dst, const void* src_global, unsigned bytes, SharedAddr barrier)
{
    load((void*)dst, src_global, bytes, (Barrier*)barrier);
}
inline void arrive_expect_tx_at(SharedAddr barrier, unsigned bytes) { arrive_expect_tx((Barrier*)barrier, bytes); }

#else

typedef uint64_t Barrier;
// shared-memory addresses in the form the copy instruction takes them (computed once, then offset)
typedef uint32_t SharedAddr;

__device__ __forceinline__ SharedAddr shared_addr(const void* p) { return (SharedAddr)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void init(Barrier* b, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(shared_addr(b)), "r"(count) : "memory");
}
// after the init, before any thread or the TMA unit uses the barrier (followed by a __syncthreads)
__device__ __forceinline__ void fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void arrive(Barrier* b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(shared_addr(b)) : "memory");
}
// one arrival that also announces `bytes` of copies which will complete on this barrier
__device__ __forceinline__ void arrive_expect_tx(Barrier* b, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(shared_addr(b)), "r"(bytes) : "memory");
}
// blocks until the phase with parity `parity` has completed (all arrivals in, all announced bytes landed)
__device__ __forceinline__ void wait(Barrier* b, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" :: "r"(shared_addr(b)), "r"(parity) : "memory");
}
// `bytes` (a multiple of 16) from 16-byte aligned global memory to 16-byte aligned shared memory;
// completion is counted on `b` (complete_tx)
__device__ __forceinline__ void load(void* dst_shared, const void* src_global, unsigned bytes, Barrier* b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(shared_addr(dst_shared)), "l"(src_global), "r"(bytes), "r"(shared_addr(b)) : "memory");
}
__device__ __forceinline__ void load_at(SharedAddr dst, const void* src_global, unsigned bytes, SharedAddr barrier)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src_global), "r"(bytes), "r"(barrier) : "memory");
}
__device__ __forceinline__ void arrive_expect_tx_at(SharedAddr barrier, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(barrier), "r"(bytes) : "memory");
}
// true in exactly one lane of a converged warp; ptxas then keeps the operands of the copies issued
// under it in uniform registers (no per-lane serialisation loop around UBLKCP)
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}

#endif

}  // namespace tma
}  // namespace slr

// dynamic shared memory of a kernel (the emulation gives every kernel the same static arena)
#if defined(SLR_CPU_EMULATION)
#define SLR_DYNAMIC_SMEM(name) alignas(1024) static unsigned char name[232448]
#else
#define SLR_DYNAMIC_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
#endif
