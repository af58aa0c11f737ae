// TEST INFRASTRUCTURE -- not part of the product.
//
// A stand-in for <cuda_runtime.h> that gives the CUDA sources of slr-sfs_b200/csrc a CPU
// meaning, so that the *same kernel text* the B200 runs can be executed (slowly) by the
// `-m "not gpu"` tests in a container without a GPU: every CUDA thread of a block is a
// cooperative fiber (tests/emu/emu_runtime.cpp), __syncthreads / warp collectives are
// scheduling points, global and shared atomics are plain read-modify-writes (one OS thread).
// Only tests/ builds or loads this; the package loads csrc/libslr_splat.so and nothing else.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>

#define SLR_CPU_EMULATION 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static

// ---------------------------------------------------------------------------
// vector types and launch geometry
// ---------------------------------------------------------------------------
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(8) uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
inline float4 make_float4(float x, float y, float z, float w) { float4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
inline float2 make_float2(float x, float y) { float2 v; v.x = x; v.y = y; return v; }
inline uint2 make_uint2(unsigned x, unsigned y) { uint2 v; v.x = x; v.y = y; return v; }

extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

namespace emu {
enum Op { kShflIdx, kShflXor, kShflUp, kShflDown, kMatchAny, kAll, kAny, kBallot, kReduceMax, kReduceMin,
          kReduceAdd, kSyncWarp };
// Runs `body` once per CUDA thread of a grid x block launch (blocks one after the other).
void launch(dim3 grid, dim3 block, const std::function<void()>& body);
uint64_t collective(Op op, unsigned mask, uint64_t value, int param);
int block_barrier(int pred, int mode);      // mode 0: plain, 1: or, 2: and, 3: count
void spin();                                // a polling loop's scheduling point: lets every other fiber run once
void misaligned(const void* p, size_t a);
template <class T> inline void check(const T* p)
{
    if ((uintptr_t)p % alignof(T)) misaligned(p, alignof(T));
}
template <class T> inline uint64_t bits(T v) { uint64_t u = 0; memcpy(&u, &v, sizeof(T)); return u; }
template <class T> inline T unbits(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }
}  // namespace emu

// ---------------------------------------------------------------------------
// memory intrinsics
// ---------------------------------------------------------------------------
template <class T> inline T __ldg(const T* p) { emu::check(p); return *p; }
template <class T> inline T __ldcs(const T* p) { emu::check(p); return *p; }
template <class T> inline T __ldcg(const T* p) { emu::check(p); return *p; }
template <class T> inline T __ldca(const T* p) { emu::check(p); return *p; }
template <class T> inline void __stcs(T* p, T v) { emu::check(p); *p = v; }
template <class T> inline void __stcg(T* p, T v) { emu::check(p); *p = v; }
template <class T> inline void __stwt(T* p, T v) { emu::check(p); *p = v; }

template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
template <class T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <class T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }

// ---------------------------------------------------------------------------
// arithmetic intrinsics (build with -ffp-contract=off: a * b + c stays two roundings)
// ---------------------------------------------------------------------------
inline unsigned __float_as_uint(float f) { return emu::unbits<unsigned>(emu::bits(f)); }
inline int __float_as_int(float f) { return emu::unbits<int>(emu::bits(f)); }
inline float __uint_as_float(unsigned u) { return emu::unbits<float>(emu::bits(u)); }
inline float __int_as_float(int u) { return emu::unbits<float>(emu::bits(u)); }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline long min(long a, long b) { return a < b ? a : b; }
inline long max(long a, long b) { return a > b ? a : b; }

// ---------------------------------------------------------------------------
// barriers and warp collectives
// ---------------------------------------------------------------------------
inline void __syncthreads() { emu::block_barrier(0, 0); }
inline int __syncthreads_or(int p) { return emu::block_barrier(p, 1); }
inline int __syncthreads_and(int p) { return emu::block_barrier(p, 2); }
inline int __syncthreads_count(int p) { return emu::block_barrier(p, 3); }
inline void __syncwarp(unsigned m = 0xffffffffu) { emu::collective(emu::kSyncWarp, m, 0, 0); }
template <class T> inline T __shfl_sync(unsigned m, T v, int lane, int = 32)
{ return emu::unbits<T>(emu::collective(emu::kShflIdx, m, emu::bits(v), lane)); }
template <class T> inline T __shfl_xor_sync(unsigned m, T v, int x, int = 32)
{ return emu::unbits<T>(emu::collective(emu::kShflXor, m, emu::bits(v), x)); }
template <class T> inline T __shfl_up_sync(unsigned m, T v, unsigned d, int = 32)
{ return emu::unbits<T>(emu::collective(emu::kShflUp, m, emu::bits(v), (int)d)); }
template <class T> inline T __shfl_down_sync(unsigned m, T v, unsigned d, int = 32)
{ return emu::unbits<T>(emu::collective(emu::kShflDown, m, emu::bits(v), (int)d)); }
template <class T> inline unsigned __match_any_sync(unsigned m, T v)
{ return (unsigned)emu::collective(emu::kMatchAny, m, emu::bits(v), 0); }
inline int __all_sync(unsigned m, int p) { return (int)emu::collective(emu::kAll, m, p != 0, 0); }
inline int __any_sync(unsigned m, int p) { return (int)emu::collective(emu::kAny, m, p != 0, 0); }
inline unsigned __ballot_sync(unsigned m, int p) { return (unsigned)emu::collective(emu::kBallot, m, p != 0, 0); }
inline int __reduce_max_sync(unsigned m, int v) { return (int)(int64_t)emu::collective(emu::kReduceMax, m, (uint64_t)(int64_t)v, 1); }
inline unsigned __reduce_max_sync(unsigned m, unsigned v) { return (unsigned)emu::collective(emu::kReduceMax, m, v, 0); }
inline int __reduce_min_sync(unsigned m, int v) { return (int)(int64_t)emu::collective(emu::kReduceMin, m, (uint64_t)(int64_t)v, 1); }
inline unsigned __reduce_min_sync(unsigned m, unsigned v) { return (unsigned)emu::collective(emu::kReduceMin, m, v, 0); }
inline int __reduce_add_sync(unsigned m, int v) { return (int)emu::collective(emu::kReduceAdd, m, (uint64_t)(int64_t)v, 0); }
inline unsigned __reduce_add_sync(unsigned m, unsigned v) { return (unsigned)emu::collective(emu::kReduceAdd, m, v, 0); }

// ---------------------------------------------------------------------------
// the sliver of the runtime API the C-ABI layer touches
// ---------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
inline cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 2; return cudaSuccess; }   // "2 SMs": small fixed grids
