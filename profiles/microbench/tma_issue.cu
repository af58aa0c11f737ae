// How fast can cp.async.bulk (UBLKCP) row copies be ISSUED?  tma_stage.cu measured ~76 cycles per copy
// from one divergent thread (ptxas wraps every UBLKCP in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop
// because the operands are not provably warp-uniform).  Variants:
//   0  `if (tid == producer)`            divergent single thread (as in tma_stage.cu)
//   1  warp-uniform branch + elect.sync  operands from a shared-memory descriptor table (the real case)
//   2  all 32 lanes of the producer warp issue one row each (descriptor per lane)
//   3  like 1 but 4 warps issue a quarter of the rows each
// Descriptors (src offset, dst offset, bytes) are read from shared memory like the gather's plan.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_issue tma_issue.cu && ./tma_issue
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}

constexpr int kRows = 64;          // rows per stage
struct Desc { uint32_t src_off, dst_off, bytes, pad; };

template <int MODE>
__global__ void __launch_bounds__(160)
issue_kernel(const char* __restrict__ g, size_t span, int row_bytes, int iters, long long* cycles)
{
    extern __shared__ __align__(128) char smem[];
    __shared__ Desc desc[kRows];
    __shared__ uint64_t full;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < kRows) {
        desc[tid].src_off = (uint32_t)(tid * 16384 + (tid & 3) * 16);
        desc[tid].dst_off = (uint32_t)(tid * row_bytes);
        desc[tid].bytes = (uint32_t)row_bytes;
    }
    if (tid == 0) { mbar_init(&full, MODE == 3 ? 4 : 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const uint32_t bar = smem_u32(&full), base = smem_u32(smem);
    const char* src0 = g + ((size_t)blockIdx.x * 2097152) % (span - 2097152);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
            if (tid == 128) {
                mbar_expect_tx(&full, kRows * row_bytes);
                for (int r = 0; r < kRows; ++r) bulk_g2s(base + desc[r].dst_off, src0 + desc[r].src_off, desc[r].bytes, bar);
            }
        } else if (MODE == 1) {
            if (warp == 4) {
                if (elect_one()) {
                    mbar_expect_tx(&full, kRows * row_bytes);
                    #pragma unroll 4
                    for (int r = 0; r < kRows; ++r) bulk_g2s(base + desc[r].dst_off, src0 + desc[r].src_off, desc[r].bytes, bar);
                }
            }
        } else if (MODE == 2) {
            if (warp == 4) {
                if (lane == 0) mbar_expect_tx(&full, kRows * row_bytes);
                __syncwarp();
                for (int r = lane; r < kRows; r += 32) bulk_g2s(base + desc[r].dst_off, src0 + desc[r].src_off, desc[r].bytes, bar);
            }
        } else {
            if (warp < 4) {
                if (elect_one()) {
                    mbar_expect_tx(&full, (kRows / 4) * row_bytes);
                    #pragma unroll 4
                    for (int r = warp * (kRows / 4); r < (warp + 1) * (kRows / 4); ++r)
                        bulk_g2s(base + desc[r].dst_off, src0 + desc[r].src_off, desc[r].bytes, bar);
                }
            }
        }
        mbar_wait(&full, it & 1);          // everybody waits for the stage (single stage: issue rate + latency)
        __syncthreads();
    }
    long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

int main()
{
    const size_t span = (size_t)256 << 20;
    char* g; long long* cyc;
    CK(cudaMalloc(&g, span)); CK(cudaMemset(g, 0, span)); CK(cudaMalloc(&cyc, 64));
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const char* names[] = {"tid==X (divergent)", "warp branch + elect.sync", "32 lanes, a row each", "4 warps x elect.sync"};
    for (int rb : {256, 1024})
        for (int ctas = 1; ctas <= 2; ++ctas)
            for (int mode = 0; mode < 4; ++mode) {
                const int iters = 300, smem = kRows * rb, grid = sms * ctas;
                auto launch = [&]() {
                    if (mode == 0) { CK(cudaFuncSetAttribute(issue_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); issue_kernel<0><<<grid, 160, smem>>>(g, span, rb, iters, cyc); }
                    if (mode == 1) { CK(cudaFuncSetAttribute(issue_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); issue_kernel<1><<<grid, 160, smem>>>(g, span, rb, iters, cyc); }
                    if (mode == 2) { CK(cudaFuncSetAttribute(issue_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); issue_kernel<2><<<grid, 160, smem>>>(g, span, rb, iters, cyc); }
                    if (mode == 3) { CK(cudaFuncSetAttribute(issue_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); issue_kernel<3><<<grid, 160, smem>>>(g, span, rb, iters, cyc); }
                };
                launch(); CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
                printf("row %4d B, %d CTA/SM, %-26s: %7.3f ms, %6.1f cycles per stage of %d rows (%.1f cycles/copy incl. wait), %.0f GB/s\n",
                       rb, ctas, names[mode], ms, (double)c / iters, kRows, (double)c / iters / kRows,
                       (double)grid * iters * kRows * rb / ms / 1e6);
            }
    return 0;
}
