// Microbenchmarks behind the TMA-staged gather.  Built here, run on the B200 box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stage tma_stage.cu && ./tma_stage
//
//  B  cp.async.bulk (UBLKCP) row copies global -> shared: aggregate GB/s and copies/s as a function of
//     the copy size (the staged source regions of a destination tile are rows of 10-60 pixels x 16 B),
//     with consumer warps reading the stage back with LDS.128.
//  C  cp.async.bulk.tensor.2d (UTMALDG) boxes of a [rows][width x float4] plane, same pipeline.
//  D  per-SM bytes/clk of the consumer side alone: LDS.128, generic LD.128 on a shared address,
//     LDG.128 on an L1-resident window at a misaligned run start, and float4 exchange by 4 x SHFL.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tensor2d_g2s(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

constexpr int kStages = 3;
constexpr int kConsumers = 128;      // 4 consumer warps + 1 producer warp

// MODE 0: rows of `row_bytes` by cp.async.bulk; MODE 1: one 2-D tensor box per stage
template <int MODE>
__global__ void __launch_bounds__(kConsumers + 32)
stage_kernel(const char* __restrict__ g, size_t span_bytes, int rows, int row_bytes, int iters, int lds_per_lane,
             const __grid_constant__ CUtensorMap tmap, int box_w, int box_h, int plane_w, int plane_h, float* sink)
{
    extern __shared__ __align__(128) char smem[];
    __shared__ uint64_t full[kStages], empty[kStages];
    const int stage_bytes = MODE == 0 ? rows * row_bytes : box_w * box_h * 16;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kConsumers / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid >= kConsumers) {
        if (tid == kConsumers) {          // the producer thread
            uint32_t rng = blockIdx.x * 2654435761u + 12345u;
            for (int it = 0; it < iters; ++it) {
                const int s = it % kStages;
                if (it >= kStages) mbar_wait(&empty[s], ((it / kStages) - 1) & 1);
                mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
                char* dst = smem + (size_t)s * stage_bytes;
                if (MODE == 0) {
                    // a region of `rows` consecutive image rows (pitch 16 KB) at a pseudo-random place
                    rng = rng * 1664525u + 1013904223u;
                    size_t base = ((size_t)(rng >> 4) * 16) % (span_bytes - (size_t)rows * 16384 - 65536);
                    base &= ~(size_t)15;
                    for (int r = 0; r < rows; ++r)
                        bulk_g2s(dst + (size_t)r * row_bytes, g + base + (size_t)r * 16384 + (size_t)(r & 3) * 16, (uint32_t)row_bytes, &full[s]);
                } else {
                    rng = rng * 1664525u + 1013904223u;
                    const int x = (int)((rng >> 8) % (unsigned)(plane_w - box_w)), y = (int)((rng >> 3) % (unsigned)(plane_h - box_h));
                    tensor2d_g2s(dst, &tmap, x * 4, y, &full[s]);
                }
            }
        }
        return;
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int n16 = stage_bytes / 16;
    for (int it = 0; it < iters; ++it) {
        const int s = it % kStages;
        mbar_wait(&full[s], (it / kStages) & 1);
        const float4* st = reinterpret_cast<const float4*>(smem + (size_t)s * stage_bytes);
        #pragma unroll 4
        for (int k = 0; k < lds_per_lane; ++k) {
            const float4 v = st[(tid + k * 37) % n16];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[s]);
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x;
}

// D: consumer-side instruction throughput on one resident window
template <int MODE>
__global__ void __launch_bounds__(256)
pipe_kernel(const float4* __restrict__ g, float* sink, long long* cycles, int iters, int shift)
{
    extern __shared__ __align__(16) float4 sm[];
    const int tid = threadIdx.x;
    const float4* base = g + (size_t)blockIdx.x * 2048;
    for (int i = tid; i < 2048 + 64; i += 256) sm[i] = base[i % 2048];
    __syncthreads();
    const float4* gen = sm;        // generic pointer to shared memory
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        #pragma unroll
        for (int j = 0; j < 8; ++j) {
            // a run of 32 consecutive float4 starting at a (mis)aligned place
            const int idx = (((it * 8 + j) * 97) & 1023) * 1 + shift + (tid & 31) + (tid >> 5) * 7;
            float4 v;
            if (MODE == 0) v = sm[idx & 2047];
            else if (MODE == 1) { asm volatile("ld.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(gen + (idx & 2047))); }
            else if (MODE == 2) v = __ldg(base + (idx & 2047));
            else {
                v = acc;
                v.x = __shfl_up_sync(0xffffffffu, acc.x + j, 1); v.y = __shfl_up_sync(0xffffffffu, acc.y, 1);
                v.z = __shfl_up_sync(0xffffffffu, acc.z, 1); v.w = __shfl_up_sync(0xffffffffu, acc.w, 1);
            }
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    const long long t1 = clock64();
    if (acc.x == 123.456f) sink[0] = acc.x;
    if (tid == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    const size_t span = (size_t)768 << 20;            // 768 MB of source: far larger than L2
    char* g; float* sink; long long* cyc;
    CK(cudaMalloc(&g, span)); CK(cudaMemset(g, 0, span));
    CK(cudaMalloc(&sink, 256)); CK(cudaMalloc(&cyc, 256));
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

    EncodeTiled encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    const int plane_w = 1024, plane_h = 768;
    CUtensorMap dummy; memset(&dummy, 0, sizeof(dummy));

    printf("== B: cp.async.bulk rows (pitch 16 KB), %d SMs, %d stages, 4 consumer warps ==\n", sms, kStages);
    const int row_sizes[] = {128, 256, 512, 1024, 2048, 4096};
    for (int ctas_per_sm = 1; ctas_per_sm <= 4; ctas_per_sm *= 2)
        for (int rs : row_sizes)
            for (int lds = 0; lds <= 48; lds += 48) {
                const int stage_target = 16384;
                const int rows = stage_target / rs;
                const int stage_bytes = rows * rs;
                const int smem = kStages * stage_bytes;
                const int iters = 400;
                CK(cudaFuncSetAttribute(stage_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                const int grid = sms * ctas_per_sm;
                for (int rep = 0; rep < 2; ++rep) {
                    CK(cudaEventRecord(e0));
                    stage_kernel<0><<<grid, kConsumers + 32, smem>>>(g, span, rows, rs, iters, lds, dummy, 0, 0, plane_w, plane_h, sink);
                    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                }
                CK(cudaGetLastError());
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                const double bytes = (double)grid * iters * stage_bytes;
                printf("ctas/SM %d row %4d B x %3d rows/stage  lds/lane/stage %2d : %7.3f ms  %7.1f GB/s  %6.2f Mcopies/s/SM  (%.1f B/clk/SM @1.9GHz)\n",
                       ctas_per_sm, rs, rows, lds, ms, bytes / ms / 1e6, (double)grid * iters * rows / ms / 1e3 / sms,
                       bytes / sms / (ms * 1e-3) / 1.9e9);
            }

    printf("== C: cp.async.bulk.tensor.2d boxes of a %dx%d float4 plane ==\n", plane_h, plane_w);
    const int boxes[][2] = {{32, 8}, {40, 12}, {48, 16}, {64, 16}, {32, 32}};
    for (auto& b : boxes) {
        CUtensorMap tmap;
        cuuint64_t dims[2] = {(cuuint64_t)plane_w * 4, (cuuint64_t)plane_h * 64};     // 64 planes stacked: 805 MB? no: keep within span
        dims[1] = span / ((size_t)plane_w * 16);
        cuuint64_t strides[1] = {(cuuint64_t)plane_w * 16};
        cuuint32_t box[2] = {(cuuint32_t)b[0] * 4, (cuuint32_t)b[1]};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d for box %dx%d\n", (int)r, b[0], b[1]); continue; }
        for (int ctas_per_sm = 1; ctas_per_sm <= 4; ctas_per_sm *= 2)
            for (int lds = 0; lds <= 48; lds += 48) {
                const int stage_bytes = b[0] * b[1] * 16;
                const int smem = kStages * stage_bytes;
                const int iters = 400;
                CK(cudaFuncSetAttribute(stage_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                const int grid = sms * ctas_per_sm;
                for (int rep = 0; rep < 2; ++rep) {
                    CK(cudaEventRecord(e0));
                    stage_kernel<1><<<grid, kConsumers + 32, smem>>>(g, span, 0, 0, iters, lds, tmap, b[0], b[1], plane_w, (int)dims[1], sink);
                    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                }
                CK(cudaGetLastError());
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                const double bytes = (double)grid * iters * stage_bytes;
                printf("box %2dx%2d px (%5d B) ctas/SM %d lds/lane/stage %2d : %7.3f ms  %7.1f GB/s  %6.2f Mboxes/s/SM\n",
                       b[0], b[1], stage_bytes, ctas_per_sm, lds, ms, bytes / ms / 1e6, (double)grid * iters / ms / 1e3 / sms);
            }
    }

    printf("== D: consumer-side pipes, 2 CTAs x 256 threads per SM, bytes/clk/SM ==\n");
    const char* names[] = {"LDS.128", "LD.128 generic->shared", "LDG.128 (L1-resident window)", "4 x SHFL.UP (float4)"};
    for (int mode = 0; mode < 4; ++mode)
        for (int shift = 0; shift < 4; shift += (mode == 3 ? 4 : 1)) {
            const int iters = 2000, grid = sms * 2, smem = (2048 + 64) * 16;
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) pipe_kernel<0><<<grid, 256, smem>>>((const float4*)g, sink, cyc, iters, shift);
                if (mode == 1) pipe_kernel<1><<<grid, 256, smem>>>((const float4*)g, sink, cyc, iters, shift);
                if (mode == 2) pipe_kernel<2><<<grid, 256, smem>>>((const float4*)g, sink, cyc, iters, shift);
                if (mode == 3) pipe_kernel<3><<<grid, 256, smem>>>((const float4*)g, sink, cyc, iters, shift);
                CK(cudaDeviceSynchronize());
            }
            long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
            printf("%-30s run start %% 8 = %d : %.1f B/clk/SM\n", names[mode], shift, 2.0 * iters * 8 * 256 * 16 / (double)c);
        }
    return 0;
}
