// Microbenchmark: bytes per clock per SM that L1-hit LDG.128 / LDG.32 and LDS.128 deliver,
// aligned and misaligned by 16 bytes.  Build: nvcc -arch=sm_100a -O3 -o l1_bw l1_bw.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(const float4* __restrict__ g, float4* out, int iters, int shift)
{
    extern __shared__ float4 sm[];
    const int tid = threadIdx.x;
    // per-CTA window of 1024 float4 (16 KB) -> L1 resident
    const float4* base = g + (size_t)blockIdx.x * 1024;
    if (MODE == 2) { for (int i = tid; i < 1024 + 8; i += 256) sm[i] = base[i % 1024]; __syncthreads(); }
    float4 acc = make_float4(0, 0, 0, 0);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        #pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int idx = ((tid + j * 256 + it * 32) & 1023) + shift;
            float4 v;
            if (MODE == 0) v = __ldg(base + idx);
            else if (MODE == 1) { const float* p = (const float*)(base) + idx; float x = __ldg(p); v = make_float4(x, x, x, x); }
            else v = sm[idx];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    long long t1 = clock64();
    if (acc.x == 123.456f) out[0] = acc;
    if (tid == 0 && blockIdx.x == 0) ((long long*)out)[8] = t1 - t0;
}

int main()
{
    float4 *g, *out;
    const int ctas = 148 * 2;
    cudaMalloc(&g, sizeof(float4) * 1040 * (size_t)ctas + 4096);
    cudaMemset(g, 0, sizeof(float4) * 1040 * (size_t)ctas + 4096);
    cudaMalloc(&out, 4096);
    const int iters = 2000;
    for (int mode = 0; mode < 3; ++mode)
        for (int shift = 0; shift < 2; ++shift) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<ctas, 256, 0>>>(g, out, iters, shift);
                if (mode == 1) k<1><<<ctas, 256, 0>>>(g, out, iters, shift);
                if (mode == 2) k<2><<<ctas, 256, 17000>>>(g, out, iters, shift);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long cyc; cudaMemcpy(&cyc, (long long*)out + 8, 8, cudaMemcpyDeviceToHost);
            const double bytes_per_cta = (double)iters * 4 * 256 * (mode == 1 ? 4 : 16);
            printf("mode %s shift %d: %.3f ms, CTA0 cycles %lld, bytes/clk/SM (2 CTAs) = %.1f\n",
                   mode == 0 ? "LDG.128" : mode == 1 ? "LDG.32 " : "LDS.128", shift, ms, cyc, 2.0 * bytes_per_cta / cyc);
        }
    return 0;
}
