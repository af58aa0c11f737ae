// LDG.128 vs LDG.256 (sm_100: ld.global.v8.f32): bytes per clock per SM a gather-like access pattern
// gets from L1 (window resident in L1) and from L2 (window far larger than L1), for runs of 32
// consecutive pixels starting at an arbitrary (misaligned) pixel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldg_width ldg_width.cu && ./ldg_width
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct f8 { float v[8]; };

__device__ __forceinline__ f8 ld256(const void* p)
{
    f8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
    return r;
}

// MODE 0: 16 B per pixel, LDG.128; MODE 1: 32 B per pixel, LDG.256; MODE 2: 32 B per pixel, 2 x LDG.128
template <int MODE>
__global__ void __launch_bounds__(128) k(const char* __restrict__ g, float* out, long long* cyc, int iters, int window_px, int n_slots)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int px_bytes = MODE == 0 ? 16 : 32;
    const char* base = g + (size_t)blockIdx.x * (size_t)window_px * px_bytes;
    float acc = 0.f;
    unsigned rng = blockIdx.x * 977u + warp * 131u + 7u;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        // n_slots independent loads of a run of 32 consecutive pixels at a pseudo-random start
        #pragma unroll 4
        for (int s = 0; s < n_slots; ++s) {
            rng = rng * 1664525u + 1013904223u;
            const unsigned start = (rng >> 8) % (unsigned)(window_px - 32);
            const char* p = base + (size_t)(start + lane) * px_bytes;
            if (MODE == 0) { const float4 v = __ldg((const float4*)p); acc += v.x + v.y + v.z + v.w; }
            else if (MODE == 1) { const f8 v = ld256(p); acc += v.v[0] + v.v[1] + v.v[2] + v.v[3] + v.v[4] + v.v[5] + v.v[6] + v.v[7]; }
            else { const float4 a = __ldg((const float4*)p), b = __ldg((const float4*)(p + 16)); acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w; }
        }
    }
    long long t1 = clock64();
    if (acc == 123.456f) out[0] = acc;
    if (tid == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main()
{
    const size_t bytes = (size_t)1 << 30;
    char* g; float* out; long long* cyc;
    CK(cudaMalloc(&g, bytes)); CK(cudaMemset(g, 0, bytes)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&cyc, 64));
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const char* names[] = {"16 B/px LDG.128", "32 B/px LDG.256", "32 B/px 2xLDG.128"};
    const int ctas = sms * 4;                    // 4 CTAs x 4 warps per SM, like rowgather_kernel
    for (int window_kb : {24, 1024}) {           // per-CTA window: 4 x 24 KB fits L1; 1 MB per CTA = 592 MB in total: L2 / HBM
        for (int mode = 0; mode < 3; ++mode) {
            const int px_bytes = mode == 0 ? 16 : 32;
            const int window_px = window_kb * 1024 / px_bytes;
            const int iters = 400, n_slots = 12;
            if ((size_t)ctas * window_px * px_bytes > bytes) { printf("window too large\n"); continue; }
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(e0));
                if (mode == 0) k<0><<<ctas, 128>>>(g, out, cyc, iters, window_px, n_slots);
                if (mode == 1) k<1><<<ctas, 128>>>(g, out, cyc, iters, window_px, n_slots);
                if (mode == 2) k<2><<<ctas, 128>>>(g, out, cyc, iters, window_px, n_slots);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            }
            CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            const double total = (double)ctas * 4 * iters * n_slots * 32 * px_bytes;
            printf("window %4d KB/CTA  %-18s : %7.3f ms  %8.1f GB/s  %6.1f B/clk/SM (at 1.9 GHz)\n", window_kb, names[mode], ms,
                   total / ms / 1e6, total / sms / (ms * 1e-3) / 1.9e9);
        }
    }
    return 0;
}
