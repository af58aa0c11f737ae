// Shared device helpers for the SLR-SFS splat kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace slr {

// Bilinear landing footprint of one source pixel: the 2x2 destination cells it
// splats into and their weights.  Arithmetic follows the reference scatter
// (/root/reference/models/softsplat.py:166-200) operation for operation so the
// fp32 weights are bit-identical: landing = pixel + flow, north-west = floor,
// weight of a corner = product of the distances to the opposite corner.
struct Footprint {
    int x0, y0;        // north-west cell
    float w[4];        // NW, NE, SW, SE
    unsigned ok;       // bit k set when corner k lies inside the frame
};

// Footprint of a landing position (ox, oy) given directly.
__device__ __forceinline__ Footprint footprint_at(float ox, float oy, int H, int W)
{
    Footprint f;
    const float flx = floorf(ox), fly = floorf(oy);
    // Anything this far out (or non-finite) misses the frame with all four
    // corners; testing in the float domain keeps the int conversion defined.
    const bool far = !(flx > -2.0f && flx < (float)W + 1.0f && fly > -2.0f && fly < (float)H + 1.0f);
    const int ix = far ? -4 : (int)flx, iy = far ? -4 : (int)fly;
    const float bx = (float)ix, by = (float)iy, ex = (float)(ix + 1), ey = (float)(iy + 1);
    f.x0 = ix; f.y0 = iy;
    f.w[0] = (ex - ox) * (ey - oy);
    f.w[1] = (ox - bx) * (ey - oy);
    f.w[2] = (ex - ox) * (oy - by);
    f.w[3] = (ox - bx) * (oy - by);
    const bool xl = ix >= 0 && ix < W, xr = ix + 1 >= 0 && ix + 1 < W;
    const bool yt = iy >= 0 && iy < H, yb = iy + 1 >= 0 && iy + 1 < H;
    f.ok = far ? 0u : ((xl && yt) ? 1u : 0u) | ((xr && yt) ? 2u : 0u) | ((xl && yb) ? 4u : 0u) | ((xr && yb) ? 8u : 0u);
    return f;
}

__device__ __forceinline__ Footprint landing(int x, int y, float fx, float fy, int H, int W)
{
    return footprint_at((float)x + fx, (float)y + fy, H, W);
}

// fp32 add with no returned value: compiles to RED.E.ADD.F32 (fire-and-forget at L2).
__device__ __forceinline__ void red_add(float* p, float v) { atomicAdd(p, v); }

// Order-independent float max on global memory.  Works for mixed signs because
// non-negative floats order like signed ints and negative floats order inversely
// like unsigned ints.
__device__ __forceinline__ void red_max(float* p, float v)
{
    // keyed on the sign BIT: -0.0 (a zero-weight corner of a negative value) must
    // still beat every negative number, as fmaxf does in the reference's CAS loop
    if (!signbit(v)) atomicMax(reinterpret_cast<int*>(p), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned*>(p), __float_as_uint(v));
}

__device__ __forceinline__ float warp_max(float v)
{
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace slr
