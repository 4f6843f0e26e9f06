// device_math.cuh — small f32 vector helpers, the counter-based RNG and the exactly-rounded
// primitives the parity-critical code uses.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ort {

// Individually rounded f32 operations: never contracted into FMA, whatever -fmad says.  The
// ray-triangle solve, the ray generation and the hit-point reconstruction are written with
// these so that their results are bit-identical to the reference arithmetic order
// (raytracer.odin:136-150, :580-586, :421, :456).
__device__ __forceinline__ float mulr(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float addr(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float subr(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float divr(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float sqrtr(float a) { return __fsqrt_rn(a); }

struct f3 {
    float x, y, z;
};
__device__ __forceinline__ f3 mk3(float x, float y, float z) { return {x, y, z}; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ f3 operator-(f3 a) { return {-a.x, -a.y, -a.z}; }
__device__ __forceinline__ f3 operator*(f3 a, f3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ f3 operator*(float s, f3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ f3 operator/(f3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 cross3(f3 a, f3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ float length3(f3 a) { return sqrtf(dot3(a, a)); }
__device__ __forceinline__ f3 normalize3(f3 a) { return a / length3(a); } // linalg.normalize: v / length(v)
__device__ __forceinline__ float sq(float x) { return x * x; }
__device__ __forceinline__ float norm_l1(f3 a) { return fabsf(a.x) + fabsf(a.y) + fabsf(a.z); }
// Odin builtin min/max semantics: select(a < b, a, b) / select(a > b, a, b)
__device__ __forceinline__ float omin(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float omax(float a, float b) { return a > b ? a : b; }

// Philox4x32-10.  counter = (pixel, sample_lo, sample_hi, block), key = seed; see the stream
// definition in DESIGN.md (block 0 = pixel jitter, block 1+bounce = that bounce's draws).
struct Philox4 {
    uint32_t r0, r1, r2, r3;
};
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return {c0, c1, c2, c3};
}
__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// 256-bit read-only global load (sm_100: LDG.E.256): one L1 wavefront per lane instead of two.
// `p` must be 32-byte aligned.
struct F8 {
    float4 lo, hi;
};
__device__ __forceinline__ F8 ldg8(const void* p) {
    F8 r;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z),
                   "=f"(r.hi.w)
                 : "l"(p));
    return r;
}

} // namespace ort
