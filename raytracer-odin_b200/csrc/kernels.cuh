// kernels.cuh — the wavefront path-tracing kernels (sm_100a).
//
//   k_raygen      render_task's inner loop head (raytracer.odin:580-586): jittered pinhole rays
//   k_trace       (traverse.cuh) cast_ray (raytracer.odin:416-430) on the 4-wide re-emission of the
//                 reference BVH, and surface_sampling_pdf_bvh_sum (shading.odin:62-94) on the light
//                 BVH: persistent warps, per-lane dynamic fetch
//   k_shade       raytrace's body (raytracer.odin:437-500) + sample/pdf/shade (shading.odin),
//                 texture fetches through texture objects, ballot/prefix-sum queue compaction
//   k_resolve     rc_set_pixel (main.odin:89-102): per-pixel accumulation in sample order
//   k_stats, k_pack_rays, k_unpack_hits, k_pack_stats, k_tonemap: small helpers
#pragma once
#include "device_math.cuh"
#include "wide_bvh.h"

namespace ort {

constexpr float RAY_EPS = 1e-3f; // raytracer.odin:418, shading.odin:66
constexpr float PI_F = 3.14159265358979323846264338327950288f;
constexpr float TAU_F = 6.28318530717958647692528676655900576f;

#ifndef ORT_SMEM_STACK
#define ORT_SMEM_STACK 16
#endif
// k_trace<light> is compiled for 8 resident CTAs per SM (63 registers, no spills; the closest-hit kernel stays at
// its own 69 / 7 CTAs, where 8 was slower): 7 -> 8 CTAs is -3.4 % light-pass time, 9 and 10 lose again
// (profiles/r2s_light_pass.md)
#ifndef ORT_LIGHT_MIN_CTAS
#define ORT_LIGHT_MIN_CTAS 8
#endif
#ifndef ORT_TRACE_MIN_CTAS
#define ORT_TRACE_MIN_CTAS 1
#endif
#ifndef ORT_TRACE_THREADS
#define ORT_TRACE_THREADS 128
#endif
constexpr int TRACE_THREADS = ORT_TRACE_THREADS;
constexpr int SMEM_STACK = ORT_SMEM_STACK;        // stack entries per thread kept in shared memory
constexpr int LOCAL_STACK = 128 - ORT_SMEM_STACK; // overflow entries per thread in local memory
constexpr int MAX_STACK = SMEM_STACK + LOCAL_STACK;

struct DevMaterial {
    float color[3], roughness;
    float emission[3], metallic;
    int32_t color_tex, emission_tex, mr_tex, normal_tex;
};
struct DevTexture {
    cudaTextureObject_t raw;    // texels as the reference's texture_index returns them, srgb = false
    cudaTextureObject_t linear; // same with pow(rgb, 2.2) applied per texel (srgb = true), 0 if unused
    int32_t w, h;
    int32_t pad0, pad1;
};

struct SceneDev {
    const float4* nodes;  // WideNode[], 8 float4 each: scene BVH (root 0) followed by the light BVH
    const float4* tris;   // TriIsect[], 4 float4 (64 B) each: scene triangles followed by the light triangles
    const float4* ltris;  // = tris + 4 * light_tri_base (light sampling, shading.odin:41-50)
    int32_t light_root;       // node index of the light BVH root inside `nodes`
    uint32_t light_tri_base;  // index of the first light triangle inside `tris`
    const float4* tshade; // TriShade[], 4 float4 each
    const float4* tuv;    // TriUV[], 2 float4 each
    const float4* ttan;   // TriTan[], 3 float4 each
    const DevMaterial* mats;
    const DevTexture* texs;
    DevTexture env;
    int32_t has_env;
    int32_t n_lights;
    float pad_scale[3];  // max |coordinate| of the scene root box (light triangles are scene triangles)
};

struct RenderParams {
    float M[12]; // rows 0..2 of pixel_to_ray_dir (raytracer.odin:534-538)
    float cam_pos[3];
    uint32_t w, h, npix;
    int32_t ray_depth;
    uint32_t n_batch_samples;
    uint64_t sample_base;
    uint64_t seed;
    int32_t tiled; // primary rays enter the queue in 8x4 pixel tiles (render) or in pixel order (probes)
    int32_t tile_w, tile_h, tile_s; // tiled >= 2: pixels x samples per warp (experiment)
};

// ------------------------------------------------------------------------------------------------
// ray setup shared by both traversal kernels
// ------------------------------------------------------------------------------------------------
struct RaySetup {
    float ox, oy, oz; // origin AFTER the RAY_EPS offset (raytracer.odin:421)
    float dx, dy, dz;
    float ix, iy, iz;       // 1/d
    float nx, ny, nz;       // -o/d - pad   (added to near-plane products)
    float fx, fy, fz;       // -o/d + pad   (added to far-plane products)
    int sx, sy, sz;         // float4 index of the near plane inside a node (0/1, 2/3, 4/5)
};

// Box tests here are CONSERVATIVE with respect to the reference's check_intersect_ray_aabb
// (raytracer.odin:119-134): that test translates the ray by box.lo and divides, so its slab
// distances carry an absolute error of a few ulp of (|o| + |box|) / |d|.  Each slab distance is
// widened by `pad` = 16 ulp of that magnitude so every box the reference enters is entered here
// too; a superset of boxes cannot change the closest hit, only the triangle test decides.
//
// Direction components that are zero or tiny.  The reference divides by d: with d.x == 0 its slab interval is
// (-inf, +inf) when the origin lies inside the slab and empty otherwise.  The form used here, plane * (1/d) +
// (-o/d -+ pad), turns that into inf - inf = NaN, and fmaxf / fminf DROP NaN operands: the x slab would never
// reject, and the ray would visit every box of its yz column — one such ray was measured at 1.2-2.0 ms of serial
// traversal on the 1 M-triangle scene (profiles/r2_zero_direction_components.md), and an axis-aligned camera
// produces a few per wave (exact cancellation in pixel_to_ray_dir on one pixel column and one row).  For the box
// tests |d| is therefore clamped to 1e-18: every product stays finite for coordinates up to 1e20, an origin inside
// the slab (within the pad) still gets an interval that covers every distance the other axes allow, and an origin
// outside gets an interval beyond them — the same decisions as the reference's, as a superset.  The triangle solve
// keeps the ray's own direction.
__device__ __forceinline__ RaySetup make_ray(float4 o4, float4 d4, const float* pad_scale) {
    RaySetup r;
    r.dx = d4.x; r.dy = d4.y; r.dz = d4.z;
    r.ox = addr(o4.x, mulr(d4.x, RAY_EPS));
    r.oy = addr(o4.y, mulr(d4.y, RAY_EPS));
    r.oz = addr(o4.z, mulr(d4.z, RAY_EPS));
    const float tiny = 1e-18f;
    const float bdx = fabsf(r.dx) < tiny ? copysignf(tiny, r.dx) : r.dx;
    const float bdy = fabsf(r.dy) < tiny ? copysignf(tiny, r.dy) : r.dy;
    const float bdz = fabsf(r.dz) < tiny ? copysignf(tiny, r.dz) : r.dz;
    r.ix = 1.0f / bdx; r.iy = 1.0f / bdy; r.iz = 1.0f / bdz;
    const float ulp16 = 16.0f * 5.9604645e-08f;
    float px = ulp16 * (fabsf(r.ox) + pad_scale[0]) * fabsf(r.ix);
    float py = ulp16 * (fabsf(r.oy) + pad_scale[1]) * fabsf(r.iy);
    float pz = ulp16 * (fabsf(r.oz) + pad_scale[2]) * fabsf(r.iz);
    float bx = -r.ox * r.ix, by = -r.oy * r.iy, bz = -r.oz * r.iz;
    r.nx = bx - px; r.fx = bx + px;
    r.ny = by - py; r.fy = by + py;
    r.nz = bz - pz; r.fz = bz + pz;
    r.sx = bdx < 0.0f ? 1 : 0; // (-0.0 counts as negative, like the clamped reciprocal)
    r.sy = bdy < 0.0f ? 3 : 2;
    r.sz = bdz < 0.0f ? 5 : 4;
    return r;
}

// intersect_ray_triangle (raytracer.odin:136-150): solve [u | v | -d] (u,v,t)^T = o - p through
// adjugate * (1/det), every operation individually rounded, sums left to right.  Returns false
// when the reference would return t = -1.  `c` is the precomputed third adjugate row.
struct TriHit {
    float t, u, v;
};
__device__ __forceinline__ void tri_det_t(const RaySetup& r, float4 a, float4 b, float4 c, float& id, float& bx,
                                          float& by, float& bz, float& t, float& a00, float& a10) {
    // a = (p.x p.y p.z u.x)  b = (u.y u.z v.x v.y)  c = (v.z c0 c1 c2)
    const float m00 = a.w, m10 = b.x, m20 = b.y;
    const float m01 = b.z, m11 = b.w, m21 = c.x;
    const float m02 = -r.dx, m12 = -r.dy, m22 = -r.dz;
    (void)m00; (void)m20; (void)m01;
    a00 = subr(mulr(m11, m22), mulr(m21, m12));
    a10 = -subr(mulr(m10, m22), mulr(m20, m12));
    const float a20 = c.y;
    // det = m00*(m11*m22 - m12*m21) + (-m01)*(m10*m22 - m12*m20) + m02*(m10*m21 - m11*m20)
    //     = m00*a00 + m01*a10 + m02*a20   (same products, same roundings, same order)
    const float det = addr(addr(mulr(m00, a00), mulr(m01, a10)), mulr(m02, a20));
    id = divr(1.0f, det);
    bx = subr(r.ox, a.x); by = subr(r.oy, a.y); bz = subr(r.oz, a.z);
    t = addr(addr(mulr(mulr(c.y, id), bx), mulr(mulr(c.z, id), by)), mulr(mulr(c.w, id), bz));
}
__device__ __forceinline__ bool tri_uv(const RaySetup& r, float4 a, float4 b, float4 c, float id, float bx, float by,
                                       float bz, float a00, float a10, float& u, float& v) {
    const float m00 = a.w, m10 = b.x, m20 = b.y;
    const float m01 = b.z, m11 = b.w, m21 = c.x;
    const float m02 = -r.dx, m12 = -r.dy, m22 = -r.dz;
    (void)m20; (void)m21;
    const float a01 = -subr(mulr(m01, m22), mulr(m21, m02));
    const float a02 = subr(mulr(m01, m12), mulr(m11, m02));
    const float a11 = subr(mulr(m00, m22), mulr(m20, m02));
    const float a12 = -subr(mulr(m00, m12), mulr(m10, m02));
    u = addr(addr(mulr(mulr(a00, id), bx), mulr(mulr(a01, id), by)), mulr(mulr(a02, id), bz));
    v = addr(addr(mulr(mulr(a10, id), bx), mulr(mulr(a11, id), by)), mulr(mulr(a12, id), bz));
    return !(u < 0.0f || v < 0.0f || addr(u, v) > 1.0f);
}

} // namespace ort
#include "traverse.cuh"
namespace ort {

// ------------------------------------------------------------------------------------------------
// k_raygen: primary rays (raytracer.odin:580-586).  slot = s_local * npix + (py*w + px).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ f3 primary_dir(const RenderParams& p, uint32_t px, uint32_t py, uint32_t pix, uint64_t sample) {
    const Philox4 rr = philox4x32_10(pix, (uint32_t)sample, (uint32_t)(sample >> 32), 0u, (uint32_t)p.seed,
                                     (uint32_t)(p.seed >> 32));
    const float x = addr((float)px, u01(rr.r0));
    const float y = addr((float)py, u01(rr.r1));
    // (M * {x, y, 0, 1}).xyz, each row summed left to right
    float v[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const float* m = p.M + 4 * r;
        v[r] = addr(addr(addr(mulr(m[0], x), mulr(m[1], y)), mulr(m[2], 0.0f)), mulr(m[3], 1.0f));
    }
    const float len = sqrtr(addr(addr(mulr(v[0], v[0]), mulr(v[1], v[1])), mulr(v[2], v[2])));
    return mk3(divr(v[0], len), divr(v[1], len), divr(v[2], len));
}

// Queue position -> pixel.  The path SLOT of a ray stays s_local * npix + (py * w + px) (RNG keys, state
// and accumulation are indexed by it), but the ORDER in which primary rays enter the queue follows 8x4
// pixel tiles, so the 32 rays a traversal warp fetches together cover a compact 8x4 footprint instead
// of a 32x1 strip (more shared nodes per request, fewer distinct sectors).  Pixels outside the region
// covered by whole tiles (w % 8 columns on the right, h % 4 rows at the bottom) follow in row order.
__device__ __forceinline__ uint32_t tiled_pixel(uint32_t q, uint32_t w, uint32_t h) {
    const uint32_t w8 = w & ~7u, h4 = h & ~3u, area = w8 * h4;
    if (q < area) {
        const uint32_t tile = q >> 5, in = q & 31u, tiles_x = w8 >> 3;
        const uint32_t ty = tile / tiles_x, tx = tile - ty * tiles_x;
        return (ty * 4u + (in >> 3)) * w + tx * 8u + (in & 7u);
    }
    uint32_t r = q - area;
    const uint32_t right = (w - w8) * h; // the right strip, all rows
    if (r < right) {
        const uint32_t y = r / (w - w8);
        return y * w + w8 + (r - y * (w - w8));
    }
    r -= right; // the bottom strip under the tiled region
    const uint32_t y = r / w8;
    return (h4 + y) * w + (r - y * w8);
}

__global__ void k_raygen(const RenderParams p, float4* __restrict__ qo, float4* __restrict__ qd,
                         uint32_t* __restrict__ count0) {
    const uint32_t n = p.n_batch_samples * p.npix;
    if (blockIdx.x == 0 && threadIdx.x == 0) *count0 = n;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t s_local, pix;
        if (p.tiled >= 2 && (p.w % (uint32_t)p.tile_w) == 0u && (p.h % (uint32_t)p.tile_h) == 0u &&
            (p.n_batch_samples % (uint32_t)p.tile_s) == 0u) {
            // a warp = tile_w x tile_h pixels x tile_s samples (product 32; default 2 x 2 x 8): the 32 rays
            // differ only by sub-pixel jitter and one pixel step, and the sample groups of one pixel tile
            // follow each other in the queue.  Measured against 8x4x1: k_trace -7 % (profiles/r1_sensitivity.md)
            const uint32_t tw = (uint32_t)p.tile_w, th = (uint32_t)p.tile_h, ts = (uint32_t)p.tile_s;
            const uint32_t tp = tw * th;
            const uint32_t block = i >> 5, in = i & 31u;
            const uint32_t sin = in / tp, pin = in - sin * tp;
            const uint32_t sgroups = p.n_batch_samples / ts;
            const uint32_t pg = block / sgroups, sg = block - pg * sgroups;
            const uint32_t tiles_x = p.w / tw;
            const uint32_t ty = pg / tiles_x, tx = pg - ty * tiles_x;
            const uint32_t iy = pin / tw, ix = pin - iy * tw;
            s_local = sg * ts + sin;
            pix = (ty * th + iy) * p.w + tx * tw + ix;
        } else {
            s_local = i / p.npix;
            pix = p.tiled ? tiled_pixel(i - s_local * p.npix, p.w, p.h) : i - s_local * p.npix;
        }
        const uint32_t py = pix / p.w, px = pix - py * p.w;
        const f3 d = primary_dir(p, px, py, pix, p.sample_base + s_local);
        qo[i] = make_float4(p.cam_pos[0], p.cam_pos[1], p.cam_pos[2], __uint_as_float(s_local * p.npix + pix));
        qd[i] = make_float4(d.x, d.y, d.z, 0.0f);
    }
}

// ------------------------------------------------------------------------------------------------
// textures (textures.odin:79-135): texel fetches through texture objects (point, unnormalised),
// repeat wrap by floored modulo, no half-texel offset, bilinear weights in f32.
// ------------------------------------------------------------------------------------------------
// Odin's `%%` (floored modulo) on the integer texel coordinate (textures.odin:120-121).  32-bit
// arithmetic: identical to the reference's i64 for |uv * dims| < 2^31 texels (the conversion
// saturates beyond that, where f32 texel coordinates carry no sub-texel information anyway).
__device__ __forceinline__ int wrap_i(float f, int m) {
    const int i = __float2int_rz(f);
    const int r = i % m;
    return r < 0 ? r + m : r;
}
__device__ __forceinline__ float lerp1(float a, float b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ float4 lerp4(float4 a, float4 b, float t) {
    return make_float4(lerp1(a.x, b.x, t), lerp1(a.y, b.y, t), lerp1(a.z, b.z, t), lerp1(a.w, b.w, t));
}
// Not inlined: five call sites (metallic-roughness, normal, colour, emission, environment) would
// otherwise replicate ~180 instructions each in a kernel that already exceeds the instruction cache.
__device__ __noinline__ float4 texture_sample(const DevTexture& tx, bool srgb, float cu, float cv) {
    const cudaTextureObject_t obj = srgb ? tx.linear : tx.raw;
    const float pcx = cu * (float)tx.w, pcy = cv * (float)tx.h;
    const float lox = floorf(pcx), loy = floorf(pcy);
    const float hix = ceilf(pcx), hiy = ceilf(pcy);
    const float tx_ = pcx - lox, ty_ = pcy - loy;
    const float x0 = (float)wrap_i(lox, tx.w) + 0.5f, y0 = (float)wrap_i(loy, tx.h) + 0.5f;
    const float x1 = (float)wrap_i(hix, tx.w) + 0.5f, y1 = (float)wrap_i(hiy, tx.h) + 0.5f;
    const float4 p00 = tex2D<float4>(obj, x0, y0);
    const float4 p01 = tex2D<float4>(obj, x0, y1);
    const float4 p10 = tex2D<float4>(obj, x1, y0);
    const float4 p11 = tex2D<float4>(obj, x1, y1);
    return lerp4(lerp4(p00, p01, ty_), lerp4(p10, p11, ty_), tx_);
}

// ------------------------------------------------------------------------------------------------
// shading.odin
// ------------------------------------------------------------------------------------------------
struct Quat {
    float w, x, y, z;
};
__device__ __forceinline__ f3 quat_mul_vec(Quat q, f3 v) { // linalg.mul(quaternion, vector)
    const f3 qv = mk3(q.x, q.y, q.z);
    const f3 t = cross3(2.0f * qv, v);
    return v + q.w * t + cross3(qv, t);
}
__device__ __forceinline__ Quat vndf_rotation(f3 n) { // shading.odin:104-106
    const float w = sqrtf((1.0f + n.z) / 2.0f);
    if (w > 0.0f) return {w, -n.y / (2.0f * w), n.x / (2.0f * w), 0.0f};
    return {0.0f, 1.0f, 0.0f, 0.0f};
}
__device__ __forceinline__ Quat qconj(Quat q) { return {q.w, -q.x, -q.y, -q.z}; }

__device__ f3 vndf_sampling(f3 n, f3 omega, float alpha, float u1, float u2) { // shading.odin:102-122
    const Quat rot = vndf_rotation(n);
    const f3 V = quat_mul_vec(qconj(rot), omega);
    const f3 Vh = normalize3(mk3(alpha * V.x, alpha * V.y, V.z));
    const float len = hypotf(Vh.x, Vh.y);
    const f3 T1 = len == 0.0f ? mk3(1, 0, 0) : mk3(-Vh.y / len, Vh.x / len, 0.0f);
    const f3 T2 = cross3(Vh, T1);
    const float r = sqrtf(u1);
    const float phi = TAU_F * u2;
    float t1, t2;
    sincosf(phi, &t1, &t2);
    t1 *= r;
    t2 *= r;
    const float s = 0.5f * (1.0f + Vh.z);
    t2 = (1.0f - s) * sqrtf(1.0f - sq(t1)) + s * t2;
    const f3 Nh = t1 * T1 + t2 * T2 + Vh * sqrtf(omax(0.0f, 1.0f - sq(t1) - sq(t2)));
    const f3 Ne = normalize3(mk3(alpha * Nh.x, alpha * Nh.y, omax(0.0f, Nh.z)));
    return quat_mul_vec(rot, Ne);
}
__device__ float vndf_sampling_pdf(f3 n, f3 omega, float alpha, f3 L) { // shading.odin:124-137
    const f3 Ne = normalize3(omega + L);
    const Quat rot = vndf_rotation(n);
    const f3 V = quat_mul_vec(qconj(rot), omega);
    const f3 N = quat_mul_vec(qconj(rot), Ne);
    const float alpha2 = sq(alpha);
    const float lambda = (-1.0f + sqrtf(1.0f + alpha2 * (sq(V.x) + sq(V.y)) / sq(V.z))) * 0.5f;
    const float G1 = 1.0f / (1.0f + lambda);
    const float D = 1.0f / (PI_F * alpha2 * sq(sq(N.x / alpha) + sq(N.y / alpha) + sq(N.z)));
    const float normal = G1 * omax(0.0f, dot3(V, N)) * D / V.z;
    return normal / (4.0f * dot3(L, Ne));
}
__device__ __forceinline__ float smith_ggx(f3 n, f3 x, float alpha2) { // shading.odin:187-190
    const float c = dot3(n, x);
    return 2.0f * omax(c, 0.0f) / (c + sqrtf(alpha2 + (1.0f - alpha2) * sq(c)));
}
__device__ f3 brdf_cos(f3 color, f3 N, float metallic, float roughness, f3 in_d, f3 L) { // shade, shading.odin:164-204
    const float alpha = sq(roughness);
    const float alpha2 = sq(alpha);
    const f3 V = -in_d;
    const f3 H = normalize3(L + V);
    const float cosine = dot3(L, N);
    const float fb = powf(1.0f - dot3(H, L), 5.0f);
    const float f_ds = 0.04f + 0.96f * fb;
    const f3 f_m = color + (mk3(1, 1, 1) - color) * fb;
    const float hn = dot3(H, N);
    const float step = hn < 0.0f ? 0.0f : 1.0f;
    const float Dt = alpha2 * step / (PI_F * sq((alpha2 - 1.0f) * sq(hn) + 1.0f));
    const float G = smith_ggx(N, L, alpha2) * smith_ggx(N, V, alpha2);
    const float ct = Dt * G / (4.0f * dot3(V, N));
    const f3 spec = ct * mk3(1, 1, 1);
    const f3 diff = color * omax(cosine, 0.0f) / PI_F;
    const f3 met = spec * f_m;
    const f3 diel = diff * (1.0f - f_ds) + spec * f_ds;
    return diel * (1.0f - metallic) + met * metallic;
}

// sphere_uniform + n, normalised (cosine_weighted, shading.odin:9-15,32-35)
__device__ __forceinline__ f3 cosine_weighted(f3 N, uint32_t r1, uint32_t r2) {
    const float phi = u01(r1) * (TAU_F - 0.0f) + 0.0f;
    const float z = u01(r2) * (1.0f - -1.0f) + -1.0f;
    float sx, sy;
    sincosf(phi, &sx, &sy);
    const float radius = sqrtf(1.0f - sq(z));
    return normalize3(mk3(sx * radius, sy * radius, z) + N);
}
__device__ __forceinline__ float cosine_weighted_pdf(f3 N, f3 omega) { return omax(dot3(N, omega) / PI_F, 0.0f); } // shading.odin:37-39
// surface_sampling (shading.odin:41-50): uniform light triangle, uniform point on it
__device__ __forceinline__ f3 surface_sampling(const SceneDev& s, f3 P, uint32_t r1, uint32_t r2, uint32_t r3) {
    const uint32_t idx = (uint32_t)(((uint64_t)r1 * (uint64_t)(uint32_t)s.n_lights) >> 32);
    const float4* lp = s.ltris + (size_t)idx * 4;
    const float4 la = ldg4(lp), lb = ldg4(lp + 1), lc = ldg4(lp + 2);
    float su = u01(r2) * (1.0f - 0.0f) + 0.0f, sv = u01(r3) * (1.0f - 0.0f) + 0.0f;
    if (su + sv > 1.0f) { su = 1.0f - su; sv = 1.0f - sv; }
    const f3 world = mk3(la.x, la.y, la.z) + su * mk3(la.w, lb.x, lb.y) + sv * mk3(lb.z, lb.w, lc.x);
    return normalize3(world - P);
}
// sample (shading.odin:139-151): r0 picks the strategy, r1.. are the strategy's draws in source order
__device__ __forceinline__ f3 sample_dir(const SceneDev& s, f3 N, f3 P, float roughness, f3 in_d, const Philox4& rr) {
    const float tsel = u01(rr.r0);
    if (tsel <= 0.33333f) return cosine_weighted(N, rr.r1, rr.r2);
    if (tsel < 0.666666f && s.n_lights > 0) return surface_sampling(s, P, rr.r1, rr.r2, rr.r3);
    const f3 hn = vndf_sampling(N, -in_d, sq(roughness), u01(rr.r1), u01(rr.r2));
    return in_d - 2.0f * dot3(hn, in_d) * hn;
}
// pdf (shading.odin:153-162) from its three terms: the light term is the light-BVH all-hit sum of the
// ray (k_trace<true>) / len(light_surfaces); vndf_term already carries the (1 if has_lights else 2)
__device__ __forceinline__ float vndf_pdf_term(const SceneDev& s, f3 N, f3 in_d, float roughness, f3 nd) {
    return vndf_sampling_pdf(N, -in_d, sq(roughness), nd) * (s.n_lights > 0 ? 1.0f : 2.0f);
}
__device__ __forceinline__ float complete_pdf(const SceneDev& s, float cos_pdf, float lsum, float vndf_term) {
    const float lp = s.n_lights > 0 ? lsum / (float)s.n_lights : 0.0f;
    return (cos_pdf + lp + vndf_term) / 3.0f;
}
// miss: equirectangular environment lookup (raytracer.odin:437-446)
__device__ __forceinline__ float4 env_lookup(const SceneDev& s, float dx, float dy, float dz) {
    const float tu = 0.5f + atan2f(dz, dx) / TAU_F;
    const float tv = 0.5f - asinf(dy) / PI_F;
    return texture_sample(s.env, false, tu, tv);
}

// ------------------------------------------------------------------------------------------------
// k_shade: one wavefront step of raytrace (raytracer.odin:432-500) in iterative form.
// Path state: L = sum_k T_k * emission_k lives in st_c BY SLOT and is only touched when a hit adds
// something (emissive surface, environment); the pending (T, value, pdf terms) of the sampled
// direction travel WITH THE RAY in queue order:  pa = (T.rgb, cosine pdf)  pb = (value.rgb, vndf pdf term).
// For bounce > 0 the pending (value, pdf) of the previous hit is completed first: the light-BVH
// sum of the ray just traced is the missing third of pdf (shading.odin:153-162), then
// `norm_l1(value)/pdf > 1e-5` (raytracer.odin:495) decides whether this hit counts at all.
//
// Two phases per block of 256 queue entries.  Phase A (one thread per entry): completes the pdf,
// handles misses (environment lookup) and dead paths.  The entries that need the expensive part
// (material fetch, sampling, BRDF) are then compacted inside the block, so phase B runs on dense
// warps: 30-40 % of the rays of an open scene miss, and their lanes would otherwise idle through
// ~600 instructions.
// ------------------------------------------------------------------------------------------------
// Can the ray reach ANY light?  Same conservative slab test as the traversal, applied to the (up
// to four) child boxes of the light BVH's root.  Rays that fail have a light-pdf sum of exactly 0
// (surface_sampling_pdf_bvh_sum never gets past shading.odin:86-89) and skip the light pass.  The
// result is the 4-bit mask of the root's children whose boxes the ray enters: it travels with the
// ray's light-queue entry (bits 28..31, LQ_MASK_SHIFT), and k_trace<light> starts at those children
// instead of visiting the root a second time — for most candidates that is one of two or three node
// round trips (profiles/r2r_light_skip_root.md).
__device__ __forceinline__ unsigned light_root_mask(const SceneDev& s, float4 o4, float4 d4) {
    const RaySetup r = make_ray(o4, d4, s.pad_scale);
    const float4* nd = s.nodes + (size_t)s.light_root * 8;
    const float4 nxp = ldg4(nd + r.sx), fxp = ldg4(nd + (r.sx ^ 1));
    const float4 nyp = ldg4(nd + r.sy), fyp = ldg4(nd + (r.sy ^ 1));
    const float4 nzp = ldg4(nd + r.sz), fzp = ldg4(nd + (r.sz ^ 1));
    const int4 ch = __ldg(reinterpret_cast<const int4*>(nd + 6));
    unsigned m = 0u;
#define ORT_BOX(k, C, BIT)                                                                        \
    {                                                                                             \
        const float tn = fmaxf(fmaxf(fmaf(nxp.k, r.ix, r.nx), fmaf(nyp.k, r.iy, r.ny)),           \
                               fmaxf(fmaf(nzp.k, r.iz, r.nz), 0.0f));                             \
        const float tf = fminf(fminf(fmaf(fxp.k, r.ix, r.fx), fmaf(fyp.k, r.iy, r.fy)),           \
                               fmaf(fzp.k, r.iz, r.fz));                                          \
        if (tn <= tf && C != WIDE_EMPTY) m |= BIT;                                                \
    }
    ORT_BOX(x, ch.x, 1u) ORT_BOX(y, ch.y, 2u) ORT_BOX(z, ch.z, 4u) ORT_BOX(w, ch.w, 8u)
#undef ORT_BOX
    return m;
}

#ifndef ORT_SHADE_MIN_CTAS
#define ORT_SHADE_MIN_CTAS 3
#endif
struct ShadeArgs {
    const float4 *qo_in, *qd_in, *hits, *pa_in, *pb_in;
    const float* lsum;
    const uint32_t* n_in_ptr;
    float4 *qo_out, *qd_out, *pa_out, *pb_out;
    uint32_t *n_out_ptr, *used_ptr;
    float4* st_c;
    float* lsum_out;
    uint32_t *lq, *lq_count;
    int bounce, bin_octants;
    int prefilter; // 0: every continuation ray is a light candidate; 1: only rays that enter a child box of the light
                   // root; 2: the same, and the entered children travel with the queue entry (LQ_MASK_SHIFT)
};

__global__ void __launch_bounds__(256, ORT_SHADE_MIN_CTAS)
k_shade(const SceneDev s, const RenderParams p, const ShadeArgs a) {
    __shared__ uint32_t s_cnt[16], s_off[16], s_wsum[8], s_qn;
    __shared__ float4 s_q[512]; // pending work items of this block: (T.rgb, queue position)
    const int bounce = a.bounce;
    const uint32_t n_in = *a.n_in_ptr;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool has_lights = s.n_lights > 0;
    if (threadIdx.x == 0) s_qn = 0u;
    __syncthreads();
    for (uint32_t base = blockIdx.x * blockDim.x;; base += gridDim.x * blockDim.x) {
        // ---- phase A: complete the pending pdf, misses, dead paths
        const bool more = base < n_in; // block-uniform: queue entries left for this block
        const uint32_t pos = base + threadIdx.x;
        bool used = false, heavy = false;
        f3 T = mk3(1, 1, 1);
        if (more && pos < n_in) {
            const float4 h4 = a.hits[pos];
            bool alive = true;
            if (bounce > 0) {
                const float4 pa = a.pa_in[pos], pb = a.pb_in[pos];
                const f3 value = mk3(pb.x, pb.y, pb.z);
                // pdf (shading.odin:158-161): (cosine + light + vndf * (1 | 2)) / 3
                const float pdf = complete_pdf(s, pa.w, has_lights ? a.lsum[pos] : 0.0f, pb.w);
                if (norm_l1(value) / pdf > 1e-5f) T = mk3(pa.x, pa.y, pa.z) * value / pdf;
                else alive = false; // exitance = emission only: L already holds it
            }
            used = alive; // this traversal is a cast_ray call the reference makes (raytracer.odin:496)
            if (alive) {
                if (__float_as_int(h4.w) < 0) {
                    // miss: equirectangular env lookup (raytracer.odin:437-446), black without a map
                    if (s.has_env) {
                        const float4 d4 = a.qd_in[pos];
                        const uint32_t slot = __float_as_uint(a.qo_in[pos].w);
                        const float4 e = env_lookup(s, d4.x, d4.y, d4.z);
                        const float4 c = a.st_c[slot];
                        const f3 L = mk3(c.x, c.y, c.z) + T * mk3(e.x, e.y, e.z);
                        a.st_c[slot] = make_float4(L.x, L.y, L.z, 0.0f);
                    }
                } else {
                    heavy = true;
                }
            }
        }
        const unsigned umask = __ballot_sync(0xffffffffu, used);
        if (umask && lane == 0) atomicAdd(a.used_ptr, (uint32_t)__popc(umask));
        // ---- append the work items to the block's queue; phase B runs once 256 are pending (or at
        //      the very end), so every warp of the block shades a full complement of hits
        const unsigned hmask = __ballot_sync(0xffffffffu, heavy);
        if (threadIdx.x < 16) s_cnt[threadIdx.x] = 0u;
        if (lane == 0) s_wsum[wid] = (uint32_t)__popc(hmask);
        __syncthreads();
        uint32_t wbase = 0, n_new = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            const uint32_t c = s_wsum[w];
            if (w < wid) wbase += c;
            n_new += c;
        }
        uint32_t qn = s_qn; // < 256 here
        if (heavy) s_q[qn + wbase + __popc(hmask & ((1u << lane) - 1u))] = make_float4(T.x, T.y, T.z, __uint_as_float(pos));
        qn += n_new;
        __syncthreads();
        const bool run_b = qn >= 256u || (!more && qn > 0u);
        const uint32_t n_work = run_b ? (qn < 256u ? qn : 256u) : 0u;
        const uint32_t q_first = qn - n_work; // the newest n_work items are taken
        if (!run_b) {
            if (threadIdx.x == 0) s_qn = qn;
            __syncthreads();
            if (!more) break;
            continue;
        }

        // ---- phase B: dense warps shade the hits
        bool emit = false;
        float4 out_o = make_float4(0, 0, 0, 0), out_d = make_float4(0, 0, 0, 0);
        float4 out_a = make_float4(0, 0, 0, 0), out_b = make_float4(0, 0, 0, 0);
        if (threadIdx.x < n_work) {
            const float4 w4 = s_q[q_first + threadIdx.x];
            const uint32_t wpos = __float_as_uint(w4.w);
            T = mk3(w4.x, w4.y, w4.z);
            const float4 o4 = a.qo_in[wpos];
            const float4 d4 = a.qd_in[wpos];
            const float4 h4 = a.hits[wpos];
            const uint32_t slot = __float_as_uint(o4.w);
            const f3 in_d = mk3(d4.x, d4.y, d4.z);
            const int tri = __float_as_int(h4.w);
            const float u = h4.y, v = h4.z;
            const float4* ts = s.tshade + (size_t)tri * 4;
            const float4 s0 = ldg4(ts), s1 = ldg4(ts + 1), s2 = ldg4(ts + 2);
            const int4 s3 = __ldg(reinterpret_cast<const int4*>(ts + 3));
            const DevMaterial m = s.mats[s3.x];
            const float4* tp = s.tris + (size_t)tri * 4;
            const float4 ta = ldg4(tp), tb = ldg4(tp + 1), tc = ldg4(tp + 2);
            // p = trig.p + trig.u*u + trig.v*v (raytracer.odin:456), individually rounded
            const f3 P = mk3(addr(addr(ta.x, mulr(ta.w, u)), mulr(tb.z, v)),
                             addr(addr(ta.y, mulr(tb.x, u)), mulr(tb.w, v)),
                             addr(addr(ta.z, mulr(tb.y, u)), mulr(tc.x, v)));
            const float w0 = 1.0f - u - v;
            const bool any_tex = m.color_tex >= 0 || m.emission_tex >= 0 || m.mr_tex >= 0 || m.normal_tex >= 0;
            float tcx = 0.0f, tcy = 0.0f;
            if (any_tex) {
                const float4 uv0 = ldg4(s.tuv + (size_t)tri * 2), uv1 = ldg4(s.tuv + (size_t)tri * 2 + 1);
                tcx = uv0.x * w0 + uv0.z * u + uv1.x * v; // raytracer.odin:454
                tcy = uv0.y * w0 + uv0.w * u + uv1.y * v;
            }
            float4 mr = make_float4(1, 1, 1, 1);
            if (m.mr_tex >= 0) mr = texture_sample(s.texs[m.mr_tex], false, tcx, tcy);
            const f3 n_interp = mk3(s0.x, s0.y, s0.z) * w0 + mk3(s1.x, s1.y, s1.z) * u + mk3(s2.x, s2.y, s2.z) * v;
            f3 N;
            if (m.normal_tex >= 0) { // raytracer.odin:458-470
                const float4* tt = s.ttan + (size_t)tri * 3;
                const float4 g0 = ldg4(tt), g1 = ldg4(tt + 1), g2 = ldg4(tt + 2);
                float t4x = g0.x * w0 + g1.x * u + g2.x * v, t4y = g0.y * w0 + g1.y * u + g2.y * v;
                float t4z = g0.z * w0 + g1.z * u + g2.z * v, t4w = g0.w * w0 + g1.w * u + g2.w * v;
                const float l4 = sqrtf(t4x * t4x + t4y * t4y + t4z * t4z + t4w * t4w); // [4]f32 normalize
                t4x /= l4; t4y /= l4; t4z /= l4; t4w /= l4;
                const f3 lx = mk3(t4x, t4y, t4z);
                const f3 lz = normalize3(n_interp);
                const f3 ly = cross3(lz, lx) * t4w;
                const float4 ns = texture_sample(s.texs[m.normal_tex], false, tcx, tcy);
                const f3 ln = mk3(ns.x, ns.y, ns.z) * 2.0f - mk3(1, 1, 1);
                N = normalize3(mk3(lx.x * ln.x + ly.x * ln.y + lz.x * ln.z, lx.y * ln.x + ly.y * ln.y + lz.y * ln.z,
                                   lx.z * ln.x + ly.z * ln.y + lz.z * ln.z));
            } else {
                N = normalize3(n_interp); // raytracer.odin:472
            }
            f3 color = mk3(m.color[0], m.color[1], m.color[2]);
            f3 emission = mk3(m.emission[0], m.emission[1], m.emission[2]);
            if (m.color_tex >= 0) {
                const float4 c = texture_sample(s.texs[m.color_tex], true, tcx, tcy);
                color = color * mk3(c.x, c.y, c.z);
            }
            if (m.emission_tex >= 0) {
                const float4 c = texture_sample(s.texs[m.emission_tex], true, tcx, tcy);
                emission = emission * mk3(c.x, c.y, c.z);
            }
            const float roughness = omax(m.roughness * mr.y, 0.03f); // raytracer.odin:480
            const float metallic = m.metallic * mr.z;
            // inside = dot(ng, d) > 0 (raytracer.odin:148), flips the shading normal (:485-488)
            const float ngd = addr(addr(mulr(s0.w, in_d.x), mulr(s1.w, in_d.y)), mulr(s2.w, in_d.z));
            if (ngd > 0.0f) N = -N;
            // L = L + T * emission: the accumulator is only touched when this hit emits (adding an
            // exact zero would leave it unchanged)
            if (emission.x != 0.0f || emission.y != 0.0f || emission.z != 0.0f) {
                const float4 c = a.st_c[slot];
                const f3 L = mk3(c.x, c.y, c.z) + T * emission;
                a.st_c[slot] = make_float4(L.x, L.y, L.z, 0.0f);
            }
            bool cont = bounce + 1 < p.ray_depth; // raytrace(.., depth_left - 1) with depth_left == 1 returns 0
            f3 nd = mk3(0, 0, 0), value = mk3(0, 0, 0);
            if (cont) {
                // sample (shading.odin:139-151)
                const uint32_t s_local = slot / p.npix, pix = slot - s_local * p.npix;
                const uint64_t smp = p.sample_base + s_local;
                const Philox4 rr = philox4x32_10(pix, (uint32_t)smp, (uint32_t)(smp >> 32), 1u + (uint32_t)bounce,
                                                 (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
                nd = sample_dir(s, N, P, roughness, in_d, rr);
                value = brdf_cos(color, N, metallic, roughness, in_d, nd);
                // norm_l1(value)/pdf > 1e-5 can only hold for norm_l1(value) > 0
                if (!(norm_l1(value) > 0.0f)) cont = false;
            }
            if (cont) {
                const float cos_pdf = cosine_weighted_pdf(N, nd);
                const float vndf_term = vndf_pdf_term(s, N, in_d, roughness, nd);
                emit = true;
                out_a = make_float4(T.x, T.y, T.z, cos_pdf);
                out_b = make_float4(value.x, value.y, value.z, vndf_term);
                out_o = make_float4(P.x, P.y, P.z, o4.w);
                out_d = make_float4(nd.x, nd.y, nd.z, 0.0f);
            }
        }
        if (a.bin_octants) {
            // Queue compaction per BLOCK with an 8-bin counting sort on the direction octant: the 256
            // paths of this iteration come from neighbouring pixels, so each run in the queue holds
            // rays with nearby origins AND the same direction signs.  Rays fetched together by a
            // traversal warp then share nodes (one L1 wavefront serves several lanes) and child order.
            const int oct = (out_d.x < 0.0f ? 1 : 0) | (out_d.y < 0.0f ? 2 : 0) | (out_d.z < 0.0f ? 4 : 0);
            bool cand = emit && has_lights;
            unsigned lmask = 0u; // 0 = "start at the root" (no prefilter)
            if (cand && a.prefilter) {
                lmask = light_root_mask(s, out_o, out_d);
                cand = lmask != 0u;
                if (a.prefilter < 2) lmask = 0u; // positions need all 32 bits: the light pass starts at the root
            }
            uint32_t rank = 0, lrank = 0;
            if (emit) rank = atomicAdd(&s_cnt[oct], 1u);
            if (cand) lrank = atomicAdd(&s_cnt[8 + oct], 1u);
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t tot = 0, ltot = 0;
                for (int i = 0; i < 8; i++) { tot += s_cnt[i]; ltot += s_cnt[8 + i]; }
                uint32_t b0 = tot ? atomicAdd(a.n_out_ptr, tot) : 0u;
                uint32_t b1 = ltot ? atomicAdd(a.lq_count, ltot) : 0u;
                for (int i = 0; i < 8; i++) { s_off[i] = b0; b0 += s_cnt[i]; s_off[8 + i] = b1; b1 += s_cnt[8 + i]; }
            }
            __syncthreads();
            if (emit) {
                const uint32_t q = s_off[oct] + rank;
                a.qo_out[q] = out_o;
                a.qd_out[q] = out_d;
                a.pa_out[q] = out_a;
                a.pb_out[q] = out_b;
                if (has_lights) {
                    if (cand) a.lq[s_off[8 + oct] + lrank] = q | (lmask << LQ_MASK_SHIFT);
                    else a.lsum_out[q] = 0.0f;
                }
            }
            if (threadIdx.x == 0) s_qn = q_first;
            __syncthreads(); // s_cnt / s_off / s_q / s_qn are reused by the next iteration
            if (!more && q_first == 0u) break;
            continue;
        }
        // queue compaction: warp ballot + prefix popcount + one atomic per warp
        const unsigned mask = __ballot_sync(0xffffffffu, emit);
        if (mask) {
            const int leader = __ffs(mask) - 1;
            uint32_t qbase = 0;
            if (lane == leader) qbase = atomicAdd(a.n_out_ptr, (uint32_t)__popc(mask));
            qbase = __shfl_sync(0xffffffffu, qbase, leader);
            uint32_t q = 0;
            if (emit) {
                q = qbase + __popc(mask & ((1u << lane) - 1u));
                a.qo_out[q] = out_o;
                a.qd_out[q] = out_d;
                a.pa_out[q] = out_a;
                a.pb_out[q] = out_b;
            }
            if (has_lights) {
                // second queue: only rays that can reach a light need the light-BVH pass
                bool cand = emit;
                unsigned lmask = 0u;
                if (emit && a.prefilter) {
                    lmask = light_root_mask(s, out_o, out_d);
                    if (lmask == 0u) { cand = false; a.lsum_out[q] = 0.0f; }
                    if (a.prefilter < 2) lmask = 0u;
                }
                const unsigned cmask = __ballot_sync(0xffffffffu, cand);
                if (cmask) {
                    const int cl = __ffs(cmask) - 1;
                    uint32_t cbase = 0;
                    if (lane == cl) cbase = atomicAdd(a.lq_count, (uint32_t)__popc(cmask));
                    cbase = __shfl_sync(0xffffffffu, cbase, cl);
                    if (cand) a.lq[cbase + __popc(cmask & ((1u << lane) - 1u))] = q | (lmask << LQ_MASK_SHIFT);
                }
            }
        }
        if (threadIdx.x == 0) s_qn = q_first;
        __syncthreads(); // s_q / s_qn / s_wsum are reused by the next iteration
        if (!more && q_first == 0u) break;
    }
}

// ------------------------------------------------------------------------------------------------
// k_resolve: rc_set_pixel (main.odin:89-102) for every sample of the wave, in sample order.
// accum planes: total.rgb | total_squared.rgb | count_lo | count_hi, index (H-1-y)*W + x.
// Sample_Stats.count is a u32 (main.odin:36): the count is kept as count_lo + 2^20 * count_hi with
// count_lo < 2^20 after every wave, so both planes stay exact integers in f32 — also under the one
// float sum-reduce over up to 16 GPUs — however many waves a caller accumulates.
// ------------------------------------------------------------------------------------------------
constexpr float COUNT_RADIX = 1048576.0f; // 2^20
__device__ __forceinline__ uint32_t accum_count(const float* accum, uint32_t npix, uint32_t i) {
    const unsigned long long c = (unsigned long long)accum[6 * npix + i] + ((unsigned long long)accum[7 * npix + i] << 20);
    return c > 0xffffffffull ? 0xffffffffu : (uint32_t)c;
}
__global__ void k_resolve(const RenderParams p, const float4* __restrict__ st_c, float* __restrict__ accum,
                          float* __restrict__ first, float* __restrict__ last, const int write_first,
                          const int write_last) {
    const uint32_t npix = p.npix;
    for (uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.x * blockDim.x) {
        const uint32_t py = pix / p.w, px = pix - py * p.w;
        const uint32_t i = (p.h - py - 1) * p.w + px;
        float tr = accum[i], tg = accum[npix + i], tb = accum[2 * npix + i];
        float qr = accum[3 * npix + i], qg = accum[4 * npix + i], qb = accum[5 * npix + i];
        float4 c = make_float4(0, 0, 0, 0);
        for (uint32_t sl = 0; sl < p.n_batch_samples; sl++) {
            c = st_c[(size_t)sl * npix + pix];
            if (sl == 0 && write_first) { first[i] = c.x; first[npix + i] = c.y; first[2 * npix + i] = c.z; }
            tr += c.x; tg += c.y; tb += c.z;
            qr += c.x * c.x; qg += c.y * c.y; qb += c.z * c.z;
        }
        accum[i] = tr; accum[npix + i] = tg; accum[2 * npix + i] = tb;
        accum[3 * npix + i] = qr; accum[4 * npix + i] = qg; accum[5 * npix + i] = qb;
        float lo = accum[6 * npix + i] + (float)p.n_batch_samples, hi = accum[7 * npix + i];
        while (lo >= COUNT_RADIX) { lo -= COUNT_RADIX; hi += 1.0f; }
        accum[6 * npix + i] = lo; accum[7 * npix + i] = hi;
        if (write_last) { last[i] = c.x; last[npix + i] = c.y; last[2 * npix + i] = c.z; }
    }
}

// counters: [0 .. D] queue sizes per bounce; used[k]: traversals of bounce k whose result counted.
// stats: [0] reference-equivalent closest rays  [1] light rays  [2] paths  [3] traversals launched
__global__ void k_stats(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ used,
                        const uint32_t* __restrict__ lcounts, const int depth, const int has_lights,
                        unsigned long long* __restrict__ stats) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long rays = 0, lrays = 0, traced = 0;
        for (int k = 0; k < depth; k++) {
            rays += used[k];
            traced += counts[k];
            if (k > 0 && has_lights) lrays += lcounts[k];
        }
        stats[0] += rays;
        stats[1] += lrays;
        stats[2] += counts[0];
        stats[3] += traced;
    }
}

// ---- probes / packing ---------------------------------------------------------------------------
// ort_probe_shading: one evaluation per thread of the device functions k_shade calls (record layouts in
// include/odinrt_b200.h).  `lsum` = light-BVH sums of the ORT_PROBE_PDF rays (k_trace<true> ran first).
__host__ __device__ inline int probe_in_floats(int kind) {
    return kind == 0 ? 14 : kind == 1 ? 9 : kind == 2 ? 10 : kind == 3 ? 14 : kind == 4 ? 13 : kind == 5 ? 4 : kind == 6 ? 5 : 3;
}
__host__ __device__ inline int probe_out_floats(int kind) {
    return kind == 2 || kind == 4 ? 1 : kind == 5 || kind == 6 ? 4 : 3;
}
__global__ void k_probe_pack(const float* __restrict__ in, const uint32_t n, float4* __restrict__ qo, float4* __restrict__ qd,
                             uint32_t* __restrict__ count0) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *count0 = n;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* x = in + (size_t)i * 13; // ORT_PROBE_PDF: n[3] pos[3] roughness in_d[3] out_d[3]
        qo[i] = make_float4(x[3], x[4], x[5], __uint_as_float(i));
        qd[i] = make_float4(x[10], x[11], x[12], 0.0f);
    }
}
__global__ void k_probe(const SceneDev s, const int kind, const float* __restrict__ in, const uint32_t n,
                        float* __restrict__ out, const float* __restrict__ lsum) {
    const int ni = probe_in_floats(kind), no = probe_out_floats(kind);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* x = in + (size_t)i * ni;
        float* o = out + (size_t)i * no;
        switch (kind) {
        case 0: {
            const f3 v = brdf_cos(mk3(x[3], x[4], x[5]), mk3(x[0], x[1], x[2]), x[6], x[7], mk3(x[8], x[9], x[10]), mk3(x[11], x[12], x[13]));
            o[0] = v.x; o[1] = v.y; o[2] = v.z;
        } break;
        case 1: {
            const f3 v = vndf_sampling(mk3(x[0], x[1], x[2]), mk3(x[3], x[4], x[5]), x[6], x[7], x[8]);
            o[0] = v.x; o[1] = v.y; o[2] = v.z;
        } break;
        case 2: o[0] = vndf_sampling_pdf(mk3(x[0], x[1], x[2]), mk3(x[3], x[4], x[5]), x[6], mk3(x[7], x[8], x[9])); break;
        case 3: {
            Philox4 rr;
            rr.r0 = __float_as_uint(x[10]); rr.r1 = __float_as_uint(x[11]); rr.r2 = __float_as_uint(x[12]); rr.r3 = __float_as_uint(x[13]);
            const f3 v = sample_dir(s, mk3(x[0], x[1], x[2]), mk3(x[3], x[4], x[5]), x[6], mk3(x[7], x[8], x[9]), rr);
            o[0] = v.x; o[1] = v.y; o[2] = v.z;
        } break;
        case 4: {
            const f3 N = mk3(x[0], x[1], x[2]), in_d = mk3(x[7], x[8], x[9]), nd = mk3(x[10], x[11], x[12]);
            o[0] = complete_pdf(s, cosine_weighted_pdf(N, nd), s.n_lights > 0 ? lsum[i] : 0.0f, vndf_pdf_term(s, N, in_d, x[6], nd));
        } break;
        case 5: {
            const int ti = __float_as_int(x[0]);
            const float4 v = texture_sample(ti < 0 ? s.env : s.texs[ti], x[1] != 0.0f, x[2], x[3]);
            o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
        } break;
        case 6: {
            const f3 N = mk3(x[0], x[1], x[2]);
            const f3 v = cosine_weighted(N, __float_as_uint(x[3]), __float_as_uint(x[4]));
            o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = cosine_weighted_pdf(N, v);
        } break;
        default: {
            const float4 v = env_lookup(s, x[0], x[1], x[2]);
            o[0] = v.x; o[1] = v.y; o[2] = v.z;
        } break;
        }
    }
}
__global__ void k_pack_rays(const float* __restrict__ rays6, const uint32_t n, float4* __restrict__ qo,
                            float4* __restrict__ qd, uint32_t* __restrict__ count0) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *count0 = n;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        qo[i] = make_float4(rays6[6 * i], rays6[6 * i + 1], rays6[6 * i + 2], __uint_as_float(i));
        qd[i] = make_float4(rays6[6 * i + 3], rays6[6 * i + 4], rays6[6 * i + 5], 0.0f);
    }
}
// ort_hit = {t, u, v, tri, material, inside}
__global__ void k_unpack_hits(const SceneDev s, const float4* __restrict__ hits, const float4* __restrict__ qd,
                              const uint32_t n, float* __restrict__ out6) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 h = hits[i];
        const int tri = __float_as_int(h.w);
        int mat = -1, inside = 0;
        if (tri >= 0) {
            const float4* ts = s.tshade + (size_t)tri * 4;
            const float4 s0 = ldg4(ts), s1 = ldg4(ts + 1), s2 = ldg4(ts + 2);
            mat = __ldg(reinterpret_cast<const int4*>(ts + 3)).x;
            const float4 d = qd[i];
            inside = addr(addr(mulr(s0.w, d.x), mulr(s1.w, d.y)), mulr(s2.w, d.z)) > 0.0f ? 1 : 0;
        }
        out6[6 * i + 0] = addr(h.x, RAY_EPS); // hit.t += RAY_EPS (raytracer.odin:428)
        out6[6 * i + 1] = h.y;
        out6[6 * i + 2] = h.z;
        out6[6 * i + 3] = __int_as_float(tri);
        out6[6 * i + 4] = __int_as_float(mat);
        out6[6 * i + 5] = __int_as_float(inside);
    }
}
__global__ void k_unpack_rays(const float4* __restrict__ qo, const float4* __restrict__ qd, const uint32_t n,
                              float* __restrict__ out6) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 o = qo[i], d = qd[i];
        out6[6 * i] = o.x; out6[6 * i + 1] = o.y; out6[6 * i + 2] = o.z;
        out6[6 * i + 3] = d.x; out6[6 * i + 4] = d.y; out6[6 * i + 5] = d.z;
    }
}
__global__ void k_scale(float* __restrict__ x, const uint32_t n, const float inv) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] = x[i] / inv;
}

// streaming read of n32 32-byte units, `iters` times (bandwidth probe: L2 when the buffer fits it)
__global__ void __launch_bounds__(256) k_read_bw(const float4* __restrict__ p, const size_t n32, const int iters,
                                                 float* __restrict__ sink) {
    float acc = 0.0f;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (int it = 0; it < iters; it++)
        for (size_t i = i0; i < n32; i += stride) {
            const F8 v = ldg8(p + 2 * i);
            acc += v.lo.x + v.hi.w;
        }
    if (acc == 123.456f) *sink = acc; // never true for the zero-filled buffer: keeps the loads alive
}

// planar accumulators (+ optional first/last planes) -> Sample_Stats AoS (13 words / pixel)
__global__ void k_pack_stats(const float* __restrict__ accum, const float* __restrict__ first,
                             const float* __restrict__ last, const uint32_t npix, uint32_t* __restrict__ out13) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        uint32_t* o = out13 + (size_t)i * 13;
        for (int c = 0; c < 3; c++) {
            o[c] = __float_as_uint(first ? first[c * npix + i] : 0.0f);
            o[4 + c] = __float_as_uint(last ? last[c * npix + i] : 0.0f);
            o[7 + c] = __float_as_uint(accum[c * npix + i]);
            o[10 + c] = __float_as_uint(accum[(3 + c) * npix + i]);
        }
        o[3] = accum_count(accum, npix, i);
    }
}

// Sample_Stats AoS -> planar accumulators + first / last planes (ort_frame_load: a resumed checkpoint)
__global__ void k_unpack_stats(const uint32_t* __restrict__ in13, const uint32_t npix, float* __restrict__ accum,
                               float* __restrict__ first, float* __restrict__ last) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        const uint32_t* o = in13 + (size_t)i * 13;
        for (int c = 0; c < 3; c++) {
            first[c * npix + i] = __uint_as_float(o[c]);
            last[c * npix + i] = __uint_as_float(o[4 + c]);
            accum[c * npix + i] = __uint_as_float(o[7 + c]);
            accum[(3 + c) * npix + i] = __uint_as_float(o[10 + c]);
        }
        accum[6 * npix + i] = (float)(o[3] & 0xfffffu);
        accum[7 * npix + i] = (float)(o[3] >> 20);
    }
}

// get_rgb_image, mode Mean (output.odin:21-80).  The gamma power is evaluated in f64 and rounded once to f32:
// that is the correctly rounded powf(tm, 1/2.2f) (bar double-rounding cases rarer than 1 in 2^27), which is what
// the oracle's glibc powf returns — CUDA's f32 powf (up to 2 ulp off) would flip a byte at rounding boundaries.
__global__ void k_tonemap(const float* __restrict__ accum, const uint32_t npix, uint8_t* __restrict__ rgb) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        const float cnt = (float)accum_count(accum, npix, i);
        for (int c = 0; c < 3; c++) {
            float x = omax(accum[c * npix + i] / cnt, 0.0f);
            float tm = (x * (2.51f * x + 0.03f)) / (x * (2.43f * x + 0.59f) + 0.14f);
            tm = fminf(fmaxf(tm, 0.0f), 1.0f);
            const float g = (float)pow((double)tm, (double)(float)(1.0 / 2.2));
            rgb[3 * (size_t)i + c] = (uint8_t)roundf(g * 255.0f);
        }
    }
}

} // namespace ort
