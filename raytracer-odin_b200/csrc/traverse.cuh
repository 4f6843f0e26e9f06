// traverse.cuh — k_trace: closest hit (cast_ray, raytracer.odin:416-430) and, as a second
// instantiation over the light-candidate queue, the light-BVH all-hit pdf sum of the same ray
// (surface_sampling_pdf_bvh_sum, shading.odin:62-94).
//
// Design notes (every choice below was measured on B200, see profiles/):
//   * Software BVH traversal here is not bandwidth bound (DRAM < 4 %, L2 ~ 20-30 % of peak).  The first
//     version (warp fetches 32 rays, runs until the slowest is done) executed with 6 of 32 lanes active and
//     was issue bound; with the scheduling below the kernel is LATENCY bound at the knee of its occupancy
//     curve: throughput follows the number of node fetches in flight per SM (warps x descending lanes).
//     Denser lanes alone do not help — a warp-level ray pool with dense node / leaf phases was level with
//     this kernel at equal occupancy (profiles/r2_pool_traversal.md).
//   * Persistent threads with PER-LANE dynamic fetch: when fewer than `refill_threshold` lanes of
//     a warp still hold a ray, the warp leaves the traversal loop (in-flight rays keep their state
//     in registers and on the stack), idle lanes claim new rays from the compacted queue with ONE
//     atomic per warp (ballot + prefix popcount), and everybody resumes.
//   * while-while traversal with an early exit from the inner-node loop: lanes that found a leaf
//     wait at the loop's end; once fewer than `inner_min` lanes are still descending, the loop is
//     left so the waiting lanes test their triangles.  (A fully warp-synchronous "vote one step
//     per iteration" variant was tried and was 20 % slower: profiles/r1_traversal_variants.md.)
//   * The light BVH is appended to the scene's node / triangle arrays; k_trace<true> walks it with
//     the same inner loop (no distance culling, no ordering needed) for the all-hit sum.
//   * 4-wide nodes of one 128-byte line read with 3 x LDG.256 (lo | hi planes of an axis) + 1 x LDG.128
//     (children), near / far planes picked in registers; stack entries are 64-bit (node, entry
//     distance): the first SMEM_STACK per thread in shared memory ([entry][thread], conflict free, one
//     LDS.64 / STS.64 per pop / push), deeper ones in thread-local memory (exact worst case checked on
//     the host); popped entries farther than the current best hit are skipped without a node fetch.
//   * What bounds it, measured (profiles/r1_sensitivity.md): ~5 warp-level L2 round trips per ray, each
//     waiting for the slowest of ~13 divergent lanes, at the knee of the occupancy curve (7 CTAs / SM).
//     Extra L1-hit loads are free, +40 % ALU per visit costs 15 %, prefetching and batched triangle loads
//     are slower.  Alternatives that were built, passed parity and lost (8-wide octant-ordered and
//     exact-order nodes, 8-bit quantised nodes, closest hit + light sum fused in one pass) were removed
//     from the tree in round 2; their write-up stays in profiles/r1_sensitivity.md and
//     profiles/r1_traversal_variants.md, their code in the history (commit 1e9815e).
//
// Numerics: box tests are conservative supersets of the reference's (see make_ray); the triangle
// solve is the reference's arithmetic bit for bit (tri_det_t / tri_uv); triangles of a leaf are
// tested in reference order with strict `<`, so the first one wins exact ties inside a leaf.
#pragma once
#include "device_math.cuh"
#include "wide_bvh.h"

namespace ort {

// Dynamic-fetch thresholds, re-tuned in round 2 after the zero-component fix (profiles/r2m_refill_sweep.md): the
// closest-hit pass refills when fewer than 16 lanes hold a ray (C2 +3.3 %, C4 +2.5 % over 22), the all-hit light pass
// keeps 22.
#ifndef ORT_REFILL_THRESHOLD
#define ORT_REFILL_THRESHOLD 16
#endif
#ifndef ORT_REFILL_LIGHT
#define ORT_REFILL_LIGHT 22
#endif
#ifndef ORT_INNER_MIN
#define ORT_INNER_MIN 12
#endif

// Unused child slots carry the box lo = +inf, hi = -inf (host_bvh.cpp): whatever the ray, the near
// plane distance of the x axis is +inf and the far one -inf, so the slab test can never pass and no
// separate validity compare is needed.

// Stack entry = (node reference, entry distance bits): one 64-bit access per push / pop.  The all-hit light pass
// never culls by distance, so its entries are the 32-bit node reference alone (half the shared memory per CTA).
template <bool NODE_ONLY> struct StackEntry { using T = uint2; };
template <> struct StackEntry<true> { using T = uint32_t; };
#ifndef ORT_LIGHT_STACK32
#define ORT_LIGHT_STACK32 1
#endif
#define ORT_PUSH(NODE, DIST)                                                                          \
    {                                                                                                 \
        E e_;                                                                                         \
        if constexpr (sizeof(E) == 4) e_ = (uint32_t)(NODE);                                          \
        else e_ = make_uint2((uint32_t)(NODE), __float_as_uint(DIST));                                \
        if (sp < SMEM_STACK) sh_stack[sp][threadIdx.x] = e_;                                          \
        else l_stack[sp - SMEM_STACK] = e_;                                                           \
        sp++;                                                                                         \
    }
#define ORT_POP(NODE, DIST)                                                                           \
    {                                                                                                 \
        sp--;                                                                                         \
        const E e_ = sp < SMEM_STACK ? sh_stack[sp][threadIdx.x] : l_stack[sp - SMEM_STACK];          \
        if constexpr (sizeof(E) == 4) { NODE = (int)e_; DIST = 0.0f; }                                \
        else { NODE = (int)e_.x; DIST = __uint_as_float(e_.y); }                                      \
    }
#define ORT_CSWAP(da, ca, db, cb)               \
    {                                           \
        const bool sw_ = db < da;               \
        const float td_ = sw_ ? db : da;        \
        const int tc_ = sw_ ? cb : ca;          \
        db = sw_ ? da : db; cb = sw_ ? ca : cb; \
        da = td_; ca = tc_;                     \
    }

// Triangle records are read once per leaf visit: they bypass L1 allocation and leave the L1 to the nodes
// (+0.7 % on C2 and C4; a prefetch of the leaf's first triangle line while the lane waits for the others to finish
// descending was 3-5 % slower: profiles/r2_zero_direction_components.md, last table).
__device__ __forceinline__ F8 ldg8_noalloc(const void* p) {
    F8 r;
    asm("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
        : "l"(p));
    return r;
}
#define ORT_LDG8_TRI(P) ldg8_noalloc(P)

struct TraceArgs {
    const float4* qo;       // ray origins (xyz) + path slot (w), compacted queue order
    const float4* qd;       // ray directions
    const uint32_t* index;  // optional: queue positions to process (light-candidate list: position | root-children
                            // mask << LQ_MASK_SHIFT when index_packed); NULL = 0..n-1
    int index_packed;       // 0: index holds plain positions (waves of more than 2^28 paths)
    const uint32_t* n_ptr;  // number of rays to process (device resident)
    uint32_t* work_ctr;     // persistent-thread work counter
    float4* hits;           // out: (t, u, v, tri)                     (closest hit)
    float* lsum;            // out: light pdf sum                      (light pass)
    int refill_threshold;   // dynamic fetch when fewer lanes than this hold a ray
    int inner_min;          // leave the inner-node loop when fewer lanes than this remain in it
};

// LIGHT = false: closest hit on the scene BVH (every bounce).  LIGHT = true: all-hit pdf sum on the
// light BVH (bounces > 0, only over the light-candidate queue).
template <bool LIGHT>
__global__ void __launch_bounds__(TRACE_THREADS, LIGHT ? ORT_LIGHT_MIN_CTAS : ORT_TRACE_MIN_CTAS)
k_trace(const SceneDev s, const TraceArgs a) {
    using E = typename StackEntry<LIGHT && ORT_LIGHT_STACK32>::T;
    __shared__ E sh_stack[SMEM_STACK][TRACE_THREADS];
    E l_stack[LOCAL_STACK];

    const uint32_t n = *a.n_ptr;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const float best_pad = 1.0f + 7.62939453125e-06f; // 1 + 2^-17: distance culling margin
    const float inf = __int_as_float(0x7f800000);

    RaySetup r;
    float best = inf, hu = 0.0f, hv = 0.0f, lsumv = 0.0f;
    float cull = inf;       // pop / box limit: best * best_pad for the closest hit, +inf for the light sum
    int htri = -1, sp = 0, cur = WIDE_EMPTY;
    uint32_t pos = 0;
    bool exhausted = false; // warp-uniform: the queue has no unclaimed rays left

    for (;;) {
        // ---- refill idle lanes (dynamic fetch)
        const bool idle = cur == WIDE_EMPTY;
        const unsigned idle_mask = __ballot_sync(0xffffffffu, idle);
        if (idle_mask != 0u && !exhausted) {
            const int cnt = __popc(idle_mask);
            const int leader = __ffs(idle_mask) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(a.work_ctr, (uint32_t)cnt);
            base = __shfl_sync(0xffffffffu, base, leader);
            if (idle) {
                uint32_t idx = base + __popc(idle_mask & lt_mask);
                if (idx < n) {
                    uint32_t rmask = 0u;
                    if (a.index) {
                        idx = __ldg(a.index + idx);
                        if (a.index_packed) { rmask = idx >> LQ_MASK_SHIFT; idx &= LQ_POS_MASK; }
                    }
                    r = make_ray(ldg4(a.qo + idx), ldg4(a.qd + idx), s.pad_scale);
                    best = inf; hu = 0.0f; hv = 0.0f; htri = -1; lsumv = 0.0f; // max_dist = +inf (raytracer.odin:435)
                    cull = inf; sp = 0; pos = idx;
                    cur = LIGHT ? s.light_root : 0;
                    if (LIGHT && rmask != 0u) {
                        // k_shade has already tested the root's child boxes for this ray (light_root_mask):
                        // start at the children it enters (any order: the sum needs none)
                        const int4 rc = __ldg(reinterpret_cast<const int4*>(s.nodes + (size_t)s.light_root * 8 + 6));
                        cur = WIDE_EMPTY;
#define ORT_START(BIT, C)                                                  \
    if (rmask & BIT) {                                                     \
        if (cur != WIDE_EMPTY) ORT_PUSH(cur, 0.0f)                         \
        cur = C;                                                           \
    }
                        ORT_START(8u, rc.w) ORT_START(4u, rc.z) ORT_START(2u, rc.y) ORT_START(1u, rc.x)
#undef ORT_START
                    }
                }
            }
            exhausted = base + (uint32_t)cnt >= n;
        }
        if (__ballot_sync(0xffffffffu, cur != WIDE_EMPTY) == 0u) break;

        // ---- traverse
        if (cur != WIDE_EMPTY) {
            for (;;) {
                while (cur >= 0) {
                    // one 128-byte node = 3 x LDG.256 (lo|hi planes of an axis) + 1 x LDG.128 (children):
                    // 4 L1 wavefronts per lane instead of 7; near / far picked in registers
                    const float4* nd = s.nodes + (size_t)cur * 8;
                    const F8 px = ldg8(nd), py = ldg8(nd + 2), pz = ldg8(nd + 4);
                    const int4 ch = __ldg(reinterpret_cast<const int4*>(nd + 6));
                    const bool ngx = r.sx & 1, ngy = r.sy & 1, ngz = r.sz & 1; // direction component negative
#define ORT_SEL4(C, A, B) make_float4(C ? A.x : B.x, C ? A.y : B.y, C ? A.z : B.z, C ? A.w : B.w)
                    const float4 nxp = ORT_SEL4(ngx, px.hi, px.lo), fxp = ORT_SEL4(ngx, px.lo, px.hi);
                    const float4 nyp = ORT_SEL4(ngy, py.hi, py.lo), fyp = ORT_SEL4(ngy, py.lo, py.hi);
                    const float4 nzp = ORT_SEL4(ngz, pz.hi, pz.lo), fzp = ORT_SEL4(ngz, pz.lo, pz.hi);
#undef ORT_SEL4
                    float d0, d1, d2, d3;
                    int c0 = ch.x, c1 = ch.y, c2 = ch.z, c3 = ch.w;
#define ORT_BOX(k, D)                                                                             \
    {                                                                                             \
        const float tn = fmaxf(fmaxf(fmaf(nxp.k, r.ix, r.nx), fmaf(nyp.k, r.iy, r.ny)),           \
                               fmaxf(fmaf(nzp.k, r.iz, r.nz), 0.0f));                             \
        const float tf = fminf(fminf(fmaf(fxp.k, r.ix, r.fx), fmaf(fyp.k, r.iy, r.fy)),           \
                               fminf(fmaf(fzp.k, r.iz, r.fz), cull));                             \
        D = tn <= tf ? tn : inf;                                                                  \
    }
                    ORT_BOX(x, d0) ORT_BOX(y, d1) ORT_BOX(z, d2) ORT_BOX(w, d3)
#undef ORT_BOX
                    const int nh = (d0 < inf) + (d1 < inf) + (d2 < inf) + (d3 < inf);
                    ORT_CSWAP(d0, c0, d1, c1) ORT_CSWAP(d2, c2, d3, c3) ORT_CSWAP(d0, c0, d2, c2)
                    ORT_CSWAP(d1, c1, d3, c3) ORT_CSWAP(d1, c1, d2, c2)
                    if (nh == 0) {
                        cur = WIDE_EMPTY; // pop, skipping entries the current best already culls
                        while (sp > 0) {
                            int nd2; float dd;
                            ORT_POP(nd2, dd)
                            if (dd <= cull) { cur = nd2; break; }
                        }
                    } else {
                        // (three predicated stores at precomputed slots instead of three branches measured
                        // the same: profiles/r1_traversal_variants.md)
                        if (nh > 3) ORT_PUSH(c3, d3)
                        if (nh > 2) ORT_PUSH(c2, d2)
                        if (nh > 1) ORT_PUSH(c1, d1)
                        cur = c0;
                    }
                    // lanes that already hold a leaf wait at the end of this loop: once too few lanes
                    // are still descending, stop and let the waiting lanes test their triangles
                    if (__popc(__activemask()) < a.inner_min) break;
                }
                if (cur < 0 && cur != WIDE_EMPTY) {
                    const uint32_t code = (uint32_t)~cur;
                    const uint32_t first = code >> 3, cnt = code & 7u;
                    if (!LIGHT) {
                        // cast_ray_through_trigs (raytracer.odin:351-369): reference order, first wins ties
                        for (uint32_t i = 0; i < cnt; i++) {
                            const float4* tp = s.tris + (size_t)(first + i) * 4;
                            const F8 tab = ORT_LDG8_TRI(tp);
                            const float4 ta = tab.lo, tb = tab.hi, tc = ldg4(tp + 2);
                            float id, bx, by, bz, t, a00, a10;
                            tri_det_t(r, ta, tb, tc, id, bx, by, bz, t, a00, a10);
                            if (t > 0.0f && t < best) { // raytracer.odin:360
                                float u, v;
                                if (tri_uv(r, ta, tb, tc, id, bx, by, bz, a00, a10, u, v)) {
                                    best = t; hu = u; hv = v; htri = (int)(first + i);
                                    cull = best * best_pad;
                                }
                            }
                        }
                    } else {
                        // surface_sampling_pdf_trigs_sum (shading.odin:52-60): the reference's intersect
                        // returns t = -1 when (u,v) is outside, then `!(t >= 0)` skips.  ng and 2 / |u x v| of a
                        // light triangle sit in the last 16 bytes of its record (TriIsect::light).
#define ORT_LIGHT_TEST(TA, TB, TCL)                                                                   \
    {                                                                                                 \
        float id, bx, by, bz, t, a00, a10, u, v;                                                      \
        tri_det_t(r, TA, TB, TCL.lo, id, bx, by, bz, t, a00, a10);                                    \
        if (t >= 0.0f && tri_uv(r, TA, TB, TCL.lo, id, bx, by, bz, a00, a10, u, v)) {                 \
            const float4 L = TCL.hi;                                                                  \
            const float weight = (t * t) / fabsf(L.x * r.dx + L.y * r.dy + L.z * r.dz);               \
            lsumv += L.w * weight;                                                                    \
        }                                                                                             \
    }
                        for (uint32_t i = 0; i < cnt; i++) {
                            const float4* tp = s.tris + (size_t)(first + i) * 4;
                            const F8 tab = ORT_LDG8_TRI(tp), tcl = ORT_LDG8_TRI(tp + 2);
                            ORT_LIGHT_TEST(tab.lo, tab.hi, tcl)
                        }
#undef ORT_LIGHT_TEST
                    }
                    cur = WIDE_EMPTY;
                    while (sp > 0) {
                        int nd2; float dd;
                        ORT_POP(nd2, dd)
                        if (dd <= cull) { cur = nd2; break; }
                    }
                }
                if (cur == WIDE_EMPTY) { // ray finished
                    if (!LIGHT) a.hits[pos] = make_float4(htri >= 0 ? best : 0.0f, hu, hv, __int_as_float(htri));
                    else a.lsum[pos] = lsumv;
                    break;
                }
                if (!exhausted && __popc(__activemask()) < a.refill_threshold) break; // go refill
            }
        }
    }
}

#undef ORT_PUSH
#undef ORT_POP
#undef ORT_CSWAP

} // namespace ort
