// traverse.cuh — k_trace: closest hit (cast_ray, raytracer.odin:416-430) and, fused behind it for
// continuation rays, the light-BVH all-hit pdf sum of the same ray
// (surface_sampling_pdf_bvh_sum, shading.odin:62-94).
//
// Design notes (every choice below was measured on B200, see profiles/):
//   * Software BVH traversal here is ISSUE bound, not bandwidth bound (DRAM < 4 %, L2 ~ 12 % of
//     peak, issue slots ~ 70 % busy), and its enemy is SIMD divergence: the first version (warp
//     fetches 32 rays, runs until the slowest is done) executed with 6 of 32 lanes active.
//   * Persistent threads with PER-LANE dynamic fetch: when fewer than `refill_threshold` lanes of
//     a warp still hold a ray, the warp leaves the traversal loop (in-flight rays keep their state
//     in registers and on the stack), idle lanes claim new rays from the compacted queue with ONE
//     atomic per warp (ballot + prefix popcount), and everybody resumes.
//   * while-while traversal with an early exit from the inner-node loop: lanes that found a leaf
//     wait at the loop's end; once fewer than `inner_min` lanes are still descending, the loop is
//     left so the waiting lanes test their triangles.  (A fully warp-synchronous "vote one step
//     per iteration" variant was tried and was 20 % slower: profiles/r1_traversal_variants.md.)
//   * The light BVH is appended to the scene's node / triangle arrays.  When a continuation ray's
//     closest-hit stack runs dry the lane switches to phase 1 and walks the light tree with the
//     same inner loop (no distance culling, no ordering needed) for the all-hit sum.
//   * 4-wide nodes of one 128-byte line read with 3 x LDG.256 (lo | hi planes of an axis) + 1 x LDG.128
//     (children), near / far planes picked in registers; stack entries are 64-bit (node, entry
//     distance): the first SMEM_STACK per thread in shared memory ([entry][thread], conflict free, one
//     LDS.64 / STS.64 per pop / push), deeper ones in thread-local memory (exact worst case checked on
//     the host); popped entries farther than the current best hit are skipped without a node fetch.
//   * What bounds it, measured (profiles/r1_sensitivity.md): ~5 warp-level L2 round trips per ray, each
//     waiting for the slowest of ~13 divergent lanes, at the knee of the occupancy curve (7 CTAs / SM).
//     Extra L1-hit loads are free, +40 % ALU per visit costs 15 %, prefetching and batched triangle loads
//     are slower, an 8-wide layout (traverse8.cuh) loses on the exact triangle solve.  What helped last:
//     queue ORDER — a warp's primary rays are 2x2 pixels x 8 samples (k_raygen).
//
// Numerics: box tests are conservative supersets of the reference's (see make_ray); the triangle
// solve is the reference's arithmetic bit for bit (tri_det_t / tri_uv); triangles of a leaf are
// tested in reference order with strict `<`, so the first one wins exact ties inside a leaf.
#pragma once
#include "device_math.cuh"
#include "wide_bvh.h"

namespace ort {

#ifndef ORT_REFILL_THRESHOLD
#define ORT_REFILL_THRESHOLD 22
#endif
#ifndef ORT_INNER_MIN
#define ORT_INNER_MIN 12
#endif

// Unused child slots carry the box lo = +inf, hi = -inf (host_bvh.cpp): whatever the ray, the near
// plane distance of the x axis is +inf and the far one -inf, so the slab test can never pass and no
// separate validity compare is needed (ORT_CHECK_EMPTY restores it).
#ifdef ORT_CHECK_EMPTY
#define ORT_SLOT_OK(HIT, C) ((HIT) && (C) != WIDE_EMPTY)
#else
#define ORT_SLOT_OK(HIT, C) (HIT)
#endif

// Stack entry = (node reference, entry distance bits): one 64-bit access per push / pop.
#define ORT_PUSH(NODE, DIST)                                                                          \
    {                                                                                                 \
        const uint2 e_ = make_uint2((uint32_t)(NODE), __float_as_uint(DIST));                         \
        if (sp < SMEM_STACK) sh_stack[sp][threadIdx.x] = e_;                                          \
        else l_stack[sp - SMEM_STACK] = e_;                                                           \
        sp++;                                                                                         \
    }
#define ORT_POP(NODE, DIST)                                                                           \
    {                                                                                                 \
        sp--;                                                                                         \
        const uint2 e_ = sp < SMEM_STACK ? sh_stack[sp][threadIdx.x] : l_stack[sp - SMEM_STACK];      \
        NODE = (int)e_.x; DIST = __uint_as_float(e_.y);                                               \
    }
#define ORT_CSWAP(da, ca, db, cb)               \
    {                                           \
        const bool sw_ = db < da;               \
        const float td_ = sw_ ? db : da;        \
        const int tc_ = sw_ ? cb : ca;          \
        db = sw_ ? da : db; cb = sw_ ? ca : cb; \
        da = td_; ca = tc_;                     \
    }

struct TraceArgs {
    const float4* qo;       // ray origins (xyz) + path slot (w), compacted queue order
    const float4* qd;       // ray directions
    const uint32_t* index;  // optional: queue positions to process (light-candidate list); NULL = 0..n-1
    const uint32_t* n_ptr;  // number of rays to process (device resident)
    uint32_t* work_ctr;     // persistent-thread work counter
    float4* hits;           // out: (t, u, v, tri)
    float* lsum;            // out: light pdf sum (only when do_light)
    int refill_threshold;   // dynamic fetch when fewer lanes than this hold a ray
    int inner_min;          // leave the inner-node loop when fewer lanes than this remain in it
};

// CLOSEST / LIGHT select the phases compiled in: <true,false> closest hit only (what render uses for
// every bounce), <false,true> light sum only (render, bounces > 0), <true,true> both fused in one
// pass (kept for comparison: 77 registers and phase-divergent leaf code make it 15 % slower than
// the two specialised launches, profiles/r1_traversal_variants.md).
//
// QUANT selects the node encoding: WideNode (f32 planes, 7 x LDG.128 per visit) or QuantNode (8-bit
// planes, 4 x LDG.128 per visit, ~35 % more ALU per visit).  Small scenes are issue bound and run
// faster on WideNode; large scenes are bound by the L1/TEX pipe and run faster on QuantNode.
template <bool CLOSEST, bool LIGHT, int QUANT>
__global__ void __launch_bounds__(TRACE_THREADS, ORT_TRACE_MIN_CTAS)
k_trace(const SceneDev s, const TraceArgs a) {
    __shared__ uint2 sh_stack[SMEM_STACK][TRACE_THREADS];
    uint2 l_stack[LOCAL_STACK];

    const uint32_t n = *a.n_ptr;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const float best_pad = 1.0f + 7.62939453125e-06f; // 1 + 2^-17: distance culling margin
    const float inf = __int_as_float(0x7f800000);

    RaySetup r;
    float best = inf, hu = 0.0f, hv = 0.0f, lsumv = 0.0f;
    float cull = inf;       // pop / box limit: best * best_pad in phase 0, +inf in phase 1
    int htri = -1, sp = 0, cur = WIDE_EMPTY;
    int phase = 0;          // 0: closest hit on the scene BVH, 1: all-hit sum on the light BVH
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
                    if (a.index) idx = __ldg(a.index + idx);
                    r = make_ray(ldg4(a.qo + idx), ldg4(a.qd + idx), s.pad_scale);
                    best = inf; hu = 0.0f; hv = 0.0f; htri = -1; lsumv = 0.0f; // max_dist = +inf (raytracer.odin:435)
                    cull = inf; sp = 0; pos = idx;
                    if (!CLOSEST) { phase = 1; cur = QUANT == 2 ? s.light_root8x : s.light_root; }
                    else { phase = 0; cur = 0; }
                }
            }
            exhausted = base + (uint32_t)cnt >= n;
        }
        if (__ballot_sync(0xffffffffu, cur != WIDE_EMPTY) == 0u) break;

        // ---- traverse
        if (cur != WIDE_EMPTY) {
            for (;;) {
                while (cur >= 0) {
                    if (QUANT == 2) {
                        // ---- exact-order 8-wide visit (Wide8xNode): eight slab tests, a 19-comparator sorting
                        // network on (entry distance, child), far ... near pushed with their distances
                        const float4* nd = s.nodes8x + (size_t)cur * 16;
                        const int ox = (r.sx & 1) * 2, oy = (r.sy & 1) * 2, oz = (r.sz & 1) * 2;
                        const F8 nxp = ldg8(nd + ox), fxp = ldg8(nd + (ox ^ 2));
                        const F8 nyp = ldg8(nd + 4 + oy), fyp = ldg8(nd + 4 + (oy ^ 2));
                        const F8 nzp = ldg8(nd + 8 + oz), fzp = ldg8(nd + 8 + (oz ^ 2));
                        const F8 chf = ldg8(nd + 12);
                        float e0, e1, e2, e3, e4, e5, e6, e7;
                        int k0 = __float_as_int(chf.lo.x), k1 = __float_as_int(chf.lo.y), k2 = __float_as_int(chf.lo.z),
                            k3 = __float_as_int(chf.lo.w), k4 = __float_as_int(chf.hi.x), k5 = __float_as_int(chf.hi.y),
                            k6 = __float_as_int(chf.hi.z), k7 = __float_as_int(chf.hi.w);
#define ORT_BOX8X(H, k, D)                                                                            \
    {                                                                                                 \
        const float tn = fmaxf(fmaxf(fmaf(nxp.H.k, r.ix, r.nx), fmaf(nyp.H.k, r.iy, r.ny)),           \
                               fmaxf(fmaf(nzp.H.k, r.iz, r.nz), 0.0f));                               \
        const float tf = fminf(fminf(fmaf(fxp.H.k, r.ix, r.fx), fmaf(fyp.H.k, r.iy, r.fy)),           \
                               fminf(fmaf(fzp.H.k, r.iz, r.fz), cull));                               \
        D = tn <= tf ? tn : inf;                                                                      \
    }
                        ORT_BOX8X(lo, x, e0) ORT_BOX8X(lo, y, e1) ORT_BOX8X(lo, z, e2) ORT_BOX8X(lo, w, e3)
                        ORT_BOX8X(hi, x, e4) ORT_BOX8X(hi, y, e5) ORT_BOX8X(hi, z, e6) ORT_BOX8X(hi, w, e7)
#undef ORT_BOX8X
                        const int nh8 = (e0 < inf) + (e1 < inf) + (e2 < inf) + (e3 < inf) + (e4 < inf) + (e5 < inf) + (e6 < inf) + (e7 < inf);
                        // Batcher odd-even merge sort, 19 comparators (misses carry +inf and sink to the end)
                        ORT_CSWAP(e0, k0, e1, k1) ORT_CSWAP(e2, k2, e3, k3) ORT_CSWAP(e0, k0, e2, k2) ORT_CSWAP(e1, k1, e3, k3) ORT_CSWAP(e1, k1, e2, k2)
                        ORT_CSWAP(e4, k4, e5, k5) ORT_CSWAP(e6, k6, e7, k7) ORT_CSWAP(e4, k4, e6, k6) ORT_CSWAP(e5, k5, e7, k7) ORT_CSWAP(e5, k5, e6, k6)
                        ORT_CSWAP(e0, k0, e4, k4) ORT_CSWAP(e1, k1, e5, k5) ORT_CSWAP(e2, k2, e6, k6) ORT_CSWAP(e3, k3, e7, k7)
                        ORT_CSWAP(e2, k2, e4, k4) ORT_CSWAP(e3, k3, e5, k5)
                        ORT_CSWAP(e1, k1, e2, k2) ORT_CSWAP(e3, k3, e4, k4) ORT_CSWAP(e5, k5, e6, k6)
                        if (nh8 == 0) {
                            cur = WIDE_EMPTY;
                            while (sp > 0) {
                                int nd2; float dd;
                                ORT_POP(nd2, dd)
                                if (dd <= cull) { cur = nd2; break; }
                            }
                        } else {
                            if (nh8 > 7) ORT_PUSH(k7, e7)
                            if (nh8 > 6) ORT_PUSH(k6, e6)
                            if (nh8 > 5) ORT_PUSH(k5, e5)
                            if (nh8 > 4) ORT_PUSH(k4, e4)
                            if (nh8 > 3) ORT_PUSH(k3, e3)
                            if (nh8 > 2) ORT_PUSH(k2, e2)
                            if (nh8 > 1) ORT_PUSH(k1, e1)
                            cur = k0;
                        }
                        if (__popc(__activemask()) < a.inner_min) break;
                        continue;
                    }
                    float d0, d1, d2, d3;
                    int c0, c1, c2, c3;
                    if (!QUANT) {
                        // one 128-byte node = 3 x LDG.256 (lo|hi planes of an axis) + 1 x LDG.128 (children):
                        // 4 L1 wavefronts per lane instead of 7; near / far picked in registers
                        const float4* nd = s.nodes + (size_t)cur * 8;
                        const F8 px = ldg8(nd), py = ldg8(nd + 2), pz = ldg8(nd + 4);
                        int4 ch = __ldg(reinterpret_cast<const int4*>(nd + 6));
#ifdef ORT_EXP_DUPLOAD
                        // sensitivity experiment: the same 128-byte node read a second time (L1 hits, but the
                        // wavefronts go through the LSU data pipe again); results folded in so nothing is elided
                        {
                            float x0, x1, x2, x3, x4, x5, x6, x7;
                            unsigned acc = 0;
#pragma unroll
                            for (int kk = 0; kk < ORT_EXP_DUPLOAD; kk++) {
                                asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                             : "=f"(x0), "=f"(x1), "=f"(x2), "=f"(x3), "=f"(x4), "=f"(x5), "=f"(x6), "=f"(x7)
                                             : "l"(nd + 2 * (kk & 3)));
                                acc |= __float_as_uint(x0) & __float_as_uint(x7) & 0x80000000u & (unsigned)cur;
                            }
                            ch.x |= (int)(acc & (acc >> 1) & 1u); // always 0 (bit 0 of a value with only bit 31 set)
                        }
#endif
                        const bool ngx = r.sx & 1, ngy = r.sy & 1, ngz = r.sz & 1; // direction component negative
#define ORT_SEL4(C, A, B) make_float4(C ? A.x : B.x, C ? A.y : B.y, C ? A.z : B.z, C ? A.w : B.w)
                        const float4 nxp = ORT_SEL4(ngx, px.hi, px.lo), fxp = ORT_SEL4(ngx, px.lo, px.hi);
                        const float4 nyp = ORT_SEL4(ngy, py.hi, py.lo), fyp = ORT_SEL4(ngy, py.lo, py.hi);
                        const float4 nzp = ORT_SEL4(ngz, pz.hi, pz.lo), fzp = ORT_SEL4(ngz, pz.lo, pz.hi);
#undef ORT_SEL4
                        c0 = ch.x; c1 = ch.y; c2 = ch.z; c3 = ch.w;
#define ORT_BOX(k, D, C)                                                                          \
    {                                                                                             \
        const float tn = fmaxf(fmaxf(fmaf(nxp.k, r.ix, r.nx), fmaf(nyp.k, r.iy, r.ny)),           \
                               fmaxf(fmaf(nzp.k, r.iz, r.nz), 0.0f));                             \
        const float tf = fminf(fminf(fmaf(fxp.k, r.ix, r.fx), fmaf(fyp.k, r.iy, r.fy)),           \
                               fminf(fmaf(fzp.k, r.iz, r.fz), cull));                             \
        D = ORT_SLOT_OK(tn <= tf, C) ? tn : inf;                                                  \
    }
                        ORT_BOX(x, d0, c0) ORT_BOX(y, d1, c1) ORT_BOX(z, d2, c2) ORT_BOX(w, d3, c3)
#ifdef ORT_EXP_DUPALU
                        // sensitivity experiment: ORT_EXP_DUPALU extra dependent FMAs per visit on the critical path
                        {
#ifdef ORT_EXP_INDEP
                            float z0 = d0, z1 = d1, z2 = d2, z3 = d3; // four independent chains: issue slots, little latency
#pragma unroll
                            for (int kk = 0; kk < ORT_EXP_DUPALU / 4; kk++) {
                                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(z0) : "f"(r.ix), "f"(r.nx));
                                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(z1) : "f"(r.ix), "f"(r.nx));
                                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(z2) : "f"(r.ix), "f"(r.nx));
                                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(z3) : "f"(r.ix), "f"(r.nx));
                            }
                            const float z = z0 + z1 + z2 + z3;
#else
                            float z = d0;
#pragma unroll
                            for (int kk = 0; kk < ORT_EXP_DUPALU; kk++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(z) : "f"(r.ix), "f"(r.nx));
#endif
                            if (z == 123.456f) d1 = z;
                        }
#endif
#undef ORT_BOX
                    } else {
                        // 64-byte node = 2 x LDG.256
                        const float4* nd = s.nodes + (size_t)cur * 4;
                        const F8 q01 = ldg8(nd), q23 = ldg8(nd + 2);
                        const uint4 v0 = make_uint4(__float_as_uint(q01.lo.x), __float_as_uint(q01.lo.y), __float_as_uint(q01.lo.z), __float_as_uint(q01.lo.w));
                        const uint4 v1 = make_uint4(__float_as_uint(q01.hi.x), __float_as_uint(q01.hi.y), __float_as_uint(q01.hi.z), __float_as_uint(q01.hi.w));
                        const uint4 v2 = make_uint4(__float_as_uint(q23.lo.x), __float_as_uint(q23.lo.y), 0u, 0u);
                        c0 = __float_as_int(q23.hi.x); c1 = __float_as_int(q23.hi.y); c2 = __float_as_int(q23.hi.z); c3 = __float_as_int(q23.hi.w);
                        // plane = origin + q * step.  q is spliced into the mantissa of 1.0f (one PRMT):
                        // f = 1 + q * 2^-15, so  t = f * (2^15 step / d) + ((origin - o) / d -+ pad - 2^15 step / d)
                        const float ax_ = __uint_as_float(((v0.w & 0xffu) + 15u) << 23) * r.ix;
                        const float ay_ = __uint_as_float((((v0.w >> 8) & 0xffu) + 15u) << 23) * r.iy;
                        const float az_ = __uint_as_float((((v0.w >> 16) & 0xffu) + 15u) << 23) * r.iz;
                        const float ox_ = __uint_as_float(v0.x), oy_ = __uint_as_float(v0.y), oz_ = __uint_as_float(v0.z);
                        const float bnx = fmaf(ox_, r.ix, r.nx) - ax_, bfx = fmaf(ox_, r.ix, r.fx) - ax_;
                        const float bny = fmaf(oy_, r.iy, r.ny) - ay_, bfy = fmaf(oy_, r.iy, r.fy) - ay_;
                        const float bnz = fmaf(oz_, r.iz, r.nz) - az_, bfz = fmaf(oz_, r.iz, r.fz) - az_;
                        const bool fx_ = r.sx & 1, fy_ = r.sy & 1, fz_ = r.sz & 1; // direction component negative
                        const uint32_t nxw = fx_ ? v1.y : v1.x, fxw = fx_ ? v1.x : v1.y;
                        const uint32_t nyw = fy_ ? v1.w : v1.z, fyw = fy_ ? v1.z : v1.w;
                        const uint32_t nzw = fz_ ? v2.y : v2.x, fzw = fz_ ? v2.x : v2.y;
#define ORT_QF(W, SEL) __uint_as_float(__byte_perm(W, 0x3F800000u, SEL))
#define ORT_BOX(SEL, D, C)                                                                        \
    {                                                                                             \
        const float tn = fmaxf(fmaxf(fmaf(ORT_QF(nxw, SEL), ax_, bnx), fmaf(ORT_QF(nyw, SEL), ay_, bny)), \
                               fmaxf(fmaf(ORT_QF(nzw, SEL), az_, bnz), 0.0f));                    \
        const float tf = fminf(fminf(fmaf(ORT_QF(fxw, SEL), ax_, bfx), fmaf(ORT_QF(fyw, SEL), ay_, bfy)), \
                               fminf(fmaf(ORT_QF(fzw, SEL), az_, bfz), cull));                    \
        D = (tn <= tf && C != WIDE_EMPTY) ? tn : inf;                                             \
    }
                        ORT_BOX(0x7604u, d0, c0) ORT_BOX(0x7614u, d1, c1) ORT_BOX(0x7624u, d2, c2) ORT_BOX(0x7634u, d3, c3)
#undef ORT_BOX
#undef ORT_QF
                    }
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
                    if (CLOSEST && (!LIGHT || phase == 0)) {
                        // cast_ray_through_trigs (raytracer.odin:351-369): reference order, first wins ties
                        for (uint32_t i = 0; i < cnt; i++) {
                            const float4* tp = s.tris + (size_t)(first + i) * 4;
                            const F8 tab = ldg8(tp);
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
                        // returns t = -1 when (u,v) is outside, then `!(t >= 0)` skips
                        for (uint32_t i = 0; i < cnt; i++) {
                            const float4* tp = s.tris + (size_t)(first + i) * 4;
                            const F8 tab = ldg8(tp);
                            const float4 ta = tab.lo, tb = tab.hi, tc = ldg4(tp + 2);
                            float id, bx, by, bz, t, a00, a10, u, v;
                            tri_det_t(r, ta, tb, tc, id, bx, by, bz, t, a00, a10);
                            if (t >= 0.0f && tri_uv(r, ta, tb, tc, id, bx, by, bz, a00, a10, u, v)) {
                                const float4 L = ldg4(s.llight + (first + i - s.light_tri_base));
                                const float weight = (t * t) / fabsf(L.x * r.dx + L.y * r.dy + L.z * r.dz);
                                lsumv += L.w * weight;
                            }
                        }
                    }
                    cur = WIDE_EMPTY;
                    while (sp > 0) {
                        int nd2; float dd;
                        ORT_POP(nd2, dd)
                        if (dd <= cull) { cur = nd2; break; }
                    }
                }
                if (cur == WIDE_EMPTY) {
                    if (CLOSEST && LIGHT && phase == 0) {
                        // closest hit known; now the light-BVH all-hit sum of the same ray
                        phase = 1; cull = inf; cur = s.light_root;
                    } else {
                        if (CLOSEST) a.hits[pos] = make_float4(htri >= 0 ? best : 0.0f, hu, hv, __int_as_float(htri));
                        if (LIGHT) a.lsum[pos] = lsumv;
                        break; // ray finished
                    }
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
