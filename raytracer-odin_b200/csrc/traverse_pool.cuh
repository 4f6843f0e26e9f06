// traverse_pool.cuh — k_trace_pool: the same traversal as k_trace (traverse.cuh: same nodes, same box
// tests, same child order, same exact triangle solve, same results bit for bit), scheduled differently.
//
// k_trace keeps one ray per lane for the ray's whole life, so a warp always holds a mix of lanes that
// descend inner nodes, lanes that wait to test a leaf and lanes that wait for a refill: ncu shows 13-15 of
// 32 lanes active per instruction on bounce rays (profiles/r1t_*).  Here a warp owns a POOL of 64 ray slots
// in shared memory (two per lane) and alternates between two dense phases:
//   * node phase: up to 32 slots that need an inner-node visit are gathered — one per lane, state in
//     registers — and run the classic inner loop until fewer than `inner_min` of them are still descending;
//     then cur / sp are scattered back;
//   * leaf phase: every (slot, triangle) pair of every slot waiting at a leaf becomes one work item, lanes
//     take 32 items at a time, run the reference's exact solve, and the winner of a slot is found with a
//     64-bit atomicMin on (t bits, index in leaf) — strict `<` against the slot's previous best, lowest index
//     wins equal t, exactly cast_ray_through_trigs' rule (raytracer.odin:351-369).  The slots' owner lanes
//     then pop their next node.
// A slot that finishes is refilled from the ray queue (one atomic per warp and refill, as in k_trace).
// The order in which a ray visits nodes and leaves depends only on its own stack and its own best hit, so
// results do not depend on the scheduling: hits are bit-identical to k_trace's, the light sums too (a leaf's
// contributions are added in leaf order by the owner lane).
#pragma once
#include "device_math.cuh"
#include "wide_bvh.h"

namespace ort {

constexpr int POOL = 64;                          // ray slots per warp
constexpr int POOL_OVF = MAX_STACK;               // entries beyond the shared-memory part: global scratch, [entry][slot] per warp
constexpr int POOL_WARPS = TRACE_THREADS / 32;
constexpr int POOL_LEAF_MAX = 7;                  // triangles per leaf the node encoding allows (the reference builds <= 4)

template <bool LIGHT>
struct PoolWarp {                // shared memory of one warp
    static constexpr int SD = LIGHT ? 8 : 12;     // stack entries per slot kept in shared memory
    float4 o[POOL];              // origin after the RAY_EPS offset | best t (closest) or pdf sum (light)
    float4 d[POOL];              // direction | cull distance
    float4 i[POOL];              // 1/d | direction sign bits (x: 1, y: 2, z: 4)
    float4 n[POOL];              // near-plane offsets | queue position
    float4 f[POOL];              // far-plane offsets | hit triangle
    float2 uv[POOL];
    int2 cs[POOL];               // cur (node >= 0, leaf < 0, WIDE_EMPTY = free slot) | stack pointer
    unsigned long long key[POOL];
    uint2 stack[SD][POOL];
    float contrib[LIGHT ? POOL * POOL_LEAF_MAX : 4]; // light pass: per-item contributions of the current leaf phase
    uint16_t items[POOL * POOL_LEAF_MAX];            // (slot << 3) | index in leaf
    uint8_t list[POOL];
};

struct PoolArgs {
    uint2* overflow;   // POOL * POOL_OVF entries per warp of the grid
    int refill_min;    // refill when at least this many slots are free
    int node_min;      // run a node phase when at least this many slots wait for one (or no slot waits at a leaf)
    int inner_min;     // leave the inner-node loop when fewer lanes than this are still descending
};

#define ORT_PPUSH(NODE, DIST)                                                                         \
    {                                                                                                 \
        const uint2 e_ = make_uint2((uint32_t)(NODE), __float_as_uint(DIST));                         \
        if (sp < SD) W.stack[sp][slot] = e_;                                                          \
        else ovf[(size_t)(sp - SD) * POOL + slot] = e_;                                               \
        sp++;                                                                                         \
    }
#define ORT_PPOP(NODE, DIST)                                                                          \
    {                                                                                                 \
        sp--;                                                                                         \
        const uint2 e_ = sp < SD ? W.stack[sp][slot] : ovf[(size_t)(sp - SD) * POOL + slot];          \
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

template <bool LIGHT>
__global__ void __launch_bounds__(TRACE_THREADS, 4)
k_trace_pool(const SceneDev s, const TraceArgs a, const PoolArgs pa) {
    extern __shared__ __align__(16) unsigned char pool_smem[];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    PoolWarp<LIGHT>& W = reinterpret_cast<PoolWarp<LIGHT>*>(pool_smem)[wid];
    constexpr int SD = PoolWarp<LIGHT>::SD;
    uint2* const ovf = pa.overflow + (size_t)(blockIdx.x * POOL_WARPS + wid) * POOL * POOL_OVF;

    const uint32_t n = *a.n_ptr;
    const unsigned lt_mask = (1u << lane) - 1u;
    const float best_pad = 1.0f + 7.62939453125e-06f; // 1 + 2^-17: distance culling margin
    const float inf = __int_as_float(0x7f800000);

    W.cs[lane] = make_int2(WIDE_EMPTY, 0);
    W.cs[lane + 32] = make_int2(WIDE_EMPTY, 0);
    __syncwarp();
    bool exhausted = false; // warp-uniform: the queue has no unclaimed rays left
    uint32_t round = 0;

    // a finished ray: the closest hit / the light sum goes out, the slot becomes free
    auto finish = [&](int slot) {
        const uint32_t pos = __float_as_uint(W.n[slot].w);
        if (!LIGHT) {
            const int htri = __float_as_int(W.f[slot].w);
            const float2 uv = W.uv[slot];
            a.hits[pos] = make_float4(htri >= 0 ? W.o[slot].w : 0.0f, uv.x, uv.y, __int_as_float(htri));
        } else {
            a.lsum[pos] = W.o[slot].w;
        }
    };
    auto init_slot = [&](int slot, uint32_t idx) {
        if (a.index) idx = __ldg(a.index + idx);
        const RaySetup r = make_ray(ldg4(a.qo + idx), ldg4(a.qd + idx), s.pad_scale);
        W.o[slot] = make_float4(r.ox, r.oy, r.oz, LIGHT ? 0.0f : inf); // max_dist = +inf (raytracer.odin:435)
        W.d[slot] = make_float4(r.dx, r.dy, r.dz, inf);
        W.i[slot] = make_float4(r.ix, r.iy, r.iz, __int_as_float((r.sx & 1) | ((r.sy & 1) << 1) | ((r.sz & 1) << 2)));
        W.n[slot] = make_float4(r.nx, r.ny, r.nz, __uint_as_float(idx));
        W.f[slot] = make_float4(r.fx, r.fy, r.fz, __int_as_float(-1));
        W.uv[slot] = make_float2(0.0f, 0.0f);
    };

    for (;; round++) {
        int2 cs0 = W.cs[lane], cs1 = W.cs[lane + 32];
        // ---- refill free slots (dynamic fetch: one atomic per warp)
        if (!exhausted) {
            const unsigned e0 = __ballot_sync(0xffffffffu, cs0.x == WIDE_EMPTY);
            const unsigned e1 = __ballot_sync(0xffffffffu, cs1.x == WIDE_EMPTY);
            const int ne = __popc(e0) + __popc(e1);
            if (ne >= pa.refill_min) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(a.work_ctr, (uint32_t)ne);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (cs0.x == WIDE_EMPTY) {
                    const uint32_t idx = base + __popc(e0 & lt_mask);
                    if (idx < n) { init_slot(lane, idx); cs0 = make_int2(LIGHT ? s.light_root : 0, 0); W.cs[lane] = cs0; }
                }
                if (cs1.x == WIDE_EMPTY) {
                    const uint32_t idx = base + __popc(e0) + __popc(e1 & lt_mask);
                    if (idx < n) { init_slot(lane + 32, idx); cs1 = make_int2(LIGHT ? s.light_root : 0, 0); W.cs[lane + 32] = cs1; }
                }
                exhausted = base + (uint32_t)ne >= n;
            }
        }
        const bool node0 = cs0.x >= 0, node1 = cs1.x >= 0;
        const bool leaf0 = cs0.x < 0 && cs0.x != WIDE_EMPTY, leaf1 = cs1.x < 0 && cs1.x != WIDE_EMPTY;
        const unsigned n0 = __ballot_sync(0xffffffffu, node0), n1 = __ballot_sync(0xffffffffu, node1);
        const unsigned l0 = __ballot_sync(0xffffffffu, leaf0), l1 = __ballot_sync(0xffffffffu, leaf1);
        const int nn = __popc(n0) + __popc(n1), nl = __popc(l0) + __popc(l1);
        if (nn + nl == 0) {
            if (exhausted) break;
            continue;
        }

        if (nn >= pa.node_min || nl == 0) {
            // ================================ node phase ================================
            // gather list: alternate which half of the pool comes first so no slot starves
            const bool flip = round & 1u;
            const unsigned na = flip ? n1 : n0, nb = flip ? n0 : n1;
            if (flip ? node1 : node0) W.list[__popc(na & lt_mask)] = (uint8_t)(flip ? lane + 32 : lane);
            if (flip ? node0 : node1) W.list[__popc(na) + __popc(nb & lt_mask)] = (uint8_t)(flip ? lane : lane + 32);
            __syncwarp();
            if (lane < nn) {
                const int slot = W.list[lane];
                const float4 O = W.o[slot], D = W.d[slot], I = W.i[slot], N = W.n[slot], F = W.f[slot];
                const int2 cs = W.cs[slot];
                int cur = cs.x, sp = cs.y;
                const float cull = D.w;
                const int sgn = __float_as_int(I.w);
                const bool ngx = sgn & 1, ngy = sgn & 2, ngz = sgn & 4; // direction component negative
                (void)O;
                while (cur >= 0) {
                    const float4* nd = s.nodes + (size_t)cur * 8;
                    const F8 px = ldg8(nd), py = ldg8(nd + 2), pz = ldg8(nd + 4);
                    const int4 ch = __ldg(reinterpret_cast<const int4*>(nd + 6));
#define ORT_SEL4(C, A, B) make_float4(C ? A.x : B.x, C ? A.y : B.y, C ? A.z : B.z, C ? A.w : B.w)
                    const float4 nxp = ORT_SEL4(ngx, px.hi, px.lo), fxp = ORT_SEL4(ngx, px.lo, px.hi);
                    const float4 nyp = ORT_SEL4(ngy, py.hi, py.lo), fyp = ORT_SEL4(ngy, py.lo, py.hi);
                    const float4 nzp = ORT_SEL4(ngz, pz.hi, pz.lo), fzp = ORT_SEL4(ngz, pz.lo, pz.hi);
#undef ORT_SEL4
                    float d0, d1, d2, d3;
                    int c0 = ch.x, c1 = ch.y, c2 = ch.z, c3 = ch.w;
#define ORT_BOX(k, DD)                                                                            \
    {                                                                                             \
        const float tn = fmaxf(fmaxf(fmaf(nxp.k, I.x, N.x), fmaf(nyp.k, I.y, N.y)),               \
                               fmaxf(fmaf(nzp.k, I.z, N.z), 0.0f));                               \
        const float tf = fminf(fminf(fmaf(fxp.k, I.x, F.x), fmaf(fyp.k, I.y, F.y)),               \
                               fminf(fmaf(fzp.k, I.z, F.z), cull));                               \
        DD = tn <= tf ? tn : inf;                                                                 \
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
                            ORT_PPOP(nd2, dd)
                            if (dd <= cull) { cur = nd2; break; }
                        }
                    } else {
                        if (nh > 3) ORT_PPUSH(c3, d3)
                        if (nh > 2) ORT_PPUSH(c2, d2)
                        if (nh > 1) ORT_PPUSH(c1, d1)
                        cur = c0;
                    }
                    if (__popc(__activemask()) < pa.inner_min) break;
                }
                if (cur == WIDE_EMPTY) finish(slot);
                W.cs[slot] = make_int2(cur, sp);
            }
            __syncwarp();
        } else {
            // ================================ leaf phase ================================
            const int k0 = leaf0 ? (int)((uint32_t)~cs0.x & 7u) : 0, k1 = leaf1 ? (int)((uint32_t)~cs1.x & 7u) : 0;
            int x0 = k0, x1 = k1; // inclusive scans over the lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y0 = __shfl_up_sync(0xffffffffu, x0, o), y1 = __shfl_up_sync(0xffffffffu, x1, o);
                if (lane >= o) { x0 += y0; x1 += y1; }
            }
            const int tot0 = __shfl_sync(0xffffffffu, x0, 31), tot1 = __shfl_sync(0xffffffffu, x1, 31);
            const int off0 = x0 - k0, off1 = tot0 + x1 - k1, T = tot0 + tot1;
            for (int k = 0; k < k0; k++) W.items[off0 + k] = (uint16_t)((lane << 3) | k);
            for (int k = 0; k < k1; k++) W.items[off1 + k] = (uint16_t)(((lane + 32) << 3) | k);
            if (!LIGHT) {
                if (leaf0) W.key[lane] = ((unsigned long long)__float_as_uint(W.o[lane].w) << 32) | 0xffffffffull;
                if (leaf1) W.key[lane + 32] = ((unsigned long long)__float_as_uint(W.o[lane + 32].w) << 32) | 0xffffffffull;
            }
            __syncwarp();
            for (int b0 = 0; b0 < T; b0 += 32) {
                const int b = b0 + lane;
                const bool act = b < T;
                int slot = 0;
                unsigned long long mykey = ~0ull;
                float u = 0.0f, v = 0.0f;
                bool cand = false;
                if (act) {
                    const int item = W.items[b];
                    slot = item >> 3;
                    const float4 O = W.o[slot], D = W.d[slot];
                    const uint32_t first = ((uint32_t)~W.cs[slot].x) >> 3;
                    const uint32_t tri = first + (uint32_t)(item & 7);
                    RaySetup r;
                    r.ox = O.x; r.oy = O.y; r.oz = O.z; r.dx = D.x; r.dy = D.y; r.dz = D.z;
                    const float4* tp = s.tris + (size_t)tri * 4;
                    const F8 tab = ldg8(tp);
                    const float4 ta = tab.lo, tb = tab.hi, tc = ldg4(tp + 2);
                    float id, bx, by, bz, t, a00, a10;
                    tri_det_t(r, ta, tb, tc, id, bx, by, bz, t, a00, a10);
                    if (!LIGHT) {
                        if (t > 0.0f && t < O.w && tri_uv(r, ta, tb, tc, id, bx, by, bz, a00, a10, u, v)) { // raytracer.odin:360
                            cand = true;
                            mykey = ((unsigned long long)__float_as_uint(t) << 32) | (unsigned long long)(item & 7);
                            atomicMin(&W.key[slot], mykey);
                        }
                    } else {
                        // surface_sampling_pdf_trigs_sum (shading.odin:52-60)
                        float c = 0.0f;
                        if (t >= 0.0f && tri_uv(r, ta, tb, tc, id, bx, by, bz, a00, a10, u, v)) {
                            const float4 L = ldg4(s.llight + (tri - s.light_tri_base));
                            const float weight = (t * t) / fabsf(L.x * r.dx + L.y * r.dy + L.z * r.dz);
                            c = L.w * weight;
                        }
                        W.contrib[b] = c;
                    }
                }
                if (!LIGHT) {
                    __syncwarp();
                    if (cand && W.key[slot] == mykey) W.uv[slot] = make_float2(u, v); // a later batch may still beat it
                    __syncwarp();
                }
            }
            __syncwarp();
            // ---- owners: commit the leaf's result, pop the next node
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const bool isleaf = h ? leaf1 : leaf0;
                if (!isleaf) continue;
                const int slot = lane + 32 * h;
                int2 cs = h ? cs1 : cs0;
                const uint32_t first = ((uint32_t)~cs.x) >> 3;
                float cull = inf;
                if (!LIGHT) {
                    const unsigned long long key = W.key[slot];
                    if ((uint32_t)key != 0xffffffffu) {
                        const float best = __uint_as_float((uint32_t)(key >> 32));
                        W.o[slot].w = best;
                        W.d[slot].w = best * best_pad;
                        W.f[slot].w = __int_as_float((int)(first + (uint32_t)key));
                    }
                    cull = W.d[slot].w;
                } else {
                    const int off = h ? off1 : off0, cnt = h ? k1 : k0;
                    float lsumv = W.o[slot].w;
                    for (int k = 0; k < cnt; k++) lsumv += W.contrib[off + k];
                    W.o[slot].w = lsumv;
                }
                int sp = cs.y, cur = WIDE_EMPTY;
                while (sp > 0) {
                    int nd2; float dd;
                    ORT_PPOP(nd2, dd)
                    if (dd <= cull) { cur = nd2; break; }
                }
                if (cur == WIDE_EMPTY) finish(slot);
                W.cs[slot] = make_int2(cur, sp);
            }
            __syncwarp();
        }
    }
}

#undef ORT_PPUSH
#undef ORT_PPOP
#undef ORT_CSWAP

} // namespace ort
