// traverse8.cuh — k_trace8: the same two queries as k_trace (closest hit = cast_ray,
// raytracer.odin:416-430; all-hit light pdf sum = surface_sampling_pdf_bvh_sum, shading.odin:62-94)
// on the 8-wide re-emission of the reference BVH (Wide8Node, wide_bvh.h).
//
// Why a second layout: profiles/r1_sensitivity.md shows k_trace bound by the number of dependent
// node round trips per ray (each waits for the slowest of ~13 divergent lanes) at the knee of its
// occupancy curve.  Eight children per visit halve the round trips; slots are pre-assigned by octant so
// the hit children are visited in ascending (slot ^ ray octant) order — no per-visit sorting network —
// and a visit pushes at most ONE 64-bit stack entry (the remaining hit siblings as a bit mask).
// Leaf children do not get stack entries at all: their triangles are tested right after the visit that
// hit their boxes.
//
// MEASURED (profiles/r1_sensitivity.md): parity-green, but 1.6x SLOWER than k_trace on C2 and C4 (trace
// 54.4 vs 34.1 ms, 103 vs 66.8 ms), at 4, 5 or 6 CTAs per SM alike.  With this renderer's exact
// (85-instruction, IEEE-division) triangle solve, testing the triangles of EVERY hit leaf of a node
// before descending costs more than the halved node round trips save; k_trace's exact near-first order
// with per-entry distance culling tests far fewer triangles.  Opt-in only: ORT_BVH8=1.
//
// Numerics are k_trace's: conservative padded slab tests (make_ray), the reference's triangle solve bit
// for bit (tri_det_t / tri_uv), triangles of a leaf in reference order with strict `<`.
#pragma once
#include "device_math.cuh"
#include "wide_bvh.h"

namespace ort {

// bit permutation i -> i ^ o of the low 8 bits (o = ray octant): hit bits in slot order become hit
// bits in visiting order and back
__device__ __forceinline__ uint32_t xor_permute8(uint32_t m, uint32_t o) {
    if (o & 1u) m = ((m & 0x55u) << 1) | ((m & 0xAAu) >> 1);
    if (o & 2u) m = ((m & 0x33u) << 2) | ((m & 0xCCu) >> 2);
    if (o & 4u) m = ((m & 0x0Fu) << 4) | ((m & 0xF0u) >> 4);
    return m;
}

template <bool CLOSEST>
__global__ void __launch_bounds__(TRACE_THREADS, ORT_TRACE8_MIN_CTAS)
k_trace8(const SceneDev s, const TraceArgs a) {
    __shared__ uint2 sh_stack[SMEM_STACK][TRACE_THREADS];
    uint2 l_stack[LOCAL_STACK];

    const uint32_t n = *a.n_ptr;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const float best_pad = 1.0f + 7.62939453125e-06f; // 1 + 2^-17: distance culling margin
    const float inf = __int_as_float(0x7f800000);

    RaySetup r;
    float best = inf, hu = 0.0f, hv = 0.0f, lsumv = 0.0f;
    float cull = inf;
    int htri = -1, sp = 0;
    int cur = -1;              // node to visit next, -1: none
    bool has_ray = false;
    uint32_t oct = 0;          // ray octant: bit a set when direction component a is negative
    uint32_t g_base = 0, g_bits = 0; // current node group: child_base | pending hit bits (visiting order, low 8) + imask << 8
    uint32_t t_base = 0, t_mask = 0; // triangles waiting to be tested
    uint32_t pos = 0;
    bool exhausted = false;

    for (;;) {
        // ---- refill idle lanes (dynamic fetch, as in k_trace)
        const bool idle = !has_ray;
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
                    oct = (uint32_t)(r.sx & 1) | ((uint32_t)(r.sy & 1) << 1) | ((uint32_t)(r.sz & 1) << 2);
                    best = inf; hu = 0.0f; hv = 0.0f; htri = -1; lsumv = 0.0f;
                    cull = inf; sp = 0; pos = idx;
                    g_bits = 0u; t_mask = 0u;
                    cur = CLOSEST ? 0 : s.light_root8;
                    has_ray = true;
                }
            }
            exhausted = base + (uint32_t)cnt >= n;
        }
        if (__ballot_sync(0xffffffffu, has_ray) == 0u) break;

        if (has_ray) {
            for (;;) {
                // ---- node stage: visit nodes until this lane has triangles to test (or runs out of nodes)
                while (cur >= 0) {
                    const float4* nd = s.nodes8 + (size_t)cur * 16;
                    const int ox = (int)(oct & 1u) * 2, oy = (int)((oct >> 1) & 1u) * 2, oz = (int)((oct >> 2) & 1u) * 2;
                    const F8 nxp = ldg8(nd + ox), fxp = ldg8(nd + (ox ^ 2));
                    const F8 nyp = ldg8(nd + 4 + oy), fyp = ldg8(nd + 4 + (oy ^ 2));
                    const F8 nzp = ldg8(nd + 8 + oz), fzp = ldg8(nd + 8 + (oz ^ 2));
                    const uint4 hd = __ldg(reinterpret_cast<const uint4*>(nd + 12));
                    const uint4 tm0 = __ldg(reinterpret_cast<const uint4*>(nd + 14));
                    const uint4 tm1 = __ldg(reinterpret_cast<const uint4*>(nd + 15));
                    uint32_t hm = 0u, tm = 0u;
#define ORT_BOX8(H, k, BIT, TM)                                                                       \
    {                                                                                                 \
        const float tn = fmaxf(fmaxf(fmaf(nxp.H.k, r.ix, r.nx), fmaf(nyp.H.k, r.iy, r.ny)),           \
                               fmaxf(fmaf(nzp.H.k, r.iz, r.nz), 0.0f));                               \
        const float tf = fminf(fminf(fmaf(fxp.H.k, r.ix, r.fx), fmaf(fyp.H.k, r.iy, r.fy)),           \
                               fminf(fmaf(fzp.H.k, r.iz, r.fz), cull));                               \
        if (tn <= tf) { hm |= (BIT); tm |= (TM); }                                                    \
    }
                    ORT_BOX8(lo, x, 1u, tm0.x) ORT_BOX8(lo, y, 2u, tm0.y) ORT_BOX8(lo, z, 4u, tm0.z) ORT_BOX8(lo, w, 8u, tm0.w)
                    ORT_BOX8(hi, x, 16u, tm1.x) ORT_BOX8(hi, y, 32u, tm1.y) ORT_BOX8(hi, z, 64u, tm1.z) ORT_BOX8(hi, w, 128u, tm1.w)
#undef ORT_BOX8
                    // the siblings still pending in the group this node came from go onto the stack
                    if (g_bits & 0xffu) {
                        const uint2 e = make_uint2(g_base, g_bits);
                        if (sp < SMEM_STACK) sh_stack[sp][threadIdx.x] = e; else l_stack[sp - SMEM_STACK] = e;
                        sp++;
                    }
                    g_base = hd.x;
                    g_bits = xor_permute8(hm & hd.z, oct) | (hd.z << 8);
                    t_base = hd.y;
                    t_mask = tm;
                    cur = -1;
                    if (t_mask == 0u) {
                        // next node: nearest pending child of the current group, else the stack
                        if ((g_bits & 0xffu) == 0u && sp > 0) {
                            sp--;
                            const uint2 e = sp < SMEM_STACK ? sh_stack[sp][threadIdx.x] : l_stack[sp - SMEM_STACK];
                            g_base = e.x; g_bits = e.y;
                        }
                        if (g_bits & 0xffu) {
                            const uint32_t p = (uint32_t)__ffs((int)(g_bits & 0xffu)) - 1u;
                            g_bits &= ~(1u << p);
                            const uint32_t slot = p ^ oct;
                            cur = (int)(g_base + (uint32_t)__popc((g_bits >> 8) & ((1u << slot) - 1u)));
                        }
                    }
                    if (__popc(__activemask()) < a.inner_min) break;
                }
                // ---- triangle stage
                if (t_mask != 0u) {
                    do {
                        const uint32_t b = (uint32_t)__ffs((int)t_mask) - 1u;
                        t_mask &= t_mask - 1u;
                        const float4* tp = s.tris8 + (size_t)(t_base + b) * 4;
                        const F8 tab = ldg8(tp);
                        const float4 ta = tab.lo, tb = tab.hi;
                        const F8 tcd = ldg8(tp + 2); // third float4 + (reference triangle id, padding)
                        const float4 tc = tcd.lo;
                        float id, bx, by, bz, t, a00, a10;
                        tri_det_t(r, ta, tb, tc, id, bx, by, bz, t, a00, a10);
                        if (CLOSEST) {
                            if (t > 0.0f && t < best) { // raytracer.odin:360
                                float u, v;
                                if (tri_uv(r, ta, tb, tc, id, bx, by, bz, a00, a10, u, v)) {
                                    best = t; hu = u; hv = v; htri = __float_as_int(tcd.hi.x);
                                    cull = best * best_pad;
                                }
                            }
                        } else {
                            float u, v;
                            if (t >= 0.0f && tri_uv(r, ta, tb, tc, id, bx, by, bz, a00, a10, u, v)) { // shading.odin:52-60
                                const float4 L = ldg4(s.llight + (__float_as_int(tcd.hi.x) - (int)s.light_tri_base));
                                const float weight = (t * t) / fabsf(L.x * r.dx + L.y * r.dy + L.z * r.dz);
                                lsumv += L.w * weight;
                            }
                        }
                    } while (t_mask != 0u);
                }
                // ---- next node for lanes that have none: current group, else the stack, else the ray is done
                if (cur < 0) {
                    if ((g_bits & 0xffu) == 0u && sp > 0) {
                        sp--;
                        const uint2 e = sp < SMEM_STACK ? sh_stack[sp][threadIdx.x] : l_stack[sp - SMEM_STACK];
                        g_base = e.x; g_bits = e.y;
                    }
                    if (g_bits & 0xffu) {
                        const uint32_t p = (uint32_t)__ffs((int)(g_bits & 0xffu)) - 1u;
                        g_bits &= ~(1u << p);
                        const uint32_t slot = p ^ oct;
                        cur = (int)(g_base + (uint32_t)__popc((g_bits >> 8) & ((1u << slot) - 1u)));
                    } else {
                        if (CLOSEST) a.hits[pos] = make_float4(htri >= 0 ? best : 0.0f, hu, hv, __int_as_float(htri));
                        else a.lsum[pos] = lsumv;
                        has_ray = false;
                        break; // ray finished
                    }
                }
                if (!exhausted && __popc(__activemask()) < a.refill_threshold) break; // go refill
            }
        }
    }
}

} // namespace ort
