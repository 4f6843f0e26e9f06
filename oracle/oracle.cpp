// oracle.cpp — CPU restatement of raytracer-odin's per-pixel path-tracing loop.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (raytracer-odin_b200/, the CUDA library,
// the CLI) may import, link or execute this file; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs do, and only as the checker / the timed CPU
// baseline.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures, it
// cannot be compiled here (no Odin toolchain), and its arithmetic leans on Odin's `core:`
// library (core:math/linalg inverse/normalize/cross/quaternion, core:sort, core:math/rand), whose
// sources are not under /root/reference and whose version is not pinned by the repo.  Where the
// operation order lives in `core:`, this file FIXES an order and says so next to the function;
// every f32 operation is individually rounded (build with -ffp-contract=off, no fast-math).
// The random streams are the build's counter-based Philox streams (BASELINE.json north_star),
// not Odin's per-task reseeded default generator (raytracer.odin:552).
//
// Each function cites the reference file:line it follows (paths relative to /root/reference).

#include "../include/odinrt_b200.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

constexpr float INF_F32 = std::numeric_limits<float>::infinity();
constexpr float PI_F = 3.14159265358979323846264338327950288f;  // f32(math.PI)
constexpr float TAU_F = 6.28318530717958647692528676655900576f; // f32(math.TAU)
constexpr float RAY_EPS = 1e-3f;                                // raytracer.odin:418, shading.odin:66

struct V3 {
    float x, y, z;
};
inline V3 v3(const float* p) { return {p[0], p[1], p[2]}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 operator/(V3 a, V3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }

// Odin builtin min/max lower to select(x < y, x, y) / select(x > y, x, y).
inline float omin(float a, float b) { return a < b ? a : b; }
inline float omax(float a, float b) { return a > b ? a : b; }
inline V3 vmin(V3 a, V3 b) { return {omin(a.x, b.x), omin(a.y, b.y), omin(a.z, b.z)}; }
inline V3 vmax(V3 a, V3 b) { return {omax(a.x, b.x), omax(a.y, b.y), omax(a.z, b.z)}; }

// linalg.dot on [3]f32: x*x' + y*y' + z*z', left to right.
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// linalg.cross: swizzle(a,1,2,0)*swizzle(b,2,0,1) - swizzle(a,2,0,1)*swizzle(b,1,2,0).
inline V3 cross(V3 a, V3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float length(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalize(V3 a) { return a / length(a); } // linalg.normalize = v / length(v)

// utils.odin:6-20
inline float sq(float x) { return x * x; }
inline float compsum(V3 a) { return a.x + a.y + a.z; }
inline float norm_l1(V3 a) { return compsum(V3{std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)}); }

struct Ray {
    V3 o, d;
};

// ------------------------------------------------------------------------------------------
// Counter-based RNG: Philox4x32-10 (Salmon et al. 2011).  Replaces rand.float32 / float32_range
// / int_max (raytracer.odin:581-582, shading.odin:10-11,42,45-46,103,142).
//   counter = (pixel = py*W+px [y up, unflipped], sample_lo, sample_hi, block), key = seed
//   block 0           -> r0,r1 = pixel jitter (raytracer.odin:581-582)
//   block 1+bounce    -> r0 = strategy (shading.odin:142); r1..r3 = the strategy's draws in
//                        source order: cosine {phi, z} (shading.odin:10-11), light {index, u, v}
//                        (shading.odin:42,45-46), VNDF {u1, u2} (shading.odin:103)
//   float32() = (r >> 8) * 2^-24;  float32_range(lo,hi) = float32()*(hi-lo)+lo;
//   int_max(n) = (u64(r) * n) >> 32.
// ------------------------------------------------------------------------------------------
struct Philox {
    uint32_t r[4];
};
inline Philox philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                            uint32_t k1) {
    for (int i = 0; i < 10; i++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return {{c0, c1, c2, c3}};
}
inline float u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }

struct Stream {
    uint32_t pixel;
    uint64_t sample;
    uint64_t seed;
    Philox block(uint32_t b) const {
        return philox4x32_10(pixel, (uint32_t)sample, (uint32_t)(sample >> 32), b, (uint32_t)seed,
                             (uint32_t)(seed >> 32));
    }
};

// ------------------------------------------------------------------------------------------
// AABB (raytracer.odin:152-209)
// ------------------------------------------------------------------------------------------
struct AABB {
    V3 lo, hi;
};
const AABB AABB_EMPTY = {{INF_F32, INF_F32, INF_F32}, {-INF_F32, -INF_F32, -INF_F32}};
inline AABB aabb_merge(AABB a, AABB b) { return {vmin(a.lo, b.lo), vmax(a.hi, b.hi)}; } // :161
inline AABB aabb_of_triangle(const ort_triangle& t) {                                    // :197
    V3 p = v3(t.p), q = p + v3(t.u), r = p + v3(t.v);
    AABB a = {p, p}; // aabb_of_points :188 — starts from points[0], merges all three
    a.lo = vmin(a.lo, p); a.hi = vmax(a.hi, p);
    a.lo = vmin(a.lo, q); a.hi = vmax(a.hi, q);
    a.lo = vmin(a.lo, r); a.hi = vmax(a.hi, r);
    return a;
}
inline float aabb_area(AABB a) { // :206 compsum(size.xyz * size.yzx)
    V3 s = a.hi - a.lo;
    return compsum(V3{s.x * s.y, s.y * s.z, s.z * s.x});
}

// ------------------------------------------------------------------------------------------
// check_intersect_ray_aabb (raytracer.odin:119-134)
// ------------------------------------------------------------------------------------------
inline bool check_intersect_ray_aabb(const Ray& g, const float* lo_, const float* hi_,
                                     float max_dist, float* t_out) {
    V3 lo = v3(lo_), hi = v3(hi_);
    Ray ray = {g.o - lo, g.d};
    V3 extent = hi - lo;
    if (length(ray.o - extent / 2.0f) - length(extent / 2.0f) > max_dist) return false; // :122
    V3 t1_raw = (extent - ray.o) / ray.d;                                                 // :125
    V3 t2_raw = (-ray.o) / ray.d;                                                         // :126
    V3 t_min = vmin(t1_raw, t2_raw);
    V3 t_max = vmax(t1_raw, t2_raw);
    float t1 = omax(omax(t_min.x, t_min.y), t_min.z);
    float t2 = omin(omin(t_max.x, t_max.y), t_max.z);
    if (t1 > t2) return false;
    if (t2 < 0) return false;
    *t_out = omax(t1, 0.0f);
    return true;
}

// ------------------------------------------------------------------------------------------
// intersect_ray_triangle (raytracer.odin:136-150)
//
// `linalg.inverse(a) * b` with a[0]=U, a[1]=V, a[2]=-d as COLUMNS.  core:math/linalg's
// matrix3x3_inverse is adjugate * (1/determinant); its source is not under /root/reference, so
// the ORDER BELOW IS THIS ORACLE'S DEFINITION (m[r][c] = row r, column c):
//   adj[0][0]=+(m11*m22-m21*m12) adj[0][1]=-(m01*m22-m21*m02) adj[0][2]=+(m01*m12-m11*m02)
//   adj[1][0]=-(m10*m22-m20*m12) adj[1][1]=+(m00*m22-m20*m02) adj[1][2]=-(m00*m12-m10*m02)
//   adj[2][0]=+(m10*m21-m20*m11) adj[2][1]=-(m00*m21-m20*m01) adj[2][2]=+(m00*m11-m10*m01)
//   det = m00*(m11*m22-m12*m21) + (-m01)*(m10*m22-m12*m20) + m02*(m10*m21-m11*m20)
//   inv[r][c] = adj[r][c] * (1/det);   (inv*b)[r] = inv[r][0]*b0 + inv[r][1]*b1 + inv[r][2]*b2
// every operation rounded to f32, sums left to right, no fused multiply-add.
// ------------------------------------------------------------------------------------------
struct GeomHit {
    bool inside;
    float t, u, v;
};
#ifdef ORC_VARIANTS
// SENSITIVITY STUDY ONLY (liboracle_alt.so, tests/test_oracle_sensitivity.py): alternative definitions of the
// operations whose order lives in Odin's `core:` / compiler, to MEASURE how much of the parity claim depends on the
// choice made above.  Never compiled into liboracle.so / liboracle_native.so.
//   bit 0  inverse = adjugate / det            (nine divisions instead of adjugate * (1/det))
//   bit 1  det expanded along the first column (m00*a00 + m10*a01 + m20*a02)
//   bit 2  every a*b - c*d contracted to fma(a, b, -(c*d))   (a compiler that contracts)
//   bit 3  matrix * vector with fused multiply-adds: fma(i2, b2, fma(i1, b1, i0*b0))   (llvm.fmuladd lowering)
//   bit 4  bvh_build's sort breaks ties in REVERSE input order (an unstable sort's other extreme)
int g_variant = 0;
inline float pd(float a, float b, float c, float d) { // a*b - c*d
    return (g_variant & 4) ? std::fma(a, b, -(c * d)) : a * b - c * d;
}
inline float mv(float i0, float i1, float i2, V3 b) {
    return (g_variant & 8) ? std::fma(i2, b.z, std::fma(i1, b.y, i0 * b.x)) : i0 * b.x + i1 * b.y + i2 * b.z;
}
#endif
inline GeomHit intersect_ray_triangle(const Ray& ray, const ort_triangle& tr) {
    V3 b = ray.o - v3(tr.p);
    float m00 = tr.u[0], m10 = tr.u[1], m20 = tr.u[2];
    float m01 = tr.v[0], m11 = tr.v[1], m21 = tr.v[2];
    float m02 = -ray.d.x, m12 = -ray.d.y, m22 = -ray.d.z;
#ifdef ORC_VARIANTS
    if (g_variant & 15) {
        float a00 = +pd(m11, m22, m21, m12), a01 = -pd(m01, m22, m21, m02), a02 = +pd(m01, m12, m11, m02);
        float a10 = -pd(m10, m22, m20, m12), a11 = +pd(m00, m22, m20, m02), a12 = -pd(m00, m12, m10, m02);
        float a20 = +pd(m10, m21, m20, m11), a21 = -pd(m00, m21, m20, m01), a22 = +pd(m00, m11, m10, m01);
        float det;
        if (g_variant & 2) det = m00 * a00 + m10 * a01 + m20 * a02;
        else det = m00 * pd(m11, m22, m12, m21) + (-m01) * pd(m10, m22, m12, m20) + m02 * pd(m10, m21, m11, m20);
        float u, v, t;
        if (g_variant & 1) {
            u = mv(a00 / det, a01 / det, a02 / det, b);
            v = mv(a10 / det, a11 / det, a12 / det, b);
            t = mv(a20 / det, a21 / det, a22 / det, b);
        } else {
            float id = 1.0f / det;
            u = mv(a00 * id, a01 * id, a02 * id, b);
            v = mv(a10 * id, a11 * id, a12 * id, b);
            t = mv(a20 * id, a21 * id, a22 * id, b);
        }
        if (u < 0 || v < 0 || u + v > 1) return {false, -1.0f, 0, 0};
        return {dot(v3(tr.ng), ray.d) > 0, t, u, v};
    }
#endif
    float a00 = +(m11 * m22 - m21 * m12), a01 = -(m01 * m22 - m21 * m02), a02 = +(m01 * m12 - m11 * m02);
    float a10 = -(m10 * m22 - m20 * m12), a11 = +(m00 * m22 - m20 * m02), a12 = -(m00 * m12 - m10 * m02);
    float a20 = +(m10 * m21 - m20 * m11), a21 = -(m00 * m21 - m20 * m01), a22 = +(m00 * m11 - m10 * m01);
    float da = m00 * (m11 * m22 - m12 * m21);
    float db = (-m01) * (m10 * m22 - m12 * m20);
    float dc = m02 * (m10 * m21 - m11 * m20);
    float det = da + db + dc;
    float id = 1.0f / det;
    float u = (a00 * id) * b.x + (a01 * id) * b.y + (a02 * id) * b.z;
    float v = (a10 * id) * b.x + (a11 * id) * b.y + (a12 * id) * b.z;
    float t = (a20 * id) * b.x + (a21 * id) * b.y + (a22 * id) * b.z;
    if (u < 0 || v < 0 || u + v > 1) return {false, -1.0f, 0, 0}; // :143
    return {dot(v3(tr.ng), ray.d) > 0, t, u, v};
}

// ------------------------------------------------------------------------------------------
// bvh_build (raytracer.odin:227-342).  `sort.sort` (core:sort) is an unstable sort whose
// source is absent; this oracle DEFINES the permutation as std::stable_sort on aabb.lo[axis]
// (SURVEY §8c-ii).  The triangles are carried as an index permutation and moved once at the end,
// which is observationally identical to swapping the 168-byte structs in place (:265-268).
// ------------------------------------------------------------------------------------------
struct Item {
    AABB box;
    int64_t idx;
};
struct Builder {
    std::vector<Item> items;
    std::vector<AABB> buf;
    std::vector<Item> tmp;
    std::vector<ort_bvh_node>* nodes;

    static float axis_of(const V3& v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

    // try_axis (:276-304)
    void try_axis(int axis, int64_t b, int64_t n, float* best_sah_out, int64_t* best_index_out) {
        Item* it = items.data() + b;
#ifdef ORC_VARIANTS
        if (g_variant & 16) {
            std::reverse(it, it + n); // ties now come out in reverse order of the (previous) sequence
            std::stable_sort(it, it + n, [axis](const Item& l, const Item& r) {
                return axis_of(l.box.lo, axis) < axis_of(r.box.lo, axis);
            });
        } else
#endif
        std::stable_sort(it, it + n, [axis](const Item& l, const Item& r) {
            return axis_of(l.box.lo, axis) < axis_of(r.box.lo, axis);
        });
        AABB* bf = buf.data() + b;
        for (int64_t i = n - 1; i >= 0; i--) { // :289-294
            bf[i] = it[i].box;
            if (i != n - 1) bf[i] = aabb_merge(bf[i], bf[i + 1]);
        }
        float best_sah = INF_F32;
        int64_t best_index = 0;
        AABB total = AABB_EMPTY;
        for (int64_t i = 1; i < n; i++) { // :297-302
            total = aabb_merge(total, it[i - 1].box);
            float sah = aabb_area(total) * (float)i + aabb_area(bf[i]) * (float)(n - i);
            if (sah < best_sah) { best_sah = sah; best_index = i; }
        }
        *best_sah_out = best_sah;
        *best_index_out = best_index;
    }

    int64_t recurse(int64_t b, int64_t n) {
        if (n <= 4) { // LEAF_NODE_THRESHOLD :230,243
            AABB a = AABB_EMPTY;
            for (int64_t i = 0; i < n; i++) a = aabb_merge(a, items[b + i].box);
            ort_bvh_node nd{};
            nd.lo[0] = a.lo.x; nd.lo[1] = a.lo.y; nd.lo[2] = a.lo.z;
            nd.hi[0] = a.hi.x; nd.hi[1] = a.hi.y; nd.hi[2] = a.hi.z;
            nd.kind = 0; nd.a = b; nd.b = n;
            nodes->push_back(nd);
            return (int64_t)nodes->size() - 1;
        }
        float sah0, sah1, sah2;
        int64_t split, dummy;
        try_axis(0, b, n, &sah0, &dummy); // :306
        AABB total = buf[b];              // :307
        try_axis(1, b, n, &sah1, &dummy);
        try_axis(2, b, n, &sah2, &dummy);
        float s;
        if (sah0 < sah1 && sah0 < sah2) try_axis(0, b, n, &s, &split);      // :311-317
        else if (sah1 < sah0 && sah1 < sah2) try_axis(1, b, n, &s, &split);
        else try_axis(2, b, n, &s, &split);
        int64_t left = recurse(b, split);              // :318
        int64_t right = recurse(b + split, n - split); // :319
        ort_bvh_node nd{};
        nd.lo[0] = total.lo.x; nd.lo[1] = total.lo.y; nd.lo[2] = total.lo.z;
        nd.hi[0] = total.hi.x; nd.hi[1] = total.hi.y; nd.hi[2] = total.hi.z;
        nd.kind = 1; nd.a = left; nd.b = right;
        nodes->push_back(nd);
        return (int64_t)nodes->size() - 1;
    }
};

// ------------------------------------------------------------------------------------------
// Traversal (raytracer.odin:344-430)
// ------------------------------------------------------------------------------------------
struct Hit {
    int64_t trig; // -1 == nil
    float t;
    bool inside;
    float u, v;
};
struct Counters {
    uint64_t rays, node_pops, box_tests, tri_tests, stack_drops, stack_high, exact_ties;
    uint64_t light_rays, light_node_pops, light_box_tests, light_tri_tests;
    float last_tie_t; // instrumentation: t of the most recent exact tie of the current ray (NaN = none)
};

struct SceneView {
    const ort_scene* s;
    int mode; // 0 = faithful (reference push order incl. duplicate-left, :395-409), 1 = ideal
};

// cast_ray_through_trigs (:351-369)
// `best_trig` (the current winner of the enclosing traversal, or -1) is instrumentation only: it
// lets the tie counter tell "another triangle produced a bit-identical t" from "the same triangle
// was re-tested" (the faithful push order re-visits leaves).
inline Hit cast_ray_through_trigs(const ort_triangle* trigs, int64_t first, int64_t count,
                                  const Ray& ray, float max_dist, Counters* c, int64_t best_trig = -1) {
    Hit hit{-1, max_dist, false, 0, 0};
    for (int64_t i = 0; i < count; i++) {
        GeomHit gh = intersect_ray_triangle(ray, trigs[first + i]);
        if (c) {
            c->tri_tests++;
            int64_t cur_best = hit.trig >= 0 ? hit.trig : best_trig;
            if (gh.t > 0 && gh.t == hit.t && cur_best >= 0 && cur_best != first + i) {
                c->exact_ties++;
                c->last_tie_t = gh.t;
            }
        }
        if (gh.t > 0 && gh.t < hit.t) hit = {first + i, gh.t, gh.inside, gh.u, gh.v}; // :360
    }
    return hit;
}

// Small_Array(64,int): append on a full array is a silent no-op (raytracer.odin:379).
struct Stack64 {
    int64_t v[64];
    int n = 0;
    bool append(int64_t x) {
        if (n >= 64) return false;
        v[n++] = x;
        return true;
    }
    int64_t pop_back() { return v[--n]; }
};

// cast_ray_through_bvh (:371-414)
inline Hit cast_ray_through_bvh(const ort_bvh_node* bvh, int64_t n_nodes, const ort_triangle* trigs,
                                const Ray& ray, float max_dist, int mode, Counters* c) {
    Hit hit{-1, max_dist, false, 0, 0};
    float tt;
    if (c) c->box_tests++;
    if (!check_intersect_ray_aabb(ray, bvh[n_nodes - 1].lo, bvh[n_nodes - 1].hi, max_dist, &tt)) return hit;
    Stack64 stack;
    stack.append(n_nodes - 1);
    auto push = [&](int64_t id) {
        if (!stack.append(id) && c) c->stack_drops++;
        if (c && (uint64_t)stack.n > c->stack_high) c->stack_high = stack.n;
    };
    while (stack.n > 0) {
        int64_t id = stack.pop_back();
        if (c) c->node_pops++;
        const ort_bvh_node& node = bvh[id];
        if (node.kind == 0) {
            Hit cur = cast_ray_through_trigs(trigs, node.a, node.b, ray, max_dist, c, hit.trig);
            if (cur.t < max_dist) { // :388
                hit = cur;
                max_dist = cur.t;
            }
        } else {
            float tl = 0, tr = 0;
            bool hl = check_intersect_ray_aabb(ray, bvh[node.a].lo, bvh[node.a].hi, max_dist, &tl);
            bool hr = check_intersect_ray_aabb(ray, bvh[node.b].lo, bvh[node.b].hi, max_dist, &tr);
            if (c) c->box_tests += 2;
            if (!hl && !hr) continue;
            if (mode == 0) {
                // :396-409 verbatim: the second `if hl` is not an `else`, so when both children
                // are hit the left child is pushed a second time.
                if (hl && hr) {
                    if (tl < tr) { push(node.a); push(node.b); }
                    else { push(node.b); push(node.a); }
                }
                if (hl) push(node.a);
                else if (hr) push(node.b);
            } else {
                // duplicate-free near-first order on the same tests
                if (hl && hr) {
                    if (tl < tr) { push(node.b); push(node.a); }
                    else { push(node.a); push(node.b); }
                } else if (hl) push(node.a);
                else push(node.b);
            }
        }
    }
    return hit;
}

// cast_ray (:416-430)
inline Hit cast_ray(const SceneView& sv, const Ray& g, float max_dist, Counters* c) {
    Ray ray = {g.o + g.d * RAY_EPS, g.d};
    if (c) c->rays++;
    Hit hit{-1, 0.0f, false, 0, 0}; // zero-initialised `hit` (:416)
    Hit hit2 = cast_ray_through_bvh(sv.s->bvh, sv.s->n_bvh_nodes, sv.s->triangles, ray, max_dist, sv.mode, c);
    if (hit2.t < max_dist) hit = hit2;
    hit.t += RAY_EPS;
    return hit;
}

// ------------------------------------------------------------------------------------------
// textures.odin:79-135
// ------------------------------------------------------------------------------------------
struct V4 {
    float x, y, z, w;
};
inline int64_t floored_mod(int64_t a, int64_t m) { // Odin %%
    int64_t r = a % m;
    return r < 0 ? r + m : r;
}
inline V4 texture_index(const ort_texture& t, int64_t cx, int64_t cy, bool srgb) { // :79-104
    int64_t index = cy * t.stride + cx * t.channels;
    float px[4] = {1, 1, 1, 1};
    if (t.data == nullptr) return {1, 1, 1, 1};
    if (!t.is_f32) {
        const uint8_t* d = (const uint8_t*)t.data;
        for (int c = 0; c < t.channels; c++) px[c] = (float)d[index + c] / 255.0f;
    } else {
        const float* d = (const float*)t.data;
        for (int c = 0; c < t.channels; c++) px[c] = d[index + c];
    }
    if (srgb) { // :99-101 linalg.pow(pixel.rgb, 2.2)
        px[0] = std::pow(px[0], 2.2f);
        px[1] = std::pow(px[1], 2.2f);
        px[2] = std::pow(px[2], 2.2f);
    }
    return {px[0], px[1], px[2], px[3]};
}
inline float lerp1(float a, float b, float t) { return a * (1 - t) + b * t; } // math.lerp
inline V4 lerp4(V4 a, V4 b, float t) {
    return {lerp1(a.x, b.x, t), lerp1(a.y, b.y, t), lerp1(a.z, b.z, t), lerp1(a.w, b.w, t)};
}
inline V4 texture_sample(const ort_texture* tex, float cu, float cv, bool srgb, V4 def) { // :106-135
    if (tex == nullptr) return def;
    float pcx = cu * (float)tex->width, pcy = cv * (float)tex->height;
    float lox = std::floor(pcx), loy = std::floor(pcy);
    float hix = std::ceil(pcx), hiy = std::ceil(pcy);
    float tx = pcx - lox, ty = pcy - loy;
    int64_t c00x = floored_mod((int64_t)lox, tex->width), c00y = floored_mod((int64_t)loy, tex->height);
    int64_t c11x = floored_mod((int64_t)hix, tex->width), c11y = floored_mod((int64_t)hiy, tex->height);
    V4 p00 = texture_index(*tex, c00x, c00y, srgb);
    V4 p01 = texture_index(*tex, c00x, c11y, srgb);
    V4 p10 = texture_index(*tex, c11x, c00y, srgb);
    V4 p11 = texture_index(*tex, c11x, c11y, srgb);
    return lerp4(lerp4(p00, p01, ty), lerp4(p10, p11, ty), tx);
}
inline const ort_texture* sampler(const ort_scene* s, int32_t idx) {
    return idx < 0 ? nullptr : &s->textures[idx];
}

// ------------------------------------------------------------------------------------------
// shading.odin
// ------------------------------------------------------------------------------------------
struct PointMaterial { // raytracer.odin:25-32
    V3 pos, color, normal, emission;
    float metallic, roughness;
};

inline V3 sphere_uniform(uint32_t r_phi, uint32_t r_z) { // shading.odin:9-15
    float phi = u01(r_phi) * (TAU_F - 0.0f) + 0.0f;
    float z = u01(r_z) * (1.0f - -1.0f) + -1.0f;
    float x = std::sin(phi), y = std::cos(phi); // x, y := math.sincos(phi)
    float radius = std::sqrt(1 - sq(z));
    return {x * radius, y * radius, z};
}
inline V3 cosine_weighted(V3 n, uint32_t r1, uint32_t r2) { // :32-35
    return normalize(sphere_uniform(r1, r2) + n);
}
inline float cosine_weighted_pdf(V3 n, V3 omega) { return omax(dot(n, omega) / PI_F, 0.0f); } // :37-39

inline V3 surface_sampling(const ort_scene* s, V3 origin, uint32_t r_idx, uint32_t r_u, uint32_t r_v) { // :41-50
    int64_t index = (int64_t)(((uint64_t)r_idx * (uint64_t)s->n_light_triangles) >> 32);
    const ort_triangle& trig = s->light_triangles[index];
    float u = u01(r_u) * (1.0f - 0.0f) + 0.0f;
    float v = u01(r_v) * (1.0f - 0.0f) + 0.0f;
    if (u + v > 1) { u = 1 - u; v = 1 - v; }
    V3 world = v3(trig.p) + u * v3(trig.u) + v * v3(trig.v);
    return normalize(world - origin);
}

inline float surface_sampling_pdf_trigs_sum(const ort_triangle* trigs, int64_t first, int64_t count,
                                            const Ray& ray, Counters* c) { // :52-60
    float p = 0;
    for (int64_t i = 0; i < count; i++) {
        const ort_triangle& trig = trigs[first + i];
        GeomHit hit = intersect_ray_triangle(ray, trig);
        if (c) c->light_tri_tests++;
        if (!(hit.t >= 0)) continue;
        float weight = sq(hit.t) / std::fabs(dot(v3(trig.ng), ray.d));
        p += 2 / length(cross(v3(trig.u), v3(trig.v))) * weight;
    }
    return p;
}

inline float surface_sampling_pdf_bvh_sum(const ort_scene* s, const Ray& g, Counters* c) { // :62-94
    Ray ray = {g.o + g.d * RAY_EPS, g.d};
    const ort_bvh_node* bvh = s->light_bvh;
    int64_t n = s->n_light_bvh_nodes;
    float p = 0, tt;
    if (c) { c->light_rays++; c->light_box_tests++; }
    if (!check_intersect_ray_aabb(ray, bvh[n - 1].lo, bvh[n - 1].hi, INF_F32, &tt)) return p;
    Stack64 stack;
    stack.append(n - 1);
    while (stack.n > 0) {
        int64_t id = stack.pop_back();
        if (c) c->light_node_pops++;
        const ort_bvh_node& node = bvh[id];
        if (node.kind == 0) {
            p += surface_sampling_pdf_trigs_sum(s->light_triangles, node.a, node.b, ray, c);
        } else {
            bool hl = check_intersect_ray_aabb(ray, bvh[node.a].lo, bvh[node.a].hi, INF_F32, &tt);
            bool hr = check_intersect_ray_aabb(ray, bvh[node.b].lo, bvh[node.b].hi, INF_F32, &tt);
            if (c) c->light_box_tests += 2;
            if (hl) stack.append(node.a);
            if (hr) stack.append(node.b);
        }
    }
    return p;
}
inline float surface_sampling_pdf(const ort_scene* s, const Ray& ray, Counters* c) { // :96-100
    return surface_sampling_pdf_bvh_sum(s, ray, c) / (float)s->n_light_triangles;
}

// quaternion helpers: linalg.mul(quaternion128, [3]f32) = v + q.w*t + cross(q.xyz, t) with
// t = cross(2*q.xyz, v)  (core:math/linalg quaternion128_mul_vector3 — order defined here).
struct Quat {
    float w, x, y, z;
};
inline V3 quat_mul_vec(Quat q, V3 v) {
    V3 qv = {q.x, q.y, q.z};
    V3 t = cross(2.0f * qv, v);
    return v + q.w * t + cross(qv, t);
}
inline Quat conj(Quat q) { return {q.w, -q.x, -q.y, -q.z}; }
inline Quat vndf_rotation(V3 n) { // shading.odin:104-106
    float w = std::sqrt((1 + n.z) / 2);
    if (w > 0) return {w, -n.y / (2 * w), n.x / (2 * w), 0};
    return {0, 1, 0, 0};
}

inline V3 vndf_sampling(V3 n, V3 omega, float alpha, float u1, float u2) { // :102-122
    Quat rotation = vndf_rotation(n);
    V3 V = quat_mul_vec(conj(rotation), omega);
    V3 Vh = normalize(V3{alpha * V.x, alpha * V.y, V.z});
    float len = std::hypot(Vh.x, Vh.y);
    V3 T1 = len == 0 ? V3{1, 0, 0} : V3{-Vh.y / len, Vh.x / len, 0};
    V3 T2 = cross(Vh, T1);
    float r = std::sqrt(u1);
    float phi = TAU_F * u2;
    float t1 = std::sin(phi), t2 = std::cos(phi); // t1, t2 := math.sincos(phi)
    t1 *= r;
    t2 *= r;
    float s = 0.5f * (1 + Vh.z);
    t2 = (1 - s) * std::sqrt(1 - sq(t1)) + s * t2;
    V3 Nh = t1 * T1 + t2 * T2 + Vh * std::sqrt(omax(0.0f, 1 - sq(t1) - sq(t2)));
    V3 Ne = normalize(V3{alpha * Nh.x, alpha * Nh.y, omax(0.0f, Nh.z)});
    return quat_mul_vec(rotation, Ne);
}

inline float vndf_sampling_pdf(V3 n, V3 omega, float alpha, V3 L) { // :124-137
    V3 Ne = normalize(omega + L);
    Quat rotation = vndf_rotation(n);
    V3 V = quat_mul_vec(conj(rotation), omega);
    V3 N = quat_mul_vec(conj(rotation), Ne);
    float alpha2 = sq(alpha);
    float lambda = (-1 + std::sqrt(1 + alpha2 * (sq(V.x) + sq(V.y)) / sq(V.z))) * 0.5f;
    float G1 = 1 / (1 + lambda);
    float D = 1 / (PI_F * alpha2 * sq(sq(N.x / alpha) + sq(N.y / alpha) + sq(N.z)));
    float normal = G1 * omax(0.0f, dot(V, N)) * D / V.z;
    return normal / (4 * dot(L, Ne));
}

// sample (:139-151)
inline V3 sample_dir(const ort_scene* s, const PointMaterial& mat, const Ray& in_ray, const Philox& r) {
    float t = u01(r.r[0]);
    if (t <= 0.33333f) {
        return cosine_weighted(mat.normal, r.r[1], r.r[2]);
    } else if (t < 0.666666f && s->n_light_triangles > 0) {
        return surface_sampling(s, mat.pos, r.r[1], r.r[2], r.r[3]);
    } else {
        V3 n = vndf_sampling(mat.normal, -in_ray.d, sq(mat.roughness), u01(r.r[1]), u01(r.r[2]));
        return in_ray.d - 2 * dot(n, in_ray.d) * n;
    }
}

// pdf (:153-162)
inline float pdf_dir(const ort_scene* s, const PointMaterial& mat, const Ray& in_ray, const Ray& out_ray,
                     Counters* c) {
    bool has_lights = s->n_light_triangles > 0;
    return (cosine_weighted_pdf(mat.normal, out_ray.d) +
            (has_lights ? surface_sampling_pdf(s, out_ray, c) : 0.0f) +
            vndf_sampling_pdf(mat.normal, -in_ray.d, sq(mat.roughness), out_ray.d) *
                (has_lights ? 1.0f : 2.0f)) / 3;
}

inline float smith_geometry_ggx(V3 n, V3 x, float alpha2) { // :187-190
    float cosine = dot(n, x);
    return 2 * omax(cosine, 0.0f) / (cosine + std::sqrt(alpha2 + (1 - alpha2) * sq(cosine)));
}
inline V3 lerp3(V3 a, V3 b, float t) { return a * (1 - t) + b * t; }
inline V3 lerp3v(V3 a, V3 b, V3 t) { return a * (V3{1, 1, 1} - t) + b * t; }

// shade (:164-204)
inline V3 shade(const PointMaterial& mat, const Ray& in_ray, const Ray& out_ray) {
    float alpha = sq(mat.roughness);
    float alpha2 = sq(alpha);
    V3 L = out_ray.d;
    V3 V = -in_ray.d;
    V3 H = normalize(L + V);
    V3 N = mat.normal;
    float cosine = dot(L, N);
    const float f0 = 0.04f;
    float frensel_base = std::pow(1 - dot(H, L), 5.0f);
    float frensel_diff_spec = f0 + 0.96f * frensel_base; // (f90 - f0) folds to 0.96
    V3 frensel_metallic = mat.color + (V3{1, 1, 1} - mat.color) * frensel_base;
    float hn = dot(H, N);
    float step = hn < 0.0f ? 0.0f : 1.0f; // math.step(0, x)
    float distribution_term = alpha2 * step / (PI_F * sq((alpha2 - 1) * sq(hn) + 1));
    float geometry_term = smith_geometry_ggx(N, L, alpha2) * smith_geometry_ggx(N, V, alpha2);
    float cook_torrance = distribution_term * geometry_term / (4 * dot(V, N));
    V3 specular = cook_torrance * V3{1, 1, 1};
    V3 diffuse = mat.color * omax(cosine, 0.0f) / PI_F;
    V3 metallic = specular * frensel_metallic;
    V3 dielectic = lerp3(diffuse, specular, frensel_diff_spec);
    return lerp3(dielectic, metallic, mat.metallic);
}

// ------------------------------------------------------------------------------------------
// raytrace (raytracer.odin:432-518)
// ------------------------------------------------------------------------------------------
inline PointMaterial fetch_material(const ort_scene* s, const Hit& hit, const Ray& ray) {
    const ort_triangle& trig = s->triangles[hit.trig];
    float u = hit.u, v = hit.v;
    const ort_material& om = s->materials[trig.material_index];
    float w0 = 1 - u - v;
    float tcx = trig.tex1[0] * w0 + trig.tex2[0] * u + trig.tex3[0] * v; // :454
    float tcy = trig.tex1[1] * w0 + trig.tex2[1] * u + trig.tex3[1] * v;
    V4 white = {1, 1, 1, 1};
    V4 mr = texture_sample(sampler(s, om.metallic_roughness_texture), tcx, tcy, false, white);
    V3 p = v3(trig.p) + v3(trig.u) * u + v3(trig.v) * v; // :456
    V3 normal;
    if (om.normal_texture >= 0) { // :458-470
        // linalg.normalize on the [4]f32 tangent: the 4-component length (w included) divides.
        float t4[4];
        for (int i = 0; i < 4; i++) t4[i] = trig.tan1[i] * w0 + trig.tan2[i] * u + trig.tan3[i] * v;
        float l4 = std::sqrt(t4[0] * t4[0] + t4[1] * t4[1] + t4[2] * t4[2] + t4[3] * t4[3]);
        for (int i = 0; i < 4; i++) t4[i] = t4[i] / l4;
        V3 local_x = {t4[0], t4[1], t4[2]};
        V3 local_z = normalize(v3(trig.n1) * w0 + v3(trig.n2) * u + v3(trig.n3) * v);
        V3 local_y = cross(local_z, local_x) * t4[3];
        V4 ns = texture_sample(sampler(s, om.normal_texture), tcx, tcy, false, V4{0.5f, 1.0f, 0.5f, 0.0f});
        V3 ln = V3{ns.x, ns.y, ns.z} * 2.0f - V3{1, 1, 1};
        // local_basis columns = (local_x, local_y, local_z); basis * v summed left to right
        V3 nb = {local_x.x * ln.x + local_y.x * ln.y + local_z.x * ln.z,
                 local_x.y * ln.x + local_y.y * ln.y + local_z.y * ln.z,
                 local_x.z * ln.x + local_y.z * ln.y + local_z.z * ln.z};
        normal = normalize(nb);
    } else {
        normal = normalize(v3(trig.n1) * w0 + v3(trig.n2) * u + v3(trig.n3) * v); // :472
    }
    V4 ct = texture_sample(sampler(s, om.color_texture), tcx, tcy, true, white);
    V4 et = texture_sample(sampler(s, om.emission_texture), tcx, tcy, true, white);
    PointMaterial mat;
    mat.pos = p;
    mat.normal = normal;
    mat.color = v3(om.color_factor) * V3{ct.x, ct.y, ct.z};
    mat.emission = v3(om.emission_factor) * V3{et.x, et.y, et.z};
    mat.roughness = omax(om.roughness_factor * mr.y, 0.03f);
    mat.metallic = om.metallic_factor * mr.z;
    if (hit.inside) mat.normal = -mat.normal; // :485-488
    (void)ray;
    return mat;
}

inline V3 env_lookup(const ort_scene* s, V3 d) { // :437-446
    float tu = 0.5f + std::atan2(d.z, d.x) / TAU_F;
    float tv = 0.5f - std::asin(d.y) / PI_F;
    V4 e = texture_sample(s->env_map, tu, tv, false, V4{0, 0, 0, 0});
    return {e.x, e.y, e.z};
}

V3 raytrace(const SceneView& sv, const Ray& ray, int32_t depth_left, int32_t ray_depth, const Stream& rng,
            Counters* c) {
    if (depth_left == 0) return {0, 0, 0};
    Hit hit = cast_ray(sv, ray, INF_F32, c);
    if (hit.trig < 0) return env_lookup(sv.s, ray.d);
    PointMaterial mat = fetch_material(sv.s, hit, ray);
    Philox r = rng.block(1u + (uint32_t)(ray_depth - depth_left));
    V3 d_reflected = sample_dir(sv.s, mat, ray, r);
    Ray reflected = {mat.pos, d_reflected};
    float pdf = pdf_dir(sv.s, mat, ray, reflected, c);
    V3 value = shade(mat, ray, reflected);
    V3 exitance;
    if (norm_l1(value) / pdf > 1e-5f) { // :495
        V3 irradiance = raytrace(sv, reflected, depth_left - 1, ray_depth, rng, c);
        exitance = value * irradiance / pdf + mat.emission;
    } else {
        exitance = mat.emission;
    }
    return exitance;
}

// pixel_to_ray_dir (raytracer.odin:529-538): mat4(basis) * scale(tan_x,tan_y,1) *
// translate(-1,-1,1) * scale(1/(w/2), 1/(h/2), 1).  4x4 products as C[i][j] = sum_k A[i][k]B[k][j],
// k = 0..3 left to right (order defined here; core:math/linalg).
struct M4 {
    float m[4][4]; // m[row][col]
};
inline M4 m4_mul(const M4& a, const M4& b) {
    M4 c;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            c.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j] + a.m[i][3] * b.m[3][j];
    return c;
}
inline M4 m4_identity() {
    M4 r{};
    for (int i = 0; i < 4; i++) r.m[i][i] = 1;
    return r;
}
inline M4 pixel_to_ray_dir(const ort_camera& cam, uint32_t w, uint32_t h) {
    float dx = (float)w, dy = (float)h;
    float aspect_ratio = dx / dy;
    float tan_fov_x = std::tan(cam.fov_x / 2);
    float tan_fov_y = tan_fov_x / aspect_ratio;
    M4 b = m4_identity();
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) b.m[r][c] = cam.basis[3 * c + r];
    M4 s1 = m4_identity();
    s1.m[0][0] = tan_fov_x; s1.m[1][1] = tan_fov_y; s1.m[2][2] = 1;
    M4 tr = m4_identity();
    tr.m[0][3] = -1; tr.m[1][3] = -1; tr.m[2][3] = 1;
    M4 s2 = m4_identity();
    s2.m[0][0] = 1.0f / (dx / 2); s2.m[1][1] = 1.0f / (dy / 2); s2.m[2][2] = 1.0f / 1.0f;
    return m4_mul(m4_mul(m4_mul(b, s1), tr), s2);
}
inline Ray primary_ray(const ort_camera& cam, const M4& M, uint32_t px, uint32_t py, const Stream& rng) { // :580-593
    Philox r = rng.block(0);
    float x = (float)px + u01(r.r[0]);
    float y = (float)py + u01(r.r[1]);
    float z = 0.0f, w = 1.0f;
    V3 raw = {M.m[0][0] * x + M.m[0][1] * y + M.m[0][2] * z + M.m[0][3] * w,
              M.m[1][0] * x + M.m[1][1] * y + M.m[1][2] * z + M.m[1][3] * w,
              M.m[2][0] * x + M.m[2][1] * y + M.m[2][2] * z + M.m[2][3] * w};
    return {v3(cam.pos), normalize(raw)};
}

// rc_set_pixel (main.odin:89-102)
inline void rc_set_pixel(ort_sample_stats* pixels, uint32_t w, uint32_t h, uint32_t px, uint32_t py, V3 color) {
    int64_t i = (int64_t)(h - py - 1) * (int64_t)w + (int64_t)px;
    ort_sample_stats* p = &pixels[i];
    if (p->count == 0) { p->first[0] = color.x; p->first[1] = color.y; p->first[2] = color.z; }
    p->count += 1;
    p->last[0] = color.x; p->last[1] = color.y; p->last[2] = color.z;
    p->total[0] += color.x; p->total[1] += color.y; p->total[2] += color.z;
    p->total_squared[0] += color.x * color.x;
    p->total_squared[1] += color.y * color.y;
    p->total_squared[2] += color.z * color.z;
}

inline uint64_t ceil_div(uint64_t x, uint64_t y) { return (x + y - 1) / y; }

} // namespace

// ==========================================================================================
// C entry points (ctypes)
// ==========================================================================================
extern "C" {

typedef struct orc_counters {
    uint64_t rays, node_pops, box_tests, tri_tests, stack_drops, stack_high, exact_ties;
    uint64_t light_rays, light_node_pops, light_box_tests, light_tri_tests;
} orc_counters;

static void add_counters(orc_counters* dst, const Counters& c) {
    dst->rays += c.rays; dst->node_pops += c.node_pops; dst->box_tests += c.box_tests;
    dst->tri_tests += c.tri_tests; dst->stack_drops += c.stack_drops;
    if (c.stack_high > dst->stack_high) dst->stack_high = c.stack_high;
    dst->exact_ties += c.exact_ties; dst->light_rays += c.light_rays;
    dst->light_node_pops += c.light_node_pops; dst->light_box_tests += c.light_box_tests;
    dst->light_tri_tests += c.light_tri_tests;
}

void orc_philox(uint32_t pixel, uint64_t sample, uint32_t block, uint64_t seed, uint32_t out[4]) {
    Stream s{pixel, sample, seed};
    Philox p = s.block(block);
    for (int i = 0; i < 4; i++) out[i] = p.r[i];
}

// bvh_build (raytracer.odin:227-342); sorts tris in place; returns node count.
int64_t orc_bvh_build(ort_triangle* tris, int64_t n, ort_bvh_node* nodes_out, int64_t cap) {
    Builder b;
    std::vector<ort_bvh_node> nodes;
    nodes.reserve((size_t)(n / 2 + 1));
    b.nodes = &nodes;
    b.items.resize((size_t)n);
    b.buf.resize((size_t)n);
    for (int64_t i = 0; i < n; i++) b.items[i] = {aabb_of_triangle(tris[i]), i};
    b.recurse(0, n);
    std::vector<ort_triangle> sorted((size_t)n);
    for (int64_t i = 0; i < n; i++) sorted[i] = tris[b.items[i].idx];
    if (n) std::memcpy(tris, sorted.data(), sizeof(ort_triangle) * (size_t)n);
    if ((int64_t)nodes.size() > cap) return -(int64_t)nodes.size();
    std::memcpy(nodes_out, nodes.data(), sizeof(ort_bvh_node) * nodes.size());
    return (int64_t)nodes.size();
}

int orc_check_intersect_ray_aabb(const float o[3], const float d[3], const float lo[3], const float hi[3],
                                 float max_dist, float* t_out) {
    Ray r = {v3(o), v3(d)};
    float t = 0;
    bool h = check_intersect_ray_aabb(r, lo, hi, max_dist, &t);
    *t_out = h ? t : 0.0f;
    return h ? 1 : 0;
}

// out = {t, u, v, inside}
void orc_intersect_ray_triangle(const float o[3], const float d[3], const ort_triangle* tri, float out[4]) {
    Ray r = {v3(o), v3(d)};
    GeomHit g = intersect_ray_triangle(r, *tri);
    out[0] = g.t; out[1] = g.u; out[2] = g.v; out[3] = g.inside ? 1.0f : 0.0f;
}

// cast_ray on n rays; mode 0 faithful / 1 ideal.
// ties_out (may be NULL): 1 where a DIFFERENT triangle produced a t bit-identical to the winner's,
// i.e. where the result depends on the visiting order (SURVEY.md §7 "exact-t ties").
void orc_trace_rays(const ort_scene* scene, const ort_ray* rays, int64_t n, int mode, ort_hit* out,
                    orc_counters* counters, int threads, uint8_t* ties_out) {
    SceneView sv{scene, mode};
    if (threads < 1) threads = 1;
    std::vector<Counters> cs((size_t)threads);
    std::vector<std::thread> pool;
    auto work = [&](int tid) {
        Counters c{};
        for (int64_t i = tid; i < n; i += threads) {
            Ray r = {v3(rays[i].o), v3(rays[i].d)};
            c.last_tie_t = std::numeric_limits<float>::quiet_NaN();
            const uint64_t drops_before = c.stack_drops;
            Hit h = cast_ray(sv, r, INF_F32, &c);
            // bit 0: exact-t tie; bit 1: the 64-entry stack (raytracer.odin:379) dropped a push on this ray
            if (ties_out) ties_out[i] = (uint8_t)(((h.trig >= 0 && c.last_tie_t + RAY_EPS == h.t) ? 1 : 0) |
                                                  (c.stack_drops > drops_before ? 2 : 0));
            out[i].t = h.t; out[i].u = h.u; out[i].v = h.v;
            out[i].tri = (int32_t)h.trig;
            out[i].material = h.trig < 0 ? -1 : (int32_t)scene->triangles[h.trig].material_index;
            out[i].inside = h.inside ? 1 : 0;
        }
        cs[tid] = c;
    };
    for (int t = 1; t < threads; t++) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
    if (counters) for (auto& c : cs) add_counters(counters, c);
}

void orc_light_pdf(const ort_scene* scene, const ort_ray* rays, int64_t n, float* out) {
    for (int64_t i = 0; i < n; i++) {
        Ray r = {v3(rays[i].o), v3(rays[i].d)};
        out[i] = scene->n_light_triangles > 0 ? surface_sampling_pdf(scene, r, nullptr) : 0.0f;
    }
}

// Primary rays + their cast_ray result for one sample index; pixel order y*w + x (unflipped).
void orc_primary_hits(const ort_scene* scene, uint32_t w, uint32_t h, uint64_t sample, uint64_t seed, int mode,
                      ort_hit* out, ort_ray* rays_out, orc_counters* counters, int threads, uint8_t* ties_out) {
    SceneView sv{scene, mode};
    M4 M = pixel_to_ray_dir(scene->cam, w, h);
    if (threads < 1) threads = 1;
    std::vector<Counters> cs((size_t)threads);
    std::vector<std::thread> pool;
    auto work = [&](int tid) {
        Counters c{};
        for (uint32_t py = tid; py < h; py += threads)
            for (uint32_t px = 0; px < w; px++) {
                Stream rng{py * w + px, sample, seed};
                Ray r = primary_ray(scene->cam, M, px, py, rng);
                int64_t i = (int64_t)py * w + px;
                if (rays_out) {
                    rays_out[i].o[0] = r.o.x; rays_out[i].o[1] = r.o.y; rays_out[i].o[2] = r.o.z;
                    rays_out[i].d[0] = r.d.x; rays_out[i].d[1] = r.d.y; rays_out[i].d[2] = r.d.z;
                }
                c.last_tie_t = std::numeric_limits<float>::quiet_NaN();
                Hit hh = cast_ray(sv, r, INF_F32, &c);
                if (ties_out) ties_out[i] = (hh.trig >= 0 && c.last_tie_t + RAY_EPS == hh.t) ? 1 : 0;
                out[i].t = hh.t; out[i].u = hh.u; out[i].v = hh.v;
                out[i].tri = (int32_t)hh.trig;
                out[i].material = hh.trig < 0 ? -1 : (int32_t)scene->triangles[hh.trig].material_index;
                out[i].inside = hh.inside ? 1 : 0;
            }
        cs[tid] = c;
    };
    for (int t = 1; t < threads; t++) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
    if (counters) for (auto& c : cs) add_counters(counters, c);
}

// render_scene / render_task (raytracer.odin:528-623), one trial.
//   schedule 0: the reference's task order — id -> (sample block of 32, tile x, tile y), y fastest
//               (:557-560), threads pull ids from one atomic counter (:551), unsynchronised
//               accumulation like the reference (use for TIMING; with threads > 1 the per-pixel
//               f32 summation order is whatever the scheduler produced, as in the reference).
//   schedule 1: same tasks, but one worker owns a tile for all its sample blocks in order, so the
//               per-pixel summation order is sample order regardless of thread count (use for
//               CHECKING).
// Pixel box [x0,x1) x [y0,y1) restricts rendering to a window (bounded CPU samples).
// tile_stride > 1 renders only the 4x4-pixel tiles whose tile coordinates are both multiples of it: a
// bounded sample SPREAD OVER THE WHOLE FRAME (sky, horizon and ground in the frame's own proportions)
// instead of a centre window, for timing the CPU arm on workloads too large to render in full.
void orc_render_strided(const ort_scene* scene, uint32_t w, uint32_t h, int32_t ray_depth, uint64_t first_sample,
                        uint64_t n_samples, uint64_t seed, int mode, int schedule, int threads, uint32_t x0,
                        uint32_t y0, uint32_t x1, uint32_t y1, uint32_t tile_stride, ort_sample_stats* out,
                        orc_counters* counters) {
    SceneView sv{scene, mode};
    M4 M = pixel_to_ray_dir(scene->cam, w, h);
    if (threads < 1) threads = 1;
    if (x1 > w) x1 = w;
    if (y1 > h) y1 = h;
    const uint64_t TILE = 4, TILE_SAMPLES = 32; // raytracer.odin:525-526
    uint64_t dim_x = ceil_div(w, TILE), dim_y = ceil_div(h, TILE);
    uint64_t sample_blocks = ceil_div(n_samples, TILE_SAMPLES);
    uint64_t total_tasks = schedule == 0 ? sample_blocks * dim_x * dim_y : dim_x * dim_y;
    std::atomic<uint64_t> tile_id{0};
    std::vector<Counters> cs((size_t)threads);
    std::vector<std::thread> pool;

    auto run_block = [&](uint64_t sample_coord, uint64_t x_coord, uint64_t y_coord, Counters* c) {
        uint64_t num_samples = std::min<uint64_t>(TILE_SAMPLES, n_samples - TILE_SAMPLES * sample_coord);
        uint32_t start_x = (uint32_t)(TILE * x_coord), end_x = std::min<uint32_t>(start_x + TILE, w);
        uint32_t start_y = (uint32_t)(TILE * y_coord), end_y = std::min<uint32_t>(start_y + TILE, h);
        if (start_x >= x1 || end_x <= x0 || start_y >= y1 || end_y <= y0) return;
        if (tile_stride > 1 && (x_coord % tile_stride != 0 || y_coord % tile_stride != 0)) return;
        for (uint64_t sample = 0; sample < num_samples; sample++)
            for (uint32_t px = start_x; px < end_x; px++)
                for (uint32_t py = start_y; py < end_y; py++) {
                    if (px < x0 || px >= x1 || py < y0 || py >= y1) continue;
                    uint64_t s = first_sample + sample_coord * TILE_SAMPLES + sample;
                    Stream rng{py * w + px, s, seed};
                    Ray r = primary_ray(scene->cam, M, px, py, rng);
                    V3 exitance = raytrace(sv, r, ray_depth, ray_depth, rng, c);
                    rc_set_pixel(out, w, h, px, py, exitance);
                }
    };
    auto work = [&](int tid) {
        Counters c{};
        for (;;) {
            uint64_t id = tile_id.fetch_add(1, std::memory_order_relaxed);
            if (id >= total_tasks) break;
            uint64_t y_coord = id % dim_y; id /= dim_y;
            uint64_t x_coord = id % dim_x; id /= dim_x;
            if (schedule == 0) {
                run_block(id, x_coord, y_coord, &c);
            } else {
                for (uint64_t sb = 0; sb < sample_blocks; sb++) run_block(sb, x_coord, y_coord, &c);
            }
        }
        cs[tid] = c;
    };
    for (int t = 1; t < threads; t++) pool.emplace_back(work, t); // threads-1 workers + caller (:609-619)
    work(0);
    for (auto& t : pool) t.join();
    if (counters) for (auto& c : cs) add_counters(counters, c);
}

void orc_render(const ort_scene* scene, uint32_t w, uint32_t h, int32_t ray_depth, uint64_t first_sample,
                uint64_t n_samples, uint64_t seed, int mode, int schedule, int threads, uint32_t x0,
                uint32_t y0, uint32_t x1, uint32_t y1, ort_sample_stats* out, orc_counters* counters) {
    orc_render_strided(scene, w, h, ray_depth, first_sample, n_samples, seed, mode, schedule, threads, x0, y0, x1, y1, 1,
                       out, counters);
}

// --- shading / texture KAT entry points -----------------------------------------------------
static PointMaterial make_mat(const float n[3], const float color[3], float metallic, float roughness) {
    PointMaterial m{};
    m.normal = v3(n); m.color = v3(color); m.metallic = metallic; m.roughness = roughness;
    m.pos = {0, 0, 0}; m.emission = {0, 0, 0};
    return m;
}
void orc_shade(const float n[3], const float color[3], float metallic, float roughness, const float in_d[3],
               const float out_d[3], float out[3]) {
    PointMaterial m = make_mat(n, color, metallic, roughness);
    Ray in{{0, 0, 0}, v3(in_d)}, o{{0, 0, 0}, v3(out_d)};
    V3 r = shade(m, in, o);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
float orc_cosine_weighted_pdf(const float n[3], const float omega[3]) { return cosine_weighted_pdf(v3(n), v3(omega)); }
float orc_vndf_sampling_pdf(const float n[3], const float omega[3], float alpha, const float L[3]) {
    return vndf_sampling_pdf(v3(n), v3(omega), alpha, v3(L));
}
void orc_vndf_sampling(const float n[3], const float omega[3], float alpha, float u1, float u2, float out[3]) {
    V3 r = vndf_sampling(v3(n), v3(omega), alpha, u1, u2);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void orc_cosine_weighted(const float n[3], uint32_t r1, uint32_t r2, float out[3]) {
    V3 r = cosine_weighted(v3(n), r1, r2);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
// sample() with an explicit Philox block (4 x u32) — shading.odin:139-151
void orc_sample(const ort_scene* scene, const float n[3], const float pos[3], float roughness, const float in_d[3],
                const uint32_t r[4], float out[3]) {
    float color[3] = {1, 1, 1};
    PointMaterial m = make_mat(n, color, 0, roughness);
    m.pos = v3(pos);
    Ray in{{0, 0, 0}, v3(in_d)};
    Philox p{{r[0], r[1], r[2], r[3]}};
    V3 d = sample_dir(scene, m, in, p);
    out[0] = d.x; out[1] = d.y; out[2] = d.z;
}
float orc_pdf(const ort_scene* scene, const float n[3], const float pos[3], float roughness, const float in_d[3],
              const float out_d[3]) {
    float color[3] = {1, 1, 1};
    PointMaterial m = make_mat(n, color, 0, roughness);
    m.pos = v3(pos);
    Ray in{{0, 0, 0}, v3(in_d)}, o{v3(pos), v3(out_d)};
    return pdf_dir(scene, m, in, o, nullptr);
}
void orc_texture_sample(const ort_texture* tex, float u, float v, int srgb, const float def[4], float out[4]) {
    V4 r = texture_sample(tex, u, v, srgb != 0, V4{def[0], def[1], def[2], def[3]});
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
void orc_env_lookup(const ort_scene* scene, const float d[3], float out[3]) {
    V3 r = env_lookup(scene, v3(d));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void orc_pixel_to_ray_dir(const ort_camera* cam, uint32_t w, uint32_t h, float out16[16]) {
    M4 M = pixel_to_ray_dir(*cam, w, h);
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) out16[4 * r + c] = M.m[r][c];
}

// get_rgb_image, mode Mean (output.odin:21-80): total/count -> max(.,0) -> ACES -> pow(1/2.2)
// -> round(*255) -> u8.
void orc_get_rgb_image(const ort_sample_stats* px, uint32_t w, uint32_t h, uint8_t* out) {
    for (int64_t i = 0; i < (int64_t)w * h; i++) {
        for (int c = 0; c < 3; c++) {
            float raw = px[i].total[c] / (float)px[i].count;
            raw = omax(raw, 0.0f);
            float x = raw;
            float tm = (x * (2.51f * x + 0.03f)) / (x * (2.43f * x + 0.59f) + 0.14f);
            tm = tm < 0.0f ? 0.0f : (tm > 1.0f ? 1.0f : tm); // linalg.clamp
            float g = std::pow(tm, (float)(1.0 / 2.2)); // untyped constant 1/2.2 folds before f32
            out[i * 3 + c] = (uint8_t)std::round(g * 255.0f);
        }
    }
}

int orc_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }
#ifdef ORC_VARIANTS
void orc_set_variant(int bits) { g_variant = bits; }
#endif

} // extern "C"
