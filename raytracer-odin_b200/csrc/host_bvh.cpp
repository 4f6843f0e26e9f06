// host_bvh.cpp — host-side (CPU, one-shot) pieces of the library:
//   * ort_bvh_build: the reference's SAH sweep builder (raytracer.odin:227-342) for hosts that do
//     not bring their own (the Python harness; Odin keeps its own bvh_build).  Same splits, same
//     triangle permutation, same post-order node array — implemented with stable LSD radix sorts
//     on (key, index) records and task-parallel subtrees instead of comparison sorts that swap
//     168-byte structs.
//   * build_wide_bvh: re-emission of the binary reference BVH as the 4-wide, cache-line sized
//     node layout the traversal kernels read (wide_bvh.h).
//   * per-triangle traversal records and the pixel->ray matrix.
// Compiled with -ffp-contract=off: the SAH costs, the adjugate row and the matrix entries are
// plain individually-rounded f32 operations.
#include "wide_bvh.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <future>
#include <limits>
#include <thread>

namespace ort {
namespace {

constexpr float kInf = std::numeric_limits<float>::infinity();

struct Box {
    float lo[3], hi[3];
};
struct Ent {
    Box box;
    uint32_t idx;
    uint32_t key;
};

inline void box_grow(Box& a, const Box& b) {
    for (int k = 0; k < 3; k++) {
        // aabb_merge(lhs = a, rhs = b) with Odin's min/max = select(l < r, l, r) / select(l > r, l, r):
        // on equal-comparing values (+0 / -0) the RIGHT operand is kept, bit for bit
        a.lo[k] = a.lo[k] < b.lo[k] ? a.lo[k] : b.lo[k];
        a.hi[k] = a.hi[k] > b.hi[k] ? a.hi[k] : b.hi[k];
    }
}
inline Box box_empty() { return {{kInf, kInf, kInf}, {-kInf, -kInf, -kInf}}; }
inline float half_area(const Box& b) { // aabb_area raytracer.odin:206: x*y + y*z + z*x
    float sx = b.hi[0] - b.lo[0], sy = b.hi[1] - b.lo[1], sz = b.hi[2] - b.lo[2];
    return sx * sy + sy * sz + sz * sx;
}

// Monotone float -> uint key; -0 is folded onto +0 so that keys are equal exactly when the
// floats compare equal (ties must stay in their previous order).
inline uint32_t sort_key(float x) {
    x = x + 0.0f;
    uint32_t u;
    std::memcpy(&u, &x, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct Build {
    Ent* ent;
    Ent* tmp;
    Box* suffix;
    std::atomic<bool> failed{false};

    // Stable ascending sort of ent[b, b+n) by box.lo[axis].
    void sort_axis(int64_t b, int64_t n, int axis) {
        Ent* e = ent + b;
        if (n <= 64) {
            for (int64_t i = 1; i < n; i++) {
                Ent x = e[i];
                float kx = x.box.lo[axis];
                int64_t j = i - 1;
                while (j >= 0 && kx < e[j].box.lo[axis]) { e[j + 1] = e[j]; j--; }
                e[j + 1] = x;
            }
            return;
        }
        for (int64_t i = 0; i < n; i++) e[i].key = sort_key(e[i].box.lo[axis]);
        Ent* src = e;
        Ent* dst = tmp + b;
        const int bits[3] = {11, 11, 10};
        int shift = 0;
        for (int pass = 0; pass < 3; pass++) {
            const uint32_t mask = (1u << bits[pass]) - 1;
            uint32_t count[2048] = {0};
            for (int64_t i = 0; i < n; i++) count[(src[i].key >> shift) & mask]++;
            if (count[(src[0].key >> shift) & mask] != (uint32_t)n) { // digit not constant
                uint32_t sum = 0;
                for (uint32_t d = 0; d <= mask; d++) { uint32_t c = count[d]; count[d] = sum; sum += c; }
                for (int64_t i = 0; i < n; i++) dst[count[(src[i].key >> shift) & mask]++] = src[i];
                std::swap(src, dst);
            }
            shift += bits[pass];
        }
        if (src != e) std::memcpy(e, src, sizeof(Ent) * (size_t)n);
    }

    // One try_axis pass (raytracer.odin:276-304) after the sort: returns the SAH minimum and
    // the first index reaching it.
    float sweep(int64_t b, int64_t n, int64_t* split) {
        Ent* e = ent + b;
        Box* s = suffix + b;
        s[n - 1] = e[n - 1].box;
        for (int64_t i = n - 2; i >= 0; i--) { s[i] = e[i].box; box_grow(s[i], s[i + 1]); }
        float best = kInf;
        int64_t best_i = 0;
        Box prefix = box_empty();
        for (int64_t i = 1; i < n; i++) {
            box_grow(prefix, e[i - 1].box);
            float cost = half_area(prefix) * (float)i + half_area(s[i]) * (float)(n - i);
            if (cost < best) { best = cost; best_i = i; }
        }
        *split = best_i;
        return best;
    }

    // Appends the subtree over ent[b, b+n) to `out` in post-order, child links relative to
    // out's own indexing.  Returns the subtree root's index in `out`.
    int64_t subtree(int64_t b, int64_t n, std::vector<ort_bvh_node>& out, int fork_levels) {
        ort_bvh_node nd;
        std::memset(&nd, 0, sizeof nd);
        if (n <= 4) {
            Box bb = box_empty();
            for (int64_t i = 0; i < n; i++) box_grow(bb, ent[b + i].box);
            std::memcpy(nd.lo, bb.lo, 12);
            std::memcpy(nd.hi, bb.hi, 12);
            nd.kind = 0; nd.a = b; nd.b = n;
            out.push_back(nd);
            return (int64_t)out.size() - 1;
        }
        int64_t s0, s1, s2, split;
        sort_axis(b, n, 0);
        float c0 = sweep(b, n, &s0);
        Box total = suffix[b];
        sort_axis(b, n, 1);
        float c1 = sweep(b, n, &s1);
        sort_axis(b, n, 2);
        float c2 = sweep(b, n, &s2);
        if (c0 < c1 && c0 < c2) { sort_axis(b, n, 0); sweep(b, n, &split); }
        else if (c1 < c0 && c1 < c2) { sort_axis(b, n, 1); sweep(b, n, &split); }
        else split = s2; // already in axis-2 order; a stable re-sort is the identity
        if (split <= 0 || split >= n) { failed = true; return -1; } // the reference would not terminate
        int64_t left, right;
        if (fork_levels > 0 && n >= 32768) {
            std::vector<ort_bvh_node> lv, rv;
            auto fut = std::async(std::launch::async, [&] { return subtree(b, split, lv, fork_levels - 1); });
            int64_t rr = subtree(b + split, n - split, rv, fork_levels - 1);
            int64_t lr = fut.get();
            if (lr < 0 || rr < 0) { failed = true; return -1; }
            auto splice = [&out](std::vector<ort_bvh_node>& v, int64_t root) {
                int64_t base = (int64_t)out.size();
                for (auto& x : v) {
                    if (x.kind == 1) { x.a += base; x.b += base; }
                    out.push_back(x);
                }
                return base + root;
            };
            left = splice(lv, lr);
            right = splice(rv, rr);
        } else {
            left = subtree(b, split, out, 0);
            if (left < 0) return -1;
            right = subtree(b + split, n - split, out, 0);
            if (right < 0) return -1;
        }
        std::memcpy(nd.lo, total.lo, 12);
        std::memcpy(nd.hi, total.hi, 12);
        nd.kind = 1; nd.a = left; nd.b = right;
        out.push_back(nd);
        return (int64_t)out.size() - 1;
    }
};

} // namespace

namespace {

// fn(first, count) over [0, n) on up to `threads` host threads (inline for small ranges)
template <typename F>
void par_for(size_t n, int threads, size_t min_per_thread, F fn) {
    size_t nt = std::min<size_t>((size_t)std::max(threads, 1), n / std::max<size_t>(min_per_thread, 1));
    if (nt <= 1) { if (n) fn((size_t)0, n); return; }
    std::vector<std::thread> pool;
    const size_t per = (n + nt - 1) / nt;
    for (size_t t = 1; t < nt; t++) {
        const size_t a = t * per, b = std::min(n, a + per);
        if (a < b) pool.emplace_back([=] { fn(a, b - a); });
    }
    fn((size_t)0, std::min(n, per));
    for (auto& th : pool) th.join();
}

} // namespace

// Level-synchronous and parallel: the nodes of one wide level are independent once their output
// indices are known, and those follow from an exclusive scan of the per-node inner-child counts in
// queue order — exactly the indices the serial breadth-first emission hands out, so the result does
// not depend on the thread count (tests/test_host.py compares 1 and 16 threads).
bool build_wide_bvh(const ort_bvh_node* bvh, int64_t n_nodes, int64_t n_tris, WideBVH* out, const char** err, int threads) {
    out->nodes.clear();
    out->depth = 0;
    out->max_stack = 0;
    if (n_nodes <= 0) { *err = "empty BVH node array"; return false; }
    if (threads <= 0) threads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    {
        std::atomic<int> bad{0};
        par_for((size_t)n_nodes, threads, 1 << 16, [&](size_t f, size_t c) {
            for (size_t i = f; i < f + c; i++) {
                const ort_bvh_node& nd = bvh[i];
                if (nd.kind == 0) {
                    if (nd.a < 0 || nd.b < 0 || nd.b > 7 || nd.a + nd.b > n_tris || nd.a >= (1 << 28)) bad.store(1);
                } else if (nd.kind == 1) {
                    // post-order: children precede their parent
                    if (nd.a < 0 || nd.b < 0 || nd.a >= (int64_t)i || nd.b >= (int64_t)i) bad.store(2);
                } else bad.store(3);
            }
        });
        if (bad.load() == 1) { *err = "BVH leaf out of range (first/count)"; return false; }
        if (bad.load() == 2) { *err = "BVH branch child index out of range"; return false; }
        if (bad.load() == 3) { *err = "BVH node kind must be 0 (leaf) or 1 (branch)"; return false; }
    }
    const ort_bvh_node& root = bvh[n_nodes - 1];
    for (int k = 0; k < 3; k++) {
        float a = std::fabs(root.lo[k]), b = std::fabs(root.hi[k]);
        float m = a > b ? a : b;
        out->max_abs[k] = std::isfinite(m) ? m : 0.0f;
    }
    auto area = [&](int64_t id) {
        const ort_bvh_node& nd = bvh[id];
        float sx = nd.hi[0] - nd.lo[0], sy = nd.hi[1] - nd.lo[1], sz = nd.hi[2] - nd.lo[2];
        float a = sx * sy + sy * sz + sz * sx;
        return std::isfinite(a) ? a : 0.0f;
    };
    struct Item { int64_t src; int64_t kids[4]; int nk; int n_inner; size_t child_off; };
    std::vector<Item> level(1), next;
    level[0].src = n_nodes - 1;
    size_t level_base = 0; // index of the first node of this level in out->nodes
    out->nodes.reserve((size_t)n_nodes / 2 + 16); // a wide node absorbs at least one binary branch besides its own: no regrowth copies
    out->nodes.resize(1);
    while (!level.empty()) {
        out->depth++;
        // 1. children of every node of the level (parallel)
        par_for(level.size(), threads, 2048, [&](size_t f, size_t c) {
            for (size_t i = f; i < f + c; i++) {
                Item& it = level[i];
                it.nk = 0;
                if (bvh[it.src].kind == 0) {
                    if (bvh[it.src].b > 0) it.kids[it.nk++] = it.src; // a lone (root) leaf; an empty leaf yields an empty node
                } else {
                    it.kids[it.nk++] = bvh[it.src].a;
                    it.kids[it.nk++] = bvh[it.src].b;
                    while (it.nk < 4) { // open the largest inner child until the node is full
                        int pick = -1;
                        float best = -1.0f;
                        for (int k = 0; k < it.nk; k++)
                            if (bvh[it.kids[k]].kind == 1 && area(it.kids[k]) > best) { best = area(it.kids[k]); pick = k; }
                        if (pick < 0) break;
                        const int64_t open = it.kids[pick];
                        it.kids[pick] = bvh[open].a;
                        it.kids[it.nk++] = bvh[open].b;
                    }
                }
                it.n_inner = 0;
                for (int k = 0; k < it.nk; k++)
                    if (bvh[it.kids[k]].kind == 1) it.n_inner++;
            }
        });
        // 2. output indices of the next level: exclusive scan in queue order
        size_t n_next = 0;
        for (Item& it : level) { it.child_off = n_next; n_next += (size_t)it.n_inner; }
        const size_t next_base = level_base + level.size();
        if (next_base + n_next >= (size_t)0x7fffffff) { *err = "wide BVH too large"; return false; }
        out->nodes.resize(next_base + n_next);
        next.resize(n_next);
        // 3. emit the nodes of this level and the work items of the next one (parallel)
        par_for(level.size(), threads, 2048, [&](size_t f, size_t c) {
            for (size_t i = f; i < f + c; i++) {
                const Item& it = level[i];
                WideNode wn;
                for (int k = 0; k < 4; k++) {
                    for (int ax = 0; ax < 3; ax++) { wn.bounds[ax][0][k] = kInf; wn.bounds[ax][1][k] = -kInf; }
                    wn.child[k] = WIDE_EMPTY;
                    wn.reserved[k] = 0;
                }
                size_t inner = 0;
                for (int k = 0; k < it.nk; k++) {
                    const ort_bvh_node& ch = bvh[it.kids[k]];
                    // empty leaf (only a host-supplied BVH can hold one below the root): the slot stays unused AND
                    // keeps its inverted infinite box — the kernels rely on the box alone to skip unused slots
                    if (ch.kind == 0 && ch.b == 0) continue;
                    for (int ax = 0; ax < 3; ax++) { wn.bounds[ax][0][k] = ch.lo[ax]; wn.bounds[ax][1][k] = ch.hi[ax]; }
                    if (ch.kind == 0) {
                        wn.child[k] = ~(int32_t)((ch.a << 3) | ch.b);
                    } else {
                        const size_t pos = it.child_off + inner++;
                        wn.child[k] = (int32_t)(next_base + pos);
                        next[pos].src = it.kids[k];
                    }
                }
                out->nodes[level_base + i] = wn;
            }
        });
        level_base = next_base;
        level.swap(next);
    }
    // worst-case stack occupancy: F(node) = (#children - 1) + max F(inner child); children have
    // larger indices than their parent (breadth-first emission), so one reverse sweep suffices.
    std::vector<int> need(out->nodes.size(), 0);
    for (int64_t i = (int64_t)out->nodes.size() - 1; i >= 0; i--) {
        int c = 0, deepest = 0;
        for (int k = 0; k < 4; k++) {
            int32_t ch = out->nodes[i].child[k];
            if (ch == WIDE_EMPTY) continue;
            c++;
            if (ch >= 0 && need[ch] > deepest) deepest = need[ch];
        }
        need[i] = (c > 0 ? c - 1 : 0) + deepest;
    }
    out->max_stack = need[0] + 1;
    return true;
}

int64_t reference_stack_need(const ort_bvh_node* bvh, int64_t n_nodes) {
    if (n_nodes <= 0) return 0;
    // post-order array: children precede their parent (validated by build_wide_bvh)
    std::vector<int32_t> branches((size_t)n_nodes, 0);
    for (int64_t i = 0; i < n_nodes; i++) {
        const ort_bvh_node& nd = bvh[i];
        if (nd.kind == 1 && nd.a >= 0 && nd.b >= 0 && nd.a < i && nd.b < i)
            branches[(size_t)i] = 1 + std::max(branches[(size_t)nd.a], branches[(size_t)nd.b]);
    }
    return 2 * (int64_t)branches[(size_t)n_nodes - 1] + 1;
}

void make_isect_records(const ort_triangle* tris, int64_t n, TriIsect* out, bool light) {
    for (int64_t i = 0; i < n; i++) {
        const ort_triangle& t = tris[i];
        TriIsect r;
        r.p[0] = t.p[0]; r.p[1] = t.p[1]; r.p[2] = t.p[2];
        r.ux = t.u[0]; r.uy = t.u[1]; r.uz = t.u[2];
        r.vx = t.v[0]; r.vy = t.v[1]; r.vz = t.v[2];
        // third adjugate row of the column matrix [u | v | -d] (m[r][c]: m00=ux m10=uy m20=uz,
        // m01=vx m11=vy m21=vz): independent of the ray.
        r.c0 = +(r.uy * r.vz - r.uz * r.vy);
        r.c1 = -(r.ux * r.vz - r.uz * r.vx);
        r.c2 = +(r.ux * r.vy - r.uy * r.vx);
        r.light[0] = r.light[1] = r.light[2] = r.light[3] = 0.0f;
        if (light) {
            // 2 / linalg.length(linalg.cross(trig.u, trig.v))   shading.odin:57
            float cx = t.u[1] * t.v[2] - t.u[2] * t.v[1];
            float cy = t.u[2] * t.v[0] - t.u[0] * t.v[2];
            float cz = t.u[0] * t.v[1] - t.u[1] * t.v[0];
            float len = std::sqrt(cx * cx + cy * cy + cz * cz);
            r.light[0] = t.ng[0]; r.light[1] = t.ng[1]; r.light[2] = t.ng[2];
            r.light[3] = 2.0f / len;
        }
        out[i] = r;
    }
}

void make_pixel_to_ray_dir(const ort_camera& cam, uint32_t w, uint32_t h, float m[16]) {
    // matrix4_from_matrix3(basis) * scale(tan_x, tan_y, 1) * translate(-1,-1,1) * scale(2/w, 2/h, 1)
    // evaluated as three general 4x4 products, each entry a left-to-right sum of four products,
    // like a generic matrix multiply would; the zero / one entries make most terms exact.
    float fw = (float)w, fh = (float)h;
    float aspect = fw / fh;
    float tan_x = std::tan(cam.fov_x / 2);
    float tan_y = tan_x / aspect;
    float A[16], B[16], C[16];
    auto ident = [](float* x) { for (int i = 0; i < 16; i++) x[i] = (i % 5 == 0) ? 1.0f : 0.0f; };
    auto mul = [](const float* a, const float* b, float* c) {
        for (int r = 0; r < 4; r++)
            for (int k = 0; k < 4; k++)
                c[r * 4 + k] = a[r * 4 + 0] * b[0 * 4 + k] + a[r * 4 + 1] * b[1 * 4 + k] +
                               a[r * 4 + 2] * b[2 * 4 + k] + a[r * 4 + 3] * b[3 * 4 + k];
    };
    ident(A);
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) A[r * 4 + c] = cam.basis[3 * c + r];
    ident(B);
    B[0] = tan_x; B[5] = tan_y; B[10] = 1.0f;
    mul(A, B, C);
    ident(B);
    B[3] = -1.0f; B[7] = -1.0f; B[11] = 1.0f;
    mul(C, B, A);
    ident(B);
    B[0] = 1.0f / (fw / 2); B[5] = 1.0f / (fh / 2); B[10] = 1.0f / 1.0f;
    mul(A, B, m);
}

} // namespace ort

extern "C" int64_t ort_bvh_build(ort_triangle* tris, int64_t n, ort_bvh_node* nodes_out, int64_t cap) {
    using namespace ort;
    if (n < 0 || (n > 0 && tris == nullptr) || nodes_out == nullptr) return -1;
    if (n >= (int64_t)1 << 31) return -2;
    std::vector<Ent> ent((size_t)n), tmp((size_t)n);
    std::vector<Box> suffix((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        const ort_triangle& t = tris[i];
        Box b;
        for (int k = 0; k < 3; k++) { // aabb_of_triangle raytracer.odin:197: p, p+u, p+v
            float p0 = t.p[k], p1 = t.p[k] + t.u[k], p2 = t.p[k] + t.v[k];
            float lo = p0, hi = p0;
            lo = lo < p0 ? lo : p0; hi = hi > p0 ? hi : p0; // aabb_of_points merges points[0] too
            lo = lo < p1 ? lo : p1; hi = hi > p1 ? hi : p1;
            lo = lo < p2 ? lo : p2; hi = hi > p2 ? hi : p2;
            b.lo[k] = lo; b.hi[k] = hi;
        }
        ent[i].box = b;
        ent[i].idx = (uint32_t)i;
        ent[i].key = 0;
    }
    Build bd;
    bd.ent = ent.data(); bd.tmp = tmp.data(); bd.suffix = suffix.data();
    std::vector<ort_bvh_node> nodes;
    nodes.reserve((size_t)(n / 2 + 2));
    unsigned hw = std::thread::hardware_concurrency();
    int fork_levels = hw >= 16 ? 4 : (hw >= 8 ? 3 : (hw >= 4 ? 2 : (hw >= 2 ? 1 : 0)));
    int64_t root = bd.subtree(0, n, nodes, fork_levels);
    if (root < 0 || bd.failed) return -3;
    if ((int64_t)nodes.size() > cap) return -4;
    std::vector<ort_triangle> sorted((size_t)n);
    for (int64_t i = 0; i < n; i++) sorted[i] = tris[ent[i].idx];
    if (n) std::memcpy(tris, sorted.data(), sizeof(ort_triangle) * (size_t)n);
    std::memcpy(nodes_out, nodes.data(), sizeof(ort_bvh_node) * nodes.size());
    return (int64_t)nodes.size();
}

extern "C" int64_t ort_wide_bvh_emit(const ort_bvh_node* bvh, int64_t n_nodes, int64_t n_tris, int32_t threads,
                                     void* nodes_out, int64_t cap, int32_t* depth, int32_t* max_stack) {
    ort::WideBVH w;
    const char* why = nullptr;
    if (!ort::build_wide_bvh(bvh, n_nodes, n_tris, &w, &why, threads)) return -1;
    if (depth) *depth = w.depth;
    if (max_stack) *max_stack = w.max_stack;
    if (nodes_out) {
        if ((int64_t)w.nodes.size() > cap) return -2;
        std::memcpy(nodes_out, w.nodes.data(), w.nodes.size() * sizeof(ort::WideNode));
    }
    return (int64_t)w.nodes.size();
}
