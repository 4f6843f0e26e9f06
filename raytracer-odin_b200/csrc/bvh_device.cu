// bvh_device.cu — SURVEY §8(f) rank 1: the reference's bvh_build (raytracer.odin:227-342) on the GPU.
//
// Same algorithm, same result, bit for bit: binary SAH sweep over the triangles sorted by
// aabb.lo[axis] for each of the three axes, first minimum wins, axis chosen with the reference's
// strict-< rule (ties fall to axis 2), split until <= 4 triangles, node box = box of the axis-0
// pass, triangles permuted in place, nodes in post-order with the root last.  The reference does
// this recursively with four comparison sorts per node; here every level of the tree is processed
// at once:
//   * the four stable sorts of ALL open segments of a level are four device-wide stable radix
//     sorts on the composite key (segment id << 32 | monotone bits of lo[axis]);
//   * the sweep's prefix / suffix boxes are two segmented scans whose operator is the reference's
//     aabb_merge with Odin's select semantics (a < b ? a : b keeps the RIGHT operand on +-0 ties),
//     which is associative, so the scan reproduces the sequential fold exactly;
//   * the SAH cost is evaluated per element with the reference's operation order and the
//     first minimum per segment is a reduce-by-key on (cost, index).
// CUB (part of the CUDA toolkit) supplies the radix sort / scan-by-key / reduce-by-key primitives;
// everything specific to the builder is in the kernels below.  Parity: tests compare nodes and
// permutation with the oracle's builder byte for byte.
#include <cuda_runtime.h>

#include <cub/cub.cuh>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/odinrt_b200.h"

namespace {

struct Box6 {
    float lo[3], hi[3];
};
struct Item {
    float cost;
    uint32_t li;
};

// aabb_merge(lhs = a, rhs = b): min/max = select(l < r, l, r) / select(l > r, l, r)
struct MergeFwd {
    __host__ __device__ Box6 operator()(const Box6& a, const Box6& b) const {
        Box6 r;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            r.lo[k] = a.lo[k] < b.lo[k] ? a.lo[k] : b.lo[k];
            r.hi[k] = a.hi[k] > b.hi[k] ? a.hi[k] : b.hi[k];
        }
        return r;
    }
};
// reversed scan: acc holds the suffix (elements to the right), x the own box: merge(x, acc)
struct MergeRev {
    __host__ __device__ Box6 operator()(const Box6& acc, const Box6& x) const { return MergeFwd()(x, acc); }
};
struct MinFirst { // first minimum: lower cost, then lower index
    __host__ __device__ Item operator()(const Item& a, const Item& b) const {
        return (b.cost < a.cost || (b.cost == a.cost && b.li < a.li)) ? b : a;
    }
};

__device__ __forceinline__ float half_area(const Box6& b) { // aabb_area raytracer.odin:206
    const float sx = b.hi[0] - b.lo[0], sy = b.hi[1] - b.lo[1], sz = b.hi[2] - b.lo[2];
    return __fadd_rn(__fadd_rn(__fmul_rn(sx, sy), __fmul_rn(sy, sz)), __fmul_rn(sz, sx));
}
__device__ __forceinline__ uint32_t sort_key(float x) { // monotone; -0 folded onto +0 (ties keep their order)
    x = x + 0.0f;
    const uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void k_boxes(const float* __restrict__ puv, int64_t n, Box6* __restrict__ box, uint32_t* __restrict__ perm,
                        uint32_t* __restrict__ seg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* t = puv + 9 * i;
    Box6 b;
#pragma unroll
    for (int k = 0; k < 3; k++) { // aabb_of_triangle raytracer.odin:188-204: points p, p+u, p+v, all merged into {p, p}
        const float p0 = t[k], p1 = __fadd_rn(t[k], t[3 + k]), p2 = __fadd_rn(t[k], t[6 + k]);
        float lo = p0, hi = p0;
        lo = lo < p0 ? lo : p0; hi = hi > p0 ? hi : p0;
        lo = lo < p1 ? lo : p1; hi = hi > p1 ? hi : p1;
        lo = lo < p2 ? lo : p2; hi = hi > p2 ? hi : p2;
        b.lo[k] = lo; b.hi[k] = hi;
    }
    box[i] = b;
    perm[i] = (uint32_t)i;
    seg[i] = 0u;
}

// per-segment tables
struct Segs {
    uint32_t* begin;
    uint32_t* count;
    uint32_t* active; // 1: more than 4 triangles, still to be split
    uint32_t* node;   // tree node id of this segment
    uint32_t* chosen; // axis of the final sort
    uint32_t* split;
    float* cost;      // [3][S]
};

__global__ void k_keys(const Box6* __restrict__ box, const uint32_t* __restrict__ seg, const Segs s, int pass, int64_t n,
                       uint64_t* __restrict__ keys, uint32_t* __restrict__ pos) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t sg = seg[i];
    uint32_t low = 0u;
    if (s.active[sg]) {
        const int axis = pass < 3 ? pass : (int)s.chosen[sg];
        low = sort_key(box[i].lo[axis]);
    }
    keys[i] = ((uint64_t)sg << 32) | low;
    pos[i] = (uint32_t)i;
}
__global__ void k_gather(const Box6* __restrict__ box, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ pos,
                         int64_t n, Box6* __restrict__ box_out, uint32_t* __restrict__ perm_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = pos[i];
    box_out[i] = box[p];
    perm_out[i] = perm[p];
}
__global__ void k_reverse(const Box6* __restrict__ box, const uint32_t* __restrict__ seg, int64_t n, Box6* __restrict__ rbox,
                          uint32_t* __restrict__ rseg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    rbox[i] = box[n - 1 - i];
    rseg[i] = seg[n - 1 - i];
}
// try_axis' sweep (raytracer.odin:295-302): sah(i) = area(prefix[0..i)) * i + area(suffix[i..n)) * (n - i)
__global__ void k_cost(const Box6* __restrict__ prefix, const Box6* __restrict__ rsuffix, const uint32_t* __restrict__ seg,
                       const Segs s, int64_t n, Item* __restrict__ items) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t sg = seg[i];
    const uint32_t li = (uint32_t)i - s.begin[sg];
    float c = __int_as_float(0x7f800000);
    if (s.active[sg] && li >= 1) {
        const uint32_t cnt = s.count[sg];
        const float v = __fadd_rn(__fmul_rn(half_area(prefix[i - 1]), (float)li),
                                  __fmul_rn(half_area(rsuffix[n - 1 - i]), (float)(cnt - li)));
        if (v == v) c = v; // `sah < best_sah` is false for NaN: never selected
    }
    items[i] = Item{c, li};
}
__global__ void k_record(const Item* __restrict__ best, const Box6* __restrict__ rsuffix, Segs s, int pass, uint32_t n_seg,
                         int64_t n, Box6* __restrict__ node_box) {
    const uint32_t sg = blockIdx.x * blockDim.x + threadIdx.x;
    if (sg >= n_seg || !s.active[sg]) return;
    const Item b = best[sg];
    if (pass < 3) s.cost[(size_t)pass * n_seg + sg] = b.cost;
    if (pass == 0) node_box[s.node[sg]] = rsuffix[n - 1 - (int64_t)s.begin[sg]]; // aabb_total := buf[0] (raytracer.odin:307)
    if (pass == 2) {
        const float c0 = s.cost[sg], c1 = s.cost[(size_t)n_seg + sg], c2 = b.cost;
        s.chosen[sg] = (c0 < c1 && c0 < c2) ? 0u : ((c1 < c0 && c1 < c2) ? 1u : 2u); // raytracer.odin:311-317
    }
    if (pass == 3) s.split[sg] = b.li;
}

// Split every open segment at its SAH minimum; children with <= 4 triangles become leaves.
__global__ void k_children(const Segs s, uint32_t n_seg, const uint32_t* __restrict__ new_base,
                           const uint32_t* __restrict__ active_rank, uint32_t next_node, const Box6* __restrict__ box,
                           Segs o, uint32_t* __restrict__ node_left, uint32_t* __restrict__ node_right,
                           uint32_t* __restrict__ node_begin, uint32_t* __restrict__ node_count, Box6* __restrict__ node_box,
                           uint32_t* __restrict__ degenerate) {
    const uint32_t sg = blockIdx.x * blockDim.x + threadIdx.x;
    if (sg >= n_seg) return;
    const uint32_t nb = new_base[sg];
    if (!s.active[sg]) {
        o.begin[nb] = s.begin[sg]; o.count[nb] = s.count[sg]; o.active[nb] = 0u; o.node[nb] = s.node[sg];
        return;
    }
    const uint32_t b = s.begin[sg], cnt = s.count[sg], k = s.split[sg];
    if (k == 0u || k >= cnt) { atomicExch(degenerate, 1u); } // the reference would recurse forever
    const uint32_t kk = (k == 0u || k >= cnt) ? cnt / 2 : k;
    const uint32_t ids[2] = {next_node + 2u * active_rank[sg], next_node + 2u * active_rank[sg] + 1u};
    const uint32_t cb[2] = {b, b + kk}, cc[2] = {kk, cnt - kk};
    node_left[s.node[sg]] = ids[0];
    node_right[s.node[sg]] = ids[1];
    for (int c = 0; c < 2; c++) {
        o.begin[nb + c] = cb[c]; o.count[nb + c] = cc[c]; o.node[nb + c] = ids[c];
        o.active[nb + c] = cc[c] > 4u ? 1u : 0u;
        node_begin[ids[c]] = cb[c];
        node_count[ids[c]] = cc[c];
        node_left[ids[c]] = 0xffffffffu;
        node_right[ids[c]] = 0xffffffffu;
        if (cc[c] <= 4u) { // leaf: aabb folded from AABB_EMPTY in triangle order (raytracer.odin:243-247)
            Box6 a;
            for (int q = 0; q < 3; q++) { a.lo[q] = __int_as_float(0x7f800000); a.hi[q] = __int_as_float(0xff800000); }
            for (uint32_t i = 0; i < cc[c]; i++) a = MergeFwd()(a, box[cb[c] + i]);
            node_box[ids[c]] = a;
        }
    }
}
__global__ void k_relabel(uint32_t* __restrict__ seg, const Segs s, const uint32_t* __restrict__ new_base, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t sg = seg[i];
    uint32_t ns = new_base[sg];
    if (s.active[sg]) {
        const uint32_t cnt = s.count[sg], k = s.split[sg];
        const uint32_t kk = (k == 0u || k >= cnt) ? cnt / 2 : k;
        if ((uint32_t)i - s.begin[sg] >= kk) ns += 1u;
    }
    seg[i] = ns;
}
__global__ void k_nchild(const uint32_t* __restrict__ active, uint32_t n_seg, uint32_t* __restrict__ nchild) {
    const uint32_t sg = blockIdx.x * blockDim.x + threadIdx.x;
    if (sg < n_seg) nchild[sg] = active[sg] ? 2u : 1u;
}

// One device allocation carved into 256-byte aligned pieces (dozens of cudaMalloc / cudaFree calls
// cost more than the whole build).
struct DevBuf {
    char* base = nullptr;
    size_t cap = 0, used = 0;
    bool sizing = true; // first pass: only add up the sizes
    template <typename T>
    T* get(size_t n) {
        const size_t bytes = (std::max<size_t>(n, 1) * sizeof(T) + 255) & ~(size_t)255;
        T* p = sizing ? nullptr : (T*)(base + used);
        used += bytes;
        return p;
    }
    bool commit() {
        cap = used; used = 0; sizing = false;
        return cudaMalloc((void**)&base, cap) == cudaSuccess;
    }
    ~DevBuf() { if (base) cudaFree(base); }
};

thread_local std::string g_err;

} // namespace

extern "C" const char* ort_bvh_build_device_error(void) { return g_err.c_str(); }

extern "C" int64_t ort_bvh_build_device(int32_t device, ort_triangle* tris, int64_t n, ort_bvh_node* nodes_out, int64_t cap) {
    if (n < 0 || (n > 0 && !tris) || !nodes_out) { g_err = "bad arguments"; return -1; }
    if (n >= ((int64_t)1 << 31)) { g_err = "too many triangles"; return -2; }
    if (n <= 4) return ort_bvh_build(tris, n, nodes_out, cap); // a single leaf: nothing to parallelise
    int prev = 0;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) { g_err = "cudaSetDevice failed (no GPU: use ort_bvh_build)"; cudaGetLastError(); return -5; }
    struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{prev};
#define BCK(x)                                                                              \
    do {                                                                                    \
        cudaError_t e_ = (x);                                                               \
        if (e_ != cudaSuccess) { g_err = std::string(#x) + ": " + cudaGetErrorString(e_); return -6; } \
    } while (0)
    cudaStream_t st = nullptr;
    BCK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct KillStream { cudaStream_t s; ~KillStream() { cudaStreamDestroy(s); } } ks{st};
    DevBuf d;
    const size_t N = (size_t)n, SMAX = N + 2, NODES = 2 * N + 2;
    float* puv; Box6 *box, *box2, *prefix, *rbox, *rsuf; uint32_t *perm, *perm2, *seg, *rseg, *pos, *pos2;
    uint64_t *keys, *keys2; Item *items, *best;
    uint32_t *uniq, *nruns, *nchild, *new_base, *arank, *degenerate, *node_left, *node_right, *node_begin, *node_count;
    Box6* node_box; Segs sg[2]; void* temp = nullptr; size_t tb = 0;
    for (int pass = 0; pass < 2; pass++) {
        puv = d.get<float>(9 * N);
        box = d.get<Box6>(N); box2 = d.get<Box6>(N); prefix = d.get<Box6>(N); rbox = d.get<Box6>(N); rsuf = d.get<Box6>(N);
        perm = d.get<uint32_t>(N); perm2 = d.get<uint32_t>(N); seg = d.get<uint32_t>(N); rseg = d.get<uint32_t>(N);
        pos = d.get<uint32_t>(N); pos2 = d.get<uint32_t>(N);
        keys = d.get<uint64_t>(N); keys2 = d.get<uint64_t>(N);
        items = d.get<Item>(N); best = d.get<Item>(SMAX);
        uniq = d.get<uint32_t>(SMAX); nruns = d.get<uint32_t>(1); nchild = d.get<uint32_t>(SMAX);
        new_base = d.get<uint32_t>(SMAX); arank = d.get<uint32_t>(SMAX); degenerate = d.get<uint32_t>(1);
        for (int i = 0; i < 2; i++) {
            sg[i].begin = d.get<uint32_t>(SMAX); sg[i].count = d.get<uint32_t>(SMAX); sg[i].active = d.get<uint32_t>(SMAX);
            sg[i].node = d.get<uint32_t>(SMAX); sg[i].chosen = d.get<uint32_t>(SMAX); sg[i].split = d.get<uint32_t>(SMAX);
            sg[i].cost = d.get<float>(3 * SMAX);
        }
        node_left = d.get<uint32_t>(NODES); node_right = d.get<uint32_t>(NODES); node_begin = d.get<uint32_t>(NODES);
        node_count = d.get<uint32_t>(NODES); node_box = d.get<Box6>(NODES);
        if (pass == 0) { // temp storage for the CUB primitives (largest request); sizes do not depend on the pointers
            size_t t1 = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, t1, keys, keys2, pos, pos2, (int)n, 0, 64, st); tb = std::max(tb, t1);
            cub::DeviceScan::InclusiveScanByKey(nullptr, t1, seg, box, prefix, MergeFwd(), (int)n, cub::Equality(), st); tb = std::max(tb, t1);
            cub::DeviceReduce::ReduceByKey(nullptr, t1, seg, uniq, items, best, nruns, MinFirst(), (int)n, st); tb = std::max(tb, t1);
            cub::DeviceScan::ExclusiveSum(nullptr, t1, nchild, new_base, (int)SMAX, st); tb = std::max(tb, t1);
        }
        temp = d.get<char>(tb);
        if (pass == 0 && !d.commit()) { g_err = "out of device memory"; cudaGetLastError(); return -7; }
    }

    const int T = 256;
    auto blocks = [&](size_t m) { return (unsigned)((m + T - 1) / T); };
    const bool dbg = getenv("ORT_BVH_DEBUG") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_start = now();
    {
        // p, u, v are the first 36 bytes of the 168-byte Triangle: pack them (a strided 2-D copy of
        // 36-byte rows from pageable memory is an order of magnitude slower than this loop + one copy)
        std::vector<float> packed(9 * N);
        for (size_t i = 0; i < N; i++) std::memcpy(&packed[9 * i], tris[i].p, 36);
        BCK(cudaMemcpyAsync(puv, packed.data(), 36 * N, cudaMemcpyHostToDevice, st));
        BCK(cudaStreamSynchronize(st));
    }
    const double t_up = now();
    k_boxes<<<blocks(N), T, 0, st>>>(puv, n, box, perm, seg);
    {
        const uint32_t h_begin = 0, h_count = (uint32_t)n, h_active = 1, h_node = 0, none = 0xffffffffu;
        BCK(cudaMemcpyAsync(sg[0].begin, &h_begin, 4, cudaMemcpyHostToDevice, st));
        BCK(cudaMemcpyAsync(sg[0].count, &h_count, 4, cudaMemcpyHostToDevice, st));
        BCK(cudaMemcpyAsync(sg[0].active, &h_active, 4, cudaMemcpyHostToDevice, st));
        BCK(cudaMemcpyAsync(sg[0].node, &h_node, 4, cudaMemcpyHostToDevice, st));
        BCK(cudaMemcpyAsync(node_begin, &h_begin, 4, cudaMemcpyHostToDevice, st));
        BCK(cudaMemcpyAsync(node_count, &h_count, 4, cudaMemcpyHostToDevice, st));
        BCK(cudaMemcpyAsync(node_left, &none, 4, cudaMemcpyHostToDevice, st));
        BCK(cudaMemcpyAsync(node_right, &none, 4, cudaMemcpyHostToDevice, st));
        BCK(cudaMemsetAsync(degenerate, 0, 4, st));
        BCK(cudaStreamSynchronize(st)); // the host scalars above live on this stack frame
    }
    uint32_t n_seg = 1, n_nodes = 1, n_active = 1;
    int cur = 0, levels = 0;
    while (n_active > 0) {
        if (++levels > 4096) { g_err = "BVH deeper than 4096 levels"; return -8; }
        Segs& S = sg[cur];
        int seg_bits = 1;
        while ((1ull << seg_bits) < n_seg) seg_bits++;
        for (int pass = 0; pass < 4; pass++) {
            k_keys<<<blocks(N), T, 0, st>>>(box, seg, S, pass, n, keys, pos);
            size_t tsz = tb;
            BCK(cub::DeviceRadixSort::SortPairs(temp, tsz, keys, keys2, pos, pos2, (int)n, 0, 32 + seg_bits, st));
            k_gather<<<blocks(N), T, 0, st>>>(box, perm, pos2, n, box2, perm2);
            std::swap(box, box2);
            std::swap(perm, perm2);
            tsz = tb;
            BCK(cub::DeviceScan::InclusiveScanByKey(temp, tsz, seg, box, prefix, MergeFwd(), (int)n, cub::Equality(), st));
            k_reverse<<<blocks(N), T, 0, st>>>(box, seg, n, rbox, rseg);
            tsz = tb;
            BCK(cub::DeviceScan::InclusiveScanByKey(temp, tsz, rseg, rbox, rsuf, MergeRev(), (int)n, cub::Equality(), st));
            k_cost<<<blocks(N), T, 0, st>>>(prefix, rsuf, seg, S, n, items);
            tsz = tb;
            BCK(cub::DeviceReduce::ReduceByKey(temp, tsz, seg, uniq, items, best, nruns, MinFirst(), (int)n, st));
            k_record<<<blocks(n_seg), T, 0, st>>>(best, rsuf, S, pass, n_seg, n, node_box);
        }
        // new segmentation
        k_nchild<<<blocks(n_seg), T, 0, st>>>(S.active, n_seg, nchild);
        size_t tsz = tb;
        BCK(cub::DeviceScan::ExclusiveSum(temp, tsz, nchild, new_base, (int)(n_seg + 1), st));
        tsz = tb;
        BCK(cub::DeviceScan::ExclusiveSum(temp, tsz, S.active, arank, (int)(n_seg + 1), st));
        Segs& O = sg[cur ^ 1];
        k_children<<<blocks(n_seg), T, 0, st>>>(S, n_seg, new_base, arank, n_nodes, box, O, node_left, node_right, node_begin,
                                                node_count, node_box, degenerate);
        k_relabel<<<blocks(N), T, 0, st>>>(seg, S, new_base, n);
        uint32_t h[2] = {0, 0};
        BCK(cudaMemcpyAsync(&h[0], new_base + n_seg, 4, cudaMemcpyDeviceToHost, st)); // new number of segments
        BCK(cudaMemcpyAsync(&h[1], arank + n_seg, 4, cudaMemcpyDeviceToHost, st));    // segments split at this level
        BCK(cudaStreamSynchronize(st));
        n_nodes += 2 * h[1];
        n_seg = h[0];
        cur ^= 1;
        // how many of the new segments are still open?
        tsz = tb;
        BCK(cub::DeviceScan::ExclusiveSum(temp, tsz, sg[cur].active, arank, (int)(n_seg + 1), st));
        BCK(cudaMemcpyAsync(&n_active, arank + n_seg, 4, cudaMemcpyDeviceToHost, st));
        BCK(cudaStreamSynchronize(st));
    }
    const double t_levels = now();
    uint32_t h_deg = 0;
    BCK(cudaMemcpy(&h_deg, degenerate, 4, cudaMemcpyDeviceToHost));
    if (h_deg) { g_err = "degenerate input: no finite SAH split (the reference would not terminate)"; return -3; }
    if ((int64_t)n_nodes > cap) { g_err = "node capacity too small"; return -4; }

    // download the tree (creation order) and the permutation; renumber in post-order, root last
    std::vector<uint32_t> L(n_nodes), R(n_nodes), B(n_nodes), C(n_nodes), P(N);
    std::vector<Box6> X(n_nodes);
    BCK(cudaMemcpy(L.data(), node_left, 4ull * n_nodes, cudaMemcpyDeviceToHost));
    BCK(cudaMemcpy(R.data(), node_right, 4ull * n_nodes, cudaMemcpyDeviceToHost));
    BCK(cudaMemcpy(B.data(), node_begin, 4ull * n_nodes, cudaMemcpyDeviceToHost));
    BCK(cudaMemcpy(C.data(), node_count, 4ull * n_nodes, cudaMemcpyDeviceToHost));
    BCK(cudaMemcpy(X.data(), node_box, sizeof(Box6) * (size_t)n_nodes, cudaMemcpyDeviceToHost));
    BCK(cudaMemcpy(P.data(), perm, 4ull * N, cudaMemcpyDeviceToHost));
#undef BCK
    std::vector<int64_t> post(n_nodes, -1);
    {
        int64_t next = 0;
        std::vector<std::pair<uint32_t, int>> stack; // (node, state): 0 = enter, 1 = left done, 2 = both done
        stack.push_back({0u, 0});
        while (!stack.empty()) {
            auto& top = stack.back();
            const uint32_t nd = top.first;
            if (L[nd] == 0xffffffffu) { post[nd] = next++; stack.pop_back(); continue; }
            if (top.second == 0) { top.second = 1; stack.push_back({L[nd], 0}); }
            else if (top.second == 1) { top.second = 2; stack.push_back({R[nd], 0}); }
            else { post[nd] = next++; stack.pop_back(); }
        }
    }
    for (uint32_t nd = 0; nd < n_nodes; nd++) {
        ort_bvh_node o;
        std::memset(&o, 0, sizeof o);
        std::memcpy(o.lo, X[nd].lo, 12);
        std::memcpy(o.hi, X[nd].hi, 12);
        if (L[nd] == 0xffffffffu) { o.kind = 0; o.a = B[nd]; o.b = C[nd]; }
        else { o.kind = 1; o.a = post[L[nd]]; o.b = post[R[nd]]; }
        nodes_out[post[nd]] = o;
    }
    {
        std::vector<ort_triangle> sorted(N);
        const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < hw; t++)
            pool.emplace_back([&, t] {
                for (size_t i = N * t / hw; i < N * (t + 1) / hw; i++) sorted[i] = tris[P[i]];
            });
        for (auto& th : pool) th.join();
        pool.clear();
        for (unsigned t = 0; t < hw; t++)
            pool.emplace_back([&, t] {
                const size_t a = N * t / hw, b = N * (t + 1) / hw;
                std::memcpy(tris + a, sorted.data() + a, sizeof(ort_triangle) * (b - a));
            });
        for (auto& th : pool) th.join();
    }
    if (dbg)
        fprintf(stderr, "[bvh_device] n=%lld levels=%d nodes=%u: upload %.1f ms, levels %.1f ms, download+renumber+permute %.1f ms\n",
                (long long)n, levels, n_nodes, t_up - t_start, t_levels - t_up, now() - t_levels);
    return (int64_t)n_nodes;
}
