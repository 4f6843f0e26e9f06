// wide_bvh.h — flattened wide-node layout the reference BVH is re-emitted into, plus the
// traversal / shading triangle records.  Shared by the host builder (host_bvh.cpp) and the
// kernels (kernels.cu).  All records are multiples of 16 bytes so every load is a 16-byte
// vector load.
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/odinrt_b200.h"

namespace ort {

// One 128-byte node = one L1/L2 cache line = 8 x LDG.128.  Four child boxes in SoA form, laid
// out [axis][lo|hi][child] so that the near / far plane of an axis is picked with a per-ray byte
// offset (0 or 16) instead of a per-box select.
//   bytes   0.. 15 lo.x[4]    16.. 31 hi.x[4]
//          32.. 47 lo.y[4]    48.. 63 hi.y[4]
//          64.. 79 lo.z[4]    80.. 95 hi.z[4]
//          96..111 child[4]  112..127 reserved
// child >= 0  : index of an inner wide node
// child <  0  : leaf, ~child = (first_triangle << 3) | triangle_count   (count 1..4: the
//               reference's LEAF_NODE_THRESHOLD, raytracer.odin:230; triangle order inside a
//               leaf is the reference's, so "first wins" ties inside a leaf are preserved)
// child == WIDE_EMPTY : unused slot
struct alignas(16) WideNode {
    float bounds[3][2][4];
    int32_t child[4];
    int32_t reserved[4];
};
static_assert(sizeof(WideNode) == 128, "WideNode must be one cache line");
constexpr int32_t WIDE_EMPTY = (int32_t)0x80000000;

// Traversal record, 64 bytes = 2 x LDG.256: p, u, v and the ray-independent third row of the
// adjugate of [u | v | -d] (raytracer.odin:138-142): c = (uy*vz - uz*vy, -(ux*vz - uz*vx),
// ux*vy - uy*vx), each product and difference individually rounded.
struct alignas(32) TriIsect {
    float p[3], ux;
    float uy, uz, vx, vy;
    float vz, c0, c1, c2;
    float pad[4];
};
static_assert(sizeof(TriIsect) == 64, "TriIsect");

// Extra record for light triangles (all-hit pdf sum, shading.odin:52-60): ng and
// k = 2 / length(cross(u, v)).
struct alignas(16) TriLight {
    float ng[3], k;
};

// Shading records.
struct alignas(16) TriShade { // 64 bytes
    float n1[3], ngx;
    float n2[3], ngy;
    float n3[3], ngz;
    int32_t material, flags, pad0, pad1;
};
struct alignas(16) TriUV { // 32 bytes
    float tex1[2], tex2[2];
    float tex3[2], pad[2];
};
struct alignas(16) TriTan { // 48 bytes
    float tan1[4], tan2[4], tan3[4];
};

// Quantised 4-wide node, 64 bytes = 4 x LDG.128 (vs 7 for WideNode).  Used for large scenes, where
// the traversal is bound by the L1/TEX pipe (ncu: l1tex throughput 90 % on the 1 M-triangle scene).
// Child planes are 8-bit offsets from the node's own lower corner in units of a per-axis power of
// two: plane = origin + q * 2^(exp-127).  The quantisation is CONSERVATIVE (lower planes rounded
// down, upper planes up, each with an extra 2^-7 step of slack for the kernel's decode rounding),
// so the decoded boxes contain the reference's boxes: supersets cannot change the closest hit.
//   bytes  0..11 origin.xyz (f32)   12..15 exp.x | exp.y << 8 | exp.z << 16
//         16..19 lo.x[4 children]   20..23 hi.x[4]   24..27 lo.y[4]   28..31 hi.y[4]
//         32..35 lo.z[4]            36..39 hi.z[4]   40..47 reserved
//         48..63 child[4]   (same encoding as WideNode)
struct alignas(16) QuantNode {
    float origin[3];
    uint32_t exps;
    uint32_t planes[3][2]; // [axis][lo|hi], byte k = child k
    uint32_t reserved[2];
    int32_t child[4];
};
static_assert(sizeof(QuantNode) == 64, "QuantNode must be half a cache line");

// 8-wide node, 256 bytes.  Child boxes in SoA form [axis][lo|hi][8 slots] (192 bytes: the near / far
// vector of an axis is one 32-byte load picked by a per-ray offset), then a header and the per-slot
// triangle masks.  Inner children are stored contiguously in slot order: child index =
// child_base + popcount(imask & ((1 << slot) - 1)).  The triangles of all leaf children of a node are
// contiguous in the TRAVERSAL-ORDER triangle array (tris8) starting at tri_base, in slot order and, inside a
// leaf, in the reference's order; trimask[slot] has one bit per triangle of that leaf (0 for inner / empty
// slots), so a node references at most 32 triangles (8 leaves of <= 4, raytracer.odin:230).  Slots are
// assigned so that slot bit a (x: 1, y: 2, z: 4) is set for children on the positive side of the node's
// centre along axis a: visiting the hit slots in ascending order of (slot ^ octant of the ray direction)
// is a near-to-far order without any per-visit sorting (Ylitie et al. 2017).
struct alignas(32) Wide8Node {
    float bounds[3][2][8];
    uint32_t child_base, tri_base, imask, pad0;
    uint32_t pad1[4];
    uint32_t trimask[8];
};
static_assert(sizeof(Wide8Node) == 256, "Wide8Node must be two cache lines");

// 8-wide node with explicit child references (same encoding as WideNode.child), for the exact-order
// 8-wide traversal (k_trace<.., .., 2>): children sorted by entry distance at every visit, stack entries
// carry their distance, leaves are stack entries — k_trace's algorithm on nodes twice as wide.
struct alignas(32) Wide8xNode {
    float bounds[3][2][8];
    int32_t child[8];
    int32_t pad[8];
};
static_assert(sizeof(Wide8xNode) == 256, "Wide8xNode must be two cache lines");
struct Wide8xBVH {
    std::vector<Wide8xNode> nodes; // root = 0
    int depth = 0, max_stack = 0;
};
bool build_wide8x_bvh(const ort_bvh_node* bvh, int64_t n_nodes, int64_t n_tris, Wide8xBVH* out, const char** err);

struct Wide8BVH {
    std::vector<Wide8Node> nodes;     // root = 0
    std::vector<uint32_t> tri_order;  // traversal-order position -> reference triangle index
    int depth = 0;                    // = worst-case stack occupancy (one entry per level)
};
// Re-emit the binary reference BVH as an 8-wide BVH.  Returns false with *err set when the input is
// malformed or a node would reference more than 32 triangles (leaves larger than the reference's 4).
bool build_wide8_bvh(const ort_bvh_node* bvh, int64_t n_nodes, int64_t n_tris, Wide8BVH* out, const char** err);

struct WideBVH {
    std::vector<WideNode> nodes; // root = 0
    int depth = 0;               // wide levels, root = 1
    int max_stack = 0;           // exact worst-case traversal stack occupancy
    float max_abs[3] = {0, 0, 0}; // max |coordinate| of the root box (box-test padding scale)
};

// Re-emit the reference's binary post-order BVH (root = last node, raytracer.odin:375) as a
// 4-wide BVH.  Returns false (with *err set) on malformed input.
bool build_wide_bvh(const ort_bvh_node* bvh, int64_t n_nodes, int64_t n_tris, WideBVH* out, const char** err,
                    int threads = 0 /* 0 = all host cores, at most 16 */);

// Conservative 8-bit re-encoding of a WideNode array (same indices, same children).
void quantize_wide_nodes(const WideNode* in, size_t n, QuantNode* out);

void make_isect_records(const ort_triangle* tris, int64_t n, TriIsect* out);
void make_light_records(const ort_triangle* tris, int64_t n, TriLight* out);

// pixel_to_ray_dir (raytracer.odin:529-538), row-major m[r*4+c].
void make_pixel_to_ray_dir(const ort_camera& cam, uint32_t w, uint32_t h, float m[16]);

} // namespace ort
