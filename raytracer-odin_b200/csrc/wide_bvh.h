// wide_bvh.h — flattened wide-node layout the reference BVH is re-emitted into, plus the
// traversal / shading triangle records.  Shared by the host builder (host_bvh.cpp) and the
// kernels (kernels.cuh, traverse.cuh).  All records are multiples of 16 bytes so every load is a 16-byte
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

// Light-candidate queue entry (k_shade -> k_trace<light>): the ray's position in the ray queue in the low 28 bits,
// the mask of the light root's children whose boxes the ray enters in the top four (0 = start at the root).
// A wave therefore holds at most 2^28 paths (the default is 2^25).
constexpr uint32_t LQ_MASK_SHIFT = 28u;
constexpr uint32_t LQ_POS_MASK = (1u << LQ_MASK_SHIFT) - 1u;

// Traversal record, 64 bytes = 2 x LDG.256: p, u, v and the ray-independent third row of the
// adjugate of [u | v | -d] (raytracer.odin:138-142): c = (uy*vz - uz*vy, -(ux*vz - uz*vx),
// ux*vy - uy*vx), each product and difference individually rounded.  The last 16 bytes are used by the
// records of the LIGHT triangles only (all-hit pdf sum, shading.odin:52-60): ng and 2 / length(cross(u, v)) arrive
// with the second half of the record instead of through a dependent load after a hit (that load was 10 % of the
// light pass's stall samples at 2 active lanes: profiles/r2s_light_pass.md).
struct alignas(32) TriIsect {
    float p[3], ux;
    float uy, uz, vx, vy;
    float vz, c0, c1, c2;
    float light[4]; // ng.xyz, 2 / |u x v|   (zero for scene triangles)
};
static_assert(sizeof(TriIsect) == 64, "TriIsect");

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

// Worst-case occupancy of the REFERENCE's traversal stack (sa.Small_Array(64, int), raytracer.odin:379) on
// this binary BVH: a both-hit branch pops one entry and appends three (:396-409), so a ray that hits both
// children along the deepest root-to-leaf path holds 2 * (branches on that path) + 1 entries.  Above 64
// the reference silently drops pushes; this library never does (ort_stats.reference_stack_need).
int64_t reference_stack_need(const ort_bvh_node* bvh, int64_t n_nodes);

void make_isect_records(const ort_triangle* tris, int64_t n, TriIsect* out, bool light = false);

// pixel_to_ray_dir (raytracer.odin:529-538), row-major m[r*4+c].
void make_pixel_to_ray_dir(const ort_camera& cam, uint32_t w, uint32_t h, float m[16]);

} // namespace ort
