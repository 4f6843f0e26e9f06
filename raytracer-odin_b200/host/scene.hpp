// scene.hpp — host-side `Scene` (raytracer.odin:51-60) for the C++ host, its loader
// (read_gltf, input.odin:13-259), finish_scene (raytracer.odin:62-91) and the view handed across the
// C ABI.  In production Odin owns this stage (BASELINE.json north_star); this is the native
// stand-in SURVEY §8(f) rank 3 asks for, so the CLI runs without an Odin toolchain.
#pragma once
#include <string>
#include <vector>

#include "../../include/odinrt_b200.h"
#include "image_io.hpp"

namespace orh {

struct HostScene {
    ort_camera cam{};
    std::vector<ort_triangle> triangles;  // scene.trigs[1:] (the dummy triangle 0 of input.odin:43 is never stored)
    std::vector<ort_material> materials;  // scene.materials INCLUDING the dummy material 0 (input.odin:44)
    std::vector<Texture> textures;        // de-duplicated by resolved path (input.odin:66-72), first-use order
    bool has_env = false;
    Texture env_map;                      // scene.env_map (main.odin:213-220)
    // filled by finish_scene
    std::vector<ort_bvh_node> bvh, light_bvh;
    std::vector<ort_triangle> light_triangles;
    bool finished = false;
};

// read_gltf (input.odin:13-259): JSON .gltf only (:28), external or data-URI buffers, TRIANGLES
// primitives with POSITION / NORMAL / TEXCOORD_0 / TANGENT, one Material per primitive instance.
bool read_gltf(const std::string& path, HostScene* out, std::string* err);

// finish_scene (raytracer.odin:62-91): emissive triangles are collected BEFORE the scene BVH build
// reorders the triangle array (:63-66); both BVHs come from the library's bvh_build
// (ort_bvh_build, or ort_bvh_build_device when device >= 0).
bool finish_scene(HostScene* s, int bvh_device, std::string* err);

// ort_scene view of a finished scene; `tex_keep` owns the ort_texture array the view points into.
void scene_view(const HostScene& s, ort_scene* out, std::vector<ort_texture>* tex_keep);

// get_rgb_image, mode Mean (output.odin:30-80) and save_result (output.odin:82-107).
void get_rgb_image(const ort_sample_stats* pixels, int w, int h, std::vector<uint8_t>* rgb);
bool save_result(const ort_sample_stats* pixels, int w, int h, const std::string& path, std::string* err);

} // namespace orh
