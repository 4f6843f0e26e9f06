// host_api.cpp — see host_api.h.
#include "host_api.h"

#include <cstdio>
#include <cstring>

#include "scene.hpp"

struct orh_scene {
    orh::HostScene scene;
    std::vector<ort_texture> tex_keep;
};

namespace {
int put_err(char* err, int n, const std::string& m) {
    if (err && n > 0) std::snprintf(err, (size_t)n, "%s", m.c_str());
    return 1;
}
} // namespace

extern "C" {

int orh_scene_load(const char* gltf_path, const char* env_map_path, orh_scene** out, char* err, int err_len) {
    if (!gltf_path || !out) return put_err(err, err_len, "orh_scene_load: NULL argument");
    *out = nullptr;
    orh_scene* s = new orh_scene();
    std::string e;
    if (!orh::read_gltf(gltf_path, &s->scene, &e)) { delete s; return put_err(err, err_len, "Failed to parse gltf: " + e); }
    if (env_map_path && *env_map_path) {
        if (!orh::load_texture(env_map_path, &s->scene.env_map, &e)) { delete s; return put_err(err, err_len, "Failed to load environment map: " + e); }
        s->scene.has_env = true;
    }
    *out = s;
    return 0;
}

void orh_scene_free(orh_scene* s) { delete s; }

int orh_scene_finish(orh_scene* s, int bvh_device, char* err, int err_len) {
    if (!s) return put_err(err, err_len, "orh_scene_finish: NULL scene");
    std::string e;
    if (!orh::finish_scene(&s->scene, bvh_device, &e)) return put_err(err, err_len, e);
    return 0;
}

int orh_scene_view(orh_scene* s, ort_scene* out) {
    if (!s || !out) return 1;
    orh::scene_view(s->scene, out, &s->tex_keep);
    return 0;
}

void orh_scene_set_fov_x(orh_scene* s, float fov_x) { if (s) s->scene.cam.fov_x = fov_x; }

void orh_get_rgb_image(const ort_sample_stats* pixels, int w, int h, uint8_t* rgb_out) {
    std::vector<uint8_t> rgb;
    orh::get_rgb_image(pixels, w, h, &rgb);
    std::memcpy(rgb_out, rgb.data(), rgb.size());
}

int orh_save_result(const ort_sample_stats* pixels, int w, int h, const char* path, char* err, int err_len) {
    std::string e;
    if (!orh::save_result(pixels, w, h, path, &e)) return put_err(err, err_len, e);
    return 0;
}

} // extern "C"
