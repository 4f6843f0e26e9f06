/* host_api.h — C hooks into the C++ host (libodinrt_host.so) so the Python test-suite can check
 * the native loader / writer against the Python stand-ins.  Not part of the drop-in boundary
 * (that is include/odinrt_b200.h); Odin keeps this stage in production. */
#ifndef ODINRT_HOST_API_H
#define ODINRT_HOST_API_H
#include "../../include/odinrt_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct orh_scene orh_scene;
/* read_gltf (+ optional --env-map). Returns 0 or 1 with a message in err. */
int  orh_scene_load(const char* gltf_path, const char* env_map_path, orh_scene** out, char* err, int err_len);
void orh_scene_free(orh_scene* s);
/* finish_scene; bvh_device < 0 = host builder */
int  orh_scene_finish(orh_scene* s, int bvh_device, char* err, int err_len);
/* ort_scene view (valid until the scene is freed or finished again) */
int  orh_scene_view(orh_scene* s, ort_scene* out);
void orh_scene_set_fov_x(orh_scene* s, float fov_x);
void orh_get_rgb_image(const ort_sample_stats* pixels, int w, int h, uint8_t* rgb_out);
int  orh_save_result(const ort_sample_stats* pixels, int w, int h, const char* path, char* err, int err_len);
#ifdef __cplusplus
}
#endif
#endif
