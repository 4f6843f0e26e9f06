/* odinrt_b200.h — C ABI of libodinrt_b200.so
 *
 * Drop-in boundary for the per-pixel path-tracing loop of elteammate/raytracer-odin.
 * The reference has no FFI seam on this path; the cut is the single call
 *     render_scene(rc, &scene, number_of_trials)          main.odin:242 -> raytracer.odin:602
 * Everything before it (glTF load input.odin:13, finish_scene raytracer.odin:62 incl. both
 * bvh_build calls raytracer.odin:227) and after it (save_result output.odin:82) stays on the host
 * side (Odin).  The structs below are plain C mirrors of what the Odin side owns at that point;
 * INTEGRATION.md shows the `foreign import` shim that fills them.
 *
 * Conventions: every function returns 0 on success, non-zero on error (ort_last_error() holds
 * the message; the reference panics on load errors main.odin:195,216 and has no recoverable
 * render error — the shim panics on non-zero).  The caller owns every host buffer passed in; the
 * library owns all device memory.  No pointer into caller memory is kept after ort_upload_scene
 * returns, except `out` / `interrupt` for the duration of a blocking render call.
 * No C++ exception or torch type crosses this boundary.
 */
#ifndef ODINRT_B200_H
#define ODINRT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORT_ABI_VERSION 2

/* ---- scene data handed over by the host (once per scene) --------------------------------- */

/* Triangle — layout-identical to the Odin `Triangle` (raytracer.odin:18-23), 168 bytes:
 * p0 u12 v24 n1 36 n2 48 n3 60 ng72 tex1 84 tex2 92 tex3 100 tan1 108 tan2 124 tan3 140
 * material_index(i64) 160.  Triangles are passed in POST-BUILD order (bvh_build sorts
 * scene.trigs[1:] in place, raytracer.odin:265-268); the dummy at index 0 (input.odin:43) is NOT
 * passed, so triangle id == index into this array. */
typedef struct ort_triangle {
    float p[3], u[3], v[3];
    float n1[3], n2[3], n3[3], ng[3];
    float tex1[2], tex2[2], tex3[2];
    float tan1[4], tan2[4], tan3[4];
    int64_t material_index; /* index into ort_scene.materials; Odin's dummy material 0 IS passed */
} ort_triangle;

/* BVH node — C mirror of the Odin `BVH_Node` (raytracer.odin:211-225): binary, post-order,
 * root == last element (raytracer.odin:375,380).  kind 0 = leaf: a = first triangle index,
 * b = triangle count (slice ptr - base)/168, len);  kind 1 = branch: a = left, b = right. */
typedef struct ort_bvh_node {
    float lo[3], hi[3];
    int32_t kind;
    int32_t _pad;
    int64_t a, b;
} ort_bvh_node;

/* Texture — mirror of `Texture` (textures.odin:14-19).  data is u8 or f32 (stb native channel
 * count, 1..4); stride = elements per row (channels * width, textures.odin:65). */
typedef struct ort_texture {
    const void* data;
    int32_t width, height;
    int32_t channels;
    int32_t is_f32; /* 0: [^]u8, 1: [^]f32 */
    int64_t stride;
} ort_texture;

/* Material — mirror of `Material` (raytracer.odin:34-43).  Samplers become indices into
 * ort_scene.textures, -1 == nil sampler. */
typedef struct ort_material {
    float color_factor[3];
    int32_t color_texture;
    float emission_factor[3];
    int32_t emission_texture;
    float metallic_factor;
    float roughness_factor;
    int32_t metallic_roughness_texture;
    int32_t normal_texture;
} ort_material;

/* Cam — mirror of `Cam` (raytracer.odin:45-49). basis is column-major: basis[3*c + r] is row r
 * of column c (Odin matrix[3,3]f32; input.odin:105-107).  fov_x already carries the aspect
 * multiplication of main.odin:202-203. */
typedef struct ort_camera {
    float pos[3];
    float basis[9];
    float fov_x;
} ort_camera;

/* Scene — mirror of `Scene` (raytracer.odin:51-60) after finish_scene (raytracer.odin:62-91). */
typedef struct ort_scene {
    ort_camera cam;
    const ort_triangle* triangles;      /* scene.trigs[1:], post-build order */
    int64_t n_triangles;
    const ort_bvh_node* bvh;            /* scene.bvh */
    int64_t n_bvh_nodes;
    const ort_triangle* light_triangles; /* scene.light_surfaces, post-build order (raytracer.odin:75) */
    int64_t n_light_triangles;
    const ort_bvh_node* light_bvh;      /* scene.light_bvh */
    int64_t n_light_bvh_nodes;
    const ort_material* materials;      /* scene.materials incl. dummy 0 (input.odin:44) */
    int64_t n_materials;
    const ort_texture* textures;        /* de-duplicated textures (input.odin:250-256) */
    int64_t n_textures;
    const ort_texture* env_map;         /* scene.env_map (main.odin:213-220) or NULL */
} ort_scene;

/* ---- render output ------------------------------------------------------------------------ */

/* Sample_Stats — layout-identical to main.odin:34-40, 52 bytes.  Pixel (x,y) lives at index
 * (H-1-y)*W + x (rc_set_pixel, main.odin:95). */
typedef struct ort_sample_stats {
    float first[3];
    uint32_t count;
    float last[3];
    float total[3];
    float total_squared[3];
} ort_sample_stats;

typedef struct ort_ray {
    float o[3];
    float d[3];
} ort_ray; /* Ray, raytracer.odin:105-107 */

/* Result of cast_ray (raytracer.odin:416-430).  tri == -1 is the reference's `trig == nil`. */
typedef struct ort_hit {
    float t;    /* includes the + RAY_EPS of raytracer.odin:428 */
    float u, v; /* hit.uv */
    int32_t tri;      /* index into ort_scene.triangles, -1 on miss */
    int32_t material; /* triangles[tri].material_index ("primitive id"), -1 on miss */
    int32_t inside;   /* hit.inside */
} ort_hit;

typedef struct ort_device_cfg {
    int32_t device;     /* CUDA device ordinal */
    int32_t _pad;
    uint64_t seed;      /* key of the counter-based per-pixel RNG streams */
    int64_t max_paths_in_flight; /* 0 = default; path-state capacity of one wave */
    int64_t max_path_bytes;      /* 0 = no limit; otherwise the library behaves as if only this much HBM were
                                    free for path state (smaller waves, fewer pipelines, then an error) */
} ort_device_cfg;

typedef struct ort_stats {
    uint64_t rays_closest;   /* cast_ray calls the reference would make (raytracer.odin:435):
                                primary rays + continuations that pass `norm_l1(value)/pdf > 1e-5` */
    uint64_t rays_traced;    /* closest-hit traversals actually launched: the continuation ray is
                                traced together with its light-pdf sum before that test can be
                                evaluated, so this is >= rays_closest (speculative rays) */
    uint64_t rays_light_pdf; /* surface_sampling_pdf traversals launched (shading.odin:98) */
    uint64_t paths;          /* pixel-samples completed */
    uint64_t kernel_launches;/* launches of this library's own kernels */
    double   render_ms;      /* device time of the last render call (CUDA events) */
    double   trace_ms;       /* device time spent in closest-hit trace kernels, last profiled call */
    double   light_ms;
    double   shade_ms;
    double   other_ms;
    /* wide-BVH facts fixed at upload time */
    int64_t  wide_nodes, wide_depth, light_wide_nodes;
    int64_t  device_bytes;
    int64_t  wide_max_stack;       /* exact worst-case stack occupancy of this library's traversal (never dropped) */
    int64_t  reference_stack_need; /* worst-case occupancy of the REFERENCE's 64-entry stack on the same binary BVH
                                      (2 * branch depth + 1, raytracer.odin:379,396-409): above 64 the reference
                                      silently drops pushes and may miss hits this library finds */
} ort_stats;

typedef struct ort_ctx ort_ctx;

/* ---- entry points ------------------------------------------------------------------------- */

int  ort_abi_version(void);

/* Create / destroy one rendering context bound to one GPU.  Fails (non-zero) when no sm_100
 * class device is present: there is no CPU fallback. */
int  ort_create(ort_ctx** out, const ort_device_cfg* cfg);
void ort_destroy(ort_ctx* ctx);
const char* ort_last_error(const ort_ctx* ctx); /* ctx may be NULL: last error of ort_create */

/* Run all subsequent work of this context on an externally owned cudaStream_t (e.g. torch's
 * current stream).  The handle is used as given — 0 is CUDA's legacy default stream — and
 * ORT_OWN_STREAM returns to the context's own non-blocking stream. */
#define ORT_OWN_STREAM ((void*)(intptr_t)-1)
int  ort_set_stream(ort_ctx* ctx, void* cuda_stream);

/* Deep-copies the scene to the device: re-emits the reference BVHs (scene + light) as flattened
 * wide-node layouts, builds traversal/shading triangle records, material table and texture
 * objects.  Replaces: the scene hand-off into render_scene (raytracer.odin:602). */
int  ort_upload_scene(ort_ctx* ctx, const ort_scene* scene);

/* Blocking render of samples [first_sample, first_sample + n_samples) for every pixel of a
 * w x h image, accumulated INTO `out` (w*h ort_sample_stats, host memory), like render_scene
 * accumulates into rc.pixels[0] across --times trials without clearing (raytracer.odin:606-610).
 * `interrupt` (may be NULL) is polled between waves like is_interrupted() (raytracer.odin:554);
 * on interrupt the samples finished so far are written and 0 is returned.
 * Replaces: render_scene / render_task / raytrace / rc_set_pixel
 * (raytracer.odin:602,528,432; main.odin:89). */
int  ort_render(ort_ctx* ctx, uint32_t w, uint32_t h, int32_t ray_depth,
                uint64_t first_sample, uint64_t n_samples,
                ort_sample_stats* out, const volatile uint8_t* interrupt);

/* Same render, device-resident: accumulates into a caller-owned DEVICE buffer of 8*w*h floats,
 * planar: total.r,g,b | total_squared.r,g,b | count_lo | count_hi, each plane w*h in
 * ort_sample_stats pixel order ((H-1-y)*W+x); count = count_lo + 2^20 * count_hi, both exact integers
 * in f32 (count_lo < 2^20 after every wave), so a float sum-reduce over the GPUs of a box keeps the u32
 * count of Sample_Stats (main.odin:36) exact.  Asynchronous on the context stream.  This is the buffer
 * multi-GPU hosts reduce (one NCCL reduce per frame) before unpacking. */
int  ort_render_device(ort_ctx* ctx, uint32_t w, uint32_t h, int32_t ray_depth,
                       uint64_t first_sample, uint64_t n_samples, float* d_accum);

/* Unpack a planar device accumulator into host Sample_Stats (adds to `out`). */
int  ort_unpack_accum(ort_ctx* ctx, uint32_t w, uint32_t h, const float* d_accum,
                      ort_sample_stats* out);

/* Parity probes.  ort_trace_rays == cast_ray (raytracer.odin:416) on n host rays;
 * ort_primary_hits generates the primary ray of `sample` for every pixel exactly as render does
 * (render_task, raytracer.odin:580-593) and returns its cast_ray result, pixel order y*w + x
 * (unflipped). */
int  ort_trace_rays(ort_ctx* ctx, const ort_ray* rays, int64_t n, ort_hit* out);
int  ort_primary_hits(ort_ctx* ctx, uint32_t w, uint32_t h, uint64_t sample, ort_hit* out,
                      ort_ray* rays_out /* may be NULL */);
/* Sum of surface_sampling_pdf_bvh_sum (shading.odin:62-94) / n_lights for n host rays. */
int  ort_light_pdf(ort_ctx* ctx, const ort_ray* rays, int64_t n, float* out);

/* Tone-map on device: get_rgb_image (output.odin:30-80, mode Mean) from a planar device
 * accumulator to w*h*3 bytes (host). SURVEY §8(f) rank 2. */
int  ort_tonemap_rgb8(ort_ctx* ctx, uint32_t w, uint32_t h, const float* d_accum, uint8_t* out_rgb);

/* Samples per pixel the last ort_render / ort_render_device / ort_frame_render call of this context
 * completed (== n_samples unless `interrupt` fired between waves, raytracer.odin:554). */
uint64_t ort_last_render_samples(const ort_ctx* ctx);

/* ---- device-resident frame (--continious, main.odin:207; live preview, debug.odin:80) ----------
 * ort_render moves 52 bytes per pixel to the host on every call.  A frame keeps the accumulators
 * (total, total_squared, count, first, last) in HBM across calls:
 *   ort_frame_begin   allocates and clears them for a w x h image (discarding a previous frame);
 *   ort_frame_load    (optional) seeds them from host Sample_Stats, e.g. a resumed checkpoint;
 *   ort_frame_render  enqueues samples [first_sample, first_sample + n_samples) and returns without
 *                     waiting for the device (with `interrupt` it stays a few waves ahead at most and
 *                     stops enqueuing once the flag is set); *done (may be NULL) = samples enqueued;
 *   ort_frame_wait    blocks until everything enqueued has been rendered;
 *   ort_frame_snapshot  copies the accumulators, as of everything enqueued so far, into the frame's second
 *                     buffer (device to device, asynchronous): samples enqueued AFTER it render while the
 *                     snapshot is tone-mapped and read back;
 *   ort_frame_preview_rgb8  get_rgb_image (output.odin:30-80) of the pending snapshot (takes one if there
 *                     is none) on the device, 3 bytes per pixel to the host; waits for the snapshot only;
 *   ort_frame_fetch   OVERWRITES `out` with the frame's Sample_Stats (52 bytes per pixel) — once, at
 *                     exit or at a checkpoint;
 *   ort_frame_end     releases the frame. */
int  ort_frame_begin(ort_ctx* ctx, uint32_t w, uint32_t h);
int  ort_frame_load(ort_ctx* ctx, const ort_sample_stats* in);
int  ort_frame_render(ort_ctx* ctx, int32_t ray_depth, uint64_t first_sample, uint64_t n_samples,
                      const volatile uint8_t* interrupt, uint64_t* done);
int  ort_frame_wait(ort_ctx* ctx);
int  ort_frame_snapshot(ort_ctx* ctx);
int  ort_frame_preview_rgb8(ort_ctx* ctx, uint8_t* out_rgb);
int  ort_frame_fetch(ort_ctx* ctx, ort_sample_stats* out);
int  ort_frame_end(ort_ctx* ctx);

/* Parity probe for the shading device functions: runs n independent evaluations of the SAME __device__
 * functions k_shade calls, one per thread.  `in` / `out` are n records of the floats listed below
 * (u32 values are passed as their bit patterns).  Kinds that need scene data (lights, textures,
 * environment map) use the uploaded scene.
 *   ORT_PROBE_SHADE        in: n[3] color[3] metallic roughness in_d[3] out_d[3]   out: value[3]    shade, shading.odin:164-204
 *   ORT_PROBE_VNDF_SAMPLE  in: n[3] omega[3] alpha u1 u2                             out: h[3]        vndf_sampling, shading.odin:102-122
 *   ORT_PROBE_VNDF_PDF     in: n[3] omega[3] alpha L[3]                              out: pdf         vndf_sampling_pdf, shading.odin:124-137
 *   ORT_PROBE_SAMPLE       in: n[3] pos[3] roughness in_d[3] r[4](u32)               out: dir[3]      sample, shading.odin:139-151
 *   ORT_PROBE_PDF          in: n[3] pos[3] roughness in_d[3] out_d[3]                out: pdf         pdf, shading.odin:153-162 (incl. the light-BVH sum)
 *   ORT_PROBE_TEXTURE      in: texture(i32 bits; -1 = environment map) srgb u v      out: rgba[4]     texture_sample, textures.odin:79-135
 *   ORT_PROBE_COSINE       in: n[3] r1(u32) r2(u32)                                  out: dir[3] pdf  cosine_weighted(_pdf), shading.odin:17-39
 *   ORT_PROBE_ENV          in: d[3]                                                  out: rgb[3]      environment lookup, raytracer.odin:437-446 */
enum {
    ORT_PROBE_SHADE = 0, ORT_PROBE_VNDF_SAMPLE = 1, ORT_PROBE_VNDF_PDF = 2, ORT_PROBE_SAMPLE = 3,
    ORT_PROBE_PDF = 4, ORT_PROBE_TEXTURE = 5, ORT_PROBE_COSINE = 6, ORT_PROBE_ENV = 7
};
int  ort_probe_shading(ort_ctx* ctx, int32_t kind, const float* in, int64_t n, float* out);

/* Diagnostic: time the traversal kernel alone.  Uploads n host rays once, launches the closest-hit
 * (mode 0) or light-sum (mode 1) kernel `iters` times over them and returns the mean device time of
 * one launch in milliseconds (CUDA events).  Used by tools/trace_bench.py to compare kernel variants. */
int  ort_bench_trace(ort_ctx* ctx, const ort_ray* rays, int64_t n, int32_t mode, int32_t iters, double* ms_per_launch);

/* Diagnostic: streaming-read bandwidth of a working set of `bytes` (256-bit read-only loads from every
 * SM, `iters` passes after one warm-up pass), in GB/s.  With a working set between the total L1 size
 * and the L2 size this is the box's L2 read bandwidth — the peak SURVEY §8(d) asks the build to measure
 * for the L2-resident scenes (C1-C4); with a multi-GB working set it is the HBM read bandwidth. */
int  ort_bench_read_bw(ort_ctx* ctx, int64_t bytes, int32_t iters, double* gb_per_s);

int  ort_get_stats(ort_ctx* ctx, ort_stats* out);
int  ort_reset_stats(ort_ctx* ctx);
/* When on, render calls time each kernel class with CUDA events (serialises the pipeline). */
int  ort_set_profiling(ort_ctx* ctx, int32_t on);

/* ---- several GPUs of one box from ONE host process (the Odin host is a single process) ------
 * Every listed device holds a full scene replica and renders a contiguous block of the sample
 * range (counter-based RNG keyed by the global sample index: any split renders the same sample
 * set).  The partial accumulators are combined once per call by a reduce kernel on devices[0]
 * that reads the other GPUs' accumulators directly through NVLink peer memory (staged
 * cudaMemcpyPeer when peer access is unavailable), then merged into `out` like ort_render.
 * The reference's analogue is its thread pool pulling tasks from one atomic counter
 * (raytracer.odin:551,609-623). */
typedef struct ort_multi ort_multi;
int  ort_multi_create(ort_multi** out, const int32_t* devices, int32_t n_devices, uint64_t seed);
void ort_multi_destroy(ort_multi* m);
const char* ort_multi_last_error(const ort_multi* m); /* m may be NULL: last error of ort_multi_create */
int  ort_multi_upload_scene(ort_multi* m, const ort_scene* scene);
int  ort_multi_render(ort_multi* m, uint32_t w, uint32_t h, int32_t ray_depth,
                      uint64_t first_sample, uint64_t n_samples,
                      ort_sample_stats* out, const volatile uint8_t* interrupt);
int  ort_multi_get_stats(ort_multi* m, ort_stats* out); /* counters summed, times = max over devices */
/* Samples per pixel the last ort_multi_render / ort_multi_frame_render completed, summed over the GPUs
 * (== n_samples unless interrupted; an interrupted call still consumes its whole index range, so no
 * sample index is ever rendered twice). */
uint64_t ort_multi_last_render_samples(const ort_multi* m);
/* Device-resident frame on several GPUs: every GPU keeps its own accumulators and renders its block of
 * each ort_multi_frame_render call; preview / fetch snapshot them on every GPU, and devices[0] sums the
 * snapshots through NVLink peer memory on a side stream WHILE the next samples render. */
int  ort_multi_frame_begin(ort_multi* m, uint32_t w, uint32_t h);
int  ort_multi_frame_load(ort_multi* m, const ort_sample_stats* in);
int  ort_multi_frame_render(ort_multi* m, int32_t ray_depth, uint64_t first_sample, uint64_t n_samples,
                            const volatile uint8_t* interrupt, uint64_t* done);
int  ort_multi_frame_wait(ort_multi* m);
int  ort_multi_frame_snapshot(ort_multi* m);
int  ort_multi_frame_preview_rgb8(ort_multi* m, uint8_t* out_rgb);
int  ort_multi_frame_fetch(ort_multi* m, ort_sample_stats* out);
int  ort_multi_frame_end(ort_multi* m);

/* Host-side scene finalisation helper (NOT used when Odin is the host): the reference
 * bvh_build (raytracer.odin:227-342) as native code, for hosts that do not have one.
 * Sorts `tris` in place like the reference; writes up to `cap` nodes, returns node count
 * (or a negative error code). */
int64_t ort_bvh_build(ort_triangle* tris, int64_t n, ort_bvh_node* nodes_out, int64_t cap);

/* Host-side helper: the 4-wide, 128-byte-node re-emission of a binary reference BVH that
 * ort_upload_scene performs internally (breadth-first, root = node 0; layout in csrc/wide_bvh.h), on
 * `threads` host threads (0 = all cores; the output does not depend on the count).  For tools and
 * tests; returns the number of wide nodes (written to nodes_out when it has room for them), -1 on
 * malformed input, -2 when cap is too small. */
int64_t ort_wide_bvh_emit(const ort_bvh_node* bvh, int64_t n_nodes, int64_t n_tris, int32_t threads,
                          void* nodes_out, int64_t cap, int32_t* depth, int32_t* max_stack);

/* The same builder on the GPU: identical splits, permutation and post-order
 * node array as ort_bvh_build / the reference, every tree level processed at once (device-wide
 * stable radix sorts on (segment, lo[axis]) keys, segmented box scans, per-segment arg-min).
 * Host buffers in, host buffers out; returns the node count or a negative error
 * (ort_bvh_build_device_error() holds the message). */
int64_t ort_bvh_build_device(int32_t device, ort_triangle* tris, int64_t n, ort_bvh_node* nodes_out, int64_t cap);
const char* ort_bvh_build_device_error(void);

#ifdef __cplusplus
}
#endif
#endif /* ODINRT_B200_H */
