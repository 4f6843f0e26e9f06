// odinrt.cu — context, scene upload, wavefront scheduling and the C ABI of libodinrt_b200.so
// (include/odinrt_b200.h).  Replaces render_scene / render_task (raytracer.odin:528-665): the
// atomic tile counter and OS threads become waves of (pixel x sample) paths advanced one bounce
// at a time by k_trace (closest hit + fused light-pdf sum) / k_shade, with queue sizes living on the device so a whole wave
// is enqueued without any host synchronisation.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "kernels.cuh"

using namespace ort;

namespace {
thread_local std::string g_create_error;
}

constexpr int MAX_PIPES = 8;

struct ort_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    uint64_t seed = 0;
    int64_t capacity_cfg = 0;
    std::string err;

    // scene
    bool has_scene = false;
    SceneDev sd{};
    ort_camera cam{};
    // Scene memory is pooled: buffers only ever grow and texture arrays are recycled by size, so
    // re-uploading a scene of the same shape performs no cudaMalloc / cudaFree (a cudaFree was
    // measured at 15-600 ms on B200 boxes: it synchronises the device and unmaps).
    struct DevBuf { void* p = nullptr; size_t cap = 0, used = 0; };
    enum { SB_NODES, SB_TRIS, SB_MATS, SB_TSHADE, SB_TUV, SB_TTAN, SB_TEXS, SB_COUNT };
    DevBuf sbuf[SB_COUNT];
    struct TexSlot { cudaArray_t arr = nullptr; cudaTextureObject_t obj = 0; size_t w = 0, h = 0; bool in_use = false; };
    std::vector<TexSlot> tex_pool;
    // pinned staging ring for host -> device uploads (records are built straight into it)
    static constexpr int STAGE_SLOTS = 3;
    static constexpr size_t STAGE_BYTES = (size_t)8 << 20;
    char* stage = nullptr;
    cudaEvent_t stage_ev[STAGE_SLOTS] = {};
    int stage_next = 0;
    int host_threads = 1;
    int64_t n_tris = 0, n_ltris = 0;
    int64_t wide_nodes = 0, wide_depth = 0, wide_max_stack = 0, lwide_nodes = 0;
    int64_t ref_stack_need = 0; // worst-case occupancy of the REFERENCE's 64-entry stack on this scene's binary BVH
    int64_t scene_bytes = 0;

    // path buffers: two independent wave pipelines (ps[1] only when waves are overlapped)
    struct PathSet {
        float4 *qo[2] = {nullptr, nullptr}, *qd[2] = {nullptr, nullptr};
        float4* hits = nullptr;
        float* lsum = nullptr;  // 2 x capacity floats: ping-pong by bounce parity
        uint32_t* lq = nullptr; // light-candidate queue (positions into the ray queue)
        float4 *pa[2] = {nullptr, nullptr}, *pb[2] = {nullptr, nullptr}; // pending (T, value, pdf terms), queue order
        float4* st_c = nullptr;                                           // accumulated radiance by path slot
        uint32_t* counters = nullptr; // 5 arrays of D+2: queue counts, trace / light work counters, used-ray counts, light-queue counts
        int64_t capacity = 0;
        int counters_depth = 0;
        cudaEvent_t resolved = nullptr; // recorded after this pipeline's k_resolve + k_stats
        bool used = false;              // `resolved` has been recorded at least once
    } ps[MAX_PIPES];
    cudaStream_t aux_stream[MAX_PIPES] = {}; // pipelines 1.. (0 runs on `stream`)
    cudaEvent_t ev_fork = nullptr, ev_join[MAX_PIPES] = {};
    int overlap = 4;                    // number of overlapped wave pipelines (env ORT_OVERLAP=1..8)
    unsigned long long* d_stats = nullptr;
    int64_t path_bytes = 0;
    int64_t max_path_bytes = 0;         // ort_device_cfg.max_path_bytes: pretend only this much HBM is free for path state
    uint64_t last_done = 0;             // samples per pixel the last render call completed
    uint64_t wave_seq = 0;              // waves enqueued so far (picks the pipeline; persistent across ort_frame_render calls)
    cudaEvent_t last_resolved = nullptr; // `resolved` of the most recently enqueued wave: the next accumulation waits for it
    bool pipes_busy = false;            // chained waves may still be running on aux streams that ctx->stream has not joined
    bool need_fork = true;              // ctx->stream holds work the other pipelines must wait for before their next wave
    int chain_pipes = 0;

    // device-resident frame (ort_frame_*): 8 accumulator planes + 3 first + 3 last, w*h floats each
    float* frame = nullptr;
    size_t frame_bytes = 0;
    uint32_t frame_w = 0, frame_h = 0;
    bool frame_has_first = false;       // the `first` planes are already written (first wave done, or loaded)
    bool frame_snapshot_valid = false;  // ort_frame_snapshot was taken and not consumed yet
    cudaStream_t side_stream = nullptr; // preview / fetch of the snapshot run here, next to the render pipelines
    cudaEvent_t snap_ev = nullptr;
    void* side_buf = nullptr;           // packed Sample_Stats / RGB8 staging of the side stream
    size_t side_bytes = 0;

    // scratch for host-facing calls
    float* scratch = nullptr;
    size_t scratch_bytes = 0;
    void* pinned = nullptr;
    size_t pinned_bytes = 0;

    int trace_grid[2] = {0, 0}; // persistent grid sizes: closest hit, light sum
    int shade_grid = 0;
    // tuning knobs: fixed at their measured optimum in the shipped library; a -DORT_TUNING build
    // (make variant) reads them from the environment for tools/tune.py
    int tiled = 2;           // 0 = primary rays in pixel order, 1 = 8x4 pixel tiles, 2 = tile_w x tile_h pixels x tile_s samples per warp
    int tile_w = 2, tile_h = 2, tile_s = 8;
    int bin_octants = 1;     // 0: plain per-warp queue compaction (no direction-octant binning)
    int light_prefilter = 2; // 0: send every continuation ray through the light pass; 1: only rays that enter a child
                             // box of the light root; 2: those, starting the light pass AT the entered children
    int refill = ORT_REFILL_THRESHOLD; // dynamic-fetch threshold of the closest-hit pass
    int refill_light = ORT_REFILL_LIGHT; // ... of the light pass
    int inner_min = ORT_INNER_MIN;     // inner-loop early-exit threshold
    bool profiling = false;
    double ms_trace = 0, ms_light = 0, ms_shade = 0, ms_other = 0, ms_render = 0;
    uint64_t launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evp0 = nullptr, evp1 = nullptr;
};

namespace {

int fail(ort_ctx* c, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return 1;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, std::string(#call) + ": " + cudaGetErrorString(e_));                  \
    } while (0)

// No C++ exception crosses the C ABI (std::thread, std::vector can throw): report it like any other error.
template <typename F>
int guarded(ort_ctx* ctx, F f) {
    try {
        return f();
    } catch (const std::exception& e) {
        return fail(ctx, std::string("exception: ") + e.what());
    } catch (...) {
        return fail(ctx, "unknown C++ exception");
    }
}

// The 4-wide re-emission of the scene and light BVHs (+ the reference's own worst-case stack need), built
// once per upload — and once for ALL devices of an ort_multi.
struct SharedWide {
    WideBVH wide, lwide;
    bool ok_scene = false, ok_light = false;
    const char* why_scene = "not built";
    const char* why_light = "not built";
    int64_t ref_stack_need = 0;
    std::mutex mu;
    std::condition_variable cv;
    bool ready = false;
    void build(const ort_scene* sc) {
        try {
            ok_scene = build_wide_bvh(sc->bvh, sc->n_bvh_nodes, sc->n_triangles, &wide, &why_scene);
            ok_light = build_wide_bvh(sc->light_bvh, sc->n_light_bvh_nodes, sc->n_light_triangles, &lwide, &why_light);
            if (ok_scene) ref_stack_need = reference_stack_need(sc->bvh, sc->n_bvh_nodes);
        } catch (...) {
            ok_scene = false; why_scene = "out of host memory";
        }
        { std::lock_guard<std::mutex> l(mu); ready = true; }
        cv.notify_all();
    }
    void wait() { std::unique_lock<std::mutex> l(mu); cv.wait(l, [&] { return ready; }); }
};
int upload_scene_impl(ort_ctx* ctx, const ort_scene* sc, SharedWide* shared);

// ORT_TIMING=1: wall-clock phases of the host-facing calls on stderr (diagnostic)
struct PhaseTimer {
    bool on;
    const char* what;
    std::chrono::steady_clock::time_point t;
    std::string line;
    explicit PhaseTimer(const char* w) : on(std::getenv("ORT_TIMING") != nullptr), what(w), t(std::chrono::steady_clock::now()) {}
    void mark(const char* name) {
        if (!on) return;
        const auto n = std::chrono::steady_clock::now();
        char buf[96];
        std::snprintf(buf, sizeof buf, " %s=%.2fms", name, std::chrono::duration<double, std::milli>(n - t).count());
        line += buf;
        t = n;
    }
    ~PhaseTimer() { if (on) std::fprintf(stderr, "[%s]%s\n", what, line.c_str()); }
};

struct Bind {
    int prev = -1;
    explicit Bind(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~Bind() { if (prev >= 0) cudaSetDevice(prev); }
};

void free_scene(ort_ctx* c) { // only at destroy: uploads recycle the pools
    for (auto& t : c->tex_pool) {
        if (t.obj) cudaDestroyTextureObject(t.obj);
        if (t.arr) cudaFreeArray(t.arr);
    }
    c->tex_pool.clear();
    for (auto& b : c->sbuf) { if (b.p) cudaFree(b.p); b = ort_ctx::DevBuf{}; }
    c->has_scene = false;
    c->scene_bytes = 0;
}
void free_paths(ort_ctx* c) {
    for (auto& P : c->ps) {
        void* ptrs[] = {P.qo[0], P.qo[1], P.qd[0], P.qd[1], P.hits, P.lsum, P.lq, P.pa[0], P.pa[1], P.pb[0], P.pb[1], P.st_c};
        for (void* p : ptrs) if (p) cudaFree(p);
        P.qo[0] = P.qo[1] = P.qd[0] = P.qd[1] = P.hits = P.pa[0] = P.pa[1] = P.pb[0] = P.pb[1] = P.st_c = nullptr;
        P.lsum = nullptr; P.lq = nullptr;
        P.capacity = 0;
    }
    c->path_bytes = 0;
}

// fn(first, count) over [0, n) on up to ctx->host_threads threads (inline when the range is small)
template <typename F>
void parallel_for(ort_ctx* ctx, size_t n, size_t min_per_thread, F fn) {
    size_t nt = std::min<size_t>((size_t)std::max(ctx->host_threads, 1), n / std::max<size_t>(min_per_thread, 1));
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

// Pooled scene buffer: grows, never shrinks.
int scene_buffer(ort_ctx* ctx, int slot, size_t bytes, void** out) {
    auto& b = ctx->sbuf[slot];
    bytes = std::max<size_t>(bytes, 16);
    if (b.cap < bytes) {
        if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
        CK(cudaMalloc(&b.p, bytes));
        b.cap = bytes;
    }
    b.used = bytes;
    *out = b.p;
    return 0;
}

int ensure_stage(ort_ctx* ctx) {
    if (ctx->stage) return 0;
    CK(cudaMallocHost((void**)&ctx->stage, ort_ctx::STAGE_BYTES * ort_ctx::STAGE_SLOTS));
    for (auto& e : ctx->stage_ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return 0;
}

// Host -> device upload of n records of `item` bytes produced by fill(first, count, dst): the
// records are built (on several host threads) directly into a ring of pinned chunks, and the copy
// of one chunk runs while the next one is being built.
template <typename F>
int staged_upload(ort_ctx* ctx, void* d_dst, size_t n, size_t item, size_t min_per_thread, F fill) {
    if (n == 0) return 0;
    if (ensure_stage(ctx)) return 1;
    const size_t per_chunk = std::max<size_t>(1, ort_ctx::STAGE_BYTES / item);
    for (size_t off = 0; off < n; off += per_chunk) {
        const size_t cnt = std::min(per_chunk, n - off);
        const int slot = ctx->stage_next;
        ctx->stage_next = (slot + 1) % ort_ctx::STAGE_SLOTS;
        char* h = ctx->stage + (size_t)slot * ort_ctx::STAGE_BYTES;
        CK(cudaEventSynchronize(ctx->stage_ev[slot])); // the copy that last used this chunk is done
        parallel_for(ctx, cnt, min_per_thread, [&](size_t a, size_t c) { fill(off + a, c, h + a * item); });
        CK(cudaMemcpyAsync((char*)d_dst + off * item, h, cnt * item, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaEventRecord(ctx->stage_ev[slot], ctx->stream));
    }
    return 0;
}

// Same, for K record types derived from the SAME source items in one pass (the 168-byte triangles are
// read once, not once per record type): fill(first, count, outs[K]) writes record type j of item
// first + i to (char*)outs[j] + i * item[j].
template <int K, typename F>
int staged_upload_multi(ort_ctx* ctx, void* const (&d_dst)[K], const size_t (&item)[K], size_t n, size_t min_per_thread, F fill) {
    if (n == 0) return 0;
    if (ensure_stage(ctx)) return 1;
    size_t total = 0;
    for (int j = 0; j < K; j++) total += item[j];
    const size_t per_chunk = std::max<size_t>(1, ort_ctx::STAGE_BYTES / total);
    for (size_t off = 0; off < n; off += per_chunk) {
        const size_t cnt = std::min(per_chunk, n - off);
        const int slot = ctx->stage_next;
        ctx->stage_next = (slot + 1) % ort_ctx::STAGE_SLOTS;
        char* h = ctx->stage + (size_t)slot * ort_ctx::STAGE_BYTES;
        char* region[K];
        size_t acc = 0;
        for (int j = 0; j < K; j++) { region[j] = h + acc; acc += item[j] * per_chunk; }
        CK(cudaEventSynchronize(ctx->stage_ev[slot]));
        parallel_for(ctx, cnt, min_per_thread, [&](size_t a, size_t c) {
            void* outs[K];
            for (int j = 0; j < K; j++) outs[j] = region[j] + a * item[j];
            fill(off + a, c, outs);
        });
        for (int j = 0; j < K; j++)
            if (d_dst[j]) CK(cudaMemcpyAsync((char*)d_dst[j] + off * item[j], region[j], cnt * item[j], cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaEventRecord(ctx->stage_ev[slot], ctx->stream));
    }
    return 0;
}

// texture_index (textures.odin:79-104) applied to every texel on the host: u8 -> /255, missing
// channels = 1, optional pow(rgb, 2.2); the result is what a point fetch returns on the device.
int make_texture(ort_ctx* ctx, const ort_texture& t, bool srgb, cudaTextureObject_t* out) {
    if (t.data == nullptr || t.width <= 0 || t.height <= 0 || t.channels < 1 || t.channels > 4)
        return fail(ctx, "invalid texture (data/size/channels)");
    const size_t w = (size_t)t.width, h = (size_t)t.height;
    if (w * 16 > ort_ctx::STAGE_BYTES) return fail(ctx, "texture too wide");
    ort_ctx::TexSlot* slot = nullptr;
    for (auto& ts : ctx->tex_pool)
        if (!ts.in_use && ts.w == w && ts.h == h) { slot = &ts; break; }
    if (!slot) {
        ort_ctx::TexSlot ts;
        ts.w = w; ts.h = h;
        cudaChannelFormatDesc desc = cudaCreateChannelDesc<float4>();
        CK(cudaMallocArray(&ts.arr, &desc, w, h));
        cudaResourceDesc rd{};
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = ts.arr;
        cudaTextureDesc td{};
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        if (cudaError_t e = cudaCreateTextureObject(&ts.obj, &rd, &td, nullptr); e != cudaSuccess) {
            cudaFreeArray(ts.arr);
            return fail(ctx, std::string("cudaCreateTextureObject: ") + cudaGetErrorString(e));
        }
        ctx->tex_pool.push_back(ts);
        slot = &ctx->tex_pool.back();
    }
    slot->in_use = true;
    ctx->scene_bytes += (int64_t)(w * h * 16);
    float lut[256];
    for (int i = 0; i < 256; i++) {
        float x = (float)i / 255.0f;
        lut[i] = srgb ? std::pow(x, 2.2f) : x;
    }
    if (ensure_stage(ctx)) return 1;
    const size_t rows_per_chunk = std::max<size_t>(1, ort_ctx::STAGE_BYTES / (w * 16));
    for (size_t y0 = 0; y0 < h; y0 += rows_per_chunk) {
        const size_t rows = std::min(rows_per_chunk, h - y0);
        const int sl = ctx->stage_next;
        ctx->stage_next = (sl + 1) % ort_ctx::STAGE_SLOTS;
        float* texels = (float*)(ctx->stage + (size_t)sl * ort_ctx::STAGE_BYTES);
        CK(cudaEventSynchronize(ctx->stage_ev[sl]));
        parallel_for(ctx, rows, 16, [&](size_t ra, size_t rc) {
            for (size_t y = y0 + ra; y < y0 + ra + rc; y++)
                for (size_t x = 0; x < w; x++) {
                    float px[4] = {1, 1, 1, 1};
                    const size_t idx = y * (size_t)t.stride + x * (size_t)t.channels;
                    for (int c = 0; c < t.channels; c++) {
                        if (t.is_f32) {
                            float v = ((const float*)t.data)[idx + c];
                            px[c] = (srgb && c < 3) ? std::pow(v, 2.2f) : v;
                        } else {
                            uint8_t v = ((const uint8_t*)t.data)[idx + c];
                            px[c] = c < 3 ? lut[v] : (float)v / 255.0f;
                        }
                    }
                    if (srgb) // linalg.pow(pixel.rgb, 2.2) also hits the default 1.0 of absent channels: pow(1, 2.2) = 1
                        for (int c = t.channels; c < 3; c++) px[c] = 1.0f;
                    std::memcpy(&texels[((y - y0) * w + x) * 4], px, 16);
                }
        });
        CK(cudaMemcpy2DToArrayAsync(slot->arr, 0, y0, texels, w * 16, w * 16, rows, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaEventRecord(ctx->stage_ev[sl], ctx->stream));
    }
    *out = slot->obj;
    return 0;
}

int ensure_paths(ort_ctx* ctx, int64_t need, int pipes = 1) {
    if (ctx->max_path_bytes > 0 && (double)need * (16 * 10 + 12) * pipes > (double)ctx->max_path_bytes)
        return fail(ctx, "cudaMalloc: out of memory (ort_device_cfg.max_path_bytes)");
    for (int i = 0; i < pipes; i++) {
        auto& P = ctx->ps[i];
        if (P.capacity >= need) continue;
        void* ptrs[] = {P.qo[0], P.qo[1], P.qd[0], P.qd[1], P.hits, P.lsum, P.lq, P.pa[0], P.pa[1], P.pb[0], P.pb[1], P.st_c};
        for (void* p : ptrs) if (p) cudaFree(p);
        P.qo[0] = P.qo[1] = P.qd[0] = P.qd[1] = P.hits = P.pa[0] = P.pa[1] = P.pb[0] = P.pb[1] = P.st_c = nullptr;
        P.lsum = nullptr; P.lq = nullptr;
        ctx->path_bytes -= P.capacity * (16 * 10 + 12);
        P.capacity = 0;
        const size_t n = (size_t)need;
        for (int k = 0; k < 2; k++) {
            CK(cudaMalloc(&P.qo[k], n * 16));
            CK(cudaMalloc(&P.qd[k], n * 16));
            CK(cudaMalloc(&P.pa[k], n * 16));
            CK(cudaMalloc(&P.pb[k], n * 16));
        }
        CK(cudaMalloc(&P.hits, n * 16));
        CK(cudaMalloc(&P.lsum, n * 8));
        CK(cudaMalloc(&P.lq, n * 4));
        CK(cudaMalloc(&P.st_c, n * 16));
        P.capacity = need;
        ctx->path_bytes += (int64_t)(n * (16 * 10 + 12));
    }
    return 0;
}
int ensure_counters(ort_ctx* ctx, int depth, int pipes = 1) {
    for (int i = 0; i < pipes; i++) {
        auto& P = ctx->ps[i];
        if (P.counters && P.counters_depth >= depth) continue;
        if (P.counters) cudaFree(P.counters);
        P.counters = nullptr;
        CK(cudaMalloc(&P.counters, sizeof(uint32_t) * 5 * (size_t)(depth + 2)));
        P.counters_depth = depth;
    }
    return 0;
}
int ensure_scratch(ort_ctx* ctx, size_t bytes) {
    if (ctx->scratch_bytes >= bytes) return 0;
    if (ctx->scratch) cudaFree(ctx->scratch);
    ctx->scratch = nullptr; ctx->scratch_bytes = 0;
    CK(cudaMalloc(&ctx->scratch, bytes));
    ctx->scratch_bytes = bytes;
    return 0;
}
int ensure_pinned(ort_ctx* ctx, size_t bytes) {
    if (ctx->pinned_bytes >= bytes) return 0;
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr; ctx->pinned_bytes = 0;
    CK(cudaMallocHost(&ctx->pinned, bytes));
    ctx->pinned_bytes = bytes;
    return 0;
}

// mode 0: closest hit, 1: light-pdf sum
void launch_trace(ort_ctx* ctx, ort_ctx::PathSet& P, cudaStream_t st, const float4* qo, const float4* qd,
                  const uint32_t* n_ptr, uint32_t* work_ctr, int mode, float* lsum = nullptr,
                  const uint32_t* index = nullptr, int index_packed = 0) {
    TraceArgs a;
    a.qo = qo; a.qd = qd; a.n_ptr = n_ptr; a.work_ctr = work_ctr; a.index = index; a.index_packed = index_packed;
    a.hits = P.hits; a.lsum = lsum ? lsum : P.lsum;
    a.refill_threshold = mode == 0 ? ctx->refill : ctx->refill_light; a.inner_min = ctx->inner_min;
    if (mode == 0) k_trace<false><<<ctx->trace_grid[0], TRACE_THREADS, 0, st>>>(ctx->sd, a);
    else k_trace<true><<<ctx->trace_grid[1], TRACE_THREADS, 0, st>>>(ctx->sd, a);
    ctx->launches++;
}

struct Prof { // profiling runs single-pipeline on ctx->stream
    ort_ctx* c;
    double* acc;
    Prof(ort_ctx* ctx, double* a) : c(ctx), acc(a) { if (c->profiling) cudaEventRecord(c->evp0, c->stream); }
    ~Prof() {
        if (c->profiling) {
            cudaEventRecord(c->evp1, c->stream);
            cudaEventSynchronize(c->evp1);
            float ms = 0;
            cudaEventElapsedTime(&ms, c->evp0, c->evp1);
            *acc += ms;
        }
    }
};

void fill_params(ort_ctx* ctx, uint32_t w, uint32_t h, int32_t depth, RenderParams* p) {
    float M[16];
    make_pixel_to_ray_dir(ctx->cam, w, h, M);
    std::memcpy(p->M, M, sizeof(float) * 12);
    std::memcpy(p->cam_pos, ctx->cam.pos, 12);
    p->w = w; p->h = h; p->npix = w * h;
    p->ray_depth = depth;
    p->seed = ctx->seed;
    p->n_batch_samples = 1;
    p->sample_base = 0;
    p->tiled = 0;
    p->tile_w = 2; p->tile_h = 2; p->tile_s = 8;
}

// One wave: n_batch_samples samples of every pixel, all bounces, then accumulation.  `wait_for` is
// the previous wave's `resolved` event (other pipeline): accumulation stays in wave order.
int launch_wave(ort_ctx* ctx, ort_ctx::PathSet& P, cudaStream_t st, cudaEvent_t wait_for, const RenderParams& p,
                float* d_accum, float* d_first, float* d_last, int write_first, int write_last) {
    const int D = p.ray_depth;
    uint32_t* counts = P.counters;
    uint32_t* wtrace = P.counters + (D + 2);
    uint32_t* wlight = P.counters + 2 * (D + 2);
    uint32_t* used = P.counters + 3 * (D + 2);
    uint32_t* lcount = P.counters + 4 * (D + 2);
    const bool lights = ctx->sd.n_lights > 0;
    // light-queue entries carry the root-children mask in their top bits when the wave's positions fit below them
    const int packed = (uint64_t)p.n_batch_samples * p.npix <= ((uint64_t)1 << LQ_MASK_SHIFT) ? 1 : 0;
    const int prefilter = ctx->light_prefilter >= 2 ? 1 + packed : (ctx->light_prefilter ? 1 : 0);
    {
        Prof pr(ctx, &ctx->ms_other);
        CK(cudaMemsetAsync(P.counters, 0, sizeof(uint32_t) * 5 * (size_t)(D + 2), st));
        // L = 0: k_shade only touches the accumulator of a path when a hit adds radiance
        CK(cudaMemsetAsync(P.st_c, 0, (size_t)p.n_batch_samples * p.npix * 16, st));
        k_raygen<<<ctx->shade_grid, 256, 0, st>>>(p, P.qo[0], P.qd[0], counts);
        ctx->launches++;
    }
    for (int k = 0; k < D; k++) {
        const int in = k & 1, out = in ^ 1;
        const bool need_light = k > 0 && lights; // bounce 0 has no pending pdf to complete
        float* lsum_in = P.lsum + (size_t)in * (size_t)P.capacity;
        float* lsum_out = P.lsum + (size_t)out * (size_t)P.capacity;
        {
            Prof pr(ctx, &ctx->ms_trace);
            launch_trace(ctx, P, st, P.qo[in], P.qd[in], counts + k, wtrace + k, 0, lsum_in);
        }
        if (need_light) {
            // only the rays k_shade queued as light candidates (the others already have lsum = 0)
            Prof pr(ctx, &ctx->ms_light);
            launch_trace(ctx, P, st, P.qo[in], P.qd[in], lcount + k, wlight + k, 1, lsum_in, P.lq, prefilter == 2);
        }
        {
            Prof pr(ctx, &ctx->ms_shade);
            ShadeArgs sa;
            sa.qo_in = P.qo[in]; sa.qd_in = P.qd[in]; sa.hits = P.hits; sa.pa_in = P.pa[in]; sa.pb_in = P.pb[in];
            sa.lsum = lsum_in; sa.n_in_ptr = counts + k;
            sa.qo_out = P.qo[out]; sa.qd_out = P.qd[out]; sa.pa_out = P.pa[out]; sa.pb_out = P.pb[out];
            sa.n_out_ptr = counts + k + 1; sa.used_ptr = used + k; sa.st_c = P.st_c;
            sa.lsum_out = lsum_out; sa.lq = P.lq; sa.lq_count = lcount + k + 1;
            sa.bounce = k; sa.prefilter = prefilter; sa.bin_octants = ctx->bin_octants;
            k_shade<<<ctx->shade_grid, 256, 0, st>>>(ctx->sd, p, sa);
            ctx->launches++;
        }
    }
    {
        Prof pr(ctx, &ctx->ms_other);
        if (wait_for) CK(cudaStreamWaitEvent(st, wait_for, 0));
        k_resolve<<<ctx->shade_grid, 256, 0, st>>>(p, P.st_c, d_accum, d_first, d_last, write_first, write_last);
        k_stats<<<1, 32, 0, st>>>(counts, used, lcount, D, lights ? 1 : 0, ctx->d_stats);
        ctx->launches += 2;
        CK(cudaEventRecord(P.resolved, st));
    }
    CK(cudaGetLastError());
    return 0;
}

// Make ctx->stream wait for everything enqueued on the other pipelines' streams.
int join_pipes(ort_ctx* ctx) {
    if (!ctx->pipes_busy) return 0;
    for (int i = 1; i < MAX_PIPES; i++) {
        CK(cudaEventRecord(ctx->ev_join[i], ctx->aux_stream[i]));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join[i], 0));
    }
    ctx->pipes_busy = false;
    ctx->last_resolved = nullptr; // later waves are ordered behind ctx->stream by the next fork
    ctx->need_fork = true;
    return 0;
}

// chain = false (ort_render, ort_render_device): fork the pipelines off ctx->stream, enqueue the waves, join
// them back — everything of this call is complete once ctx->stream reaches the end of the call.
// chain = true (ort_frame_render): the pipelines keep running ACROSS calls — wave k of the frame runs on
// pipeline k % pipes whichever call enqueued it, its accumulation waits for wave k-1's, and nothing joins
// at the end of the call, so consecutive calls leave no drain / ramp-up gap on the device.
int render_impl(ort_ctx* ctx, uint32_t w, uint32_t h, int32_t depth, uint64_t first_sample, uint64_t n_samples,
                float* d_accum, float* d_first, float* d_last, const volatile uint8_t* interrupt, uint64_t* done_out,
                bool chain = false) {
    if (!ctx->has_scene) return fail(ctx, "ort_upload_scene has not been called");
    if (w == 0 || h == 0) return fail(ctx, "width and height must be non-zero");
    if (depth < 0) return fail(ctx, "ray_depth must be >= 0");
    const uint64_t npix = (uint64_t)w * h;
    if (npix > (1ull << 31)) return fail(ctx, "image too large");
    int64_t cap = ctx->capacity_cfg > 0 ? ctx->capacity_cfg : ((int64_t)1 << 25); // paths per wave (x up to 4 overlapped waves)
    if ((uint64_t)cap < npix) cap = (int64_t)npix;
    uint64_t per_wave = std::max<uint64_t>(1, (uint64_t)cap / npix);
    if (!chain && per_wave > n_samples) per_wave = std::max<uint64_t>(n_samples, 1);
    // several wave pipelines on separate streams: the tail of every persistent kernel of one wave (a
    // few warps finishing their last rays) is filled by the other waves' kernels
    int pipes = 1, overlap = ctx->overlap; // a fallback below de-tunes THIS call only
    ctx->last_done = 0;
    if (done_out) *done_out = 0;
    if (!chain && join_pipes(ctx)) return 1; // waves of an earlier ort_frame_render may still be in flight
    for (;;) {
        const uint64_t n_waves = (n_samples + per_wave - 1) / per_wave;
        pipes = std::max(1, std::min(overlap, MAX_PIPES));
        if (ctx->profiling || depth == 0) pipes = 1;
        if (!chain && (uint64_t)pipes > n_waves) pipes = (int)n_waves;
        if (pipes < 1) pipes = 1;
        if (ctx->pipes_busy && (pipes != ctx->chain_pipes || (int64_t)(per_wave * npix) > ctx->ps[0].capacity) && join_pipes(ctx)) return 1;
        if (ensure_paths(ctx, (int64_t)(per_wave * npix), pipes) == 0) break;
        // not enough free HBM for this many paths in flight (other contexts / processes on the GPU):
        // fall back to smaller waves, then to fewer pipelines, before giving up
        cudaGetLastError();
        if (join_pipes(ctx)) return 1;
        CK(cudaStreamSynchronize(ctx->stream));
        free_paths(ctx);
        if (per_wave > 1) per_wave = (per_wave + 1) / 2;
        else if (overlap > 1) overlap = 1;
        else return 1; // ctx->err holds the cudaMalloc message
    }
    if (ensure_counters(ctx, depth, pipes)) return 1;
    RenderParams p;
    fill_params(ctx, w, h, depth, &p);
    p.tiled = ctx->tiled;
    p.tile_w = ctx->tile_w; p.tile_h = ctx->tile_h; p.tile_s = ctx->tile_s;
    if (!chain) CK(cudaEventRecord(ctx->ev0, ctx->stream));
    if (pipes > 1 && (!chain || ctx->need_fork)) {
        // the other pipelines start behind whatever ctx->stream already holds (scene upload, cleared accumulators)
        CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
        for (int i = 1; i < pipes; i++) CK(cudaStreamWaitEvent(ctx->aux_stream[i], ctx->ev_fork, 0));
    }
    ctx->need_fork = false;
    if (!chain) { ctx->wave_seq = 0; ctx->last_resolved = nullptr; }
    uint64_t done = 0;
    while (done < n_samples) {
        if (interrupt && *interrupt) break; // is_interrupted(), raytracer.odin:554
        const uint64_t nb = std::min<uint64_t>(per_wave, n_samples - done);
        p.sample_base = first_sample + done;
        p.n_batch_samples = (uint32_t)nb;
        const int pi = (int)(ctx->wave_seq % (uint64_t)pipes);
        cudaStream_t st = pi ? ctx->aux_stream[pi] : ctx->stream;
        // with an interrupt flag the host stays at most `pipes` waves ahead of the device (it waits for
        // the wave that last used this pipeline), so the poll above sees a SIGINT within a few waves
        // while the pipelines still overlap
        if (interrupt && ctx->ps[pi].used) CK(cudaEventSynchronize(ctx->ps[pi].resolved));
        if (depth == 0) {
            // raytrace(depth_left = 0) returns 0 (raytracer.odin:433): count the samples, add nothing
            CK(cudaMemsetAsync(ctx->ps[0].st_c, 0, (size_t)(nb * npix) * 16, ctx->stream));
            k_resolve<<<ctx->shade_grid, 256, 0, ctx->stream>>>(p, ctx->ps[0].st_c, d_accum, d_first, d_last,
                                                                d_first && done == 0, d_last && done + nb == n_samples);
            ctx->launches++;
        } else {
            if (launch_wave(ctx, ctx->ps[pi], st, pipes > 1 ? ctx->last_resolved : nullptr, p, d_accum, d_first, d_last,
                            d_first && done == 0, d_last != nullptr))
                return 1;
            ctx->last_resolved = ctx->ps[pi].resolved;
            ctx->ps[pi].used = true;
        }
        done += nb;
        ctx->wave_seq++;
    }
    if (chain) {
        if (pipes > 1 && done > 0) { ctx->pipes_busy = true; ctx->chain_pipes = pipes; }
    } else {
        for (int i = 1; i < pipes; i++) {
            CK(cudaEventRecord(ctx->ev_join[i], ctx->aux_stream[i]));
            CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join[i], 0));
        }
        ctx->last_resolved = nullptr;
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
    }
    ctx->last_done = done;
    if (done_out) *done_out = done;
    return 0;
}


// Device accumulators (+ first / last planes) -> packed Sample_Stats -> pinned host -> merged into
// `out` like repeated rc_set_pixel calls would (main.odin:96-101).
int pack_and_merge(ort_ctx* ctx, const float* accum, const float* first, const float* last, size_t npix,
                   uint32_t* packed, ort_sample_stats* out, PhaseTimer* pt = nullptr) {
    k_pack_stats<<<ctx->shade_grid, 256, 0, ctx->stream>>>(accum, first, last, (uint32_t)npix, packed);
    ctx->launches++;
    if (ensure_pinned(ctx, npix * 52)) return 1;
    CK(cudaMemcpyAsync(ctx->pinned, packed, npix * 52, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (pt) pt->mark("render+d2h");
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->ms_render = ms; else cudaGetLastError();
    const ort_sample_stats* src = (const ort_sample_stats*)ctx->pinned;
    parallel_for(ctx, npix, 1 << 16, [&](size_t f, size_t n) {
        for (size_t i = f; i < f + n; i++) {
            if (src[i].count == 0) continue;
            ort_sample_stats& d = out[i];
            if (d.count == 0) std::memcpy(d.first, src[i].first, 12);
            d.count += src[i].count;
            std::memcpy(d.last, src[i].last, 12);
            for (int c = 0; c < 3; c++) { d.total[c] += src[i].total[c]; d.total_squared[c] += src[i].total_squared[c]; }
        }
    });
    if (pt) pt->mark("merge");
    return 0;
}

// Second buffer of the frame: a device-to-device copy of the accumulators as of everything enqueued so far.
// Preview / fetch work on it from the side stream while the pipelines go on accumulating.
int frame_snapshot(ort_ctx* ctx) {
    const size_t npix = (size_t)ctx->frame_w * ctx->frame_h;
    if (join_pipes(ctx)) return 1;
    if (!ctx->side_stream) {
        CK(cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ctx->snap_ev, cudaEventDisableTiming));
    }
    // (the previous snapshot's readers have finished: preview / fetch synchronise the side stream before returning)
    CK(cudaMemcpyAsync(ctx->frame + 14 * npix, ctx->frame, npix * 14 * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaEventRecord(ctx->snap_ev, ctx->stream));
    ctx->frame_snapshot_valid = true;
    ctx->need_fork = true;
    return 0;
}
int ensure_side(ort_ctx* ctx, size_t bytes) {
    if (ctx->side_bytes >= bytes) return 0;
    if (ctx->side_buf) { CK(cudaStreamSynchronize(ctx->side_stream)); cudaFree(ctx->side_buf); }
    ctx->side_buf = nullptr; ctx->side_bytes = 0;
    CK(cudaMalloc(&ctx->side_buf, bytes));
    ctx->side_bytes = bytes;
    return 0;
}

} // namespace

// =================================================================================================
extern "C" {

int ort_abi_version(void) { return ORT_ABI_VERSION; }

const char* ort_last_error(const ort_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int ort_create(ort_ctx** out, const ort_device_cfg* cfg) {
    ort_ctx* ctx = nullptr; // CK reports into g_create_error while ctx is null
    if (!out) return fail(nullptr, "ort_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)");
    const int dev = cfg ? cfg->device : 0;
    if (dev < 0 || dev >= n) return fail(nullptr, "ort_create: device ordinal out of range");
    cudaDeviceProp prop{};
    CK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) return fail(nullptr, "ort_create: this library is built for sm_100a (B200) only");
    Bind b(dev);
    ort_ctx* c = new ort_ctx();
    c->device = dev;
    c->sm_count = prop.multiProcessorCount;
    c->seed = cfg ? cfg->seed : 0;
    c->capacity_cfg = cfg ? cfg->max_paths_in_flight : 0;
    c->host_threads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    if (const char* e2 = std::getenv("ORT_HOST_THREADS")) c->host_threads = std::max(1, std::atoi(e2));
    c->max_path_bytes = cfg ? cfg->max_path_bytes : 0;
    if (const char* e2 = std::getenv("ORT_OVERLAP")) c->overlap = std::atoi(e2);
    if (const char* e2 = std::getenv("ORT_WAVE_PATHS")) c->capacity_cfg = std::atoll(e2);
#ifdef ORT_TUNING
#include "tuning_env.inl" // make variant EXTRA=-DORT_TUNING: knobs from the environment for tools/tune.py
#endif
    ctx = c;
    auto bail = [&](const char* what, cudaError_t err) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
        delete c;
        return 1;
    };
    if ((e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    c->stream = c->own_stream;
    if ((e = cudaEventCreate(&c->ev0)) != cudaSuccess) return bail("cudaEventCreate", e);
    cudaEventCreate(&c->ev1); cudaEventCreate(&c->evp0); cudaEventCreate(&c->evp1);
    cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    for (int i = 1; i < MAX_PIPES; i++) {
        cudaStreamCreateWithFlags(&c->aux_stream[i], cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming);
    }
    for (auto& P : c->ps) cudaEventCreateWithFlags(&P.resolved, cudaEventDisableTiming);
    if ((e = cudaMalloc(&c->d_stats, 8 * sizeof(unsigned long long))) != cudaSuccess) return bail("cudaMalloc", e);
    cudaMemset(c->d_stats, 0, 8 * sizeof(unsigned long long));
    // persistent grids: as many CTAs as stay resident, a multiple of the SM count
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trace<false>, TRACE_THREADS, 0);
    c->trace_grid[0] = c->sm_count * std::max(occ, 1);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trace<true>, TRACE_THREADS, 0);
    c->trace_grid[1] = c->sm_count * std::max(occ, 1);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_shade, 256, 0);
    c->shade_grid = c->sm_count * std::max(occ, 1);
    *out = c;
    return 0;
}

void ort_destroy(ort_ctx* ctx) {
    if (!ctx) return;
    Bind b(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_scene(ctx);
    free_paths(ctx);
    for (auto& P : ctx->ps) { if (P.counters) cudaFree(P.counters); if (P.resolved) cudaEventDestroy(P.resolved); }
    cudaEventDestroy(ctx->ev_fork);
    for (int i = 1; i < MAX_PIPES; i++) {
        if (ctx->aux_stream[i]) { cudaStreamSynchronize(ctx->aux_stream[i]); cudaStreamDestroy(ctx->aux_stream[i]); }
        if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]);
    }
    if (ctx->d_stats) cudaFree(ctx->d_stats);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->frame) cudaFree(ctx->frame);
    if (ctx->side_stream) { cudaStreamSynchronize(ctx->side_stream); cudaStreamDestroy(ctx->side_stream); }
    if (ctx->side_buf) cudaFree(ctx->side_buf);
    if (ctx->snap_ev) cudaEventDestroy(ctx->snap_ev);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->stage) cudaFreeHost(ctx->stage);
    for (auto& e : ctx->stage_ev) if (e) cudaEventDestroy(e);
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1); cudaEventDestroy(ctx->evp0); cudaEventDestroy(ctx->evp1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int ort_set_stream(ort_ctx* ctx, void* cuda_stream) {
    if (!ctx) return 1;
    Bind b(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream == ORT_OWN_STREAM ? ctx->own_stream : (cudaStream_t)cuda_stream;
    return 0;
}

int ort_set_profiling(ort_ctx* ctx, int32_t on) {
    if (!ctx) return 1;
    ctx->profiling = on != 0;
    return 0;
}

int ort_upload_scene(ort_ctx* ctx, const ort_scene* sc) {
    if (!ctx) return 1;
    return guarded(ctx, [&] { return upload_scene_impl(ctx, sc, nullptr); });
}

} // extern "C"

namespace {
int upload_scene_impl(ort_ctx* ctx, const ort_scene* sc, SharedWide* shared) {
    if (!sc) return fail(ctx, "ort_upload_scene: scene is NULL");
    Bind b(ctx->device);
    PhaseTimer pt("ort_upload_scene");
    if (join_pipes(ctx)) return 1; // chained waves of a frame may still be reading the scene on the other pipelines
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->has_scene = false;
    ctx->scene_bytes = 0;
    for (auto& ts : ctx->tex_pool) ts.in_use = false;
    for (auto& sb : ctx->sbuf) sb.used = 0;
    if (sc->n_triangles < 0 || sc->n_light_triangles < 0 || sc->n_materials < 0 || sc->n_textures < 0)
        return fail(ctx, "ort_upload_scene: negative count");
    if (sc->n_triangles >= (1 << 28)) return fail(ctx, "ort_upload_scene: more than 2^28 triangles");
    if ((uint64_t)sc->n_triangles + (uint64_t)sc->n_light_triangles >= (1u << 28))
        return fail(ctx, "ort_upload_scene: more than 2^28 traversal triangles");
    if ((sc->n_triangles > 0 && !sc->triangles) || (sc->n_light_triangles > 0 && !sc->light_triangles) ||
        (sc->n_materials > 0 && !sc->materials) || (sc->n_textures > 0 && !sc->textures) ||
        (sc->n_bvh_nodes > 0 && !sc->bvh) || (sc->n_light_bvh_nodes > 0 && !sc->light_bvh))
        return fail(ctx, "ort_upload_scene: NULL array with a non-zero count");
    if (sc->n_bvh_nodes < 0 || sc->n_light_bvh_nodes < 0) return fail(ctx, "ort_upload_scene: negative count");
    const size_t nt = (size_t)sc->n_triangles, nlt = (size_t)sc->n_light_triangles;

    // The wide-BVH re-emission runs on its own host thread while this thread builds and uploads the
    // per-triangle records, which do not depend on it.  With several GPUs (ort_multi_upload_scene) the
    // emission is done ONCE and shared: `shared` then carries the result and this call only waits for it.
    SharedWide own;
    SharedWide* sw = shared ? shared : &own;
    std::thread wide_thread;
    if (!shared) wide_thread = std::thread([&] { own.build(sc); });
    struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{wide_thread};

    ctx->cam = sc->cam;
    ctx->n_tris = sc->n_triangles;
    ctx->n_ltris = sc->n_light_triangles;
    SceneDev sd{};
    void* d = nullptr;

    // materials first: they decide which optional per-triangle records are needed
    bool any_tex = false, any_normal = false;
    std::vector<DevMaterial> mats((size_t)sc->n_materials);
    std::vector<char> used_raw((size_t)sc->n_textures, 0), used_lin((size_t)sc->n_textures, 0);
    for (int64_t i = 0; i < sc->n_materials; i++) {
        const ort_material& m = sc->materials[i];
        DevMaterial dm{};
        std::memcpy(dm.color, m.color_factor, 12);
        std::memcpy(dm.emission, m.emission_factor, 12);
        dm.roughness = m.roughness_factor;
        dm.metallic = m.metallic_factor;
        const int32_t ids[4] = {m.color_texture, m.emission_texture, m.metallic_roughness_texture, m.normal_texture};
        for (int k = 0; k < 4; k++) {
            if (ids[k] >= sc->n_textures) return fail(ctx, "material texture index out of range");
            if (ids[k] >= 0) { any_tex = true; (k < 2 ? used_lin : used_raw)[ids[k]] = 1; }
        }
        if (m.normal_texture >= 0) any_normal = true;
        dm.color_tex = ids[0] < 0 ? -1 : ids[0]; dm.emission_tex = ids[1] < 0 ? -1 : ids[1];
        dm.mr_tex = ids[2] < 0 ? -1 : ids[2]; dm.normal_tex = ids[3] < 0 ? -1 : ids[3];
        mats[(size_t)i] = dm;
    }
    if (scene_buffer(ctx, ort_ctx::SB_MATS, mats.size() * sizeof(DevMaterial), &d)) return 1;
    sd.mats = (const DevMaterial*)d;
    if (staged_upload(ctx, d, mats.size(), sizeof(DevMaterial), 1 << 20,
                      [&](size_t f, size_t c, void* o) { std::memcpy(o, mats.data() + f, c * sizeof(DevMaterial)); }))
        return 1;

    // per-triangle records of the scene triangles, all in ONE pass over the 168-byte source structs:
    // traversal (TriIsect), shading (TriShade, + the material_index range check), and — only when a
    // material needs them — texture coordinates (TriUV) and tangents (TriTan).  The traversal array
    // continues with the light triangles (a ray walks both trees with the same code).
    std::atomic<int> bad_material{0};
    {
        void* d_isect = nullptr; void* d_shade = nullptr; void* d_uv = nullptr; void* d_tan = nullptr;
        if (scene_buffer(ctx, ort_ctx::SB_TRIS, (nt + nlt) * sizeof(TriIsect), &d_isect)) return 1;
        if (scene_buffer(ctx, ort_ctx::SB_TSHADE, nt * sizeof(TriShade), &d_shade)) return 1;
        if (any_tex && scene_buffer(ctx, ort_ctx::SB_TUV, nt * sizeof(TriUV), &d_uv)) return 1;
        if (any_normal && scene_buffer(ctx, ort_ctx::SB_TTAN, nt * sizeof(TriTan), &d_tan)) return 1;
        sd.tris = (const float4*)d_isect;
        sd.ltris = sd.tris + nt * 4;
        sd.light_tri_base = (uint32_t)nt;
        sd.tshade = (const float4*)d_shade;
        sd.tuv = (const float4*)d_uv;
        sd.ttan = (const float4*)d_tan;
        void* const dsts[4] = {d_isect, d_shade, d_uv, d_tan};
        const size_t items[4] = {sizeof(TriIsect), sizeof(TriShade), any_tex ? sizeof(TriUV) : 0, any_normal ? sizeof(TriTan) : 0};
        const bool want_uv = any_tex, want_tan = any_normal;
        if (staged_upload_multi<4>(ctx, dsts, items, nt, 4096, [&](size_t f, size_t c, void** outs) {
                make_isect_records(sc->triangles + f, (int64_t)c, (TriIsect*)outs[0]);
                TriShade* rs = (TriShade*)outs[1];
                TriUV* ru = (TriUV*)outs[2];
                TriTan* rt = (TriTan*)outs[3];
                for (size_t i = 0; i < c; i++) {
                    const ort_triangle& t = sc->triangles[f + i];
                    TriShade& r = rs[i];
                    std::memcpy(r.n1, t.n1, 12); std::memcpy(r.n2, t.n2, 12); std::memcpy(r.n3, t.n3, 12);
                    r.ngx = t.ng[0]; r.ngy = t.ng[1]; r.ngz = t.ng[2];
                    if (t.material_index < 0 || t.material_index >= sc->n_materials) bad_material.store(1, std::memory_order_relaxed);
                    r.material = (int32_t)t.material_index; r.flags = 0; r.pad0 = r.pad1 = 0;
                    if (want_uv) {
                        TriUV& u = ru[i];
                        std::memcpy(u.tex1, t.tex1, 8); std::memcpy(u.tex2, t.tex2, 8); std::memcpy(u.tex3, t.tex3, 8);
                        u.pad[0] = u.pad[1] = 0;
                    }
                    if (want_tan) {
                        std::memcpy(rt[i].tan1, t.tan1, 16); std::memcpy(rt[i].tan2, t.tan2, 16); std::memcpy(rt[i].tan3, t.tan3, 16);
                    }
                }
            }))
            return 1;
        if (bad_material.load()) { cudaStreamSynchronize(ctx->stream); return fail(ctx, "triangle material_index out of range"); }
        if (staged_upload(ctx, (char*)d_isect + nt * sizeof(TriIsect), nlt, sizeof(TriIsect), 4096,
                          [&](size_t f, size_t c, void* o) { make_isect_records(sc->light_triangles + f, (int64_t)c, (TriIsect*)o, true); }))
            return 1;
    }
    pt.mark("triangle_records");
    {
        std::vector<DevTexture> texs((size_t)sc->n_textures);
        for (int64_t i = 0; i < sc->n_textures; i++) {
            DevTexture dt{};
            dt.w = sc->textures[i].width; dt.h = sc->textures[i].height;
            if (used_raw[(size_t)i] && make_texture(ctx, sc->textures[i], false, &dt.raw)) return 1;
            if (used_lin[(size_t)i] && make_texture(ctx, sc->textures[i], true, &dt.linear)) return 1;
            texs[(size_t)i] = dt;
        }
        if (scene_buffer(ctx, ort_ctx::SB_TEXS, texs.size() * sizeof(DevTexture), &d)) return 1;
        sd.texs = (const DevTexture*)d;
        if (staged_upload(ctx, d, texs.size(), sizeof(DevTexture), 1 << 20,
                          [&](size_t f, size_t c, void* o) { std::memcpy(o, texs.data() + f, c * sizeof(DevTexture)); }))
            return 1;
    }
    if (sc->env_map) {
        sd.env.w = sc->env_map->width; sd.env.h = sc->env_map->height;
        if (make_texture(ctx, *sc->env_map, false, &sd.env.raw)) return 1;
        sd.has_env = 1;
    }
    pt.mark("textures");

    // nodes: scene BVH (root 0) followed by the light BVH, whose references are rebased
    if (shared) shared->wait(); else wide_thread.join();
    pt.mark("wait_wide_bvh");
    if (!sw->ok_scene) { cudaStreamSynchronize(ctx->stream); return fail(ctx, std::string("scene BVH: ") + sw->why_scene); }
    if (!sw->ok_light) { cudaStreamSynchronize(ctx->stream); return fail(ctx, std::string("light BVH: ") + sw->why_light); }
    const WideBVH& wide = sw->wide;
    const WideBVH& lwide = sw->lwide;
    if (wide.max_stack > MAX_STACK || lwide.max_stack > MAX_STACK) {
        cudaStreamSynchronize(ctx->stream);
        return fail(ctx, "BVH too deep for the traversal stack (worst case " + std::to_string(wide.max_stack) + " > " +
                             std::to_string(MAX_STACK) + ")");
    }
    {
        const size_t ns = wide.nodes.size(), nl = lwide.nodes.size();
        auto light_node = [&](size_t i) {
            WideNode w = lwide.nodes[i];
            for (int k = 0; k < 4; k++) {
                if (w.child[k] == WIDE_EMPTY) continue;
                if (w.child[k] >= 0) w.child[k] += (int32_t)ns;
                else {
                    const uint32_t code = (uint32_t)~w.child[k];
                    w.child[k] = ~(int32_t)((((code >> 3) + (uint32_t)sc->n_triangles) << 3) | (code & 7u));
                }
            }
            return w;
        };
        if (scene_buffer(ctx, ort_ctx::SB_NODES, (ns + nl) * sizeof(WideNode), &d)) return 1;
        if (staged_upload(ctx, d, ns, sizeof(WideNode), 8192,
                          [&](size_t f, size_t c, void* o) { std::memcpy(o, wide.nodes.data() + f, c * sizeof(WideNode)); }))
            return 1;
        if (staged_upload(ctx, (char*)d + ns * sizeof(WideNode), nl, sizeof(WideNode), 8192, [&](size_t f, size_t c, void* o) {
                for (size_t i = 0; i < c; i++) ((WideNode*)o)[i] = light_node(f + i);
            }))
            return 1;
        sd.nodes = (const float4*)d;
        sd.light_root = (int32_t)ns;
    }
    sd.n_lights = (int32_t)sc->n_light_triangles;
    std::memcpy(sd.pad_scale, wide.max_abs, 12);
    ctx->wide_nodes = (int64_t)wide.nodes.size(); ctx->wide_depth = wide.depth; ctx->wide_max_stack = wide.max_stack;
    ctx->lwide_nodes = (int64_t)lwide.nodes.size();
    ctx->ref_stack_need = sw->ref_stack_need;
    CK(cudaStreamSynchronize(ctx->stream));
    pt.mark("nodes+drain");
    // recycle what the new scene does not use (only happens when the scene's shape changed)
    for (size_t i = 0; i < ctx->tex_pool.size();) {
        if (ctx->tex_pool[i].in_use) { i++; continue; }
        cudaDestroyTextureObject(ctx->tex_pool[i].obj);
        cudaFreeArray(ctx->tex_pool[i].arr);
        ctx->tex_pool.erase(ctx->tex_pool.begin() + (long)i);
    }
    for (auto& sb : ctx->sbuf) ctx->scene_bytes += (int64_t)sb.used;
    ctx->sd = sd;
    ctx->has_scene = true;
    return 0;
}
} // namespace

extern "C" {

int ort_render_device(ort_ctx* ctx, uint32_t w, uint32_t h, int32_t ray_depth, uint64_t first_sample,
                      uint64_t n_samples, float* d_accum) {
    if (!ctx) return 1;
    if (!d_accum) return fail(ctx, "ort_render_device: d_accum is NULL");
    Bind b(ctx->device);
    return guarded(ctx, [&] { return render_impl(ctx, w, h, ray_depth, first_sample, n_samples, d_accum, nullptr, nullptr, nullptr, nullptr); });
}

int ort_render(ort_ctx* ctx, uint32_t w, uint32_t h, int32_t ray_depth, uint64_t first_sample, uint64_t n_samples,
               ort_sample_stats* out, const volatile uint8_t* interrupt) {
    if (!ctx) return 1;
    if (!out) return fail(ctx, "ort_render: out is NULL");
    Bind b(ctx->device);
    return guarded(ctx, [&]() -> int {
    const size_t npix = (size_t)w * h;
    // planes: 8 accum + 3 first + 3 last, then the packed Sample_Stats
    if (ensure_scratch(ctx, npix * (14 * 4 + 52))) return 1;
    float* accum = ctx->scratch;
    float* first = accum + 8 * npix;
    float* last = first + 3 * npix;
    uint32_t* packed = (uint32_t*)(last + 3 * npix);
    PhaseTimer pt("ort_render");
    CK(cudaMemsetAsync(accum, 0, npix * 14 * 4, ctx->stream));
    uint64_t done = 0;
    if (render_impl(ctx, w, h, ray_depth, first_sample, n_samples, accum, first, last, interrupt, &done)) return 1;
    pt.mark("enqueue");
    const int rc = pack_and_merge(ctx, accum, first, last, npix, packed, out, &pt);
    return rc;
    });
}

uint64_t ort_last_render_samples(const ort_ctx* ctx) { return ctx ? ctx->last_done : 0; }

// ---- device-resident frame ---------------------------------------------------------------------------
int ort_frame_begin(ort_ctx* ctx, uint32_t w, uint32_t h) {
    if (!ctx) return 1;
    if (w == 0 || h == 0) return fail(ctx, "ort_frame_begin: width and height must be non-zero");
    Bind b(ctx->device);
    return guarded(ctx, [&]() -> int {
        if (join_pipes(ctx)) return 1;
        const size_t npix = (size_t)w * h, bytes = npix * 14 * 4;
        if (ctx->frame_bytes < 2 * bytes) { // accumulators + their snapshot
            CK(cudaStreamSynchronize(ctx->stream));
            if (ctx->frame) cudaFree(ctx->frame);
            ctx->frame = nullptr; ctx->frame_bytes = 0;
            CK(cudaMalloc(&ctx->frame, 2 * bytes));
            ctx->frame_bytes = 2 * bytes;
        }
        CK(cudaMemsetAsync(ctx->frame, 0, bytes, ctx->stream));
        ctx->frame_w = w; ctx->frame_h = h;
        ctx->frame_has_first = false;
        ctx->frame_snapshot_valid = false;
        ctx->need_fork = true;
        return 0;
    });
}

int ort_frame_end(ort_ctx* ctx) {
    if (!ctx) return 1;
    Bind b(ctx->device);
    return guarded(ctx, [&]() -> int {
        if (join_pipes(ctx)) return 1;
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->side_stream) CK(cudaStreamSynchronize(ctx->side_stream));
        if (ctx->frame) cudaFree(ctx->frame);
        ctx->frame = nullptr; ctx->frame_bytes = 0; ctx->frame_w = ctx->frame_h = 0;
        return 0;
    });
}

int ort_frame_load(ort_ctx* ctx, const ort_sample_stats* in) {
    if (!ctx) return 1;
    if (!ctx->frame_w) return fail(ctx, "ort_frame_load: no frame (ort_frame_begin)");
    if (!in) return fail(ctx, "ort_frame_load: in is NULL");
    Bind b(ctx->device);
    return guarded(ctx, [&]() -> int {
        const size_t npix = (size_t)ctx->frame_w * ctx->frame_h;
        if (join_pipes(ctx)) return 1;
        if (ensure_scratch(ctx, npix * 52)) return 1;
        CK(cudaMemcpyAsync(ctx->scratch, in, npix * 52, cudaMemcpyHostToDevice, ctx->stream));
        k_unpack_stats<<<ctx->shade_grid, 256, 0, ctx->stream>>>((const uint32_t*)ctx->scratch, (uint32_t)npix, ctx->frame,
                                                                 ctx->frame + 8 * npix, ctx->frame + 11 * npix);
        ctx->launches++;
        CK(cudaStreamSynchronize(ctx->stream)); // `in` is the caller's pageable memory
        ctx->frame_has_first = in[0].count > 0; // whole-frame rendering: every pixel holds the same count
        ctx->need_fork = true;
        return 0;
    });
}

int ort_frame_render(ort_ctx* ctx, int32_t ray_depth, uint64_t first_sample, uint64_t n_samples,
                     const volatile uint8_t* interrupt, uint64_t* done) {
    if (!ctx) return 1;
    if (!ctx->frame_w) return fail(ctx, "ort_frame_render: no frame (ort_frame_begin)");
    Bind b(ctx->device);
    return guarded(ctx, [&]() -> int {
        const size_t npix = (size_t)ctx->frame_w * ctx->frame_h;
        uint64_t d = 0;
        if (render_impl(ctx, ctx->frame_w, ctx->frame_h, ray_depth, first_sample, n_samples, ctx->frame,
                        ctx->frame_has_first ? nullptr : ctx->frame + 8 * npix, ctx->frame + 11 * npix, interrupt, &d, true))
            return 1;
        if (d > 0) ctx->frame_has_first = true;
        if (done) *done = d;
        return 0;
    });
}

int ort_frame_wait(ort_ctx* ctx) {
    if (!ctx) return 1;
    Bind b(ctx->device);
    return guarded(ctx, [&]() -> int {
        if (join_pipes(ctx)) return 1;
        CK(cudaStreamSynchronize(ctx->stream));
        return 0;
    });
}

int ort_frame_snapshot(ort_ctx* ctx) {
    if (!ctx) return 1;
    if (!ctx->frame_w) return fail(ctx, "ort_frame_snapshot: no frame (ort_frame_begin)");
    Bind b(ctx->device);
    return guarded(ctx, [&] { return frame_snapshot(ctx); });
}

int ort_frame_preview_rgb8(ort_ctx* ctx, uint8_t* out_rgb) {
    if (!ctx) return 1;
    if (!ctx->frame_w) return fail(ctx, "ort_frame_preview_rgb8: no frame (ort_frame_begin)");
    if (!out_rgb) return fail(ctx, "ort_frame_preview_rgb8: out_rgb is NULL");
    Bind b(ctx->device);
    return guarded(ctx, [&]() -> int {
        const size_t npix = (size_t)ctx->frame_w * ctx->frame_h;
        if (!ctx->frame_snapshot_valid && frame_snapshot(ctx)) return 1;
        // tone-map the snapshot on the side stream: the render pipelines keep running meanwhile
        if (ensure_side(ctx, npix * 52)) return 1;
        CK(cudaStreamWaitEvent(ctx->side_stream, ctx->snap_ev, 0));
        k_tonemap<<<ctx->shade_grid, 256, 0, ctx->side_stream>>>(ctx->frame + 14 * npix, (uint32_t)npix, (uint8_t*)ctx->side_buf);
        ctx->launches++;
        CK(cudaMemcpyAsync(out_rgb, ctx->side_buf, npix * 3, cudaMemcpyDeviceToHost, ctx->side_stream));
        CK(cudaStreamSynchronize(ctx->side_stream));
        ctx->frame_snapshot_valid = false;
        return 0;
    });
}

int ort_frame_fetch(ort_ctx* ctx, ort_sample_stats* out) {
    if (!ctx) return 1;
    if (!ctx->frame_w) return fail(ctx, "ort_frame_fetch: no frame (ort_frame_begin)");
    if (!out) return fail(ctx, "ort_frame_fetch: out is NULL");
    Bind b(ctx->device);
    return guarded(ctx, [&]() -> int {
        const size_t npix = (size_t)ctx->frame_w * ctx->frame_h;
        if (frame_snapshot(ctx)) return 1;
        if (ensure_side(ctx, npix * 52)) return 1;
        const float* snap = ctx->frame + 14 * npix;
        CK(cudaStreamWaitEvent(ctx->side_stream, ctx->snap_ev, 0));
        k_pack_stats<<<ctx->shade_grid, 256, 0, ctx->side_stream>>>(snap, snap + 8 * npix, snap + 11 * npix, (uint32_t)npix, (uint32_t*)ctx->side_buf);
        ctx->launches++;
        CK(cudaMemcpyAsync(out, ctx->side_buf, npix * 52, cudaMemcpyDeviceToHost, ctx->side_stream));
        CK(cudaStreamSynchronize(ctx->side_stream));
        ctx->frame_snapshot_valid = false;
        return 0;
    });
}

int ort_probe_shading(ort_ctx* ctx, int32_t kind, const float* in, int64_t n, float* out) {
    if (!ctx) return 1;
    if (!ctx->has_scene) return fail(ctx, "ort_upload_scene has not been called");
    if (kind < 0 || kind > 7 || n < 0 || n > (1 << 24) || (n > 0 && (!in || !out))) return fail(ctx, "ort_probe_shading: bad arguments");
    if (kind == ORT_PROBE_ENV && !ctx->sd.has_env) return fail(ctx, "ort_probe_shading: the scene has no environment map");
    if (n == 0) return 0;
    Bind b(ctx->device);
    return guarded(ctx, [&]() -> int {
        const size_t ni = (size_t)probe_in_floats(kind), no = (size_t)probe_out_floats(kind);
        if (ensure_scratch(ctx, (size_t)n * (ni + no) * 4)) return 1;
        float* d_in = ctx->scratch;
        float* d_out = ctx->scratch + (size_t)n * ni;
        CK(cudaMemcpyAsync(d_in, in, (size_t)n * ni * 4, cudaMemcpyHostToDevice, ctx->stream));
        const float* lsum = nullptr;
        if (kind == ORT_PROBE_PDF && ctx->sd.n_lights > 0) {
            // pdf's light term: the light-BVH all-hit sum of (pos, out_d), by the same k_trace<true> render uses
            if (join_pipes(ctx)) return 1;
            if (ensure_paths(ctx, n)) return 1;
            if (ensure_counters(ctx, 1)) return 1;
            auto& P = ctx->ps[0];
            CK(cudaMemsetAsync(P.counters, 0, sizeof(uint32_t) * 8, ctx->stream));
            k_probe_pack<<<ctx->shade_grid, 256, 0, ctx->stream>>>(d_in, (uint32_t)n, P.qo[0], P.qd[0], P.counters);
            launch_trace(ctx, P, ctx->stream, P.qo[0], P.qd[0], P.counters, P.counters + 1, 1);
            ctx->launches++;
            lsum = P.lsum;
        }
        k_probe<<<ctx->shade_grid, 256, 0, ctx->stream>>>(ctx->sd, kind, d_in, (uint32_t)n, d_out, lsum);
        ctx->launches++;
        CK(cudaMemcpyAsync(out, d_out, (size_t)n * no * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaGetLastError());
        return 0;
    });
}

int ort_unpack_accum(ort_ctx* ctx, uint32_t w, uint32_t h, const float* d_accum, ort_sample_stats* out) {
    if (!ctx) return 1;
    if (!d_accum || !out) return fail(ctx, "ort_unpack_accum: NULL argument");
    Bind b(ctx->device);
    const size_t npix = (size_t)w * h;
    if (ensure_scratch(ctx, npix * 52)) return 1;
    k_pack_stats<<<ctx->shade_grid, 256, 0, ctx->stream>>>(d_accum, nullptr, nullptr, (uint32_t)npix, (uint32_t*)ctx->scratch);
    ctx->launches++;
    if (ensure_pinned(ctx, npix * 52)) return 1;
    CK(cudaMemcpyAsync(ctx->pinned, ctx->scratch, npix * 52, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const ort_sample_stats* src = (const ort_sample_stats*)ctx->pinned;
    parallel_for(ctx, npix, 1 << 16, [&](size_t f, size_t n) {
    for (size_t i = f; i < f + n; i++) {
        if (src[i].count == 0) continue;
        ort_sample_stats& d = out[i];
        const float mean[3] = {src[i].total[0] / (float)src[i].count, src[i].total[1] / (float)src[i].count,
                               src[i].total[2] / (float)src[i].count};
        if (d.count == 0) std::memcpy(d.first, mean, 12);
        d.count += src[i].count;
        std::memcpy(d.last, mean, 12);
        for (int c = 0; c < 3; c++) { d.total[c] += src[i].total[c]; d.total_squared[c] += src[i].total_squared[c]; }
    }
    });
    return 0;
}

int ort_trace_rays(ort_ctx* ctx, const ort_ray* rays, int64_t n, ort_hit* out) {
    if (!ctx) return 1;
    if (!ctx->has_scene) return fail(ctx, "ort_upload_scene has not been called");
    if (n < 0 || (n > 0 && (!rays || !out))) return fail(ctx, "ort_trace_rays: bad arguments");
    Bind b(ctx->device);
    const int64_t chunk = 1 << 22;
    if (ensure_paths(ctx, std::min<int64_t>(std::max<int64_t>(n, 1), chunk))) return 1;
    if (ensure_counters(ctx, 1)) return 1;
    if (ensure_scratch(ctx, (size_t)std::min<int64_t>(std::max<int64_t>(n, 1), chunk) * 24 * 2)) return 1;
    for (int64_t off = 0; off < n; off += chunk) {
        const uint32_t m = (uint32_t)std::min<int64_t>(chunk, n - off);
        float* d_in = ctx->scratch;
        float* d_out = ctx->scratch + (size_t)std::min<int64_t>(n, chunk) * 6;
        CK(cudaMemcpyAsync(d_in, rays + off, (size_t)m * 24, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(ctx->ps[0].counters, 0, sizeof(uint32_t) * 8, ctx->stream));
        k_pack_rays<<<ctx->shade_grid, 256, 0, ctx->stream>>>(d_in, m, ctx->ps[0].qo[0], ctx->ps[0].qd[0], ctx->ps[0].counters);
        launch_trace(ctx, ctx->ps[0], ctx->stream, ctx->ps[0].qo[0], ctx->ps[0].qd[0], ctx->ps[0].counters, ctx->ps[0].counters + 1, 0);
        k_unpack_hits<<<ctx->shade_grid, 256, 0, ctx->stream>>>(ctx->sd, ctx->ps[0].hits, ctx->ps[0].qd[0], m, d_out);
        ctx->launches += 2;
        CK(cudaMemcpyAsync(out + off, d_out, (size_t)m * 24, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    CK(cudaGetLastError());
    return 0;
}

int ort_bench_trace(ort_ctx* ctx, const ort_ray* rays, int64_t n, int32_t mode, int32_t iters, double* ms_per_launch) {
    if (!ctx) return 1;
    if (!ctx->has_scene) return fail(ctx, "ort_upload_scene has not been called");
    if (n <= 0 || n > ((int64_t)1 << 26) || !rays || !ms_per_launch || iters < 1 || mode < 0 || mode > 1)
        return fail(ctx, "ort_bench_trace: bad arguments");
    if (mode == 1 && ctx->sd.n_lights == 0) return fail(ctx, "ort_bench_trace: scene has no lights");
    Bind b(ctx->device);
    if (ensure_paths(ctx, n)) return 1;
    if (ensure_counters(ctx, 1)) return 1;
    if (ensure_scratch(ctx, (size_t)n * 24)) return 1;
    auto& P = ctx->ps[0];
    CK(cudaMemcpyAsync(ctx->scratch, rays, (size_t)n * 24, cudaMemcpyHostToDevice, ctx->stream));
    k_pack_rays<<<ctx->shade_grid, 256, 0, ctx->stream>>>(ctx->scratch, (uint32_t)n, P.qo[0], P.qd[0], P.counters);
    for (int it = -2; it < iters; it++) { // two warm-up launches
        if (it == 0) CK(cudaEventRecord(ctx->evp0, ctx->stream));
        CK(cudaMemsetAsync(P.counters + 1, 0, sizeof(uint32_t), ctx->stream));
        launch_trace(ctx, P, ctx->stream, P.qo[0], P.qd[0], P.counters, P.counters + 1, mode);
    }
    CK(cudaEventRecord(ctx->evp1, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->evp0, ctx->evp1));
    *ms_per_launch = (double)ms / iters;
    return 0;
}

int ort_bench_read_bw(ort_ctx* ctx, int64_t bytes, int32_t iters, double* gb_per_s) {
    if (!ctx) return 1;
    if (bytes < 32 || iters < 1 || !gb_per_s) return fail(ctx, "ort_bench_read_bw: bad arguments");
    Bind b(ctx->device);
    const size_t n32 = (size_t)bytes / 32;
    if (ensure_scratch(ctx, n32 * 32 + 256)) return 1;
    CK(cudaMemsetAsync(ctx->scratch, 0, n32 * 32 + 256, ctx->stream));
    const int grid = ctx->sm_count * 8;
    float* sink = ctx->scratch + n32 * 8;
    k_read_bw<<<grid, 256, 0, ctx->stream>>>((const float4*)ctx->scratch, n32, 1, sink); // warm-up: fills L2
    CK(cudaEventRecord(ctx->evp0, ctx->stream));
    k_read_bw<<<grid, 256, 0, ctx->stream>>>((const float4*)ctx->scratch, n32, iters, sink);
    CK(cudaEventRecord(ctx->evp1, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    ctx->launches += 2;
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->evp0, ctx->evp1));
    *gb_per_s = (double)n32 * 32.0 * iters / ((double)ms * 1e-3) / 1e9;
    return 0;
}

int ort_light_pdf(ort_ctx* ctx, const ort_ray* rays, int64_t n, float* out) {
    if (!ctx) return 1;
    if (!ctx->has_scene) return fail(ctx, "ort_upload_scene has not been called");
    if (n < 0 || (n > 0 && (!rays || !out))) return fail(ctx, "ort_light_pdf: bad arguments");
    Bind b(ctx->device);
    if (ctx->sd.n_lights == 0) { std::memset(out, 0, sizeof(float) * (size_t)n); return 0; }
    const int64_t chunk = 1 << 22;
    if (ensure_paths(ctx, std::min<int64_t>(std::max<int64_t>(n, 1), chunk))) return 1;
    if (ensure_counters(ctx, 1)) return 1;
    if (ensure_scratch(ctx, (size_t)std::min<int64_t>(std::max<int64_t>(n, 1), chunk) * 24)) return 1;
    for (int64_t off = 0; off < n; off += chunk) {
        const uint32_t m = (uint32_t)std::min<int64_t>(chunk, n - off);
        CK(cudaMemcpyAsync(ctx->scratch, rays + off, (size_t)m * 24, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(ctx->ps[0].counters, 0, sizeof(uint32_t) * 8, ctx->stream));
        k_pack_rays<<<ctx->shade_grid, 256, 0, ctx->stream>>>(ctx->scratch, m, ctx->ps[0].qo[0], ctx->ps[0].qd[0], ctx->ps[0].counters);
        launch_trace(ctx, ctx->ps[0], ctx->stream, ctx->ps[0].qo[0], ctx->ps[0].qd[0], ctx->ps[0].counters, ctx->ps[0].counters + 1, 1);
        k_scale<<<ctx->shade_grid, 256, 0, ctx->stream>>>(ctx->ps[0].lsum, m, (float)ctx->sd.n_lights);
        ctx->launches += 2;
        CK(cudaMemcpyAsync(out + off, ctx->ps[0].lsum, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    CK(cudaGetLastError());
    return 0;
}

int ort_primary_hits(ort_ctx* ctx, uint32_t w, uint32_t h, uint64_t sample, ort_hit* out, ort_ray* rays_out) {
    if (!ctx) return 1;
    if (!ctx->has_scene) return fail(ctx, "ort_upload_scene has not been called");
    if (!out || w == 0 || h == 0) return fail(ctx, "ort_primary_hits: bad arguments");
    Bind b(ctx->device);
    const size_t npix = (size_t)w * h;
    if (ensure_paths(ctx, (int64_t)npix)) return 1;
    if (ensure_counters(ctx, 1)) return 1;
    if (ensure_scratch(ctx, npix * 24)) return 1;
    RenderParams p;
    fill_params(ctx, w, h, 1, &p);
    p.sample_base = sample;
    p.n_batch_samples = 1;
    CK(cudaMemsetAsync(ctx->ps[0].counters, 0, sizeof(uint32_t) * 8, ctx->stream));
    k_raygen<<<ctx->shade_grid, 256, 0, ctx->stream>>>(p, ctx->ps[0].qo[0], ctx->ps[0].qd[0], ctx->ps[0].counters);
    launch_trace(ctx, ctx->ps[0], ctx->stream, ctx->ps[0].qo[0], ctx->ps[0].qd[0], ctx->ps[0].counters, ctx->ps[0].counters + 1, 0);
    k_unpack_hits<<<ctx->shade_grid, 256, 0, ctx->stream>>>(ctx->sd, ctx->ps[0].hits, ctx->ps[0].qd[0], (uint32_t)npix, ctx->scratch);
    ctx->launches += 2;
    CK(cudaMemcpyAsync(out, ctx->scratch, npix * 24, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (rays_out) {
        k_unpack_rays<<<ctx->shade_grid, 256, 0, ctx->stream>>>(ctx->ps[0].qo[0], ctx->ps[0].qd[0], (uint32_t)npix, ctx->scratch);
        ctx->launches++;
        CK(cudaMemcpyAsync(rays_out, ctx->scratch, npix * 24, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    CK(cudaGetLastError());
    return 0;
}

int ort_tonemap_rgb8(ort_ctx* ctx, uint32_t w, uint32_t h, const float* d_accum, uint8_t* out_rgb) {
    if (!ctx) return 1;
    if (!d_accum || !out_rgb) return fail(ctx, "ort_tonemap_rgb8: NULL argument");
    Bind b(ctx->device);
    const size_t npix = (size_t)w * h;
    if (ensure_scratch(ctx, npix * 3)) return 1;
    k_tonemap<<<ctx->shade_grid, 256, 0, ctx->stream>>>(d_accum, (uint32_t)npix, (uint8_t*)ctx->scratch);
    ctx->launches++;
    CK(cudaMemcpyAsync(out_rgb, ctx->scratch, npix * 3, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int ort_get_stats(ort_ctx* ctx, ort_stats* out) {
    if (!ctx || !out) return 1;
    Bind b(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    unsigned long long s[8];
    CK(cudaMemcpy(s, ctx->d_stats, sizeof s, cudaMemcpyDeviceToHost));
    std::memset(out, 0, sizeof *out);
    out->rays_closest = s[0]; out->rays_light_pdf = s[1]; out->paths = s[2]; out->rays_traced = s[3];
    out->kernel_launches = ctx->launches;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->ms_render = ms; else cudaGetLastError();
    out->render_ms = ctx->ms_render;
    out->trace_ms = ctx->ms_trace; out->light_ms = ctx->ms_light; out->shade_ms = ctx->ms_shade; out->other_ms = ctx->ms_other;
    out->wide_nodes = ctx->wide_nodes;
    out->wide_depth = ctx->wide_depth;
    out->light_wide_nodes = ctx->lwide_nodes;
    out->wide_max_stack = ctx->wide_max_stack;
    out->reference_stack_need = ctx->ref_stack_need;
    out->device_bytes = ctx->scene_bytes + ctx->path_bytes;
    return 0;
}

int ort_reset_stats(ort_ctx* ctx) {
    if (!ctx) return 1;
    Bind b(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemset(ctx->d_stats, 0, 8 * sizeof(unsigned long long)));
    ctx->ms_trace = ctx->ms_light = ctx->ms_shade = ctx->ms_other = ctx->ms_render = 0;
    ctx->launches = 0;
    return 0;
}

} // extern "C"

// =================================================================================================
// Several GPUs from one process: sample-index split + one peer-memory reduce per frame.
// =================================================================================================
struct ort_multi {
    std::vector<ort_ctx*> ctx;
    std::vector<char> peer; // devices[0] can dereference ctx[g]'s memory directly
    std::string err;
    uint64_t last_done = 0;
    // device-resident frame
    uint32_t fw = 0, fh = 0;
    int last_g = 0;                  // GPU that rendered the highest sample block of the latest call
    bool snapshot_valid = false;
    cudaStream_t rstream = nullptr;  // reduce / preview / fetch stream on devices[0]
    float* rbuf = nullptr;           // devices[0]: reduced accumulators (8 planes) + last (3 planes) + staging (8 planes)
    size_t rbuf_bytes = 0;
};

namespace {
thread_local std::string g_multi_create_error;

struct PeerPtrs {
    const float* p[16];
};
// dst[i] = (accumulate ? dst[i] : 0) + sum over the sources; 8 planes (total, total_squared, count_lo, count_hi)
// of npix floats each.  Sources on other GPUs are read straight through NVLink peer memory.
__global__ void k_reduce_peers(float* __restrict__ dst, const PeerPtrs src, const int n_src, const size_t n, const int accumulate) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float acc = accumulate ? dst[i] : 0.0f;
        for (int k = 0; k < n_src; k++) acc += src.p[k][i];
        dst[i] = acc;
    }
}
int mfail(ort_multi* m, const std::string& msg) {
    if (m) m->err = msg; else g_multi_create_error = msg;
    return 1;
}
template <typename F>
int mguarded(ort_multi* m, F f) {
    try {
        return f();
    } catch (const std::exception& e) {
        return mfail(m, std::string("exception: ") + e.what());
    } catch (...) {
        return mfail(m, "unknown C++ exception");
    }
}
// contiguous sample blocks, sizes differ by at most one
void split_samples(uint64_t first_sample, uint64_t n_samples, int G, std::vector<uint64_t>* first, std::vector<uint64_t>* cnt) {
    first->assign((size_t)G, 0); cnt->assign((size_t)G, 0);
    const uint64_t base = n_samples / (uint64_t)G, rem = n_samples % (uint64_t)G;
    uint64_t f = first_sample;
    for (int g = 0; g < G; g++) { (*cnt)[g] = base + ((uint64_t)g < rem ? 1 : 0); (*first)[g] = f; f += (*cnt)[g]; }
}
// fn(g) on one host thread per GPU; returns the first failing GPU's index or -1
template <typename F>
int per_gpu(ort_multi* m, F fn) {
    const int G = (int)m->ctx.size();
    std::vector<int> rc((size_t)G, 0);
    auto run = [&](int g) {
        try { rc[g] = fn(g); } catch (const std::exception& e) { rc[g] = fail(m->ctx[g], std::string("exception: ") + e.what()); }
        catch (...) { rc[g] = fail(m->ctx[g], "unknown C++ exception"); }
    };
    std::vector<std::thread> pool;
    for (int g = 1; g < G; g++) pool.emplace_back(run, g);
    run(0);
    for (auto& t : pool) t.join();
    for (int g = 0; g < G; g++) if (rc[g]) return g;
    return -1;
}
int mfail_gpu(ort_multi* m, int g) { return mfail(m, "device " + std::to_string(m->ctx[g]->device) + ": " + ort_last_error(m->ctx[g])); }

// devices[0]: dst (8 planes) = sum over g of src_of(g) (each GPU's 8 accumulator planes), on stream st.
template <typename SrcOf>
int reduce_to_dev0(ort_multi* m, cudaStream_t st, float* dst, float* staging, size_t npix, const std::vector<char>& take, SrcOf src_of) {
    ort_ctx* ctx = m->ctx[0];
    const int G = (int)m->ctx.size();
    PeerPtrs pp{};
    int np = 0;
    for (int g = 0; g < G; g++)
        if (take[g] && m->peer[g]) pp.p[np++] = src_of(g);
    k_reduce_peers<<<ctx->shade_grid, 256, 0, st>>>(dst, pp, np, npix * 8, 0);
    ctx->launches++;
    for (int g = 1; g < G; g++) { // no peer access: stage, add, repeat
        if (!take[g] || m->peer[g]) continue;
        CK(cudaMemcpyPeerAsync(staging, ctx->device, src_of(g), m->ctx[g]->device, npix * 8 * 4, st));
        PeerPtrs one{};
        one.p[0] = staging;
        k_reduce_peers<<<ctx->shade_grid, 256, 0, st>>>(dst, one, 1, npix * 8, 1);
        ctx->launches++;
    }
    CK(cudaGetLastError());
    return 0;
}
int ensure_rbuf(ort_multi* m, size_t npix) {
    ort_ctx* ctx = m->ctx[0];
    const size_t bytes = npix * (8 + 3 + 3 + 8) * 4;
    if (!m->rstream) CK(cudaStreamCreateWithFlags(&m->rstream, cudaStreamNonBlocking));
    if (m->rbuf_bytes >= bytes) return 0;
    if (m->rbuf) { CK(cudaStreamSynchronize(m->rstream)); cudaFree(m->rbuf); }
    m->rbuf = nullptr; m->rbuf_bytes = 0;
    CK(cudaMalloc(&m->rbuf, bytes));
    m->rbuf_bytes = bytes;
    return 0;
}
// Snapshot every GPU's frame and start the peer reduce on devices[0]'s side stream; returns without waiting.
int multi_snapshot(ort_multi* m) {
    const int G = (int)m->ctx.size();
    const size_t npix = (size_t)m->fw * m->fh;
    if (m->rstream) { // an earlier snapshot's reduce may still be reading the per-GPU snapshot buffers
        Bind b(m->ctx[0]->device);
        cudaStreamSynchronize(m->rstream);
    }
    for (int g = 0; g < G; g++) {
        ort_ctx* ctx = m->ctx[g];
        Bind b(ctx->device);
        if (frame_snapshot(ctx)) return mfail_gpu(m, g);
    }
    ort_ctx* ctx = m->ctx[0];
    Bind b(ctx->device);
    auto body = [&]() -> int {
    if (ensure_rbuf(m, npix)) return 1;
    for (int g = 0; g < G; g++) CK(cudaStreamWaitEvent(m->rstream, m->ctx[g]->snap_ev, 0));
    float* red = m->rbuf;
    float* firstp = red + 8 * npix;
    float* lastp = firstp + 3 * npix;
    float* staging = lastp + 3 * npix;
    std::vector<char> take((size_t)G, 1);
    if (reduce_to_dev0(m, m->rstream, red, staging, npix, take, [&](int g) { return (const float*)(m->ctx[g]->frame + 14 * npix); }))
        return 1;
    // first: GPU 0 rendered the frame's first sample; last: the GPU holding the highest block of the latest call
    CK(cudaMemcpyAsync(firstp, m->ctx[0]->frame + 14 * npix + 8 * npix, npix * 3 * 4, cudaMemcpyDeviceToDevice, m->rstream));
    const int lg = m->last_g;
    if (lg == 0) CK(cudaMemcpyAsync(lastp, m->ctx[0]->frame + 14 * npix + 11 * npix, npix * 3 * 4, cudaMemcpyDeviceToDevice, m->rstream));
    else CK(cudaMemcpyPeerAsync(lastp, ctx->device, m->ctx[lg]->frame + 14 * npix + 11 * npix, m->ctx[lg]->device, npix * 3 * 4, m->rstream));
    return 0;
    };
    if (body()) return mfail_gpu(m, 0);
    m->snapshot_valid = true;
    for (int g = 0; g < G; g++) m->ctx[g]->frame_snapshot_valid = false; // consumed by the reduce above
    return 0;
}
} // namespace

extern "C" {

const char* ort_multi_last_error(const ort_multi* m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }

void ort_multi_destroy(ort_multi* m) {
    if (!m) return;
    if (!m->ctx.empty()) {
        Bind b(m->ctx[0]->device);
        if (m->rstream) { cudaStreamSynchronize(m->rstream); cudaStreamDestroy(m->rstream); }
        if (m->rbuf) cudaFree(m->rbuf);
    }
    for (ort_ctx* c : m->ctx) ort_destroy(c);
    delete m;
}

int ort_multi_create(ort_multi** out, const int32_t* devices, int32_t n_devices, uint64_t seed) {
    if (!out || !devices || n_devices < 1 || n_devices > 16) return mfail(nullptr, "ort_multi_create: need 1..16 devices");
    *out = nullptr;
    ort_multi* m = nullptr;
    try { m = new ort_multi(); } catch (...) { return mfail(nullptr, "out of host memory"); }
    for (int g = 0; g < n_devices; g++) {
        ort_device_cfg cfg{};
        cfg.device = devices[g];
        cfg.seed = seed;
        ort_ctx* c = nullptr;
        if (ort_create(&c, &cfg) != 0) {
            g_multi_create_error = std::string("device ") + std::to_string(devices[g]) + ": " + ort_last_error(nullptr);
            ort_multi_destroy(m);
            return 1;
        }
        m->ctx.push_back(c);
    }
    m->peer.assign((size_t)n_devices, 0);
    m->peer[0] = 1;
    for (int g = 1; g < n_devices; g++) {
        if (m->ctx[g]->device == m->ctx[0]->device) { m->peer[g] = 1; continue; }
        int can = 0;
        cudaDeviceCanAccessPeer(&can, m->ctx[0]->device, m->ctx[g]->device);
        if (can) {
            Bind b(m->ctx[0]->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(m->ctx[g]->device, 0);
            if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) m->peer[g] = 1;
            cudaGetLastError();
        }
    }
    *out = m;
    return 0;
}

// All GPUs upload concurrently (one host thread each); the wide-BVH re-emission is done once and shared.
int ort_multi_upload_scene(ort_multi* m, const ort_scene* scene) {
    if (!m) return 1;
    if (!scene) return mfail(m, "ort_multi_upload_scene: scene is NULL");
    return mguarded(m, [&]() -> int {
        SharedWide shared;
        const bool arrays_ok = (scene->n_bvh_nodes <= 0 || scene->bvh) && (scene->n_light_bvh_nodes <= 0 || scene->light_bvh);
        std::thread builder([&] {
            if (arrays_ok) shared.build(scene);
            else { std::lock_guard<std::mutex> l(shared.mu); shared.ready = true; shared.cv.notify_all(); }
        });
        const int bad = per_gpu(m, [&](int g) { return upload_scene_impl(m->ctx[g], scene, &shared); });
        builder.join();
        return bad >= 0 ? mfail_gpu(m, bad) : 0;
    });
}

int ort_multi_render(ort_multi* m, uint32_t w, uint32_t h, int32_t ray_depth, uint64_t first_sample,
                     uint64_t n_samples, ort_sample_stats* out, const volatile uint8_t* interrupt) {
    if (!m) return 1;
    if (!out) return mfail(m, "ort_multi_render: out is NULL");
    return mguarded(m, [&]() -> int {
    const int G = (int)m->ctx.size();
    const size_t npix = (size_t)w * h;
    std::vector<uint64_t> first, cnt, done((size_t)G, 0);
    split_samples(first_sample, n_samples, G, &first, &cnt);
    const int bad = per_gpu(m, [&](int g) -> int {
        ort_ctx* ctx = m->ctx[g];
        Bind b(ctx->device);
        // planes: 8 accum + 3 first + 3 last (+ packed Sample_Stats and a staging area on device 0)
        if (ensure_scratch(ctx, npix * (14 * 4 + 52) + (g == 0 ? npix * 8 * 4 : 0))) return 1;
        float* accum = ctx->scratch;
        CK(cudaMemsetAsync(accum, 0, npix * 14 * 4, ctx->stream));
        if (cnt[g] > 0 && render_impl(ctx, w, h, ray_depth, first[g], cnt[g], accum, g == 0 ? accum + 8 * npix : nullptr,
                                      accum + 11 * npix, interrupt, &done[g]))
            return 1;
        CK(cudaStreamSynchronize(ctx->stream));
        return 0;
    });
    if (bad >= 0) return mfail_gpu(m, bad);
    m->last_done = 0;
    int last_g = 0; // `last` comes from the highest block that actually rendered something (interrupts)
    std::vector<char> take((size_t)G, 0);
    for (int g = 0; g < G; g++) { m->last_done += done[g]; if (done[g] > 0) { last_g = g; take[g] = 1; } }
    take[0] = 1;

    ort_ctx* ctx = m->ctx[0];
    Bind b(ctx->device);
    float* accum = ctx->scratch;
    float* firstp = accum + 8 * npix;
    float* lastp = firstp + 3 * npix;
    uint32_t* packed = (uint32_t*)(lastp + 3 * npix);
    float* staging = (float*)((char*)packed + npix * 52);
    // the one reduce of the frame: devices[0] sums its peers' accumulators into its own
    {
        PeerPtrs pp{};
        int np = 0;
        for (int g = 1; g < G; g++) {
            if (!take[g]) continue;
            if (m->peer[g]) { pp.p[np++] = m->ctx[g]->scratch; continue; }
            CK(cudaMemcpyPeerAsync(staging, ctx->device, m->ctx[g]->scratch, m->ctx[g]->device, npix * 8 * 4, ctx->stream));
            PeerPtrs one{};
            one.p[0] = staging;
            k_reduce_peers<<<ctx->shade_grid, 256, 0, ctx->stream>>>(accum, one, 1, npix * 8, 1);
            ctx->launches++;
        }
        if (np > 0) {
            k_reduce_peers<<<ctx->shade_grid, 256, 0, ctx->stream>>>(accum, pp, np, npix * 8, 1);
            ctx->launches++;
        }
    }
    if (last_g != 0)
        CK(cudaMemcpyPeerAsync(lastp, ctx->device, m->ctx[last_g]->scratch + 11 * npix, m->ctx[last_g]->device,
                               npix * 3 * 4, ctx->stream));
    CK(cudaGetLastError());
    if (pack_and_merge(ctx, accum, firstp, lastp, npix, packed, out)) return mfail(m, ort_last_error(ctx));
    return 0;
    });
}

uint64_t ort_multi_last_render_samples(const ort_multi* m) { return m ? m->last_done : 0; }

int ort_multi_frame_begin(ort_multi* m, uint32_t w, uint32_t h) {
    if (!m) return 1;
    return mguarded(m, [&]() -> int {
        const int bad = per_gpu(m, [&](int g) { return ort_frame_begin(m->ctx[g], w, h); });
        if (bad >= 0) return mfail_gpu(m, bad);
        m->fw = w; m->fh = h; m->last_g = 0; m->snapshot_valid = false;
        return 0;
    });
}

int ort_multi_frame_end(ort_multi* m) {
    if (!m) return 1;
    return mguarded(m, [&]() -> int {
        if (m->rstream) { Bind b(m->ctx[0]->device); cudaStreamSynchronize(m->rstream); }
        const int bad = per_gpu(m, [&](int g) { return ort_frame_end(m->ctx[g]); });
        m->fw = m->fh = 0;
        return bad >= 0 ? mfail_gpu(m, bad) : 0;
    });
}

// The loaded image goes to devices[0]; the other GPUs start from zero (their partial sums are added at fetch).
int ort_multi_frame_load(ort_multi* m, const ort_sample_stats* in) {
    if (!m) return 1;
    if (!m->fw) return mfail(m, "ort_multi_frame_load: no frame (ort_multi_frame_begin)");
    if (ort_frame_load(m->ctx[0], in)) return mfail_gpu(m, 0);
    return 0;
}

int ort_multi_frame_render(ort_multi* m, int32_t ray_depth, uint64_t first_sample, uint64_t n_samples,
                           const volatile uint8_t* interrupt, uint64_t* done_out) {
    if (!m) return 1;
    if (!m->fw) return mfail(m, "ort_multi_frame_render: no frame (ort_multi_frame_begin)");
    return mguarded(m, [&]() -> int {
        const int G = (int)m->ctx.size();
        std::vector<uint64_t> first, cnt, done((size_t)G, 0);
        split_samples(first_sample, n_samples, G, &first, &cnt);
        const int bad = per_gpu(m, [&](int g) -> int {
            return cnt[g] > 0 ? ort_frame_render(m->ctx[g], ray_depth, first[g], cnt[g], interrupt, &done[g]) : 0;
        });
        if (bad >= 0) return mfail_gpu(m, bad);
        m->last_done = 0;
        for (int g = 0; g < G; g++) { m->last_done += done[g]; if (done[g] > 0) m->last_g = g; } // ascending: ends on the highest block
        if (done_out) *done_out = m->last_done;
        return 0;
    });
}

int ort_multi_frame_wait(ort_multi* m) {
    if (!m) return 1;
    return mguarded(m, [&]() -> int {
        const int bad = per_gpu(m, [&](int g) { return ort_frame_wait(m->ctx[g]); });
        return bad >= 0 ? mfail_gpu(m, bad) : 0;
    });
}

int ort_multi_frame_snapshot(ort_multi* m) {
    if (!m) return 1;
    if (!m->fw) return mfail(m, "ort_multi_frame_snapshot: no frame (ort_multi_frame_begin)");
    return mguarded(m, [&] { return multi_snapshot(m); });
}

int ort_multi_frame_preview_rgb8(ort_multi* m, uint8_t* out_rgb) {
    if (!m) return 1;
    if (!m->fw) return mfail(m, "ort_multi_frame_preview_rgb8: no frame (ort_multi_frame_begin)");
    if (!out_rgb) return mfail(m, "ort_multi_frame_preview_rgb8: out_rgb is NULL");
    return mguarded(m, [&]() -> int {
        if (!m->snapshot_valid && multi_snapshot(m)) return 1;
        ort_ctx* ctx = m->ctx[0];
        Bind b(ctx->device);
        const size_t npix = (size_t)m->fw * m->fh;
        if (ensure_side(ctx, npix * 52)) return mfail_gpu(m, 0);
        k_tonemap<<<ctx->shade_grid, 256, 0, m->rstream>>>(m->rbuf, (uint32_t)npix, (uint8_t*)ctx->side_buf);
        ctx->launches++;
        if (cudaMemcpyAsync(out_rgb, ctx->side_buf, npix * 3, cudaMemcpyDeviceToHost, m->rstream) != cudaSuccess ||
            cudaStreamSynchronize(m->rstream) != cudaSuccess)
            return mfail(m, std::string("preview: ") + cudaGetErrorString(cudaGetLastError()));
        m->snapshot_valid = false;
        return 0;
    });
}

int ort_multi_frame_fetch(ort_multi* m, ort_sample_stats* out) {
    if (!m) return 1;
    if (!m->fw) return mfail(m, "ort_multi_frame_fetch: no frame (ort_multi_frame_begin)");
    if (!out) return mfail(m, "ort_multi_frame_fetch: out is NULL");
    return mguarded(m, [&]() -> int {
        if (multi_snapshot(m)) return 1;
        ort_ctx* ctx = m->ctx[0];
        Bind b(ctx->device);
        const size_t npix = (size_t)m->fw * m->fh;
        if (ensure_side(ctx, npix * 52)) return mfail_gpu(m, 0);
        k_pack_stats<<<ctx->shade_grid, 256, 0, m->rstream>>>(m->rbuf, m->rbuf + 8 * npix, m->rbuf + 11 * npix, (uint32_t)npix, (uint32_t*)ctx->side_buf);
        ctx->launches++;
        if (cudaMemcpyAsync(out, ctx->side_buf, npix * 52, cudaMemcpyDeviceToHost, m->rstream) != cudaSuccess ||
            cudaStreamSynchronize(m->rstream) != cudaSuccess)
            return mfail(m, std::string("fetch: ") + cudaGetErrorString(cudaGetLastError()));
        m->snapshot_valid = false;
        return 0;
    });
}

int ort_multi_get_stats(ort_multi* m, ort_stats* out) {
    if (!m || !out) return 1;
    std::memset(out, 0, sizeof *out);
    for (size_t g = 0; g < m->ctx.size(); g++) {
        ort_stats s;
        if (ort_get_stats(m->ctx[g], &s)) return mfail(m, ort_last_error(m->ctx[g]));
        out->rays_closest += s.rays_closest; out->rays_traced += s.rays_traced; out->rays_light_pdf += s.rays_light_pdf;
        out->paths += s.paths; out->kernel_launches += s.kernel_launches;
        out->render_ms = std::max(out->render_ms, s.render_ms);
        out->trace_ms = std::max(out->trace_ms, s.trace_ms); out->light_ms = std::max(out->light_ms, s.light_ms);
        out->shade_ms = std::max(out->shade_ms, s.shade_ms); out->other_ms = std::max(out->other_ms, s.other_ms);
        out->wide_nodes = s.wide_nodes; out->wide_depth = s.wide_depth; out->light_wide_nodes = s.light_wide_nodes;
        out->wide_max_stack = s.wide_max_stack; out->reference_stack_need = s.reference_stack_need;
        out->device_bytes += s.device_bytes;
    }
    return 0;
}

} // extern "C"
