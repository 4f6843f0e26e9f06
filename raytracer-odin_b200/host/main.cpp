// odinrt — C++ command line of the B200 renderer, mirroring the reference's (main.odin:174-253):
//
//   odinrt <input.gltf> [output.ppm|png] --width W --height H --ray-depth D --num-samples N
//          [--env-map file.hdr] [--times T] [--continious] [--threads n]
//          [--gpus 0,1,...] [--seed S] [--bvh host|device] [--checkpoint f] [--resume f] [--chunk spp]
//          [--preview file.png|ppm] [--preview-every seconds] [--duration seconds]
//
// The accumulators live on the GPU(s) for the whole run (ort_frame_*): --continious enqueues chunks of
// `--chunk` samples PER GPU back to back (default 64), the periodic preview is tone-mapped on the device
// (3 bytes per pixel to the host, written by a background thread), and the 52-byte-per-pixel Sample_Stats
// cross the bus once, at exit.  --duration stops a continuous run after that many seconds (like a SIGINT).
//
// read_gltf -> (width/height/fov/env-map overrides) -> finish_scene -> render_scene -> save_result,
// with render_scene being the C ABI of libodinrt_b200.so (include/odinrt_b200.h).  Like the
// reference, render parameters default to zero when omitted (main.odin:199-206).  --threads is
// accepted and ignored (no CPU workers on this path); --debug (SDL window) is not part of this build.
// Errors are fatal like the reference's panics (main.odin:195,216): message on stderr, exit code 1.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "scene.hpp"

namespace {

volatile uint8_t g_interrupt = 0; // async_interrupt (main.odin:170-172)
void on_sigint(int) { g_interrupt = 1; }

[[noreturn]] void die(const std::string& msg) {
    std::fprintf(stderr, "%s\n", msg.c_str());
    std::exit(1);
}

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Args {
    std::string input_file, output_file, env_map, gpus = "0", bvh = "host", checkpoint, resume, preview;
    long times = 0, threads = 0, width = 0, height = 0, ray_depth = 0, num_samples = 0, chunk = 64;
    double preview_every = 5.0, duration = 0.0;
    unsigned long long seed = 0;
    bool continious = false, debug = false;
};

Args parse_args(int argc, char** argv) {
    Args a;
    int positional = 0;
    for (int i = 1; i < argc; i++) {
        std::string s = argv[i];
        if (s.size() > 1 && s[0] == '-') {
            s = s.substr(s[1] == '-' ? 2 : 1);
            std::string val;
            bool has_val = false;
            const size_t eq = s.find_first_of("=:");
            if (eq != std::string::npos) { val = s.substr(eq + 1); s = s.substr(0, eq); has_val = true; }
            auto value = [&]() -> std::string {
                if (has_val) return val;
                if (i + 1 >= argc) die("missing value for flag --" + s);
                return argv[++i];
            };
            auto num = [&]() { return std::strtol(value().c_str(), nullptr, 10); };
            if (s == "continious") a.continious = true;
            else if (s == "debug") a.debug = true;
            else if (s == "times") a.times = num();
            else if (s == "threads") a.threads = num();
            else if (s == "width") a.width = num();
            else if (s == "height") a.height = num();
            else if (s == "ray-depth" || s == "ray_depth") a.ray_depth = num();
            else if (s == "num-samples" || s == "num_samples") a.num_samples = num();
            else if (s == "env-map" || s == "env_map") a.env_map = value();
            else if (s == "gpus") a.gpus = value();
            else if (s == "seed") a.seed = std::strtoull(value().c_str(), nullptr, 10);
            else if (s == "bvh") a.bvh = value();
            else if (s == "checkpoint") a.checkpoint = value();
            else if (s == "resume") a.resume = value();
            else if (s == "chunk") a.chunk = num();
            else if (s == "preview") a.preview = value();
            else if (s == "preview-every" || s == "preview_every") a.preview_every = std::strtod(value().c_str(), nullptr);
            else if (s == "duration") a.duration = std::strtod(value().c_str(), nullptr);
            else if (s == "help" || s == "h") {
                std::printf("usage: odinrt <input.gltf> [output.ppm|png] --width W --height H --ray-depth D --num-samples N\n"
                            "       [--env-map f.hdr] [--times T] [--continious] [--threads n] [--gpus 0,1,..] [--seed S]\n"
                            "       [--bvh host|device] [--checkpoint f] [--resume f] [--chunk spp-per-gpu]\n"
                            "       [--preview f.png] [--preview-every s] [--duration s]\n");
                std::exit(0);
            } else die("unknown flag: --" + s);
        } else if (positional == 0) { a.input_file = s; positional++; }
        else if (positional == 1) { a.output_file = s; positional++; }
        else die("unexpected argument: " + s);
    }
    if (a.input_file.empty()) die("missing required argument: input_file");
    return a;
}

// Raw accumulator checkpoint (same file format as api.save_checkpoint):
// "ORTCKPT1" | u32 width | u32 height | u64 next_sample | width*height Sample_Stats (52 bytes each)
bool save_checkpoint(const std::string& path, uint32_t w, uint32_t h, uint64_t next, const std::vector<ort_sample_stats>& px) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    bool ok = std::fwrite("ORTCKPT1", 1, 8, f) == 8 && std::fwrite(&w, 4, 1, f) == 1 && std::fwrite(&h, 4, 1, f) == 1 &&
              std::fwrite(&next, 8, 1, f) == 1 && std::fwrite(px.data(), sizeof(ort_sample_stats), px.size(), f) == px.size();
    return std::fclose(f) == 0 && ok;
}
bool load_checkpoint(const std::string& path, uint32_t w, uint32_t h, uint64_t* next, std::vector<ort_sample_stats>* px) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    char magic[8];
    uint32_t fw = 0, fh = 0;
    bool ok = std::fread(magic, 1, 8, f) == 8 && !std::memcmp(magic, "ORTCKPT1", 8) && std::fread(&fw, 4, 1, f) == 1 &&
              std::fread(&fh, 4, 1, f) == 1 && std::fread(next, 8, 1, f) == 1 && fw == w && fh == h;
    if (ok) ok = std::fread(px->data(), sizeof(ort_sample_stats), px->size(), f) == px->size();
    std::fclose(f);
    return ok;
}

} // namespace

int main(int argc, char** argv) {
    std::signal(SIGINT, on_sigint);
    const Args a = parse_args(argc, argv);
    if (a.debug) std::fprintf(stderr, "--debug: the SDL debug window is not part of this build; ignored\n");

    orh::HostScene scene;
    std::string err;
    if (!orh::read_gltf(a.input_file, &scene, &err)) die("Failed to parse gltf: " + err); // main.odin:195
    if (a.height != 0) { // main.odin:200-204
        const float aspect = (float)a.width / (float)a.height;
        scene.cam.fov_x *= aspect;
    }
    if (!a.env_map.empty()) { // main.odin:213-220
        if (!orh::load_texture(a.env_map, &scene.env_map, &err)) die("Failed to load environment map: " + err);
        scene.has_env = true;
    }
    std::vector<int32_t> devices;
    for (size_t p = 0; p <= a.gpus.size();) {
        const size_t q = std::min(a.gpus.find(',', p), a.gpus.size());
        if (q > p) devices.push_back((int32_t)std::strtol(a.gpus.substr(p, q - p).c_str(), nullptr, 10));
        p = q + 1;
    }
    if (devices.empty()) die("--gpus: no device given");

    double t0 = now_s();
    if (!orh::finish_scene(&scene, a.bvh == "device" ? devices[0] : -1, &err)) die("finish_scene: " + err);
    std::printf("Scene + light BVH built in %.1fms (%zu triangles, %zu lights)\n", (now_s() - t0) * 1e3, scene.triangles.size(),
                scene.light_triangles.size());

    ort_scene view;
    std::vector<ort_texture> tex_keep;
    orh::scene_view(scene, &view, &tex_keep);

    ort_ctx* ctx = nullptr;
    ort_multi* multi = nullptr;
    if (devices.size() == 1) {
        ort_device_cfg cfg{};
        cfg.device = devices[0];
        cfg.seed = a.seed;
        if (ort_create(&ctx, &cfg)) die(std::string("ort_create: ") + ort_last_error(nullptr));
        if (ort_upload_scene(ctx, &view)) die(std::string("ort_upload_scene: ") + ort_last_error(ctx));
    } else {
        if (ort_multi_create(&multi, devices.data(), (int32_t)devices.size(), a.seed)) die(std::string("ort_multi_create: ") + ort_multi_last_error(nullptr));
        if (ort_multi_upload_scene(multi, &view)) die(std::string("ort_multi_upload_scene: ") + ort_multi_last_error(multi));
    }
    // the frame: accumulators stay in HBM until the end of the run
    const uint32_t w = (uint32_t)a.width, h = (uint32_t)a.height;
    auto check = [&](int rc, const char* what) {
        if (rc) die(std::string(what) + ": " + (ctx ? ort_last_error(ctx) : ort_multi_last_error(multi)));
    };
    auto frame_render = [&](uint64_t first, uint64_t n) -> uint64_t {
        uint64_t done = 0;
        check(ctx ? ort_frame_render(ctx, (int32_t)a.ray_depth, first, n, &g_interrupt, &done)
                  : ort_multi_frame_render(multi, (int32_t)a.ray_depth, first, n, &g_interrupt, &done), "render");
        return done;
    };
    auto frame_wait = [&] { check(ctx ? ort_frame_wait(ctx) : ort_multi_frame_wait(multi), "wait"); };
    check(ctx ? ort_frame_begin(ctx, w, h) : ort_multi_frame_begin(multi, w, h), "frame_begin");

    std::vector<ort_sample_stats> pixels((size_t)w * h); // create_rendering_context: zeroed Sample_Stats
    uint64_t first = 0;
    if (!a.resume.empty()) {
        if (!load_checkpoint(a.resume, w, h, &first, &pixels))
            die("checkpoint " + a.resume + " does not match a " + std::to_string(w) + "x" + std::to_string(h) + " Sample_Stats image");
        check(ctx ? ort_frame_load(ctx, pixels.data()) : ort_multi_frame_load(multi, pixels.data()), "frame_load");
    }

    if (a.continious) { // samples = max(int): render until interrupted (main.odin:207)
        const uint64_t chunk = (uint64_t)std::max(1L, a.chunk) * devices.size(); // --chunk samples per GPU and call
        uint64_t rendered = 0;
        std::vector<uint8_t> rgb((size_t)w * h * 3), rgb_writing;
        std::thread writer;
        double host_blocked = 0; // time the render loop spent in preview calls (device keeps rendering meanwhile)
        int previews = 0;
        t0 = now_s();
        double next_preview = t0 + a.preview_every;
        // progress every 10 s: samples ENQUEUED (the device lags by at most a few waves per GPU, the host is
        // throttled by the interrupt flag), i.e. the sustained rate of the whole loop incl. previews
        double last_report = t0;
        uint64_t last_rendered = 0;
        std::vector<double> interval_rates;
        while (!g_interrupt) {
            rendered += frame_render(first, chunk);
            first += chunk; // an interrupted call still consumes its index range: no sample index is ever reused
            const double now = now_s();
            if (now - last_report >= 10.0) {
                const double rate = (double)(rendered - last_rendered) * (double)w * (double)h / (now - last_report) / 1e6;
                std::printf("t=%.1fs  %llu samples per pixel  %.1f Msamples/s over the last %.1fs\n", now - t0,
                            (unsigned long long)rendered, rate, now - last_report);
                std::fflush(stdout);
                interval_rates.push_back(rate);
                last_report = now;
                last_rendered = rendered;
            }
            if (a.duration > 0 && now - t0 >= a.duration) g_interrupt = 1;
            if (!a.preview.empty() && now >= next_preview && !g_interrupt) {
                // get_rgb_image on the device (output.odin:30-80): snapshot now, enqueue the next chunk, THEN wait
                // for the 3-byte-per-pixel image — the GPUs never idle while the host reads or encodes it
                check(ctx ? ort_frame_snapshot(ctx) : ort_multi_frame_snapshot(multi), "snapshot");
                rendered += frame_render(first, chunk);
                first += chunk;
                const double tp = now_s();
                check(ctx ? ort_frame_preview_rgb8(ctx, rgb.data()) : ort_multi_frame_preview_rgb8(multi, rgb.data()), "preview");
                host_blocked += now_s() - tp;
                if (writer.joinable()) writer.join();
                rgb_writing = rgb;
                writer = std::thread([&, path = a.preview] {
                    std::string e;
                    const bool png = path.size() > 4 && path.compare(path.size() - 4, 4, ".png") == 0;
                    if (!(png ? orh::write_png(path, (int)w, (int)h, rgb_writing.data(), &e) : orh::write_ppm(path, (int)w, (int)h, rgb_writing.data(), &e)))
                        std::fprintf(stderr, "preview: %s\n", e.c_str());
                });
                previews++;
                next_preview = now_s() + a.preview_every;
            }
        }
        frame_wait();
        const double dt = now_s() - t0;
        if (writer.joinable()) writer.join();
        ort_stats st{};
        if (ctx) ort_get_stats(ctx, &st); else ort_multi_get_stats(multi, &st);
        std::printf("Rendered %llu samples in %.2fs\n", (unsigned long long)rendered, dt);
        std::printf("%.1f Mrays/s sustained, %.1f Msamples/s, %d previews (host blocked %.1f ms in total)\n",
                    (double)st.rays_closest / dt / 1e6, (double)st.paths / dt / 1e6, previews, host_blocked * 1e3);
        if (interval_rates.size() > 1) { // steady state: the 10-s intervals after the first
            std::vector<double> ss(interval_rates.begin() + 1, interval_rates.end());
            std::sort(ss.begin(), ss.end());
            const double med = ss[ss.size() / 2], rays_per_sample = st.paths ? (double)st.rays_closest / (double)st.paths : 0.0;
            std::printf("steady state after 10 s: median %.1f Msamples/s = %.1f Mrays/s (min %.1f, max %.1f Msamples/s over %zu intervals)\n",
                        med, med * rays_per_sample, ss.front(), ss.back(), ss.size());
        }
    } else {
        const long trials = a.times > 0 ? a.times : 1;
        std::vector<double> timings;
        uint64_t last_done = 0;
        for (long trial = 0; trial < trials; trial++) { // render_scene (raytracer.odin:606-624): trials replay the same samples
            t0 = now_s();
            last_done = frame_render(first, (uint64_t)std::max(0L, a.num_samples));
            frame_wait();
            timings.push_back(now_s() - t0);
            std::printf("Trial %ld >>> Rendered in %.3fms\n", trial, timings.back() * 1e3);
        }
        // one GPU renders a contiguous prefix, so an interrupted run resumes exactly where it stopped; several GPUs leave
        // gaps inside their blocks: the whole range counts as consumed (no index is ever rendered twice; the per-pixel
        // count in the checkpoint says how many samples the image really holds)
        first += (ctx && g_interrupt) ? last_done : (uint64_t)std::max(0L, a.num_samples);
        ort_stats st{};
        if (ctx) ort_get_stats(ctx, &st); else ort_multi_get_stats(multi, &st);
        double total = 0;
        for (double t : timings) total += t;
        std::printf("%.1f Mrays/s, %.1f Msamples/s\n", (double)st.rays_closest / total / 1e6, (double)st.paths / total / 1e6);
        if (trials > 1) { // raytracer.odin:625-664
            std::vector<double> ts = timings;
            std::sort(ts.begin(), ts.end());
            double mean = total / (double)ts.size(), var = 0;
            for (double t : ts) var += (t - mean) * (t - mean);
            const double sd = std::sqrt(var / (double)(ts.size() - 1));
            const double med = ts.size() % 2 ? ts[ts.size() / 2] : 0.5 * (ts[ts.size() / 2 - 1] + ts[ts.size() / 2]);
            std::printf(">>>>>>>>> Performance Summary <<<<<<<<<\nTrials: %ld\nTime: %.02f±%.02fms\n"
                        "Best: %.02fms, Median: %.02fms, Worst: %.02fms\n>>>>>>>>> Performance Summary <<<<<<<<<\n",
                        trials, mean * 1e3, sd * 1e3, ts.front() * 1e3, med * 1e3, ts.back() * 1e3);
        }
    }
    // the one 52-byte-per-pixel transfer of the run
    t0 = now_s();
    check(ctx ? ort_frame_fetch(ctx, pixels.data()) : ort_multi_frame_fetch(multi, pixels.data()), "frame_fetch");
    std::printf("Sample_Stats fetched in %.1fms\n", (now_s() - t0) * 1e3);
    if (!a.checkpoint.empty() && !save_checkpoint(a.checkpoint, w, h, first, pixels)) die("failed to write checkpoint " + a.checkpoint);
    if (!a.output_file.empty() && !orh::save_result(pixels.data(), (int)w, (int)h, a.output_file, &err)) die(err);
    if (ctx) ort_destroy(ctx);
    if (multi) ort_multi_destroy(multi);
    return 0;
}
