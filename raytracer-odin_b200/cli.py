"""CLI mirroring the reference (main.odin:174-253):

    python -m raytracer_odin_b200.cli <gltf> <out.ppm|png> --width W --height H --ray-depth D \
        --num-samples N [--env-map file.hdr] [--times T] [--continious] [--gpus 0,1,..] [--seed S]
        [--checkpoint acc.npy] [--resume acc.npy]

Like the reference, render parameters default to zero when omitted (main.odin:199-206), so
--width/--height/--ray-depth/--num-samples are effectively mandatory.  --threads is accepted and
ignored (the GPU path has no CPU worker threads); --debug (SDL window) is not part of this build.
"""
import argparse
import signal
import time

import numpy as np

from . import api, cabi, gltf, output
from .scene import device_bvh_build, native_bvh_build


def main(argv=None):
    ap = argparse.ArgumentParser(prog="raytracer_odin_b200")
    ap.add_argument("input_file")
    ap.add_argument("output_file", nargs="?", default="")
    ap.add_argument("--times", "-times", type=int, default=0)
    ap.add_argument("--continious", "-continious", action="store_true")
    ap.add_argument("--threads", "-threads", type=int, default=0)
    ap.add_argument("--width", "-width", type=int, default=0)
    ap.add_argument("--height", "-height", type=int, default=0)
    ap.add_argument("--ray-depth", "-ray-depth", type=int, default=0)
    ap.add_argument("--num-samples", "-num-samples", type=int, default=0)
    ap.add_argument("--env-map", "-env-map", default="")
    ap.add_argument("--gpus", default="0", help="comma separated CUDA device ordinals (one scene replica each)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--bvh", choices=("host", "device"), default="host", help="where finish_scene builds the BVHs")
    ap.add_argument("--checkpoint", default="", help="write the raw Sample_Stats accumulators here on exit (.npy)")
    ap.add_argument("--resume", default="", help="continue from accumulators written by --checkpoint")
    a = ap.parse_args(argv)

    scene = gltf.read_gltf(a.input_file)
    if a.height:
        scene.fov_x = scene.apply_render_config(a.width, a.height)  # main.odin:200-204
    if a.env_map:
        scene.env_map = gltf.load_texture(a.env_map)
    t0 = time.perf_counter()
    scene.finish(native_bvh_build if a.bvh == "host" else device_bvh_build)
    print(f"Scene + light BVH built in {(time.perf_counter() - t0) * 1e3:.1f}ms")

    interrupt = np.zeros(1, np.uint8)
    signal.signal(signal.SIGINT, lambda *_: interrupt.__setitem__(0, 1))  # main.odin:170-172

    devices = [int(x) for x in a.gpus.split(",")]
    r = (api.Renderer(device=devices[0], seed=a.seed) if len(devices) == 1
         else api.MultiRenderer(devices, seed=a.seed)).upload_scene(scene)
    w, h = a.width, a.height
    # the frame: accumulators stay in HBM for the whole run (ort_frame_*), like host/main.cpp
    r.frame_begin(w, h)
    resume_at = 0
    if a.resume:  # raw accumulators + the next sample index: the counter-based streams continue seamlessly
        pixels, resume_at = api.load_checkpoint(a.resume, w, h)
        r.frame_load(pixels)
    if a.continious:  # samples = max(int): render chunks until interrupted (main.odin:207)
        first, chunk, rendered = resume_at, 64 * len(devices), 0
        t0 = time.perf_counter()
        while not interrupt[0]:
            rendered += r.frame_render(a.ray_depth, first, chunk, interrupt)
            first += chunk
        r.frame_wait()
        print(f"Rendered {rendered} samples in {time.perf_counter() - t0:.2f}s")
    else:
        trials = a.times if a.times > 0 else 1
        timings = []
        for trial in range(trials):
            t0 = time.perf_counter()
            r.frame_render(a.ray_depth, resume_at, a.num_samples, interrupt)
            r.frame_wait()
            timings.append(time.perf_counter() - t0)
            print(f"Trial {trial} >>> Rendered in {timings[-1] * 1e3:.3f}ms")
        first = resume_at + a.num_samples
        st = r.stats()
        total = sum(timings)
        print(f"{st['rays_closest'] / total / 1e6:.1f} Mrays/s, {st['paths'] / total / 1e6:.1f} Msamples/s")
    pixels = r.frame_fetch()  # the one 52-byte-per-pixel transfer of the run
    if a.checkpoint:
        api.save_checkpoint(a.checkpoint, pixels, first, w, h)
    if a.output_file:
        output.save_result(pixels, w, h, a.output_file)
    r.close()


if __name__ == "__main__":
    main()
