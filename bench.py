#!/usr/bin/env python
"""bench.py — Mrays/s of the path-tracing hot path on BASELINE.json's configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C4] [--spp S] [--impl reference]

A *step* is one FRAME of the named config: `spp` samples of every pixel (default C4, the configuration
BASELINE.json's target is quoted on: procedural 1 M-triangle scene, 1920x1080, ray-depth 10, 4096 spp —
configs[3]).  The scene replica is resident in HBM before the timed region.  With N GPUs the frame's
sample range is DIVIDED over the ranks (strong scaling: total work fixed, `multigpu.sample_partition`),
every rank renders its block into device accumulators, and ONE NCCL reduce per frame combines them on
rank 0 — inside the timed region.  --config C2 / C3 / C5 select the other BASELINE configs.

  value     whole-job Mrays/s (closest-hit cast_ray calls the reference would make / s / 1e6),
            device-timed with CUDA events over exactly K frames, max over ranks.
  e2e       the same metric through the public host calls with HOST buffers, per frame:
            ort_upload_scene (scene H2D) + render + reduce + ONE host Sample_Stats image (D2H, 52 B/px) on
            rank 0, wall clock, max over ranks.  N = 1: ort_render; N > 1: ort_render_device + the NCCL
            reduce + ort_unpack_accum.
  roofline  dominant kernel k_trace<closest>: algorithmic bytes/ray (32 + 16 + 24*N_box +
            36*N_tri, N_* from the oracle's duplicate-free traversal of the same BVH) x rays / CUDA-event
            time of its launches (a separate pass with per-kernel-class events on the library's stream).
            bound = "l2" when the traversal working set (nodes + triangle records) is L2 resident (C1-C4):
            peak = the L2 read bandwidth measured in this run at that working-set size; bound = "hbm" (C5):
            peak = MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the CPU restatement of the reference (oracle, faithful traversal, reference task
            scheduling, all host threads) on a bounded sample of the same frame — every k-th 4x4 tile over
            the WHOLE frame at a few spp — plus `ideal_value`: the same with the reference's duplicate-left
            push (raytracer.odin:395-409) removed.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED = 20261018
L2_RESIDENT_BYTES = 100 << 20  # traversal working sets below this are served from the 126 MB L2


def build_scene(config, finish_with, scale=None):
    from raytracer_odin_b200 import gltf, scenegen

    cfg = dict(scenegen.CONFIGS[config])
    d = tempfile.mkdtemp(prefix=f"ort_{config}_")
    path, env = scenegen.generate(config, d, **(scale or {}))
    s = gltf.read_gltf(path)
    s.fov_x = s.apply_render_config(cfg["width"], cfg["height"])
    if env:
        s.env_map = gltf.load_texture(env)
    s.finish(finish_with)
    return s, cfg


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md, no MEASURED_PEAKS.json)"


def cpu_leg(scene, cfg, target_s):
    """The CPU restatement on a bounded sample of the frame: every `stride`-th 4x4 tile in x and y over the
    WHOLE frame, `spp` samples, sized for ~target_s seconds of wall time.  Returns run(mode) -> (Mrays/s,
    counters, seconds), the thread count and a description of the sample."""
    from oracle import binding as orc

    w, h, depth = cfg["width"], cfg["height"], cfg["ray_depth"]
    nat = orc.OracleScene(scene, native=True)
    threads = max(1, nat.lib.orc_hardware_threads())
    # pilot: every 16th tile in x and y (1/256 of the frame), 1 spp -> size the sample for ~target_s of wall time
    t0 = time.perf_counter()
    _, c = nat.render(w, h, depth, 1, seed=SEED, mode=0, schedule=0, threads=threads, tile_stride=16)
    pilot_s = max(time.perf_counter() - t0, 1e-4)
    full_frame_1spp_s = pilot_s * 256.0          # the whole frame at 1 spp at the pilot's rate
    spp = 4
    stride = int(max(1, min(16, round((full_frame_1spp_s * spp / target_s) ** 0.5))))
    if stride == 1:                               # even the whole frame is too short: more samples per pixel
        spp = int(max(4, min(64, round(target_s / full_frame_1spp_s))))

    def run(mode=0):
        t0 = time.perf_counter()
        _, c = nat.render(w, h, depth, spp, seed=SEED, mode=mode, schedule=0, threads=threads, tile_stride=stride)
        dt = time.perf_counter() - t0
        return c["rays"] / dt / 1e6, c, dt

    sample = (f"rate on a sample: one 4x4-pixel tile of every {stride}x{stride} tiles over the whole {w}x{h} frame "
              f"(1/{stride * stride} of the pixels), {spp} spp, depth {depth}")
    return run, threads, sample, nat


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path.  The Odin toolchain
    does not exist in this image, so this is the C++ restatement (oracle/, kind 'port')."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding as orc

    scene, cfg = build_scene(args.config, orc.bvh_build)  # the oracle's own bvh_build: nothing of the library on this arm
    run, threads, sample, _ = cpu_leg(scene, cfg, target_s=6.0)
    for _ in range(args.warmup):
        run()
    vals, t0 = [], time.perf_counter()
    for _ in range(args.steps):
        vals.append(run())
    total = time.perf_counter() - t0
    rays = sum(v[1]["rays"] for v in vals)
    value = rays / sum(v[2] for v in vals) / 1e6
    ideal_v, _, _ = run(mode=1)
    spp = args.spp or cfg["spp"] or 64
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.config, cfg, scene, spp),
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample,
                         "ideal_value": ideal_v,
                         "note": "CPU restatement of the reference render loop (faithful traversal order incl. the "
                                 "duplicate-left push, reference 4x4x32 task scheduling); Odin toolchain unavailable. "
                                 "ideal_value: the same loop with the duplicate push removed (BASELINE.md §3)"},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(name, cfg, scene, spp):
    return {"workload": f"BASELINE config {name}: generated {len(scene.triangles)}-triangle glTF "
                        f"({cfg['gen']}), {cfg['width']}x{cfg['height']}, ray-depth {cfg['ray_depth']}, "
                        f"{spp} spp per frame (one step = one frame, sample-split over the GPUs)",
            "triangles": int(len(scene.triangles)), "light_triangles": int(len(scene.light_triangles)),
            "width": cfg["width"], "height": cfg["height"], "ray_depth": cfg["ray_depth"],
            "spp_per_step": spp,
            "l2": "per-step path state (>1 GB per wave) and queues exceed the 126 MB L2; no explicit flush"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="C4")
    ap.add_argument("--spp", type=int, default=0, help="samples per frame (default: the config's spp; C5: 64)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--bvh", choices=("host", "device"), default="host",
                    help="builder used by finish_scene before the timed region (identical output)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    from raytracer_odin_b200 import api, multigpu
    from raytracer_odin_b200.scene import device_bvh_build, native_bvh_build

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL writes its banner ("NCCL version ...") to stdout when the communicator is created: keep
        # stdout for the one JSON line by pointing fd 1 at stderr until the first collective has run
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.all_reduce(torch.zeros(1, device="cuda"))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    scene, cfg = build_scene(args.config, native_bvh_build if args.bvh == "host"
                             else (lambda t: device_bvh_build(t, local)))
    w, h, depth = cfg["width"], cfg["height"], cfg["ray_depth"]
    spp = args.spp or cfg["spp"] or 64
    npix = w * h

    r = api.Renderer(device=local, seed=SEED)
    r.upload_scene(scene)
    stream = torch.cuda.Stream()  # explicit stream: the library, the events and NCCL all order on it
    torch.cuda.set_stream(stream)
    r.set_stream(stream.cuda_stream)
    accum = torch.zeros(8, npix, device="cuda", dtype=torch.float32)

    def step(i):
        # one frame: the frame's spp samples are divided over the ranks (strong scaling); trials replay
        # disjoint sample ranges (frame i renders samples [i*spp, (i+1)*spp))
        first, cnt = multigpu.sample_partition(i * spp, spp, rank, world)
        accum.zero_()
        if cnt > 0:
            r.render_device(w, h, depth, first, cnt, accum.data_ptr())
        if world > 1:
            multigpu.reduce_accum(accum, 0)  # the one collective of the path: one NCCL reduce per frame

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    r.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    with ClockSampler(local) as clocks:
        torch.cuda.synchronize()
        e0.record(stream)
        for i in range(args.steps):
            step(args.warmup + i)
            marks[i].record(stream)  # per-frame times: best / median / worst like the reference's summary
        e1.record(stream)
        torch.cuda.synchronize()
    step_ms = [a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks)]
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    st = r.stats()
    t = torch.tensor([ms, float(st["rays_closest"]), float(st["rays_traced"]), float(st["paths"]),
                      float(st["kernel_launches"]), float(st["rays_light_pdf"])], device="cuda", dtype=torch.float64)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms = float(tmax[0])
    rays, traced, paths, launches, light_rays = (float(x) for x in t[1:])
    value = rays / (ms * 1e-3) / 1e6
    # the frame on rank 0 holds every rank's samples: count plane == spp everywhere
    if rank == 0:
        cnt_plane = accum[6] + accum[7] * float(1 << 20)
        assert float(cnt_plane.min()) == float(cnt_plane.max()) == float(spp), "reduced frame does not hold spp samples per pixel"

    # ---- e2e: public host calls, HOST buffers; per frame: scene H2D + render + reduce + ONE Sample_Stats image D2H
    h2d = (scene.triangles.nbytes + scene.bvh.nbytes + scene.light_triangles.nbytes + scene.light_bvh.nbytes +
           scene.materials.nbytes + sum(t_.nbytes for t_ in scene.textures) +
           (scene.env_map.nbytes if scene.env_map is not None else 0))
    out = np.zeros(npix, dtype=api.cabi.STATS_DTYPE)
    e2e_steps = max(2, min(args.steps, 3))
    e2e_parts = []
    if world == 1:
        r.set_stream(None)

    def e2e_step(i):
        first, cnt = multigpu.sample_partition(i * spp, spp, rank, world)
        ta = time.perf_counter()
        r.upload_scene(scene)
        tb = time.perf_counter()
        if world == 1:
            r.render(w, h, depth, cnt, first, out)
        else:
            accum.zero_()
            if cnt > 0:
                r.render_device(w, h, depth, first, cnt, accum.data_ptr())
            multigpu.reduce_accum(accum, 0)
            if rank == 0:
                r.unpack_accum(w, h, accum.data_ptr(), out)  # the frame's ONE host image (synchronises)
            else:
                torch.cuda.synchronize()
        e2e_parts.append(((tb - ta) * 1e3, (time.perf_counter() - tb) * 1e3))

    e2e_step(args.warmup + args.steps)
    out[:] = 0
    r.reset_stats()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(args.warmup + args.steps + 1 + i)
    if world > 1:
        dist.barrier()
    e2e_s = time.perf_counter() - t0
    e2e_rays = float(r.stats()["rays_closest"])
    te = torch.tensor([e2e_s, e2e_rays], device="cuda", dtype=torch.float64)
    if world > 1:
        tm = te.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(te, op=dist.ReduceOp.SUM)
        e2e_s = float(tm[0])
    e2e_value = float(te[1]) / e2e_s / 1e6
    if rank == 0:
        assert int(out["count"].min()) == int(out["count"].max()) == spp * e2e_steps, "e2e image is not ONE full frame per step"

    # ---- roofline of the dominant kernel (rank 0): a separate pass with per-kernel-class CUDA events
    roof, cpu = None, None
    if rank == 0:
        r.set_stream(stream.cuda_stream)
        r.reset_stats()
        r.set_profiling(True)
        wave_paths = int(os.environ.get("ORT_WAVE_PATHS", 1 << 25))  # library default: 2^25 paths per wave
        spp_per_wave = max(1, wave_paths // npix)
        prof_spp = min(spp, 16 * spp_per_wave)  # 16 waves: per-launch times do not depend on the frame's length
        accum.zero_()
        r.render_device(w, h, depth, 10_000_000, prof_spp, accum.data_ptr())
        torch.cuda.synchronize()
        ps = r.stats()
        r.set_profiling(False)
        hbm_peak, hbm_src = measured_peak()
        counts = None
        if not args.no_cpu:
            run, threads, sample, nat = cpu_leg(scene, cfg, target_s=10.0)
            v, c, dt = run(0)
            vi, ci, dti = run(1)
            cpu = {"value": v, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample,
                   "seconds": dt, "rays": c["rays"],
                   "ideal_value": vi, "ideal_seconds": dti,
                   "reference_node_pops_per_ray": c["node_pops"] / c["rays"],
                   "reference_tri_tests_per_ray": c["tri_tests"] / c["rays"],
                   "ideal_node_pops_per_ray": ci["node_pops"] / ci["rays"],
                   "ideal_tri_tests_per_ray": ci["tri_tests"] / ci["rays"],
                   "gpu_over_cpu": {"faithful": value / v, "ideal": value / vi, "n_gpus": world, "cores": threads},
                   "note": "value: the reference as written (duplicate-left push, raytracer.odin:395-409); "
                           "ideal_value: the same loop without the duplicate push"}
            counts = {"n_box": ci["box_tests"] / ci["rays"], "n_tri": ci["tri_tests"] / ci["rays"]}
        bpr = (32 + 16 + 24 * counts["n_box"] + 36 * counts["n_tri"]) if counts else None
        trace_s = ps["trace_ms"] * 1e-3
        n_launch = depth * ((prof_spp + spp_per_wave - 1) // spp_per_wave)  # one k_trace<closest> per bounce per wave
        traffic = None
        tpath = os.path.join(ROOT, "profiles", f"ncu_traffic_{args.config.lower()}.json")
        if os.path.exists(tpath):  # dram__bytes_read+write per launch from the committed ncu --set full capture
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("spp_per_wave") == spp_per_wave:
                traffic = tj["k_trace_closest"]["dram_bytes_per_launch_avg"]
        achieved = (bpr * ps["rays_traced"] / trace_s / 1e9) if bpr else None
        # the traversal working set (nodes + triangle records) of C1-C4 is L2 resident, so the roof
        # is the box's L2 read bandwidth, measured here with the library's streaming-read probe on a working set
        # of the scene's size (at least 32 MB to stay out of L1); C5 (2.3 GB) streams from HBM
        ws = ps["wide_nodes"] * 128 + ps["light_wide_nodes"] * 128 + (len(scene.triangles) + len(scene.light_triangles)) * 64
        alg_per_launch = (bpr * ps["rays_traced"] / n_launch) if bpr else None
        # ... or whose MEASURED DRAM traffic (ncu) is a small part of the algorithmic bytes: C5's 2.3 GB do not fit the
        # L2, but its hot upper levels do — ncu shows ~2 GB of DRAM traffic per launch against ~14 GB algorithmic
        served_by_l2 = traffic is not None and alg_per_launch is not None and traffic < 0.25 * alg_per_launch
        if achieved is not None and achieved > hbm_peak:  # more algorithmic bytes per second than HBM can deliver:
            served_by_l2 = True                            # they cannot be streaming from HBM
        l2_bound = ws < L2_RESIDENT_BYTES or served_by_l2
        probe = min(max(ws, 32 << 20), 96 << 20) if l2_bound else min(max(ws, 1 << 30), 4 << 30)
        read_gbs = r.bench_read_bw(probe, 20 if probe <= (256 << 20) else 5)
        if l2_bound:
            peak, peak_src = read_gbs, ("measured in this run: ort_bench_read_bw (256-bit read-only loads from all SMs) over "
                                        f"a {probe >> 20} MB working set = L2 read bandwidth"
                                        + ("" if ws < L2_RESIDENT_BYTES else
                                           f"; the {ws >> 20} MB working set exceeds the L2, but "
                                           + (f"ncu's DRAM traffic per launch is {traffic / alg_per_launch:.0%} of the algorithmic bytes"
                                              if traffic else "the algorithmic rate exceeds the HBM copy peak")
                                           + ": the traversal is served by the L2 (hot upper levels of the tree)"))
        else:
            peak, peak_src = hbm_peak, hbm_src
        roof = {"bound": "l2" if l2_bound else "hbm", "kernel": "k_trace<closest>", "unit": "GB/s",
                "achieved": achieved, "peak": peak, "frac": (achieved / peak) if achieved else None,
                "peak_source": peak_src, "traffic": traffic,
                "working_set_bytes": int(ws),
                "hbm": {"peak": hbm_peak, "peak_source": hbm_src, "unit": "GB/s",
                        "achieved": (traffic / (ps["trace_ms"] / n_launch * 1e-3) / 1e9) if traffic else None,
                        "frac": (traffic / (ps["trace_ms"] / n_launch * 1e-3) / 1e9 / hbm_peak) if traffic else None,
                        "what": "MEASURED DRAM bytes per launch (ncu, profiles/ncu_traffic_*.json) / average launch time of this run"},
                "bytes_per_ray": bpr, "n_box": counts["n_box"] if counts else None,
                "n_tri": counts["n_tri"] if counts else None,
                "profiled_pass": f"{prof_spp} spp ({n_launch // depth} waves of {spp_per_wave} spp), kernel classes serialised",
                "rays_per_pass": ps["rays_traced"], "launches_per_pass": n_launch,
                "kernel_ms_per_pass": ps["trace_ms"], "avg_launch_ms": ps["trace_ms"] / n_launch,
                "algorithmic_bytes_per_launch": alg_per_launch,
                "trace_grays_per_s": ps["rays_traced"] / trace_s / 1e9,
                "pass_breakdown_ms": {"trace": ps["trace_ms"], "light": ps["light_ms"], "shade": ps["shade_ms"],
                                      "other": ps["other_ms"]},
                "note": "the limiting resource of the traversal is issue slots x SIMD efficiency (ncu: 13-15 of 32 lanes "
                        "per instruction on bounce rays), not bytes: frac says how far the kernel sits from a design that "
                        "streams the algorithmic bytes at the roof; `traffic` (ncu dram bytes per launch) far below "
                        "algorithmic_bytes_per_launch means the bytes come from L2/L1"}
        # compulsory HBM traffic of the wavefront design per traced ray and bounce — k_trace reads the ray (32 B) and
        # writes the hit (16 B); k_shade reads ray, hit, pending payload and light sum (84 B) and writes the next ray +
        # payload + light-queue entry (68 B) — against the HBM peak
        state_bpr = 32 + 16 + 84 + 68
        roof["state_traffic"] = {"bytes_per_ray_bounce": state_bpr, "unit": "GB/s",
                                 "achieved": state_bpr * traced / world / (ms * 1e-3) / 1e9,  # per GPU
                                 "frac": state_bpr * traced / world / (ms * 1e-3) / 1e9 / hbm_peak}

    if rank == 0:
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.config, cfg, scene, spp),
            "samples_per_s": paths / (ms * 1e-3), "frames_1080p_spp_per_s": paths / (ms * 1e-3) / 2073600.0,
            "rays_traced_per_s": traced / (ms * 1e-3), "light_pdf_rays_per_s": light_rays / (ms * 1e-3),
            "step_ms": {"best": min(step_ms), "median": float(np.median(step_ms)), "worst": max(step_ms)},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d) * world,
                    "d2h_bytes_per_step": int(npix * 52), "steps": e2e_steps,
                    "upload_ms": [round(a, 2) for a, _ in e2e_parts[1:]],
                    "render_ms": [round(b, 2) for _, b in e2e_parts[1:]],
                    "what": ("ort_upload_scene (host scene -> HBM, wide-BVH re-emission) + ort_render into host Sample_Stats"
                             if world == 1 else
                             "per rank ort_upload_scene + ort_render_device of its sample block, ONE NCCL reduce, "
                             "ort_unpack_accum into ONE host Sample_Stats image on rank 0") + "; wall clock, max over ranks"},
            "gpu_launches": int(launches), "clocks": clocks.summary(),
            "roofline": roof, "cpu_baseline": cpu,
            "wide_bvh": {"nodes": st["wide_nodes"], "depth": st["wide_depth"], "device_bytes": st["device_bytes"],
                         "max_stack": st["wide_max_stack"], "reference_stack_need": st["reference_stack_need"]},
        }
        print(json.dumps(line), flush=True)
    r.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
