#!/usr/bin/env python
"""bench.py — Mrays/s of the path-tracing hot path on BASELINE.json's configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl reference]

A *step* is one pass of the hot path over one batch: `spp` samples of every pixel of the named
config (default C2: generated ~100k-triangle instanced-sphere glTF, 1920x1080, ray-depth 8,
256 spp — BASELINE.json configs[1]).  The scene replica is resident in HBM before the timed region.

  value     whole-job Mrays/s (closest-hit cast_ray calls the reference would make / s / 1e6),
            device-timed with CUDA events over exactly K steps, max over ranks.
  e2e       the same metric through the public host call (ort_upload_scene + ort_render with HOST
            buffers): scene H2D and Sample_Stats D2H inside the timed region.
  roofline  dominant kernel k_trace: algorithmic bytes/ray (32 + 16 + 24*N_box +
            36*N_tri, N_* from the oracle's duplicate-free traversal of the same BVH) x rays /
            CUDA-event time of the k_trace launches of one step; peak = MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the CPU restatement of the reference (oracle, faithful traversal, reference task
            scheduling, all host threads) on a bounded pixel window of the same workload.

N > 1 (torchrun): every rank holds a scene replica and renders its own block of sample indices
(weak scaling: per-GPU work fixed); one NCCL reduce of the accumulators per step.
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


def cpu_leg(scene, cfg, target_s, ideal_counts=True):
    """Time the CPU restatement on a centred pixel window sized for ~target_s seconds of wall time;
    also count the duplicate-free traversal's box / triangle tests per ray (algorithmic need)."""
    from oracle import binding as orc

    w, h, depth = cfg["width"], cfg["height"], cfg["ray_depth"]
    nat = orc.OracleScene(scene, native=True)
    threads = max(1, nat.lib.orc_hardware_threads())

    def window(nx, ny):
        return (w // 2 - nx // 2, h // 2 - ny // 2, w // 2 - nx // 2 + nx, h // 2 - ny // 2 + ny)

    # pilot: 64x36 window, 1 spp
    t0 = time.perf_counter()
    _, c = nat.render(w, h, depth, 1, seed=SEED, mode=0, schedule=0, threads=threads, window=window(64, 36))
    pilot_s = max(time.perf_counter() - t0, 1e-4)
    rate = c["rays"] / pilot_s
    rays_per_px = c["rays"] / (64 * 36)
    want_px = max(64 * 36, rate * target_s / max(rays_per_px, 1e-9))
    spp = 4
    nx = int(min(w, max(64, (want_px / spp * 16 / 9) ** 0.5))) // 4 * 4
    ny = int(min(h, max(36, nx * 9 // 16))) // 4 * 4
    win = window(nx, ny)

    def run():
        t0 = time.perf_counter()
        _, c = nat.render(w, h, depth, spp, seed=SEED, mode=0, schedule=0, threads=threads, window=win)
        dt = time.perf_counter() - t0
        return c["rays"] / dt / 1e6, c, dt

    sample = f"{nx}x{ny} centre window of the {w}x{h} frame, {spp} spp, depth {depth}"
    counts = None
    if ideal_counts:
        chk = orc.OracleScene(scene, native=True)
        _, ci = chk.render(w, h, depth, 1, seed=SEED, mode=1, schedule=0, threads=threads, window=window(256, 144))
        counts = {"n_box": ci["box_tests"] / ci["rays"], "n_tri": ci["tri_tests"] / ci["rays"],
                  "node_pops": ci["node_pops"] / ci["rays"], "rays": ci["rays"]}
    return run, threads, sample, counts


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path.  The Odin toolchain
    does not exist in this image, so this is the C++ restatement (oracle/, kind 'port')."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding as orc

    scene, cfg = build_scene(args.config, orc.bvh_build)
    run, threads, sample, _ = cpu_leg(scene, cfg, target_s=8.0, ideal_counts=False)
    for _ in range(args.warmup):
        run()
    vals, t0 = [], time.perf_counter()
    for _ in range(args.steps):
        vals.append(run())
    total = time.perf_counter() - t0
    rays = sum(v[1]["rays"] for v in vals)
    value = rays / sum(v[2] for v in vals) / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.config, cfg, scene, None),
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "CPU restatement of the reference render loop (faithful traversal order incl. the "
                                 "duplicate-left push, reference 4x4x32 task scheduling); Odin toolchain unavailable"},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(name, cfg, scene, spp):
    return {"workload": f"BASELINE config {name}: generated {len(scene.triangles)}-triangle glTF "
                        f"({cfg['gen']}), {cfg['width']}x{cfg['height']}, ray-depth {cfg['ray_depth']}, "
                        f"{spp if spp else cfg['spp']} spp per step",
            "triangles": int(len(scene.triangles)), "light_triangles": int(len(scene.light_triangles)),
            "width": cfg["width"], "height": cfg["height"], "ray_depth": cfg["ray_depth"],
            "spp_per_step": spp if spp else cfg["spp"],
            "l2": "per-step path state (>1 GB) and queues exceed the 126 MB L2; no explicit flush"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="C2")
    ap.add_argument("--spp", type=int, default=0, help="samples per step (default: the config's spp)")
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
        # weak scaling: every rank renders `spp` samples of its own block of the global sample axis
        first, cnt = multigpu.sample_partition(i * spp * world, spp * world, rank, world)
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
            marks[i].record(stream)  # per-step times: best / median / worst like the reference's summary
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

    # ---- e2e: public host call, HOST buffers, scene H2D + Sample_Stats D2H inside the timed region
    r.set_stream(None)
    cs, keep = scene.to_c()
    h2d = (scene.triangles.nbytes + scene.bvh.nbytes + scene.light_triangles.nbytes + scene.light_bvh.nbytes +
           scene.materials.nbytes + sum(t_.nbytes for t_ in scene.textures) +
           (scene.env_map.nbytes if scene.env_map is not None else 0))
    out = np.zeros(npix, dtype=api.cabi.STATS_DTYPE)
    e2e_steps = max(2, min(args.steps, 3))

    e2e_parts = []

    def e2e_step(i):
        first, cnt = multigpu.sample_partition(i * spp * world, spp * world, rank, world)
        ta = time.perf_counter()
        r.upload_scene(scene)
        tb = time.perf_counter()
        r.render(w, h, depth, cnt, first, out)
        e2e_parts.append(((tb - ta) * 1e3, (time.perf_counter() - tb) * 1e3))

    e2e_step(0)
    r.reset_stats()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(1 + i)
    e2e_s = time.perf_counter() - t0
    e2e_rays = float(r.stats()["rays_closest"])
    te = torch.tensor([e2e_s, e2e_rays], device="cuda", dtype=torch.float64)
    if world > 1:
        tm = te.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(te, op=dist.ReduceOp.SUM)
        e2e_s = float(tm[0])
    e2e_value = float(te[1]) / e2e_s / 1e6

    # ---- roofline of the dominant kernel (rank 0): one extra step with per-kernel-class CUDA events
    roof, cpu = None, None
    if rank == 0:
        r.set_stream(stream.cuda_stream)
        r.reset_stats()
        r.set_profiling(True)
        r.render_device(w, h, depth, 10_000_000, spp, accum.data_ptr())
        torch.cuda.synchronize()
        ps = r.stats()
        r.set_profiling(False)
        peak, peak_src = measured_peak()
        counts = None
        if not args.no_cpu:
            run, threads, sample, counts = cpu_leg(scene, cfg, target_s=12.0)
            v, c, dt = run()
            cpu = {"value": v, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample,
                   "seconds": dt, "rays": c["rays"],
                   "reference_node_pops_per_ray": c["node_pops"] / c["rays"],
                   "reference_tri_tests_per_ray": c["tri_tests"] / c["rays"]}
        if counts:
            bpr = 32 + 16 + 24 * counts["n_box"] + 36 * counts["n_tri"]
        else:
            bpr = None
        trace_s = ps["trace_ms"] * 1e-3
        wave_paths = int(os.environ.get("ORT_WAVE_PATHS", 1 << 25))  # library default: 2^25 paths per wave
        spp_per_wave = max(1, wave_paths // npix)
        n_launch = depth * ((spp + spp_per_wave - 1) // spp_per_wave)  # one k_trace<closest> per bounce per wave
        traffic = None
        tpath = os.path.join(ROOT, "profiles", f"ncu_traffic_{args.config.lower()}.json")
        if os.path.exists(tpath):  # dram__bytes_read+write per launch from the committed ncu --set full capture
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("spp_per_wave") == spp_per_wave:
                traffic = tj["k_trace_closest"]["dram_bytes_per_launch_avg"]
        roof = {"bound": "hbm", "kernel": "k_trace", "unit": "GB/s", "peak": peak, "peak_source": peak_src,
                "bytes_per_ray": bpr, "n_box": counts["n_box"] if counts else None,
                "n_tri": counts["n_tri"] if counts else None,
                "rays_per_step": ps["rays_traced"], "launches_per_step": n_launch,
                "kernel_ms_per_step": ps["trace_ms"], "avg_launch_ms": ps["trace_ms"] / n_launch,
                "achieved": (bpr * ps["rays_traced"] / trace_s / 1e9) if bpr else None,
                "traffic": traffic,
                "algorithmic_bytes_per_launch": (bpr * ps["rays_traced"] / n_launch) if bpr else None,
                "trace_grays_per_s": ps["rays_traced"] / trace_s / 1e9,
                "step_breakdown_ms": {"trace": ps["trace_ms"], "light": ps["light_ms"], "shade": ps["shade_ms"],
                                      "other": ps["other_ms"]},
                "note": "C1-C4 traversal working sets sit in the 126 MB L2 (SURVEY §8d), so achieved "
                        "algorithmic GB/s may exceed the HBM copy peak; ncu dram bytes are in profiles/"}
        roof["frac"] = (roof["achieved"] / peak) if roof["achieved"] else None
        # second fraction: compulsory HBM traffic of the wavefront design per traced ray and bounce —
        # k_trace reads the ray (32 B) and writes the hit (16 B); k_shade reads ray, hit, pending payload and
        # light sum (84 B) and writes the next ray + payload + light-queue entry (68 B) — against the HBM peak
        state_bpr = 32 + 16 + 84 + 68
        roof["state_traffic"] = {"bytes_per_ray_bounce": state_bpr, "unit": "GB/s",
                                 "achieved": state_bpr * traced / world / (ms * 1e-3) / 1e9,  # per GPU
                                 "frac": state_bpr * traced / world / (ms * 1e-3) / 1e9 / peak}
        # C1-C4 traversal working sets are L2 resident, so the same algorithmic GB/s is also
        # held against the box's L2 read bandwidth, measured here with the library's streaming-read probe
        # on a working set of the scene's size (nodes + traversal triangles, at least 32 MB to stay out of L1)
        ws = ps["wide_nodes"] * 128 + ps["light_wide_nodes"] * 128 + (len(scene.triangles) + len(scene.light_triangles)) * 64
        probe = max(ws, 32 << 20)
        l2_gbs = r.bench_read_bw(probe, 20 if probe <= (256 << 20) else 5)
        roof["l2"] = {"working_set_bytes": int(ws), "probe_bytes": int(probe), "peak": l2_gbs, "unit": "GB/s",
                      "peak_source": "measured in this run: ort_bench_read_bw, 256-bit read-only loads from all SMs",
                      "frac": (roof["achieved"] / l2_gbs) if roof["achieved"] else None}

    if rank == 0:
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.config, cfg, scene, spp),
            "samples_per_s": paths / (ms * 1e-3), "frames_1080p_spp_per_s": paths / (ms * 1e-3) / 2073600.0,
            "rays_traced_per_s": traced / (ms * 1e-3), "light_pdf_rays_per_s": light_rays / (ms * 1e-3),
            "step_ms": {"best": min(step_ms), "median": float(np.median(step_ms)), "worst": max(step_ms)},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(npix * 52), "steps": e2e_steps,
                    "upload_ms": [round(a, 2) for a, _ in e2e_parts[1:]],
                    "render_ms": [round(b, 2) for _, b in e2e_parts[1:]],
                    "what": "ort_upload_scene (host scene -> HBM, wide-BVH re-emission) + ort_render into host "
                            "Sample_Stats, wall clock"},
            "gpu_launches": int(launches), "clocks": clocks.summary(),
            "roofline": roof, "cpu_baseline": cpu,
            "wide_bvh": {"nodes": st["wide_nodes"], "depth": st["wide_depth"], "device_bytes": st["device_bytes"]},
        }
        print(json.dumps(line), flush=True)
    r.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
