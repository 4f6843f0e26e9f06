"""Sweep a tunable (env var read at ort_create) and print Mrays/s + per-kernel-class ms on a config."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from raytracer_odin_b200 import api
from raytracer_odin_b200.scene import native_bvh_build

def main():
    config = sys.argv[1] if len(sys.argv) > 1 else "C2"
    spp = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    var = sys.argv[3] if len(sys.argv) > 3 else "ORT_REFILL"
    values = sys.argv[4].split(",") if len(sys.argv) > 4 else ["0", "16", "20", "24", "28"]
    scene, cfg = bench.build_scene(config, native_bvh_build)
    w, h, depth = cfg["width"], cfg["height"], cfg["ray_depth"]
    acc = torch.zeros(8, w * h, device="cuda")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    for v in values:
        for kv in v.split("+"):
            k, _, val = kv.rpartition("=")
            os.environ[k or var] = val
        r = api.Renderer(seed=1).upload_scene(scene)
        r.set_stream(torch.cuda.current_stream().cuda_stream)
        r.render_device(w, h, depth, 0, spp, acc.data_ptr()); torch.cuda.synchronize()
        r.reset_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r.render_device(w, h, depth, 100, spp, acc.data_ptr()); r.render_device(w, h, depth, 200, spp, acc.data_ptr()); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1); st = r.stats()
        r.reset_stats(); r.set_profiling(True)
        r.render_device(w, h, depth, 300, spp, acc.data_ptr()); torch.cuda.synchronize()
        ps = r.stats()
        print(json.dumps({"set": v, "Mrays/s": round(st["rays_closest"] / ms / 1e3, 1), "ms": round(ms / 2, 2),
                          "trace_ms": round(ps["trace_ms"], 2), "light_ms": round(ps["light_ms"], 2),
                          "shade_ms": round(ps["shade_ms"], 2), "other_ms": round(ps["other_ms"], 2),
                          "trace_Grays/s": round(ps["rays_traced"] / ps["trace_ms"] / 1e6, 3)}), flush=True)
        r.close()

if __name__ == "__main__":
    main()
