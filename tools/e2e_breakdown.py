"""Where the end-to-end (host-buffer) step spends its time: to_c, ort_upload_scene, ort_render vs the
device-timed ort_render_device of the same samples.  python tools/e2e_breakdown.py [C2] [spp]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from raytracer_odin_b200 import api  # noqa: E402
from raytracer_odin_b200.scene import native_bvh_build  # noqa: E402


def main():
    config = sys.argv[1] if len(sys.argv) > 1 else "C2"
    spp = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    scene, cfg = bench.build_scene(config, native_bvh_build)
    w, h, depth = cfg["width"], cfg["height"], cfg["ray_depth"]
    npix = w * h
    r = api.Renderer(seed=1)
    out = np.zeros(npix, api.cabi.STATS_DTYPE)
    acc = torch.zeros(8, npix, device="cuda")
    rows = []
    for it in range(4):
        t0 = time.perf_counter()
        scene.to_c()
        t1 = time.perf_counter()
        r.upload_scene(scene)
        t2 = time.perf_counter()
        r.render(w, h, depth, spp, it * spp, out)
        t3 = time.perf_counter()
        r.render_device(w, h, depth, it * spp, spp, acc.data_ptr())
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        st = r.stats()
        rows.append({"to_c_ms": (t1 - t0) * 1e3, "upload_ms": (t2 - t1) * 1e3, "render_host_ms": (t3 - t2) * 1e3,
                     "render_device_ms": (t4 - t3) * 1e3, "render_ms_events": st["render_ms"]})
        print(json.dumps(rows[-1]), flush=True)
    r.close()


if __name__ == "__main__":
    main()
