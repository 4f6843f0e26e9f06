"""Bisect the slow chunks of the primary queue (tools/primary_chunks.py) down to single rays and print them."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from raytracer_odin_b200 import api, cabi
from raytracer_odin_b200.scene import native_bvh_build

config = sys.argv[1] if len(sys.argv) > 1 else "C4"
scene, cfg = bench.build_scene(config, native_bvh_build)
w, h = cfg["width"], cfg["height"]
spp = max(8, (1 << 25) // (w * h) // 8 * 8)
r = api.Renderer(seed=bench.SEED).upload_scene(scene)
rays = np.stack([r.primary_hits(w, h, s, want_rays=True)[1] for s in range(spp)])
n = spp * w * h
i = np.arange(n, dtype=np.int64)
block, inn = i >> 5, i & 31
sin, pin = inn // 4, inn % 4
sgroups = spp // 8
pg, sg = block // sgroups, block % sgroups
tiles_x = w // 2
ty, tx = pg // tiles_x, pg % tiles_x
pix = (ty * 2 + pin // 2) * w + tx * 2 + pin % 2
smp = sg * 8 + sin
q = rays[smp, pix]
C = 64
per = n // C
slow = []
for c in range(C):
    ms = r.bench_trace(q[c * per:(c + 1) * per], 0, 2)
    if ms > 0.6:
        slow.append((c, ms))
print("slow chunks", slow, flush=True)
st = r.stats()
print("wide depth", st["wide_depth"], "max_stack", st["wide_max_stack"], "root box", scene.bvh[-1]["lo"], scene.bvh[-1]["hi"], "cam", scene.cam_pos)
for c, ms in slow:
    lo, hi = c * per, (c + 1) * per
    while hi - lo > 1:
        mid = (lo + hi) // 2
        a = r.bench_trace(q[lo:mid], 0, 2)
        b = r.bench_trace(q[mid:hi], 0, 2)
        if a >= b: hi = mid
        else: lo = mid
        if max(a, b) < 0.3: break
    idx = np.arange(lo, hi)
    for k in idx[:4]:
        ray = q[k]
        t1 = r.bench_trace(q[k:k + 1], 0, 2)
        hit = r.trace_rays(q[k:k + 1])
        print(json.dumps({"chunk": c, "queue_index": int(k), "pixel": [int(pix[k] % w), int(pix[k] // w)], "sample": int(smp[k]),
                          "o": [float(x) for x in ray["o"]], "d": [repr(float(x)) for x in ray["d"]],
                          "d_hex": [hex(int(x)) for x in ray["d"].view(np.uint32)], "single_ray_ms": round(t1, 3),
                          "hit": [float(hit["t"][0]), int(hit["tri"][0])]}), flush=True)
