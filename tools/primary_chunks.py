"""Where does the primary-ray launch of a config spend its time?  Rebuilds the render's primary queue (2x2 pixels x
8 samples per warp, 16 spp), times k_trace<closest> on the whole queue and on 64 consecutive chunks of it.
   usage: primary_chunks.py CONFIG"""
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
per_sample = [r.primary_hits(w, h, s, want_rays=True)[1] for s in range(spp)]
rays = np.stack(per_sample)  # [sample][pixel]
n = spp * w * h
i = np.arange(n, dtype=np.int64)
block, inn = i >> 5, i & 31
sin, pin = inn // 4, inn % 4
sgroups = spp // 8
pg, sg = block // sgroups, block % sgroups
tiles_x = w // 2
ty, tx = pg // tiles_x, pg % tiles_x
pix = (ty * 2 + pin // 2) * w + tx * 2 + pin % 2
q = rays[sg * 8 + sin, pix]
del rays, per_sample
whole = r.bench_trace(q, 0, 5)
print(json.dumps({"config": config, "rays": int(n), "whole_ms": round(whole, 3), "Grays/s": round(n / whole / 1e6, 3)}), flush=True)
rng = np.random.default_rng(0)
# the same rays with the WARPS of the queue shuffled (32-ray groups keep their coherence, the image-space order goes)
perm = rng.permutation(n // 32)
qs = q.reshape(-1, 32)[perm].reshape(-1)
ms = r.bench_trace(qs, 0, 5)
print(json.dumps({"warp_groups_shuffled_ms": round(ms, 3), "Grays/s": round(n / ms / 1e6, 3)}), flush=True)
del qs
C = 64
per = n // C
out = []
for c in range(C):
    ms = r.bench_trace(q[c * per:(c + 1) * per], 0, 3)
    out.append(round(ms, 3))
print(json.dumps({"chunks": C, "rays_per_chunk": int(per), "chunk_ms": out, "sum_ms": round(sum(out), 3)}), flush=True)
