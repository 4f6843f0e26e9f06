"""Dump a finished scene (triangles, binary BVH) and two ray sets (primary, cosine-distributed bounce rays) for
tools/collapse_study.cpp:  python tools/collapse_study_dump.py C2 /tmp/study/C2 ;  g++ -O2 -std=c++17
tools/collapse_study.cpp -o /tmp/study/study ;  /tmp/study/study /tmp/study/C2"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from raytracer_odin_b200 import cabi
from raytracer_odin_b200.scene import native_bvh_build
from oracle import binding as orc

cfgname, out = sys.argv[1], sys.argv[2]
scene, cfg = bench.build_scene(cfgname, native_bvh_build)
w, h = 480, 270
o = orc.OracleScene(scene, native=True)
hits, rays, _ = o.primary_hits(w, h, sample=0, seed=1, mode=1, threads=8)
ok = hits["tri"] >= 0
tri = scene.triangles[hits["tri"][ok]]
u, v = hits["u"][ok][:, None], hits["v"][ok][:, None]
P = tri["p"] + tri["u"] * u + tri["v"] * v
N = tri["n1"] * (1 - u - v) + tri["n2"] * u + tri["n3"] * v
N /= np.linalg.norm(N, axis=1, keepdims=True)
N[(N * rays["d"][ok]).sum(1) > 0] *= -1
rng = np.random.default_rng(0)
s = rng.normal(size=P.shape); s /= np.linalg.norm(s, axis=1, keepdims=True)
D = s + N; D /= np.linalg.norm(D, axis=1, keepdims=True)
b1 = np.zeros(len(P), cabi.RAY_DTYPE); b1["o"] = (P + 1e-3 * D).astype(np.float32); b1["d"] = D.astype(np.float32)
os.makedirs(os.path.dirname(out), exist_ok=True)
scene.triangles.tofile(out + "_tris.bin"); scene.bvh.tofile(out + "_bvh.bin")
rays.tofile(out + "_primary.bin"); b1.tofile(out + "_bounce.bin")
print(cfgname, len(scene.triangles), len(scene.bvh), len(rays), len(b1))
