"""Build time of the three builders on a BASELINE scene: oracle (comparison sorts), host (radix,
task-parallel), device (level-synchronous).  All three must agree byte for byte."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from raytracer_odin_b200 import gltf, scenegen
from raytracer_odin_b200.scene import native_bvh_build, device_bvh_build
from oracle import binding as orc
import tempfile
cfgs = sys.argv[1:] or ["C2", "C4"]
for c in cfgs:
    path, _ = scenegen.generate(c, tempfile.mkdtemp())
    s = gltf.read_gltf(path)
    res = {}
    for name, fn in (("device", device_bvh_build), ("device(2nd)", device_bvh_build), ("host", native_bvh_build)) + ((("oracle", orc.bvh_build),) if len(s.triangles) < 2_000_000 else ()):
        t = s.triangles.copy(); t0 = time.perf_counter(); nodes = fn(t); dt = time.perf_counter() - t0
        res[name] = (dt, nodes.tobytes(), t["p"].tobytes())
        print(f"{c} {len(t)} tris  {name:12s} {dt*1e3:9.1f} ms  nodes {len(nodes)}", flush=True)
    ref = res["host"]
    for k, v in res.items():
        assert v[1] == ref[1] and v[2] == ref[2], f"{k} differs from host"
    print(c, "all builders identical")
