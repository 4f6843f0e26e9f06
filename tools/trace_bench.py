"""Traversal micro-benchmark: realistic ray sets resident on the device, event-timed launches.
   usage: trace_bench.py CONFIG [lib.so ...]   (each lib runs in its own process via ORT_LIB)
Ray sets: primary (coherent), bounce1 (cosine-distributed directions from the primary hit points,
in queue = pixel order), bounce1 shuffled (fully incoherent)."""
import os, sys, json, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def worker(config):
    import numpy as np
    import bench
    from raytracer_odin_b200 import api, cabi
    from raytracer_odin_b200.scene import native_bvh_build
    scene, cfg = bench.build_scene(config, native_bvh_build)
    w, h = cfg["width"], cfg["height"]
    r = api.Renderer(seed=1).upload_scene(scene)
    hits, rays = r.primary_hits(w, h, 0, want_rays=True)
    ok = hits["tri"] >= 0
    tri = scene.triangles[hits["tri"][ok]]
    u, v = hits["u"][ok][:, None], hits["v"][ok][:, None]
    P = tri["p"] + tri["u"] * u + tri["v"] * v
    N = tri["n1"] * (1 - u - v) + tri["n2"] * u + tri["n3"] * v
    N /= np.linalg.norm(N, axis=1, keepdims=True)
    N[(N * rays["d"][ok]).sum(1) > 0] *= -1
    rng = np.random.default_rng(0)
    REP = 4  # four cosine-distributed directions per hit point, samples adjacent like in the render's queue
    P = np.repeat(P, REP, axis=0); N = np.repeat(N, REP, axis=0)
    s = rng.normal(size=P.shape); s /= np.linalg.norm(s, axis=1, keepdims=True)
    D = s + N; D /= np.linalg.norm(D, axis=1, keepdims=True)
    b1 = np.zeros(len(P), cabi.RAY_DTYPE); b1["o"] = P.astype(np.float32); b1["d"] = D.astype(np.float32)
    out = {"config": config, "lib": os.environ.get("ORT_LIB", "default"), "n_primary": len(rays), "n_bounce1": len(b1)}
    def binned(rs, seg, nbits):
        """reorder inside consecutive segments of `seg` rays by a direction-bin key (stable)"""
        d = rs["d"]
        key = (d[:, 0] < 0) * 1 + (d[:, 1] < 0) * 2 + (d[:, 2] < 0) * 4
        if nbits > 3:
            ad = np.abs(d)
            key = key * 4 + (ad[:, 0] > ad[:, 1]) * 2 + (ad[:, 2] > np.maximum(ad[:, 0], ad[:, 1])) * 1
        segid = np.arange(len(rs)) // seg
        order = np.lexsort((key, segid))
        return rs[order]
    sets = [("primary", rays), ("bounce1", b1), ("bounce1_shuffled", b1[rng.permutation(len(b1))])]
    if os.environ.get("ORT_BENCH_ORDERS"):
        sets += [("b1_oct256", binned(b1, 256, 3)), ("b1_oct4096", binned(b1, 4096, 3)), ("b1_dir32_1024", binned(b1, 1024, 5)),
                 ("b1_dir32_4096", binned(b1, 4096, 5)), ("b1_dir32_65536", binned(b1, 65536, 5)), ("b1_dir32_global", binned(b1, 1 << 30, 5))]
    for name, rs in sets:
        ms = r.bench_trace(rs, 0, 10)
        out[name + "_ms"] = round(ms, 3); out[name + "_Grays/s"] = round(len(rs) / ms / 1e6, 3)
    if len(scene.light_triangles):
        ms = r.bench_trace(b1, 1, 10)
        out["light_bounce1_ms"] = round(ms, 3)
    print(json.dumps(out), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--worker":
        worker(sys.argv[2])
    else:
        config = sys.argv[1] if len(sys.argv) > 1 else "C2"
        libs = sys.argv[2:] or [""]
        for lib in libs:
            env = dict(os.environ)
            if lib: env["ORT_LIB"] = os.path.abspath(lib)
            subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", config], env=env)
