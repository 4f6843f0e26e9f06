"""Regenerates tests/golden/*.npz from the CPU oracle.

The reference ships no golden vectors, known-answer tests or fixtures and cannot be
run here (Odin toolchain absent), so these fixtures PIN THE ORACLE against regressions; they are
not outputs of the reference.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import binding as orc  # noqa: E402
from raytracer_odin_b200 import cabi, gltf, scenegen  # noqa: E402

SEED = 99


def soup(n, seed):
    rng = np.random.default_rng(seed)
    t = np.zeros(n, cabi.TRI_DTYPE)
    c = rng.uniform(-1, 1, (n, 3))
    t["p"] = c.astype(np.float32)
    t["u"] = rng.normal(scale=0.15, size=(n, 3)).astype(np.float32)
    t["v"] = rng.normal(scale=0.15, size=(n, 3)).astype(np.float32)
    ng = np.cross(t["u"], t["v"])
    t["ng"] = (ng / np.linalg.norm(ng, axis=1, keepdims=True)).astype(np.float32)
    t["n1"] = t["n2"] = t["n3"] = t["ng"]
    t["material_index"] = 1
    return t


def main():
    # (1) BVH build KAT on 64 triangles
    t = soup(64, 1)
    nodes = orc.bvh_build(t)
    np.savez_compressed(os.path.join(HERE, "bvh_kat.npz"), tris_in=soup(64, 1), tris_out=t, nodes=nodes)
    # (2) Cornell: primary hits + 8 spp image at 64x64
    with tempfile.TemporaryDirectory() as d:
        s = gltf.read_gltf(scenegen.cornell(os.path.join(d, "c1.gltf")))
    w = h = 64
    s.fov_x = s.apply_render_config(w, h)
    s.finish(orc.bvh_build)
    o = orc.OracleScene(s)
    hits, rays, _ = o.primary_hits(w, h, sample=3, seed=SEED, mode=0)
    px, c = o.render(w, h, 6, 8, seed=SEED, mode=0, schedule=1, threads=1)
    np.savez_compressed(os.path.join(HERE, "cornell_64.npz"), hits=hits, rays=rays, total=px["total"],
                        total_squared=px["total_squared"], count=px["count"], first=px["first"], last=px["last"],
                        n_rays=np.array([c["rays"]]))
    # (3) traversal KAT on a 5k soup: 20k rays
    t = soup(5000, 2)
    sc = gltf.Scene()
    sc.triangles = t
    sc.materials = np.array([((0, 0, 0), -1, (0, 0, 0), -1, 0, 0, -1, -1), ((0.8, 0.8, 0.8), -1, (0, 0, 0), -1, 0, 1, -1, -1)],
                            cabi.MAT_DTYPE)
    sc.finish(orc.bvh_build)
    rng = np.random.default_rng(3)
    rays = np.zeros(20000, cabi.RAY_DTYPE)
    rays["o"] = rng.uniform(-1.2, 1.2, (20000, 3)).astype(np.float32)
    dd = rng.normal(size=(20000, 3))
    rays["d"] = (dd / np.linalg.norm(dd, axis=1, keepdims=True)).astype(np.float32)
    hits, _ = orc.OracleScene(sc).trace_rays(rays, mode=0)
    np.savez_compressed(os.path.join(HERE, "soup_trace.npz"), rays=rays, hits=hits)
    # (4) BASELINE config C1 as quoted: Cornell box, 256x256, ray-depth 6, 64 spp (schedule 1: per-pixel sums in
    #     sample order whatever the thread count)
    with tempfile.TemporaryDirectory() as d:
        s = gltf.read_gltf(scenegen.cornell(os.path.join(d, "c1.gltf")))
    w = h = 256
    s.fov_x = s.apply_render_config(w, h)
    s.finish(orc.bvh_build)
    px, c = orc.OracleScene(s).render(w, h, 6, 64, seed=SEED, mode=0, schedule=1, threads=8)
    assert np.all(px["count"] == 64)
    np.savez_compressed(os.path.join(HERE, "cornell_c1_256.npz"), total=px["total"], seed=np.array(SEED),
                        n_rays=np.array([c["rays"]]))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
