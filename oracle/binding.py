"""ctypes binding of the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE ONLY — importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "rays", "node_pops", "box_tests", "tri_tests", "stack_drops", "stack_high", "exact_ties",
        "light_rays", "light_node_pops", "light_box_tests", "light_tri_tests")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_libs = {}


def build(force=False):
    if force or not all(os.path.exists(os.path.join(HERE, n)) for n in ("liboracle.so", "liboracle_native.so",
                                                                         "liboracle_alt.so")):
        subprocess.check_call(["make", "-C", HERE] + (["-B"] if force else []), stdout=subprocess.DEVNULL)


def load(native=False):
    """native=False: -O2 build used as the CHECKER; native=True: -O3 -march=native build used as
    the TIMED CPU baseline (the reference's `brrr` flags analogue, justfile:37-41).  Both are
    -ffp-contract=off.  native="alt": the ORC_VARIANTS build of the sensitivity study (orc_set_variant)."""
    key = native if native == "alt" else bool(native)
    if key in _libs:
        return _libs[key]
    name = "liboracle_alt.so" if native == "alt" else ("liboracle_native.so" if native else "liboracle.so")
    path = os.path.join(HERE, name)
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    vp, f, u32, u64, i32, i64 = C.c_void_p, C.c_float, C.c_uint32, C.c_uint64, C.c_int, C.c_int64
    pf = C.POINTER(C.c_float)
    sig = {
        "orc_philox": (None, [u32, u64, u32, u64, C.POINTER(u32)]),
        "orc_bvh_build": (i64, [vp, i64, vp, i64]),
        "orc_check_intersect_ray_aabb": (i32, [pf, pf, pf, pf, f, pf]),
        "orc_intersect_ray_triangle": (None, [pf, pf, vp, pf]),
        "orc_trace_rays": (None, [vp, vp, i64, i32, vp, C.POINTER(Counters), i32, vp]),
        "orc_light_pdf": (None, [vp, vp, i64, vp]),
        "orc_primary_hits": (None, [vp, u32, u32, u64, u64, i32, vp, vp, C.POINTER(Counters), i32, vp]),
        "orc_render": (None, [vp, u32, u32, C.c_int32, u64, u64, u64, i32, i32, i32, u32, u32, u32, u32, vp,
                              C.POINTER(Counters)]),
        "orc_render_strided": (None, [vp, u32, u32, C.c_int32, u64, u64, u64, i32, i32, i32, u32, u32, u32, u32, u32, vp,
                                      C.POINTER(Counters)]),
        "orc_shade": (None, [pf, pf, f, f, pf, pf, pf]),
        "orc_cosine_weighted_pdf": (f, [pf, pf]),
        "orc_vndf_sampling_pdf": (f, [pf, pf, f, pf]),
        "orc_vndf_sampling": (None, [pf, pf, f, f, f, pf]),
        "orc_cosine_weighted": (None, [pf, u32, u32, pf]),
        "orc_sample": (None, [vp, pf, pf, f, pf, C.POINTER(u32), pf]),
        "orc_pdf": (f, [vp, pf, pf, f, pf, pf]),
        "orc_texture_sample": (None, [vp, f, f, i32, pf, pf]),
        "orc_env_lookup": (None, [vp, pf, pf]),
        "orc_pixel_to_ray_dir": (None, [vp, u32, u32, pf]),
        "orc_get_rgb_image": (None, [vp, u32, u32, vp]),
        "orc_hardware_threads": (i32, []),
    }
    if key == "alt":
        sig["orc_set_variant"] = (None, [i32])
    for n, (res, args) in sig.items():
        fn = getattr(lib, n)
        fn.restype, fn.argtypes = res, args
    _libs[key] = lib
    return lib


def _fa(a):
    a = np.ascontiguousarray(a, np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def bvh_build(tris: np.ndarray, native=False) -> np.ndarray:
    """bvh_build (raytracer.odin:227-342) — sorts `tris` in place, returns the node array."""
    from raytracer_odin_b200 import cabi

    lib = load(native)
    n = len(tris)
    cap = max(2 * n, 1)
    nodes = np.zeros(cap, cabi.NODE_DTYPE)
    cnt = lib.orc_bvh_build(tris.ctypes.data if n else None, n, nodes.ctypes.data, cap)
    assert cnt > 0
    return nodes[:cnt].copy()


class OracleScene:
    """Holds the ort_scene view of a finished Scene for oracle calls."""

    def __init__(self, scene, native=False):
        self.scene = scene
        self.cs, self._keep = scene.to_c()
        self.lib = load(native)
        self.ref = C.byref(self.cs)

    def trace_rays(self, rays, mode=0, threads=8):
        from raytracer_odin_b200 import cabi

        rays = np.ascontiguousarray(rays, cabi.RAY_DTYPE)
        out = np.zeros(len(rays), cabi.HIT_DTYPE)
        c = Counters()
        self.ties = np.zeros(len(rays), np.uint8)  # 1 = result depends on visiting order (exact-t tie)
        self.lib.orc_trace_rays(self.ref, rays.ctypes.data, len(rays), mode, out.ctypes.data, C.byref(c), threads,
                                self.ties.ctypes.data)
        return out, c.as_dict()

    def light_pdf(self, rays):
        from raytracer_odin_b200 import cabi

        rays = np.ascontiguousarray(rays, cabi.RAY_DTYPE)
        out = np.zeros(len(rays), np.float32)
        self.lib.orc_light_pdf(self.ref, rays.ctypes.data, len(rays), out.ctypes.data)
        return out

    def primary_hits(self, w, h, sample=0, seed=0, mode=0, threads=8):
        from raytracer_odin_b200 import cabi

        out = np.zeros(w * h, cabi.HIT_DTYPE)
        rays = np.zeros(w * h, cabi.RAY_DTYPE)
        c = Counters()
        self.ties = np.zeros(w * h, np.uint8)
        self.lib.orc_primary_hits(self.ref, w, h, sample, seed, mode, out.ctypes.data, rays.ctypes.data,
                                  C.byref(c), threads, self.ties.ctypes.data)
        return out, rays, c.as_dict()

    def render(self, w, h, ray_depth, n_samples, first_sample=0, seed=0, mode=0, schedule=1, threads=8,
               window=None, out=None, tile_stride=1):
        """Returns (Sample_Stats array [h*w], counters).  tile_stride > 1: only every tile_stride-th 4x4 tile
        in x and y (a bounded sample spread over the whole frame)."""
        from raytracer_odin_b200 import cabi

        if out is None:
            out = np.zeros(w * h, cabi.STATS_DTYPE)
        x0, y0, x1, y1 = window if window else (0, 0, w, h)
        c = Counters()
        self.lib.orc_render_strided(self.ref, w, h, ray_depth, first_sample, n_samples, seed, mode, schedule, threads,
                                    x0, y0, x1, y1, tile_stride, out.ctypes.data, C.byref(c))
        return out, c.as_dict()


def get_rgb_image(stats, w, h):
    lib = load()
    out = np.zeros(w * h * 3, np.uint8)
    stats = np.ascontiguousarray(stats)
    lib.orc_get_rgb_image(stats.ctypes.data, w, h, out.ctypes.data)
    return out.reshape(h, w, 3)
