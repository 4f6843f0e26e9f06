"""ctypes view of the C++ host (raytracer-odin_b200/host): the native `read_gltf` / `finish_scene` /
`save_result` that the `odinrt` command line uses, exposed so tests can compare it with the Python
stand-ins (gltf.py, output.py) and so Python callers can load big scenes at native speed."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import cabi
from .scene import Scene

HOST_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host")
HOST_LIB = os.path.join(HOST_DIR, "libodinrt_host.so")
CLI_PATH = os.path.join(HOST_DIR, "odinrt")

_lib = None


def build():
    cabi.load_library()  # the host links against libodinrt_b200.so
    subprocess.check_call(["make", "-C", HOST_DIR], stdout=subprocess.DEVNULL)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_LIB) or not os.path.exists(CLI_PATH):
            build()
        lib = C.CDLL(HOST_LIB)
        lib.orh_scene_load.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p), C.c_char_p, C.c_int]
        lib.orh_scene_free.argtypes = [C.c_void_p]
        lib.orh_scene_free.restype = None
        lib.orh_scene_finish.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        lib.orh_scene_view.argtypes = [C.c_void_p, C.POINTER(cabi.OrtScene)]
        lib.orh_scene_set_fov_x.argtypes = [C.c_void_p, C.c_float]
        lib.orh_scene_set_fov_x.restype = None
        lib.orh_get_rgb_image.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.orh_get_rgb_image.restype = None
        lib.orh_save_result.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_int]
        _lib = lib
    return _lib


def _np(ptr, n, dtype):
    if not ptr or n == 0:
        return np.zeros(0, dtype)
    buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr if isinstance(ptr, int) else C.cast(ptr, C.c_void_p).value)
    return np.frombuffer(buf, dtype=dtype, count=n).copy()


def _tex(t: cabi.OrtTexture) -> np.ndarray:
    dt = np.dtype(np.float32 if t.is_f32 else np.uint8)
    n = t.width * t.height * t.channels
    return _np(t.data, n, dt).reshape(t.height, t.width, t.channels)


def read_gltf(path: str, env_map: str = "", finish: bool = False, bvh_device: int = -1, fov_x=None) -> Scene:
    """read_gltf (+ --env-map, + finish_scene) through the C++ host, returned as a numpy `Scene`."""
    lib = load()
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    if lib.orh_scene_load(path.encode(), env_map.encode() if env_map else None, C.byref(h), err, 512) != 0:
        raise RuntimeError(err.value.decode())
    try:
        if fov_x is not None:
            lib.orh_scene_set_fov_x(h, fov_x)
        if finish and lib.orh_scene_finish(h, bvh_device, err, 512) != 0:
            raise RuntimeError(err.value.decode())
        v = cabi.OrtScene()
        lib.orh_scene_view(h, C.byref(v))
        s = Scene()
        s.cam_pos = np.array(list(v.cam.pos), np.float32)
        s.cam_basis = np.array(list(v.cam.basis), np.float32).reshape(3, 3).T.copy()
        s.fov_x = float(v.cam.fov_x)
        s.triangles = _np(v.triangles, v.n_triangles, cabi.TRI_DTYPE)
        s.materials = _np(v.materials, v.n_materials, cabi.MAT_DTYPE)
        s.textures = [_tex(v.textures[i]) for i in range(v.n_textures)]
        if v.env_map:
            s.env_map = _tex(v.env_map.contents)
        if finish:
            s.bvh = _np(v.bvh, v.n_bvh_nodes, cabi.NODE_DTYPE)
            s.light_triangles = _np(v.light_triangles, v.n_light_triangles, cabi.TRI_DTYPE)
            s.light_bvh = _np(v.light_bvh, v.n_light_bvh_nodes, cabi.NODE_DTYPE)
        return s
    finally:
        lib.orh_scene_free(h)


def get_rgb_image(stats: np.ndarray, width: int, height: int) -> np.ndarray:
    lib = load()
    stats = np.ascontiguousarray(stats, cabi.STATS_DTYPE)
    out = np.zeros((height, width, 3), np.uint8)
    lib.orh_get_rgb_image(stats.ctypes.data, width, height, out.ctypes.data)
    return out


def save_result(stats: np.ndarray, width: int, height: int, path: str):
    lib = load()
    stats = np.ascontiguousarray(stats, cabi.STATS_DTYPE)
    err = C.create_string_buffer(512)
    if lib.orh_save_result(stats.ctypes.data, width, height, path.encode(), err, 512) != 0:
        raise RuntimeError(err.value.decode())
