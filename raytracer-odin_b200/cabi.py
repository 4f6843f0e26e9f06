"""ctypes mirrors of include/odinrt_b200.h (the C ABI the Odin shim binds with `foreign import`)."""
import ctypes as C
import os

import numpy as np

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libodinrt_b200.so")

# numpy mirror of ort_triangle == Odin `Triangle` (raytracer.odin:18-23), 168 bytes.
TRI_DTYPE = np.dtype(
    [
        ("p", "<f4", 3), ("u", "<f4", 3), ("v", "<f4", 3),
        ("n1", "<f4", 3), ("n2", "<f4", 3), ("n3", "<f4", 3), ("ng", "<f4", 3),
        ("tex1", "<f4", 2), ("tex2", "<f4", 2), ("tex3", "<f4", 2),
        ("tan1", "<f4", 4), ("tan2", "<f4", 4), ("tan3", "<f4", 4),
        ("material_index", "<i8"),
    ],
    align=True,
)
assert TRI_DTYPE.itemsize == 168 and TRI_DTYPE.fields["material_index"][1] == 160

# ort_bvh_node (BVH_Node, raytracer.odin:211-225), 48 bytes.
NODE_DTYPE = np.dtype(
    [("lo", "<f4", 3), ("hi", "<f4", 3), ("kind", "<i4"), ("_pad", "<i4"), ("a", "<i8"), ("b", "<i8")],
    align=True,
)
assert NODE_DTYPE.itemsize == 48

# ort_material (Material, raytracer.odin:34-43)
MAT_DTYPE = np.dtype(
    [
        ("color_factor", "<f4", 3), ("color_texture", "<i4"),
        ("emission_factor", "<f4", 3), ("emission_texture", "<i4"),
        ("metallic_factor", "<f4"), ("roughness_factor", "<f4"),
        ("metallic_roughness_texture", "<i4"), ("normal_texture", "<i4"),
    ],
    align=True,
)
assert MAT_DTYPE.itemsize == 48

# ort_sample_stats (Sample_Stats, main.odin:34-40), 52 bytes.
STATS_DTYPE = np.dtype(
    [("first", "<f4", 3), ("count", "<u4"), ("last", "<f4", 3), ("total", "<f4", 3), ("total_squared", "<f4", 3)]
)
assert STATS_DTYPE.itemsize == 52

RAY_DTYPE = np.dtype([("o", "<f4", 3), ("d", "<f4", 3)])
HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("tri", "<i4"), ("material", "<i4"), ("inside", "<i4")])
assert RAY_DTYPE.itemsize == 24 and HIT_DTYPE.itemsize == 24


class OrtTexture(C.Structure):
    _fields_ = [
        ("data", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32),
        ("channels", C.c_int32), ("is_f32", C.c_int32), ("stride", C.c_int64),
    ]


class OrtCamera(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("basis", C.c_float * 9), ("fov_x", C.c_float)]


class OrtScene(C.Structure):
    _fields_ = [
        ("cam", OrtCamera),
        ("triangles", C.c_void_p), ("n_triangles", C.c_int64),
        ("bvh", C.c_void_p), ("n_bvh_nodes", C.c_int64),
        ("light_triangles", C.c_void_p), ("n_light_triangles", C.c_int64),
        ("light_bvh", C.c_void_p), ("n_light_bvh_nodes", C.c_int64),
        ("materials", C.c_void_p), ("n_materials", C.c_int64),
        ("textures", C.POINTER(OrtTexture)), ("n_textures", C.c_int64),
        ("env_map", C.POINTER(OrtTexture)),
    ]


class OrtDeviceCfg(C.Structure):
    _fields_ = [("device", C.c_int32), ("_pad", C.c_int32), ("seed", C.c_uint64), ("max_paths_in_flight", C.c_int64),
                ("max_path_bytes", C.c_int64)]


class OrtStats(C.Structure):
    _fields_ = [
        ("rays_closest", C.c_uint64), ("rays_traced", C.c_uint64), ("rays_light_pdf", C.c_uint64), ("paths", C.c_uint64),
        ("kernel_launches", C.c_uint64),
        ("render_ms", C.c_double), ("trace_ms", C.c_double), ("light_ms", C.c_double),
        ("shade_ms", C.c_double), ("other_ms", C.c_double),
        ("wide_nodes", C.c_int64), ("wide_depth", C.c_int64), ("light_wide_nodes", C.c_int64),
        ("device_bytes", C.c_int64), ("wide_max_stack", C.c_int64), ("reference_stack_need", C.c_int64),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# name -> (restype, argtypes); every symbol include/odinrt_b200.h declares.
ABI = {
    "ort_abi_version": (C.c_int, []),
    "ort_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(OrtDeviceCfg)]),
    "ort_destroy": (None, [C.c_void_p]),
    "ort_last_error": (C.c_char_p, [C.c_void_p]),
    "ort_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ort_upload_scene": (C.c_int, [C.c_void_p, C.POINTER(OrtScene)]),
    "ort_render": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int32, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "ort_render_device": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int32, C.c_uint64, C.c_uint64, C.c_void_p]),
    "ort_unpack_accum": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "ort_trace_rays": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ort_primary_hits": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p]),
    "ort_light_pdf": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ort_tonemap_rgb8": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "ort_bench_trace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_double)]),
    "ort_wide_bvh_emit": (C.c_int64, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int64,
                                      C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "ort_bench_read_bw": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.c_double)]),
    "ort_get_stats": (C.c_int, [C.c_void_p, C.POINTER(OrtStats)]),
    "ort_reset_stats": (C.c_int, [C.c_void_p]),
    "ort_set_profiling": (C.c_int, [C.c_void_p, C.c_int32]),
    "ort_bvh_build": (C.c_int64, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
    "ort_bvh_build_device": (C.c_int64, [C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
    "ort_bvh_build_device_error": (C.c_char_p, []),
    "ort_multi_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_int32, C.c_uint64]),
    "ort_multi_destroy": (None, [C.c_void_p]),
    "ort_multi_last_error": (C.c_char_p, [C.c_void_p]),
    "ort_multi_upload_scene": (C.c_int, [C.c_void_p, C.POINTER(OrtScene)]),
    "ort_multi_render": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int32, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "ort_multi_get_stats": (C.c_int, [C.c_void_p, C.POINTER(OrtStats)]),
    "ort_last_render_samples": (C.c_uint64, [C.c_void_p]),
    "ort_frame_begin": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "ort_frame_load": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ort_frame_render": (C.c_int, [C.c_void_p, C.c_int32, C.c_uint64, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64)]),
    "ort_frame_wait": (C.c_int, [C.c_void_p]),
    "ort_frame_snapshot": (C.c_int, [C.c_void_p]),
    "ort_frame_preview_rgb8": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ort_frame_fetch": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ort_frame_end": (C.c_int, [C.c_void_p]),
    "ort_probe_shading": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "ort_multi_last_render_samples": (C.c_uint64, [C.c_void_p]),
    "ort_multi_frame_begin": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "ort_multi_frame_load": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ort_multi_frame_render": (C.c_int, [C.c_void_p, C.c_int32, C.c_uint64, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64)]),
    "ort_multi_frame_wait": (C.c_int, [C.c_void_p]),
    "ort_multi_frame_snapshot": (C.c_int, [C.c_void_p]),
    "ort_multi_frame_preview_rgb8": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ort_multi_frame_fetch": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ort_multi_frame_end": (C.c_int, [C.c_void_p]),
}

# ort_probe_shading: kind -> (floats in, floats out) per record (include/odinrt_b200.h)
PROBE = {"shade": (0, 14, 3), "vndf_sample": (1, 9, 3), "vndf_pdf": (2, 10, 1), "sample": (3, 14, 3), "pdf": (4, 13, 1),
         "texture": (5, 4, 4), "cosine": (6, 5, 4), "env": (7, 3, 3)}

_lib = None


def load_library(path=None):
    """dlopen libodinrt_b200.so and type every entry point.  Fails loudly if the library was not
    built (python __graft_entry__.py build / make -C raytracer-odin_b200/csrc)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("ORT_LIB") or LIB_PATH  # ORT_LIB: a tuning variant built by `make variant`
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} is missing: the CUDA library is the product and there is no CPU fallback. "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'`."
        )
    lib = C.CDLL(p)
    for name, (res, args) in ABI.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.ort_abi_version() != 2:
        raise RuntimeError("libodinrt_b200.so ABI version mismatch")
    if path is None:
        _lib = lib
    return lib


def ptr(a):
    """void* of a numpy array (must stay alive while C uses it)."""
    return a.ctypes.data_as(C.c_void_p)
