"""Host-side `Scene` (raytracer.odin:51-60) as numpy arrays, plus `finish_scene`
(raytracer.odin:62-91) and the conversion to the C structs that cross the boundary."""
import ctypes as C
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np

from . import cabi


@dataclass
class Scene:
    # Cam (raytracer.odin:45-49): basis[:, c] is column c of the Odin matrix[3,3]f32
    cam_pos: np.ndarray = field(default_factory=lambda: np.zeros(3, np.float32))
    cam_basis: np.ndarray = field(default_factory=lambda: np.eye(3, dtype=np.float32))
    fov_x: float = 0.0
    # scene.trigs[1:] — the dummy triangle 0 (input.odin:43) is never stored
    triangles: np.ndarray = field(default_factory=lambda: np.zeros(0, cabi.TRI_DTYPE))
    # scene.materials INCLUDING the dummy material 0 (input.odin:44)
    materials: np.ndarray = field(default_factory=lambda: np.array(
        [((0, 0, 0), -1, (0, 0, 0), -1, 0.0, 0.0, -1, -1)], cabi.MAT_DTYPE))
    textures: List[np.ndarray] = field(default_factory=list)  # HxWxC uint8 or float32, C in 1..4
    env_map: Optional[np.ndarray] = None
    # filled by finish()
    bvh: Optional[np.ndarray] = None
    light_triangles: Optional[np.ndarray] = None
    light_bvh: Optional[np.ndarray] = None

    def apply_render_config(self, width: int, height: int):
        """main.odin:199-204: when --height is given, fov_x (= glTF yfov, input.odin:108) is
        multiplied by the aspect ratio.  Returns a new fov_x; idempotence is the caller's job."""
        aspect = np.float32(width) / np.float32(height)
        return float(np.float32(self.fov_x) * aspect)

    def finish(self, bvh_build: Callable[[np.ndarray], np.ndarray]):
        """finish_scene (raytracer.odin:62-91): collect emissive triangles BEFORE the scene BVH
        build reorders scene.trigs (:63-66), then build both BVHs (each sorts its triangle array
        in place, :72,75)."""
        m = self.materials[self.triangles["material_index"]]
        l1 = np.abs(m["emission_factor"]).sum(axis=1, dtype=np.float32)  # norm_l1 (utils.odin:10)
        self.light_triangles = self.triangles[l1 > np.float32(1e-6)].copy()
        self.triangles = np.ascontiguousarray(self.triangles)
        self.bvh = bvh_build(self.triangles)
        self.light_bvh = bvh_build(self.light_triangles)
        return self

    @property
    def finished(self):
        return self.bvh is not None

    def to_c(self):
        """Build the ort_scene view. Returns (OrtScene, keepalive)."""
        if not self.finished:
            raise RuntimeError("Scene.finish() must run before the scene crosses the boundary")
        keep = []

        def tex_struct(img):
            img = np.ascontiguousarray(img)
            if img.ndim == 2:
                img = img[:, :, None]
            assert img.dtype in (np.uint8, np.float32) and 1 <= img.shape[2] <= 4
            keep.append(img)
            t = cabi.OrtTexture()
            t.data = img.ctypes.data
            t.width, t.height, t.channels = img.shape[1], img.shape[0], img.shape[2]
            t.is_f32 = 1 if img.dtype == np.float32 else 0
            t.stride = img.shape[1] * img.shape[2]
            return t

        s = cabi.OrtScene()
        s.cam.pos[:] = [float(x) for x in self.cam_pos]
        s.cam.basis[:] = [float(self.cam_basis[r, c]) for c in range(3) for r in range(3)]
        s.cam.fov_x = float(self.fov_x)
        for name, arr, dt in (
            ("triangles", self.triangles, cabi.TRI_DTYPE),
            ("bvh", self.bvh, cabi.NODE_DTYPE),
            ("light_triangles", self.light_triangles, cabi.TRI_DTYPE),
            ("light_bvh", self.light_bvh, cabi.NODE_DTYPE),
            ("materials", self.materials, cabi.MAT_DTYPE),
        ):
            a = np.ascontiguousarray(arr, dtype=dt)
            keep.append(a)
            setattr(s, name, a.ctypes.data if len(a) else None)
        s.n_triangles = len(self.triangles)
        s.n_bvh_nodes = len(self.bvh)
        s.n_light_triangles = len(self.light_triangles)
        s.n_light_bvh_nodes = len(self.light_bvh)
        s.n_materials = len(self.materials)
        if self.textures:
            arr = (cabi.OrtTexture * len(self.textures))(*[tex_struct(t) for t in self.textures])
            keep.append(arr)
            s.textures = C.cast(arr, C.POINTER(cabi.OrtTexture))
        s.n_textures = len(self.textures)
        if self.env_map is not None:
            e = tex_struct(self.env_map)
            keep.append(e)
            s.env_map = C.pointer(e)
        return s, keep


def native_bvh_build(tris: np.ndarray) -> np.ndarray:
    """bvh_build (raytracer.odin:227-342) through the library's host helper ort_bvh_build.
    Sorts `tris` in place like the reference."""
    lib = cabi.load_library()
    n = len(tris)
    cap = max(2 * n, 1)
    nodes = np.zeros(cap, cabi.NODE_DTYPE)
    cnt = lib.ort_bvh_build(cabi.ptr(tris) if n else None, n, cabi.ptr(nodes), cap)
    if cnt < 0:
        raise RuntimeError(f"ort_bvh_build failed ({cnt})")
    return nodes[:cnt].copy()


def device_bvh_build(tris: np.ndarray, device: int = 0) -> np.ndarray:
    """bvh_build on the GPU (ort_bvh_build_device): same nodes and permutation as native_bvh_build."""
    lib = cabi.load_library()
    n = len(tris)
    cap = max(2 * n, 1)
    nodes = np.zeros(cap, cabi.NODE_DTYPE)
    cnt = lib.ort_bvh_build_device(device, cabi.ptr(tris) if n else None, n, cabi.ptr(nodes), cap)
    if cnt < 0:
        raise RuntimeError(f"ort_bvh_build_device failed ({cnt}): {lib.ort_bvh_build_device_error().decode()}")
    return nodes[:cnt].copy()
