"""Stand-in for `read_gltf` / `load_texture` (input.odin:13-259, textures.odin:25-68).

In production Odin keeps this stage (BASELINE.json north_star); this module exists so the test
harness, the CLI and bench.py can load the same glTF subset without an Odin toolchain.  It
follows populate_scene's flattening rules: JSON .gltf only (input.odin:28), node transform chain
parent*local (:99-101), camera from the node that carries one (:103-109), one Material per
primitive instance (:137-162), positions to world space, tangents normalised (:191-196),
ng = normalize(cross(e1,e2)) (:197), normals through cofactor(mat3(transform)) (:203-206).
"""
import base64
import json
import os
from typing import Dict, List, Optional
from urllib.parse import unquote

import numpy as np

from . import cabi
from .scene import Scene

_COMP = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}

f32 = np.float32


def load_texture(path: str) -> np.ndarray:
    """load_texture (textures.odin:25-68): stb_image semantics — HDR files decode to f32 RGB,
    everything else to u8 with the file's native channel count."""
    import cv2

    with open(path, "rb") as f:
        head = f.read(16)
    is_hdr = head.startswith(b"#?RADIANCE") or head.startswith(b"#?RGBE")
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise RuntimeError("Failed to parse texture")  # textures.odin:56
    if is_hdr:
        return np.ascontiguousarray(img[:, :, ::-1].astype(np.float32))
    if img.dtype == np.uint16:  # stb converts 16-bit to 8-bit
        img = (img >> 8).astype(np.uint8)
    if img.ndim == 2:
        return np.ascontiguousarray(img[:, :, None])
    if img.shape[2] == 3:
        return np.ascontiguousarray(img[:, :, ::-1])
    if img.shape[2] == 4:
        return np.ascontiguousarray(img[:, :, [2, 1, 0, 3]])
    return np.ascontiguousarray(img)


def _node_transform_local(node: dict) -> np.ndarray:
    """cgltf_node_transform_local (called at input.odin:100). Returns a 4x4 with m[r, c]."""
    if "matrix" in node:
        return np.array(node["matrix"], f32).reshape(4, 4).T.copy()  # glTF stores column-major
    tx, ty, tz = [f32(x) for x in node.get("translation", (0, 0, 0))]
    qx, qy, qz, qw = [f32(x) for x in node.get("rotation", (0, 0, 0, 1))]
    sx, sy, sz = [f32(x) for x in node.get("scale", (1, 1, 1))]
    one, two = f32(1), f32(2)
    lm = np.zeros(16, f32)
    lm[0] = (one - two * qy * qy - two * qz * qz) * sx
    lm[1] = (two * qx * qy + two * qz * qw) * sx
    lm[2] = (two * qx * qz - two * qy * qw) * sx
    lm[4] = (two * qx * qy - two * qz * qw) * sy
    lm[5] = (one - two * qx * qx - two * qz * qz) * sy
    lm[6] = (two * qy * qz + two * qx * qw) * sy
    lm[8] = (two * qx * qz + two * qy * qw) * sz
    lm[9] = (two * qy * qz - two * qx * qw) * sz
    lm[10] = (one - two * qx * qx - two * qy * qy) * sz
    lm[12], lm[13], lm[14], lm[15] = tx, ty, tz, one
    return lm.reshape(4, 4).T.copy()  # lm is column-major


def _cofactor3(m: np.ndarray) -> np.ndarray:
    """linalg.cofactor of a 3x3 (input.odin:203)."""
    c = np.zeros((3, 3), f32)
    for i in range(3):
        for j in range(3):
            r = [k for k in range(3) if k != i]
            s = [k for k in range(3) if k != j]
            minor = m[r[0], s[0]] * m[r[1], s[1]] - m[r[0], s[1]] * m[r[1], s[0]]
            c[i, j] = minor if (i + j) % 2 == 0 else -minor
    return c


def _normalize_rows(v: np.ndarray) -> np.ndarray:
    """linalg.normalize per row: v / sqrt((x*x + y*y) + z*z), every f32 operation rounded once."""
    with np.errstate(invalid="ignore", divide="ignore"):
        v = v.astype(f32)
        ln = np.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2]).astype(f32)
        return (v / ln[:, None]).astype(f32)


def _mat4_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """4x4 f32 product, entries summed left to right with no fused multiply-add (numpy's matmul may
    call a BLAS kernel with a different summation order; the C++ host uses exactly this order)."""
    o = np.zeros((4, 4), f32)
    for r in range(4):
        for c in range(4):
            o[r, c] = ((a[r, 0] * b[0, c] + a[r, 1] * b[1, c]) + a[r, 2] * b[2, c]) + a[r, 3] * b[3, c]
    return o


def _transform_rows(m3: np.ndarray, tr: np.ndarray, v: np.ndarray, w: float) -> np.ndarray:
    """(transform * {v, w}).xyz for every row of v: ((m0*x + m1*y) + m2*z) + t*w, elementwise f32."""
    v = v.astype(f32)
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    out = np.empty((len(v), 3), f32)
    for r in range(3):
        out[:, r] = (m3[r, 0] * x + m3[r, 1] * y) + m3[r, 2] * z
        if tr is not None:
            out[:, r] = out[:, r] + tr[r] * f32(w)
    return out


class _Gltf:
    def __init__(self, path: str):
        self.path = path
        self.root = os.path.dirname(os.path.abspath(path))
        with open(path, "r") as f:
            self.j = json.load(f)
        self.buffers: List[bytes] = []
        for b in self.j.get("buffers", []):
            uri = b["uri"]
            if uri.startswith("data:"):
                self.buffers.append(base64.b64decode(uri.split(",", 1)[1]))
            else:
                with open(os.path.join(self.root, unquote(uri)), "rb") as f:
                    self.buffers.append(f.read())
        self._acc: Dict[int, np.ndarray] = {}

    def accessor(self, idx: int) -> np.ndarray:
        if idx in self._acc:
            return self._acc[idx]
        a = self.j["accessors"][idx]
        bv = self.j["bufferViews"][a["bufferView"]]
        dt = np.dtype(_COMP[a["componentType"]])
        nc = _NCOMP[a["type"]]
        off = bv.get("byteOffset", 0) + a.get("byteOffset", 0)
        stride = bv.get("byteStride", 0) or dt.itemsize * nc
        buf = self.buffers[bv["buffer"]]
        arr = np.ndarray((a["count"], nc), dt, buf, off, (stride, dt.itemsize))
        self._acc[idx] = arr
        return arr

    def accessor_float(self, idx: int) -> np.ndarray:
        """cgltf_accessor_read_float semantics (float passthrough, normalised ints scaled)."""
        arr = self.accessor(idx)
        a = self.j["accessors"][idx]
        if arr.dtype == np.float32:
            return arr
        if a.get("normalized"):
            info = np.iinfo(arr.dtype)
            return np.maximum(arr.astype(f32) / f32(info.max), f32(-1))
        return arr.astype(f32)


def read_gltf(gltf_path: str) -> Scene:
    """read_gltf (input.odin:13-259).  Returns an un-finished Scene (call Scene.finish)."""
    g = _Gltf(gltf_path)
    j = g.j
    scene = Scene()
    tris: List[np.ndarray] = []
    mats: List[tuple] = [((0, 0, 0), -1, (0, 0, 0), -1, 0.0, 0.0, -1, -1)]  # dummy Material{} 0: nil samplers
    texture_cache: Dict[str, int] = {}
    textures: List[np.ndarray] = []

    def load_sampler(view: Optional[dict]) -> int:  # load_sampler/load_image input.odin:50-90
        if view is None:
            return -1
        tex = j["textures"][view["index"]]
        img = j["images"][tex["source"]]
        p = os.path.join(g.root, unquote(img["uri"]))
        if p not in texture_cache:
            texture_cache[p] = len(textures)
            textures.append(load_texture(p))
        return texture_cache[p]

    def populate(node_idx: int, parent: np.ndarray):  # populate_scene input.odin:92-233
        node = j["nodes"][node_idx]
        transform = _mat4_mul(parent, _node_transform_local(node))
        if "camera" in node:  # :103-109
            scene.cam_pos = transform[:3, 3].copy()
            basis = np.zeros((3, 3), f32)
            basis[:, 0] = transform[:3, 0]
            basis[:, 1] = transform[:3, 1]
            basis[:, 2] = -transform[:3, 2]
            scene.cam_basis = basis
            scene.fov_x = float(f32(j["cameras"][node["camera"]]["perspective"]["yfov"]))
        if "mesh" in node:
            for prim in j["meshes"][node["mesh"]]["primitives"]:
                attrs = prim["attributes"]
                if "POSITION" not in attrs:
                    raise RuntimeError("No position accessor found in mesh primitive")
                gm = j["materials"][prim["material"]]  # nil material derefs in the reference (:138)
                pbr = gm.get("pbrMetallicRoughness", {})
                color = pbr.get("baseColorFactor", [1, 1, 1, 1])
                emission = np.array(gm.get("emissiveFactor", [0, 0, 0]), f32)
                ext = gm.get("extensions", {}).get("KHR_materials_emissive_strength")
                ct = load_sampler(pbr.get("baseColorTexture"))
                et = load_sampler(gm.get("emissiveTexture"))
                mrt = load_sampler(pbr.get("metallicRoughnessTexture"))
                nt = load_sampler(gm.get("normalTexture"))
                if ext is not None:  # :157-159
                    emission = emission * f32(ext.get("emissiveStrength", 1.0))
                material_index = len(mats)
                mats.append((
                    tuple(f32(c) for c in color[:3]), ct, tuple(emission), et,
                    f32(pbr.get("metallicFactor", 1.0)), f32(pbr.get("roughnessFactor", 1.0)), mrt, nt,
                ))

                pos = g.accessor_float(attrs["POSITION"])
                if "indices" in prim:
                    idx = g.accessor(prim["indices"])[:, 0].astype(np.int64)
                else:
                    idx = np.arange(len(pos), dtype=np.int64)
                nt_ = len(idx) // 3
                idx = idx[: nt_ * 3].reshape(nt_, 3)

                m3 = transform[:3, :3]
                tr = transform[:3, 3]
                wp = _transform_rows(m3, tr, pos[:, :3], 1.0)  # transform * (p,1)
                P = wp[idx]  # nt x 3 x 3
                e1 = (P[:, 1] - P[:, 0]).astype(f32)
                e2 = (P[:, 2] - P[:, 0]).astype(f32)
                ng = _normalize_rows(np.cross(e1, e2).astype(f32))

                T = np.zeros(nt_, cabi.TRI_DTYPE)
                T["p"], T["u"], T["v"], T["ng"] = P[:, 0], e1, e2, ng
                if "NORMAL" in attrs:
                    cof = _cofactor3(m3)
                    wn = _normalize_rows(_transform_rows(cof, None, g.accessor_float(attrs["NORMAL"])[:, :3], 0.0))
                    N = wn[idx]
                    T["n1"], T["n2"], T["n3"] = N[:, 0], N[:, 1], N[:, 2]
                else:
                    T["n1"] = T["n2"] = T["n3"] = ng
                if "TEXCOORD_0" in attrs:
                    uv = g.accessor_float(attrs["TEXCOORD_0"])[:, :2].astype(f32)[idx]
                    T["tex1"], T["tex2"], T["tex3"] = uv[:, 0], uv[:, 1], uv[:, 2]
                if "TANGENT" in attrs:
                    tg = g.accessor_float(attrs["TANGENT"]).astype(f32)
                    wt = _normalize_rows(_transform_rows(m3, tr, tg[:, :3], 0.0))
                    tg4 = np.concatenate([wt, tg[:, 3:4]], axis=1)[idx]
                else:
                    # the reference normalises a zero tangent -> NaN xyz, w = 0 (:193-195)
                    tg4 = np.zeros((nt_, 3, 4), f32)
                    tg4[:, :, :3] = np.nan
                T["tan1"], T["tan2"], T["tan3"] = tg4[:, 0], tg4[:, 1], tg4[:, 2]
                T["material_index"] = material_index
                tris.append(T)
        for child in node.get("children", []):
            populate(child, transform)

    ident = np.eye(4, dtype=f32)
    if "scene" in j:  # input.odin:236-248
        roots = j["scenes"][j["scene"]]["nodes"]
    elif j.get("scenes"):
        roots = j["scenes"][0]["nodes"]
    else:
        roots = list(range(len(j.get("nodes", []))))
    for r in roots:
        populate(r, ident)

    scene.triangles = np.concatenate(tris) if tris else np.zeros(0, cabi.TRI_DTYPE)
    scene.materials = np.array(mats, dtype=cabi.MAT_DTYPE)
    scene.textures = textures
    return scene
