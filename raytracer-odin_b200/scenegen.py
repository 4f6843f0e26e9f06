"""Synthetic glTF scene generators for the BASELINE.json configs.

Everything stays inside the glTF subset input.odin understands: JSON .gltf + external .bin,
TRIANGLES, POSITION/NORMAL/TEXCOORD_0/TANGENT as f32, u32 indices, every primitive has a material,
images by relative URI (PNG / Radiance .hdr), one perspective camera (only yfov is read).
All generators are deterministic in their seed.

  C1 cornell()   36 tris, area light           256x256,   depth 6,  64 spp
  C2 spheres()   ~100k tris instanced spheres  1920x1080, depth 8,  256 spp
  C3 textured()  textured PBR + HDR env map    1920x1080, depth 8,  1024 spp
  C4 terrain()   ~1M tris terrain + spheres    1920x1080, depth 10, 4096 spp
  C5 terrain(grid=1500, n_spheres=4300)  ~10M  3840x2160, depth 12, continuous
"""
import json
import os
from typing import Optional

import numpy as np

f32 = np.float32


class GltfWriter:
    def __init__(self):
        self.bin = bytearray()
        self.j = {
            "asset": {"version": "2.0", "generator": "raytracer-odin_b200.scenegen"},
            "scene": 0, "scenes": [{"nodes": []}], "nodes": [], "meshes": [], "materials": [],
            "accessors": [], "bufferViews": [], "buffers": [], "cameras": [],
        }
        self.ext_used = set()

    def _view(self, data: bytes, target: Optional[int]) -> int:
        while len(self.bin) % 4:
            self.bin.append(0)
        off = len(self.bin)
        self.bin += data
        bv = {"buffer": 0, "byteOffset": off, "byteLength": len(data)}
        if target:
            bv["target"] = target
        self.j["bufferViews"].append(bv)
        return len(self.j["bufferViews"]) - 1

    def accessor(self, arr: np.ndarray, kind: str, index: bool = False) -> int:
        arr = np.ascontiguousarray(arr)
        comp = {np.dtype(np.float32): 5126, np.dtype(np.uint32): 5125, np.dtype(np.uint16): 5123}[arr.dtype]
        bv = self._view(arr.tobytes(), 34963 if index else 34962)
        a = {"bufferView": bv, "componentType": comp, "count": int(arr.shape[0]), "type": kind}
        if kind == "VEC3" and not index:
            a["min"] = [float(x) for x in arr.min(axis=0)]
            a["max"] = [float(x) for x in arr.max(axis=0)]
        self.j["accessors"].append(a)
        return len(self.j["accessors"]) - 1

    def geometry(self, pos, idx, normal=None, uv=None, tangent=None) -> dict:
        attrs = {"POSITION": self.accessor(np.asarray(pos, f32), "VEC3")}
        if normal is not None:
            attrs["NORMAL"] = self.accessor(np.asarray(normal, f32), "VEC3")
        if uv is not None:
            attrs["TEXCOORD_0"] = self.accessor(np.asarray(uv, f32), "VEC2")
        if tangent is not None:
            attrs["TANGENT"] = self.accessor(np.asarray(tangent, f32), "VEC4")
        return {"attributes": attrs, "indices": self.accessor(np.asarray(idx, np.uint32).reshape(-1), "SCALAR", True), "mode": 4}

    def material(self, color=(1, 1, 1), metallic=0.0, roughness=1.0, emissive=(0, 0, 0), strength=None,
                 color_tex=None, mr_tex=None, normal_tex=None, emissive_tex=None) -> int:
        m = {"pbrMetallicRoughness": {"baseColorFactor": [float(c) for c in color] + [1.0],
                                      "metallicFactor": float(metallic), "roughnessFactor": float(roughness)}}
        if any(emissive):
            m["emissiveFactor"] = [float(e) for e in emissive]
        if strength is not None:
            m["extensions"] = {"KHR_materials_emissive_strength": {"emissiveStrength": float(strength)}}
            self.ext_used.add("KHR_materials_emissive_strength")
        if color_tex is not None:
            m["pbrMetallicRoughness"]["baseColorTexture"] = {"index": color_tex}
        if mr_tex is not None:
            m["pbrMetallicRoughness"]["metallicRoughnessTexture"] = {"index": mr_tex}
        if normal_tex is not None:
            m["normalTexture"] = {"index": normal_tex}
        if emissive_tex is not None:
            m["emissiveTexture"] = {"index": emissive_tex}
        self.j["materials"].append(m)
        return len(self.j["materials"]) - 1

    def texture(self, uri: str) -> int:
        self.j.setdefault("images", []).append({"uri": uri})
        self.j.setdefault("textures", []).append({"source": len(self.j["images"]) - 1})
        return len(self.j["textures"]) - 1

    def mesh(self, geom: dict, material: int) -> int:
        p = dict(geom)
        p["material"] = material
        self.j["meshes"].append({"primitives": [p]})
        return len(self.j["meshes"]) - 1

    def node(self, mesh=None, translation=None, rotation=None, scale=None, matrix=None, camera=None,
             children=None, root=True) -> int:
        n = {}
        if mesh is not None:
            n["mesh"] = mesh
        if camera is not None:
            n["camera"] = camera
        if matrix is not None:
            n["matrix"] = [float(x) for x in np.asarray(matrix, f32).T.reshape(-1)]  # column-major
        if translation is not None:
            n["translation"] = [float(x) for x in translation]
        if rotation is not None:
            n["rotation"] = [float(x) for x in rotation]
        if scale is not None:
            n["scale"] = [float(x) for x in scale]
        if children:
            n["children"] = list(children)
        self.j["nodes"].append(n)
        i = len(self.j["nodes"]) - 1
        if root:
            self.j["scenes"][0]["nodes"].append(i)
        return i

    def camera_look_at(self, eye, target, up=(0, 1, 0), yfov=0.69):
        eye, target, up = (np.asarray(v, np.float64) for v in (eye, target, up))
        z = eye - target
        z /= np.linalg.norm(z)  # glTF cameras look down -Z
        x = np.cross(up, z)
        x /= np.linalg.norm(x)
        y = np.cross(z, x)
        m = np.eye(4)
        m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = x, y, z, eye
        self.j["cameras"].append({"type": "perspective", "perspective": {"yfov": float(yfov), "znear": 0.01}})
        return self.node(camera=len(self.j["cameras"]) - 1, matrix=m)

    def save(self, path: str):
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        bin_name = os.path.splitext(os.path.basename(path))[0] + ".bin"
        self.j["buffers"] = [{"uri": bin_name, "byteLength": len(self.bin)}]
        if self.ext_used:
            self.j["extensionsUsed"] = sorted(self.ext_used)
        with open(os.path.join(os.path.dirname(os.path.abspath(path)), bin_name), "wb") as f:
            f.write(bytes(self.bin))
        with open(path, "w") as f:
            json.dump(self.j, f)
        return path


# --- primitives ---------------------------------------------------------------------------------

def quad(p0, p1, p2, p3):
    """Two triangles (p0,p1,p2), (p0,p2,p3); geometric normal = cross(p1-p0, p2-p0)."""
    return np.array([p0, p1, p2, p3], f32), np.array([[0, 1, 2], [0, 2, 3]], np.uint32)


def box(lo, hi):
    """Axis-aligned box, outward-facing, 12 triangles, flat (no NORMAL attribute)."""
    lo, hi = np.asarray(lo, f32), np.asarray(hi, f32)
    c = np.array([[lo[0], lo[1], lo[2]], [hi[0], lo[1], lo[2]], [hi[0], hi[1], lo[2]], [lo[0], hi[1], lo[2]],
                  [lo[0], lo[1], hi[2]], [hi[0], lo[1], hi[2]], [hi[0], hi[1], hi[2]], [lo[0], hi[1], hi[2]]], f32)
    faces = [(0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (2, 3, 7, 6), (1, 2, 6, 5), (0, 4, 7, 3)]
    idx = []
    for a, b, cc, d in faces:
        idx += [[a, b, cc], [a, cc, d]]
    return c, np.array(idx, np.uint32)


def icosphere(subdiv: int):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    v = [np.array(p, np.float64) / np.linalg.norm(p) for p in v]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2),
             (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11),
             (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    for _ in range(subdiv):
        cache, nf = {}, []

        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[k] = len(v) - 1
            return cache[k]

        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = nf
    pos = np.array(v, f32)
    return pos, np.array(faces, np.uint32), pos.copy()  # unit sphere: normal == position


def uv_sphere(nu: int, nv: int, radius=1.0):
    u = np.linspace(0, 1, nu + 1)
    v = np.linspace(0, 1, nv + 1)
    uu, vv = np.meshgrid(u, v, indexing="xy")
    phi, theta = uu * 2 * np.pi, vv * np.pi
    n = np.stack([np.sin(theta) * np.cos(phi), np.cos(theta), np.sin(theta) * np.sin(phi)], -1)
    tg = np.stack([-np.sin(phi), np.zeros_like(phi), np.cos(phi), np.ones_like(phi)], -1)
    pos = n * radius
    idx = []
    for jv in range(nv):
        for iu in range(nu):
            a = jv * (nu + 1) + iu
            b, c, d = a + 1, a + nu + 1, a + nu + 2
            if jv != 0:
                idx.append([a, b, c])
            if jv != nv - 1:
                idx.append([b, d, c])
    return (pos.reshape(-1, 3).astype(f32), np.array(idx, np.uint32), n.reshape(-1, 3).astype(f32),
            np.stack([uu, vv], -1).reshape(-1, 2).astype(f32), tg.reshape(-1, 4).astype(f32))


def torus(nu: int, nv: int, R=1.0, r=0.35):
    u = np.linspace(0, 1, nu + 1)
    v = np.linspace(0, 1, nv + 1)
    uu, vv = np.meshgrid(u, v, indexing="xy")
    a, b = uu * 2 * np.pi, vv * 2 * np.pi
    n = np.stack([np.cos(b) * np.cos(a), np.sin(b), np.cos(b) * np.sin(a)], -1)
    center = np.stack([R * np.cos(a), np.zeros_like(a), R * np.sin(a)], -1)
    pos = center + r * n
    tg = np.stack([-np.sin(a), np.zeros_like(a), np.cos(a), np.ones_like(a)], -1)
    idx = []
    for jv in range(nv):
        for iu in range(nu):
            p = jv * (nu + 1) + iu
            q, s, t = p + 1, p + nu + 1, p + nu + 2
            idx += [[p, s, q], [q, s, t]]
    return (pos.reshape(-1, 3).astype(f32), np.array(idx, np.uint32), n.reshape(-1, 3).astype(f32),
            np.stack([uu * 4, vv * 2], -1).reshape(-1, 2).astype(f32), tg.reshape(-1, 4).astype(f32))


def value_noise(rng, res: int, octaves=5):
    out = np.zeros((res, res))
    amp, tot = 1.0, 0.0
    for o in range(octaves):
        n = 2 ** (o + 2)
        g = rng.random((n + 1, n + 1))
        g[-1, :], g[:, -1] = g[0, :], g[:, 0]  # tileable
        x = np.linspace(0, n, res, endpoint=False)
        xi = x.astype(int)
        xf = x - xi
        xf = xf * xf * (3 - 2 * xf)
        a = g[np.ix_(xi, xi)] * (1 - xf)[None, :] + g[np.ix_(xi, xi + 1)] * xf[None, :]
        b = g[np.ix_(xi + 1, xi)] * (1 - xf)[None, :] + g[np.ix_(xi + 1, xi + 1)] * xf[None, :]
        out += amp * (a * (1 - xf)[:, None] + b * xf[:, None])
        tot += amp
        amp *= 0.5
    return out / tot


def quat_from_axis_angle(axis, angle):
    axis = np.asarray(axis, np.float64)
    axis = axis / np.linalg.norm(axis)
    s = np.sin(angle / 2)
    return [axis[0] * s, axis[1] * s, axis[2] * s, np.cos(angle / 2)]


# --- scenes -------------------------------------------------------------------------------------

def cornell(path: str) -> str:
    """C1: Cornell box — 5 wall quads + 2 boxes + ceiling light quad = 36 triangles."""
    w = GltfWriter()
    white = w.material((0.73, 0.73, 0.73), 0.0, 1.0)
    red = w.material((0.65, 0.05, 0.05), 0.0, 1.0)
    green = w.material((0.12, 0.45, 0.15), 0.0, 1.0)
    light = w.material((0.0, 0.0, 0.0), 0.0, 1.0, emissive=(1, 1, 1), strength=15.0)
    glossy = w.material((0.8, 0.8, 0.85), 1.0, 0.25)

    def add(pos_idx, mat):
        pos, idx = pos_idx
        w.node(mesh=w.mesh(w.geometry(pos, idx), mat))

    add(quad((-1, -1, 1), (1, -1, 1), (1, -1, -1), (-1, -1, -1)), white)   # floor, normal +y
    add(quad((-1, 1, 1), (-1, 1, -1), (1, 1, -1), (1, 1, 1)), white)       # ceiling, normal -y
    add(quad((-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1)), white)   # back, normal +z
    add(quad((-1, -1, 1), (-1, -1, -1), (-1, 1, -1), (-1, 1, 1)), red)     # left, normal +x
    add(quad((1, -1, 1), (1, 1, 1), (1, 1, -1), (1, -1, -1)), green)       # right, normal -x
    # light just below the ceiling (not coplanar: no exact-t ties with the ceiling quad)
    add(quad((-0.25, 0.995, 0.25), (-0.25, 0.995, -0.25), (0.25, 0.995, -0.25), (0.25, 0.995, 0.25)), light)
    pos, idx = box((-0.3, -0.6, -0.3), (0.3, 0.6, 0.3))
    w.node(mesh=w.mesh(w.geometry(pos, idx), white), translation=(-0.35, -0.397, -0.3),  # 3 mm above the floor: no coplanar faces
           rotation=quat_from_axis_angle((0, 1, 0), 0.3))
    pos, idx = box((-0.3, -0.3, -0.3), (0.3, 0.3, 0.3))
    w.node(mesh=w.mesh(w.geometry(pos, idx), glossy), translation=(0.35, -0.697, 0.3),
           rotation=quat_from_axis_angle((0, 1, 0), -0.3))
    w.camera_look_at((0, 0, 3.9), (0, 0, 0), yfov=0.69)
    return w.save(path)


def spheres(path: str, n_spheres=78, subdiv=3, seed=1234, box_size=20.0, n_emissive=4) -> str:
    """C2: ground quad + instanced icospheres with random TRS, random roughness/metallic,
    a few emissive spheres.  n_spheres=78, subdiv=3 -> 2 + 78*1280 = 99 842 triangles."""
    rng = np.random.default_rng(seed)
    w = GltfWriter()
    pos, idx, nrm = icosphere(subdiv)
    geom = w.geometry(pos, idx, normal=nrm)
    h = box_size / 2
    ground = w.material((0.6, 0.6, 0.6), 0.0, 0.8)
    gp, gi = quad((-2 * h, -h, 2 * h), (2 * h, -h, 2 * h), (2 * h, -h, -2 * h), (-2 * h, -h, -2 * h))
    w.node(mesh=w.mesh(w.geometry(gp, gi), ground))
    for i in range(n_spheres):
        if i < n_emissive:
            mat = w.material((0, 0, 0), 0.0, 1.0, emissive=tuple(0.5 + 0.5 * rng.random(3)), strength=12.0)
        else:
            mat = w.material(tuple(0.2 + 0.75 * rng.random(3)), float(rng.integers(0, 2)),
                             float(rng.uniform(0.05, 1.0)))
        s = rng.uniform(0.8, 2.2) * (box_size / 20.0)
        sc = s * (1 + 0.3 * (rng.random(3) - 0.5))
        axis = rng.normal(size=3)
        w.node(mesh=w.mesh(geom, mat), translation=tuple(rng.uniform(-h * 0.85, h * 0.85, 3)),
               rotation=quat_from_axis_angle(axis, rng.uniform(0, 2 * np.pi)), scale=tuple(sc))
    w.camera_look_at((0.0, h * 0.45, h * 2.9), (0, -h * 0.15, 0), yfov=0.62)
    return w.save(path)


def write_env_hdr(path: str, width=2048, height=1024, sun_peak=5e3) -> str:
    """Sky gradient + sun disk as a Radiance .hdr (decoded as 3 x f32 like stbi.loadf)."""
    import cv2

    v = (np.arange(height) + 0.5) / height
    u = (np.arange(width) + 0.5) / width
    el = (0.5 - v) * np.pi
    az = (u - 0.5) * 2 * np.pi
    t = np.clip(np.sin(el) * 0.5 + 0.5, 0, 1)[:, None]
    sky = (1 - t) * np.array([0.9, 0.85, 0.8]) + t * np.array([0.25, 0.45, 0.95])
    img = np.broadcast_to(sky[:, None, :], (height, width, 3)).copy()
    img[el < 0] *= 0.35
    d = np.stack([np.cos(el)[:, None] * np.cos(az)[None, :], np.broadcast_to(np.sin(el)[:, None], (height, width)),
                  np.cos(el)[:, None] * np.sin(az)[None, :]], -1)
    sun = np.array([0.5, 0.6, 0.62])
    sun /= np.linalg.norm(sun)
    c = d @ sun
    img += (np.clip((c - 0.9985) / (1 - 0.9985), 0, 1)[..., None] ** 2) * sun_peak * np.array([1.0, 0.95, 0.85])
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    cv2.imwrite(path, img[:, :, ::-1].astype(np.float32))
    return path


def textured(path: str, seed=7, tex_res=1024, detail=1.0) -> str:
    """C3: UV-sphere + torus + plane with baseColor / metallicRoughness / normal / emissive PNG maps
    and TANGENT attributes (≈50k triangles at detail=1)."""
    import cv2

    rng = np.random.default_rng(seed)
    root = os.path.dirname(os.path.abspath(path))
    os.makedirs(root, exist_ok=True)
    stem = os.path.splitext(os.path.basename(path))[0]
    n1, n2 = value_noise(rng, tex_res), value_noise(rng, tex_res)
    yy, xx = np.mgrid[0:tex_res, 0:tex_res]
    checker = (((xx * 8 // tex_res) + (yy * 8 // tex_res)) % 2).astype(np.float64)
    base = np.stack([0.25 + 0.7 * n1, 0.3 + 0.5 * checker, 0.9 - 0.6 * n2], -1)
    mr = np.stack([np.ones_like(n1), 0.08 + 0.8 * n2, (n1 > 0.5).astype(np.float64)], -1)
    gy, gx = np.gradient(n1 * 6.0)
    nm = np.stack([-gx * tex_res / 64, -gy * tex_res / 64, np.ones_like(n1)], -1)
    nm /= np.linalg.norm(nm, axis=-1, keepdims=True)
    nm = nm * 0.5 + 0.5
    em = np.zeros((tex_res, tex_res, 3))
    em[(n2 > 0.72)] = [1.0, 0.55, 0.2]

    def save_png(name, img):
        cv2.imwrite(os.path.join(root, name), (np.clip(img, 0, 1) * 255 + 0.5).astype(np.uint8)[:, :, ::-1])
        return name

    w = GltfWriter()
    t_base = w.texture(save_png(f"{stem}_basecolor.png", base))
    t_mr = w.texture(save_png(f"{stem}_mr.png", mr))
    t_nm = w.texture(save_png(f"{stem}_normal.png", nm))
    t_em = w.texture(save_png(f"{stem}_emissive.png", em))
    m_full = w.material((1, 1, 1), 1.0, 1.0, color_tex=t_base, mr_tex=t_mr, normal_tex=t_nm)
    m_glow = w.material((1, 1, 1), 1.0, 1.0, emissive=(1, 1, 1), strength=6.0, color_tex=t_base, mr_tex=t_mr,
                        normal_tex=t_nm, emissive_tex=t_em)
    m_plane = w.material((0.8, 0.8, 0.8), 0.0, 1.0, color_tex=t_base, mr_tex=t_mr)
    nu = max(8, int(128 * detail))
    p, i, n, uv, tg = uv_sphere(nu, nu // 2, 1.0)
    w.node(mesh=w.mesh(w.geometry(p, i, n, uv * np.array([2, 1], f32), tg), m_full), translation=(-1.3, 0.0, 0))
    p, i, n, uv, tg = torus(nu, max(8, int(64 * detail)))
    w.node(mesh=w.mesh(w.geometry(p, i, n, uv, tg), m_glow), translation=(1.4, -0.3, 0.2),
           rotation=quat_from_axis_angle((1, 0, 0.3), 0.9))
    ng_ = max(2, int(32 * detail))
    gx_, gz_ = np.meshgrid(np.linspace(-6, 6, ng_ + 1), np.linspace(-6, 6, ng_ + 1), indexing="xy")
    pp = np.stack([gx_, np.full_like(gx_, -1.0), gz_], -1).reshape(-1, 3)
    pi_ = []
    for a in range(ng_):
        for b in range(ng_):
            q = a * (ng_ + 1) + b
            pi_ += [[q, q + ng_ + 1, q + 1], [q + 1, q + ng_ + 1, q + ng_ + 2]]
    puv = np.stack([gx_ / 3, gz_ / 3], -1).reshape(-1, 2)
    w.node(mesh=w.mesh(w.geometry(pp, np.array(pi_), np.tile([0, 1, 0], (len(pp), 1)), puv,
                                  np.tile([1, 0, 0, 1], (len(pp), 1))), m_plane))
    w.camera_look_at((0.0, 1.2, 5.2), (0, -0.1, 0), yfov=0.6)
    return w.save(path)


def terrain(path: str, grid=500, n_spheres=391, subdiv=3, seed=42, n_emissive=16, extent=100.0) -> str:
    """C4/C5: displaced-grid terrain (2*grid^2 tris) + instanced icospheres.
    grid=500, n_spheres=391 -> 500 000 + 500 480 ≈ 1.0M;  grid=1500, n_spheres=4300 -> ≈10M."""
    rng = np.random.default_rng(seed)
    w = GltfWriter()
    res = grid + 1
    hmap = value_noise(rng, res, octaves=6)
    xs = np.linspace(-extent / 2, extent / 2, res)
    X, Z = np.meshgrid(xs, xs, indexing="xy")
    Y = (hmap - 0.5) * extent * 0.18
    pos = np.stack([X, Y, Z], -1).reshape(-1, 3)
    gz, gx = np.gradient(Y, xs, xs)
    nrm = np.stack([-gx, np.ones_like(gx), -gz], -1)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    q = (np.arange(grid)[:, None] * res + np.arange(grid)[None, :]).reshape(-1)
    idx = np.stack([np.stack([q, q + res, q + 1], -1), np.stack([q + 1, q + res, q + res + 1], -1)], 1).reshape(-1, 3)
    ground = w.material((0.45, 0.5, 0.35), 0.0, 0.9)
    w.node(mesh=w.mesh(w.geometry(pos, idx, nrm.reshape(-1, 3)), ground))
    sp, si, sn = icosphere(subdiv)
    geom = w.geometry(sp, si, normal=sn)
    for i in range(n_spheres):
        if i < n_emissive:
            mat = w.material((0, 0, 0), 0.0, 1.0, emissive=tuple(0.5 + 0.5 * rng.random(3)), strength=40.0)
        else:
            mat = w.material(tuple(0.2 + 0.75 * rng.random(3)), float(rng.integers(0, 2)),
                             float(rng.uniform(0.05, 1.0)))
        x, z = rng.uniform(-extent * 0.45, extent * 0.45, 2)
        ix = int((x / extent + 0.5) * grid)
        iz = int((z / extent + 0.5) * grid)
        s = rng.uniform(0.6, 2.4) * (extent / 100.0) * (1.0 if n_spheres < 1000 else 0.45)
        y = Y[iz, ix] + s * rng.uniform(0.6, 3.0)
        w.node(mesh=w.mesh(geom, mat), translation=(x, y, z),
               rotation=quat_from_axis_angle(rng.normal(size=3), rng.uniform(0, 2 * np.pi)),
               scale=(s, s * rng.uniform(0.8, 1.2), s))
    w.camera_look_at((0.0, extent * 0.16, extent * 0.62), (0, -extent * 0.02, 0), yfov=0.6)
    return w.save(path)


# BASELINE.json configs -> (generator kwargs, width, height, ray_depth, spp)
CONFIGS = {
    "C1": dict(gen="cornell", kwargs={}, width=256, height=256, ray_depth=6, spp=64, env=False),
    "C2": dict(gen="spheres", kwargs={}, width=1920, height=1080, ray_depth=8, spp=256, env=False),
    "C3": dict(gen="textured", kwargs={}, width=1920, height=1080, ray_depth=8, spp=1024, env=True),
    "C4": dict(gen="terrain", kwargs={}, width=1920, height=1080, ray_depth=10, spp=4096, env=False),
    "C5": dict(gen="terrain", kwargs=dict(grid=1500, n_spheres=4300, seed=43), width=3840, height=2160,
               ray_depth=12, spp=None, env=False),
}


def generate(config: str, out_dir: str, **overrides):
    """Write the glTF (and env map) of a BASELINE config into out_dir; returns (gltf, env or None)."""
    cfg = CONFIGS[config]
    kwargs = dict(cfg["kwargs"])
    kwargs.update(overrides)
    path = os.path.join(out_dir, f"{config.lower()}.gltf")
    globals()[cfg["gen"]](path, **kwargs)
    env = None
    if cfg["env"]:
        env = write_env_hdr(os.path.join(out_dir, f"{config.lower()}_env.hdr"),
                            *( (512, 256) if overrides.get("tex_res", 1024) < 1024 else (2048, 1024) ))
    return path, env
