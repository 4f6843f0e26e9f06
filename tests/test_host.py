"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header
declares (no compute calls without a GPU), the product's bvh_build equals the oracle's, the glTF
stand-in loader follows input.odin, the generators hit BASELINE's triangle counts."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import binding as orc
from raytracer_odin_b200 import api, cabi, gltf, multigpu, output, scenegen
from raytracer_odin_b200.scene import native_bvh_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    with open(os.path.join(ROOT, "include", "odinrt_b200.h")) as f:
        hdr = f.read()
    declared = set(re.findall(r"\b(ort_[a-z0-9_]+)\s*\(", hdr))
    assert {"ort_create", "ort_upload_scene", "ort_render", "ort_render_device", "ort_trace_rays",
            "ort_primary_hits", "ort_bvh_build"} <= declared
    lib = cabi.load_library()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(cabi.ABI), "cabi.ABI and the header disagree"
    assert lib.ort_abi_version() == 2


def test_struct_layouts_match_the_reference():
    # Triangle 168 B (raytracer.odin:18-23), Sample_Stats 52 B (main.odin:34-40)
    off = {n: cabi.TRI_DTYPE.fields[n][1] for n in cabi.TRI_DTYPE.names}
    assert off == {"p": 0, "u": 12, "v": 24, "n1": 36, "n2": 48, "n3": 60, "ng": 72, "tex1": 84, "tex2": 92,
                   "tex3": 100, "tan1": 108, "tan2": 124, "tan3": 140, "material_index": 160}
    assert cabi.STATS_DTYPE.fields["count"][1] == 12 and cabi.STATS_DTYPE.fields["total"][1] == 28
    assert C.sizeof(cabi.OrtTexture) == 32 and C.sizeof(cabi.OrtCamera) == 52


def test_no_gpu_means_loud_failure():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.OrtError, match="no CUDA device|CPU fallback"):
        api.Renderer()


def _same_build(tris):
    a, b = tris.copy(), tris.copy()
    na, nb = orc.bvh_build(a), native_bvh_build(b)
    assert na.tobytes() == nb.tobytes()
    for f in cabi.TRI_DTYPE.names:
        assert np.array_equal(a[f], b[f], equal_nan=True), f
    return na


def test_native_bvh_build_equals_oracle(scene_dir):
    from tests.golden.make_golden import soup

    _same_build(soup(3000, 8))
    _same_build(soup(70000, 9))  # exercises the radix path and the task-parallel subtrees
    # heavy ties: a regular grid of identical quads (equal lo keys everywhere) and -0.0 / +0.0 keys
    g = soup(4096, 10)
    ix = np.arange(4096)
    g["p"] = np.stack([(ix % 16) - 8.0, ((ix // 16) % 16) - 8.0, (ix // 256) * 0.0], 1).astype(np.float32)
    g["p"][::7, 2] = -0.0
    g["u"], g["v"] = np.float32([1, 0, 0]), np.float32([0, 1, 0])
    _same_build(g)
    for name, kw in (("cornell", {}), ("spheres", dict(n_spheres=10, subdiv=2)), ("terrain", dict(grid=40, n_spheres=8, subdiv=1))):
        s = gltf.read_gltf(getattr(scenegen, name)(os.path.join(scene_dir, f"nb_{name}.gltf"), **kw))
        _same_build(s.triangles)
    assert len(native_bvh_build(np.zeros(0, cabi.TRI_DTYPE))) == 1


def test_gltf_loader_follows_input_odin(scene_dir):
    p = scenegen.cornell(os.path.join(scene_dir, "ld_c1.gltf"))
    s = gltf.read_gltf(p)
    assert len(s.triangles) == 36 and len(s.materials) == 1 + 8  # dummy + one per primitive instance
    assert np.all(s.materials[0]["color_factor"] == 0) and s.materials[0]["color_texture"] == -1
    assert s.triangles["material_index"].min() == 1
    # camera: pos = transform column 3, basis column 2 = -Z of the node (input.odin:104-107)
    np.testing.assert_allclose(s.cam_pos, [0, 0, 3.9], atol=1e-6)
    np.testing.assert_allclose(s.cam_basis, np.diag([1, 1, -1]), atol=1e-6)
    assert abs(s.fov_x - 0.69) < 1e-6
    assert abs(s.apply_render_config(1920, 1080) - 0.69 * 1920 / 1080) < 1e-5  # main.odin:202-203
    # emissive strength multiplies the factor (input.odin:157-159); lights collected by emission_factor
    assert np.allclose(s.materials["emission_factor"].max(axis=1).max(), 15.0)
    s.finish(native_bvh_build)
    assert len(s.light_triangles) == 2 and len(s.light_bvh) == 1
    # no NORMAL attribute -> vertex normals are the geometric normal (input.odin:198-201)
    assert np.array_equal(s.triangles["n1"], s.triangles["ng"])
    # no TANGENT attribute -> normalize(0) = NaN like the reference (input.odin:193-195)
    assert np.isnan(s.triangles["tan1"][:, :3]).all()
    ng = np.cross(s.triangles["u"], s.triangles["v"])
    np.testing.assert_allclose(s.triangles["ng"], ng / np.linalg.norm(ng, axis=1, keepdims=True), atol=1e-6)


def test_gltf_textures_and_env(scene_dir):
    p = scenegen.textured(os.path.join(scene_dir, "ld_c3.gltf"), tex_res=32, detail=0.1)
    s = gltf.read_gltf(p)
    assert len(s.textures) == 4 and all(t.dtype == np.uint8 and t.shape == (32, 32, 3) for t in s.textures)
    m = s.materials[1]
    assert m["color_texture"] >= 0 and m["normal_texture"] >= 0 and m["metallic_roughness_texture"] >= 0
    assert np.isfinite(s.triangles["tan1"]).all() and np.all(np.abs(s.triangles["tan1"][:, 3]) == 1)
    env = gltf.load_texture(scenegen.write_env_hdr(os.path.join(scene_dir, "e.hdr"), 64, 32))
    assert env.dtype == np.float32 and env.shape == (32, 64, 3) and env.max() > 100


def test_scenegen_counts(scene_dir):
    s = gltf.read_gltf(scenegen.spheres(os.path.join(scene_dir, "cnt_c2.gltf")))
    assert len(s.triangles) == 2 + 78 * 1280 == 99842
    s = gltf.read_gltf(scenegen.terrain(os.path.join(scene_dir, "cnt_c4.gltf"), grid=50, n_spheres=3, subdiv=1))
    assert len(s.triangles) == 2 * 50 * 50 + 3 * 80


def test_output_and_partition(tmp_path):
    px = np.zeros(6, cabi.STATS_DTYPE)
    px["count"] = 2
    px["total"] = np.float32([[0, 0, 0], [2, 2, 2], [1e3, 0, 0], [0.2, 0.4, 0.8], [1, 1, 1], [4, 4, 4]])
    output.save_result(px, 3, 2, str(tmp_path / "o.ppm"))
    raw = (tmp_path / "o.ppm").read_bytes()
    assert raw.startswith(b"P6\n3 2\n255\n") and len(raw) == 11 + 18
    with pytest.raises(RuntimeError):
        output.save_result(px, 3, 2, str(tmp_path / "o.jpg"))
    # sample partition: contiguous, disjoint, complete, balanced to one sample
    for n, w in ((4096, 8), (10, 4), (3, 8), (0, 2)):
        parts = [multigpu.sample_partition(5, n, r, w) for r in range(w)]
        assert parts[0][0] == 5 and sum(c for _, c in parts) == n
        for (f0, c0), (f1, _) in zip(parts, parts[1:]):
            assert f0 + c0 == f1
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_wide_bvh_emission_is_thread_count_independent(scene_dir):
    """The level-synchronous parallel re-emission (ort_upload_scene's host stage) gives byte-identical
    nodes, depth and stack bound on 1, 3 and 16 threads; breadth-first order, root = node 0."""
    from tests.golden.make_golden import soup

    lib = cabi.load_library()
    s = gltf.read_gltf(scenegen.terrain(os.path.join(scene_dir, "wide_c4.gltf"), grid=120, n_spheres=40, subdiv=2, seed=3))
    for tris in (soup(1, 5), soup(9, 6), soup(20000, 7), s.triangles.copy()):
        nodes = native_bvh_build(tris)
        outs = []
        for threads in (1, 3, 16):
            buf = np.zeros((2 * len(tris) + 4) * 128, np.uint8)
            d, m = C.c_int32(), C.c_int32()
            n = lib.ort_wide_bvh_emit(cabi.ptr(nodes), len(nodes), len(tris), threads, cabi.ptr(buf), len(buf) // 128,
                                      C.byref(d), C.byref(m))
            assert n > 0
            outs.append((n, d.value, m.value, buf[: n * 128].tobytes()))
        assert outs[0] == outs[1] == outs[2]
        n, depth, max_stack, raw = outs[0]
        child = np.frombuffer(raw, np.int32).reshape(n, 32)[:, 24:28]
        inner = child[(child >= 0)]
        assert sorted(inner.tolist()) == list(range(1, n))  # every node but the root is referenced exactly once
        assert np.all(child[child >= 0].reshape(-1) > np.repeat(np.arange(n), 4).reshape(n, 4)[child >= 0])  # children after parents
        leaves = child[(child < 0) & (child != np.int32(-2**31))]
        cnt = (~leaves) & 7
        assert cnt.sum() == len(tris) and 1 <= depth <= max_stack
    assert lib.ort_wide_bvh_emit(cabi.ptr(nodes), 0, len(tris), 1, None, 0, None, None) == -1
