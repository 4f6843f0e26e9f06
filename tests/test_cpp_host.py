"""The C++ host above the C ABI (raytracer-odin_b200/host): native read_gltf / load_texture /
finish_scene / get_rgb_image / save_result and the `odinrt` command line with the reference's flags
(main.odin:174-253).  CPU tests compare it bit for bit with the Python stand-ins; the GPU test runs
the binary end to end and compares its accumulators with the Python CLI's."""
import os
import subprocess

import numpy as np
import pytest

from raytracer_odin_b200 import api, cabi, gltf, hostlib, output, scenegen
from raytracer_odin_b200.scene import native_bvh_build


def _same_scene(a, b, finished=False):
    for f in cabi.TRI_DTYPE.names:
        assert np.array_equal(a.triangles[f], b.triangles[f], equal_nan=True), f
    assert a.materials.tobytes() == b.materials.tobytes()
    assert np.array_equal(a.cam_pos, b.cam_pos) and np.array_equal(a.cam_basis, b.cam_basis) and a.fov_x == b.fov_x
    assert len(a.textures) == len(b.textures)
    for x, y in zip(a.textures, b.textures):
        assert x.dtype == y.dtype and np.array_equal(x, y)
    if finished:
        assert a.bvh.tobytes() == b.bvh.tobytes() and a.light_bvh.tobytes() == b.light_bvh.tobytes()
        assert len(a.light_triangles) == len(b.light_triangles)
        for f in ("p", "u", "v", "material_index"):
            assert np.array_equal(a.light_triangles[f], b.light_triangles[f]), f


@pytest.mark.parametrize("name", ["cornell", "spheres", "terrain"])
def test_native_loader_equals_python_loader(scene_dir, name):
    p = os.path.join(scene_dir, f"cpp_{name}.gltf")
    if name == "cornell":
        scenegen.cornell(p)
    elif name == "spheres":
        scenegen.spheres(p, n_spheres=14, subdiv=2, seed=5)
    else:
        scenegen.terrain(p, grid=48, n_spheres=24, subdiv=2, seed=3, n_emissive=3)
    _same_scene(gltf.read_gltf(p), hostlib.read_gltf(p))
    # finish_scene: same light list, same BVHs, same post-build triangle order
    a = gltf.read_gltf(p).finish(native_bvh_build)
    b = hostlib.read_gltf(p, finish=True)
    _same_scene(a, b, finished=True)


def test_native_loader_full_size_c4(scene_dir):
    """BASELINE config 4 at full size (1 000 480 triangles in 392 primitives): bit-identical to the Python
    loader, and in about a second (a per-primitive exact reserve once made this quadratic: 50 s)."""
    import time

    p = scenegen.terrain(os.path.join(scene_dir, "cpp_c4_full.gltf"))
    t0 = time.perf_counter()
    b = hostlib.read_gltf(p)
    dt = time.perf_counter() - t0
    assert len(b.triangles) == 1_000_480 and dt < 15.0, dt
    _same_scene(gltf.read_gltf(p), b)


def test_native_textures_png_and_hdr(scene_dir, tmp_path):
    p = scenegen.textured(os.path.join(scene_dir, "cpp_c3.gltf"), tex_res=64, detail=0.15)
    env = scenegen.write_env_hdr(os.path.join(scene_dir, "cpp_env.hdr"), 128, 64)
    a, b = gltf.read_gltf(p), hostlib.read_gltf(p, env)
    _same_scene(a, b)
    assert len(b.textures) == 4 and np.isfinite(b.triangles["tan1"]).all()
    e = gltf.load_texture(env)
    assert b.env_map.dtype == np.float32 and np.array_equal(e, b.env_map)
    # PNG flavours stb decodes to their native channel count: gray, gray+alpha, RGBA, 16-bit, palette
    import cv2

    rng = np.random.default_rng(1)
    for name, img in (("g.png", rng.integers(0, 256, (5, 7), np.uint8)),
                      ("ga.png", None),
                      ("rgba.png", rng.integers(0, 256, (6, 4, 4), np.uint8)),
                      ("rgb16.png", rng.integers(0, 65536, (3, 9, 3), np.uint16))):
        path = str(tmp_path / name)
        if img is None:
            continue
        cv2.imwrite(path, img)
        gl = {"asset": {"version": "2.0"}, "scenes": [{"nodes": []}]}
        gp = str(tmp_path / (name + ".gltf"))
        import json

        with open(gp, "w") as f:
            json.dump(gl, f)
        got = hostlib.read_gltf(gp, path).env_map
        assert np.array_equal(got, gltf.load_texture(path)), name
    with pytest.raises(RuntimeError, match="Failed to read texture file"):
        hostlib.read_gltf(p, str(tmp_path / "missing.hdr"))


def test_native_loader_errors(tmp_path):
    with pytest.raises(RuntimeError, match="Failed to open input file"):
        hostlib.read_gltf(str(tmp_path / "nope.gltf"))
    bad = tmp_path / "bad.gltf"
    bad.write_text("{ not json")
    with pytest.raises(RuntimeError, match="Failed to parse .gltf file"):
        hostlib.read_gltf(str(bad))
    # primitive without a material: the reference dereferences nil (input.odin:138); reported here
    import json

    nomat = {"asset": {"version": "2.0"}, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
             "meshes": [{"primitives": [{"attributes": {"POSITION": 0}}]}],
             "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"}],
             "bufferViews": [{"buffer": 0, "byteLength": 36}],
             "buffers": [{"uri": "data:application/octet-stream;base64," + "A" * 48, "byteLength": 36}]}
    p = tmp_path / "nomat.gltf"
    p.write_text(json.dumps(nomat))
    with pytest.raises(RuntimeError, match="without a material"):
        hostlib.read_gltf(str(p))


def test_native_output_equals_python(tmp_path):
    rng = np.random.default_rng(2)
    w, h = 17, 9
    px = np.zeros(w * h, cabi.STATS_DTYPE)
    px["count"] = rng.integers(1, 9, w * h)
    px["total"] = (rng.gamma(0.6, 2.0, (w * h, 3)) * px["count"][:, None]).astype(np.float32)
    px["count"][3] = 0  # 0/0 -> NaN -> 0 like linalg.max(raw, 0) then to_u8
    a, b = output.get_rgb_image(px, w, h), hostlib.get_rgb_image(px, w, h)
    assert np.abs(a.astype(int) - b.astype(int)).max() <= 1 and (a != b).mean() < 0.01  # pow() ulps at .5 boundaries
    hostlib.save_result(px, w, h, str(tmp_path / "o.ppm"))
    raw = (tmp_path / "o.ppm").read_bytes()
    assert raw.startswith(b"P6\n17 9\n255\n") and raw[len(b"P6\n17 9\n255\n"):] == b.tobytes()
    hostlib.save_result(px, w, h, str(tmp_path / "o.png"))
    import cv2

    back = cv2.imread(str(tmp_path / "o.png"), cv2.IMREAD_UNCHANGED)[:, :, ::-1]
    assert np.array_equal(back, b)
    with pytest.raises(RuntimeError, match="Unsupported file format"):
        hostlib.save_result(px, w, h, str(tmp_path / "o.jpg"))


def test_cli_without_gpu_fails_loudly(scene_dir):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    hostlib.load()
    p = scenegen.cornell(os.path.join(scene_dir, "cpp_cli_nogpu.gltf"))
    r = subprocess.run([hostlib.CLI_PATH, p, "--width", "8", "--height", "8", "--ray-depth", "2", "--num-samples", "1"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr  # no CPU fallback
    r = subprocess.run([hostlib.CLI_PATH], capture_output=True, text=True)
    assert r.returncode == 1 and "input_file" in r.stderr


@pytest.mark.gpu
def test_cpp_cli_equals_python_cli(scene_dir, tmp_path):
    """odinrt (C++) and the Python mirror render the same accumulators: same loader output, same
    BVH, same counter-based streams; --times, checkpoint/resume and the image writers included."""
    from raytracer_odin_b200 import cli

    hostlib.load()
    p = scenegen.textured(os.path.join(scene_dir, "cpp_cli_c3.gltf"), tex_res=64, detail=0.15)
    env = scenegen.write_env_hdr(os.path.join(scene_dir, "cpp_cli_env.hdr"), 128, 64)
    common = ["--width", "48", "--height", "32", "--ray-depth", "4", "--seed", "3", "--env-map", env]
    ck_c, ck_p = str(tmp_path / "c.ckpt"), str(tmp_path / "p.ckpt")
    out_c, out_p = str(tmp_path / "c.ppm"), str(tmp_path / "p.ppm")
    r = subprocess.run([hostlib.CLI_PATH, p, out_c] + common + ["--num-samples", "6", "--checkpoint", ck_c],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Trial 0 >>> Rendered in" in r.stdout and "Mrays/s" in r.stdout
    cli.main([p, out_p] + common + ["--num-samples", "6", "--checkpoint", ck_p])
    a, na = api.load_checkpoint(ck_c, 48, 32)
    b, nb = api.load_checkpoint(ck_p, 48, 32)
    assert na == nb == 6 and a.tobytes() == b.tobytes()
    ia, ib = np.frombuffer(open(out_c, "rb").read()[-48 * 32 * 3:], np.uint8), np.frombuffer(open(out_p, "rb").read()[-48 * 32 * 3:], np.uint8)
    assert np.abs(ia.astype(int) - ib.astype(int)).max() <= 1
    # resume 3 + 3 == 6, --times 2 replays the same samples into uncleared accumulators, device BVH build
    r = subprocess.run([hostlib.CLI_PATH, p] + common + ["--num-samples", "3", "--checkpoint", ck_c, "--bvh", "device"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([hostlib.CLI_PATH, p, str(tmp_path / "c.png")] + common + ["--num-samples", "3", "--resume", ck_c, "--checkpoint", ck_c],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    c, nc = api.load_checkpoint(ck_c, 48, 32)
    assert nc == 6 and np.all(c["count"] == 6)
    np.testing.assert_allclose(c["total"], a["total"], rtol=1e-5, atol=1e-6)
    r = subprocess.run([hostlib.CLI_PATH, p] + common + ["--num-samples", "2", "--times", "2", "--gpus", "0,0", "--checkpoint", ck_c],
                       capture_output=True, text=True)
    assert r.returncode == 0 and "Performance Summary" in r.stdout, r.stderr
    t2, _ = api.load_checkpoint(ck_c, 48, 32)
    assert np.all(t2["count"] == 4)


@pytest.mark.gpu
def test_cpp_cli_continious_until_sigint(scene_dir, tmp_path):
    """--continious (main.odin:207): render 16-sample chunks until SIGINT, then write the image and the
    checkpoint; the sample count is whatever finished, every pixel got the same number of samples."""
    import signal
    import time

    hostlib.load()
    p = scenegen.cornell(os.path.join(scene_dir, "cpp_cli_cont.gltf"))
    out, ck = str(tmp_path / "cont.ppm"), str(tmp_path / "cont.ckpt")
    proc = subprocess.Popen([hostlib.CLI_PATH, p, out, "--width", "64", "--height", "48", "--ray-depth", "4",
                             "--continious", "--checkpoint", ck], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    time.sleep(6.0)
    proc.send_signal(signal.SIGINT)
    stdout, stderr = proc.communicate(timeout=60)
    assert proc.returncode == 0, stderr
    assert "Rendered" in stdout and "samples in" in stdout
    px, nxt = api.load_checkpoint(ck, 64, 48)
    assert nxt > 0 and nxt % 16 == 0
    # an interrupted last chunk may be partial, but whole waves only: counts are uniform over the image
    assert px["count"].min() == px["count"].max() and 0 < px["count"][0] <= nxt
    assert os.path.getsize(out) == len(b"P6\n64 48\n255\n") + 64 * 48 * 3
