"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Integer / index work is bit-exact; radiance tolerances are written in the test."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 1234


@pytest.fixture(scope="module")
def orc():
    from oracle import binding

    return binding


def _renderer(scene, **kw):
    from raytracer_odin_b200 import api

    return api.Renderer(device=0, seed=SEED, **kw).upload_scene(scene)


def _hits_equal(g, o, what, ties=None, max_tie_frac=1e-5):
    """Bit-exact gate.  `ties` (from the oracle) marks rays where a DIFFERENT triangle produced a
    t bit-identical to the winner's: there the reference keeps whichever its own pop order visits
    first (strict `<`, raytracer.odin:360,388), which no other traversal order can reproduce, so
    the gate is: every tie-free ray identical in id / primitive / inside / t / u / v bits, every
    tied ray identical in t bits, and ties rarer than max_tie_frac."""
    free = np.ones(len(o), bool) if ties is None else ties == 0
    n_ties = int((~free).sum())
    assert n_ties <= max(max_tie_frac * len(o), 0), f"{what}: {n_ties} exact-t ties in {len(o)} rays"
    bad = np.sum(g["tri"][free] != o["tri"][free])
    assert bad == 0, f"{what}: triangle ids differ on {bad} tie-free rays"
    assert np.array_equal(g["material"][free], o["material"][free]), f"{what}: primitive (material) ids differ"
    assert np.array_equal(g["inside"][free], o["inside"][free]), f"{what}: inside flags differ"
    assert np.array_equal(g["tri"] >= 0, o["tri"] >= 0), f"{what}: hit/miss differs"
    hit = o["tri"] >= 0
    assert np.array_equal(g["t"][hit].view(np.uint32), o["t"][hit].view(np.uint32)), f"{what}: t bits differ"
    for f in ("u", "v"):
        m = hit & free
        assert np.array_equal(g[f][m].view(np.uint32), o[f][m].view(np.uint32)), f"{what}: {f} bits differ"
    return n_ties


@pytest.mark.parametrize("name,w,h", [("cornell", 256, 256), ("spheres_small", 320, 180), ("terrain_small", 320, 180),
                                      ("textured_small", 160, 90), ("spheres_nolight", 64, 64)])
def test_primary_hits_bit_exact(scenes, orc, name, w, h):
    scene = scenes(name, w, h)
    o = orc.OracleScene(scene)
    with _renderer(scene) as r:
        for sample in (0, 17):
            g, grays = r.primary_hits(w, h, sample, want_rays=True)
            ref, orays, c = o.primary_hits(w, h, sample=sample, seed=SEED, mode=0)
            assert np.array_equal(grays["d"].view(np.uint32), orays["d"].view(np.uint32)), "primary ray directions differ"
            assert np.array_equal(grays["o"].view(np.uint32), orays["o"].view(np.uint32))
            _hits_equal(g, ref, f"{name} sample {sample}", o.ties)
            assert c["stack_drops"] == 0


@pytest.mark.parametrize("name", ["cornell", "spheres_small", "terrain_small"])
def test_trace_random_rays_bit_exact(scenes, orc, name):
    """cast_ray on incoherent rays: origins inside the scene box, uniform directions."""
    from raytracer_odin_b200 import cabi

    scene = scenes(name)
    rng = np.random.default_rng(11)
    n = 200_000
    lo = scene.bvh[-1]["lo"]
    hi = scene.bvh[-1]["hi"]
    rays = np.zeros(n, cabi.RAY_DTYPE)
    rays["o"] = (lo + (hi - lo) * rng.random((n, 3))).astype(np.float32)
    d = rng.normal(size=(n, 3))
    rays["d"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    o = orc.OracleScene(scene)
    ref, c = o.trace_rays(rays, mode=0)
    with _renderer(scene) as r:
        g = r.trace_rays(rays)
    _hits_equal(g, ref, name, o.ties)
    assert (ref["tri"] >= 0).mean() > 0.2


def test_trace_edge_cases(scenes, orc):
    """Empty input, axis-parallel directions (zero components), rays starting on geometry."""
    from raytracer_odin_b200 import cabi

    scene = scenes("cornell")
    with _renderer(scene) as r:
        assert len(r.trace_rays(np.zeros(0, cabi.RAY_DTYPE))) == 0
        rays = np.zeros(7, cabi.RAY_DTYPE)
        rays["o"] = [[0, 0, 0], [0, 0, 0], [0, 0, 0], [0.1, 0.2, 0.3], [0, -1, 0], [0, 0, 3.9], [5, 5, 5]]
        rays["d"] = [[1, 0, 0], [0, -1, 0], [0, 0, -1], [0, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 0]]
        g = r.trace_rays(rays)
        o = orc.OracleScene(scene)
        ref, _ = o.trace_rays(rays, mode=0)
        _hits_equal(g, ref, "edge cases", o.ties, max_tie_frac=0.5)
        assert g["tri"][-1] == -1 and g["tri"][-2] == -1


@pytest.mark.parametrize("name", ["cornell", "spheres_small", "terrain_small"])
def test_light_pdf_matches_oracle(scenes, orc, name):
    """surface_sampling_pdf (shading.odin:96-100): all-hit sum over the light BVH. f32 sums in a
    different order: relative tolerance 1e-5."""
    from raytracer_odin_b200 import cabi

    scene = scenes(name)
    rng = np.random.default_rng(3)
    n = 50_000
    lo, hi = scene.bvh[-1]["lo"], scene.bvh[-1]["hi"]
    rays = np.zeros(n, cabi.RAY_DTYPE)
    rays["o"] = (lo + (hi - lo) * rng.random((n, 3))).astype(np.float32)
    lt = scene.light_triangles
    pick = rng.integers(0, len(lt), n)
    tgt = lt["p"][pick] + lt["u"][pick] * 0.3 + lt["v"][pick] * 0.3
    d = tgt - rays["o"]
    rays["d"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    ref = orc.OracleScene(scene).light_pdf(rays)
    with _renderer(scene) as r:
        g = r.light_pdf(rays)
    assert (ref > 0).mean() > 0.5
    np.testing.assert_allclose(g, ref, rtol=1e-5, atol=1e-12)


@pytest.mark.parametrize("name,w,h,depth,spp", [("cornell", 64, 64, 6, 16), ("spheres_small", 96, 54, 8, 8),
                                                ("terrain_small", 96, 54, 5, 8), ("textured_small", 96, 54, 8, 8),
                                                ("spheres_nolight", 48, 48, 4, 8)])
def test_render_matches_oracle_same_streams(scenes, orc, name, w, h, depth, spp):
    """Radiance parity with common random numbers: both sides draw the same Philox streams, so
    paths coincide except where an f32 difference (CUDA vs glibc transcendentals) flips a branch.
    Tolerance: relRMSE <= 1e-2 and mean luminance within 0.5 % (BASELINE.json north_star), and
    the ray counts agree to 0.1 %."""
    from raytracer_odin_b200 import api

    scene = scenes(name, w, h)
    with _renderer(scene) as r:
        px = r.render(w, h, depth, spp)
        st = r.stats()
    opx, c = orc.OracleScene(scene).render(w, h, depth, spp, seed=SEED, mode=0, schedule=1)
    assert np.array_equal(px["count"], opx["count"]) and int(px["count"][0]) == spp
    a, b = api.mean_image(px, w, h), api.mean_image(opx, w, h)
    rmse, lum = api.rel_rmse(a, b)
    assert rmse <= 1e-2, (name, rmse, lum)
    assert abs(lum - 1) <= 5e-3, (name, rmse, lum)
    assert abs(st["rays_closest"] - c["rays"]) <= 1e-3 * c["rays"], (st["rays_closest"], c["rays"])
    assert st["rays_traced"] >= st["rays_closest"]
    # most pixels agree to f32 noise
    close = np.isclose(a, b, rtol=1e-3, atol=1e-5).all(axis=2).mean()
    assert close > 0.97, close


def test_render_deterministic_and_split_invariant(scenes):
    """Same seed -> bit-identical accumulators; rendering [0,8) equals [0,4) then [4,8) into the
    same pixels (the sample-split used across GPUs) up to f32 summation order."""
    scene = scenes("spheres_small", 96, 54)
    w, h = 96, 54
    with _renderer(scene) as r:
        a = r.render(w, h, 6, 8)
        b = r.render(w, h, 6, 8)
        c = r.render(w, h, 6, 4, first_sample=0)
        r.render(w, h, 6, 4, first_sample=4, out=c)
    assert a.tobytes() == b.tobytes()
    assert np.array_equal(a["count"], c["count"])
    np.testing.assert_allclose(a["total"], c["total"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(a["first"], c["first"]) and np.array_equal(a["last"], c["last"])
    # a wave capacity smaller than the frame forces many waves: same samples, same sums
    with _renderer(scene, max_paths_in_flight=w * h * 3) as r2:
        d = r2.render(w, h, 6, 8)
    np.testing.assert_allclose(a["total"], d["total"], rtol=1e-5, atol=1e-6)


def test_accumulates_like_times(scenes):
    """--times N re-renders the same samples into uncleared accumulators (raytracer.odin:606-610)."""
    scene = scenes("cornell", 32, 32)
    with _renderer(scene) as r:
        one = r.render(32, 32, 4, 4)
        two, _ = r.render_scene(32, 32, 4, 4, number_of_trials=2, log=None)
    assert np.array_equal(two["count"], 2 * one["count"])
    np.testing.assert_allclose(two["total"], 2 * one["total"], rtol=1e-6)


def test_depth_zero_and_interrupt(scenes):
    scene = scenes("cornell", 32, 32)
    with _renderer(scene) as r:
        z = r.render(32, 32, 0, 3)
        assert np.all(z["count"] == 3) and np.all(z["total"] == 0)
        flag = np.ones(1, np.uint8)  # already interrupted: nothing is rendered, call still succeeds
        out = r.render(32, 32, 4, 8, interrupt=flag)
        assert np.all(out["count"] == 0)


def test_device_accum_and_tonemap(scenes, orc):
    """ort_render_device into a torch-owned buffer, unpack and device tonemap (output.odin:30-80)."""
    import torch

    scene = scenes("cornell", 64, 64)
    w = h = 64
    with _renderer(scene) as r:
        acc = torch.zeros(8, h * w, device="cuda", dtype=torch.float32)
        r.set_stream(torch.cuda.current_stream().cuda_stream)
        r.render_device(w, h, 5, 0, 8, acc.data_ptr())
        torch.cuda.synchronize()
        px = r.unpack_accum(w, h, acc.data_ptr())
        rgb = r.tonemap_rgb8(w, h, acc.data_ptr())
        host = r.render(w, h, 5, 8)
    assert np.array_equal(px["total"], host["total"]) and np.array_equal(px["count"], host["count"])
    ref = orc.get_rgb_image(host, w, h)
    assert np.abs(rgb.astype(int) - ref.astype(int)).max() <= 1


def test_errors_are_reported(scenes):
    from raytracer_odin_b200 import api, cabi

    r = api.Renderer(device=0)
    with pytest.raises(api.OrtError):
        r.render(8, 8, 2, 1)  # no scene
    with pytest.raises(api.OrtError):
        api.Renderer(device=99)
    scene = scenes("cornell")
    bad = scene.bvh.copy()
    bad["a"][-1] = 10_000
    import copy

    s2 = copy.copy(scene)
    s2.bvh = bad
    with pytest.raises(api.OrtError):
        r.upload_scene(s2)
    r.close()


def test_full_size_c2_primary_hits(scenes, orc):
    """BASELINE config 2 at full size (1920x1080, ~100k triangles): every primary hit id and t
    bit-identical to the faithful reference traversal."""
    w, h = 1920, 1080
    scene = scenes("spheres_c2", w, h)
    o = orc.OracleScene(scene)
    ref, _, c = o.primary_hits(w, h, sample=0, seed=SEED, mode=0, threads=orc.load().orc_hardware_threads())
    with _renderer(scene) as r:
        g = r.primary_hits(w, h, 0)
    n_ties = _hits_equal(g, ref, "C2 1080p", o.ties)
    print(f"C2 1080p: {len(ref)} primary rays, {n_ties} exact-t ties, all tie-free rays bit-identical")
    assert c["stack_drops"] == 0


def test_full_size_c4_subsampled_hits_and_split(scenes, orc):
    """BASELINE config 4 at full size (1920x1080, ~1.0 M triangles): 300k of the GPU's own primary
    rays re-traced by the faithful oracle must agree bit for bit (tie-free rays), the stack never
    overflows, and rendering in many small waves gives the same sums as one big wave."""
    w, h = 1920, 1080
    scene = scenes("terrain_c4", w, h)
    assert len(scene.triangles) > 1_000_000
    o = orc.OracleScene(scene)
    with _renderer(scene) as r:
        g, rays = r.primary_hits(w, h, 5, want_rays=True)
        pick = np.random.default_rng(2).choice(w * h, 300_000, replace=False)
        ref, c = o.trace_rays(rays[pick], mode=0, threads=orc.load().orc_hardware_threads())
        n_ties = _hits_equal(g[pick], ref, "C4 1080p", o.ties, max_tie_frac=1e-4)
        assert c["stack_drops"] == 0, c
        a = r.render(w, h, 10, 2)
        st = r.stats()
    with _renderer(scene, max_paths_in_flight=w * h) as r2:
        b = r2.render(w, h, 10, 2)
    assert np.array_equal(a["count"], b["count"])
    np.testing.assert_allclose(a["total"], b["total"], rtol=1e-5, atol=1e-6)
    assert st["wide_depth"] < 40 and np.isfinite(a["total"]).all()
    print(f"C4: {n_ties} ties in 300k rays; reference stack high-water {c['stack_high']}; wide depth {st['wide_depth']}")


def test_full_size_c3_window_radiance(scenes, orc):
    """BASELINE config 3 (textured metallic-roughness + normal maps + HDR env map) at 1920x1080:
    a 160x90 window rendered by the oracle with the same streams; relRMSE <= 1e-2, luminance 0.5 %."""
    from raytracer_odin_b200 import api

    w, h, depth, spp = 1920, 1080, 8, 4
    scene = scenes("textured_c3", w, h)
    win = (880, 495, 1040, 585)
    with _renderer(scene) as r:
        px = r.render(w, h, depth, spp)
    opx, _ = orc.OracleScene(scene).render(w, h, depth, spp, seed=SEED, mode=1, schedule=1,
                                           threads=orc.load().orc_hardware_threads(), window=win)
    m = opx["count"] > 0
    assert m.sum() == 160 * 90
    a = (px["total"][m] / spp).reshape(1, -1, 3)
    b = (opx["total"][m] / spp).reshape(1, -1, 3)
    rmse, lum = api.rel_rmse(a, b)
    assert rmse <= 1e-2 and abs(lum - 1) <= 5e-3, (rmse, lum)


def test_multi_device_library_split(scenes):
    """ort_multi_*: sample split + one peer reduce from a single process.  With one physical GPU the
    two contexts share it (devices=[0, 0]); with two or more, real NVLink peers are used.  The sum
    over the blocks equals the single-context render up to f32 summation order, first/last are
    those of the first / last sample, and the ray count is identical."""
    import torch

    from raytracer_odin_b200 import api

    scene = scenes("spheres_small", 96, 54)
    w, h, depth, spp = 96, 54, 6, 9
    with _renderer(scene) as r:
        one = r.render(w, h, depth, spp)
        rays_one = r.stats()["rays_closest"]
    n = torch.cuda.device_count()
    for devices in ([0], [0, 0], [0, 0, 0, 0]) + (([0, 1],) if n >= 2 else ()) + ((list(range(n)),) if n > 2 else ()):
        with api.MultiRenderer(devices, seed=SEED).upload_scene(scene) as m:
            got = m.render(w, h, depth, spp)
            st = m.stats()
        assert np.array_equal(got["count"], one["count"]), devices
        np.testing.assert_allclose(got["total"], one["total"], rtol=1e-5, atol=1e-6, err_msg=str(devices))
        np.testing.assert_allclose(got["total_squared"], one["total_squared"], rtol=1e-5, atol=1e-5)
        assert np.array_equal(got["first"], one["first"]) and np.array_equal(got["last"], one["last"]), devices
        assert st["rays_closest"] == rays_one, devices
    with pytest.raises(api.OrtError):
        api.MultiRenderer([0, 99])


def test_cli_end_to_end_and_checkpoint(scene_dir, tmp_path):
    """The reference's command line (main.odin:174-253) through the Python mirror: PPM output,
    --times accumulation, checkpoint + resume == one uninterrupted run."""
    import os

    from raytracer_odin_b200 import api, cli, scenegen

    gltf_path = scenegen.cornell(os.path.join(scene_dir, "cli_c1.gltf"))
    out = str(tmp_path / "a.ppm")
    ck = str(tmp_path / "a.npy")
    common = [gltf_path, out, "--width", "48", "--height", "32", "--ray-depth", "4", "--seed", "3"]
    cli.main(common + ["--num-samples", "4", "--checkpoint", ck])
    raw = open(out, "rb").read()
    assert raw.startswith(b"P6\n48 32\n255\n") and len(raw) == len(b"P6\n48 32\n255\n") + 48 * 32 * 3
    cli.main(common + ["--num-samples", "4", "--resume", ck, "--checkpoint", ck])
    resumed, nxt = api.load_checkpoint(ck, 48, 32)
    assert nxt == 8 and np.all(resumed["count"] == 8)
    ck2 = str(tmp_path / "b.npy")
    cli.main(common + ["--num-samples", "8", "--checkpoint", ck2])
    full, _ = api.load_checkpoint(ck2, 48, 32)
    np.testing.assert_allclose(resumed["total"], full["total"], rtol=1e-5, atol=1e-6)
    cli.main(common + ["--num-samples", "2", "--times", "3", "--gpus", "0,0", "--checkpoint", ck2])
    t3, _ = api.load_checkpoint(ck2, 48, 32)
    assert np.all(t3["count"] == 6)


def test_converged_4096spp_cornell(scenes, orc):
    """BASELINE.json north_star: converged radiance within relative RMSE <= 1 % and mean luminance
    within 0.5 % at 4096 spp (same counter-based streams on both sides)."""
    from raytracer_odin_b200 import api

    w = h = 48
    scene = scenes("cornell", w, h)
    with _renderer(scene) as r:
        px = r.render(w, h, 6, 4096)
    opx, _ = orc.OracleScene(scene).render(w, h, 6, 4096, seed=SEED, mode=1, schedule=1,
                                           threads=orc.load().orc_hardware_threads())
    assert np.all(px["count"] == 4096)
    rmse, lum = api.rel_rmse(api.mean_image(px, w, h), api.mean_image(opx, w, h))
    assert rmse <= 1e-2 and abs(lum - 1) <= 5e-3, (rmse, lum)
    print(f"4096 spp: relRMSE {rmse:.2e}, luminance ratio {lum:.6f}")


def _tiny_scene(n_tris, emissive=False):
    from raytracer_odin_b200 import cabi
    from raytracer_odin_b200.scene import Scene

    s = Scene()
    s.cam_pos = np.float32([0, 0, 3])
    s.cam_basis = np.diag(np.float32([1, 1, -1]))
    s.fov_x = 0.8
    t = np.zeros(n_tris, cabi.TRI_DTYPE)
    for i in range(n_tris):
        t["p"][i] = [-1 + 0.1 * i, -1, -0.2 * i]
        t["u"][i] = [2, 0, 0]
        t["v"][i] = [0, 2, 0]
    t["ng"] = t["n1"] = t["n2"] = t["n3"] = np.float32([0, 0, 1])
    t["material_index"] = 1
    s.triangles = t
    s.materials = np.array([((0, 0, 0), -1, (0, 0, 0), -1, 0, 0, -1, -1),
                            ((0.7, 0.6, 0.5), -1, ((2, 2, 2) if emissive else (0, 0, 0)), -1, 0.2, 0.6, -1, -1)],
                           cabi.MAT_DTYPE)
    return s


@pytest.mark.parametrize("n_tris,emissive", [(0, False), (1, False), (1, True), (3, True)])
def test_degenerate_scenes_and_odd_sizes(orc, n_tris, emissive):
    """Ragged / minimal inputs: empty scene (one empty leaf, raytracer.odin:243-254), a single
    triangle (root is a leaf), image sizes that are not multiples of the 4x4 tile, 1 spp, depth 1
    and a depth far beyond the scene's needs, zero samples."""
    from raytracer_odin_b200 import api

    scene = _tiny_scene(n_tris, emissive).finish(orc.bvh_build)
    o = orc.OracleScene(scene)
    with _renderer(scene) as r:
        for (w, h, depth, spp) in ((5, 3, 1, 1), (7, 9, 3, 2), (33, 2, 40, 3)):
            g = r.primary_hits(w, h, 0)
            ref, _, _ = o.primary_hits(w, h, sample=0, seed=SEED, mode=0)
            _hits_equal(g, ref, f"tiny {n_tris} {w}x{h}", o.ties, max_tie_frac=1.0)
            px = r.render(w, h, depth, spp)
            opx, c = o.render(w, h, depth, spp, seed=SEED, mode=0)
            assert np.array_equal(px["count"], opx["count"])
            np.testing.assert_allclose(px["total"], opx["total"], rtol=2e-4, atol=1e-6)
        z = r.render(4, 4, 3, 0)
        assert np.all(z["count"] == 0)


def test_device_bvh_build_equals_oracle(scene_dir, orc):
    """ort_bvh_build_device reproduces bvh_build (raytracer.odin:227-342) exactly —
    same post-order node array (boxes, leaf ranges, child links) and same triangle permutation as
    the oracle's builder, on meshes, soups, heavy key ties and signed zeros."""
    import os

    from raytracer_odin_b200 import cabi, gltf, scenegen
    from raytracer_odin_b200.scene import device_bvh_build
    from tests.golden.make_golden import soup

    def same(tris, what):
        a, b = tris.copy(), tris.copy()
        na, nb = orc.bvh_build(a), device_bvh_build(b)
        assert len(na) == len(nb), what
        assert na.tobytes() == nb.tobytes(), what
        for f in cabi.TRI_DTYPE.names:
            assert np.array_equal(a[f], b[f], equal_nan=True), (what, f)

    same(soup(5, 1), "n=5")
    same(soup(3000, 8), "soup 3000")
    same(soup(70000, 9), "soup 70000")
    g = soup(4096, 10)
    ix = np.arange(4096)
    g["p"] = np.stack([(ix % 16) - 8.0, ((ix // 16) % 16) - 8.0, (ix // 256) * 0.0], 1).astype(np.float32)
    g["p"][::7, 2] = -0.0
    g["u"], g["v"] = np.float32([1, 0, 0]), np.float32([0, 1, 0])
    same(g, "grid with ties and signed zeros")
    for name, kw in (("cornell", {}), ("spheres", dict(n_spheres=10, subdiv=2)), ("terrain", dict(grid=40, n_spheres=8, subdiv=1))):
        s = gltf.read_gltf(getattr(scenegen, name)(os.path.join(scene_dir, f"db_{name}.gltf"), **kw))
        same(s.triangles, name)
    assert len(device_bvh_build(np.zeros(0, cabi.TRI_DTYPE))) == 1


@pytest.mark.gpu
def test_falls_back_to_smaller_waves_when_memory_is_short(scenes):
    """When the path buffers of the default wave size do not fit (other contexts on the GPU), the render
    uses smaller waves / fewer pipelines instead of failing, with identical results (the counter-based
    streams make the wave split invisible)."""
    from raytracer_odin_b200 import api

    s = scenes("cornell")
    with api.Renderer(seed=9).upload_scene(s) as r:
        ref = r.render(64, 64, 4, 32)
    with api.Renderer(seed=9, max_path_bytes=64 * 64 * 172 * 5).upload_scene(s) as r:  # room for 5 samples in flight
        small = r.render(64, 64, 4, 32)
        again = r.render(64, 64, 4, 32)  # the fallback de-tunes one call, not the context
    assert np.array_equal(small["total"], again["total"])
    with api.Renderer(seed=9, max_path_bytes=1000).upload_scene(s) as r:  # not even one sample per pixel fits
        with pytest.raises(api.OrtError, match="out of memory"):
            r.render(64, 64, 4, 32)
    assert np.array_equal(ref["count"], small["count"])
    np.testing.assert_allclose(ref["total"], small["total"], rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_ten_million_rays_bit_exact_c2(scenes, orc):
    """SURVEY §7 step 3 gate: triangle id and the bits of t / u / v equal the faithful reference
    traversal on >= 1e7 rays of one scene (C2 at full size): three more jittered primary passes plus
    4.2 M incoherent rays started inside the scene's bounding box."""
    w, h = 1920, 1080
    scene = scenes("spheres_c2", w, h)
    o = orc.OracleScene(scene)
    threads = orc.load().orc_hardware_threads()
    total = ties = 0
    with _renderer(scene) as r:
        for sample in (1, 2, 3):
            ref, _, c = o.primary_hits(w, h, sample=sample, seed=SEED, mode=0, threads=threads)
            ties += _hits_equal(r.primary_hits(w, h, sample), ref, f"C2 primary sample {sample}", o.ties)
            assert c["stack_drops"] == 0
            total += len(ref)
        rng = np.random.default_rng(77)
        n = 4_200_000
        rays = np.zeros(n, api_mod().cabi.RAY_DTYPE)
        rays["o"] = rng.uniform(-10, 10, (n, 3)).astype(np.float32)
        d = rng.normal(size=(n, 3))
        rays["d"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        ref, c = o.trace_rays(rays, mode=0, threads=threads)
        ties += _hits_equal(r.trace_rays(rays), ref, "C2 incoherent", o.ties, max_tie_frac=1e-4)
        assert c["stack_drops"] == 0
        total += n
    assert total >= 10_000_000
    print(f"C2: {total} rays bit-identical on every tie-free ray, {ties} exact-t ties")


def api_mod():
    from raytracer_odin_b200 import api

    return api
