"""GPU parity tests added in round 2: the shading DEVICE FUNCTIONS one by one against the oracle's
(ort_probe_shading), device-resident frames, byte-exact tonemap, the reference's 64-entry stack on a deep
BVH, radiance at BASELINE sizes (C1 fixture, C4 window), C5 at full size."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 1234
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def orc():
    from oracle import binding

    return binding


def _renderer(scene, **kw):
    from raytracer_odin_b200 import api

    return api.Renderer(device=0, seed=SEED, **kw).upload_scene(scene)


def fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ulps(a, b):
    """Distance in units of the last place between two f32 arrays (same-sign finite values; 0 vs -0 = 0)."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib)


def _unit(rng, n):
    v = rng.normal(size=(n, 3))
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def _bits(u32):
    return np.asarray(u32, np.uint32).view(np.float32)


# ------------------------------------------------------------------------------------------------
# SURVEY §8(c) fixtures (4) and (5) on the GPU: the same __device__ functions k_shade calls
# ------------------------------------------------------------------------------------------------
def test_probe_shade_vndf_cosine(scenes, orc):
    """brdf_cos (shade, shading.odin:164-204), vndf_sampling / vndf_sampling_pdf (:102-137) and cosine_weighted
    (:32-39) on random and degenerate inputs.  Tolerance: <= 4 ulp where only + - * / sqrt are involved
    (vndf_sampling_pdf); relative 2e-5 through powf (shade); sampled unit vectors go through sincosf / hypotf
    (CUDA's differ from glibc's by <= 2 ulp, amplified by the cancellations downstream): 5e-5 absolute per
    component, 2e-6 on 99.5 % of them."""
    lib = orc.load()
    rng = np.random.default_rng(41)
    n = 4000
    with _renderer(scenes("cornell", 32, 32)) as r:
        # ---- shade
        N, V, L = _unit(rng, n), _unit(rng, n), _unit(rng, n)
        flipv = (np.einsum("ij,ij->i", V, N) < 0.05)
        V[flipv] = -V[flipv]
        flipl = (np.einsum("ij,ij->i", L, N) < 0.05)
        L[flipl] = -L[flipl]
        color = rng.uniform(0, 1, (n, 3)).astype(np.float32)
        met = rng.uniform(0, 1, n).astype(np.float32)
        rough = rng.uniform(0.03, 1, n).astype(np.float32)
        rough[:50] = 0.03  # the roughness floor of raytracer.odin:480
        met[:25] = 0.0
        met[25:50] = 1.0
        rec = np.concatenate([N, color, met[:, None], rough[:, None], -V, L], 1)
        got = r.probe_shading("shade", rec)
        want = np.zeros((n, 3), np.float32)
        for i in range(n):
            lib.orc_shade(fp(N[i]), fp(color[i]), met[i], rough[i], fp(np.ascontiguousarray(-V[i])), fp(L[i]), fp(want[i]))
        fin = np.isfinite(want).all(1)
        assert fin.mean() > 0.95 and np.array_equal(np.isfinite(got).all(1), fin)
        np.testing.assert_allclose(got[fin], want[fin], rtol=2e-5, atol=1e-7)

        # ---- vndf_sampling: random + n = (0,0,-1) (w == 0: the quaternion fallback, shading.odin:104-106)
        #      + omega parallel to n (Vh.xy == 0: `len == 0` tangent, :110)
        Nv, Om = _unit(rng, n), _unit(rng, n)
        flip = (np.einsum("ij,ij->i", Om, Nv) < 0.02)
        Om[flip] = -Om[flip]
        Nv[:40] = np.float32([0, 0, -1])
        Om[:40] = _unit(rng, 40) * np.float32([1, 1, 0]) + np.float32([0, 0, -0.8])
        Om[:40] /= np.linalg.norm(Om[:40], axis=1, keepdims=True)
        Nv[40:80] = np.float32([0, 0, 1])
        Om[40:80] = np.float32([0, 0, 1])
        Nv[80:120] = _unit(rng, 40)
        Om[80:120] = Nv[80:120]
        alpha = (rng.uniform(0.03, 1, n) ** 2).astype(np.float32)
        u1, u2 = rng.uniform(0, 1, n).astype(np.float32), rng.uniform(0, 1, n).astype(np.float32)
        u1[120:130] = 0.0
        u2[130:140] = 0.0
        got = r.probe_shading("vndf_sample", np.concatenate([Nv, Om, alpha[:, None], u1[:, None], u2[:, None]], 1))
        want = np.zeros((n, 3), np.float32)
        for i in range(n):
            lib.orc_vndf_sampling(fp(Nv[i]), fp(Om[i]), alpha[i], u1[i], u2[i], fp(want[i]))
        fin = np.isfinite(want).all(1)
        assert fin[:120].all(), "degenerate frames must still produce finite half vectors"
        err = np.abs(got[fin] - want[fin])
        assert err.max() <= 5e-5 and (err <= 2e-6).mean() > 0.995, (err.max(), (err <= 2e-6).mean())
        assert np.array_equal(np.isfinite(got).all(1), fin)

        # ---- vndf_sampling_pdf: only + - * / sqrt -> <= 4 ulp
        Lp = _unit(rng, n)
        got = r.probe_shading("vndf_pdf", np.concatenate([Nv, Om, alpha[:, None], Lp], 1))[:, 0]
        want = np.array([lib.orc_vndf_sampling_pdf(fp(Nv[i]), fp(Om[i]), alpha[i], fp(Lp[i])) for i in range(n)], np.float32)
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), fin) and np.array_equal(np.isnan(got), np.isnan(want))
        u = _ulps(got[fin], want[fin])
        print(f"vndf_sampling_pdf: max {u.max()} ulp over {fin.sum()} finite values")
        assert u.max() <= 4
        assert np.array_equal(got[~fin & ~np.isnan(want)], want[~fin & ~np.isnan(want)])  # same infinities

        # ---- cosine_weighted + its pdf, explicit Philox words
        r1 = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
        r2 = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
        r1[:4] = [0, 0xFFFFFFFF, 0, 0xFFFFFFFF]
        r2[:4] = [0, 0, 0xFFFFFFFF, 0xFFFFFFFF]
        got = r.probe_shading("cosine", np.concatenate([N, _bits(r1)[:, None], _bits(r2)[:, None]], 1))
        want = np.zeros((n, 3), np.float32)
        for i in range(n):
            lib.orc_cosine_weighted(fp(N[i]), int(r1[i]), int(r2[i]), fp(want[i]))
        fin = np.isfinite(want).all(1)
        err = np.abs(got[fin, :3] - want[fin])
        assert err.max() <= 5e-5 and (err <= 2e-6).mean() > 0.995, (err.max(), (err <= 2e-6).mean())
        wpdf = np.array([lib.orc_cosine_weighted_pdf(fp(N[i]), fp(np.ascontiguousarray(got[i, :3]))) for i in np.nonzero(fin)[0]], np.float32)
        assert _ulps(got[fin, 3], wpdf).max() <= 4


@pytest.mark.parametrize("name", ["cornell", "spheres_nolight"])
def test_probe_sample_and_pdf(scenes, orc, name):
    """sample (shading.odin:139-151) with explicit Philox words — all three strategies, light sampling only
    when the scene has lights — and pdf (:153-162) INCLUDING the light-BVH all-hit sum."""
    lib = orc.load()
    scene = scenes(name, 64, 64)
    o = orc.OracleScene(scene)
    rng = np.random.default_rng(43)
    n = 3000
    N, in_d = _unit(rng, n), _unit(rng, n)
    flip = np.einsum("ij,ij->i", in_d, N) > -0.05  # the incoming ray arrives against the normal
    in_d[flip] = -in_d[flip]
    pos = rng.uniform(-0.8, 0.8, (n, 3)).astype(np.float32)
    rough = rng.uniform(0.03, 1, n).astype(np.float32)
    words = rng.integers(0, 2 ** 32, (n, 4), dtype=np.uint64).astype(np.uint32)
    # strategy thresholds (t <= 0.33333, t < 0.666666): r0 at and around them
    for k, t in enumerate((0.33333, 0.333331, 0.666665, 0.666667, 0.0, 0.99999994)):
        words[k, 0] = np.uint32(int(np.float32(t) * 2 ** 24) << 8)
    with _renderer(scene) as r:
        got = r.probe_shading("sample", np.concatenate([N, pos, rough[:, None], in_d, _bits(words)], 1))
        want = np.zeros((n, 3), np.float32)
        for i in range(n):
            w4 = (C.c_uint32 * 4)(*[int(x) for x in words[i]])
            lib.orc_sample(o.ref, fp(N[i]), fp(pos[i]), rough[i], fp(in_d[i]), w4, fp(want[i]))
        fin = np.isfinite(want).all(1)
        assert fin.mean() > 0.99 and np.array_equal(np.isfinite(got).all(1), fin)
        # unit vectors through sincosf / hypotf: a few ulp of the transcendental, amplified where sphere + n nearly
        # cancels (cosine_weighted) — 5e-5 absolute on every component, 2e-6 on 99.5 % of them
        err = np.abs(got[fin] - want[fin])
        assert err.max() <= 5e-5, err.max()
        assert (err <= 2e-6).mean() > 0.995, (err <= 2e-6).mean()
        # pdf of those directions
        out_d = np.ascontiguousarray(want)
        out_d[~fin] = N[~fin]
        g = r.probe_shading("pdf", np.concatenate([N, pos, rough[:, None], in_d, out_d], 1))[:, 0]
        w = np.array([lib.orc_pdf(o.ref, fp(N[i]), fp(pos[i]), rough[i], fp(in_d[i]), fp(out_d[i])) for i in range(n)], np.float32)
        ok = np.isfinite(w)
        assert ok.mean() > 0.98 and np.array_equal(np.isfinite(g), ok)
        np.testing.assert_allclose(g[ok], w[ok], rtol=2e-5, atol=1e-9)
        if len(scene.light_triangles):
            from raytracer_odin_b200 import cabi

            rr = np.zeros(n, cabi.RAY_DTYPE)
            rr["o"], rr["d"] = pos, out_d
            lit = r.light_pdf(rr)
            assert (lit > 0).mean() > 0.05, "the light term must actually be exercised"


def _texture_scene(scenes, imgs, env=None):
    """Cornell + the given textures; material 2+i references texture i as colour (sRGB copy) and as
    metallic-roughness (raw copy), so both device copies exist."""
    import copy

    from raytracer_odin_b200 import cabi

    s = copy.copy(scenes("cornell", 32, 32))
    s.textures = list(imgs)
    mats = list(s.materials)
    for i in range(len(imgs)):
        mats.append(((1, 1, 1), i, (0, 0, 0), -1, 1.0, 1.0, i, -1))
    s.materials = np.array([tuple(m) for m in mats], cabi.MAT_DTYPE)
    s.env_map = env
    return s


def test_probe_texture_sample_and_env(scenes, orc):
    """texture_sample (textures.odin:79-135) through the texture objects: 1 / 3 / 4-channel u8 and 3-channel f32,
    raw and sRGB copies, negative uv (floored modulo), integer texel coordinates (floor == ceil), uv > 1; and the
    equirectangular environment lookup (raytracer.odin:437-446).  Bit-exact: texels are converted on the host with
    the oracle's own expressions and the bilinear weights are plain f32 arithmetic (no FMA)."""
    from raytracer_odin_b200 import cabi

    lib = orc.load()
    rng = np.random.default_rng(6)
    imgs = [rng.integers(0, 256, (5, 7, c)).astype(np.uint8) for c in (1, 3, 4)]
    imgs.append(rng.uniform(0, 4, (5, 7, 3)).astype(np.float32))
    imgs.append(rng.integers(0, 256, (64, 32, 3)).astype(np.uint8))
    env = rng.uniform(0, 8, (16, 32, 3)).astype(np.float32)
    scene = _texture_scene(scenes, imgs, env)
    uv = [(0.3, 0.6), (-0.25, -1.4), (3 / 7, 2 / 5), (1.0, 1.0), (2.7, -0.01), (0.0, 0.0), (-1.0, 5.0), (6 / 7, 4 / 5)]
    uv += [tuple(x) for x in rng.uniform(-3, 3, (400, 2))]
    uv = np.float32(uv)
    default = np.float32([1, 1, 1, 1])
    worst = 0
    with _renderer(scene) as r:
        for ti, img in enumerate(imgs):
            t = cabi.OrtTexture()
            keep = np.ascontiguousarray(img)
            t.data = keep.ctypes.data
            t.width, t.height, t.channels = img.shape[1], img.shape[0], img.shape[2]
            t.is_f32 = 1 if img.dtype == np.float32 else 0
            t.stride = img.shape[1] * img.shape[2]
            for srgb in (0, 1):
                rec = np.zeros((len(uv), 4), np.float32)
                rec[:, 0] = _bits(np.full(len(uv), ti, np.uint32))
                rec[:, 1] = srgb
                rec[:, 2:] = uv
                got = r.probe_shading("texture", rec)
                want = np.zeros((len(uv), 4), np.float32)
                for i, (u, v) in enumerate(uv):
                    lib.orc_texture_sample(C.byref(t), u, v, srgb, fp(default), fp(want[i]))
                u_ = _ulps(got, want)
                worst = max(worst, int(u_.max()))
                assert u_.max() == 0, (ti, srgb, int(u_.max()), got[u_.max(1).argmax()], want[u_.max(1).argmax()])
        # environment lookup: atan2f / asinf differ from glibc by an ulp or two -> the texel coordinate moves by
        # ~1e-6 of a texel; compare to the oracle with a tolerance scaled to the map's largest texel step
        o = orc.OracleScene(scene)
        d = _unit(rng, 2000)
        d[:6] = np.float32([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]])
        got = r.probe_shading("env", d)
        want = np.zeros((len(d), 3), np.float32)
        for i in range(len(d)):
            lib.orc_env_lookup(o.ref, fp(d[i]), fp(want[i]))
        close = np.isclose(got, want, rtol=1e-4, atol=1e-4).all(1)
        assert close.mean() > 0.995, close.mean()  # the rest straddle a texel boundary (floor / ceil flip)
    print(f"texture_sample: max {worst} ulp over {len(imgs) * 2 * len(uv)} samples")


# ------------------------------------------------------------------------------------------------
# device-resident frames, previews, byte-exact tonemap
# ------------------------------------------------------------------------------------------------
def test_tonemap_bytes_equal_the_oracle(scenes, orc):
    """get_rgb_image (output.odin:30-80): every byte of the device tonemap equals the oracle's on 2 M values
    spanning black .. overexposed (incl. the byte rounding boundaries: a dense ramp), negative totals and zeros."""
    import torch

    w, h = 2048, 1024
    rng = np.random.default_rng(8)
    total = np.concatenate([np.linspace(0, 3.0, w * h // 2), 10.0 ** rng.uniform(-6, 2, w * h // 2)]).astype(np.float32)
    total = np.stack([total, total[::-1], rng.permutation(total)], 1)
    total[:5] = [[-1, 0, 1e9], [0, 0, 0], [1e-30, 1, 2], [0.18, 0.18, 0.18], [255, 254, 3]]
    cnt = rng.integers(1, 5000, w * h).astype(np.uint32)
    from raytracer_odin_b200 import cabi

    px = np.zeros(w * h, cabi.STATS_DTYPE)
    px["total"] = total * cnt[:, None]
    px["count"] = cnt
    acc = torch.zeros(8, w * h, device="cuda", dtype=torch.float32)
    acc[:3] = torch.from_numpy(np.ascontiguousarray(px["total"].T)).cuda()
    acc[6] = torch.from_numpy((cnt & 0xFFFFF).astype(np.float32)).cuda()
    acc[7] = torch.from_numpy((cnt >> 20).astype(np.float32)).cuda()
    with _renderer(scenes("cornell", 32, 32)) as r:
        r.set_stream(torch.cuda.current_stream().cuda_stream)
        rgb = r.tonemap_rgb8(w, h, acc.data_ptr())
    ref = orc.get_rgb_image(px, w, h)
    diff = rgb.astype(int) - ref.astype(int)
    assert np.count_nonzero(diff) == 0, (np.count_nonzero(diff), np.abs(diff).max())


def test_frame_api_matches_render_and_resumes(scenes, orc):
    """ort_frame_*: accumulators stay on the device across calls.  4 + 4 samples in two calls == one 8-sample
    ort_render (same per-pixel summation order -> byte-identical); a snapshot freezes the preview while more samples
    render; frame_load (resume) + more samples == the uninterrupted frame, byte for byte."""
    scene = scenes("spheres_small", 96, 54)
    w, h, depth = 96, 54, 6
    with _renderer(scene) as r:
        ref8 = r.render(w, h, depth, 8)
        ref4 = r.render(w, h, depth, 4)
        r.frame_begin(w, h)
        assert r.frame_render(depth, 0, 4) == 4
        r.frame_snapshot()  # preview state: 4 samples
        assert r.frame_render(depth, 4, 4) == 4
        rgb4 = r.frame_preview_rgb8()
        got8 = r.frame_fetch()
        rgb8 = r.frame_preview_rgb8()
        assert got8.tobytes() == ref8.tobytes()
        assert np.array_equal(rgb4, orc.get_rgb_image(ref4, w, h))
        assert np.array_equal(rgb8, orc.get_rgb_image(ref8, w, h))
        # resume: load the 4-sample state, render the other 4
        r.frame_begin(w, h)
        r.frame_load(ref4)
        r.frame_render(depth, 4, 4)
        assert r.frame_fetch().tobytes() == ref8.tobytes()
        # interrupt already set: nothing is enqueued, the call succeeds and reports 0
        flag = np.ones(1, np.uint8)
        assert r.frame_render(depth, 8, 4, interrupt=flag) == 0 and r.last_render_samples() == 0
        # many small calls chain the wave pipelines across calls: same sums as one big call
        r.frame_begin(w, h)
        for k in range(8):
            r.frame_render(depth, k, 1, interrupt=np.zeros(1, np.uint8))
        many = r.frame_fetch()
        r.frame_end()
        assert many.tobytes() == ref8.tobytes()
        # and ort_render still works after frames (pipelines re-joined)
        assert r.render(w, h, depth, 8).tobytes() == ref8.tobytes()


def test_frame_large_chained_pipelines(scenes):
    """A frame big enough that every call spans several waves and all four pipelines: calls of 24 spp at 640x360
    (waves of 145 spp are cut to the call, so force small waves) — sums equal one ort_render."""
    scene = scenes("terrain_small", 640, 360)
    w, h, depth = 640, 360, 5
    with _renderer(scene, max_paths_in_flight=w * h * 2) as r:  # 2 spp per wave -> 12 waves per call
        ref = r.render(w, h, depth, 72)
        r.frame_begin(w, h)
        done = sum(r.frame_render(depth, 24 * k, 24) for k in range(3))
        got = r.frame_fetch()
        r.frame_end()
    assert done == 72 and np.array_equal(got["count"], ref["count"])
    assert got.tobytes() == ref.tobytes()


def test_multi_frame_api(scenes, orc):
    """ort_multi_frame_*: every GPU keeps its own accumulators; fetch / preview reduce the snapshots on devices[0]
    through peer memory.  (One physical GPU: the contexts alias it; with more, real peers are used.)"""
    import torch

    from raytracer_odin_b200 import api

    scene = scenes("spheres_small", 96, 54)
    w, h, depth, spp = 96, 54, 6, 12
    with _renderer(scene) as r:
        one = r.render(w, h, depth, spp)
        rays_one = r.stats()["rays_closest"]
    n = torch.cuda.device_count()
    for devices in ([0, 0], [0, 0, 0]) + (([0, 1],) if n >= 2 else ()) + ((list(range(n)),) if n > 2 else ()):
        with api.MultiRenderer(devices, seed=SEED).upload_scene(scene) as m:
            m.frame_begin(w, h)
            assert m.frame_render(depth, 0, 6) == 6
            m.frame_snapshot()
            assert m.frame_render(depth, 6, 6) == 6 and m.last_render_samples() == 6
            rgb6 = m.frame_preview_rgb8()
            got = m.frame_fetch()
            rgb12 = m.frame_preview_rgb8()
            st = m.stats()
            m.frame_end()
        assert np.array_equal(got["count"], one["count"]), devices
        np.testing.assert_allclose(got["total"], one["total"], rtol=1e-5, atol=1e-6, err_msg=str(devices))
        np.testing.assert_allclose(got["total_squared"], one["total_squared"], rtol=1e-5, atol=1e-5)
        assert np.array_equal(got["first"], one["first"]) and np.array_equal(got["last"], one["last"]), devices
        assert st["rays_closest"] == rays_one, devices
        assert np.array_equal(rgb12, orc.get_rgb_image(got, w, h))
        assert np.abs(rgb6.astype(int) - rgb12.astype(int)).max() > 0  # the snapshot really is the earlier state


def test_count_plane_stays_exact_beyond_2_pow_24(scenes):
    """Sample_Stats.count is a u32 (main.odin:36): the planar count (lo + 2^20 * hi) stays exact when a caller
    keeps accumulating past 2^24 samples per pixel (depth 0: samples are counted, nothing is traced)."""
    import torch

    w = h = 4
    with _renderer(scenes("cornell", 32, 32), max_paths_in_flight=1 << 22) as r:
        acc = torch.zeros(8, w * h, device="cuda", dtype=torch.float32)
        r.set_stream(torch.cuda.current_stream().cuda_stream)
        big = (1 << 24) + 12345
        r.render_device(w, h, 0, 0, big, acc.data_ptr())
        r.render_device(w, h, 0, big, 3, acc.data_ptr())
        torch.cuda.synchronize()
        px = r.unpack_accum(w, h, acc.data_ptr())
    assert np.all(px["count"] == big + 3)


# ------------------------------------------------------------------------------------------------
# the reference's 64-entry stack (raytracer.odin:379) on a deep BVH
# ------------------------------------------------------------------------------------------------
def deep_scene(n=150, ratio=1.5, s0=1e-12):
    """Triangles in a geometric progression of size and position: the SAH builder peels a few of the largest off
    at every split, so the binary tree is a chain of depth ~n/4.5 (33 branches for n = 150)."""
    from oracle import binding as orc
    from raytracer_odin_b200 import cabi, gltf

    t = np.zeros(n, cabi.TRI_DTYPE)
    s = s0 * ratio ** np.arange(n)
    t["p"] = np.stack([1.5 * s, s, s], 1).astype(np.float32)
    t["u"] = np.stack([0 * s, 2 * s, 0 * s], 1).astype(np.float32)
    t["v"] = np.stack([0.3 * s, 0 * s, 2 * s], 1).astype(np.float32)
    ng = np.cross(t["u"].astype(np.float64), t["v"].astype(np.float64))
    t["ng"] = (ng / np.linalg.norm(ng, axis=1, keepdims=True)).astype(np.float32)
    t["n1"] = t["n2"] = t["n3"] = t["ng"]
    t["material_index"] = 1
    sc = gltf.Scene()
    sc.triangles = t
    sc.materials = np.array([((0, 0, 0), -1, (0, 0, 0), -1, 0, 0, -1, -1), ((0.8, 0.8, 0.8), -1, (0, 0, 0), -1, 0, 1, -1, -1)],
                            cabi.MAT_DTYPE)
    sc.finish(orc.bvh_build)
    return sc, s


def test_deep_bvh_reference_stack_need(orc):
    """A BVH whose worst-case need of the REFERENCE's stack (2 * branch depth + 1 = 67) exceeds its 64 entries.
    The library reports that at upload (ort_stats.reference_stack_need), never drops a push itself, and returns
    the exact closest hit: bit-identical to the duplicate-free oracle on every tie-free ray, and to the faithful
    oracle on every ray on which the reference dropped no push.  (Filling 64 entries takes 32 nested both-children-
    hit branches on the reference's left-first path, and its duplicate-left push doubles the work at each of them
    for rays that find no early hit: on this scene the faithful oracle already needs ~3e5 node pops per ray at a
    high-water of 39.)"""
    from raytracer_odin_b200 import cabi

    sc, s = deep_scene()
    rng = np.random.default_rng(1)
    m = 1500
    j = rng.integers(0, len(s), m)
    axis = np.array([1.65, 2, 2])
    rays = np.zeros(m, cabi.RAY_DTYPE)
    rays["o"] = (axis[None, :] * (0.55 * s[j])[:, None] * (1 + rng.uniform(-0.05, 0.05, (m, 3)))).astype(np.float32)
    d = axis[None, :] * rng.choice([-1.0, 1.0], m)[:, None] + rng.uniform(-0.25, 0.25, (m, 3))
    rays["d"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    o = orc.OracleScene(sc)
    threads = orc.load().orc_hardware_threads()
    ideal, ci = o.trace_rays(rays, mode=1, threads=threads)
    ideal_flags = o.ties.copy()
    faithful, cf = o.trace_rays(rays, mode=0, threads=threads)
    flags = o.ties.copy()
    with _renderer(sc) as r:
        g = r.trace_rays(rays)
        st = r.stats()
    assert st["reference_stack_need"] == 67 and st["reference_stack_need"] > 64
    assert st["wide_max_stack"] <= 128
    assert ci["stack_drops"] == 0
    free = ideal_flags == 0
    for f in ("tri", "inside"):
        assert np.array_equal(g[f][free], ideal[f][free])
    assert np.array_equal(g["t"].view(np.uint32), ideal["t"].view(np.uint32))
    kept = (flags & 2) == 0  # rays on which the reference dropped no push
    assert np.array_equal(g["tri"][kept & (flags == 0)], faithful["tri"][kept & (flags == 0)])
    assert np.array_equal(g["t"][kept].view(np.uint32), faithful["t"][kept].view(np.uint32))
    differ = g["tri"] != faithful["tri"]
    assert not (differ & kept & (flags == 0)).any()
    print(f"deep BVH: reference stack need {st['reference_stack_need']}, faithful high-water {cf['stack_high']}, "
          f"{cf['stack_drops']} dropped pushes on {(~kept).sum()} rays, {differ.sum()} hits differ from the faithful oracle; "
          f"faithful {cf['node_pops'] / m:.0f} vs ideal {ci['node_pops'] / m:.0f} node pops per ray; hit fraction {(g['tri'] >= 0).mean():.2f}")


def test_zero_and_tiny_direction_components(scenes, orc):
    """Rays with a direction component that is exactly 0, -0, denormal or tiny (an axis-aligned camera produces a few
    exact zeros per wave by cancellation in pixel_to_ray_dir): the reference's slab test divides by it and gets an
    unbounded or an empty interval; the library clamps |d| for its box tests (kernels.cuh make_ray).  Hits must
    equal the faithful oracle bit for bit, and such a ray must not degenerate into a walk over the whole tree
    (measured before the fix: 1.2-2.0 ms for ONE ray on the 1 M-triangle scene)."""
    from raytracer_odin_b200 import cabi

    scene = scenes("terrain_c4", 1920, 1080)
    o = orc.OracleScene(scene)
    rng = np.random.default_rng(12)
    n = 6000
    rays = np.zeros(n, cabi.RAY_DTYPE)
    rays["o"] = rng.uniform(-40, 40, (n, 3)).astype(np.float32)
    rays["o"][:, 1] = rng.uniform(-3, 20, n).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    special = np.float32([0.0, -0.0, 1e-30, -1e-30, 1e-42, -1e-42, 1e-19, -1e-19, 1e-17, -1e-17])
    for k in range(n):
        ax = k % 3
        d[k, ax] = special[(k // 3) % len(special)]
        if k % 7 == 0:  # two degenerate components
            d[k, (ax + 1) % 3] = special[(k // 21) % len(special)]
        rest = [a for a in range(3) if abs(d[k, a]) > 1e-10]
        d[k, rest] /= np.float32(np.linalg.norm(d[k, rest]))
    rays["d"] = d
    # the camera-like cases of the benchmark scene: origin on the axis, component exactly zero
    rays["o"][:8] = np.float32([0.0, 16.0, 62.0])
    rays["d"][0] = np.float32([0.0, -0.5062409043312073, -0.8623921275138855])
    rays["d"][1] = np.float32([0.35083362460136414, 0.0, -0.9364379048347473])
    rays["d"][2] = np.float32([-0.0, -0.5062409043312073, -0.8623921275138855])
    rays["d"][3] = np.float32([0.35083362460136414, -0.0, -0.9364379048347473])
    ref, c = o.trace_rays(rays, mode=0, threads=orc.load().orc_hardware_threads())
    flags = o.ties.copy()
    with _renderer(scene) as r:
        g = r.trace_rays(rays)
        ms_special = r.bench_trace(rays[:4], 0, 3)
        generic = rays[:4].copy()
        generic["d"] = np.float32([[0.01, -0.506, -0.8624], [0.35, 0.01, -0.9364], [-0.01, -0.506, -0.8624], [0.35, -0.01, -0.9364]])
        ms_generic = r.bench_trace(generic, 0, 3)
    _hits_equal_flags(g, ref, flags, "zero direction components", max_tie_frac=1e-2)
    assert (ref["tri"] >= 0).mean() > 0.2
    print(f"4 camera rays with an exactly-zero component: {ms_special:.3f} ms per launch; 4 generic neighbours: {ms_generic:.3f} ms")
    assert ms_special < 10 * ms_generic + 0.05, (ms_special, ms_generic)


# ------------------------------------------------------------------------------------------------
# radiance at BASELINE sizes
# ------------------------------------------------------------------------------------------------
def test_c1_at_baseline_size_against_committed_fixture(scenes):
    """BASELINE config C1 as quoted: 256x256, ray-depth 6, 64 spp, against the oracle image committed under
    tests/golden (cornell_c1_256.npz, written by tests/golden/make_golden.py with the same seed).  relRMSE <= 1 %,
    mean luminance within 0.5 % (north_star); same streams, so almost every pixel agrees to f32 noise."""
    from raytracer_odin_b200 import api

    g = np.load(os.path.join(GOLD, "cornell_c1_256.npz"))
    w = h = 256
    scene = scenes("cornell", w, h)
    with api.Renderer(device=0, seed=int(g["seed"])).upload_scene(scene) as r:
        px = r.render(w, h, 6, 64)
        st = r.stats()
    assert np.all(px["count"] == 64)
    a = api.mean_image(px, w, h)
    b = (g["total"].astype(np.float32) / np.float32(64)).reshape(h, w, 3)
    rmse, lum = api.rel_rmse(a, b)
    print(f"C1 256x256 / 6 / 64 spp: relRMSE {rmse:.2e}, luminance ratio {lum:.6f}")
    assert rmse <= 1e-2 and abs(lum - 1) <= 5e-3, (rmse, lum)
    assert np.isclose(a, b, rtol=1e-3, atol=1e-5).all(axis=2).mean() > 0.97
    assert abs(st["rays_closest"] - int(g["n_rays"][0])) <= 1e-3 * int(g["n_rays"][0])


def test_full_size_c4_window_radiance(scenes, orc):
    """BASELINE config 4 (1 M triangles, 20 480 emissive, ray-depth 10) at 1920x1080: a 96x54 window rendered by the
    oracle with the same streams; relRMSE <= 1e-2, luminance within 0.5 % (raytracer.odin:432-518 at depth 10)."""
    from raytracer_odin_b200 import api

    w, h, depth, spp = 1920, 1080, 10, 8
    scene = scenes("terrain_c4", w, h)
    win = (912, 540, 1008, 594)
    with _renderer(scene) as r:
        px = r.render(w, h, depth, spp)
    opx, c = orc.OracleScene(scene).render(w, h, depth, spp, seed=SEED, mode=1, schedule=1,
                                           threads=orc.load().orc_hardware_threads(), window=win)
    m = opx["count"] > 0
    assert m.sum() == 96 * 54
    a = (px["total"][m] / spp).reshape(1, -1, 3)
    b = (opx["total"][m] / spp).reshape(1, -1, 3)
    rmse, lum = api.rel_rmse(a, b)
    print(f"C4 window: relRMSE {rmse:.2e}, luminance ratio {lum:.6f}, {c['rays']} oracle rays")
    assert rmse <= 1e-2 and abs(lum - 1) <= 5e-3, (rmse, lum)
    assert (b.sum(axis=2) > 0).mean() > 0.25, "the window must see lit geometry"


def test_full_size_c5_hits_and_stack(scene_dir, orc):
    """BASELINE config 5 (10 M triangles, 3840x2160, ray-depth 12): 200 k of the GPU's own primary rays and 100 k
    bounce rays (cosine-distributed around the primary hit normals) re-traced by the faithful oracle agree bit for
    bit on tie-free rays; the reference's stack high-water and drops on those rays are reported next to its
    worst-case need (2 * branch depth + 1)."""
    from raytracer_odin_b200 import api, cabi, gltf, scenegen
    from raytracer_odin_b200.scene import device_bvh_build

    cfg = scenegen.CONFIGS["C5"]
    w, h = cfg["width"], cfg["height"]
    path, _ = scenegen.generate("C5", os.path.join(scene_dir, "c5"))
    scene = gltf.read_gltf(path)
    scene.fov_x = scene.apply_render_config(w, h)
    scene.finish(device_bvh_build)  # byte-identical to the oracle's builder (test_device_bvh_build_equals_oracle)
    assert len(scene.triangles) > 9_000_000
    o = orc.OracleScene(scene)
    threads = orc.load().orc_hardware_threads()
    rng = np.random.default_rng(5)
    with _renderer(scene) as r:
        st = r.stats()
        g, rays = r.primary_hits(w, h, 2, want_rays=True)
        pick = rng.choice(w * h, 200_000, replace=False)
        ref, c = o.trace_rays(rays[pick], mode=0, threads=threads)
        flags = o.ties.copy()
        n_ties = _hits_equal_flags(g[pick], ref, flags, "C5 primary")
        # bounce rays from the primary hit points
        hit = np.nonzero(g["tri"] >= 0)[0]
        sel = rng.choice(hit, 100_000, replace=False)
        P = rays["o"][sel] + rays["d"][sel] * g["t"][sel][:, None]
        ng = scene.triangles["ng"][g["tri"][sel]]
        ng = np.where((np.einsum("ij,ij->i", ng, rays["d"][sel]) > 0)[:, None], -ng, ng)
        dd = ng + _unit(rng, len(sel)) * np.float32(0.999)
        b = np.zeros(len(sel), cabi.RAY_DTYPE)
        b["o"] = P.astype(np.float32)
        b["d"] = (dd / np.linalg.norm(dd, axis=1, keepdims=True)).astype(np.float32)
        gb = r.trace_rays(b)
        refb, cb = o.trace_rays(b, mode=0, threads=threads)
        n_ties += _hits_equal_flags(gb, refb, o.ties.copy(), "C5 bounce")
    drops = c["stack_drops"] + cb["stack_drops"]
    high = max(c["stack_high"], cb["stack_high"])
    print(f"C5: {len(scene.triangles)} triangles, wide depth {st['wide_depth']}, library stack need {st['wide_max_stack']}; reference "
          f"stack worst case {st['reference_stack_need']} (64 entries), high-water on 300k rays {high}, dropped pushes {drops}; "
          f"{n_ties} exact-t ties")
    assert st["wide_max_stack"] <= 128
    assert high <= 64 and drops == 0, "the reference dropped pushes on real C5 rays: compare with the flags"


def _hits_equal_flags(g, o, flags, what, max_tie_frac=1e-4):
    """Bit-exact on every ray without an exact-t tie (flag bit 0) and without a reference stack drop (bit 1)."""
    free = flags == 0
    n_ties = int(((flags & 1) > 0).sum())
    assert n_ties <= max_tie_frac * len(o), f"{what}: {n_ties} ties"
    for f in ("tri", "material", "inside"):
        assert np.array_equal(g[f][free], o[f][free]), f"{what}: {f} differs"
    kept = (flags & 2) == 0
    assert np.array_equal((g["tri"] >= 0)[kept], (o["tri"] >= 0)[kept]), f"{what}: hit/miss differs"
    hit = (o["tri"] >= 0) & kept
    assert np.array_equal(g["t"][hit].view(np.uint32), o["t"][hit].view(np.uint32)), f"{what}: t bits differ"
    for f in ("u", "v"):
        mm = hit & free
        assert np.array_equal(g[f][mm].view(np.uint32), o[f][mm].view(np.uint32)), f"{what}: {f} bits differ"
    return n_ties
