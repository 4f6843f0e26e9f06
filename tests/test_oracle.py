"""CPU tests of the oracle (oracle/oracle.cpp): against the committed golden fixtures (regression
pins — the reference ships none), against published known answers where one exists
(Philox), and against independent float64 restatements of the same formulas."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import binding as orc
from raytracer_odin_b200 import cabi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PF = C.POINTER(C.c_float)


def fp(a):
    return np.ascontiguousarray(a, np.float32).ctypes.data_as(PF)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    lib = orc.load()
    out = (C.c_uint32 * 4)()
    lib.orc_philox(0, 0, 0, 0, out)
    assert list(out) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    lib.orc_philox(0xFFFFFFFF, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFFFFFFFFFF, out)
    assert list(out) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    lib.orc_philox(0x243F6A88, 0x13198A2E85A308D3, 0x03707344, 0x299F31D0A4093822, out)
    assert list(out) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def _tri(p, u, v):
    t = np.zeros(1, cabi.TRI_DTYPE)
    t["p"], t["u"], t["v"] = p, u, v
    n = np.cross(np.asarray(u, float), np.asarray(v, float))
    t["ng"] = n / np.linalg.norm(n)
    return t


def test_intersect_ray_triangle_matches_float64():
    lib = orc.load()
    rng = np.random.default_rng(0)
    out = np.zeros(4, np.float32)
    n_hit = 0
    for _ in range(3000):
        p, u, v = rng.normal(size=3), rng.normal(size=3), rng.normal(size=3)
        o = rng.normal(size=3) * 2
        tgt = p + u * rng.uniform(-0.3, 1.3) + v * rng.uniform(-0.3, 1.3)
        d = tgt - o
        d /= np.linalg.norm(d)
        t = _tri(p, u, v)
        lib.orc_intersect_ray_triangle(fp(o), fp(d), t.ctypes.data, fp(out))
        o32, d32 = np.float32(o).astype(float), np.float32(d).astype(float)
        A = np.stack([t["u"][0].astype(float), t["v"][0].astype(float), -d32], axis=1)
        uu, vv, tt = np.linalg.solve(A, o32 - t["p"][0].astype(float))
        inside = uu >= 0 and vv >= 0 and uu + vv <= 1
        margin = min(uu, vv, 1 - uu - vv)
        if abs(margin) < 1e-4:
            continue  # on an edge: f32 may fall either way
        if inside:
            n_hit += 1
            assert out[0] != -1
            np.testing.assert_allclose(out[:3], [tt, uu, vv], rtol=2e-3, atol=2e-4)
            assert bool(out[3]) == (np.dot(t["ng"][0].astype(float), d32) > 0)
        else:
            assert out[0] == -1
    assert n_hit > 500


def test_intersect_edge_cases():
    lib = orc.load()
    out = np.zeros(4, np.float32)
    t = _tri([0, 0, 0], [1, 0, 0], [0, 1, 0])
    # front / back face
    lib.orc_intersect_ray_triangle(fp([0.25, 0.25, 1]), fp([0, 0, -1]), t.ctypes.data, fp(out))
    assert out[0] == 1 and out[1] == 0.25 and out[2] == 0.25 and out[3] == 0
    lib.orc_intersect_ray_triangle(fp([0.25, 0.25, -1]), fp([0, 0, 1]), t.ctypes.data, fp(out))
    assert out[0] == 1 and out[3] == 1
    # behind the origin: negative t is returned as is (the caller filters t > 0, raytracer.odin:360)
    lib.orc_intersect_ray_triangle(fp([0.25, 0.25, 1]), fp([0, 0, 1]), t.ctypes.data, fp(out))
    assert out[0] == -1 or out[0] < 0
    # outside: t = -1
    lib.orc_intersect_ray_triangle(fp([2, 2, 1]), fp([0, 0, -1]), t.ctypes.data, fp(out))
    assert out[0] == -1
    # parallel ray: det = 0 -> NaN/inf comparisons are all false -> not rejected, t is non-finite
    lib.orc_intersect_ray_triangle(fp([0.25, 0.25, 1]), fp([1, 0, 0]), t.ctypes.data, fp(out))
    assert not np.isfinite(out[0]) or out[0] == -1


def test_check_intersect_ray_aabb():
    lib = orc.load()
    rng = np.random.default_rng(1)
    t = C.c_float()
    agree = total = 0
    for _ in range(4000):
        lo = rng.uniform(-1, 0, 3)
        hi = lo + rng.uniform(0.1, 1.5, 3)
        o = rng.uniform(-3, 3, 3)
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        got = lib.orc_check_intersect_ray_aabb(fp(o), fp(d), fp(lo), fp(hi), np.float32(np.inf), C.byref(t))
        t1 = (lo - o) / d
        t2 = (hi - o) / d
        tn, tf = np.minimum(t1, t2).max(), np.maximum(t1, t2).min()
        want = tn <= tf and tf >= 0
        if abs(tn - tf) < 1e-4 or abs(tf) < 1e-4:
            continue
        total += 1
        agree += int(bool(got) == bool(want))
        if got and want:
            assert abs(t.value - max(tn, 0)) < 1e-3
    assert agree == total and total > 3000
    # origin inside: distance 0; zero-extent (flat) box of an axis-aligned triangle still hit
    assert lib.orc_check_intersect_ray_aabb(fp([0, 0, 0]), fp([0, 0, 1]), fp([-1, -1, -1]), fp([1, 1, 1]), np.float32(np.inf), C.byref(t))
    assert t.value == 0
    assert lib.orc_check_intersect_ray_aabb(fp([0.2, 0.3, 2]), fp([0, 0, -1]), fp([0, 0, 0]), fp([1, 1, 0]), np.float32(np.inf), C.byref(t))
    assert t.value == 2
    # sphere pre-cull (raytracer.odin:122): box farther than max_dist is rejected
    assert not lib.orc_check_intersect_ray_aabb(fp([0, 0, 10]), fp([0, 0, -1]), fp([-1, -1, -1]), fp([1, 1, 1]), np.float32(5.0), C.byref(t))
    assert not lib.orc_check_intersect_ray_aabb(fp([0, 0, 3]), fp([0, 0, 1]), fp([-1, -1, -1]), fp([1, 1, 1]), np.float32(np.inf), C.byref(t))


def test_bvh_build_golden_and_invariants():
    g = np.load(os.path.join(GOLD, "bvh_kat.npz"))
    tris = g["tris_in"].copy()
    nodes = orc.bvh_build(tris)
    assert nodes.tobytes() == g["nodes"].tobytes()
    for f in ("p", "u", "v"):
        assert np.array_equal(tris[f], g["tris_out"][f])
    # invariants of raytracer.odin:227-342: post-order, root last, leaves <= 4, every triangle once
    covered = np.zeros(len(tris), int)
    for i, nd in enumerate(nodes):
        if nd["kind"] == 0:
            assert 1 <= nd["b"] <= 4
            covered[nd["a"]:nd["a"] + nd["b"]] += 1
            pts = np.concatenate([tris["p"][nd["a"]:nd["a"] + nd["b"]],
                                  tris["p"][nd["a"]:nd["a"] + nd["b"]] + tris["u"][nd["a"]:nd["a"] + nd["b"]],
                                  tris["p"][nd["a"]:nd["a"] + nd["b"]] + tris["v"][nd["a"]:nd["a"] + nd["b"]]])
            assert np.all(pts >= nd["lo"]) and np.all(pts <= nd["hi"])
        else:
            assert nd["a"] < i and nd["b"] < i
            for ch in (nd["a"], nd["b"]):
                assert np.all(nodes[ch]["lo"] >= nd["lo"]) and np.all(nodes[ch]["hi"] <= nd["hi"])
    assert np.all(covered == 1) and nodes[-1]["kind"] == 1
    # empty input: one empty leaf with AABB_EMPTY (light BVH of a scene without lights)
    e = orc.bvh_build(np.zeros(0, cabi.TRI_DTYPE))
    assert len(e) == 1 and e[0]["kind"] == 0 and e[0]["b"] == 0 and np.all(np.isinf(e[0]["lo"]))


def test_traversal_golden_and_orders_agree(scenes):
    g = np.load(os.path.join(GOLD, "soup_trace.npz"))
    from tests.golden.make_golden import soup
    from raytracer_odin_b200 import gltf

    sc = gltf.Scene()
    sc.triangles = soup(5000, 2)
    sc.materials = np.array([((0, 0, 0), -1, (0, 0, 0), -1, 0, 0, -1, -1), ((0.8, 0.8, 0.8), -1, (0, 0, 0), -1, 0, 1, -1, -1)],
                            cabi.MAT_DTYPE)
    sc.finish(orc.bvh_build)
    o = orc.OracleScene(sc)
    faithful, cf = o.trace_rays(g["rays"], mode=0)
    ties = o.ties.copy()
    assert faithful.tobytes() == g["hits"].tobytes()
    ideal, ci = o.trace_rays(g["rays"], mode=1)
    free = ties == 0
    assert np.array_equal(faithful["tri"][free], ideal["tri"][free])
    assert np.array_equal(faithful["t"].view(np.uint32), ideal["t"].view(np.uint32))
    # the reference's duplicate-left push (raytracer.odin:395-409) costs many times the node pops
    assert cf["node_pops"] > 3 * ci["node_pops"] and cf["stack_drops"] == 0
    # brute force over all triangles == BVH traversal on a subset
    lib = orc.load()
    out = np.zeros(4, np.float32)
    for i in range(0, 300):
        r = g["rays"][i]
        oo = (r["o"] + r["d"] * np.float32(1e-3)).astype(np.float32)
        best, bi = np.inf, -1
        for k in range(len(sc.triangles)):
            lib.orc_intersect_ray_triangle(fp(oo), fp(r["d"]), sc.triangles[k:k + 1].ctypes.data, fp(out))
            if out[0] > 0 and out[0] < best:
                best, bi = out[0], k
        if not ties[i]:
            assert bi == faithful["tri"][i]


def test_cornell_golden(scenes):
    g = np.load(os.path.join(GOLD, "cornell_64.npz"))
    s = scenes("cornell", 64, 64)
    o = orc.OracleScene(s)
    hits, rays, _ = o.primary_hits(64, 64, sample=3, seed=99, mode=0)
    assert hits.tobytes() == g["hits"].tobytes() and rays.tobytes() == g["rays"].tobytes()
    px, c = o.render(64, 64, 6, 8, seed=99, mode=0, schedule=1, threads=1)
    assert np.array_equal(px["count"], g["count"]) and c["rays"] == int(g["n_rays"][0])
    np.testing.assert_allclose(px["total"], g["total"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(px["first"], g["first"], rtol=1e-6, atol=1e-7)
    # the multithreaded checker schedule gives the same per-pixel sums (sample order preserved)
    px4, _ = o.render(64, 64, 6, 8, seed=99, mode=0, schedule=1, threads=4)
    assert px4.tobytes() == px.tobytes()
    # ideal traversal order renders the same image
    pxi, _ = o.render(64, 64, 6, 8, seed=99, mode=1, schedule=1, threads=4)
    np.testing.assert_allclose(pxi["total"], px["total"], rtol=1e-5, atol=1e-6)


def _shade64(n, color, metallic, roughness, in_d, L):
    n, color, in_d, L = (np.asarray(x, float) for x in (n, color, in_d, L))
    a2 = roughness ** 4
    V = -in_d
    H = (L + V) / np.linalg.norm(L + V)
    fb = (1 - H @ L) ** 5
    fds = 0.04 + 0.96 * fb
    fm = color + (1 - color) * fb
    hn = H @ n
    D = a2 * (0.0 if hn < 0 else 1.0) / (np.pi * ((a2 - 1) * hn * hn + 1) ** 2)

    def g(x):
        c = n @ x
        return 2 * max(c, 0) / (c + np.sqrt(a2 + (1 - a2) * c * c))

    ct = D * g(L) * g(V) / (4 * (V @ n))
    spec = ct * np.ones(3)
    diff = color * max(L @ n, 0) / np.pi
    diel = diff * (1 - fds) + spec * fds
    return diel * (1 - metallic) + spec * fm * metallic


def test_shade_and_pdfs_match_float64():
    lib = orc.load()
    rng = np.random.default_rng(4)
    out = np.zeros(3, np.float32)
    for _ in range(2000):
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        V = rng.normal(size=3)
        V /= np.linalg.norm(V)
        if V @ n < 0.05:
            V = -V
        if V @ n < 0.05:
            continue
        L = rng.normal(size=3)
        L /= np.linalg.norm(L)
        if L @ n < 0.05:
            L = -L
        if L @ n < 0.05:
            continue
        color, metallic, rough = rng.uniform(0, 1, 3), rng.uniform(), rng.uniform(0.1, 1)
        n32, V32, L32, c32 = (np.float32(x) for x in (n, V, L, color))
        lib.orc_shade(fp(n32), fp(c32), np.float32(metallic), np.float32(rough), fp(-V32), fp(L32), fp(out))
        want = _shade64(n32, c32, float(np.float32(metallic)), float(np.float32(rough)), -V32.astype(float), L32)
        np.testing.assert_allclose(out, want, rtol=2e-3, atol=1e-6)
        assert abs(lib.orc_cosine_weighted_pdf(fp(n32), fp(L32)) - max(n32.astype(float) @ L32.astype(float) / np.pi, 0)) < 1e-6


def test_sample_pdf_consistency_and_energy(scenes):
    """sample() draws from the density pdf() reports (shading.odin:139-162): for a lobe-covering
    test function f, E[f(w)/pdf(w)] over sampled w equals the integral of f over the sphere; and
    the one-sample estimator E[shade/pdf] stays <= ~1 (no energy gain) for a white surface."""
    lib = orc.load()
    s = scenes("spheres_nolight")  # no lights: cosine 1/3, VNDF 2/3
    o = orc.OracleScene(s)
    rng = np.random.default_rng(5)
    n = np.float32([0.0, 0.0, 1.0])
    pos = np.float32([0, 50, 0])  # far above the scene
    in_d = np.float32([0.6, 0.0, -0.8])
    out = np.zeros(3, np.float32)
    sh = np.zeros(3, np.float32)
    white = np.float32([1, 1, 1])
    for rough in (0.3, 0.8):
        acc_one, acc_energy, m = 0.0, 0.0, 20000
        r = (C.c_uint32 * 4)()
        for i in range(m):
            for k in range(4):
                r[k] = int(rng.integers(0, 2 ** 32))
            lib.orc_sample(o.ref, fp(n), fp(pos), np.float32(rough), fp(in_d), r, fp(out))
            p = lib.orc_pdf(o.ref, fp(n), fp(pos), np.float32(rough), fp(in_d), fp(out))
            if out[2] <= 0 or not p > 0:
                continue
            acc_one += max(out[2], 0) / np.pi / p  # integral of cos/pi over the hemisphere = 1
            lib.orc_shade(fp(n), fp(white), np.float32(0.0), np.float32(rough), fp(in_d), fp(out), fp(sh))
            acc_energy += float(sh.mean()) / p
        assert abs(acc_one / m - 1.0) < 0.05, (rough, acc_one / m)
        assert acc_energy / m < 1.05, (rough, acc_energy / m)


def _tex(img):
    img = np.ascontiguousarray(img)
    t = cabi.OrtTexture()
    t.data = img.ctypes.data
    t.width, t.height, t.channels = img.shape[1], img.shape[0], img.shape[2]
    t.is_f32 = 1 if img.dtype == np.float32 else 0
    t.stride = img.shape[1] * img.shape[2]
    return t, img


def _sample64(img, u, v, srgb):
    h, w, c = img.shape

    def texel(x, y):
        px = np.ones(4)
        px[:c] = img[y, x].astype(float) / (255.0 if img.dtype == np.uint8 else 1.0)
        if srgb:
            px[:3] = px[:3] ** 2.2
        return px

    pcx, pcy = np.float32(u) * np.float32(w), np.float32(v) * np.float32(h)
    lx, ly, hx, hy = np.floor(pcx), np.floor(pcy), np.ceil(pcx), np.ceil(pcy)
    tx, ty = float(pcx - lx), float(pcy - ly)
    x0, y0, x1, y1 = int(lx) % w, int(ly) % h, int(hx) % w, int(hy) % h
    a = texel(x0, y0) * (1 - ty) + texel(x0, y1) * ty
    b = texel(x1, y0) * (1 - ty) + texel(x1, y1) * ty
    return a * (1 - tx) + b * tx


@pytest.mark.parametrize("channels,dtype", [(1, np.uint8), (3, np.uint8), (4, np.uint8), (3, np.float32)])
def test_texture_sample(channels, dtype):
    lib = orc.load()
    rng = np.random.default_rng(6)
    img = rng.integers(0, 256, (5, 7, channels)).astype(np.uint8) if dtype == np.uint8 else rng.uniform(0, 4, (5, 7, channels)).astype(np.float32)
    t, keep = _tex(img)
    out = np.zeros(4, np.float32)
    default = np.float32([1, 1, 1, 1])
    # negative uv wraps by floored modulo, integer texel coords (floor == ceil), > 1 wraps
    for (u, v) in [(0.3, 0.6), (-0.25, -1.4), (3 / 7, 2 / 5), (1.0, 1.0), (2.7, -0.01), (0.0, 0.0)] + [tuple(x) for x in rng.uniform(-2, 2, (50, 2))]:
        for srgb in (0, 1):
            lib.orc_texture_sample(C.byref(t), np.float32(u), np.float32(v), srgb, fp(default), fp(out))
            np.testing.assert_allclose(out, _sample64(keep, u, v, bool(srgb)), rtol=1e-4, atol=1e-5)
    # nil sampler returns the caller's default (textures.odin:110-112)
    lib.orc_texture_sample(None, 0.5, 0.5, 0, fp(np.float32([0.5, 1.0, 0.5, 0.0])), fp(out))
    assert list(out) == [0.5, 1.0, 0.5, 0.0]


def test_pixel_to_ray_dir_and_tonemap(scenes):
    lib = orc.load()
    s = scenes("cornell", 64, 64)
    cs, keep = s.to_c()
    M = np.zeros(16, np.float32)
    lib.orc_pixel_to_ray_dir(C.byref(cs.cam), 64, 64, fp(M))
    M = M.reshape(4, 4)
    # centre pixel looks along the camera's forward axis (basis column 2 = -Z of the node)
    d = M @ np.array([32, 32, 0, 1], np.float32)
    d = d[:3] / np.linalg.norm(d[:3])
    np.testing.assert_allclose(d, s.cam_basis[:, 2], atol=1e-6)
    # corner pixel: tan(fov_x / 2) along x, tan_x / aspect along y
    c = M @ np.array([64, 64, 0, 1], np.float32)
    tx = np.tan(np.float32(s.fov_x) / 2)
    np.testing.assert_allclose(c[:3], s.cam_basis @ np.array([tx, tx, 1], np.float32), rtol=1e-5)
    # get_rgb_image (output.odin:30-80) against the numpy mirror
    from raytracer_odin_b200 import output

    px = np.zeros(16, cabi.STATS_DTYPE)
    px["count"] = 4
    px["total"] = np.random.default_rng(7).uniform(0, 20, (16, 3)).astype(np.float32)
    px["total"][0] = [-1, 0, 1e9]
    a = orc.get_rgb_image(px, 4, 4)
    b = output.get_rgb_image(px, 4, 4)
    assert np.abs(a.astype(int) - b.astype(int)).max() <= 1


def test_strided_tile_sample_is_a_subset_of_the_frame(scenes):
    """orc_render_strided (bench.py's bounded CPU sample): every stride-th 4x4 tile in x and y, spread over the
    whole frame; the rendered pixels equal the full render's, all others stay untouched."""
    s = scenes("cornell", 40, 24)
    o = orc.OracleScene(s)
    full, cf = o.render(40, 24, 3, 2, seed=3, schedule=1, threads=2)
    part, cp = o.render(40, 24, 3, 2, seed=3, schedule=1, threads=2, tile_stride=2)
    full, part = full.reshape(24, 40), part.reshape(24, 40)
    yy, xx = np.mgrid[0:24, 0:40]
    sel = ((xx // 4) % 2 == 0) & ((yy // 4) % 2 == 0)
    sel = sel[::-1]  # Sample_Stats rows are flipped (main.odin:95)
    assert sel.sum() == 5 * 3 * 16
    assert np.array_equal(part[sel], full[sel])
    assert np.all(part["count"][~sel] == 0)
    assert 0 < cp["rays"] < cf["rays"]


def test_check_intersect_ray_aabb_zero_direction_components():
    """SURVEY §8(c) fixture (1), `d_i = 0`: the reference divides by d (raytracer.odin:125-126), so a zero component
    gives an unbounded slab interval when the origin lies inside that slab and an empty one otherwise — also for -0.0.
    (The GPU's box test clamps |d| instead and must stay a superset: tests/test_gpu_round2.py.)"""
    lib = orc.load()
    t = C.c_float()
    lo, hi = [-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]
    inf = np.float32(np.inf)
    for zero in (0.0, -0.0):
        # origin inside the x slab, travelling along +z towards the box: hit at z distance 2
        assert lib.orc_check_intersect_ray_aabb(fp([0.5, 0.2, -3]), fp([zero, 0.0, 1.0]), fp(lo), fp(hi), inf, C.byref(t))
        assert t.value == 2
        # origin outside the x slab: never hit, whatever the other axes say
        assert not lib.orc_check_intersect_ray_aabb(fp([1.5, 0.2, -3]), fp([zero, 0.0, 1.0]), fp(lo), fp(hi), inf, C.byref(t))
        # two zero components, origin inside both slabs, box behind the origin: rejected by t2 < 0
        assert not lib.orc_check_intersect_ray_aabb(fp([0.5, 0.2, 3]), fp([zero, zero, 1.0]), fp(lo), fp(hi), inf, C.byref(t))
    # the benchmark camera's degenerate rays (profiles/r2_zero_direction_components.md) against the C4 root box
    root_lo, root_hi = [-50.0, -6.262298, -50.0], [50.0, 12.434973, 50.0]
    assert lib.orc_check_intersect_ray_aabb(fp([0, 16, 62]), fp([0.0, -0.5062409043312073, -0.8623921275138855]), fp(root_lo), fp(root_hi), inf, C.byref(t))
    assert not lib.orc_check_intersect_ray_aabb(fp([0, 16, 62]), fp([0.35083362460136414, 0.0, -0.9364379048347473]), fp(root_lo), fp(root_hi), inf, C.byref(t))


def test_c1_golden_at_baseline_size(scenes):
    """SURVEY §8(c) fixture (6): the C1 Cornell image at its BASELINE size (256x256, depth 6, 64 spp), committed as raw
    f32 totals (tests/golden/cornell_c1_256.npz) — the end-to-end regression pin of the oracle; the GPU is held
    against the same file in tests/test_gpu_round2.py."""
    g = np.load(os.path.join(GOLD, "cornell_c1_256.npz"))
    s = scenes("cornell", 256, 256)
    px, c = orc.OracleScene(s).render(256, 256, 6, 64, seed=int(g["seed"]), mode=0, schedule=1, threads=8)
    assert np.all(px["count"] == 64) and c["rays"] == int(g["n_rays"][0])
    assert px["total"].tobytes() == g["total"].tobytes()
