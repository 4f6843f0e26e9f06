import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the CUDA library and the oracle if a checkout has no binaries yet."""
    from raytracer_odin_b200 import cabi
    from oracle import binding

    if not os.path.exists(cabi.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    binding.build()


@pytest.fixture(scope="session")
def scene_dir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("scenes"))


def _load(path, w, h, builder, env=None):
    from raytracer_odin_b200 import gltf

    s = gltf.read_gltf(path)
    s.fov_x = s.apply_render_config(w, h)
    if env:
        s.env_map = gltf.load_texture(env)
    return s.finish(builder)


@pytest.fixture(scope="session")
def scenes(scene_dir):
    """Lazily generated, finished (oracle-built BVH) test scenes keyed by name."""
    from oracle import binding as orc
    from raytracer_odin_b200 import scenegen

    cache = {}

    def get(name, w=64, h=64):
        key = (name, w, h)
        if key in cache:
            return cache[key]
        d = os.path.join(scene_dir, name)
        env = None
        if name == "cornell":
            p = scenegen.cornell(os.path.join(d, "s.gltf"))
        elif name == "spheres_small":
            p = scenegen.spheres(os.path.join(d, "s.gltf"), n_spheres=14, subdiv=2, seed=5)
        elif name == "spheres_nolight":
            p = scenegen.spheres(os.path.join(d, "s.gltf"), n_spheres=10, subdiv=1, seed=9, n_emissive=0)
            env = scenegen.write_env_hdr(os.path.join(d, "env.hdr"), 64, 32, sun_peak=50.0)
        elif name == "spheres_c2":
            p = scenegen.spheres(os.path.join(d, "s.gltf"))
        elif name == "terrain_c4":
            p = scenegen.terrain(os.path.join(d, "s.gltf"))
        elif name == "textured_c3":
            p = scenegen.textured(os.path.join(d, "s.gltf"))
            env = scenegen.write_env_hdr(os.path.join(d, "env.hdr"), 2048, 1024)
        elif name == "terrain_small":
            p = scenegen.terrain(os.path.join(d, "s.gltf"), grid=48, n_spheres=24, subdiv=2, seed=3, n_emissive=3)
        elif name == "textured_small":
            p = scenegen.textured(os.path.join(d, "s.gltf"), tex_res=64, detail=0.15)
            env = scenegen.write_env_hdr(os.path.join(d, "env.hdr"), 128, 64)
        else:
            raise KeyError(name)
        from raytracer_odin_b200.scene import native_bvh_build

        # the two builders are tested equal (tests/test_host.py); the big scenes use the faster one
        cache[key] = _load(p, w, h, native_bvh_build if name.endswith(("_c4", "_c3", "_c2")) else orc.bvh_build, env)
        return cache[key]

    return get
