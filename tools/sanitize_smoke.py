"""Small end-to-end run for compute-sanitizer --tool memcheck (one tool per gpurun call)."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytracer_odin_b200 import api, gltf, scenegen, cabi
from raytracer_odin_b200.scene import native_bvh_build
d = tempfile.mkdtemp()
for name, kw, env in (("cornell", {}, False), ("spheres", dict(n_spheres=6, subdiv=1, seed=2), False), ("textured", dict(tex_res=32, detail=0.1), True)):
    s = gltf.read_gltf(getattr(scenegen, name)(os.path.join(d, name + ".gltf"), **kw))
    s.fov_x = s.apply_render_config(40, 24)
    if env:
        s.env_map = gltf.load_texture(scenegen.write_env_hdr(os.path.join(d, "e.hdr"), 32, 16))
    s.finish(native_bvh_build)
    with api.Renderer(seed=1, max_paths_in_flight=40 * 24 * 2).upload_scene(s) as r:
        px = r.render(40, 24, 5, 7)
        hits = r.primary_hits(40, 24, 1)
        rays = np.zeros(100, cabi.RAY_DTYPE); rays["d"] = [0, 0, -1]; rays["o"] = [0, 0, 3]
        r.trace_rays(rays); r.light_pdf(rays)
    with api.MultiRenderer([0, 0], seed=1).upload_scene(s) as m:
        m.render(40, 24, 4, 5)
    print(name, "ok", float(px["total"].sum()))
