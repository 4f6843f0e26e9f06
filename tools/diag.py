import os, sys, numpy as np, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import binding as orc
from raytracer_odin_b200 import api, gltf, scenegen, cabi
SEED=1234
d = tempfile.mkdtemp()
def load(p,w,h,env=None):
    s = gltf.read_gltf(p); s.fov_x = s.apply_render_config(w,h)
    if env: s.env_map = gltf.load_texture(env)
    return s.finish(orc.bvh_build)
def show(g, o, rays, idx):
    for i in idx[:8]:
        print('  ray', i, rays[i], '\n    gpu', g[i], '\n    orc', o[i])
# 1. cornell random rays
s = load(scenegen.cornell(d+'/c1.gltf'),64,64)
rng = np.random.default_rng(11); n=200000
lo,hi = s.bvh[-1]['lo'], s.bvh[-1]['hi']
rays = np.zeros(n, cabi.RAY_DTYPE)
rays['o'] = (lo+(hi-lo)*rng.random((n,3))).astype(np.float32)
dd = rng.normal(size=(n,3)); rays['d'] = (dd/np.linalg.norm(dd,axis=1,keepdims=True)).astype(np.float32)
o = orc.OracleScene(s)
ref,c = o.trace_rays(rays, mode=0); ref1,c1 = o.trace_rays(rays, mode=1)
print('counters faithful', c); print('counters ideal', c1)
with api.Renderer(seed=SEED).upload_scene(s) as r:
    g = r.trace_rays(rays)
    for f in ('tri','material','inside'):
        print(f, 'diff', (g[f]!=ref[f]).sum(), 'ideal-vs-faithful', (ref1[f]!=ref[f]).sum())
    hit = ref['tri']>=0
    for f in ('t','u','v'):
        bad = (g[f].view(np.uint32)!=ref[f].view(np.uint32)) & hit
        print(f, 'bitdiff', bad.sum())
    bad = np.nonzero((g['tri']!=ref['tri']) | ((g['t'].view(np.uint32)!=ref['t'].view(np.uint32))&hit) | (g['inside']!=ref['inside']))[0]
    show(g, ref, rays, bad)
    # 2. render cornell
    for (w,h,depth,spp) in [(64,64,6,16),(64,64,1,4),(64,64,2,4),(64,64,3,4)]:
        s2 = load(d+'/c1.gltf', w,h)
        r.upload_scene(s2); r.reset_stats()
        px = r.render(w,h,depth,spp); st = r.stats()
        opx, oc = orc.OracleScene(s2).render(w,h,depth,spp,seed=SEED)
        a,b = api.mean_image(px,w,h), api.mean_image(opx,w,h)
        print('render', depth, spp, api.rel_rmse(a,b), 'rays', st['rays_closest'], oc['rays'], 'lrays', st['rays_light_pdf'], oc['light_rays'],
              'close', np.isclose(a,b,rtol=1e-3,atol=1e-5).all(axis=2).mean(), 'nan', np.isnan(a).sum(), np.isnan(b).sum())
        badpix = np.nonzero(~np.isclose(a,b,rtol=1e-3,atol=1e-5).all(axis=2))
        for y,x in list(zip(*badpix))[:5]:
            print('   pix', y,x, a[y,x], b[y,x])
# 3. C2 full
w,h=1920,1080
s = load(scenegen.spheres(d+'/c2.gltf'), w,h)
o = orc.OracleScene(s)
ref, orays, c = o.primary_hits(w,h,0,SEED,0,threads=64)
ref1, _, c1 = o.primary_hits(w,h,0,SEED,1,threads=64)
print('C2 faithful', c); print('C2 ideal', c1, 'ideal-vs-faithful diff', (ref1['tri']!=ref['tri']).sum())
with api.Renderer(seed=SEED).upload_scene(s) as r:
    g = r.primary_hits(w,h,0)
    print(r.stats())
bad = np.nonzero(g['tri']!=ref['tri'])[0]
show(g, ref, orays, bad)
for i in bad[:4]:
    one = orays[i:i+1]
    # brute force all triangles through a flat single-leaf "BVH": use oracle intersect on both candidates
    import ctypes as C
    lib = orc.load()
    for tri in (g['tri'][i], ref['tri'][i]):
        if tri < 0: continue
        out = np.zeros(4, np.float32)
        oo = (one['o'][0] + one['d'][0]*np.float32(1e-3)).astype(np.float32)
        lib.orc_intersect_ray_triangle(oo.ctypes.data_as(C.POINTER(C.c_float)), np.ascontiguousarray(one['d'][0]).ctypes.data_as(C.POINTER(C.c_float)), s.triangles[tri:tri+1].ctypes.data, out.ctypes.data_as(C.POINTER(C.c_float)))
        print('   tri', tri, 'oracle isect', out, out.view(np.uint32))
