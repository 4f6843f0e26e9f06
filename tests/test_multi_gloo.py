"""N > 1 path on CPU: world_size-2 gloo.  Each rank renders its block of the sample axis (with the
CPU oracle standing in for the GPU replica — this is a test), packs the planar accumulator the
library uses (total | total_squared | count_lo | count_hi), and ONE reduce to rank 0 must reproduce the
single-process render of the whole sample range up to f32 summation order — the strong-scaling split
bench.py uses: one fixed frame, its sample range divided over the ranks."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _planar(px, npix):
    acc = np.zeros((8, npix), np.float32)
    acc[0:3] = px["total"].T
    acc[3:6] = px["total_squared"].T
    acc[6] = px["count"] & 0xFFFFF  # count = lo + 2^20 * hi: both planes exact in f32, also under the sum-reduce
    acc[7] = px["count"] >> 20
    return acc


def _worker(rank, world, port, gltf_path, w, h, depth, spp, seed, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import binding as orc
    from raytracer_odin_b200 import gltf, multigpu

    s = gltf.read_gltf(gltf_path)
    s.fov_x = s.apply_render_config(w, h)
    s.finish(orc.bvh_build)
    first, cnt = multigpu.sample_partition(0, spp, rank, world)
    px, _ = orc.OracleScene(s).render(w, h, depth, cnt, first_sample=first, seed=seed, threads=2)
    acc = torch.from_numpy(_planar(px, w * h))
    multigpu.reduce_accum(acc, dst=0)
    if rank == 0:
        np.save(out_path, acc.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sample_split_reduce_world2(scene_dir, tmp_path):
    from oracle import binding as orc
    from raytracer_odin_b200 import gltf, scenegen

    w, h, depth, spp, seed = 32, 32, 4, 6, 5
    path = scenegen.cornell(os.path.join(scene_dir, "gloo_c1.gltf"))
    out = str(tmp_path / "acc.npy")
    mp.spawn(_worker, args=(2, 29500 + os.getpid() % 2000, path, w, h, depth, spp, seed, out), nprocs=2, join=True)
    got = np.load(out)
    s = gltf.read_gltf(path)
    s.fov_x = s.apply_render_config(w, h)
    s.finish(orc.bvh_build)
    px, _ = orc.OracleScene(s).render(w, h, depth, spp, seed=seed, threads=2)
    want = _planar(px, w * h)
    assert np.array_equal(got[6] + got[7] * (1 << 20), want[6] + want[7] * (1 << 20)) and np.all(got[6] == spp)
    np.testing.assert_allclose(got[:6], want[:6], rtol=1e-5, atol=1e-6)
