"""One wave of a BASELINE config for an `ncu --set full` capture: a warm-up wave, then the wave to capture.
   usage: ncu_wave.py CONFIG [waves]     (run with ORT_OVERLAP=1 so the launches of a wave are contiguous:
   k_raygen, then per bounce k_trace<closest> [k_trace<light>] k_shade, then k_resolve, k_stats)
   ncu ... --kernel-name regex:'k_trace|k_shade' --launch-skip <3*depth-1> --launch-count <3*depth-1> python tools/ncu_wave.py C4"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("ORT_OVERLAP", "1")
import torch
import bench
from raytracer_odin_b200 import api
from raytracer_odin_b200.scene import device_bvh_build

def main():
    config = sys.argv[1] if len(sys.argv) > 1 else "C4"
    waves = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    scene, cfg = bench.build_scene(config, device_bvh_build)
    w, h, depth = cfg["width"], cfg["height"], cfg["ray_depth"]
    spp_per_wave = max(1, (1 << 25) // (w * h))
    acc = torch.zeros(8, w * h, device="cuda")
    r = api.Renderer(seed=bench.SEED).upload_scene(scene)
    r.set_stream(torch.cuda.current_stream().cuda_stream)
    r.render_device(w, h, depth, 0, spp_per_wave, acc.data_ptr())          # warm-up wave (skipped by ncu)
    torch.cuda.synchronize()
    r.render_device(w, h, depth, 1000, spp_per_wave * waves, acc.data_ptr())  # captured
    torch.cuda.synchronize()
    st = r.stats()
    print(config, "spp_per_wave", spp_per_wave, "rays", st["rays_closest"], "launches", st["kernel_launches"])
    r.close()

if __name__ == "__main__":
    main()
