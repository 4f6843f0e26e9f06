"""BASELINE config 5 (or another config) in --continious mode through the C++ `odinrt` command line: generates the
scene, runs `odinrt --continious --gpus ... --duration S --preview ...` and passes its output through.
   usage: sustained.py CONFIG GPUS(e.g. 0,1,2,3,4,5,6,7) DURATION_S [CHUNK] [extra odinrt flags ...]"""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytracer_odin_b200 import scenegen

config, gpus, duration = sys.argv[1], sys.argv[2], sys.argv[3]
chunk = sys.argv[4] if len(sys.argv) > 4 else "64"
extra = sys.argv[5:]
cfg = scenegen.CONFIGS[config]
d = tempfile.mkdtemp(prefix=f"ort_{config}_")
t0 = time.time()
path, env = scenegen.generate(config, d)
print(f"generated {path} in {time.time() - t0:.1f}s", flush=True)
out = os.path.join(ROOT, "gpurun_out", f"sustained_{config.lower()}_{gpus.count(',') + 1}gpu")
cmd = [os.path.join(ROOT, "raytracer-odin_b200", "host", "odinrt"), path, out + ".png", "--width", str(cfg["width"]),
       "--height", str(cfg["height"]), "--ray-depth", str(cfg["ray_depth"]), "--continious", "--gpus", gpus,
       "--duration", duration, "--chunk", chunk, "--bvh", "device", "--preview", out + "_preview.png", "--preview-every", "10"] + extra
if env:
    cmd += ["--env-map", env]
print(" ".join(cmd), flush=True)
t0 = time.time()
rc = subprocess.call(cmd)
print(f"odinrt exit code {rc}, wall {time.time() - t0:.1f}s", flush=True)
for f in (out + ".png", out + "_preview.png"):
    if os.path.exists(f):
        print(f, os.path.getsize(f), "bytes")
        if os.path.getsize(f) > 6 << 20:
            os.remove(f)  # keep gpurun_out small; the run's numbers are on stdout
sys.exit(rc)
