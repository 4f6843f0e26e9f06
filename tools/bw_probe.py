"""Streaming-read bandwidth vs working-set size (ort_bench_read_bw): the L2 plateau is the peak the
L2-resident scenes (C1-C4) should be held against, the multi-GB tail is HBM."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytracer_odin_b200 import api

r = api.Renderer()
out = {}
for mb in (8, 16, 24, 32, 48, 64, 80, 96, 112, 128, 192, 256, 1024, 4096):
    out[mb] = round(r.bench_read_bw(mb << 20, 20 if mb <= 256 else 5), 1)
    print(mb, "MB", out[mb], "GB/s", flush=True)
print(json.dumps(out))
