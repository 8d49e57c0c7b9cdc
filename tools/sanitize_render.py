#!/usr/bin/env python3
"""Tiny renders of every kernel family, for compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_render.py
    compute-sanitizer --tool racecheck python tools/sanitize_render.py
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hb = importlib.import_module("hai719-raytracing_b200")
W, H, SPP = 64, 36, 2
os.environ["HAI719_CHUNK_LOG2"] = "16"      # several chunks: queues and counters are reused
for name in ("random_spheres", "config5", "backrooms_pool", "flamingo_pond", "cornell_box"):
    s = hb.Scene(name, aspect=W / H)
    for v in (6, 5, 3, 1):
        out = s.render(W, H, SPP, seed=1, variant=v, stats=(v == 6))
        print(name, v, float(out["linear"].sum()), flush=True)
    s.update_device()
    s.render_rgb8(W, H, SPP, seed=1)
    s.close()
print("done")
