#!/usr/bin/env python3
"""Repeatability probe: kernel_ms of the same render, several times, for a few variants, re-uploading the scene in between."""
import importlib, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS
hb = importlib.import_module("hai719-raytracing_b200")
wl = WORKLOADS[sys.argv[1]]; spp = int(sys.argv[2]); variants = [int(v) for v in sys.argv[3:]]
w, h = wl["w"], wl["h"]
s = hb.Scene(wl["scene"], aspect=w / h, seed=0)
for rep in range(3):
    for v in variants:
        ts = []
        for i in range(4):
            ts.append(s.render(w, h, spp, seed=0, want_linear=False, variant=v)["stats"]["kernel_ms"])
        print("upload %d variant %6d kernel_ms %s" % (rep, v, " ".join(("%.3f" if max(ts) < 5 else "%.1f") % t for t in ts)), flush=True)
    s.invalidate_device()
