#!/usr/bin/env python3
"""Speculative shadow samples: share of pending samples each configuration of rounds validates (rt_debug_counters).
    python tools/spec_probe.py c5 2 1 2 3 6"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS
hb = importlib.import_module("hai719-raytracing_b200")
wl = WORKLOADS[sys.argv[1]]; spp = int(sys.argv[2]); rounds = [int(v) for v in sys.argv[3:]]
w, h = wl["w"], wl["h"]
s = hb.Scene(wl["scene"], aspect=w / h, seed=0)
for r in rounds:
    st = s.render(w, h, spp, seed=0, want_linear=False, variant=6 | (r << 20), stats=True)["stats"]
    c = s.debug_counters()
    print("%s x%d rounds %2d: shadow rays %d, validated speculatively %d (%.1f %% of all shadow rays), pending examined %d, validated/pending %.3f"
          % (sys.argv[1], spp, r, st["n_shadow_rays"], c[13], 100.0 * c[13] / max(1, st["n_shadow_rays"]), c[14], c[13] / max(1, c[14])), flush=True)
