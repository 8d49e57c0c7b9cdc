#!/usr/bin/env python3
"""Smallest program that launches the render kernels of one BASELINE workload, for ncu.

    python tools/profile_render.py --workload c2 --spp 4 [--crop X0 Y0 X1 Y1] [--reps 2] [--variant V]
Prints kernel time and the work counters; run it plain first, then the same command line under ncu
(B200_PROFILING.md recipe). A number printed under ncu is never a bench value.
"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS, flops_and_bytes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--crop", type=int, nargs=4)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--no-stats", action="store_true", help="skip the counting pass (a capture of one frame's kernels then holds exactly one frame)")
    a = ap.parse_args()
    hb = importlib.import_module("hai719-raytracing_b200")
    wl = WORKLOADS[a.workload]
    w, h, spp = wl["w"], wl["h"], a.spp or wl["spp"]
    s = hb.Scene(wl["scene"], aspect=w / h, seed=0)
    crop = tuple(a.crop) if a.crop else None
    if a.no_stats:
        for _ in range(a.reps):
            t = s.render(w, h, spp, seed=0, crop=crop, want_linear=False, variant=a.variant)["stats"]
        rays = t["n_closest_rays"] + t["n_shadow_rays"]
        print(json.dumps({"workload": a.workload, "spp": spp, "crop": crop, "kernel_ms": t["kernel_ms"], "launches": t["n_launches"], "rays": rays,
                          "paths": t["n_samples"], "mrays_per_s": rays / t["kernel_ms"] / 1e3 if rays else None}))
        return
    st = s.render(w, h, spp, seed=0, crop=crop, stats=True, want_linear=False, variant=a.variant)["stats"]
    rays, flops, byts = flops_and_bytes(st)
    for _ in range(a.reps):
        t = s.render(w, h, spp, seed=0, crop=crop, want_linear=False, variant=a.variant)["stats"]
    print(json.dumps({"workload": a.workload, "spp": spp, "crop": crop, "kernel_ms": t["kernel_ms"], "launches": t["n_launches"],
                      "rays": rays, "mrays_per_s": rays / t["kernel_ms"] / 1e3, "alg_flops_per_ray": flops / rays,
                      "alg_bytes_per_ray": byts / rays, "work": st}))


if __name__ == "__main__":
    main()
