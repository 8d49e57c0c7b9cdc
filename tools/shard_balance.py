#!/usr/bin/env python3
"""Load balance of the tile sharding, measured on ONE GPU: render each rank's shard of an N-rank job in turn and
print the kernel time per shard. max/mean is the efficiency bound tile assignment puts on an N-GPU run.

    python tools/shard_balance.py c2 [--ranks 8] [--spp 0] [--tile 32 32] [--variant 0]
"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("--ranks", type=int, default=8)
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--tile", type=int, nargs=2, default=[32, 32])
    ap.add_argument("--variant", type=int, default=0)
    a = ap.parse_args()
    hb = importlib.import_module("hai719-raytracing_b200")
    wl = WORKLOADS[a.workload]
    w, h, spp = wl["w"], wl["h"], a.spp or wl["spp"]
    s = hb.Scene(wl["scene"], aspect=w / h, seed=0)
    s.render(w, h, 1, want_linear=False)   # warm-up: upload, module load
    whole = min(s.render(w, h, spp, want_linear=False, variant=a.variant, tile=tuple(a.tile))["stats"]["kernel_ms"] for _ in range(2))
    ms = []
    for r in range(a.ranks):
        t = min(s.render(w, h, spp, want_linear=False, variant=a.variant, rank=r, n_ranks=a.ranks, tile=tuple(a.tile))["stats"]["kernel_ms"]
                for _ in range(2))
        ms.append(t)
    mean = sum(ms) / len(ms)
    print(json.dumps({"workload": a.workload, "spp": spp, "ranks": a.ranks, "tile": a.tile, "variant": a.variant, "whole_ms": whole,
                      "shard_ms": ms, "max_over_mean": max(ms) / mean, "sum_over_whole": sum(ms) / whole,
                      "scaling_bound": whole / (a.ranks * max(ms))}))


if __name__ == "__main__":
    main()
