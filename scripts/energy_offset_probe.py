#!/usr/bin/env python3
"""Total-energy offset between the deterministic counter stream and the reference's mt19937, on the CPU.

The GPU renders the bits of oracle/_ref/libref_det.so (tests/test_gpu_parity.py), so `det` below IS the GPU's image.
    det      libref_det.so, seeds 100..100+n-1
    stock1   libref_stock.so with ONE thread: the reference's generator, no data race
    stockN   libref_stock.so with 16 threads: the shared racy generator (what round 1 compared against)
Prints z of the total energy for every pair, n renders a side.   python scripts/energy_offset_probe.py cornell_box 96 54 64 64
"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import oracle_ref

name, w, h, spp, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
det = oracle_ref.Ref().scene(name, aspect=w / h)
stk = oracle_ref.Ref(stock=True).scene(name, aspect=w / h)
D = np.stack([det.render(w, h, spp, seed=100 + i, threads=0, want_ids=False)["linear"].astype(np.float64) for i in range(n)])
S1 = np.stack([stk.render(w, h, spp, seed=i, threads=1, want_ids=False)["linear"].astype(np.float64) for i in range(n)])
SN = np.stack([stk.render(w, h, spp, seed=i, threads=16, want_ids=False)["linear"].astype(np.float64) for i in range(n)])


def z_total(A, B):
    ta, tb = A.sum((1, 2, 3)), B.sum((1, 2, 3))
    se = np.sqrt(ta.var(ddof=1) / len(ta) + tb.var(ddof=1) / len(tb))
    return (ta.mean() - tb.mean()) / se, 100 * (ta.mean() - tb.mean()) / tb.mean(), 100 * se / tb.mean()


for la, A, lb, B in (("det", D, "stock1", S1), ("det", D, "stock16", SN), ("stock16", SN, "stock1", S1)):
    z, pct, se = z_total(A, B)
    print("%s %dx%dx%d n=%d  %s - %s: z_total %+.2f  (%+.4f %% +- %.4f %%)" % (name, w, h, spp, n, la, lb, z, pct, se), flush=True)
