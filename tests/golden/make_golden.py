#!/usr/bin/env python3
"""Generate the committed golden vectors from the oracle (the reference itself, oracle/_ref).

Run in the build container (needs /root/reference compiled by oracle/Makefile and assets/_ref staged):
    python tests/golden/make_golden.py
Writes, per scene, <scene>_96x54x2.npz = {linear, gamma, ids} for a 96x54 render at 2 spp, seed 0,
default camera; plus rays_<scene>.npz = closest-hit known answers for 4096 seeded random rays
(origins on a shell around the scene, aimed at jittered points near the origin, random times) and
the Scene::rayTrace colour of the first 512 of them. Files are zip-compressed float32/uint32 arrays.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_ref  # noqa: E402

W, H, SPP = 96, 54, 2
RAY_SCENES = ["cornell_box", "random_spheres", "flamingo_pond", "backrooms_pool", "raccoon", "mesh"]


def random_rays(n, seed):
    rng = np.random.RandomState(seed)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    org = (d * rng.uniform(3.0, 12.0, size=(n, 1)) + np.array([0.0, 0.5, -4.0])).astype(np.float32)
    target = (rng.normal(size=(n, 3)) * np.array([3.0, 1.5, 4.0]) + np.array([0.0, -1.0, -4.0])).astype(np.float32)
    dirs = (target - org).astype(np.float32)          # deliberately NOT unit length: the Ray ctor normalises
    time = rng.uniform(0, 1, size=n).astype(np.float32)
    return org, dirs, time


def main():
    R = oracle_ref.Ref()
    for name in oracle_ref.SCENES:
        sc = R.scene(name, aspect=W / H, seed=0)
        r = sc.render(W, H, SPP, seed=0, threads=0)
        np.savez_compressed(os.path.join(HERE, "%s_%dx%dx%d.npz" % (name, W, H, SPP)), linear=r["linear"],
                            gamma=r["gamma"], ids=r["ids"])
        if name in RAY_SCENES:
            org, dirs, time = random_rays(4096, 719 + len(name))
            ids, aux = sc.trace_rays(org, dirs, time)
            rgb = sc.shade_rays(org[:512], dirs[:512], time[:512], seed=5)
            np.savez_compressed(os.path.join(HERE, "rays_%s.npz" % name), org=org, dirs=dirs, time=time, ids=ids,
                                aux=aux, rgb=rgb)
        sc.close()
        print("golden:", name)


if __name__ == "__main__":
    main()
