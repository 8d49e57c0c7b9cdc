#!/usr/bin/env python3
"""Stage the scene inputs (meshes, textures) under assets/_ref/ — git-ignored, shipped by gpurun.

The reference's scene builders open "mesh/*.off" and "img/**/*.ppm" relative to the working
directory (Scene.h:360,424,755,...). Those files are DATA of the reference checkout; they are
copied here at build time (never committed) so that the GPU box, which has no /root/reference,
sees the same inputs. Six blobs are absent from the checkout (.MISSING_LARGE_BLOBS); the three
that a scene builder actually opens are synthesised deterministically:

  img/textures/sky.ppm              1024x512 P6 procedural sky (gradient + sun + cloud bands)
  img/textures/space.ppm            1024x512 P6 procedural star field
  mesh/flamingo_float_colored.off   flamingo_float.off rewritten as COFF with the constant
                                    colour 237 149 218 255 (the scene's own material colour,
                                    Scene.h:1828)

Both the oracle (reference build) and the product's host API read this one directory, so the
substitutes cancel out of every parity comparison.
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
REF = os.environ.get("HAI719_REFERENCE", "/root/reference")


def write_p6(path, img):
    h, w, _ = img.shape
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(np.ascontiguousarray(img, dtype=np.uint8).tobytes())


def make_sky(path, w=1024, h=512):
    v = (np.arange(h, dtype=np.float64) + 0.5) / h            # 0 = zenith, 1 = nadir
    u = (np.arange(w, dtype=np.float64) + 0.5) / w
    V, U = np.meshgrid(v, u, indexing="ij")
    horizon = np.exp(-((V - 0.5) / 0.12) ** 2)
    top = np.stack([0.25 + 0.35 * V, 0.45 + 0.4 * V, 0.95 - 0.1 * V], -1)
    ground = np.stack([0.35 + 0 * V, 0.32 + 0 * V, 0.28 + 0 * V], -1)
    img = np.where((V < 0.5)[..., None], top, ground)
    img = img + horizon[..., None] * np.array([0.35, 0.3, 0.2])
    clouds = 0.5 + 0.5 * np.sin(U * 37.0 + 5.0 * np.sin(V * 23.0)) * np.sin(V * 41.0 + 3.0 * np.sin(U * 17.0))
    clouds = np.clip((clouds - 0.62) * 4.0, 0, 1) * (V < 0.47)
    img = img * (1 - clouds[..., None]) + clouds[..., None] * 0.97
    sun = np.exp(-(((U - 0.3) * 2.0) ** 2 + (V - 0.3) ** 2) / 0.0008)
    img = img + sun[..., None] * np.array([1.0, 0.95, 0.8])
    write_p6(path, np.clip(img * 255.0 + 0.5, 0, 255).astype(np.uint8))


def make_space(path, w=1024, h=512):
    rng = np.random.RandomState(719)
    img = np.zeros((h, w, 3), dtype=np.float64)
    img += np.array([0.01, 0.01, 0.03])
    n = 2500
    xs = rng.randint(0, w, n)
    ys = rng.randint(0, h, n)
    mag = rng.rand(n) ** 3
    tint = 0.7 + 0.3 * rng.rand(n, 3)
    img[ys, xs] = mag[:, None] * tint
    yy, xx = np.mgrid[0:h, 0:w]
    neb = np.exp(-(((xx - 700) / 180.0) ** 2 + ((yy - 200) / 90.0) ** 2))
    img += neb[..., None] * np.array([0.25, 0.05, 0.3])
    write_p6(path, np.clip(img * 255.0 + 0.5, 0, 255).astype(np.uint8))


def make_flamingo_colored(src, dst):
    with open(src) as f:
        tok = f.read().split()
    assert tok[0] == "OFF"
    nv, nt = int(tok[1]), int(tok[2])
    p = 4
    out = ["COFF", "%d %d 0" % (nv, nt)]
    for i in range(nv):
        out.append("%s %s %s 237 149 218 255 " % (tok[p], tok[p + 1], tok[p + 2]))
        p += 3
    for i in range(nt):
        assert tok[p] == "3"
        out.append("3 %s %s %s " % (tok[p + 1], tok[p + 2], tok[p + 3]))
        p += 4
    with open(dst, "w") as f:
        f.write("\n".join(out) + "\n")


def stage(force=False):
    if not os.path.isdir(REF):
        if os.path.isdir(os.path.join(DST, "mesh")):
            return DST            # GPU box: use what travelled with the snapshot
        raise SystemExit("no reference checkout at %s and nothing staged in %s" % (REF, DST))
    for sub in ("mesh", "img/normalMaps", "img/planeTextures", "img/sphereTextures", "img/textures"):
        os.makedirs(os.path.join(DST, sub), exist_ok=True)
    for sub in ("mesh", "img/normalMaps", "img/planeTextures", "img/sphereTextures"):
        for name in sorted(os.listdir(os.path.join(REF, sub))):
            s = os.path.join(REF, sub, name)
            d = os.path.join(DST, sub, name)
            if force or not os.path.exists(d) or os.path.getsize(d) != os.path.getsize(s):
                shutil.copyfile(s, d)
    sky = os.path.join(DST, "img/textures/sky.ppm")
    if force or not os.path.exists(sky):
        make_sky(sky)
    space = os.path.join(DST, "img/textures/space.ppm")
    if force or not os.path.exists(space):
        make_space(space)
    fl = os.path.join(DST, "mesh/flamingo_float_colored.off")
    if force or not os.path.exists(fl):
        make_flamingo_colored(os.path.join(DST, "mesh/flamingo_float.off"), fl)
    return DST


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
