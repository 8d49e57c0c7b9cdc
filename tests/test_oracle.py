"""The oracle (reference built headless) against the committed golden vectors, and its own invariants.
CPU only. These pin the checker: if the oracle build, the asset staging or the deterministic stream drift,
the goldens stop matching."""
import ctypes
import os

import numpy as np
import pytest

from conftest import ALL_SCENES, GOLDEN

W, H, SPP = 96, 54, 2
FAST = ["single_sphere", "single_square", "cornell_box", "mesh", "random_spheres", "debug_refraction", "raccoon",
        "flamingo_pond", "backrooms_pool", "config5"]


@pytest.mark.parametrize("name", FAST)
def test_oracle_reproduces_golden(ref, name):
    g = np.load(os.path.join(GOLDEN, "%s_%dx%dx%d.npz" % (name, W, H, SPP)))
    sc = ref.scene(name, aspect=W / H, seed=0)
    r = sc.render(W, H, SPP, seed=0, threads=0)
    sc.close()
    assert np.array_equal(r["ids"], g["ids"])
    assert np.array_equal(r["linear"].view(np.uint32), g["linear"].view(np.uint32))
    assert np.array_equal(r["gamma"].view(np.uint32), g["gamma"].view(np.uint32))


def test_oracle_thread_count_invariant(ref):
    sc = ref.scene("random_spheres", aspect=W / H)
    a = sc.render(W, H, 3, seed=3, threads=1)
    b = sc.render(W, H, 3, seed=3, threads=5)
    sc.close()
    assert np.array_equal(a["linear"].view(np.uint32), b["linear"].view(np.uint32))
    assert a["n_random"] == b["n_random"]


def test_oracle_crop_equals_full(ref):
    sc = ref.scene("cornell_box", aspect=W / H)
    full = sc.render(W, H, 2, seed=1)
    crop = sc.render(W, H, 2, seed=1, crop=(10, 7, 50, 33))
    sc.close()
    assert np.array_equal(crop["linear"].view(np.uint32), full["linear"][7:33, 10:50].view(np.uint32))
    assert np.array_equal(crop["ids"], full["ids"][7:33, 10:50])


def test_oracle_seed_changes_image(ref):
    sc = ref.scene("random_spheres", aspect=W / H)
    a = sc.render(W, H, 1, seed=0)["linear"]
    b = sc.render(W, H, 1, seed=1)["linear"]
    sc.close()
    assert not np.array_equal(a, b)


def test_oracle_libm_pin_is_correctly_rounded(ref):
    L = ref.lib
    L.ref_probe_asinf.restype = ctypes.c_float
    L.ref_probe_asinf.argtypes = [ctypes.c_float]
    L.ref_probe_atan2f.restype = ctypes.c_float
    L.ref_probe_atan2f.argtypes = [ctypes.c_float, ctypes.c_float]
    rng = np.random.RandomState(1)
    xs = rng.uniform(-1, 1, 4000).astype(np.float32)
    ys = rng.uniform(-1, 1, 4000).astype(np.float32)
    for x, y in zip(xs, ys):
        assert np.float32(L.ref_probe_asinf(float(x))) == np.float32(np.arcsin(np.float64(x)))
        assert np.float32(L.ref_probe_atan2f(float(y), float(x))) == np.float32(np.arctan2(np.float64(y), np.float64(x)))


def test_oracle_ray_kats_match_golden(ref):
    for name in ["cornell_box", "flamingo_pond"]:
        g = np.load(os.path.join(GOLDEN, "rays_%s.npz" % name))
        sc = ref.scene(name, aspect=W / H)
        ids, aux = sc.trace_rays(g["org"], g["dirs"], g["time"])
        rgb = sc.shade_rays(g["org"][:512], g["dirs"][:512], g["time"][:512], seed=5)
        sc.close()
        assert np.array_equal(ids, g["ids"])
        assert np.array_equal(aux.view(np.uint32), g["aux"].view(np.uint32))
        assert np.array_equal(rgb.view(np.uint32), g["rgb"].view(np.uint32))
