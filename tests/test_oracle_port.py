"""Pins the plain C++ restatement (oracle/port) — against the reference itself (oracle/_ref) when that is built,
and always against the committed golden vectors, which the reference produced. CPU only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ALL_SCENES, GOLDEN, ROOT

PORT_DIR = os.path.join(ROOT, "oracle", "port")


@pytest.fixture(scope="module")
def port(hb):
    subprocess.check_call(["make", "-s"], cwd=PORT_DIR)
    L = C.CDLL(os.path.join(PORT_DIR, "_build", "librt_port.so"))
    L.port_scene_create.restype = C.c_void_p
    L.port_scene_create.argtypes = [C.POINTER(hb.RtSceneDesc)]
    L.port_scene_destroy.argtypes = [C.c_void_p]
    L.port_render.argtypes = [C.c_void_p, C.POINTER(hb.RtCamera), C.POINTER(hb.RtRenderParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    return L


def run_port(hb, port, name, w, h, spp, seed, scene_seed=0):
    s = hb.Scene(name, aspect=w / h, seed=scene_seed)
    hnd = port.port_scene_create(s.flatten())
    cam = hb.default_camera(w, h)
    p = hb.render_params(w, h, spp, seed=seed)
    lin = np.zeros((h, w, 3), np.float32)
    gam = np.zeros((h, w, 3), np.float32)
    ids = np.zeros((h, w, 4), np.uint32)
    port.port_render(hnd, C.byref(cam), C.byref(p), lin.ctypes.data, gam.ctypes.data, ids.ctypes.data, 0)
    port.port_scene_destroy(hnd)
    return lin, gam, ids


@pytest.mark.parametrize("name", ALL_SCENES)
def test_port_reproduces_golden(hb, assets, port, name):
    g = np.load(os.path.join(GOLDEN, "%s_96x54x2.npz" % name))
    lin, gam, ids = run_port(hb, port, name, 96, 54, 2, 0)
    assert np.array_equal(ids, g["ids"])
    assert np.array_equal(lin.view(np.uint32), g["linear"].view(np.uint32))
    assert np.array_equal(gam.view(np.uint32), g["gamma"].view(np.uint32))


@pytest.mark.parametrize("name", ["cornell_box", "random_spheres", "flamingo_pond", "backrooms_pool", "raccoon", "config5"])
def test_port_equals_reference(hb, ref, assets, port, name):
    w, h, spp = 80, 45, 3
    a = ref.scene(name, aspect=w / h, seed=4)
    want = a.render(w, h, spp, seed=13, threads=0)
    a.close()
    lin, gam, ids = run_port(hb, port, name, w, h, spp, 13, scene_seed=4)
    assert np.array_equal(ids, want["ids"])
    assert np.array_equal(lin.view(np.uint32), want["linear"].view(np.uint32))
    assert np.array_equal(gam.view(np.uint32), want["gamma"].view(np.uint32))
