"""DEBUG AID, CPU only: the device functions of csrc/rt_core.cuh compiled as plain C++ (tests/hostsim)
and compared with the oracle. This is how the path logic is checked in a container without a GPU; it is
not a product path (nothing under hai719-raytracing_b200/ links or loads it) and it does not replace the
-m gpu parity tests, which exercise the real sm_100a kernels through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

SIM_DIR = os.path.join(ROOT, "tests", "hostsim")


def _load_sim(hb, so_name, defs):
    so = os.path.join(SIM_DIR, so_name)
    src = os.path.join(SIM_DIR, "hostsim.cpp")
    core = os.path.join(ROOT, "hai719-raytracing_b200", "csrc", "rt_core.cuh")
    deps = [src] + [os.path.join(os.path.dirname(core), f) for f in os.listdir(os.path.dirname(core)) if f.endswith((".cuh", ".hpp"))]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in deps):
        subprocess.check_call(["/usr/bin/g++", "-O3", "-fPIC", "-shared", "-std=c++17", "-I", os.path.join(ROOT, "include"),
                               "-I", os.path.dirname(core)] + defs + [src, "-o", so, "-lpthread"])
    L = C.CDLL(so)
    L.sim_scene_create.restype = C.c_void_p
    L.sim_scene_create.argtypes = [C.POINTER(hb.RtSceneDesc)]
    L.sim_scene_destroy.argtypes = [C.c_void_p]
    L.sim_render.argtypes = [C.c_void_p, C.POINTER(hb.RtCamera), C.POINTER(hb.RtRenderParams), C.c_void_p, C.c_void_p,
                             C.c_void_p, C.c_int]
    return L


@pytest.fixture(scope="module")
def sim(hb):
    return _load_sim(hb, "libhostsim.so", [])


@pytest.fixture(scope="module")
def sim_fma(hb):
    """The same functions with the box tests of the culling hierarchies in the OTHER form: subtract-then-multiply planes
    instead of one FMA per plane (RT_OPT_BOXFMA / RT_OPT_CONEFMA = 0; the product build has both on since round 2): a
    different conservative filter, so the same bits."""
    return _load_sim(hb, "libhostsim_nofma.so", ["-DRT_OPT_BOXFMA=0", "-DRT_OPT_CONEFMA=0"])


@pytest.fixture(scope="module")
def sim_bvh4(hb):
    """The mesh walk over the 4-wide copy of the culling hierarchies (RT_OPT_BVH4): another shape of the same
    conservative filter, so the same bits."""
    return _load_sim(hb, "libhostsim_bvh4.so", ["-DRT_OPT_BVH4=1"])


def _check_against_oracle(hb, ref, sim, name, variant):
    W, H, SPP = 64, 36, 2
    a = ref.scene(name, aspect=W / H)
    want = a.render(W, H, SPP, seed=2, threads=0)
    a.close()
    s = hb.Scene(name, aspect=W / H)
    h = sim.sim_scene_create(s.flatten())
    cam = hb.default_camera(W, H)
    p = hb.render_params(W, H, SPP, seed=2, variant=variant)
    lin = np.zeros((H, W, 3), np.float32)
    gam = np.zeros((H, W, 3), np.float32)
    ids = np.zeros((H, W, 4), np.uint32)
    sim.sim_render(h, C.byref(cam), C.byref(p), lin.ctypes.data, gam.ctypes.data, ids.ctypes.data, 0)
    sim.sim_scene_destroy(h)
    assert np.array_equal(ids, want["ids"])
    assert np.array_equal(lin.view(np.uint32), want["linear"].view(np.uint32))
    assert np.array_equal(gam.view(np.uint32), want["gamma"].view(np.uint32))


@pytest.mark.parametrize("name", ["random_spheres", "flamingo_pond", "backrooms_pool", "config5"])
@pytest.mark.parametrize("variant", [3, 5, 6])
def test_fma_box_tests_keep_the_bits(hb, ref, assets, sim_fma, name, variant):
    _check_against_oracle(hb, ref, sim_fma, name, variant)


@pytest.mark.parametrize("name", ["cornell_box", "random_spheres", "flamingo_pond", "backrooms_pool", "raccoon",
                                  "rt_in_a_weekend", "flamingo_lake", "config5"])
@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, 6])
def test_core_functions_on_cpu_match_oracle(hb, ref, assets, sim, name, variant):
    _check_against_oracle(hb, ref, sim, name, variant)


@pytest.mark.parametrize("name", ["flamingo_pond", "backrooms_pool", "config5", "raccoon", "mesh"])
@pytest.mark.parametrize("variant", [3, 6])
def test_four_wide_mesh_walk_keeps_the_bits(hb, ref, assets, sim_bvh4, name, variant):
    _check_against_oracle(hb, ref, sim_bvh4, name, variant)
