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


def test_hull_filter_never_drops_a_sphere_a_sample_could_hit(sim):
    """lc_hull_misses (the occluder-candidate filter for spheres) against the reference's own sphere test, on the functions as compiled:
    spheres placed at the boundary of the hull of the sample rays, +- relative offsets from 1e-6 to 1; whenever the filter drops one, none
    of the shadow samples (generated as path_shadow_sample_body generates them) may hit it. Both verdicts must occur, and kept spheres
    must actually get hit, or the check would be vacuous."""
    sim.sim_check_hull.argtypes = [C.c_uint, C.c_int, C.c_int, C.POINTER(C.c_ulonglong)]
    out = (C.c_ulonglong * 3)()
    bad = sim.sim_check_hull(20260419, 40000, 48, out)
    assert bad == 0
    assert out[0] > 5000 and out[1] > 5000 and out[2] > 50000, list(out)


def test_always_tested_triangle_bounds_contain_every_accepted_point(sim):
    """always_bound_of against triangle_t on the stored constants of precompute_triangle: nearly and exactly collinear triangles (a quarter
    exactly on a line before rounding, 30 % with equally spaced corners: the shapes whose barycentric test accepts points far from the
    triangle). Every accepted point lies inside both slabs; a triangle flagged "never" accepts nothing."""
    sim.sim_check_always.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_float, C.POINTER(C.c_ulonglong)]
    out = (C.c_ulonglong * 3)()
    bad = sim.sim_check_always(7, 20000, 400, 50.0, out)
    assert bad == 0
    assert out[0] > 10000 and out[1] > 10000, list(out)


def test_box_padding_covers_what_the_triangle_test_accepts_up_to_the_conditioning_limit(sim):
    """rt_bvh.hpp boxes a triangle when its conditioning number kappa = d00 d11 / denom is <= RT_BVH_KAPPA_MAX = 4e5 (1e5 until round 2) and
    pads the box by slop = 128 eps kappa lmax. Checked against triangle_t on the stored constants: every accepted point lies within that
    slop of the true triangle (distance in double), for kappa from 1e3 to 4e5; out[2] reports how much of the slop was ever used."""
    sim.sim_check_slop.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_ulonglong)]
    out = (C.c_ulonglong * 3)()
    bad = sim.sim_check_slop(11, 20000, 300, 1e3, 4e5, out)
    assert bad == 0
    assert out[0] > 10000 and out[1] > 100000, list(out)
    assert out[2] < 500000, "more than half of the slop used: %d ppm" % out[2]


@pytest.mark.parametrize("name", ["flamingo_pond", "backrooms_pool", "config5", "raccoon", "flamingo_lake", "mesh"])
def test_host_replay_of_the_stored_denominator_matches_the_device_function(hb, assets, sim, name):
    """rt_bvh.hpp drops a triangle whose STORED barycentric denominator is 0 or not finite (it can never report a hit); it computes that
    constant on the host (bvh_detail::stored_denominator). Compared bit for bit with precompute_triangle — the function the device runs
    — on every leaf reference of the scene's meshes."""
    sim.sim_check_den_replay.argtypes = [C.c_void_p, C.POINTER(hb.RtSceneDesc), C.POINTER(C.c_ulonglong)]
    s = hb.Scene(name, aspect=16 / 9)
    d = s.flatten()
    h = sim.sim_scene_create(d)
    n = C.c_ulonglong(0)
    bad = sim.sim_check_den_replay(h, d, C.byref(n))
    sim.sim_scene_destroy(h)
    assert bad == 0 and n.value > 0, (bad, n.value)


@pytest.mark.parametrize("name", ["random_spheres", "config5", "rt_in_a_weekend", "raccoon", "backrooms_pool", "cornell_box"])
def test_sphere_centres_lie_in_the_ball_the_padding_measures_from(hb, assets, sim, name):
    """The quadratic term of the per-ray box padding needs D >= |o - c| for every sphere centre c at every time in [0, 1]; D = |o - Cs| + Rs with
    (Cs, Rs) = AnalyticAccel::center_s / radius_s is that bound iff every centre lies inside the ball."""
    sim.sim_check_centre_ball.argtypes = [C.c_void_p, C.POINTER(hb.RtSceneDesc), C.POINTER(C.c_float)]
    s = hb.Scene(name, aspect=16 / 9)
    d = s.flatten()
    h = sim.sim_scene_create(d)
    slack = C.c_float(0)
    bad = sim.sim_check_centre_ball(h, d, C.byref(slack))
    sim.sim_scene_destroy(h)
    assert bad == 0 and (d.contents.n_spheres == 0 or slack.value >= 0.0), (bad, slack.value)


def test_always_tested_triangles_of_the_assets(hb, assets, sim):
    """Which triangles are still tested for every ray (rt_bvh.hpp): with the conditioning limit at 4e5 the pond and pool meshes have none (five and
    two at 1e5), the triceratops one (its second collinear sliver stores a zero denominator and is dropped: it can never report a hit)."""
    sim.sim_mesh_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int]
    want = {"flamingo_pond": [0, 0], "backrooms_pool": [0, 0, 0], "config5": [1, 0]}
    for name, counts in want.items():
        s = hb.Scene(name, aspect=16 / 9)
        h = sim.sim_scene_create(s.flatten())
        out = (C.c_int * 32)()
        n = sim.sim_mesh_info(h, out, 32)
        sim.sim_scene_destroy(h)
        assert [out[2 * i + 1] for i in range(n)] == counts, name
