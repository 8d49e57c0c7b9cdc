"""Parity tests proper: the sm_100a kernels, called through the C ABI, against the oracle.

Bars (BASELINE.json north_star): primary-hit ids identical on >= 99.99 % of pixels; final RGB max-abs error
<= 1e-3 and PSNR >= 50 dB in deterministic mode. The kernels reproduce the reference's operation order
and precision (no FMA, fp64 where the reference promotes), so the tests also assert the stronger property
that almost every pixel is BIT-identical before gamma; the allowance covers CUDA's fp64 acos/atan2/asin/pow
being 1-2 ulp functions where glibc's are <1 ulp."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ALL_SCENES, GOLDEN

pytestmark = pytest.mark.gpu
W, H, SPP = 96, 54, 2


def psnr(a, b):
    mse = float(np.mean((np.clip(a, 0, 1).astype(np.float64) - np.clip(b, 0, 1).astype(np.float64)) ** 2))
    return 99.0 if mse == 0 else 10 * np.log10(1.0 / mse)


def compare(got, want, ids_got=None, ids_want=None, min_bitexact=0.999):
    """Returns a report string; asserts the north-star tolerances plus the bit-exact fraction."""
    lin_same = (got["linear"].view(np.uint32) == want["linear"].view(np.uint32)).all(-1)
    gerr = np.abs(np.nan_to_num(got["gamma"]) - np.nan_to_num(want["gamma"]))
    p = psnr(got["gamma"], want["gamma"])
    rep = "bit-identical linear %.4f%%, gamma max-abs %.3g, PSNR %.1f dB" % (100 * lin_same.mean(), gerr.max(), p)
    if ids_got is not None:
        same_id = (ids_got[..., :3] == ids_want[..., :3]).all(-1)
        rep += ", ids %.4f%%" % (100 * same_id.mean())
        assert same_id.mean() >= 0.9999, rep
    assert gerr.max() <= 1e-3, rep
    assert p >= 50.0, rep
    assert lin_same.mean() >= min_bitexact, rep
    return rep


@pytest.fixture(scope="module")
def gpu(hb):
    if hb.device_count() < 1:
        pytest.fail("no sm_100 device: the -m gpu tests need a B200 (there is no CPU fallback to test)")
    return hb


@pytest.mark.parametrize("name", ALL_SCENES)
def test_render_matches_golden(gpu, assets, name):
    g = np.load(os.path.join(GOLDEN, "%s_%dx%dx%d.npz" % (name, W, H, SPP)))
    s = gpu.Scene(name, aspect=W / H, seed=0)
    out = s.render(W, H, SPP, seed=0)
    ids = s.trace_primary(W, H, seed=0)
    print(name, compare(out, g, ids, g["ids"]))
    assert np.array_equal(ids[..., 3], g["ids"][..., 3]) or (ids[..., 3] == g["ids"][..., 3]).mean() > 0.9999


@pytest.mark.parametrize("name,w,h,spp", [("cornell_box", 425, 240, 4), ("random_spheres", 320, 180, 8),
                                          ("flamingo_pond", 192, 108, 2), ("backrooms_pool", 192, 108, 4),
                                          ("config5", 240, 135, 4), ("flamingo_lake", 192, 108, 2)])
def test_render_matches_oracle(gpu, ref, assets, name, w, h, spp):
    a = ref.scene(name, aspect=w / h, seed=3)
    want = a.render(w, h, spp, seed=11, threads=0)
    a.close()
    s = gpu.Scene(name, aspect=w / h, seed=3)
    out = s.render(w, h, spp, seed=11)
    ids = s.trace_primary(w, h, seed=11)
    print(name, compare(out, want, ids, want["ids"]))


@pytest.mark.parametrize("name", ["cornell_box", "flamingo", "flamingo_pond", "backrooms_pool"])
def test_primary_hit_ids_at_reference_resolution(gpu, ref, assets, name):
    """850x480 (main.cpp:52-53) primary hits; scenes 7 and 9 hold the 64 zero-area triangles whose NaN
    normals FMA contraction would turn into spurious hits."""
    w, h = 850, 480
    a = ref.scene(name, aspect=w / h)
    want = a.render(w, h, 1, seed=0, threads=0, crop=(0, 200, w, 280))["ids"]
    a.close()
    ids = gpu.Scene(name, aspect=w / h).trace_primary(w, h, seed=0, crop=(0, 200, w, 280))
    same = (ids == want).all(-1).mean()
    print(name, "ids+t bit-identical on %.5f%% of %d pixels" % (100 * same, ids.shape[0] * ids.shape[1]))
    assert same >= 0.9999


@pytest.mark.parametrize("name", ["cornell_box", "random_spheres", "flamingo_pond", "backrooms_pool", "raccoon", "mesh"])
def test_ray_known_answers(gpu, assets, name):
    g = np.load(os.path.join(GOLDEN, "rays_%s.npz" % name))
    s = gpu.Scene(name, aspect=W / H)
    ids, aux = s.trace_rays(g["org"], g["dirs"], g["time"])
    assert (ids == g["ids"]).all(-1).mean() >= 0.9999
    hit = g["ids"][:, 0] > 0
    assert hit.sum() > 100
    # theta/phi of sphere hits go through fp64 acos/atan2: equal to within 1 float ulp; everything else exact
    close = np.isclose(aux, g["aux"], rtol=3e-7, atol=1e-7).all(-1)
    assert close.mean() >= 0.9999
    rgb = s.shade_rays(g["org"][:512], g["dirs"][:512], g["time"][:512], seed=5)
    assert np.abs(rgb - g["rgb"]).max() <= 1e-3
    assert (rgb.view(np.uint32) == g["rgb"].view(np.uint32)).all(-1).mean() >= 0.99


@pytest.mark.parametrize("name", ["cornell_box", "random_spheres", "flamingo_pond", "backrooms_pool", "rt_in_a_weekend", "config5"])
def test_kernel_variants_agree_bit_for_bit(gpu, assets, name):
    """variant 1 = one path per lane to completion; 2 = ray-level state machine with path regeneration (threshold
    1 and 16) walking the reference's KD order; 3 = the same with the exact culling hierarchy; 0 = auto."""
    g = np.load(os.path.join(GOLDEN, "%s_%dx%dx%d.npz" % (name, W, H, SPP)))
    s = gpu.Scene(name, aspect=W / H, seed=0)
    a = s.render(W, H, SPP, seed=0, variant=1, stats=True)
    b = s.render(W, H, SPP, seed=0, variant=2 | (1 << 8), stats=True)
    c = s.render(W, H, SPP, seed=0, variant=2 | (16 << 8))
    d = s.render(W, H, SPP, seed=0, variant=0)
    e = s.render(W, H, SPP, seed=0, variant=3, stats=True)
    f = s.render(W, H, SPP, seed=0, variant=4 | (2 << 16))
    assert np.array_equal(a["linear"].view(np.uint32), e["linear"].view(np.uint32))
    assert np.array_equal(a["linear"].view(np.uint32), f["linear"].view(np.uint32))
    # variant 5 (occluder candidates per hit and light; falls back to 3 without lights / analytic hierarchy),
    # traversal thresholds 1, 16 (default) and 32, with and without counters
    for v in (5 | (1 << 20), 5, 5 | (32 << 20) | (2 << 16)):
        q = s.render(W, H, SPP, seed=0, variant=v)
        assert np.array_equal(a["linear"].view(np.uint32), q["linear"].view(np.uint32)), v
    q = s.render(W, H, SPP, seed=0, variant=5, stats=True)
    assert np.array_equal(a["linear"].view(np.uint32), q["linear"].view(np.uint32))
    for k in ("n_closest_rays", "n_shadow_rays", "n_random", "n_tex_fetches"):
        assert a["stats"][k] == q["stats"][k], k
    # variant 6: wavefront (camera rays / trace+shade / light+scatter kernels per bounce level)
    for st in (False, True):
        q = s.render(W, H, SPP, seed=0, variant=6, stats=st)
        assert np.array_equal(a["linear"].view(np.uint32), q["linear"].view(np.uint32)), st
    for k in ("n_closest_rays", "n_shadow_rays", "n_random", "n_tex_fetches"):
        assert a["stats"][k] == q["stats"][k], k
    for k in ("n_closest_rays", "n_shadow_rays", "n_random", "n_tex_fetches"):
        assert a["stats"][k] == e["stats"][k], k
    assert np.array_equal(a["linear"].view(np.uint32), b["linear"].view(np.uint32))
    assert np.array_equal(a["linear"].view(np.uint32), c["linear"].view(np.uint32))
    assert np.array_equal(a["linear"].view(np.uint32), d["linear"].view(np.uint32))
    assert np.array_equal(b["linear"].view(np.uint32), g["linear"].view(np.uint32))
    for k in ("n_closest_rays", "n_shadow_rays", "n_random", "n_tex_fetches"):
        assert a["stats"][k] == b["stats"][k], k


@pytest.mark.parametrize("name,w,h,spp", [("flamingo_pond", 480, 270, 4), ("backrooms_pool", 480, 270, 8), ("raccoon", 480, 270, 4),
                                          ("config5", 480, 270, 8), ("flamingo", 480, 270, 2), ("mesh", 480, 270, 4)])
def test_exact_culling_equals_reference_order(gpu, assets, name, w, h, spp):
    """Variant 3 never walks the reference's KD-tree; it must still pick the same triangle at the same t for every
    ray of every bounce (depth-100 trees, dropped triangles, NaN normals, transparent shadow casters included)."""
    s = gpu.Scene(name, aspect=w / h, seed=2)
    a = s.render(w, h, spp, seed=8, variant=1)
    b = s.render(w, h, spp, seed=8, variant=3)
    assert np.array_equal(a["linear"].view(np.uint32), b["linear"].view(np.uint32))


@pytest.mark.parametrize("name,w,h,spp", [("flamingo_pond", 1600, 900, 4), ("backrooms_pool", 1600, 900, 8), ("config5", 1600, 900, 4),
                                          ("raccoon", 1280, 720, 4), ("flamingo_lake", 1280, 720, 2), ("random_spheres", 1920, 1080, 8),
                                          ("rt_in_a_weekend", 1280, 720, 4)])
def test_exact_culling_at_scale(gpu, assets, name, w, h, spp):
    """Tens of millions of rays per scene: the culled traversal (variant 3) against the reference-order traversal
    (variant 1). Images must be bit-identical and the ray / random-draw counters equal — one wrong hit anywhere
    changes the continuation of its path and therefore the counts (this is how a 1-in-1e8 miss was caught:
    profiles/r01_notes.md, 'sliver triangles')."""
    s = gpu.Scene(name, aspect=w / h, seed=0)
    a = s.render(w, h, spp, seed=21, variant=1, stats=True)
    b = s.render(w, h, spp, seed=21, variant=3, stats=True)
    c = s.render(w, h, spp, seed=21, variant=4, stats=True)
    d = s.render(w, h, spp, seed=21, variant=5, stats=True)
    e = s.render(w, h, spp, seed=21, variant=6, stats=True)
    for k in ("n_closest_rays", "n_shadow_rays", "n_random", "n_tex_fetches"):
        assert a["stats"][k] == b["stats"][k], (k, a["stats"][k], b["stats"][k])
        assert a["stats"][k] == c["stats"][k], (k, a["stats"][k], c["stats"][k])
        assert a["stats"][k] == d["stats"][k], (k, a["stats"][k], d["stats"][k])
        assert a["stats"][k] == e["stats"][k], (k, a["stats"][k], e["stats"][k])
    diff = (a["linear"].view(np.uint32) != d["linear"].view(np.uint32)).any(-1)
    assert not diff.any(), np.argwhere(diff)[:8]
    diff = (a["linear"].view(np.uint32) != e["linear"].view(np.uint32)).any(-1)
    assert not diff.any(), np.argwhere(diff)[:8]
    diff = (a["linear"].view(np.uint32) != b["linear"].view(np.uint32)).any(-1)
    assert not diff.any(), np.argwhere(diff)[:8]
    diff = (a["linear"].view(np.uint32) != c["linear"].view(np.uint32)).any(-1)
    assert not diff.any(), np.argwhere(diff)[:8]
    print(name, "rays %d, reference-order %.0f ms, culled %.0f ms" % (a["stats"]["n_closest_rays"] + a["stats"]["n_shadow_rays"],
                                                                   a["stats"]["kernel_ms"], b["stats"]["kernel_ms"]))


def test_crop_tiles_and_ranks_do_not_change_pixels(gpu, assets):
    w, h, spp = 200, 120, 3
    s = gpu.Scene("config5", aspect=w / h)
    full = s.render(w, h, spp, seed=4)
    crop = s.render(w, h, spp, seed=4, crop=(37, 11, 150, 97))
    assert np.array_equal(crop["linear"].view(np.uint32), full["linear"][11:97, 37:150].view(np.uint32))
    other = s.render(w, h, spp, seed=4, tile=(16, 8))
    assert np.array_equal(other["linear"].view(np.uint32), full["linear"].view(np.uint32))
    for n_ranks in (2, 3, 8):
        acc = np.zeros_like(full["gamma"])
        cnt = 0
        for r in range(n_ranks):
            part = s.render(w, h, spp, seed=4, rank=r, n_ranks=n_ranks, tile=(32, 32))
            touched = (part["gamma"] != 0).any(-1)
            assert not (touched & (acc != 0).any(-1)).any()      # ranks own disjoint tiles
            acc += part["gamma"]
            cnt += part["stats"]["n_tiles"]
        assert cnt == ((w + 31) // 32) * ((h + 31) // 32)
        assert np.array_equal(acc.view(np.uint32), full["gamma"].view(np.uint32))


def test_chunked_render_equals_oracle_crop(gpu, ref, assets, monkeypatch):
    """Several chunks (4 Mi paths each, forced through the tuning override the library reads per call); compare a
    window with the oracle at the same full size."""
    w, h, spp = 1024, 576, 32
    s = gpu.Scene("random_spheres", aspect=w / h)
    monkeypatch.setenv("HAI719_CHUNK_LOG2", "22")
    out = s.render(w, h, spp, seed=9, stats=True)
    assert out["stats"]["n_samples"] == w * h * spp and out["stats"]["n_chunks"] == 5
    a = ref.scene("random_spheres", aspect=w / h)
    want = a.render(w, h, spp, seed=9, crop=(500, 300, 532, 316), want_ids=False)
    a.close()
    got = {"linear": out["linear"][300:316, 500:532], "gamma": out["gamma"][300:316, 500:532]}
    print(compare(got, want))
    r = out["stats"]
    rays_per_sample = (r["n_closest_rays"] + r["n_shadow_rays"]) / r["n_samples"]
    assert 5.0 < rays_per_sample < 20.0          # 10.4 on this scene
    assert r["n_sphere_tests"] == 82 * (r["n_closest_rays"] + r["n_shadow_rays"]) or r["n_sphere_tests"] > 0


@pytest.mark.parametrize("name", ["random_spheres", "config5", "backrooms_pool"])
def test_chunk_size_does_not_change_pixels(gpu, assets, monkeypatch, name):
    """Chunks of 64 Ki paths (dozens of chunks, queues and counters reused) against one chunk, wavefront and state machine."""
    w, h, spp = 320, 180, 6
    s = gpu.Scene(name, aspect=w / h)
    one = s.render(w, h, spp, seed=12, variant=6, stats=True)
    assert one["stats"]["n_chunks"] == 1
    monkeypatch.setenv("HAI719_CHUNK_LOG2", "16")
    for v in (6, 5, 3):
        many = s.render(w, h, spp, seed=12, variant=v, stats=True)
        assert many["stats"]["n_chunks"] >= 5
        assert np.array_equal(one["linear"].view(np.uint32), many["linear"].view(np.uint32)), v
        for k in ("n_closest_rays", "n_shadow_rays", "n_random"):
            assert one["stats"][k] == many["stats"][k], (v, k)


def test_full_size_config2_window_matches_oracle(gpu, ref, assets):
    """BASELINE config 2 at its real size (1920x1080, 64 spp): a 24x16 window rendered by both sides."""
    w, h, spp = 1920, 1080, 64
    crop = (948, 600, 972, 616)
    a = ref.scene("random_spheres", aspect=w / h)
    want = a.render(w, h, spp, seed=0, crop=crop, want_ids=False)
    a.close()
    got = gpu.Scene("random_spheres", aspect=w / h).render(w, h, spp, seed=0, crop=crop)
    print(compare(got, want, min_bitexact=0.99))


def test_full_size_config3_window_matches_oracle(gpu, ref, assets):
    w, h, spp = 3840, 2160, 16
    crop = (2300, 1100, 2316, 1108)
    a = ref.scene("flamingo_pond", aspect=w / h)
    want = a.render(w, h, spp, seed=0, crop=crop, want_ids=False)
    a.close()
    got = gpu.Scene("flamingo_pond", aspect=w / h).render(w, h, spp, seed=0, crop=crop)
    print(compare(got, want, min_bitexact=0.99))


def test_same_seed_same_bits_and_ppm(gpu, assets, tmp_path):
    s = gpu.Scene("cornell_box", aspect=850 / 480.0)
    a = s.ray_trace_from_camera(170, 96, 2, seed=1, ppm_path=str(tmp_path / "rendu.ppm"))
    b = s.ray_trace_from_camera(170, 96, 2, seed=1)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    tok = open(tmp_path / "rendu.ppm").read().split()
    assert tok[:4] == ["P3", "170", "96", "255"] and len(tok) == 4 + 170 * 96 * 3
    want = (255.0 * np.minimum(1.0, a.astype(np.float32))).astype(np.int32).ravel()     # main.cpp:260
    assert np.array_equal(np.array(tok[4:], dtype=np.int32), want)


@pytest.mark.parametrize("name", ["cornell_box", "random_spheres", "backrooms_pool"])
def test_output_stage_rgb8_p3_p6(gpu, assets, tmp_path, name):
    """8-bit quantisation on the device. The bytes equal (int)(255.f*min(1.f,c)) of the float render
    (main.cpp:258), the P3 file written from them is byte-identical to the reference-format file written from floats,
    and the P6 file holds the same bytes. Crops and rank shards quantise the same pixels."""
    w, h, spp = 170, 96, 2
    s = gpu.Scene(name, aspect=w / h)
    f = s.render(w, h, spp, seed=3)["gamma"]
    want = (255.0 * np.minimum(1.0, f.astype(np.float32))).astype(np.int32)
    assert want.min() >= 0 and want.max() <= 255
    b = s.render_rgb8(w, h, spp, seed=3)
    assert b.dtype == np.uint8 and np.array_equal(b.astype(np.int32), want)
    crop = (20, 10, 150, 90)
    assert np.array_equal(s.render_rgb8(w, h, spp, seed=3, crop=crop), b[10:90, 20:150])
    shards = np.zeros_like(b)
    for r in range(3):
        part = s.render_rgb8(w, h, spp, seed=3, rank=r, n_ranks=3, tile=(16, 16))
        shards = np.maximum(shards, part)     # other ranks' tiles are left 0
    assert np.array_equal(shards, b)
    # files through the host API (default camera = the one render() uses)
    s.ray_trace_from_camera(w, h, spp, seed=3, ppm_path=str(tmp_path / "ref_p3.ppm"))
    c = s.ray_trace_from_camera_rgb8(w, h, spp, seed=3, ppm_path=str(tmp_path / "fast_p3.ppm"), p6=False)
    d = s.ray_trace_from_camera_rgb8(w, h, spp, seed=3, ppm_path=str(tmp_path / "out_p6.ppm"), p6=True)
    assert np.array_equal(c, b) and np.array_equal(d, b)
    assert open(tmp_path / "ref_p3.ppm", "rb").read() == open(tmp_path / "fast_p3.ppm", "rb").read()
    raw = open(tmp_path / "out_p6.ppm", "rb").read()
    head = b"P6\n%d %d\n255\n" % (w, h)
    assert raw.startswith(head) and raw[len(head):] == b.tobytes()


@pytest.mark.parametrize("name", ["random_spheres", "config5", "raccoon"])
def test_incremental_update_equals_fresh_upload(gpu, assets, name):
    """rt_scene_update_analytic rewrites spheres / squares / lights and their hierarchy in place (meshes and
    textures stay on the device). Frame after frame it must equal a fresh upload of the same host scene."""
    w, h, spp = 160, 90, 3
    s = gpu.Scene(name, aspect=w / h)
    first = s.render(w, h, spp, seed=5)
    bytes0 = s.device_bytes(0)
    handle0 = s.device_handle(0)
    for frame in range(3):
        s.move_sphere(1, 0.35, 0.1 * frame, -0.2)
        s.move_sphere(0, -0.2, 0.0, 0.15)
        s.update_device()
        assert s.device_handle(0) == handle0 and s.device_bytes(0) == bytes0      # same device scene, nothing re-allocated
        got = s.render(w, h, spp, seed=5)
        fresh = gpu.Scene(name, aspect=w / h)
        for f in range(frame + 1):
            fresh.move_sphere(1, 0.35, 0.1 * f, -0.2)
            fresh.move_sphere(0, -0.2, 0.0, 0.15)
        want = fresh.render(w, h, spp, seed=5)
        fresh.close()
        assert np.array_equal(got["linear"].view(np.uint32), want["linear"].view(np.uint32)), frame
        assert not np.array_equal(got["linear"].view(np.uint32), first["linear"].view(np.uint32))
    # a different number of primitives is refused
    other = gpu.Scene("cornell_box", aspect=w / h)
    assert gpu.rt.rt_scene_update_analytic(handle0, other.flatten()) == -1 and gpu.rt.rt_last_error()


@pytest.mark.parametrize("name,w,h,spp", [("cornell_box", 96, 54, 64), ("debug_refraction", 96, 54, 64), ("mesh", 64, 36, 32)])
def test_stochastic_render_agrees_with_the_unmodified_reference(gpu, assets, name, w, h, spp):
    """north_star: 'multi-spp stochastic renders must agree with the reference mean within a stated per-pixel confidence
    band'. The other side here is the reference with its OWN generator (oracle/_ref/libref_stock.so: time-seeded mt19937,
    nothing of the deterministic stream), so this checks the counter-based stream itself, not just the kernels.
    Eight independent renders on each side give a mean and a standard error per pixel and channel; the band:
      * z = (mean_gpu - mean_ref) / sqrt(se_gpu^2 + se_ref^2) has |mean| < 0.15 and standard deviation < 1.4 over the image
        (1.0 for perfectly normal estimates; path-traced pixels have heavier tails: measured 1.07-1.10 on these scenes),
      * |z| <= 4 for at least 98 % of the values (measured: 99.8-99.9 %), pixels without noise on both sides are equal,
      * the total energy of the two mean images agrees within 4 standard errors, the standard error taken from the
        spread of the per-render totals (no assumption about pixels).
    The reference side runs on ONE thread here. Round 1 compared against the 16-thread stock build and saw the GPU
    ~0.03 % brighter (+1.4 to +4.6 standard errors, always positive). Cause, measured on the CPU with 64 renders a
    side (scripts/energy_offset_probe.py, profiles/r02_notes.md): the reference's shared mt19937 is advanced by all
    scanline threads WITHOUT a lock (Functions.cpp:4-8), and that data race makes the reference itself render darker
    than its own race-free single-thread run (debug_refraction: stock16 - stock1 = -0.023 % +- 0.006 %, z = -4.0;
    deterministic - stock1 = +0.005 % +- 0.006 %, z = +1.0; Cornell: +0.03 % +- 0.18 %, z = +0.2). The counter-based
    stream agrees with the race-free generator; the offset was the reference's race, not the product's stream."""
    import oracle_ref
    if not oracle_ref.available(stock=True):
        pytest.skip("oracle/_ref/libref_stock.so not built")
    n = 12
    s = gpu.Scene(name, aspect=w / h)
    G = np.stack([s.render(w, h, spp, seed=100 + i)["linear"].astype(np.float64) for i in range(n)])
    r = oracle_ref.Ref(stock=True).scene(name, aspect=w / h)
    R = np.stack([r.render(w, h, spp, seed=i, threads=1, want_ids=False)["linear"].astype(np.float64) for i in range(n)])
    r.close()
    assert not np.array_equal(R[0], R[1])                                    # really stochastic
    mg, mr = G.mean(0), R.mean(0)
    se = np.sqrt(G.var(0, ddof=1) / n + R.var(0, ddof=1) / n)
    noisy = se > 0
    z = (mg - mr)[noisy] / se[noisy]
    tg, tr = G.sum((1, 2, 3)), R.sum((1, 2, 3))
    z_total = (tg.mean() - tr.mean()) / np.sqrt(tg.var(ddof=1) / n + tr.var(ddof=1) / n)
    print(name, "z mean %.3f std %.3f, |z|>4: %.3f %%, total-energy z %.2f" % (z.mean(), z.std(), 100 * (np.abs(z) > 4).mean(), z_total))
    assert abs(z.mean()) < 0.15 and z.std() < 1.4
    assert (np.abs(z) <= 4).mean() >= 0.98
    assert np.abs(mg - mr)[~noisy].max(initial=0.0) <= 1e-6
    assert abs(z_total) < 4.0


@pytest.mark.parametrize("name,kw", [("cornell_box", {}), ("random_spheres", {}), ("flamingo_pond", {}), ("backrooms_pool", {}),
                                     ("random_spheres", {"variant": 3}), ("config5", {"rank": 1, "n_ranks": 2, "tile": (16, 8)}),
                                     ("raccoon", {"crop": (10, 7, 93, 50)})])
def test_progressive_passes_equal_one_render(gpu, assets, name, kw):
    """rt_accum_add continues every pixel's sample sum and random streams, so passes of 1 + 2 + 3 samples
    give the bits of ONE render at 6 spp - floats before and after gamma and the 8-bit output - whatever the kernel,
    the rectangle or the sharding; after a reset the accumulator starts over."""
    w, h = 112, 63
    s = gpu.Scene(name, aspect=w / h)
    acc = s.accumulator(w, h, seed=21, **kw)
    with pytest.raises(gpu.RtError):
        acc.read()                                   # nothing accumulated yet
    total = 0
    for k in (1, 2, 3):
        total += k
        assert acc.add(k) == total
        one = s.render(w, h, total, seed=21, **kw)
        got = acc.read()
        assert np.array_equal(got["linear"].view(np.uint32), one["linear"].view(np.uint32)), (name, total)
        assert np.array_equal(got["gamma"].view(np.uint32), one["gamma"].view(np.uint32)), (name, total)
        assert np.array_equal(got["rgb8"], s.render_rgb8(w, h, total, seed=21, **kw)), (name, total)
    acc.reset()
    assert acc.samples == 0
    acc.add(2)
    two = s.render(w, h, 2, seed=21, **kw)
    assert np.array_equal(acc.read()["linear"].view(np.uint32), two["linear"].view(np.uint32))
    with pytest.raises(gpu.RtError):
        acc.add(0)
    acc.close()


def test_preview_mouse_handlers_drive_the_camera_and_restart_the_frame(gpu, assets):
    """host/Preview.h: the reference's mouse()/motion() logic (main.cpp:344-388) on the reference's Camera; passes refine
    the frame while the camera rests, any camera change restarts it, and every frame equals a one-shot render from the
    same camera at the accumulated sample count."""
    w, h = 128, 72
    s = gpu.Scene("cornell_box", aspect=w / h)
    pv = s.preview(w, h, seed=5)
    assert pv.render_pass(2) == 2 and pv.render_pass(2) == 4
    f0 = pv.frame()
    one = s.render(w, h, 4, seed=5)                  # default camera = Camera() + move(0, 0, -3.1), main.cpp:418
    assert np.array_equal(f0["gamma"].view(np.uint32), one["gamma"].view(np.uint32))
    assert np.array_equal(f0["rgb8"], s.render_rgb8(w, h, 4, seed=5))
    cams = [bytes(pv.camera())]
    # left button: trackball rotation; right button: pan; middle button: zoom
    for button, (dx, dy) in ((0, (25, 9)), (2, (-14, 6)), (1, (0, 11))):
        pv.mouse(button, 0, 60, 30)
        pv.motion(60 + dx, 30 + dy)
        pv.mouse(button, 1, 60 + dx, 30 + dy)
        cams.append(bytes(pv.camera()))
        assert cams[-1] != cams[-2], button        # the camera moved ...
        assert pv.render_pass(3) == 3                # ... so the frame restarted
        f = pv.frame()
        ref_frame = s.render(w, h, 3, seed=5, camera=pv.camera())
        assert np.array_equal(f["gamma"].view(np.uint32), ref_frame["gamma"].view(np.uint32)), button
        assert not np.array_equal(f["rgb8"], f0["rgb8"])
    pv.motion(5, 5)                                  # no button held: nothing happens
    assert bytes(pv.camera()) == cams[-1] and pv.render_pass(1) == 4
    pv.invalidate()
    assert pv.render_pass(1) == 1
    pv.resize(96, 54)
    assert pv.render_pass(2) == 2
    f = pv.frame()
    assert f["rgb8"].shape == (54, 96, 3)
    assert np.array_equal(f["gamma"].view(np.uint32), s.render(96, 54, 2, seed=5, camera=pv.camera())["gamma"].view(np.uint32))
    pv.close()


def test_device_output_and_untile(gpu, assets):
    """rt_render_device + rt_untile_device with torch-owned device buffers (what bench.py's multi-GPU path does)."""
    torch = pytest.importorskip("torch")
    w, h, spp, n_ranks = 160, 100, 2, 2
    s = gpu.Scene("random_spheres", aspect=w / h)
    full = s.render(w, h, spp, seed=6)
    cam = gpu.default_camera(w, h)
    packed, offsets = [], [0]
    stream = torch.cuda.current_stream().cuda_stream
    for r in range(n_ranks):
        p = gpu.render_params(w, h, spp, seed=6, rank=r, n_ranks=n_ranks)
        n = gpu.rt.rt_render_pixel_count(C.byref(p))
        buf = torch.zeros(n * 3, dtype=torch.float32, device="cuda:0")
        st = gpu.RtStats()
        rc = gpu.rt.rt_render_device(s.device_handle(0), C.byref(cam), C.byref(p), buf.data_ptr(), None, stream, C.byref(st))
        assert rc == 0, gpu.rt.rt_last_error()
        packed.append(buf)
        offsets.append(offsets[-1] + n)
    allp = torch.cat(packed)
    img = torch.zeros(h * w * 3, dtype=torch.float32, device="cuda:0")
    off = np.array(offsets[:-1], np.int64)
    p = gpu.render_params(w, h, spp, seed=6, rank=0, n_ranks=n_ranks)
    assert gpu.rt.rt_untile_device(C.byref(p), allp.data_ptr(), off.ctypes.data, img.data_ptr(), 0, stream) == 0
    torch.cuda.synchronize()
    assert np.array_equal(img.cpu().numpy().reshape(h, w, 3).view(np.uint32), full["gamma"].view(np.uint32))


def test_bad_arguments_are_rejected(gpu, assets):
    s = gpu.Scene("single_square")
    for kw in ({"spp": 0}, {"max_bounces": 17}, {"nb_ech": 0}, {"crop": (5, 5, 3, 9)}, {"crop": (0, 0, 999, 4)},
               {"rank": 2, "n_ranks": 2}):
        args = dict(width=32, height=18, spp=1)
        args.update(kw)
        crop = args.pop("crop", None)
        p = gpu.render_params(args.pop("width"), args.pop("height"), args.pop("spp"), crop=crop, **args)
        cam = gpu.default_camera(32, 18)
        buf = np.zeros((64, 64, 3), np.float32)
        rc = gpu.rt.rt_render(s.device_handle(0), C.byref(cam), C.byref(p), buf.ctypes.data, None, None)
        assert rc == -1, (kw, rc)
        assert gpu.rt.rt_last_error()
    d = s.flatten().contents
    d.abi_version = 99
    out = C.c_void_p()
    assert gpu.rt.rt_scene_create(C.byref(d), 0, C.byref(out)) == -1
    d.abi_version = 3
    assert gpu.rt.rt_scene_create(C.byref(d), 64, C.byref(out)) == -1


def test_max_bounces_and_shadow_samples_are_runtime(gpu, assets):
    """MAXBOUNCES / NB_ECH are compile-time in the reference (Constants.h:11-12); here they are parameters.
    One bounce = direct light only, so the image must differ from the 6-bounce one but keep its hit mask."""
    s = gpu.Scene("single_square")
    a = s.render(64, 36, 2, max_bounces=6)["linear"]
    b = s.render(64, 36, 2, max_bounces=1)["linear"]
    c = s.render(64, 36, 2, max_bounces=6, nb_ech=3)["linear"]
    assert not np.array_equal(a, b) and not np.array_equal(a, c)
    assert np.isfinite(b).all()
