"""Round-2 GPU parity tests: full-size windows of configs 4 and 5, multi-GPU / shared-framebuffer paths, the degenerate
known-answer rays of SURVEY 4(viii), a scene-file scene on the device, the output stage on non-finite pixels, and the
ray unit of bench.py (the product's counts against the reference's own calls)."""
import ctypes as C
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from test_gpu_parity import compare

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gpu(hb):
    if hb.device_count() < 1:
        pytest.fail("no sm_100 device: the -m gpu tests need a B200 (there is no CPU fallback to test)")
    return hb


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


# ---- full-size windows of the sharded configs (C4 / C5 were never compared at their real size) --------------
def test_full_size_config4_window_matches_oracle(gpu, ref, assets):
    """BASELINE config 4 at its real size (backrooms pool 3840x2160, 256 spp): a 16x8 window by both sides."""
    w, h, spp = 3840, 2160, 256
    crop = (1912, 1290, 1928, 1298)
    a = ref.scene("backrooms_pool", aspect=w / h)
    want = a.render(w, h, spp, seed=0, crop=crop, want_ids=False)
    a.close()
    got = gpu.Scene("backrooms_pool", aspect=w / h).render(w, h, spp, seed=0, crop=crop)
    print(compare(got, want, min_bitexact=0.99))


def test_full_size_config5_window_matches_oracle_through_two_rank_sharding(gpu, ref, assets):
    """BASELINE config 5 at its real size (7680x4320, 1024 spp; pixel index up to 3.3e7, sample index up to 1023): a 16x8
    window, rendered as TWO rank shards (8x8 tiles round-robin) whose pixels are merged, against the oracle."""
    w, h, spp = 7680, 4320, 1024
    crop = (4700, 2900, 4716, 2908)
    a = ref.scene("config5", aspect=w / h)
    want = a.render(w, h, spp, seed=0, crop=crop, want_ids=False)
    a.close()
    s = gpu.Scene("config5", aspect=w / h)
    parts = [s.render(w, h, spp, seed=0, crop=crop, rank=r, n_ranks=2, tile=(8, 8)) for r in range(2)]
    assert parts[0]["stats"]["n_tiles"] == 1 and parts[1]["stats"]["n_tiles"] == 1
    got = {k: parts[0][k] + parts[1][k] for k in ("linear", "gamma")}      # each rank leaves the other's tile at 0
    print(compare(got, want, min_bitexact=0.99))
    one = s.render(w, h, spp, seed=0, crop=crop)
    assert np.array_equal(bits(one["linear"]), bits(got["linear"]))


@pytest.mark.parametrize("name", ["mesh", "flamingo_pond", "flamingo_lake"])
def test_sky_scenes_match_the_reference_with_the_boxes_own_libm(gpu, assets, name):
    """Scene::skyboxTexture calls atan2f / asinf (Scene.h:155-156). The oracle used everywhere else pins both to the
    correctly rounded value (oracle/libm_pin.cpp); this test uses oracle/_ref/libref_glibc.so, the reference linked
    against the box's OWN glibc with no pin. Measured (tests/test_oracle.py::test_libm_pin_changes_no_pixel): the two
    oracles agree on every pixel, so the device must agree with the unpinned one inside the same bars."""
    import oracle_ref
    if not oracle_ref.available(kind="glibc"):
        pytest.skip("oracle/_ref/libref_glibc.so not built")
    w, h, spp = 320, 180, 4
    a = oracle_ref.Ref(kind="glibc").scene(name, aspect=w / h)
    want = a.render(w, h, spp, seed=0, threads=0, want_ids=False)
    a.close()
    got = gpu.Scene(name, aspect=w / h).render(w, h, spp, seed=0)
    print(name, compare(got, want, min_bitexact=0.9999))


# ---- one framebuffer, many writers ------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,w,h,spp", [("random_spheres", 200, 120, 3), ("config5", 160, 90, 2), ("cornell_box", 170, 96, 2)])
def test_image_mode_resolve_equals_packed_plus_untile(gpu, assets, name, w, h, spp):
    """rt_render_device_image (every pixel stored at its place in the row-major rectangle, rank by rank into ONE device
    image) against rt_render's own output; crops and odd tile sizes included."""
    torch = pytest.importorskip("torch")
    s = gpu.Scene(name, aspect=w / h)
    cam = gpu.default_camera(w, h)
    stream = torch.cuda.current_stream().cuda_stream
    for crop, tile, n_ranks in ((None, (32, 32), 1), (None, (32, 32), 3), ((13, 7, 150, 80), (16, 8), 2)):
        full = s.render(w, h, spp, seed=6, crop=crop)
        rh, rw = full["gamma"].shape[:2]
        img = torch.full((rh * rw * 3,), -1.0, dtype=torch.float32, device="cuda:0")
        lin = torch.full((rh * rw * 3,), -1.0, dtype=torch.float32, device="cuda:0")
        for r in range(n_ranks):
            p = gpu.render_params(w, h, spp, seed=6, crop=crop, rank=r, n_ranks=n_ranks, tile=tile)
            rc = gpu.rt.rt_render_device_image(s.device_handle(0), C.byref(cam), C.byref(p), img.data_ptr(), lin.data_ptr(), stream, None)
            assert rc == 0, gpu.rt.rt_last_error()
        torch.cuda.synchronize()
        assert np.array_equal(bits(img.cpu().numpy().reshape(rh, rw, 3)), bits(full["gamma"])), (crop, tile, n_ranks)
        assert np.array_equal(bits(lin.cpu().numpy().reshape(rh, rw, 3)), bits(full["linear"])), (crop, tile, n_ranks)


@pytest.mark.parametrize("name,w,h,spp", [("random_spheres", 320, 180, 4), ("backrooms_pool", 192, 108, 2)])
def test_render_multi_equals_single_device(gpu, assets, name, w, h, spp):
    """rt_render_multi / hai_render_multi (the drop-in for main.cpp:229-238 on a multi-GPU box): on every device count the
    box offers the image is bit-identical to the single-device render. With one visible GPU only n = 1 runs (threads,
    image-mode resolve, pooled framebuffer), the peer-mapped stores need >= 2 devices."""
    s = gpu.Scene(name, aspect=w / h)
    one = s.render(w, h, spp, seed=9, stats=True)
    counts = [n for n in (1, 2, 3, 4, 8) if n <= gpu.device_count()]
    for n in counts:
        got = s.render_multi(list(range(n)), w, h, spp, seed=9, stats=True)
        assert np.array_equal(bits(got["gamma"]), bits(one["gamma"])), n
        assert np.array_equal(bits(got["linear"]), bits(one["linear"])), n
        for k in ("n_samples", "n_closest_rays", "n_shadow_rays", "n_random"):
            assert got["stats"][k] == one["stats"][k], (n, k)
        crop = (33, 20, 150, 100)
        part = s.render_multi(list(range(n)), w, h, spp, seed=9, crop=crop)
        assert np.array_equal(bits(part["gamma"]), bits(one["gamma"][20:100, 33:150])), n
    img = s.ray_trace_from_camera_multi(list(range(counts[-1])), w, h, spp, seed=9)
    assert np.array_equal(bits(img), bits(s.ray_trace_from_camera(w, h, spp, seed=9)))
    print(name, "device counts tested:", counts)


def test_wavefront_ray_tally_equals_work_counters(gpu, assets):
    """Ray counts without the counter build (k_wf_tally: queue counters) = the counters of the STATS kernels."""
    w, h, spp = 320, 180, 6
    for name in ("random_spheres", "config5", "backrooms_pool"):
        s = gpu.Scene(name, aspect=w / h)
        a = s.render(w, h, spp, seed=3, stats=True)["stats"]
        b = s.render(w, h, spp, seed=3, stats=False)["stats"]
        assert b["n_closest_rays"] == a["n_closest_rays"] > 0, name
        assert b["n_shadow_rays"] == a["n_shadow_rays"], name


IPC_CHILD = textwrap.dedent("""
    import ctypes as C, importlib, sys
    sys.path.insert(0, %r)
    hb = importlib.import_module("hai719-raytracing_b200")
    name, w, h, spp, rank, n_ranks, dev = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7])
    handle = (C.c_ubyte * 64)(*bytes.fromhex(sys.argv[8]))
    ptr = C.c_void_p()
    assert hb.rt.rt_ipc_open(dev, handle, C.byref(ptr)) == 0, hb.rt.rt_last_error()
    s = hb.Scene(name, aspect=w / h)
    cam = hb.default_camera(w, h)
    p = hb.render_params(w, h, spp, seed=6, rank=rank, n_ranks=n_ranks, tile=(32, 32))
    st = hb.RtStats()
    rc = hb.rt.rt_render_device_image(s.device_handle(dev), C.byref(cam), C.byref(p), ptr, None, None, C.byref(st))
    assert rc == 0, hb.rt.rt_last_error()
    assert hb.rt.rt_ipc_close(dev, ptr) == 0
    print("child ok", st.n_tiles)
""")


def test_two_processes_write_one_framebuffer_through_cuda_ipc(gpu, assets):
    """bench.py's N > 1 path: rank 0 allocates the framebuffer (rt_ipc_alloc), another PROCESS maps it (rt_ipc_open) and
    its resolve kernel stores its tiles into it. Here the second process runs on the last visible device (the same one
    on a one-GPU box: the mapping and the ordering are the same, the stores then do not cross NVLink)."""
    torch = pytest.importorskip("torch")
    name, w, h, spp = "random_spheres", 200, 120, 3
    s = gpu.Scene(name, aspect=w / h)
    full = s.render(w, h, spp, seed=6)
    ptr = C.c_void_p()
    hbuf = (C.c_ubyte * 64)()
    assert gpu.rt.rt_ipc_alloc(0, h * w * 3 * 4, C.byref(ptr), hbuf) == 0, gpu.rt.rt_last_error()
    try:
        cam = gpu.default_camera(w, h)
        p = gpu.render_params(w, h, spp, seed=6, rank=0, n_ranks=2, tile=(32, 32))
        st = gpu.RtStats()
        assert gpu.rt.rt_render_device_image(s.device_handle(0), C.byref(cam), C.byref(p), ptr, None, None, C.byref(st)) == 0
        dev = gpu.device_count() - 1
        r = subprocess.run([sys.executable, "-c", IPC_CHILD % ROOT, name, str(w), str(h), str(spp), "1", "2", str(dev), bytes(hbuf).hex()],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
        assert r.returncode == 0 and "child ok" in r.stdout, r.stdout
        out = np.zeros((h, w, 3), np.float32)
        rt = C.CDLL("libcudart.so.12")
        assert rt.cudaMemcpy(C.c_void_p(out.ctypes.data), ptr, C.c_size_t(out.nbytes), 2) == 0
        assert np.array_equal(bits(out), bits(full["gamma"]))
    finally:
        assert gpu.rt.rt_ipc_free(0, ptr) == 0


# ---- degenerate known-answer rays ----------------------------------------------------------------------
def _mesh_arrays(scene, mesh):
    d = scene.flatten().contents
    m = d.meshes[mesh]
    pos = np.ctypeslib.as_array(m.positions, shape=(m.n_vertices, 3)).copy()
    tri = np.ctypeslib.as_array(m.triangles, shape=(m.n_triangles, 3)).copy()
    return pos, tri, np.array(list(m.root_bmin)), np.array(list(m.root_bmax))


def _same_hits(gpu_scene, ref_scene, org, dirs, what):
    org = np.ascontiguousarray(org, np.float32)
    dirs = np.ascontiguousarray(dirs, np.float32)
    ids, aux = gpu_scene.trace_rays(org, dirs)
    want, waux = ref_scene.trace_rays(org, dirs)
    same = (ids == want).all(-1)
    assert same.all(), (what, int((~same).sum()), ids[~same][:4], want[~same][:4])
    return ids


def test_degenerate_rays_parallel_to_a_slab(gpu, ref, assets):
    """AABB::intersects with a direction component of exactly 0 (AABB.h:50-53): 1.0/0 = inf, (lo - o) * inf = +-inf or
    NaN when the origin lies ON a face. Rays along each axis, origins inside / outside / exactly on the slab planes of the
    KD root box and of the mesh AABB, against the reference, for a depth-100 tree (pond) and a well-built one."""
    rng = np.random.default_rng(5)
    for name in ("flamingo_pond", "raccoon", "mesh"):
        g, r = gpu.Scene(name, aspect=16 / 9), ref.scene(name, aspect=16 / 9)
        for mesh in range(g.counts()["meshes"]):
            pos, tri, lo, hi = _mesh_arrays(g, mesh)
            org, dirs = [], []
            for axis in range(3):
                for sign in (1.0, -1.0):
                    d = np.zeros(3, np.float32); d[axis] = sign
                    n = 400
                    o = rng.uniform(lo - 0.2 * (hi - lo), hi + 0.2 * (hi - lo), size=(n, 3)).astype(np.float32)
                    o[:, axis] = (lo[axis] - 1.0) if sign > 0 else (hi[axis] + 1.0)
                    # a quarter of the origins sit EXACTLY on a slab plane of another axis, a quarter on a vertex coordinate
                    other = (axis + 1) % 3
                    o[: n // 8, other] = lo[other]; o[n // 8: n // 4, other] = hi[other]
                    o[n // 4: n // 2, other] = pos[rng.integers(0, len(pos), n // 4), other]
                    org.append(o); dirs.append(np.repeat(d[None], n, 0))
                    # and two zero components with the third tiny (almost parallel to two slabs)
                    d2 = np.zeros(3, np.float32); d2[axis] = sign; d2[other] = 1e-30
                    org.append(o); dirs.append(np.repeat(d2[None], n, 0))
            ids = _same_hits(g, r, np.concatenate(org), np.concatenate(dirs), (name, mesh))
            assert (ids[:, 0] == 3).sum() > 50, (name, mesh)      # the sweep really hits the mesh
        g.close(); r.close()


def test_degenerate_ray_origin_inside_a_sphere_and_on_its_surface(gpu, ref, assets):
    """Sphere::intersect takes the NEAR root only (Sphere.h:112-123): from inside a sphere that sphere is never hit, the
    ray goes on to whatever lies behind. Origins at the centres, inside, and exactly on the surfaces."""
    rng = np.random.default_rng(7)
    for name in ("random_spheres", "rt_in_a_weekend", "single_sphere"):
        g, r = gpu.Scene(name, aspect=16 / 9), ref.scene(name, aspect=16 / 9)
        d = g.flatten().contents
        c = np.array([[d.spheres[i].center[0], d.spheres[i].center[1], d.spheres[i].center[2], d.spheres[i].radius] for i in range(d.n_spheres)], np.float32)
        org, dirs = [], []
        for k in range(24):
            u = rng.normal(size=(len(c), 3)).astype(np.float32)
            u /= np.linalg.norm(u, axis=1, keepdims=True)
            v = rng.normal(size=(len(c), 3)).astype(np.float32)
            scale = (0.0, 0.5, 0.999, 1.0)[k % 4]
            org.append(c[:, :3] + scale * c[:, 3:4] * u); dirs.append(v)
        org, dirs = np.concatenate(org), np.concatenate(dirs)
        ids = _same_hits(g, r, org, dirs, name)
        inside = np.tile(np.arange(len(c)), 24)
        strictly = np.repeat(np.arange(24) % 4 < 3, len(c))
        hit_own = (ids[:, 0] == 1) & (ids[:, 1] == inside) & strictly
        assert not hit_own.any(), name                           # near-root-only: never the sphere the origin is inside of
        g.close(); r.close()


def test_degenerate_zero_area_triangles_are_never_hit(gpu, ref, assets):
    """The flamingo meshes hold 64 zero-area triangles whose normal is 0/0 = NaN (Triangle.h:26-37): no ray may report
    them, also not one aimed exactly at their vertices (FMA contraction turned 4.8 % of such pixels)."""
    for name in ("flamingo", "flamingo_pond"):
        g, r = gpu.Scene(name, aspect=16 / 9), ref.scene(name, aspect=16 / 9)
        found = 0
        for mesh in range(g.counts()["meshes"]):
            pos, tri, lo, hi = _mesh_arrays(g, mesh)
            a, b, c = pos[tri[:, 0]], pos[tri[:, 1]], pos[tri[:, 2]]
            area = np.linalg.norm(np.cross(b - a, c - a), axis=1)
            zero = np.where(area == 0)[0]
            if len(zero) == 0:
                continue
            found += len(zero)
            targets = np.concatenate([a[zero], b[zero], c[zero], (a[zero] + b[zero] + c[zero]) / 3])
            eye = np.array([0.0, 0.0, 3.1], np.float32)
            org = np.concatenate([np.repeat(eye[None], len(targets), 0), targets + np.float32([0, 2, 0]), targets + np.float32([1, 1, 1])])
            dirs = np.concatenate([targets - eye, np.repeat(np.float32([[0, -1, 0]]), len(targets), 0), np.repeat(np.float32([[-1, -1, -1]]), len(targets), 0)])
            ids = _same_hits(g, r, org, dirs, (name, mesh))
            hit = ids[:, 0] == 3
            assert not np.isin(ids[hit & (ids[:, 1] == mesh), 2], zero).any(), (name, mesh)
        assert found >= 32, (name, found)
        g.close(); r.close()


def test_degenerate_empty_leaves_and_depth_100_tree(gpu, ref, assets):
    """pond.off builds to depth 100 with empty leaves and dropped triangles: rays through the thin
    end of the tree — a dense fan from just above the mesh, grazing — must pick the reference's triangle and t."""
    g, r = gpu.Scene("flamingo_pond", aspect=16 / 9), ref.scene("flamingo_pond", aspect=16 / 9)
    stats = [g.kd_stats(m) for m in range(g.counts()["meshes"])]
    deep = int(np.argmax([s["max_depth"] for s in stats]))
    # SURVEY's "empty leaves" are nodes where KDTree::build gave up (depth 100 / equal halves, KDTree.cpp:142): no children
    # and no triangles. A full binary tree with n nodes has (n + 1) / 2 leaf positions; those not filled by a real leaf are
    # the childless non-leaf nodes (1788 - 1786 = 2 here).
    childless = (stats[deep]["nodes"] + 1) // 2 - stats[deep]["leaves"]
    assert stats[deep]["max_depth"] >= 100 and childless + stats[deep]["empty_leaves"] == 2, stats
    pos, tri, lo, hi = _mesh_arrays(g, deep)
    rng = np.random.default_rng(11)
    n = 20000
    org = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    org[:, 1] = hi[1] + rng.uniform(0.0, 0.05, n).astype(np.float32) * (hi[1] - lo[1])
    tgt = pos[rng.integers(0, len(pos), n)]
    ids = _same_hits(g, r, org, tgt - org, "pond fan")
    assert (ids[:, 0] == 3).mean() > 0.3
    g.close(); r.close()


# ---- scene file on the device (closes SURVEY 8(f)-3) ----------------------------------------------------------------------
@pytest.mark.parametrize("which", ["mesh", "debug_refraction"])
def test_scene_file_scene_renders_like_the_builtin(gpu, ref, assets, tmp_path, which):
    """A scene description file (host/SceneFile.cpp) that restates a built-in scene is uploaded and rendered on the device:
    the bits of the built-in builder's render, and the oracle's (mesh + glass + mirror + sky image / refraction)."""
    import test_host_scene as ths
    text = {"mesh": ths.MESH_SCENE_FILE, "debug_refraction": ths.REFRACTION_SCENE_FILE}[which]
    path = tmp_path / "scene.txt"
    path.write_text(text)
    w, h, spp = 160, 90, 3
    s = gpu.Scene(None)
    s.load_file(str(path))
    got = s.render(w, h, spp, seed=2)
    builtin = gpu.Scene(which, aspect=w / h).render(w, h, spp, seed=2)
    assert np.array_equal(bits(got["linear"]), bits(builtin["linear"]))
    assert np.array_equal(bits(got["gamma"]), bits(builtin["gamma"]))
    a = ref.scene(which, aspect=w / h)
    want = a.render(w, h, spp, seed=2, threads=0, want_ids=False)
    a.close()
    print(which, compare(got, want))
    ids = s.trace_primary(w, h, seed=2)
    assert (ids[..., 0] > 0).mean() > 0.2


# ---- output stage on non-finite pixels -----------------------------------------------------------------------
def test_quantize_matches_the_reference_cast_on_nan_inf_negative(gpu):
    """main.cpp:258 writes (int)(255.f * std::min<float>(1.f, c)): NaN -> 1.f -> 255 (std::min returns its first argument
    unless the second is smaller), +inf -> 255, > 1 -> 255; negative values are written as 0 (no PPM reader accepts the
    reference's minus sign). Device kernel against the host writer's arithmetic."""
    torch = pytest.importorskip("torch")
    v = np.array([np.nan, np.inf, -np.inf, -0.0, 0.0, -1e-3, 1e-9, 0.5, 0.999999, 1.0, 1.0000001, 7.0, 254.9 / 255, 1 / 255.0, 0.99 / 255], np.float32)
    d = torch.from_numpy(v).cuda()
    out = torch.zeros(len(v), dtype=torch.uint8, device="cuda")
    assert gpu.rt.rt_quantize_device(d.data_ptr(), len(v), out.data_ptr(), 0, torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()
    want = []
    for c in v:
        m = np.float32(255.0) * (np.float32(1.0) if not (c < np.float32(1.0)) else c)     # std::min<float>(1.f, c)
        want.append(int(m) if m > 0 else 0)
    assert out.cpu().numpy().tolist() == want, (out.cpu().numpy().tolist(), want)
    assert want[0] == 255 and want[1] == 255 and want[2] == 0


# ---- the ray unit of bench.py ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,w,h,spp", [("random_spheres", 240, 136, 4), ("cornell_box", 240, 136, 2), ("flamingo_pond", 96, 54, 2),
                                          ("backrooms_pool", 96, 54, 2)])
def test_ray_counts_equal_the_references_own_calls(gpu, assets, name, w, h, spp):
    """bench.py divides rays by seconds on both arms. The product's rays are its kernels' counts; the reference's are
    counted INSIDE the reference (libref_count.so: gcc's function-entry hook on Scene::computeIntersection /
    computeShadow, oracle/ref_driver.cpp). Same deterministic stream => the two counts must be equal, ray for ray."""
    import oracle_ref
    if not oracle_ref.available(kind="count"):
        pytest.skip("oracle/_ref/libref_count.so not built")
    a = oracle_ref.Ref(kind="count").scene(name, aspect=w / h)
    want = a.render(w, h, spp, seed=0, threads=0, want_ids=False)
    a.close()
    st = gpu.Scene(name, aspect=w / h).render(w, h, spp, seed=0, stats=True)["stats"]
    assert st["n_closest_rays"] == want["n_closest_rays"], name
    assert st["n_shadow_rays"] == want["n_shadow_rays"], name
    assert st["n_random"] == want["n_random"], name


# ---- queue-position records, block reservation, no-light trace ---------------------------------------------------------------
@pytest.mark.parametrize("name,w,h,spp", [("random_spheres", 400, 225, 3), ("backrooms_pool", 240, 136, 2), ("config5", 240, 136, 2),
                                          ("cornell_box", 240, 136, 2), ("flamingo_pond", 240, 136, 2), ("raccoon", 160, 90, 2)])
def test_wavefront_block_sizes_and_chunking_do_not_change_a_bit(gpu, assets, name, w, h, spp, monkeypatch):
    """The wavefront stores path state at queue positions that warps reserve a block at a time (32 / 64 / 256 positions by
    chunk size) and pads the unused tail of a block with invalid entries; scenes without lights (pool, Cornell) scatter
    inside the trace kernel and alternate two live queues. Small chunks (HAI719_CHUNK_LOG2: many chunks, queues reused,
    mostly padding) and large ones must give the bits of the one-path-per-lane kernel, and the same ray counts."""
    s = gpu.Scene(name, aspect=w / h, seed=0)
    want = s.render(w, h, spp, seed=4, variant=1, stats=True)
    for log2 in (None, 16, 18):
        if log2 is None:
            monkeypatch.delenv("HAI719_CHUNK_LOG2", raising=False)
        else:
            monkeypatch.setenv("HAI719_CHUNK_LOG2", str(log2))
        for st in (False, True):
            # bit 29: scenes without lights scatter inside the trace kernel; bit 28: no separate mesh-walk kernel; bit 27: the general
            # light kernel instead of k_wf_scatter for scenes without lights
            for v in (6, 6 | (1 << 29), 6 | (1 << 28), 6 | (1 << 27), 6 | (1 << 27) | (1 << 28)):
                got = s.render(w, h, spp, seed=4, variant=v, stats=st)
                assert np.array_equal(bits(want["linear"]), bits(got["linear"])), (name, log2, st, v)
                for k in ("n_closest_rays", "n_shadow_rays") + (("n_random", "n_tex_fetches") if st else ()):
                    assert want["stats"][k] == got["stats"][k], (name, log2, st, v, k)


# ---- constant-bank scene binding, device image cache ----------------------------------------------------------------------
def test_scenes_alternating_on_one_device_and_two_streams_keep_their_own_constants(gpu, assets):
    """The kernels read the scene descriptor from ONE constant-bank copy per device (rt_core.cuh : c_scene); every render binds
    its scene on its own stream and waits for the previous binding's event. Two scenes rendered alternately, and two scenes
    enqueued back to back on two different streams without a host synchronisation in between, must each give the bits of
    their solo render."""
    torch = pytest.importorskip("torch")
    w, h, spp = 160, 90, 2
    names = ["random_spheres", "cornell_box", "flamingo_pond"]
    scenes = [gpu.Scene(n, aspect=w / h, seed=0) for n in names]
    solo = [s.render(w, h, spp, seed=5)["linear"] for s in scenes]
    for rep in range(3):
        for s, want in zip(scenes, solo):
            assert np.array_equal(bits(s.render(w, h, spp, seed=5)["linear"]), bits(want))
    cam = gpu.default_camera(w, h)
    p = gpu.render_params(w, h, spp, seed=5)
    streams = [torch.cuda.Stream() for _ in scenes]
    imgs = [torch.zeros(h * w * 3, dtype=torch.float32, device="cuda:0") for _ in scenes]
    for rep in range(3):
        for s, st, img in zip(scenes, streams, imgs):
            rc = gpu.rt.rt_render_device_image(s.device_handle(0), C.byref(cam), C.byref(p), None, img.data_ptr(), st.cuda_stream, None)
            assert rc == 0, gpu.rt.rt_last_error()
        torch.cuda.synchronize()
        for img, want in zip(imgs, solo):
            assert np.array_equal(bits(img.cpu().numpy().reshape(h, w, 3)), bits(want))


def test_device_image_cache_skips_resident_images_and_release_frees_them(gpu, assets):
    """RtImage::content_id (ABI 3): a re-upload of the Cornell box (5 images, 8.4 MB) copies only its small arrays; the same
    description with the ids zeroed copies everything; after rt_release_cached_memory the images travel again. Same pixels
    every time."""
    w, h, spp = 120, 68, 2
    s = gpu.Scene("cornell_box", aspect=w / h)
    want = s.render(w, h, spp, seed=1)["linear"]
    first = s.h2d_bytes(0)
    s.invalidate_device()
    again = s.render(w, h, spp, seed=1)["linear"]
    second = s.h2d_bytes(0)
    assert np.array_equal(bits(want), bits(again))
    assert first > 8_000_000 and second < 200_000 and s.device_bytes(0) > 8_000_000, (first, second)   # every Scene object loads (and numbers) its own files
    # ids zeroed: "contents unknown", every image is uploaded
    d = s.flatten().contents
    keep = [(d.textures[i].content_id) for i in range(d.n_textures)] + [(d.normal_maps[i].content_id) for i in range(d.n_normal_maps)]
    assert all(k != 0 for k in keep)
    for i in range(d.n_textures):
        d.textures[i].content_id = 0
    for i in range(d.n_normal_maps):
        d.normal_maps[i].content_id = 0
    out = C.c_void_p()
    assert gpu.rt.rt_scene_create(C.byref(d), 0, C.byref(out)) == 0, gpu.rt.rt_last_error()
    assert int(gpu.rt.rt_scene_h2d_bytes(out)) > 8_000_000
    gpu.rt.rt_scene_destroy(out)
    # released: the next upload copies the images again
    s.invalidate_device()
    assert gpu.rt.rt_release_cached_memory(0) == 0
    third = s.render(w, h, spp, seed=1)["linear"]
    assert np.array_equal(bits(want), bits(third))
    assert s.h2d_bytes(0) > 8_000_000
