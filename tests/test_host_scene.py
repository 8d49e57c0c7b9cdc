"""Host API (GL-free C++ Scene/Camera/Mesh/KDTree/loaders + flatten) against the reference. CPU only.

The strongest check is the canonical dump: every float of every vertex, KD node, leaf list, material
and image hash that the tracer reads must equal, bit for bit, what the reference's own classes hold after
the same setup_*() call (oracle/ref_driver.cpp : dump_scene)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ALL_SCENES, ROOT


def test_libraries_export_every_declared_symbol(hb):
    for lib, names in ((hb.rt, hb.RT_SYMBOLS), (hb.host, hb.HOST_SYMBOLS)):
        for n in names:
            assert hasattr(lib, n), n
    # and the headers declare nothing else
    import re
    for header, names, prefix in (("hai719_rt.h", hb.RT_SYMBOLS, "rt_"), ("hai719_host.h", hb.HOST_SYMBOLS, "hai_")):
        src = open(os.path.join(ROOT, "include", header)).read()
        declared = set(re.findall(r"\b(%s[a-z_0-9]+)\s*\(" % prefix, src))
        assert declared == set(names), (declared ^ set(names))
    assert hb.rt.rt_abi_version() == 3


def test_struct_sizes_match_the_c_header(hb, tmp_path):
    """ctypes mirrors vs sizeof() from the real header, via a tiny C program."""
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include "hai719_rt.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                    "sizeof(RtMaterial),sizeof(RtSphere),sizeof(RtSquare),sizeof(RtLight),sizeof(RtImage),sizeof(RtKdNode),"
                    "sizeof(RtTriRef),sizeof(RtSceneMesh),sizeof(RtSceneDesc),sizeof(RtCamera),sizeof(RtRenderParams),sizeof(RtStats));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(t) for t in (hb.RtMaterial, hb.RtSphere, hb.RtSquare, hb.RtLight, hb.RtImage, hb.RtKdNode,
                                  hb.RtTriRef, hb.RtSceneMesh, hb.RtSceneDesc, hb.RtCamera, hb.RtRenderParams, hb.RtStats)]
    assert got == want


@pytest.mark.parametrize("name", ALL_SCENES)
def test_scene_dump_equals_reference(hb, ref, assets, name):
    a = ref.scene(name, aspect=850 / 480.0, seed=0)
    da = a.dump()
    a.close()
    db = hb.Scene(name, aspect=850 / 480.0, seed=0).dump()
    assert da.size == db.size
    assert np.array_equal(da, db)


@pytest.mark.parametrize("seed", [1, 12345])
def test_random_scene_is_seeded_like_the_oracle(hb, ref, assets, seed):
    for name in ("random_spheres", "config5"):
        a = ref.scene(name, seed=seed)
        da = a.dump()
        a.close()
        assert np.array_equal(da, hb.Scene(name, seed=seed).dump())
    assert not np.array_equal(hb.Scene("random_spheres", seed=seed).dump(), hb.Scene("random_spheres", seed=seed + 1).dump())


@pytest.mark.parametrize("wh", [(850, 480), (1920, 1080), (3840, 2160), (7680, 4320), (96, 54), (33, 77)])
def test_camera_matrices_equal_reference(hb, ref, wh):
    mv, pr, dr = ref.camera(*wh)
    cam = hb.default_camera(*wh)
    assert np.array_equal(mv, np.array(cam.modelview_inverse))
    assert np.array_equal(pr, np.array(cam.projection_inverse))
    assert cam.depth_near == dr[0] == 0.0


def test_cornell_geometry_follows_aspect_ratio(hb, assets):
    a = hb.Scene("cornell_box", aspect=1.0).dump()
    b = hb.Scene("cornell_box", aspect=2.0).dump()
    assert a.size == b.size and not np.array_equal(a, b)


def test_kd_tree_shapes_match_survey(hb, assets):
    """topology of the trees as built in the scenes."""
    s = hb.Scene("flamingo")
    assert s.kd_stats(0) == {"nodes": 1, "leaves": 1, "empty_leaves": 0, "refs": 832, "max_leaf": 832, "max_depth": 0}
    s = hb.Scene("flamingo_pond")
    k = s.kd_stats(0)
    assert (k["nodes"], k["refs"], k["max_depth"]) == (3575, 53772, 100)
    k = s.kd_stats(1)
    assert (k["nodes"], k["leaves"], k["refs"], k["max_leaf"], k["max_depth"]) == (161, 81, 2637, 106, 11)
    s = hb.Scene("backrooms_pool")
    assert s.counts()["squares"] == 28 and s.counts()["meshes"] == 3 and s.counts()["lights"] == 0
    assert (s.kd_stats(0)["nodes"], s.kd_stats(1)["nodes"], s.kd_stats(2)["nodes"]) == (5801, 671, 429)


def test_flatten_is_consistent(hb, assets):
    s = hb.Scene("flamingo_pond")
    d = s.flatten().contents
    assert d.abi_version == 3 and d.n_meshes == 2 and d.n_squares == 1 and d.n_lights == 1
    for i in range(d.n_meshes):
        m = d.meshes[i]
        nodes = np.ctypeslib.as_array(C.cast(m.nodes, C.POINTER(C.c_uint32)), shape=(m.n_nodes, 10))
        skip, first, nrefs, leaf = nodes[:, 3], nodes[:, 7], nodes[:, 8], nodes[:, 9]
        idx = np.arange(m.n_nodes)
        assert (skip > idx).all() and (skip <= m.n_nodes).all()
        assert (skip[leaf == 1] == idx[leaf == 1] + 1).all()
        # leaf ranges tile the ref array in order
        lf = np.nonzero(leaf == 1)[0]
        assert (first[lf] == np.concatenate([[0], np.cumsum(nrefs[lf])[:-1]])).all()
        assert nrefs[lf].sum() == m.n_leaf_refs


MESH_SCENE_FILE = """hai719scene 1
# Scene::setup_mesh (Scene.h / host Scene.cpp), restated as a scene description file
sky image img/textures/space.ppm
light 0 3 2
material green kd 0.1 0.6 0.2
material mir   mirror kd 0.8 0.8 0.8
material water glass kd 0.1 0.2 0.5 ior 1.333 transparency 0.9
material white kd 1 1 1
material black kd 0 0 0
material floor kd 0.8 0.8 0
sphere green 0 0 -16 2
sphere mir 4 0 -8 2
mesh water mesh/blob-closed.off translate 0 0.9 -4 scale 1.5 1.5 1.5 rotate_x 180 rotate_y 180
sphere white 0.2 -1 -4.8 0.3      # eye
sphere black 0.2 -1 -4.55 0.1     # pupil
sphere white -0.7 -1 -4.95 0.3
sphere black -0.7 -1 -4.7 0.1
square floor  -1 -0.2 0  1 0 0  0 1 0  2 2  translate 0 0 -2 scale 50 50 1 rotate_x -90
"""

REFRACTION_SCENE_FILE = """hai719scene 1
sky gradient
light -1 8 2
material red kd 1 0 0
material green kd 0 1 0
material blue kd 0 0 1
material white kd 1 1 1
material lens glass kd 1 1 1 ior 1.4 transparency 1
square red   -1 -1 0  1 0 0  0 1 0  2 2  scale 2 2 1 translate -2  2 -2
square green -1 -1 0  1 0 0  0 1 0  2 2  scale 2 2 1 translate -2 -2 -2
square blue  -1 -1 0  1 0 0  0 1 0  2 2  scale 2 2 1 translate  2  2 -2
square white -1 -1 0  1 0 0  0 1 0  2 2  scale 2 2 1 translate  2 -2 -2
sphere lens 0 0 0 0.75
"""


@pytest.mark.parametrize("builtin,text", [("mesh", MESH_SCENE_FILE), ("debug_refraction", REFRACTION_SCENE_FILE)])
def test_scene_file_rebuilds_a_builtin_scene_bit_for_bit(hb, assets, tmp_path, builtin, text):
    """a scene description file goes through the same host calls as the reference's hard-coded
    builders, so restating a builder in the file format must give the identical canonical dump (every vertex, KD node,
    leaf list, material, image hash)."""
    f = tmp_path / "scene.txt"
    f.write_text(text)
    want = hb.Scene(builtin).dump()
    got = hb.Scene().load_file(str(f)).dump()
    assert got.size == want.size and np.array_equal(got, want)


@pytest.mark.parametrize("text,needle", [
    ("", ":0: empty file"),
    ("hai719scene 2\n", ":1: expected the header"),
    ("hai719scene 1\nsphere nope 0 0 0 1\n", ":2: unknown material 'nope'"),
    ("hai719scene 1\nmaterial m kd 1 0\n", ":2: missing kd"),
    ("hai719scene 1\nmaterial m kd 1 0 x\n", "is not a number"),
    ("hai719scene 1\nmaterial m image 0 1 1\n", ":2: texture index out of range"),
    ("hai719scene 1\nmaterial m\nmesh m mesh/does_not_exist.off\n", ":3:"),
    ("hai719scene 1\nmaterial m\nsquare m 0 0 0 1 0 0 0 1 0 1 1 spin 3\n", ":3: unknown transform 'spin'"),
    ("hai719scene 1\nfrobnicate\n", ":2: unknown statement 'frobnicate'"),
])
def test_scene_file_errors_carry_line_numbers_and_never_exit(hb, assets, tmp_path, text, needle):
    f = tmp_path / "bad.txt"
    f.write_text(text)
    with pytest.raises(hb.RtError) as e:
        hb.Scene().load_file(str(f))
    assert needle in str(e.value), str(e.value)


def test_missing_mesh_is_an_error_not_an_exit(hb, tmp_path):
    s = hb.Scene(assets=str(tmp_path))
    with pytest.raises(hb.RtError):
        s.setup("flamingo")


def test_loaders_handle_format_variants(hb, tmp_path):
    """OFF plain / face-coloured / COFF, blank line after header; PPM P3 + P6 with comments."""
    os.makedirs(tmp_path / "mesh")
    os.makedirs(tmp_path / "img" / "textures")
    # the flamingo scene loads mesh/flamingo_lowpoly_colored.off: feed it a COFF tetrahedron
    (tmp_path / "mesh" / "flamingo_lowpoly_colored.off").write_text(
        "COFF\n4 4 0\n0 0 0 255 0 0 255 \n1 0 0 0 255 0 255 \n0 1 0 0 0 255 255 \n0 0 1 255 255 255 255 \n"
        "3 0 1 2 \n3 0 1 3 \n3 0 2 3 \n3 1 2 3 \n")
    s = hb.Scene(assets=str(tmp_path))
    s.setup("flamingo")
    d = s.flatten().contents
    m = d.meshes[0]
    assert (m.n_vertices, m.n_triangles, m.color_type) == (4, 4, 0)
    vc = np.ctypeslib.as_array(m.vert_colors, shape=(4, 3))
    assert np.array_equal(vc[0], np.array([1, 0, 0], np.float32)) and np.array_equal(vc[3], np.ones(3, np.float32))
    # face colours, detected from the first face line
    (tmp_path / "mesh" / "flamingo_lowpoly_colored.off").write_text(
        "OFF\n4 4 0\n0 0 0\n1 0 0\n0 1 0\n0 0 1\n3 0 1 2 255 0 0\n3 0 1 3 0 255 0\n3 0 2 3 0 0 255\n3 1 2 3 51 102 153\n")
    s.setup("flamingo")
    m = s.flatten().contents.meshes[0]
    assert m.color_type == 1
    fc = np.ctypeslib.as_array(m.face_colors, shape=(4, 3))
    assert np.allclose(fc[3], np.array([51, 102, 153]) / 255.0, atol=1e-7)
    # PPM: P3 with a comment line, as sky for single_sphere
    (tmp_path / "img" / "textures" / "space.ppm").write_text("P3\n# made by hand\n2 2\n255\n255 0 0  0 255 0\n0 0 255  9 8 7\n")
    s.setup("single_sphere")
    d = s.flatten().contents
    assert (d.skybox.w, d.skybox.h) == (2, 2)
    px = np.ctypeslib.as_array(C.cast(d.skybox.rgb, C.POINTER(C.c_uint8)), shape=(4, 3))
    assert px.tolist() == [[255, 0, 0], [0, 255, 0], [0, 0, 255], [9, 8, 7]]
    # P6 with a comment after the magic
    with open(tmp_path / "img" / "textures" / "space.ppm", "wb") as f:
        f.write(b"P6\n# c\n2 1\n255\n" + bytes([1, 2, 3, 4, 5, 6]))
    s.setup("single_sphere")
    d = s.flatten().contents
    assert (d.skybox.w, d.skybox.h) == (2, 1)
    # unreadable image: empty, not an error (reference prints and carries on)
    os.remove(tmp_path / "img" / "textures" / "space.ppm")
    s.setup("single_sphere")
    assert s.flatten().contents.skybox.w == 0


def test_scene_check_accepts_every_builtin_scene(hb, assets):
    for name in ["cornell_box", "random_spheres", "flamingo_pond", "backrooms_pool", "flamingo_lake", "config5"]:
        s = hb.Scene(name, aspect=16 / 9)
        assert hb.rt.rt_scene_check(s.flatten()) == 0, hb.rt.rt_last_error()


def test_scene_check_rejects_inconsistent_descriptions(hb, assets):
    """rt_scene_check = the checks rt_scene_create runs before anything reaches the device: every index the kernels
    follow (material images, KD skip links, leaf ranges, leaf-ref and triangle vertex indices) must be in range."""
    s = hb.Scene("flamingo_pond", aspect=16 / 9)
    d = s.flatten().contents
    m = d.meshes[0]
    leaf = next(k for k in range(m.n_nodes) if m.nodes[k].is_leaf)
    inner = next(k for k in range(m.n_nodes) if not m.nodes[k].is_leaf)

    def rejected(obj, field, value, needle, index=None):
        old = getattr(obj, field) if index is None else getattr(obj, field)[index]
        if index is None: setattr(obj, field, value)
        else: getattr(obj, field)[index] = value
        rc = hb.rt.rt_scene_check(C.byref(d))
        msg = hb.rt.rt_last_error().decode()
        if index is None: setattr(obj, field, old)
        else: getattr(obj, field)[index] = old
        assert rc == -1 and needle in msg, (field, rc, msg)

    rejected(d, "abi_version", d.abi_version + 1, "abi_version")
    rejected(m.material, "image", d.n_textures, "image index")
    rejected(m.material, "type", 7, "enum")
    rejected(d.squares[0].material, "normal_map", d.n_normal_maps + 3, "image index")
    rejected(m.nodes[leaf], "first_ref", m.n_leaf_refs, "leaf range")
    rejected(m.nodes[inner], "skip", m.n_nodes + 1, "skip link")
    rejected(m.nodes[inner], "skip", inner, "skip link")
    rejected(m.leaf_refs[5], "tri_index", m.n_triangles, "triangle index")
    rejected(m.leaf_refs[5], "v", m.n_vertices, "vertex index", index=1)
    rejected(m, "triangles", m.n_vertices, "triangle vertex index", index=3 * (m.n_triangles - 1) + 2)
    assert hb.rt.rt_scene_check(C.byref(d)) == 0     # everything restored
    assert hb.rt.rt_scene_check(None) == -1


_OFF4 = "4 {nt} 0\n0 0 0\n1 0 0\n0 1 0\n0 0 1\n"


@pytest.mark.parametrize("text,needle", [
    ("OFF\n" + _OFF4.format(nt=1) + "3 0 1 9\n", "names vertex 9 of 4"),
    ("OFF\n" + _OFF4.format(nt=1) + "3 0 1 -2\n", "names vertex -2"),
    ("OFF\n4 1 0\n0 0 0\n1 0 0\n", "do not fit"),
    ("OFF\n" + _OFF4.format(nt=3) + "3 0 1 2\n", "face 1 of 3 is missing"),
    ("OFF\n2000000000 2000000000 0\n0 0 0\n", "do not fit"),
    ("OFF\n-4 -1 0\n", "do not fit"),
    ("PLY\n4 1 0\n", "not an OFF file"),
    ("", "not an OFF file"),
    ("OFF\n4 1 0\n0 0 zero\n1 0 0\n0 1 0\n0 0 1\n3 0 1 2\n", "vertex list"),
])
def test_malformed_off_is_an_error_not_a_crash(hb, tmp_path, text, needle):
    """files on which Mesh::loadOFF (Mesh.cpp:9-74) would index out of bounds or allocate blindly."""
    os.makedirs(tmp_path / "mesh")
    (tmp_path / "mesh" / "flamingo_lowpoly_colored.off").write_text(text)
    s = hb.Scene(assets=str(tmp_path))
    with pytest.raises(hb.RtError) as e:
        s.setup("flamingo")
    assert needle in str(e.value)


@pytest.mark.parametrize("data", [
    b"P6\n4 4\n255\n\x01\x02\x03", b"P3\n4 4\n255\n1 2 3 4", b"P6\n2000000000 2000000000\n255\n", b"P6\n-4 4\n255\n",
    b"P6\n0 0\n255\n", b"P9\n2 2\n255\n", b"P3\n1 1\n255\nred green blue\n", b"",
])
def test_malformed_ppm_loads_as_no_image(hb, tmp_path, data):
    """ppmLoader::load_ppm (imageLoader.cpp:21-103) prints and carries on with an empty image; the hardened loader does
    the same for pixel blocks that end early and sizes that cannot fit in the file, without allocating for them."""
    os.makedirs(tmp_path / "img" / "textures")
    (tmp_path / "img" / "textures" / "space.ppm").write_bytes(data)
    s = hb.Scene(assets=str(tmp_path))
    s.setup("single_sphere")
    d = s.flatten().contents
    assert (d.skybox.w, d.skybox.h) == (0, 0)


def _png_decode(data):
    """Minimal PNG reader for the test: checks signature and every chunk CRC, inflates the IDAT stream with zlib."""
    import struct
    import zlib
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, []
    while pos < len(data):
        n, typ = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        (crc,) = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])
        assert crc == zlib.crc32(typ + body) & 0xFFFFFFFF, typ
        chunks.append((typ, body))
        pos += 12 + n
    assert chunks[0][0] == b"IHDR" and chunks[-1] == (b"IEND", b"")
    w, h, depth, colour, comp, filt, lace = struct.unpack(">IIBBBBB", chunks[0][1])
    assert (depth, colour, comp, filt, lace) == (8, 2, 0, 0, 0)
    idat = [b for t, b in chunks if t == b"IDAT"]
    assert all(len(b) <= 1 << 20 for b in idat)
    raw = np.frombuffer(zlib.decompress(b"".join(idat)), np.uint8).reshape(h, 1 + 3 * w)
    assert not raw[:, 0].any()      # filter type 0 on every row
    return raw[:, 1:].reshape(h, w, 3), len(idat)


@pytest.mark.parametrize("wh", [(1, 1), (5, 3), (213, 103), (700, 520)])
def test_output_stage_writes_png_p6_and_p3(hb, tmp_path, wh):
    """the same 8-bit values as PNG (stored blocks: sizes below and above one 65535-byte block and one
    1 MiB IDAT chunk), binary P6, and the reference's P3 text (main.cpp:252-262)."""
    w, h = wh
    rgb = np.random.default_rng(w * 1000 + h).integers(0, 256, (h, w, 3), dtype=np.uint8)
    png = tmp_path / "a.png"
    hb.write_image_rgb8(png, rgb, "png")
    back, n_idat = _png_decode(png.read_bytes())
    assert np.array_equal(back, rgb)
    assert n_idat == (2 if w * h * 3 > (1 << 20) else 1)
    try:
        from PIL import Image
        assert np.array_equal(np.asarray(Image.open(png).convert("RGB")), rgb)
    except ImportError:
        pass
    p6 = tmp_path / "a.ppm"
    hb.write_image_rgb8(p6, rgb, "p6")
    assert p6.read_bytes() == b"P6\n%d %d\n255\n" % (w, h) + rgb.tobytes()
    p3 = tmp_path / "b.ppm"
    hb.write_image_rgb8(p3, rgb, "p3")
    assert p3.read_bytes() == b"P3\n%d %d\n255\n" % (w, h) + b"".join(b"%d " % v for v in rgb.ravel()) + b"\n"
    with pytest.raises(hb.RtError):
        hb.write_image_rgb8(tmp_path / "no_such_dir" / "a.png", rgb, "png")


@pytest.mark.parametrize("wh", [(1, 1), (7, 5), (213, 103)])
def test_exr_writer_round_trips_floats(hb, tmp_path, wh):
    """lossless fp32 output. The file is parsed here by the published layout (magic, attributes, offset
    table, one scanline per block, channels B G R) and, where OpenCV was built with OpenEXR, read back with it."""
    import struct
    w, h = wh
    rng = np.random.default_rng(w + h)
    rgb = (rng.standard_normal((h, w, 3)) * 10.0 ** rng.integers(-6, 6, (h, w, 3))).astype(np.float32)
    rgb[0, 0] = (0.0, 1.0, 65504.0)
    path = tmp_path / "a.exr"
    hb.write_exr(path, rgb)
    d = path.read_bytes()
    assert struct.unpack("<II", d[:8]) == (20000630, 2)
    pos, attrs = 8, {}
    while d[pos] != 0:
        e = d.index(b"\0", pos); name = d[pos:e].decode(); pos = e + 1
        e = d.index(b"\0", pos); typ = d[pos:e].decode(); pos = e + 1
        (n,) = struct.unpack("<I", d[pos:pos + 4]); pos += 4
        attrs[name] = (typ, d[pos:pos + n]); pos += n
    pos += 1
    assert attrs["compression"] == ("compression", b"\0") and attrs["lineOrder"] == ("lineOrder", b"\0")
    assert struct.unpack("<4i", attrs["dataWindow"][1]) == (0, 0, w - 1, h - 1) == struct.unpack("<4i", attrs["displayWindow"][1])
    ch = attrs["channels"][1]
    assert [ch[18 * i:18 * i + 1] for i in range(3)] == [b"B", b"G", b"R"] and ch[-1] == 0 and len(ch) == 55
    assert all(struct.unpack("<i4xii", ch[18 * i + 2:18 * i + 18]) == (2, 1, 1) for i in range(3))
    offsets = struct.unpack("<%dQ" % h, d[pos:pos + 8 * h])
    back = np.zeros_like(rgb)
    for y, off in enumerate(offsets):
        yy, n = struct.unpack("<ii", d[off:off + 8])
        assert (yy, n) == (y, 12 * w)
        planes = np.frombuffer(d[off + 8:off + 8 + n], "<f4").reshape(3, w)
        back[y] = planes[::-1].T
    assert offsets[-1] + 8 + 12 * w == len(d)
    assert np.array_equal(back.view(np.uint32), rgb.view(np.uint32))
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    try:
        import cv2
        img = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
    except Exception:
        img = None
    if img is not None:
        assert img.dtype == np.float32 and np.array_equal(img[:, :, ::-1].view(np.uint32), rgb.view(np.uint32))
    with pytest.raises(hb.RtError):
        hb.write_exr(tmp_path / "no_such_dir" / "a.exr", rgb)


def test_preview_and_accumulator_need_a_device_too(hb, assets):
    """SURVEY 8(f)-4 entry points: same rule, no device - an error with a text, never a CPU render."""
    if hb.device_count() > 0:
        pytest.skip("a GPU is present")
    s = hb.Scene("single_square")
    with pytest.raises(hb.RtError) as e:
        s.accumulator(16, 9)
    assert e.value.status == -2 or "no CUDA device" in str(e.value)
    with pytest.raises(hb.RtError) as e:
        s.preview(16, 9)
    assert "no CUDA device" in str(e.value) or "device" in str(e.value)
    # null handles are rejected, not dereferenced
    assert hb.rt.rt_accum_add(None, None, 1, None) == -1 and hb.rt.rt_accum_read(None, None, None, None) == -1
    assert hb.rt.rt_accum_samples(None) == 0
    hb.rt.rt_accum_destroy(None)
    assert hb.host.hai_preview_pass(None, 1, None) != 0 and hb.host.hai_preview_motion(None, 0, 0) != 0
    hb.host.hai_preview_free(None)


def test_no_gpu_means_error_not_fallback(hb, assets):
    """On the CPU-only build box the render path must fail loudly."""
    if hb.device_count() > 0:
        pytest.skip("a GPU is present")
    s = hb.Scene("single_square")
    with pytest.raises(hb.RtError) as e:
        s.render(16, 9, 1)
    assert e.value.status == -2 or "no CUDA device" in str(e.value)
