"""hai719-raytracing_b200 — ctypes bindings of the B200 render path.

Two in-tree shared libraries (built by ``make -C hai719-raytracing_b200`` or ``__graft_entry__.build()``):

* ``lib/libhai719_rt.so``   hand-written sm_100a CUDA kernels behind the C ABI of ``include/hai719_rt.h``
* ``lib/libhai719_host.so`` the GL-free C++ host API (Scene / Camera / Mesh / Material / KDTree, OFF and PPM
  loaders, flatten, ``ray_trace_from_camera``) behind ``include/hai719_host.h``

This module only marshals arguments. There is no Python or CPU implementation of the render path: if the
libraries are missing, importing raises; if no B200 is present, every render call raises ``RtError``.
The package directory name contains a hyphen, so import it with
``importlib.import_module("hai719-raytracing_b200")``.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_RT = os.path.join(HERE, "lib", "libhai719_rt.so")
LIB_HOST = os.path.join(HERE, "lib", "libhai719_host.so")
DEFAULT_ASSETS = os.path.join(ROOT, "assets", "_ref")

# scene ids of Scene::setup_by_id (main.cpp:421-432 order; 11 = flamingo_lake; 100 = BASELINE config 5)
SCENES = {
    "single_sphere": 0, "single_square": 1, "cornell_box": 2, "mesh": 3, "rt_in_a_weekend": 4,
    "random_spheres": 5, "debug_refraction": 6, "flamingo": 7, "raccoon": 8, "flamingo_pond": 9,
    "backrooms_pool": 10, "flamingo_lake": 11, "config5": 100,
}


class RtError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("hai719_rt error %d: %s" % (status, msg))
        self.status = status


class RtMaterial(C.Structure):
    _fields_ = [("type", C.c_int32), ("texture_type", C.c_int32), ("diffuse", C.c_float * 3),
                ("transparency", C.c_float), ("index_medium", C.c_float), ("checker1", C.c_float * 3),
                ("checker2", C.c_float * 3), ("texture_scale_x", C.c_float), ("texture_scale_y", C.c_float),
                ("emissive", C.c_int32), ("light_color", C.c_float * 3), ("light_intensity", C.c_float),
                ("image", C.c_int32), ("normal_map", C.c_int32), ("motion", C.c_float * 3)]


class RtSphere(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("radius", C.c_float), ("material", RtMaterial)]


class RtSquare(C.Structure):
    _fields_ = [("v0", C.c_float * 3), ("v1", C.c_float * 3), ("v3", C.c_float * 3), ("right", C.c_float * 3),
                ("up", C.c_float * 3), ("material", RtMaterial)]


class RtLight(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("radius", C.c_float), ("color", C.c_float * 3)]


class RtImage(C.Structure):
    _fields_ = [("w", C.c_int32), ("h", C.c_int32), ("rgb", C.c_void_p), ("content_id", C.c_uint64)]


class RtKdNode(C.Structure):
    _fields_ = [("bmin", C.c_float * 3), ("skip", C.c_uint32), ("bmax", C.c_float * 3), ("first_ref", C.c_uint32),
                ("n_refs", C.c_uint32), ("is_leaf", C.c_uint32)]


class RtTriRef(C.Structure):
    _fields_ = [("v", C.c_uint32 * 3), ("tri_index", C.c_uint32)]


class RtSceneMesh(C.Structure):
    _fields_ = [("n_vertices", C.c_uint32), ("n_triangles", C.c_uint32), ("positions", C.POINTER(C.c_float)),
                ("triangles", C.POINTER(C.c_uint32)), ("color_type", C.c_int32),
                ("vert_colors", C.POINTER(C.c_float)), ("face_colors", C.POINTER(C.c_float)),
                ("root_bmin", C.c_float * 3), ("root_bmax", C.c_float * 3), ("n_nodes", C.c_uint32),
                ("nodes", C.POINTER(RtKdNode)), ("n_leaf_refs", C.c_uint32), ("leaf_refs", C.POINTER(RtTriRef)),
                ("material", RtMaterial)]


class RtSceneDesc(C.Structure):
    _fields_ = [("abi_version", C.c_uint32),
                ("n_spheres", C.c_uint32), ("spheres", C.POINTER(RtSphere)),
                ("n_squares", C.c_uint32), ("squares", C.POINTER(RtSquare)),
                ("n_meshes", C.c_uint32), ("meshes", C.POINTER(RtSceneMesh)),
                ("n_lights", C.c_uint32), ("lights", C.POINTER(RtLight)),
                ("n_textures", C.c_uint32), ("textures", C.POINTER(RtImage)),
                ("n_normal_maps", C.c_uint32), ("normal_maps", C.POINTER(RtImage)),
                ("skybox", RtImage), ("dark_sky", C.c_int32)]


class RtCamera(C.Structure):
    _fields_ = [("modelview_inverse", C.c_double * 16), ("projection_inverse", C.c_double * 16),
                ("depth_near", C.c_double)]


class RtRenderParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32), ("max_bounces", C.c_int32),
                ("nb_ech", C.c_int32), ("seed", C.c_uint32), ("x0", C.c_int32), ("y0", C.c_int32),
                ("x1", C.c_int32), ("y1", C.c_int32), ("rank", C.c_int32), ("n_ranks", C.c_int32),
                ("tile_w", C.c_int32), ("tile_h", C.c_int32), ("collect_stats", C.c_int32), ("variant", C.c_int32)]


class RtStats(C.Structure):
    _fields_ = [("n_samples", C.c_uint64), ("n_closest_rays", C.c_uint64), ("n_shadow_rays", C.c_uint64),
                ("n_sphere_tests", C.c_uint64), ("n_square_tests", C.c_uint64), ("n_mesh_tests", C.c_uint64),
                ("n_node_visits", C.c_uint64), ("n_tri_tests", C.c_uint64), ("n_tri_full", C.c_uint64),
                ("n_tex_fetches", C.c_uint64), ("n_random", C.c_uint64), ("kernel_ms", C.c_double),
                ("n_launches", C.c_uint32), ("n_tiles", C.c_uint32), ("n_chunks", C.c_uint32), ("reserved", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol the two headers declare — tests check the libraries export exactly these
RT_SYMBOLS = ["rt_abi_version", "rt_device_count", "rt_last_error", "rt_scene_create", "rt_scene_destroy", "rt_scene_retain",
              "rt_render_device_image", "rt_render_multi", "rt_render_multi_device", "rt_ipc_alloc", "rt_ipc_open", "rt_ipc_close", "rt_ipc_free", "rt_release_cached_memory", "rt_scene_update_analytic",
              "rt_scene_check", "rt_scene_device_bytes", "rt_scene_h2d_bytes", "rt_render_pixel_count", "rt_tile_layout", "rt_render", "rt_render_rgb8", "rt_quantize_device",
              "rt_render_device", "rt_untile_device",
              "rt_accum_create", "rt_accum_destroy", "rt_accum_reset", "rt_accum_add", "rt_accum_samples", "rt_accum_read",
              "rt_trace_primary", "rt_trace_rays", "rt_shade_rays", "rt_measure_fp32_peak"]
HOST_SYMBOLS = ["hai_last_error", "hai_scene_new", "hai_scene_free", "hai_scene_setup", "hai_scene_dump",
                "hai_scene_flatten", "hai_scene_kd_stats", "hai_scene_counts", "hai_default_camera", "hai_render",
                "hai_render_multi", "hai_ray_trace_from_camera_multi",
                "hai_scene_device", "hai_scene_invalidate_device", "hai_scene_move_sphere", "hai_scene_update_device", "hai_scene_load_file", "hai_ray_trace_from_camera", "hai_ray_trace_from_camera_rgb8", "hai_write_image_rgb8", "hai_write_exr",
                "hai_preview_new", "hai_preview_free", "hai_preview_mouse", "hai_preview_motion", "hai_preview_resize", "hai_preview_invalidate",
                "hai_preview_pass", "hai_preview_frame", "hai_preview_camera"]

if not (os.path.exists(LIB_RT) and os.path.exists(LIB_HOST)):
    raise ImportError("hai719-raytracing_b200: native libraries not built (%s). Run `make -C %s` or "
                      "`python -c 'import __graft_entry__ as g; g.build()'` — there is no Python fallback." % (LIB_RT, HERE))

rt = C.CDLL(LIB_RT)
host = C.CDLL(LIB_HOST)

rt.rt_last_error.restype = C.c_char_p
rt.rt_scene_create.argtypes = [C.POINTER(RtSceneDesc), C.c_int, C.POINTER(C.c_void_p)]
rt.rt_scene_check.argtypes = [C.POINTER(RtSceneDesc)]
rt.rt_scene_destroy.argtypes = [C.c_void_p]
rt.rt_scene_update_analytic.argtypes = [C.c_void_p, C.POINTER(RtSceneDesc)]
rt.rt_scene_device_bytes.restype = C.c_size_t
rt.rt_scene_device_bytes.argtypes = [C.c_void_p]
rt.rt_scene_h2d_bytes.restype = C.c_size_t
rt.rt_scene_h2d_bytes.argtypes = [C.c_void_p]
rt.rt_render_pixel_count.restype = C.c_int64
rt.rt_render_pixel_count.argtypes = [C.POINTER(RtRenderParams)]
rt.rt_tile_layout.restype = C.c_int64
rt.rt_tile_layout.argtypes = [C.POINTER(RtRenderParams), C.c_void_p, C.c_int64]
rt.rt_render.argtypes = [C.c_void_p, C.POINTER(RtCamera), C.POINTER(RtRenderParams), C.c_void_p, C.c_void_p,
                         C.POINTER(RtStats)]
rt.rt_render_rgb8.argtypes = [C.c_void_p, C.POINTER(RtCamera), C.POINTER(RtRenderParams), C.c_void_p, C.c_void_p]
rt.rt_quantize_device.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p]
rt.rt_render_device.argtypes = [C.c_void_p, C.POINTER(RtCamera), C.POINTER(RtRenderParams), C.c_void_p, C.c_void_p,
                                C.c_void_p, C.POINTER(RtStats)]
rt.rt_render_device_image.argtypes = rt.rt_render_device.argtypes
rt.rt_scene_retain.argtypes = [C.c_void_p]
rt.rt_scene_retain.restype = None
rt.rt_scene_destroy.restype = None
rt.rt_render_multi.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(RtCamera), C.POINTER(RtRenderParams), C.c_void_p, C.c_void_p,
                               C.POINTER(RtStats)]
rt.rt_render_multi_device.argtypes = rt.rt_render_multi.argtypes
rt.rt_ipc_alloc.argtypes = [C.c_int, C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]
rt.rt_ipc_open.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
rt.rt_ipc_close.argtypes = [C.c_int, C.c_void_p]
rt.rt_ipc_free.argtypes = [C.c_int, C.c_void_p]
rt.rt_untile_device.argtypes = [C.POINTER(RtRenderParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
rt.rt_trace_primary.argtypes = [C.c_void_p, C.POINTER(RtCamera), C.POINTER(RtRenderParams), C.c_void_p]
rt.rt_trace_rays.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
rt.rt_shade_rays.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(RtRenderParams),
                             C.c_void_p]

rt.rt_measure_fp32_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
rt.rt_accum_create.argtypes = [C.c_void_p, C.POINTER(RtRenderParams), C.POINTER(C.c_void_p)]
rt.rt_accum_destroy.argtypes = [C.c_void_p]
rt.rt_accum_reset.argtypes = [C.c_void_p]
rt.rt_accum_add.argtypes = [C.c_void_p, C.POINTER(RtCamera), C.c_int32, C.POINTER(RtStats)]
rt.rt_accum_samples.restype = C.c_uint32
rt.rt_accum_samples.argtypes = [C.c_void_p]
rt.rt_accum_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]

host.hai_last_error.restype = C.c_char_p
host.hai_scene_new.restype = C.c_void_p
host.hai_scene_new.argtypes = [C.c_char_p]
host.hai_scene_free.argtypes = [C.c_void_p]
host.hai_scene_setup.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_uint32]
host.hai_scene_dump.restype = C.c_size_t
host.hai_scene_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
host.hai_scene_flatten.restype = C.POINTER(RtSceneDesc)
host.hai_scene_flatten.argtypes = [C.c_void_p]
host.hai_scene_kd_stats.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
host.hai_scene_counts.argtypes = [C.c_void_p, C.c_void_p]
host.hai_default_camera.argtypes = [C.c_int, C.c_int, C.POINTER(RtCamera)]
host.hai_render.argtypes = [C.c_void_p, C.c_int, C.POINTER(RtCamera), C.POINTER(RtRenderParams), C.c_void_p,
                            C.c_void_p, C.POINTER(RtStats)]
host.hai_render_multi.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.POINTER(RtCamera), C.POINTER(RtRenderParams), C.c_void_p,
                                  C.c_void_p, C.POINTER(RtStats)]
host.hai_ray_trace_from_camera_multi.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_char_p,
                                                 C.c_void_p]
host.hai_scene_device.restype = C.c_void_p
host.hai_scene_device.argtypes = [C.c_void_p, C.c_int]
host.hai_scene_invalidate_device.argtypes = [C.c_void_p]
host.hai_ray_trace_from_camera.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_char_p,
                                           C.c_void_p]
host.hai_scene_load_file.argtypes = [C.c_void_p, C.c_char_p]
host.hai_preview_new.restype = C.c_void_p
host.hai_preview_new.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32]
host.hai_preview_free.argtypes = [C.c_void_p]
host.hai_preview_mouse.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
host.hai_preview_motion.argtypes = [C.c_void_p, C.c_int, C.c_int]
host.hai_preview_resize.argtypes = [C.c_void_p, C.c_int, C.c_int]
host.hai_preview_invalidate.argtypes = [C.c_void_p]
host.hai_preview_pass.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint32)]
host.hai_preview_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
host.hai_preview_camera.argtypes = [C.c_void_p, C.POINTER(RtCamera)]
host.hai_scene_move_sphere.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float]
host.hai_scene_update_device.argtypes = [C.c_void_p]
host.hai_ray_trace_from_camera_rgb8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_char_p,
                                                C.c_int, C.c_void_p]
host.hai_write_image_rgb8.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
host.hai_write_exr.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p]


def _rt_check(rc):
    if rc != 0:
        raise RtError(rc, rt.rt_last_error().decode(errors="replace"))


IMAGE_FORMATS = {"p3": 0, "p6": 1, "png": 2}


def write_image_rgb8(path, rgb8, fmt="png"):
    """Output stage on its own (host, no GPU): h x w x 3 uint8 -> P3 text (the reference's file), binary P6 or PNG."""
    a = np.ascontiguousarray(rgb8, dtype=np.uint8)
    if a.ndim != 3 or a.shape[2] != 3:
        raise ValueError("rgb8 must be h x w x 3")
    _host_check(host.hai_write_image_rgb8(str(path).encode(), IMAGE_FORMATS[fmt], a.shape[1], a.shape[0], a.ctypes.data))


def write_exr(path, rgb):
    """h x w x 3 float32 -> uncompressed scanline OpenEXR (host, no GPU)."""
    a = np.ascontiguousarray(rgb, dtype=np.float32)
    if a.ndim != 3 or a.shape[2] != 3:
        raise ValueError("rgb must be h x w x 3")
    _host_check(host.hai_write_exr(str(path).encode(), a.shape[1], a.shape[0], a.ctypes.data))


def _host_check(rc):
    if rc != 0:
        raise RtError(rc, host.hai_last_error().decode(errors="replace"))


def device_count():
    """Number of sm_100 devices the CUDA library can use (0 on a CPU-only box)."""
    return int(rt.rt_device_count())


def measure_fp32_peak(device=0):
    """(unfused FMUL+FADD, fused FFMA) Tflop/s measured on `device`."""
    a, b = C.c_double(0), C.c_double(0)
    _rt_check(rt.rt_measure_fp32_peak(device, C.byref(a), C.byref(b)))
    return a.value, b.value


def render_params(width, height, spp, max_bounces=6, nb_ech=10, seed=0, crop=None, rank=0, n_ranks=1, tile=(0, 0),
                  collect_stats=False, variant=0):
    p = RtRenderParams()
    p.width, p.height, p.spp, p.max_bounces, p.nb_ech, p.seed = width, height, spp, max_bounces, nb_ech, seed
    if crop:
        p.x0, p.y0, p.x1, p.y1 = crop
    p.rank, p.n_ranks, p.tile_w, p.tile_h = rank, n_ranks, tile[0], tile[1]
    p.collect_stats, p.variant = int(collect_stats), variant
    return p


def tile_layout(params):
    """(n, 4) int32 array {x0, y0, w, h} of the tiles `params.rank` renders, in packed order (host arithmetic only)."""
    n = rt.rt_tile_layout(C.byref(params), None, 0)
    if n < 0:
        raise RtError(-1, rt.rt_last_error().decode(errors="replace"))
    out = np.zeros((n, 4), np.int32)
    rt.rt_tile_layout(C.byref(params), out.ctypes.data, n)
    return out


def default_camera(width, height):
    """The reference's start-up camera (Camera() + move(0,0,-3.1), main.cpp:418) for a width x height window."""
    cam = RtCamera()
    _host_check(host.hai_default_camera(width, height, C.byref(cam)))
    return cam


class Scene:
    """A host-side Scene (C++ object) built by one of the reference's setup_*() builders."""

    def __init__(self, name=None, aspect=850.0 / 480.0, seed=0, assets=DEFAULT_ASSETS):
        self.h = host.hai_scene_new(assets.encode() if assets else None)
        if name is not None:
            self.setup(name, aspect, seed)

    def setup(self, name, aspect=850.0 / 480.0, seed=0):
        sid = SCENES[name] if isinstance(name, str) else int(name)
        _host_check(host.hai_scene_setup(self.h, sid, aspect, seed))
        return self

    def load_file(self, filename):
        """Build the scene from a scene description file (host/SceneFile.cpp)."""
        _host_check(host.hai_scene_load_file(self.h, filename.encode()))
        return self

    def close(self):
        if self.h:
            host.hai_scene_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dump(self):
        n = host.hai_scene_dump(self.h, None, 0)
        out = np.zeros(n, np.uint32)
        host.hai_scene_dump(self.h, out.ctypes.data, n)
        return out

    def counts(self):
        o = np.zeros(8, np.uint32)
        host.hai_scene_counts(self.h, o.ctypes.data)
        return dict(zip(["spheres", "squares", "meshes", "lights", "textures", "normals", "sky_w", "sky_h"], o.tolist()))

    def kd_stats(self, mesh):
        o = np.zeros(6, np.uint64)
        _host_check(host.hai_scene_kd_stats(self.h, mesh, o.ctypes.data))
        return dict(zip(["nodes", "leaves", "empty_leaves", "refs", "max_leaf", "max_depth"], o.tolist()))

    def flatten(self):
        d = host.hai_scene_flatten(self.h)
        if not d:
            raise RtError(-1, host.hai_last_error().decode(errors="replace"))
        return d

    def invalidate_device(self):
        host.hai_scene_invalidate_device(self.h)

    def move_sphere(self, index, dx, dy, dz):
        _host_check(host.hai_scene_move_sphere(self.h, index, dx, dy, dz))

    def update_device(self):
        """Push the host scene's spheres / squares / lights to its device copies in place (no mesh / texture upload)."""
        _host_check(host.hai_scene_update_device(self.h))

    def h2d_bytes(self, device=0):
        """Bytes the last upload of this scene to `device` really copied (cached images are not copied again)."""
        return int(rt.rt_scene_h2d_bytes(self.device_handle(device)))

    def device_bytes(self, device=0):
        return int(rt.rt_scene_device_bytes(self.device_handle(device)))

    def device_handle(self, device=0):
        h = host.hai_scene_device(self.h, device)
        if not h:
            raise RtError(-1, host.hai_last_error().decode(errors="replace"))
        return h

    # ---- render path (GPU only) -------------------------------------------------------------------
    def render(self, width, height, spp, seed=0, device=0, camera=None, want_linear=True, stats=False, **kw):
        """hai_render(): upload (cached) + rt_render() with host output buffers. Returns a dict with
        'gamma' (the reference's `image`), 'linear' and 'stats'."""
        cam = camera or default_camera(width, height)
        p = render_params(width, height, spp, seed=seed, collect_stats=stats, **kw)
        x0, y0, x1, y1 = (p.x0, p.y0, p.x1, p.y1) if (p.x0 | p.y0 | p.x1 | p.y1) else (0, 0, width, height)
        gam = np.zeros((y1 - y0, x1 - x0, 3), np.float32)
        lin = np.zeros_like(gam) if want_linear else None
        st = RtStats()
        _host_check(host.hai_render(self.h, device, C.byref(cam), C.byref(p), gam.ctypes.data,
                                    lin.ctypes.data if lin is not None else None, C.byref(st)))
        return {"gamma": gam, "linear": lin, "stats": st.as_dict()}

    def render_multi(self, devices, width, height, spp, seed=0, camera=None, want_linear=True, stats=False, **kw):
        """hai_render_multi(): the frame on several GPUs of one box inside one call (rt_render_multi)."""
        cam = camera or default_camera(width, height)
        p = render_params(width, height, spp, seed=seed, collect_stats=stats, **kw)
        x0, y0, x1, y1 = (p.x0, p.y0, p.x1, p.y1) if (p.x0 | p.y0 | p.x1 | p.y1) else (0, 0, width, height)
        gam = np.zeros((y1 - y0, x1 - x0, 3), np.float32)
        lin = np.zeros_like(gam) if want_linear else None
        st = RtStats()
        devs = (C.c_int * len(devices))(*devices)
        _host_check(host.hai_render_multi(self.h, devs, len(devices), C.byref(cam), C.byref(p), gam.ctypes.data,
                                          lin.ctypes.data if lin is not None else None, C.byref(st)))
        return {"gamma": gam, "linear": lin, "stats": st.as_dict()}

    def ray_trace_from_camera_multi(self, devices, width, height, nsamples, seed=0, ppm_path=None):
        """ray_trace_from_camera() with RenderOptions::devices (upload to every device, render, optional P3 file)."""
        out = np.zeros((height, width, 3), np.float32)
        devs = (C.c_int * len(devices))(*devices)
        _host_check(host.hai_ray_trace_from_camera_multi(self.h, devs, len(devices), width, height, nsamples, seed,
                                                         ppm_path.encode() if ppm_path else None, out.ctypes.data))
        return out

    def trace_primary(self, width, height, seed=0, device=0, camera=None, crop=None):
        cam = camera or default_camera(width, height)
        p = render_params(width, height, 1, seed=seed, crop=crop)
        x0, y0, x1, y1 = crop if crop else (0, 0, width, height)
        ids = np.zeros((y1 - y0, x1 - x0, 4), np.uint32)
        _rt_check(rt.rt_trace_primary(self.device_handle(device), C.byref(cam), C.byref(p), ids.ctypes.data))
        return ids

    def trace_rays(self, org, dirs, time=None, device=0):
        org = np.ascontiguousarray(org, np.float32)
        dirs = np.ascontiguousarray(dirs, np.float32)
        t = None if time is None else np.ascontiguousarray(time, np.float32)
        n = org.shape[0]
        ids = np.zeros((n, 4), np.uint32)
        aux = np.zeros((n, 8), np.float32)
        _rt_check(rt.rt_trace_rays(self.device_handle(device), n, org.ctypes.data, dirs.ctypes.data,
                                   t.ctypes.data if t is not None else None, ids.ctypes.data, aux.ctypes.data))
        return ids, aux

    def shade_rays(self, org, dirs, time=None, seed=0, device=0, max_bounces=6, nb_ech=10):
        org = np.ascontiguousarray(org, np.float32)
        dirs = np.ascontiguousarray(dirs, np.float32)
        t = None if time is None else np.ascontiguousarray(time, np.float32)
        n = org.shape[0]
        rgb = np.zeros((n, 3), np.float32)
        p = render_params(1, 1, 1, max_bounces=max_bounces, nb_ech=nb_ech, seed=seed)
        _rt_check(rt.rt_shade_rays(self.device_handle(device), n, org.ctypes.data, dirs.ctypes.data,
                                   t.ctypes.data if t is not None else None, C.byref(p), rgb.ctypes.data))
        return rgb

    def render_rgb8(self, width, height, spp, seed=0, device=0, camera=None, **kw):
        """rt_render_rgb8(): the bytes the reference writes to rendu.ppm, quantised on the GPU. (h, w, 3) uint8."""
        cam = camera or default_camera(width, height)
        p = render_params(width, height, spp, seed=seed, **kw)
        x0, y0, x1, y1 = (p.x0, p.y0, p.x1, p.y1) if (p.x0 | p.y0 | p.x1 | p.y1) else (0, 0, width, height)
        out = np.zeros((y1 - y0, x1 - x0, 3), np.uint8)
        st = RtStats()
        _rt_check(rt.rt_render_rgb8(self.device_handle(device), C.byref(cam), C.byref(p), out.ctypes.data, C.byref(st)))
        return out

    def ray_trace_from_camera_rgb8(self, width, height, nsamples, seed=0, device=0, ppm_path=None, p6=True):
        """ray_trace_from_camera() with the output stage on the GPU; optional P6 (binary) or P3 (reference text) file."""
        out = np.zeros((height, width, 3), np.uint8)
        _host_check(host.hai_ray_trace_from_camera_rgb8(self.h, device, width, height, nsamples, seed,
                                                        ppm_path.encode() if ppm_path else None, int(bool(p6)), out.ctypes.data))
        return out

    def ray_trace_from_camera(self, width, height, nsamples, seed=0, device=0, ppm_path=None):
        """The whole of the reference's ray_trace_from_camera(): default camera, render, optional P3 file."""
        out = np.zeros((height, width, 3), np.float32)
        _host_check(host.hai_ray_trace_from_camera(self.h, device, width, height, nsamples, seed,
                                                   ppm_path.encode() if ppm_path else None, out.ctypes.data))
        return out

    def accumulator(self, width, height, seed=0, device=0, **kw):
        """rt_accum_*: progressive accumulation over this scene's device copy (the accumulator holds its own reference on it)."""
        return Accumulator(self, width, height, seed=seed, device=device, **kw)

    def preview(self, width, height, seed=0, device=0):
        """host/Preview.h: the reference's mouse handlers + progressive passes (the preview keeps refining the device copy it was created on)."""
        return Preview(self, width, height, seed=seed, device=device)


class Accumulator:
    """Progressive accumulation through the C ABI: add(spp) traces spp MORE samples per pixel; read()
    returns the current mean, bit-identical to one render at the total sample count."""

    def __init__(self, scene, width, height, seed=0, device=0, **kw):
        self.scene = scene
        self.params = render_params(width, height, 1, seed=seed, **kw)
        p = self.params
        x0, y0, x1, y1 = (p.x0, p.y0, p.x1, p.y1) if (p.x0 | p.y0 | p.x1 | p.y1) else (0, 0, width, height)
        self.shape = (y1 - y0, x1 - x0, 3)
        self.h = C.c_void_p()
        _rt_check(rt.rt_accum_create(scene.device_handle(device), C.byref(p), C.byref(self.h)))

    def close(self):
        if self.h:
            rt.rt_accum_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        _rt_check(rt.rt_accum_reset(self.h))

    @property
    def samples(self):
        return int(rt.rt_accum_samples(self.h))

    def add(self, spp, camera=None, stats=False):
        cam = camera or default_camera(self.params.width, self.params.height)
        st = RtStats()
        _rt_check(rt.rt_accum_add(self.h, C.byref(cam), spp, C.byref(st)))
        return st.as_dict() if stats else self.samples

    def read(self):
        gam = np.zeros(self.shape, np.float32)
        lin = np.zeros(self.shape, np.float32)
        rgb8 = np.zeros(self.shape, np.uint8)
        _rt_check(rt.rt_accum_read(self.h, gam.ctypes.data, lin.ctypes.data, rgb8.ctypes.data))
        return {"gamma": gam, "linear": lin, "rgb8": rgb8}


class Preview:
    """host/Preview.h through its C wrappers. Buttons: 0 left (rotate), 1 middle (zoom), 2 right (move); state 0 down, 1 up."""

    def __init__(self, scene, width, height, seed=0, device=0):
        self.scene = scene
        self.width, self.height = width, height
        self.h = host.hai_preview_new(scene.h, device, width, height, seed)
        if not self.h:
            raise RtError(-1, host.hai_last_error().decode(errors="replace"))

    def close(self):
        if self.h:
            host.hai_preview_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def mouse(self, button, state, x, y):
        _host_check(host.hai_preview_mouse(self.h, button, state, x, y))

    def motion(self, x, y):
        _host_check(host.hai_preview_motion(self.h, x, y))

    def resize(self, width, height):
        _host_check(host.hai_preview_resize(self.h, width, height))
        self.width, self.height = width, height

    def invalidate(self):
        _host_check(host.hai_preview_invalidate(self.h))

    def render_pass(self, spp=1):
        n = C.c_uint32()
        _host_check(host.hai_preview_pass(self.h, spp, C.byref(n)))
        return n.value

    def camera(self):
        cam = RtCamera()
        _host_check(host.hai_preview_camera(self.h, C.byref(cam)))
        return cam

    def frame(self):
        rgb8 = np.zeros((self.height, self.width, 3), np.uint8)
        gam = np.zeros((self.height, self.width, 3), np.float32)
        _host_check(host.hai_preview_frame(self.h, rgb8.ctypes.data, gam.ctypes.data))
        return {"rgb8": rgb8, "gamma": gam}
