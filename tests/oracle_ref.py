"""ctypes binding of the oracle (the reference itself, built headless by oracle/Makefile).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. The product package never imports this module.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSETS = os.path.join(ROOT, "assets", "_ref")
LIB_DET = os.path.join(ROOT, "oracle", "_ref", "libref_det.so")
LIB_STOCK = os.path.join(ROOT, "oracle", "_ref", "libref_stock.so")
LIB_COUNT = os.path.join(ROOT, "oracle", "_ref", "libref_count.so")   # deterministic build that counts the reference's ray calls (never timed)
LIB_GLIBC = os.path.join(ROOT, "oracle", "_ref", "libref_glibc.so")   # deterministic build without oracle/libm_pin.cpp
KINDS = {"det": LIB_DET, "stock": LIB_STOCK, "count": LIB_COUNT, "glibc": LIB_GLIBC}

# scene ids: main.cpp:421-432 order; 11 = setup_flamingo_lake (unregistered); 100 = config 5
SCENES = {
    "single_sphere": 0, "single_square": 1, "cornell_box": 2, "mesh": 3, "rt_in_a_weekend": 4,
    "random_spheres": 5, "debug_refraction": 6, "flamingo": 7, "raccoon": 8, "flamingo_pond": 9,
    "backrooms_pool": 10, "flamingo_lake": 11, "config5": 100,
}


def available(stock=False, kind=None):
    path = KINDS[kind] if kind else (LIB_STOCK if stock else LIB_DET)
    return os.path.exists(path) and os.path.isdir(os.path.join(ASSETS, "mesh"))


class Ref:
    def __init__(self, stock=False, kind=None):
        self.kind = kind or ("stock" if stock else "det")
        self.lib = C.CDLL(KINDS[self.kind])
        L = self.lib
        L.ref_ray_counts.argtypes = [C.c_void_p]
        L.ref_render_rows.restype = C.c_double
        L.ref_render_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ref_scene_create.restype = C.c_void_p
        L.ref_scene_create.argtypes = [C.c_int, C.c_float, C.c_uint32, C.c_char_p]
        L.ref_scene_destroy.argtypes = [C.c_void_p]
        L.ref_scene_dump.restype = C.c_size_t
        L.ref_scene_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.ref_camera.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_render.restype = C.c_double
        L.ref_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_trace_rays.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p]
        L.ref_shade_rays.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                     C.c_void_p]

    def scene(self, name, aspect=850.0 / 480.0, seed=0):
        sid = SCENES[name] if isinstance(name, str) else int(name)
        h = self.lib.ref_scene_create(sid, aspect, seed, ASSETS.encode())
        if not h:
            raise RuntimeError("oracle could not build scene %r" % (name,))
        return RefScene(self, h)

    def camera(self, w, h):
        mv = np.zeros(16, np.float64)
        pr = np.zeros(16, np.float64)
        dr = np.zeros(2, np.float64)
        self.lib.ref_camera(w, h, mv.ctypes.data, pr.ctypes.data, dr.ctypes.data)
        return mv, pr, dr


class RefScene:
    def __init__(self, ref, handle):
        self.ref = ref
        self.h = handle

    def close(self):
        if self.h:
            self.ref.lib.ref_scene_destroy(self.h)
            self.h = None

    def dump(self):
        n = self.ref.lib.ref_scene_dump(self.h, None, 0)
        out = np.zeros(n, np.uint32)
        self.ref.lib.ref_scene_dump(self.h, out.ctypes.data, n)
        return out

    def render(self, w, h, spp, seed=0, threads=0, crop=None, want_ids=True):
        x0, y0, x1, y1 = crop if crop else (0, 0, w, h)
        cw, ch = x1 - x0, y1 - y0
        lin = np.zeros((ch, cw, 3), np.float32)
        gam = np.zeros((ch, cw, 3), np.float32)
        ids = np.zeros((ch, cw, 4), np.uint32) if want_ids else None
        nrand = C.c_uint64(0)
        secs = self.ref.lib.ref_render(self.h, w, h, spp, seed, threads, x0, y0, x1, y1, lin.ctypes.data,
                                       gam.ctypes.data, ids.ctypes.data if want_ids else None, C.byref(nrand))
        out = {"linear": lin, "gamma": gam, "ids": ids, "seconds": secs, "n_random": nrand.value}
        if self.ref.kind == "count":
            c = np.zeros(2, np.uint64)
            self.ref.lib.ref_ray_counts(c.ctypes.data)
            out["n_closest_rays"], out["n_shadow_rays"] = int(c[0]), int(c[1])
        return out

    def render_rows(self, w, h, spp, crop=None, want_image=False):
        """The reference's own threading (main.cpp:229-238): one std::thread per scanline of the crop, all at once;
        jitter from a thread_local mt19937. With the stock library random_float() is the shared racy generator."""
        x0, y0, x1, y1 = crop if crop else (0, 0, w, h)
        gam = np.zeros((y1 - y0, x1 - x0, 3), np.float32) if want_image else None
        secs = self.ref.lib.ref_render_rows(self.h, w, h, spp, x0, y0, x1, y1, gam.ctypes.data if want_image else None)
        return {"gamma": gam, "seconds": secs}

    def trace_rays(self, org, dirs, time=None):
        org = np.ascontiguousarray(org, np.float32)
        dirs = np.ascontiguousarray(dirs, np.float32)
        n = org.shape[0]
        t = None if time is None else np.ascontiguousarray(time, np.float32)
        out = np.zeros((n, 4), np.uint32)
        aux = np.zeros((n, 8), np.float32)
        self.ref.lib.ref_trace_rays(self.h, n, org.ctypes.data, dirs.ctypes.data,
                                    t.ctypes.data if t is not None else None, out.ctypes.data, aux.ctypes.data)
        return out, aux

    def shade_rays(self, org, dirs, time=None, seed=0):
        org = np.ascontiguousarray(org, np.float32)
        dirs = np.ascontiguousarray(dirs, np.float32)
        n = org.shape[0]
        t = None if time is None else np.ascontiguousarray(time, np.float32)
        rgb = np.zeros((n, 3), np.float32)
        self.ref.lib.ref_shade_rays(self.h, n, org.ctypes.data, dirs.ctypes.data,
                                    t.ctypes.data if t is not None else None, seed, rgb.ctypes.data)
        return rgb
