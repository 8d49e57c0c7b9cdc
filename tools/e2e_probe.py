#!/usr/bin/env python3
"""Where does an end-to-end step go? upload (rt_scene_create) / render (device) / D2H, wall clock per phase with a
synchronize after each.    python tools/e2e_probe.py c1 [iterations]"""
import ctypes as C, importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import WORKLOADS
hb = importlib.import_module("hai719-raytracing_b200")
wl = WORKLOADS[sys.argv[1]]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
w, h, spp = wl["w"], wl["h"], wl["spp"]
s = hb.Scene(wl["scene"], aspect=w / h, seed=0)
cam = hb.default_camera(w, h)
img = torch.zeros(h * w * 3, dtype=torch.float32, device="cuda:0")
host = torch.empty(h * w * 3, dtype=torch.float32).pin_memory()
p = hb.render_params(w, h, spp, seed=0)
stream = torch.cuda.current_stream().cuda_stream
acc = [0.0, 0.0, 0.0, 0.0]
for it in range(n + 5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.invalidate_device()
    handle = s.device_handle(0)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    rc = hb.rt.rt_render_device_image(handle, C.byref(cam), C.byref(p), img.data_ptr(), None, stream, None)
    assert rc == 0
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    host.copy_(img, non_blocking=True)
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    if it >= 5:
        for k, d in enumerate((t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
            acc[k] += d * 1e3 / n
print("%s: upload %.3f ms, render submit %.3f ms, render wait %.3f ms, D2H %.3f ms, sum %.3f ms (scene bytes %d)"
      % (sys.argv[1], acc[0], acc[1], acc[2], acc[3], sum(acc), s.device_bytes(0)))
