#!/usr/bin/env python3
"""bench.py — Mrays/s of the render hot path on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--scenes c1,c3,c4,c5|none] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A STEP is one full render of a workload (every pixel x every sample: camera rays, closest-hit and shadow
traversal, shading, resolve). Ray = one Scene::computeIntersection or Scene::computeShadow call of the reference
("reference-equivalent rays": shadow rays that the kernels PROVE unoccluded without tracing them still count, as the
reference traces them; the image is bit-identical either way). Counts come from the kernels themselves.

The headline (top-level keys; K timed steps after W warm-up steps) is config 2, random spheres 1920x1080x64 — the
configuration BASELINE.json's metric is quoted on. `per_scene` then holds EVERY config of BASELINE.json at its stated
size (c1 850x480x1, c2, c3 3840x2160x16, c4 3840x2160x256, c5 7680x4320x1024), each with its own small step count
(the long ones run ONE full-size step: c5 is ~35 s on one GPU), with the same keys: value, e2e, roofline, cpu_baseline.

  value  whole-job Mrays/s, scene resident in HBM, framebuffer left in HBM (device-timed with CUDA events on the
         launching stream, barrier + synchronize on both sides, max over ranks)
  e2e    the same through the reference-facing call with HOST buffers: every step re-UPLOADS the scene
         (rt_scene_create: H2D + precompute + hierarchy build), renders and copies the framebuffer into pinned host
         memory (D2H); wall clock, max over ranks
  N > 1  one process per GPU; image tiles (32x32, round-robin) are sharded over the ranks. Rank 0 owns the framebuffer
         (rt_ipc_alloc); the other ranks map it (CUDA IPC over NVLink) and their resolve kernels store their tiles
         straight into it: no packed buffers, no gather collective, no untile pass (torch.distributed carries the
         64-byte handle and the barriers). Total work is fixed => "scaling": "strong".
  roofline      executed-work fractions of the HBM and FP32 roofs + the issue-side figure of the committed ncu capture;
                the contract-algorithm figure of SURVEY 8(d) is reported as `algorithmic_speedup`
  cpu_baseline  the reference itself (oracle/_ref) on the box's host cores on a bounded crop of the same workload:
                the deterministic build on a row pool AND the stock build threaded as the reference threads it
                (one std::thread per scanline, shared racy mt19937). A reported baseline, not the target.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# BASELINE.json configs. c2 is the headline single-GPU workload (configs[1]).
WORKLOADS = {
    "c1": dict(scene="cornell_box", w=850, h=480, spp=1, label="Cornell box 850x480 1spp (configs[0])"),
    "c2": dict(scene="random_spheres", w=1920, h=1080, spp=64, label="Random spheres seed 0, 1920x1080 64spp 6 bounces (configs[1])"),
    "c3": dict(scene="flamingo_pond", w=3840, h=2160, spp=16, label="Flamingo pond (pond.off + flamingo KD-trees, sky) 3840x2160 16spp (configs[2])"),
    "c4": dict(scene="backrooms_pool", w=3840, h=2160, spp=256, label="Backrooms pool 3840x2160 256spp (configs[3])"),
    "c5": dict(scene="config5", w=7680, h=4320, spp=1024, label="Motion-blur spheres + triceratops/gorilla 7680x4320 1024spp (configs[4])"),
}


# per-scene measurement plan: (warm-up steps, timed steps, e2e steps, warm-up spp or 0 = full). The long workloads
# warm up at reduced spp (same kernels, same chunk size, scratch allocated) and time ONE full-size step, which is
# also their e2e step (upload + render + D2H, device time from the CUDA events inside it).
PLAN = {"c1": (5, 20, 20, 0), "c2": (3, 5, 3, 0), "c3": (2, 3, 2, 4), "c4": (2, 1, 1, 8), "c5": (2, 1, 1, 1)}


def kernel_of(counts, variant, paths=1 << 30):
    """Name of the render kernel(s) rt_render_device selects (same rule as csrc/rt_capi.cu); paths = paths of one render call."""
    kind = variant & 0xFF
    n_an = counts["spheres"] + counts["squares"]
    abvh = 24 <= n_an <= 128 or (1 <= n_an < 24 and counts["lights"] > 0 and counts["meshes"] > 0)
    if kind == 0:
        kind = 6 if abvh and (n_an >= 24 or paths >= (24 << 20)) else (3 if counts["meshes"] > 0 else 1)
    if kind == 5 and not (abvh and counts["lights"] > 0):
        kind = 3
    return {1: "k_render_paths", 2: "k_render_regen<ACCEL=0>", 3: "k_render_regen<ACCEL=1> (exact culling hierarchies)",
            4: "k_render_regen<ACCEL=2> (warp-voted walk)", 5: "k_render_regen<ACCEL=3> (occluder candidates per light)",
            6: "wavefront: k_camera_rays + max_bounces x (k_wf_trace + k_wf_light) per chunk; dominant kernel k_wf_light"}[kind]


def flops_and_bytes(st):
    """SURVEY §8(d) contract formulas, from the device work counters of one step."""
    rays = st["n_closest_rays"] + st["n_shadow_rays"]
    culled = st["n_tri_tests"] - st["n_tri_full"]
    flops = (31 * st["n_sphere_tests"] + 44 * st["n_square_tests"] + 12 * st["n_node_visits"] + 41 * st["n_tri_full"]
             + 12 * culled + 150 * st["n_closest_rays"])
    byts = 64 * rays + 8 * st["n_node_visits"] + 52 * st["n_tri_tests"] + 32 * st["n_tex_fetches"]
    return rays, flops, byts


class ClockSampler:
    """SM clock and throttle reasons sampled every 20 ms (NVML) / 200 ms (nvidia-smi fallback) while the timed region runs.

    Uses NVML in-process (nvidia_ml_py): two cheap queries per sample. An `nvidia-smi -lms 200` child, as in the
    profiling recipe, was measured to add 50-200 ms of launch/synchronise latency to EVERY ~190 ms step here
    (kernel time 184.1 ms, step time 241-390 ms: profiles/r01_notes.md), so it is only the fallback."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop_flag = threading.Event()
        self.thread = None
        self.proc = None
        self.how = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.idx).uuid)
            except Exception:
                pass
            h = None
            if uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.how = "nvml"

            def loop():
                while not self.stop_flag.is_set():
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        for bit, name in self.REASONS.items():
                            if r & bit:
                                self.reasons.add(name)
                    except Exception:
                        pass
                    self.stop_flag.wait(0.02)   # NVML in-process: two cheap queries; short timed regions (8 GPUs: 45 ms) still get samples
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception:
            self.how = "nvidia-smi"
            q = ("index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            try:
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                threading.Thread(target=self._pump, daemon=True).start()
            except OSError:
                self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                self.sm.append(float(f[1])); self.mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def stop(self):
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=1.0)
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "how": self.how}



def ref_sample(wl, budget_s, ref_scene, cw_max=240, ch_max=136):
    """Centre crop + spp of the CPU sample so that one pass is ~budget_s of wall time (calibrated with a 1-spp pass of a
    240x136 crop): first the spp grows up to the workload's, then the crop up to the whole image."""
    w, h, spp = wl["w"], wl["h"], wl["spp"]
    cw, ch = min(w, cw_max), min(h, ch_max)
    crop = ((w - cw) // 2, (h - ch) // 2, (w - cw) // 2 + cw, (h - ch) // 2 + ch)
    t0 = time.perf_counter(); ref_scene.render(w, h, 1, seed=0, threads=0, crop=crop, want_ids=False); dt = time.perf_counter() - t0
    s_spp = int(max(1, min(spp, round(budget_s / max(dt, 1e-3)))))
    grow = (budget_s / max(dt * s_spp, 1e-3)) ** 0.5
    if s_spp == spp and grow > 1.2:
        cw, ch = min(w, int(cw * grow)), min(h, int(ch * grow))
        crop = ((w - cw) // 2, (h - ch) // 2, (w - cw) // 2 + cw, (h - ch) // 2 + ch)
    return crop, cw, ch, s_spp


def count_ref_rays(wl, crop, s_spp):
    """Rays of the CPU sample, counted by the reference itself: oracle/_ref/libref_count.so is the deterministic build with
    gcc's function-entry hook on Scene::computeIntersection / Scene::computeShadow (oracle/ref_driver.cpp); same stream,
    same rays as the timed deterministic build. Never timed."""
    import oracle_ref
    w, h = wl["w"], wl["h"]
    rs = oracle_ref.Ref(kind="count").scene(wl["scene"], aspect=w / h, seed=0)
    r = rs.render(w, h, s_spp, seed=0, threads=0, crop=crop, want_ids=False)
    rs.close()
    return r["n_closest_rays"] + r["n_shadow_rays"]


def cpu_baseline_of(wl, budget_s=4.0):
    """cpu_baseline object for one workload (rank 0, N = 1 only): pooled deterministic reference + the reference's own
    thread-per-scanline / shared-mt19937 mode (main.cpp:229-238, Functions.cpp:4-8)."""
    import oracle_ref
    if not oracle_ref.available():
        return None
    cores = os.cpu_count() or 1
    w, h, spp = wl["w"], wl["h"], wl["spp"]
    try:
        ref = oracle_ref.Ref().scene(wl["scene"], aspect=w / h, seed=0)
        crop, cw, ch, s_spp = ref_sample(wl, budget_s, ref)
        r = ref.render(w, h, s_spp, seed=0, threads=0, crop=crop, want_ids=False)
        dt = r["seconds"]
        ref.close()
        rays = count_ref_rays(wl, crop, s_spp) if oracle_ref.available(kind="count") else None
        out = {"value": (rays / dt / 1e6) if rays else None, "unit": "Mrays/s", "cores": cores, "kind": "reference",
               "sample": "centre crop %dx%d of the %dx%d image plane at %d of %d spp: %d samples, %s rays (counted in the reference: libref_count), "
                         "%.2f s wall; reference sources built headless (oracle/_ref), deterministic RNG shim, %d-thread row pool"
                         % (cw, ch, w, h, s_spp, spp, cw * ch * s_spp, rays, dt, cores),
               "msamples_per_s": cw * ch * s_spp / dt / 1e6, "rays_per_sample": (rays / (cw * ch * s_spp)) if rays else None}
        if oracle_ref.available(stock=True):
            st = oracle_ref.Ref(stock=True).scene(wl["scene"], aspect=w / h, seed=0)
            rr = st.render_rows(w, h, s_spp, crop=crop)
            st.close()
            # the stock stream differs from the deterministic one, so its ray count is the deterministic sample's
            # rays-per-sample x its samples (the mean over >= 3e4 samples; stated, not exact)
            out["thread_per_row_stock"] = {
                "value": (rays / rr["seconds"] / 1e6) if rays else None, "unit": "Mrays/s", "msamples_per_s": cw * ch * s_spp / rr["seconds"] / 1e6,
                "threads": ch, "cores": cores, "seconds": rr["seconds"],
                "what": "the reference's own threading and RNG: one std::thread per scanline (%d threads at once, main.cpp:229-238), "
                        "random_float() = ONE shared time-seeded mt19937 without a lock (Functions.cpp:4-8), unmodified; same crop and spp; "
                        "rays = samples x rays/sample of the deterministic sample" % ch}
        return out
    except Exception as e:  # the baseline must never take the GPU number down with it
        return {"value": None, "unit": "Mrays/s", "cores": cores, "kind": "reference", "sample": "failed: %r" % (e,)}


def run_reference(args, wl):
    """--impl reference: the reference's own CPU render (oracle/_ref, deterministic RNG shim) on all host threads, each
    step a bounded crop of the same workload. Loads nothing of the product: the crop's rays are counted by the
    reference itself (libref_count.so)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_ref
    if not oracle_ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built and assets/_ref not staged"}))
        return
    cores = os.cpu_count() or 1
    w, h, spp = wl["w"], wl["h"], wl["spp"]
    ref = oracle_ref.Ref().scene(wl["scene"], aspect=w / h, seed=0)
    per_step = max(0.5, min(3.0, 150.0 / max(1, args.steps + args.warmup)))   # the whole run ends within a few minutes
    crop, cw, ch, s_spp = ref_sample(wl, per_step, ref)
    rays = count_ref_rays(wl, crop, s_spp)
    for _ in range(args.warmup):
        ref.render(w, h, s_spp, seed=0, threads=0, crop=crop, want_ids=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.render(w, h, s_spp, seed=0, threads=0, crop=crop, want_ids=False)
    dt = (time.perf_counter() - t0) / args.steps
    samples = cw * ch * s_spp
    val = rays / dt / 1e6
    sample = ("centre crop %dx%d of the %dx%d image plane at %d of %d spp (%d samples/step, %d rays/step counted in the reference itself); the "
              "per-ray rate of a CPU whose cost per ray does not depend on the image size" % (cw, ch, w, h, s_spp, spp, samples, rays))
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/s", "value": val, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_of(wl, 1, 0, sample=sample),
        "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": cores, "kind": "reference", "sample": sample,
                         "msamples_per_s": samples / dt / 1e6},
        "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def config_of(wl, world, variant, **extra):
    c = {"workload": wl["label"], "scene": wl["scene"], "width": wl["w"], "height": wl["h"], "spp": wl["spp"], "max_bounces": 6, "nb_ech": 10,
         "seed": 0, "tiles": "32x32 round-robin over ranks" if world > 1 else "32x32", "variant": variant,
         "ray_unit": "reference-equivalent rays: computeIntersection + computeShadow calls of the reference for this image"}
    c.update(extra)
    return c


class Bench:
    """Everything one process (rank) needs to measure workloads: torch / NCCL plumbing, the shared framebuffer."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.hb = importlib.import_module("hai719-raytracing_b200")
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.hb.device_count() < 1:
            raise SystemExit("bench.py: no sm_100 device; the render path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)   # > 126 MB L2
        self.stream = torch.cuda.current_stream()
        self.fp32_peaks = None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        self.hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        self.hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        try:
            self.prof = json.load(open(os.path.join(ROOT, "profiles", "latest.json")))
        except (OSError, ValueError):
            self.prof = {}

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.item()

    def sum_over_ranks(self, xs):
        t = self.torch.tensor([float(x) for x in xs], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t)
        return [int(v) for v in t.tolist()]

    # ---- the framebuffer all ranks write into -----------------------------------------------------------------------
    def framebuffer(self, n_floats):
        """Device pointer of an n_floats image on rank 0's GPU, valid on this rank. N == 1: a torch tensor. N > 1: rank 0
        allocates with rt_ipc_alloc, the 64-byte handle travels through torch.distributed, peers map it (rt_ipc_open)."""
        torch, hb = self.torch, self.hb
        if self.world == 1:
            t = torch.zeros(n_floats, dtype=torch.float32, device=self.dev)
            return t.data_ptr(), t, None
        handle = torch.zeros(64, dtype=torch.uint8, device=self.dev)
        ptr = C.c_void_p()
        if self.rank == 0:
            hbuf = (C.c_ubyte * 64)()
            rc = hb.rt.rt_ipc_alloc(self.local, n_floats * 4, C.byref(ptr), hbuf)
            if rc != 0:
                raise RuntimeError(hb.rt.rt_last_error().decode())
            handle.copy_(torch.tensor(list(hbuf), dtype=torch.uint8))
        self.dist.broadcast(handle, src=0)
        if self.rank != 0:
            hbuf = (C.c_ubyte * 64)(*handle.cpu().tolist())
            rc = hb.rt.rt_ipc_open(self.local, hbuf, C.byref(ptr))
            if rc != 0:
                raise RuntimeError("rt_ipc_open: " + hb.rt.rt_last_error().decode())
        return ptr.value, None, ptr

    def release_framebuffer(self, keep, ptr):
        if ptr is None:
            return
        self.barrier()
        if self.rank == 0:
            self.hb.rt.rt_ipc_free(self.local, ptr)
        else:
            self.hb.rt.rt_ipc_close(self.local, ptr)
        self.barrier()

    # ---- one workload ------------------------------------------------------------------------------------------------
    def measure(self, key, wl, warmup, steps, e2e_steps, warm_spp, headline=False, with_cpu=True):
        torch, hb, args = self.torch, self.hb, self.args
        rank, world, local, dev, stream = self.rank, self.world, self.local, self.dev, self.stream
        w, h, spp = wl["w"], wl["h"], wl["spp"]
        variant = args.variant if headline else 0
        scene = hb.Scene(wl["scene"], aspect=w / h, seed=0)
        cam = hb.default_camera(w, h)
        handle = scene.device_handle(local)
        n_img = h * w * 3
        img_ptr, keep, ipc_ptr = self.framebuffer(n_img)
        host_img = torch.empty(n_img, dtype=torch.float32).pin_memory() if rank == 0 else None

        def params(c_spp, **kw):
            return hb.render_params(w, h, c_spp, seed=0, rank=rank, n_ranks=world, tile=(32, 32), **kw)

        def render(p, stats=None):
            rc = hb.rt.rt_render_device_image(handle, C.byref(cam), C.byref(p), img_ptr, None, stream.cuda_stream, stats)
            if rc != 0:
                raise RuntimeError(hb.rt.rt_last_error().decode())

        p_full = params(spp, variant=variant)
        n_px = int(hb.rt.rt_render_pixel_count(C.byref(p_full)))
        keys = ["n_samples", "n_closest_rays", "n_shadow_rays", "n_sphere_tests", "n_square_tests", "n_mesh_tests", "n_node_visits",
                "n_tri_tests", "n_tri_full", "n_tex_fetches", "n_random"]

        def count(c_variant, c_spp, cw=None, chh=None):
            # counting pass (untimed) into a scratch image of its own size; cw x chh = a smaller image plane (same camera)
            if cw is None:
                pc = params(c_spp, collect_stats=True, variant=c_variant)
                st = hb.RtStats()
                render(pc, C.byref(st))
                return st.as_dict()
            pc = hb.render_params(cw, chh, c_spp, seed=0, rank=rank, n_ranks=world, tile=(32, 32), collect_stats=True, variant=c_variant)
            st = hb.RtStats()
            tmp = torch.zeros(cw * chh * 3, dtype=torch.float32, device=dev)
            ccam = hb.default_camera(cw, chh)
            rc = hb.rt.rt_render_device_image(handle, C.byref(ccam), C.byref(pc), tmp.data_ptr(), None, stream.cuda_stream, C.byref(st))
            if rc != 0:
                raise RuntimeError(hb.rt.rt_last_error().decode())
            return st.as_dict()

        # ---- warm-up (untimed) ----
        p_warm = params(warm_spp, variant=variant) if warm_spp else p_full
        for _ in range(warmup):
            self.flush.zero_()
            render(p_warm)
        self.barrier()

        # ---- timed steps: device time by CUDA events on the launching stream, max over ranks ----
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        kst = hb.RtStats()
        kernel_ms, launches, chunks = 0.0, 0, 1
        e2e_each = []
        scene_bytes = scene.h2d_bytes(local)       # refreshed after the e2e steps: what a re-upload really copies (cached images stay)
        fused_e2e = steps == 1 and e2e_steps == 1      # long workloads: the ONE full-size step is both the device-timed and the e2e step
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if fused_e2e:
            self.barrier()
            t0 = time.perf_counter()
            scene.invalidate_device()
            handle = scene.device_handle(local)        # rt_scene_create: H2D + precompute kernels + hierarchy build
            ev0.record(stream)
            render(p_full, C.byref(kst))
            ev1.record(stream)
            torch.cuda.synchronize()
            if world > 1:
                self.dist.barrier()                    # every rank's tiles are in rank 0's framebuffer
            if rank == 0:
                if world > 1:
                    rc = C.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(C.c_void_p(host_img.data_ptr()), C.c_void_p(img_ptr), C.c_size_t(n_img * 4), 2)
                    if rc != 0:
                        raise RuntimeError("cudaMemcpy D2H failed: %d" % rc)
                else:
                    host_img.copy_(keep, non_blocking=True)
                torch.cuda.synchronize()
            e2e_each.append((time.perf_counter() - t0) * 1e3)
            self.barrier()
            kernel_ms += kst.kernel_ms; launches += kst.n_launches; chunks = max(1, kst.n_chunks)
        else:
            self.barrier()
            ev0.record(stream)
            for _ in range(steps):
                self.flush.zero_()
                render(p_full, C.byref(kst))   # stats != NULL makes the call wait on its own end event: kernel_ms is that kernel time
                kernel_ms += kst.kernel_ms; launches += kst.n_launches; chunks = max(1, kst.n_chunks)
            ev1.record(stream)
            self.barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms_per_step = self.max_over_ranks(ev0.elapsed_time(ev1)) / steps

        # ---- ray counts: the wavefront tallies them for free (k_wf_tally); other kernels need a counting pass ----
        my_closest, my_shadow = int(kst.n_closest_rays), int(kst.n_shadow_rays)
        counted = "queue counters of the timed steps (k_wf_tally)"
        if my_closest == 0:
            c = count(variant, spp)
            my_closest, my_shadow = c["n_closest_rays"], c["n_shadow_rays"]
            counted = "work counters of an untimed pass of the same deterministic render"
        rays, samples = self.sum_over_ranks([my_closest + my_shadow, n_px * spp])
        value = rays / (ms_per_step * 1e-3) / 1e6

        # ---- end to end: upload scene (H2D) + render + D2H of the framebuffer into pinned host memory, per step ----
        if not fused_e2e:
            self.barrier()
            for _ in range(e2e_steps):
                t0 = time.perf_counter()
                scene.invalidate_device()
                handle = scene.device_handle(local)
                render(p_full)
                if world > 1:
                    torch.cuda.synchronize()
                    self.dist.barrier()
                if rank == 0:
                    if world > 1:
                        rc = C.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(C.c_void_p(host_img.data_ptr()), C.c_void_p(img_ptr), C.c_size_t(n_img * 4), 2)
                        if rc != 0:
                            raise RuntimeError("cudaMemcpy D2H failed: %d" % rc)
                    else:
                        host_img.copy_(keep, non_blocking=True)
                torch.cuda.synchronize()
                e2e_each.append((time.perf_counter() - t0) * 1e3)
                if world > 1:
                    self.dist.barrier()
            self.barrier()
        scene_bytes = scene.h2d_bytes(local)
        e2e_ms = self.max_over_ranks(sum(e2e_each) / len(e2e_each))
        e2e_value = rays / (e2e_ms * 1e-3) / 1e6
        checksum = float(host_img.double().sum().item()) if rank == 0 else 0.0

        # ---- work per ray for the roofline (untimed, bounded): executed = the timed variant's own counters at <= 4 spp;
        #      contract = the reference-order traversal (variant 1) on a 960x540 image plane of the same camera at <= 4 spp ----
        c_spp = min(spp, 4)
        ex = count(variant, c_spp)
        cw, chh = min(w, 960), min(h, 540)
        alg = count(1, c_spp, cw, chh)
        ex_tot = dict(zip(keys, self.sum_over_ranks([ex[k] for k in keys])))
        alg_tot = dict(zip(keys, self.sum_over_ranks([alg[k] for k in keys])))
        self.release_framebuffer(keep, ipc_ptr)
        if rank != 0:
            return None

        ex_rays, ex_flops, ex_bytes = flops_and_bytes(ex_tot)
        al_rays, al_flops, al_bytes = flops_and_bytes(alg_tot)
        if self.fp32_peaks is None:
            self.fp32_peaks = hb.measure_fp32_peak(local)
        fp32_unfused, fp32_fused = self.fp32_peaks
        k_ms = kernel_ms / steps                                  # this rank's render kernels per step
        my_rays = my_closest + my_shadow
        t_step = k_ms * 1e-3
        ex_fpr, ex_bpr = ex_flops / max(1, ex_rays), ex_bytes / max(1, ex_rays)
        al_fpr, al_bpr = al_flops / max(1, al_rays), al_bytes / max(1, al_rays)
        hbm_ach = my_rays * ex_bpr / t_step / 1e9                 # GB/s of executed algorithmic bytes
        fp_ach = my_rays * ex_fpr / t_step / 1e12
        hbm_frac, fp_frac = hbm_ach / self.hbm_peak, fp_ach / fp32_unfused
        bound = "hbm" if hbm_frac >= fp_frac else "fp32"
        t_roof_contract = max(my_rays * al_fpr / (fp32_unfused * 1e12), my_rays * al_bpr / (self.hbm_peak * 1e9))
        prof = self.prof.get(key, {})
        roof = {
            "bound": bound,
            "achieved": hbm_ach if bound == "hbm" else fp_ach, "peak": self.hbm_peak if bound == "hbm" else fp32_unfused,
            "unit": "GB/s" if bound == "hbm" else "TFLOP/s", "frac": max(hbm_frac, fp_frac),
            "what": "EXECUTED work: the tests the timed kernels really perform (their own counters at %d spp, scaled per ray) at the per-test "
                    "flop/byte figures of SURVEY 8(d) + 64 B of path state per ray; the slower roof is reported" % c_spp,
            "hbm": {"achieved_gbs": hbm_ach, "peak_gbs": self.hbm_peak, "frac": hbm_frac, "bytes_per_ray": ex_bpr},
            "fp32": {"achieved_tflops": fp_ach, "peak_tflops": fp32_unfused, "frac": fp_frac, "flops_per_ray": ex_fpr},
            "traffic": None,
            "kernel": kernel_of(scene.counts(), variant, w * h * spp // max(1, world)),
            "kernel_ms_per_launch": k_ms / chunks, "launches_per_step": chunks,
            "launch": "one chunk of <= 32 Mi (wavefront) / 16 Mi paths: every render kernel of the chunk, CUDA events around them on the launching stream",
            "peak_source": "fp32 unfused FMUL+FADD measured live by rt_measure_fp32_peak (fused: %.1f TFLOP/s; the parity build may not fuse); hbm %s" % (fp32_fused, self.hbm_src),
            "algorithmic_speedup": {
                "value": t_roof_contract / t_step,
                "what": "time the SURVEY 8(d) CONTRACT algorithm (the reference's brute-force loops and both-children KD walk, counted by variant 1 on a "
                        "%dx%d plane at %d spp) would need at the roofline / measured kernel time; > 1 because the culling hierarchies and the light-cone "
                        "proof skip most of those tests exactly. A speed-up over the contract algorithm, NOT a utilisation figure." % (cw, chh, c_spp),
                "contract_flops_per_ray": al_fpr, "contract_bytes_per_ray": al_bpr},
        }
        if prof.get("dram_bytes_per_path") is not None:
            roof["traffic"] = prof["dram_bytes_per_path"] * (n_px * spp) / chunks
            roof["traffic_source"] = "profiles/latest.json (%s): %.0f B of DRAM traffic per path (ncu) x paths per chunk" % (prof.get("source", "?"), prof["dram_bytes_per_path"])
        if prof.get("issue") is not None:
            roof["issue"] = prof["issue"]
        out = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(wl, world, variant, rays_per_step=rays, samples_per_step=samples, rays_per_sample=rays / max(1, samples),
                                rays_counted_by=counted, warmup_spp=warm_spp or spp,
                                l2="256 MiB buffer rewritten before every step (L2 flush); path state per step is far larger than L2" if not fused_e2e
                                   else "one step of >= 10 GB of path state per chunk: inputs far larger than L2",
                                framebuffer="torch tensor" if world == 1 else "rank 0's image mapped by every rank (CUDA IPC), written by the resolve kernels over NVLink"),
            "msamples_per_s": samples / (ms_per_step * 1e-3) / 1e6,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(scene_bytes) * world,
                    "d2h_bytes_per_step": int(n_img * 4), "steps": len(e2e_each),
                    "what": "rt_scene_create (upload of everything but images already resident in the device image cache + precompute + hierarchy build) + rt_render_device_image + D2H to pinned host, per step; wall clock" +
                            ("; the same step as `value`" if fused_e2e else "")},
            "gpu_launches": int(launches) * world,
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu_baseline_of(wl) if (with_cpu and world == 1 and not args.no_cpu_baseline) else None,
            "work_executed_per_ray": {k: ex_tot[k] / max(1, ex_rays) for k in keys[3:]},
            "image_checksum": checksum,
        }
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--scenes", default="c1,c3,c4,c5", help="per_scene entries besides the headline workload ('none' = headline only)")
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel of the headline workload (diagnostics only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", type=int, default=0)
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.spp:
        wl["spp"] = args.spp
        wl["label"] += " [spp overridden to %d]" % args.spp
    if args.impl == "reference":
        return run_reference(args, wl)

    # stdout must carry exactly one JSON line: NCCL (and anything else in this process) may print banners on fd 1,
    # so fd 1 is pointed at stderr for the duration of the run and the JSON goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    b = Bench(args)
    t_start = time.perf_counter()
    warm = max(3, args.warmup)
    long_run = args.workload in ("c4", "c5") and not args.spp
    if long_run:
        w_, s_, e_, ws_ = PLAN[args.workload]
        out = b.measure(args.workload, wl, w_, s_, e_, ws_, headline=True)
    else:
        out = b.measure(args.workload, wl, warm, args.steps, max(3, min(args.steps, 5)), 0, headline=True)
    per_scene = {}
    if b.rank == 0:
        per_scene[args.workload] = {k: out[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "msamples_per_s", "e2e", "roofline", "cpu_baseline", "config", "gpu_launches", "clocks")}
    if args.scenes != "none":
        for key in [k for k in args.scenes.split(",") if k and k != args.workload]:
            w_, s_, e_, ws_ = PLAN[key]
            r = b.measure(key, dict(WORKLOADS[key]), w_, s_, e_, ws_)
            if b.rank == 0:
                per_scene[key] = {k: r[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "msamples_per_s", "e2e", "roofline", "cpu_baseline", "config", "gpu_launches", "clocks")}
            if os.environ.get("BENCH_DEBUG") and b.rank == 0:
                sys.stderr.write("%s: %.0f Mrays/s, %.1f ms/step, e2e %.0f (%.0f s since start)\n" % (key, r["value"], r["ms_per_step"], r["e2e"]["value"], time.perf_counter() - t_start))
    if b.rank == 0:
        out["per_scene"] = {k: per_scene[k] for k in sorted(per_scene)}
        out["bench_wall_s"] = time.perf_counter() - t_start
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if b.world > 1:
        b.dist.destroy_process_group()


if __name__ == "__main__":
    main()
