#!/usr/bin/env python3
"""bench.py — Mrays/s of the render hot path on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c4|c5] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A STEP is one full render of the workload (every pixel x every sample: camera rays, closest-hit and shadow
traversal, shading, resolve). Ray = one Scene::computeIntersection or Scene::computeShadow call; the count
comes from the kernels' own work counters on an untimed pass of the same deterministic render.

  value  whole-job Mrays/s, scene resident in HBM, output left in HBM (device-timed with CUDA events on
         the launching stream, barrier + synchronize on both sides, max over ranks)
  e2e    the same through the reference-facing host API with HOST buffers: every step re-flattens nothing
         but re-UPLOADS the scene (rt_scene_create, H2D), renders, gathers and copies the framebuffer to
         pinned host memory (D2H)
  N > 1  image tiles (32x32, round-robin) are sharded over the ranks; each rank renders its tiles, rank 0
         gathers the packed tiles with NCCL and untiles them. Total work is fixed => "scaling": "strong".
  roofline / cpu_baseline: see DESIGN.md §5. The cpu_baseline is the reference itself (oracle/_ref) on the
         box's host cores on a bounded crop of the same workload; it is a reported baseline, not the target.
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


def kernel_of(counts, variant):
    """Name of the render kernel(s) rt_render_device selects (same rule as csrc/rt_capi.cu)."""
    kind = variant & 0xFF
    n_an = counts["spheres"] + counts["squares"]
    abvh = 24 <= n_an <= 128 or (1 <= n_an < 24 and counts["lights"] > 0 and counts["meshes"] > 0)
    if kind == 0:
        kind = 6 if 24 <= n_an <= 128 else (3 if counts["meshes"] > 0 else 1)
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


def run_reference(args, wl):
    """--impl reference: the reference's own CPU render (oracle/_ref, deterministic RNG shim) on all host threads,
    each step a bounded crop of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_ref
    if not oracle_ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built and assets/_ref not staged"}))
        return
    hb = importlib.import_module("hai719-raytracing_b200")
    cores = os.cpu_count() or 1
    w, h, spp = wl["w"], wl["h"], wl["spp"]
    ref = oracle_ref.Ref().scene(wl["scene"], aspect=w / h, seed=0)
    cw, ch = min(w, 240), min(h, 136)
    crop = ((w - cw) // 2, (h - ch) // 2, (w - cw) // 2 + cw, (h - ch) // 2 + ch)
    # calibrate samples-per-pixel of the sample so that one step is ~3 s of wall time
    t0 = time.perf_counter(); ref.render(w, h, 1, seed=0, threads=0, crop=crop, want_ids=False); dt = time.perf_counter() - t0
    s_spp = int(max(1, min(spp, round(3.0 / max(dt, 1e-3)))))
    rays_per_sample = None
    if hb.device_count() > 0:   # ray count of exactly this crop from the GPU's work counters (identical rays: parity)
        st = hb.Scene(wl["scene"], aspect=w / h, seed=0).render(w, h, s_spp, seed=0, crop=crop, stats=True, want_linear=False)["stats"]
        rays = st["n_closest_rays"] + st["n_shadow_rays"]
    else:
        rays = None
    for _ in range(args.warmup):
        ref.render(w, h, s_spp, seed=0, threads=0, crop=crop, want_ids=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.render(w, h, s_spp, seed=0, threads=0, crop=crop, want_ids=False)
    dt = (time.perf_counter() - t0) / args.steps
    samples = cw * ch * s_spp
    if rays is None:
        rays = samples * 10.4
    val = rays / dt / 1e6
    sample = "centre crop %dx%d of the %dx%d image plane at %d of %d spp (%d samples/step, %d rays/step)" % (cw, ch, w, h, s_spp, spp, samples, rays)
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/s", "value": val, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": wl["label"], "sample": sample},
        "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": cores, "kind": "reference", "sample": sample,
                         "msamples_per_s": samples / dt / 1e6},
        "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (diagnostics only)")
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
    import torch
    import torch.distributed as dist
    hb = importlib.import_module("hai719-raytracing_b200")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if hb.device_count() < 1:
        raise SystemExit("bench.py: no sm_100 device; the render path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w, h, spp = wl["w"], wl["h"], wl["spp"]
    scene = hb.Scene(wl["scene"], aspect=w / h, seed=0)
    cam = hb.default_camera(w, h)
    handle = scene.device_handle(local)
    p = hb.render_params(w, h, spp, seed=0, rank=rank, n_ranks=world, tile=(32, 32), variant=args.variant)
    n_px = int(hb.rt.rt_render_pixel_count(C.byref(p)))
    counts = [int(hb.rt.rt_render_pixel_count(C.byref(hb.render_params(w, h, spp, rank=r, n_ranks=world, tile=(32, 32))))) for r in range(world)]
    max_px = max(counts)
    packed = torch.zeros(max_px * 3, dtype=torch.float32, device=dev)     # equal-sized so all_gather_into_tensor works
    gathered = torch.zeros(world * max_px * 3, dtype=torch.float32, device=dev) if (world > 1 and rank == 0) else None
    image = torch.zeros(h * w * 3, dtype=torch.float32, device=dev) if rank == 0 else None
    offsets = np.array([r * max_px for r in range(world)], np.int64)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2
    host_img = torch.empty(h * w * 3, dtype=torch.float32).pin_memory() if rank == 0 else None
    stream = torch.cuda.current_stream()

    def render_step(stats=None):
        rc = hb.rt.rt_render_device(handle, C.byref(cam), C.byref(p), packed.data_ptr(), None, stream.cuda_stream, stats)
        if rc != 0:
            raise RuntimeError(hb.rt.rt_last_error().decode())
        if world > 1:
            dist.gather(packed, list(gathered.view(world, -1).unbind(0)) if rank == 0 else None, dst=0)
        if rank == 0:
            src = gathered if world > 1 else packed
            rc = hb.rt.rt_untile_device(C.byref(p), src.data_ptr(), offsets.ctypes.data, image.data_ptr(), local, stream.cuda_stream)
            if rc != 0:
                raise RuntimeError(hb.rt.rt_last_error().decode())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # untimed counting passes (deterministic => the same rays as the timed steps):
    #  A. the timed variant with its counters on, full size: exact ray / sample counts of this rank's shard
    #  B. the reference-order traversal (variant 1) at <= 4 spp: the ALGORITHMIC work per ray of SURVEY 8(d)
    #     (spheres, squares, KD nodes and triangles the reference's traversal visits), scaled to the full spp
    keys = ["n_samples", "n_closest_rays", "n_shadow_rays", "n_sphere_tests", "n_square_tests", "n_mesh_tests", "n_node_visits",
            "n_tri_tests", "n_tri_full", "n_tex_fetches", "n_random"]

    def count(variant, c_spp):
        pc = hb.render_params(w, h, c_spp, seed=0, rank=rank, n_ranks=world, tile=(32, 32), collect_stats=True, variant=variant)
        st = hb.RtStats()
        rc = hb.rt.rt_render_device(handle, C.byref(cam), C.byref(pc), packed.data_ptr(), None, stream.cuda_stream, C.byref(st))
        if rc != 0:
            raise RuntimeError(hb.rt.rt_last_error().decode())
        return st.as_dict()

    executed = count(args.variant, spp)
    c_spp = min(spp, 4)
    alg = count(1, c_spp)
    scale = (executed["n_closest_rays"] + executed["n_shadow_rays"]) / max(1, alg["n_closest_rays"] + alg["n_shadow_rays"])
    my = {k: (executed[k] if k in ("n_samples", "n_closest_rays", "n_shadow_rays", "n_random") else int(round(alg[k] * scale))) for k in keys}
    tot = torch.tensor([float(my[k]) for k in keys] + [float(executed[k]) for k in keys], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    total = {k: int(v) for k, v in zip(keys, tot.tolist()[:len(keys)])}
    total_executed = {k: int(v) for k, v in zip(keys, tot.tolist()[len(keys):])}
    rays, flops, byts = flops_and_bytes(total)
    my_rays, my_flops, my_byts = flops_and_bytes(my)
    ex_rays, ex_flops, ex_byts = flops_and_bytes(executed)

    for _ in range(args.warmup):
        flush.zero_()
        render_step()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kst = hb.RtStats()
    kernel_ms = 0.0
    launches = 0
    ev0.record(stream)
    for _ in range(args.steps):
        flush.zero_()
        render_step(C.byref(kst))      # stats != NULL makes the call wait on its own end event: kernel_ms is that kernel time
        kernel_ms += kst.kernel_ms
        launches += kst.n_launches + (1 if rank == 0 else 0)
        if os.environ.get("BENCH_DEBUG"):
            sys.stderr.write("rank %d step kernel_ms %.2f\n" % (rank, kst.kernel_ms))
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = ms.item() / args.steps
    value = rays / (ms_per_step * 1e-3) / 1e6

    # end to end: upload scene (H2D) + render + gather + untile + D2H of the framebuffer into pinned host memory
    e2e_steps = max(3, args.steps)
    scene_bytes = scene.device_bytes(local)
    barrier()
    e2e_each = []
    for _ in range(e2e_steps):
        t0 = time.perf_counter()
        scene.invalidate_device()
        handle = scene.device_handle(local)            # flatten() is cached; this is rt_scene_create: H2D + precompute kernels
        t1 = time.perf_counter()
        render_step()
        t2 = time.perf_counter()
        if rank == 0:
            host_img.copy_(image, non_blocking=True)
        torch.cuda.synchronize()
        e2e_each.append((time.perf_counter() - t0) * 1e3)
        if os.environ.get("BENCH_DEBUG"):
            sys.stderr.write("rank %d e2e step: upload %.2f ms, submit %.2f ms, wait %.2f ms\n" % (rank, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (time.perf_counter() - t2) * 1e3))
    barrier()
    if os.environ.get("BENCH_DEBUG"):
        sys.stderr.write("rank %d e2e ms per step: %s\n" % (rank, " ".join("%.1f" % t for t in e2e_each)))
    # mean over the steps; every step is a complete upload + render + gather + untile + D2H
    e2e_ms = torch.tensor([sum(e2e_each) / len(e2e_each)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = rays / (e2e_ms.item() * 1e-3) / 1e6

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    fp32_unfused, fp32_fused = hb.measure_fp32_peak(local)
    # dominant kernel = k_render_paths; per-launch figures of rank 0's shard
    n_launch = max(1, kst.n_chunks)
    k_ms = kernel_ms / args.steps / n_launch
    ach_tflops = my_flops / n_launch / (k_ms * 1e-3) / 1e12
    ach_gbs = my_byts / n_launch / (k_ms * 1e-3) / 1e9
    t_fp = my_flops / (fp32_unfused * 1e12)
    t_mem = my_byts / (hbm_peak * 1e9)
    bound = "fp32" if t_fp >= t_mem else "hbm"
    roof = {
        "bound": bound,
        "achieved": ach_tflops if bound == "fp32" else ach_gbs,
        "peak": fp32_unfused if bound == "fp32" else hbm_peak,
        "unit": "TFLOP/s" if bound == "fp32" else "GB/s",
        "frac": (ach_tflops / fp32_unfused) if bound == "fp32" else (ach_gbs / hbm_peak),
        "traffic": None,
        "kernel": kernel_of(scene.counts(), args.variant),
        "kernel_ms_per_launch": k_ms, "launches_per_step": n_launch,
        "launch": "one chunk of <= 32 Mi (wavefront) / 16 Mi paths: every render kernel of the chunk, CUDA events around them on the launching stream",
        "peak_source": "fp32 unfused FMUL+FADD measured live by rt_measure_fp32_peak (fused: %.1f TFLOP/s); hbm %s" % (fp32_fused, hbm_src),
        "fp32": {"achieved_tflops": ach_tflops, "peak_tflops": fp32_unfused, "frac": ach_tflops / fp32_unfused,
                 "algorithmic_flops_per_ray": my_flops / max(1, my_rays)},
        "hbm": {"achieved_gbs": ach_gbs, "peak_gbs": hbm_peak, "frac": ach_gbs / hbm_peak,
                "algorithmic_bytes_per_ray": my_byts / max(1, my_rays)},
        "executed": {"note": "work the timed variant actually performed (variant 3 culls tests exactly; node/triangle counts are then "
                             "those of its own hierarchy), same per-test flop/byte figures",
                     "flops_per_ray": ex_flops / max(1, ex_rays), "bytes_per_ray": ex_byts / max(1, ex_rays),
                     "tflops": ex_flops / n_launch / (k_ms * 1e-3) / 1e12, "frac_fp32": ex_flops / n_launch / (k_ms * 1e-3) / 1e12 / fp32_unfused},
        "roofline_mrays_per_s": my_rays / max(t_fp, t_mem) / 1e6,
        "frac_of_roofline_rays": (my_rays / (k_ms * 1e-3 * n_launch)) / (my_rays / max(t_fp, t_mem)),
    }
    # DRAM traffic of the last committed ncu capture (bytes per path), scaled to the paths of one launch group (chunk)
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "latest.json")))
        per_path = prof.get(args.workload, {}).get("dram_bytes_per_path")
        if per_path is not None:
            roof["traffic"] = per_path * (n_px * spp) / n_launch
            roof["traffic_source"] = "profiles/latest.json: %.0f B of DRAM traffic per path (ncu --set full) x paths per chunk" % per_path
    except (OSError, ValueError):
        pass

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            import oracle_ref
            if oracle_ref.available():
                cores = os.cpu_count() or 1
                ref = oracle_ref.Ref().scene(wl["scene"], aspect=w / h, seed=0)
                cw, ch = min(w, 480), min(h, 270)
                crop = ((w - cw) // 2, (h - ch) // 2, (w - cw) // 2 + cw, (h - ch) // 2 + ch)
                t0 = time.perf_counter(); ref.render(w, h, 1, seed=0, threads=0, crop=crop, want_ids=False); dt1 = time.perf_counter() - t0
                s_spp = int(max(1, min(spp, round(15.0 / max(dt1, 1e-3)))))
                t0 = time.perf_counter(); ref.render(w, h, s_spp, seed=0, threads=0, crop=crop, want_ids=False); dt = time.perf_counter() - t0
                cst = scene.render(w, h, s_spp, seed=0, crop=crop, stats=True, want_linear=False)["stats"]
                crays = cst["n_closest_rays"] + cst["n_shadow_rays"]
                cpu = {"value": crays / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "reference",
                       "sample": "centre crop %dx%d of the %dx%d image plane at %d of %d spp: %d samples, %d rays, %.1f s wall; reference sources "
                                 "built headless (oracle/_ref), deterministic RNG shim, %d-thread row pool" % (cw, ch, w, h, s_spp, spp, cw * ch * s_spp, crays, dt, cores),
                       "msamples_per_s": cw * ch * s_spp / dt / 1e6}
        except Exception as e:  # the baseline must never take the GPU number down with it
            cpu = {"value": None, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %r" % (e,)}

    out = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": wl["label"], "scene": wl["scene"], "width": w, "height": h, "spp": spp, "max_bounces": 6, "nb_ech": 10,
                   "seed": 0, "tiles": "32x32 round-robin over ranks" if world > 1 else "32x32",
                   "l2": "256 MiB buffer rewritten before every step (L2 flush); the scene itself is a few MB and is re-read from L2 by design",
                   "rays_per_step": rays, "samples_per_step": total["n_samples"], "rays_per_sample": rays / total["n_samples"],
                   "variant": args.variant},
        "msamples_per_s": total["n_samples"] / (ms_per_step * 1e-3) / 1e6,
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "ms_per_step": e2e_ms.item(), "h2d_bytes_per_step": int(scene_bytes),
                "d2h_bytes_per_step": int(h * w * 3 * 4), "steps": e2e_steps,
                "what": "rt_scene_create (upload + precompute) + rt_render_device + gather + untile + D2H to pinned host, per step"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
        "work": total,
        "work_executed": total_executed,
    }
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
