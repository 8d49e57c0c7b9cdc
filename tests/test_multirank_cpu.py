"""N > 1 host-side logic on the CPU: tile sharding arithmetic and the gather/untile bookkeeping, with two real
processes over gloo. The pixels themselves are synthetic (a function of their coordinates) — rendering needs the GPU —
but the layout, the offsets and the reassembly are the ones bench.py's multi-GPU path uses."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_tiles_partition_the_rectangle(hb):
    for (w, h, crop, tile, n) in [(200, 120, None, (32, 32), 3), (1920, 1080, None, (32, 32), 8), (97, 41, (5, 3, 90, 40), (16, 8), 2),
                                  (64, 64, None, (64, 64), 4), (33, 33, None, (32, 32), 1)]:
        x0, y0, x1, y1 = crop if crop else (0, 0, w, h)
        seen = np.zeros((h, w), np.int32)
        total = 0
        for r in range(n):
            p = hb.render_params(w, h, 1, crop=crop, rank=r, n_ranks=n, tile=tile)
            t = hb.tile_layout(p)
            assert int(hb.rt.rt_render_pixel_count(p)) == int((t[:, 2] * t[:, 3]).sum())
            total += len(t)
            for tx, ty, tw, th in t:
                assert 0 < tw <= tile[0] and 0 < th <= tile[1]
                seen[ty:ty + th, tx:tx + tw] += 1
        assert total == -(-(x1 - x0) // tile[0]) * -(-(y1 - y0) // tile[1])
        assert (seen[y0:y1, x0:x1] == 1).all() and seen.sum() == (x1 - x0) * (y1 - y0)      # every pixel exactly once
    # round-robin along rows, rows shifted against each other (rt_capi.cu : tile_owner): rank loads differ by a tile or two of 255
    counts = [len(hb.tile_layout(hb.render_params(1920, 1080, 1, rank=r, n_ranks=8, tile=(32, 32)))) for r in range(8)]
    assert max(counts) - min(counts) <= 2
    # ... and no rank owns whole columns of tiles when the rank count divides the tiles per row (3840 / 32 = 120, 8 ranks)
    t = hb.tile_layout(hb.render_params(3840, 2160, 1, rank=0, n_ranks=8, tile=(32, 32)))
    assert len(np.unique(t[:, 0])) == 120


WORKER = r'''
import importlib, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
hb = importlib.import_module("hai719-raytracing_b200")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
w, h = 150, 70
def pix(x, y):
    return np.stack([x + 1000.0 * y, 2.0 * x - y, x * 0 + rank * 0 + 7.0], -1).astype(np.float32)
def packed_for(r):
    t = hb.tile_layout(hb.render_params(w, h, 1, rank=r, n_ranks=world, tile=(32, 32)))
    out = []
    for tx, ty, tw, th in t:
        yy, xx = np.mgrid[ty:ty + th, tx:tx + tw]
        out.append(pix(xx, yy).reshape(-1, 3))
    return np.concatenate(out) if out else np.zeros((0, 3), np.float32), t
counts = [int(hb.rt.rt_render_pixel_count(hb.render_params(w, h, 1, rank=r, n_ranks=world, tile=(32, 32)))) for r in range(world)]
max_px = max(counts)
mine, _ = packed_for(rank)
buf = torch.zeros(max_px * 3)
buf[:mine.size] = torch.from_numpy(mine.ravel())
gathered = [torch.zeros(max_px * 3) for _ in range(world)] if rank == 0 else None
dist.gather(buf, gathered, dst=0)
if rank == 0:
    img = np.full((h, w, 3), -1, np.float32)
    for r in range(world):
        _, tiles = packed_for(r)
        src = gathered[r].numpy().reshape(-1, 3)
        o = 0
        for tx, ty, tw, th in tiles:
            img[ty:ty + th, tx:tx + tw] = src[o:o + tw * th].reshape(th, tw, 3)
            o += tw * th
        assert o == counts[r]
    yy, xx = np.mgrid[0:h, 0:w]
    assert np.array_equal(img, pix(xx, yy))
    print("OK")
dist.destroy_process_group()
'''


def test_two_rank_gather_and_untile_over_gloo(hb, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29517")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:]


WORKER_SHARED = r'''
import importlib, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
hb = importlib.import_module("hai719-raytracing_b200")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
w, h, crop = 150, 70, (7, 3, 140, 66)
x0, y0, x1, y1 = crop
rw, rh = x1 - x0, y1 - y0
# rank 0 owns the framebuffer and publishes a 64-byte handle (bench.py: rt_ipc_alloc + dist.broadcast of the CUDA IPC
# handle; here a file stands in for the device allocation and its name, padded to 64 bytes, for the handle)
handle = torch.zeros(64, dtype=torch.uint8)
if rank == 0:
    path = os.path.join(sys.argv[2], "fb.bin")
    np.full(rh * rw * 3, -1.0, np.float32).tofile(path)
    raw = os.path.basename(path).encode()
    handle[:len(raw)] = torch.tensor(list(raw), dtype=torch.uint8)
dist.broadcast(handle, src=0)
name = bytes(handle.tolist()).rstrip(b"\0").decode()
fb = np.memmap(os.path.join(sys.argv[2], name), np.float32, "r+", shape=(rh * rw * 3,))
# every rank stores its tiles at the place k_resolve's image mode computes: 3 * ((y - ry0) * rect_w + (x - rx0))
tiles = hb.tile_layout(hb.render_params(w, h, 1, crop=crop, rank=rank, n_ranks=world, tile=(32, 32)))
for tx, ty, tw, th in tiles:
    for y in range(ty, ty + th):
        for x in range(tx, tx + tw):
            o = 3 * ((y - y0) * rw + (x - x0))
            assert fb[o] == -1.0                       # nobody else writes this pixel
            fb[o:o + 3] = (x + 1000.0 * y, 2.0 * x - y, float(rank))
fb.flush()
dist.barrier()                                         # bench.py: stream sync + barrier before rank 0 reads
if rank == 0:
    img = np.fromfile(os.path.join(sys.argv[2], name), np.float32).reshape(rh, rw, 3)
    yy, xx = np.mgrid[y0:y1, x0:x1]
    assert np.array_equal(img[..., 0], (xx + 1000.0 * yy).astype(np.float32)) and np.array_equal(img[..., 1], (2.0 * xx - yy).astype(np.float32))
    ntx = -(-rw // 32)
    k = 3 if world % 3 else (5 if world % 5 else 1)
    owner = ((xx - x0) // 32 + k * ((yy - y0) // 32)) % world
    assert np.array_equal(img[..., 2], owner.astype(np.float32))      # tile (tx, ty) belongs to rank (tx + K ty) % world (rt_capi.cu : tile_owner)
    print("OK")
dist.barrier()
dist.destroy_process_group()
'''


def test_two_ranks_store_into_one_shared_framebuffer_over_gloo(hb, tmp_path):
    """The N > 1 path of bench.py without a GPU: handle broadcast, each rank's tiles stored at their row-major place in
    rank 0's framebuffer, barrier, read-back. No packed buffers, no gather, no untile."""
    script = tmp_path / "worker_shared.py"
    script.write_text(WORKER_SHARED)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29519")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29519", str(script), ROOT, str(tmp_path)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:]
