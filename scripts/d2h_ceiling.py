"""Host-transfer ceiling with every rank copying at once (torchrun): device->host and host->device GB/s per rank and in
total, for pinned memory from cudaHostAlloc (torch pin_memory) and for 2 MB-aligned, MADV_HUGEPAGE-backed memory
registered with cudaHostRegister.  Sizes follow one bench step (168 MB down, 33.6 MB up per rank)."""
import ctypes
import mmap
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
if "--numa" in sys.argv:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gym_sted_b200 import dist as sd
    sd.bind_to_gpu_numa(local)
NB = int(os.environ.get("D2H_MB", 168)) << 20
REPS = 12


def huge_pinned(nbytes):
    align = 2 << 20
    mm = mmap.mmap(-1, nbytes + align, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    base = ctypes.addressof(ctypes.c_char.from_buffer(mm))
    off = (-base) % align
    try:
        mm.madvise(mmap.MADV_HUGEPAGE, off, nbytes)
    except Exception as e:          # noqa: BLE001
        print("madvise failed:", e)
    arr = np.frombuffer(mm, dtype=np.uint8, count=nbytes, offset=off)
    arr[:] = 0                      # fault the pages in
    rc = torch.cuda.cudart().cudaHostRegister(arr.ctypes.data, nbytes, 0)
    assert int(rc) == 0, rc
    t = torch.from_numpy(arr)
    return t, mm


def anon_huge_kb():
    try:
        for line in open("/proc/self/smaps_rollup"):
            if line.startswith("AnonHugePages"):
                return int(line.split()[1])
    except OSError:
        pass
    return -1


def run(name, host):
    d = torch.empty(NB, dtype=torch.uint8, device=dev)
    s = torch.cuda.Stream()
    out = {}
    for direction in ("d2h", "h2d"):
        for _ in range(2):
            with torch.cuda.stream(s):
                (host.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(host, non_blocking=True))
        s.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s):
            for _ in range(REPS):
                (host.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(host, non_blocking=True))
        s.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        out[direction] = NB * REPS / float(dt.item()) / 1e9
    if rank == 0:
        print(f"{name:34s} ranks {world}: d2h {out['d2h']:6.1f} GB/s per rank = {out['d2h'] * world:6.1f} total | "
              f"h2d {out['h2d']:6.1f} per rank = {out['h2d'] * world:6.1f} total", flush=True)


if rank == 0:
    print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), "| nproc", os.cpu_count(), flush=True)
run("cudaHostAlloc (torch pin_memory)", torch.empty(NB, dtype=torch.uint8, pin_memory=True))
before = anon_huge_kb()
t, keep = huge_pinned(NB)
if rank == 0:
    print(f"AnonHugePages {before} -> {anon_huge_kb()} kB", flush=True)
run("MADV_HUGEPAGE + cudaHostRegister", t)
if world > 1:
    dist.destroy_process_group()
