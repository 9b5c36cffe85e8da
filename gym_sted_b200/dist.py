"""Environment sharding over the GPUs of one box.

Environments are independent, so the only multi-GPU structure is the partition: env ``i`` of ``N`` lives on
rank ``i // ceil(N / world)``; every rank runs its shard with ``env_base`` = first global id, which is what
the Philox counters carry — results do not depend on the world size.  No collective touches the data path;
the one optional collective gathers the per-env rewards (and small observation vectors) for a central
learner.  One process per GPU (``torchrun``), NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_envs: int, rank: int, world: int):
    """Contiguous block [lo, hi) of global env ids owned by ``rank``."""
    per = -(-n_envs // world)
    lo = min(rank * per, n_envs)
    return lo, min(lo + per, n_envs)


def bind_to_gpu_numa(gpu_index: int) -> bool:
    """Pin the calling process to the CPUs NVML reports as local to GPU ``gpu_index`` (its NUMA node), so pinned host
    buffers allocated afterwards and the H2D/D2H copies of the host-buffer API stay on the socket the GPU hangs off.
    Matters when 8 ranks stream host buffers at once; returns False (and changes nothing) when NVML is unavailable."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(gpu_index))
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False


def init_from_env(backend=None, device=None):
    """``torchrun`` rendezvous (RANK / WORLD_SIZE / MASTER_* from the environment)."""
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return dist.get_rank(), dist.get_world_size()


def gather_rewards(local: torch.Tensor, n_envs: int) -> torch.Tensor:
    """All ranks receive the (n_envs, ...) tensor of per-env values in global env order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    per = -(-n_envs // world)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat(out, dim=0)[:n_envs]
