"""World-size-2 gloo test of the env sharding and the reward gather (CPU; the N > 1 path of bench.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from gym_sted_b200 import dist as sd


def test_shard_ranges_partition_the_envs():
    for n, world in [(256, 1), (256, 2), (2048, 8), (10, 4), (3, 8), (7, 2)]:
        ids = []
        for r in range(world):
            lo, hi = sd.shard_range(n, r, world)
            assert 0 <= lo <= hi <= n
            ids += list(range(lo, hi))
        assert ids == list(range(n))


def _worker(rank, world, port, n_envs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sd.init_from_env("gloo")
    lo, hi = sd.shard_range(n_envs, rank, world)
    # per-env "reward" = function of the GLOBAL env id, as the Philox keying guarantees on the GPU path
    local = torch.arange(lo, hi, dtype=torch.float32) * 0.5 + 1.0
    obs = torch.stack([torch.arange(lo, hi, dtype=torch.float32)] * 3, dim=1)
    full = sd.gather_rewards(local, n_envs)
    full_obs = sd.gather_rewards(obs, n_envs)
    q.put((rank, full.numpy(), full_obs.numpy()))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("n_envs", [8, 7])
def test_reward_gather_world_size_2(n_envs):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_envs, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.arange(n_envs, dtype=np.float32) * 0.5 + 1.0
    for rank, full, full_obs in got:
        np.testing.assert_array_equal(full, expect)
        np.testing.assert_array_equal(full_obs[:, 1], np.arange(n_envs, dtype=np.float32))


def test_numa_binding_is_a_safe_noop_without_nvml_devices():
    """``bind_to_gpu_numa`` never raises: it returns False (affinity untouched) when NVML or the GPU is absent."""
    import os
    from gym_sted_b200 import dist as sd
    before = os.sched_getaffinity(0)
    ok = sd.bind_to_gpu_numa(0)
    assert isinstance(ok, bool)
    if not ok:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
