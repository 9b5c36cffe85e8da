"""Wall-clock cost of full vectorised env.step() calls (acquisitions + rewards + observation packing)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_sted_b200 import envs  # noqa: E402

B = int(os.environ.get("PROBE_B", 256))
for env_id, f1 in (("ContextualMOSTED-easy-hslb-v0", False), ("ContextualMOSTED-easy-hslb-v0", True)):
    vec = envs.make_vec(env_id, B, compute_f1=f1, seed=0)
    vec.reset(seed=0)
    rng = np.random.default_rng(0)
    acts = rng.uniform(vec.action_space.low, vec.action_space.high, (B, 3)).astype(np.float32)
    acts[:, 0] = np.minimum(acts[:, 0], 0.2)
    for _ in range(2):
        vec.step(acts)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        vec.step(acts)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"{env_id} B={B} f1={f1}: {dt * 1e3:.1f} ms/step = {B / dt:.0f} env-steps/s = {4 * B / dt:.0f} acquisitions/s", flush=True)
    vec.reset()
