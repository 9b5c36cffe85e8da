"""Episodes of the vectorised ContextualMOSTED-hard-v0 (config 3: every reset fits a fluorophore per environment):
reset + 10 steps, wall time per episode, with the fits of the next episode queued ahead or not
(STED_B200_PREFETCH_FLUO)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_sted_b200 import envs  # noqa: E402

B = int(os.environ.get("PROBE_B", 256)); SIZE = int(os.environ.get("PROBE_HW", 64))
vec = envs.make_vec("ContextualMOSTED-hard-v0", B, compute_f1=True, seed=0, image_size=SIZE)
rng = np.random.default_rng(0)
acts = rng.uniform(vec.action_space.low, vec.action_space.high, (B, 3)).astype(np.float32)
vec.reset(seed=0)
for _ in range(3):
    vec.step(acts)
eps, resets, steps, rewards = [], [], [], []
for e in range(6):
    t0 = time.perf_counter()
    vec.reset()
    t1 = time.perf_counter()
    for _ in range(10):
        out = vec.step(acts)
        rewards.append(float(out[1].mean()))
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    eps.append(t2 - t0); resets.append(t1 - t0); steps.append((t2 - t1) / 10)
print(f"prefetch={vec.prefetch_fluorophores}: episode {np.median(eps) * 1e3:.1f} ms  reset {np.median(resets) * 1e3:.1f} ms  "
      f"step {np.median(steps) * 1e3:.2f} ms  mean reward {np.mean(rewards):.6f}  sigma0 {vec.bleach_sampler.criteria[0]['signal']['target']:.4f}")
