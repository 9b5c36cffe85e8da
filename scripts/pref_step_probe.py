"""Step time of the vectorised PreferenceCountRateMOSTED environment (config 4: 96 x 96, PrefNet reward), B = 256."""
import os
import sys
import time
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_sted_b200 import envs  # noqa: E402

B = int(os.environ.get("PROBE_B", 256))
vec = envs.make_vec("PreferenceCountRateMOSTED-hard-v0", B, seed=0)
t0 = time.perf_counter(); vec.reset(seed=0); torch.cuda.synchronize()
print(f"reset (fluorophore fits for {B} environments + structures + count-rate loop): {(time.perf_counter() - t0) * 1e3:.1f} ms")
acts = np.random.default_rng(0).uniform(vec.action_space.low, vec.action_space.high, (B, 3)).astype(np.float32)
for _ in range(3):
    vec.step(acts)
ts = []
for _ in range(10):
    t0 = time.perf_counter(); out = vec.step(acts); ts.append(time.perf_counter() - t0)
print(f"step: median {np.median(ts) * 1e3:.2f} ms  min {min(ts) * 1e3:.2f}  mean reward {np.nanmean(out[1]):.3f}")
vec.step_timing = True
tt = [vec.step(acts)[4]["timing"] for _ in range(10)]
print("host ms per phase (median): " + ", ".join(f"{k} {np.median([t[k] for t in tt]) * 1e3:.2f}" for k in tt[0]))
