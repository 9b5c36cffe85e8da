"""Host-side phases of VectorContextualMOSTED.step (queueing / waiting for the device / unpacking), B = 256."""
import os
import sys
import numpy as np
import torch

os.environ["STED_B200_STEP_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_sted_b200 import envs  # noqa: E402

B = int(os.environ.get("PROBE_B", 256)); SIZE = int(os.environ.get("PROBE_HW", 64))
vec = envs.make_vec("ContextualMOSTED-easy-hslb-v0", B, compute_f1=True, seed=0, image_size=SIZE)
vec.reset(seed=0)
rng = np.random.default_rng(0)
acts = rng.uniform(vec.action_space.low, vec.action_space.high, (B, 3)).astype(np.float32)
for _ in range(3):
    vec.step(acts)
import time
rows, walls = [], []
for _ in range(6):
    t0 = time.perf_counter()
    info = vec.step(acts)[-1]
    walls.append(time.perf_counter() - t0)
    rows.append(info["timing"])
print("median ms/step: " + "  ".join(f"{k} {np.median([r[k] for r in rows]) * 1e3:.2f}" for k in rows[0]) + f"  | wall {np.median(walls) * 1e3:.2f}")
