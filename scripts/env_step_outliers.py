"""Which kernel makes a slow VectorContextualMOSTED.step: random actions every step (as bench.py --workload env_step
draws them), per-step wall time, and for every kernel the longest single launch over the run (torch.profiler)."""
import os
import sys
import time
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_sted_b200 import envs  # noqa: E402

B = int(os.environ.get("PROBE_B", 256)); SIZE = int(os.environ.get("PROBE_HW", 64)); N = int(os.environ.get("PROBE_STEPS", 30))
# same construction as bench.py run_env_step (rank 0): seeds, env_base, float64 actions
vec = envs.make_vec("ContextualMOSTED-easy-hslb-v0", B, image_size=SIZE, device=torch.device("cuda:0"), seed=1234, env_base=0,
                    max_episode_steps=N + 5)
vec.reset(seed=1234)
rng = np.random.default_rng(0)
acts = rng.uniform(vec.action_space.low, vec.action_space.high, (N + 3, B, 3))
for i in range(3):
    vec.step(acts[i])
walls = []
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(N):
        t0 = time.perf_counter()
        vec.step(acts[3 + i])
        walls.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
print("step ms:", " ".join(f"{w:.1f}" for w in walls))
rows = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        n, t, mx = rows.get(ev.name, (0, 0.0, 0.0))
        rows[ev.name] = (n + 1, t + ev.device_time, max(mx, ev.device_time))
print(f"{'kernel':90s} launches  mean us   max us")
for name, (n, t, mx) in sorted(rows.items(), key=lambda kv: -kv[1][2])[:12]:
    print(f"{name[:90]:90s} {n:5d} {t / n:9.1f} {mx:9.1f}")
