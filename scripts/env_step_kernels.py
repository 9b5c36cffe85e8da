"""Kernels launched by ONE vector env step (PROBE_ENV, default ContextualMOSTED-easy-hslb-v0) (torch.profiler / CUPTI): name, launches, device time."""
import os
import sys
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_sted_b200 import envs  # noqa: E402

B = int(os.environ.get("PROBE_B", 256)); SIZE = int(os.environ.get("PROBE_HW", 64))
ENV = os.environ.get("PROBE_ENV", "ContextualMOSTED-easy-hslb-v0")
vec = (envs.make_vec(ENV, B, seed=0) if ENV.startswith("Preference")
       else envs.make_vec(ENV, B, compute_f1=True, seed=0, image_size=SIZE))
vec.reset(seed=0)
acts = np.random.default_rng(0).uniform(vec.action_space.low, vec.action_space.high, (B, 3)).astype(np.float32)
for _ in range(3):
    vec.step(acts)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    vec.step(acts)
    torch.cuda.synchronize()
rows = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        n, t = rows.get(ev.name, (0, 0.0))
        rows[ev.name] = (n + 1, t + ev.device_time)
print(f"{'kernel':110s} launches   us")
for name, (n, t) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:110]:110s} {n:5d} {t:8.1f}")

if os.environ.get("PROBE_TIMELINE"):
    # start / end of every kernel and copy of the step relative to the first one, by stream (which ones overlap)
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    t0 = min(e.time_range.start for e in evs)
    print(f"\n{'start us':>9s} {'end us':>9s} {'stream':>7s}  kernel")
    for e in sorted(evs, key=lambda e: e.time_range.start):
        print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - t0:9.1f} {getattr(e, 'device_index', 0)!s:>3s}/{getattr(e, 'device_resource_id', '?')!s:>3s}  {e.name[:70]}")
