import os, sys, time, collections
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_sted_b200 import envs, rewards, synapse
from gym_sted_b200.batch import BatchedMicroscope
T = collections.defaultdict(float)
def wrap(obj, name, label=None):
    f = getattr(obj, name)
    def g(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = f(*a, **k)
        torch.cuda.synchronize(); T[label or name] += time.perf_counter() - t0
        return r
    setattr(obj, name, g)
wrap(BatchedMicroscope, "get_signal_and_bleach")
wrap(BatchedMicroscope, "set_datamaps")
wrap(rewards, "foreground_objectives")
wrap(rewards, "resolution_device")
wrap(rewards, "foreground_objectives_device")
wrap(rewards, "nanodomain_f1_launch")
wrap(rewards, "nanodomain_f1_batch")
wrap(synapse.DeviceSynapseGenerator, "generate")
wrap(rewards, "_candidates_device")
wrap(rewards, "gaussfit_validate")
wrap(rewards, "match_counts")
wrap(rewards, "detect_nanodomains_batch")
B = 256
SIZE = int(os.environ.get("PROBE_HW", 64))
vec = envs.make_vec("ContextualMOSTED-easy-hslb-v0", B, compute_f1=True, seed=0, image_size=SIZE)
vec.reset(seed=0)
rng = np.random.default_rng(0)
acts = rng.uniform(vec.action_space.low, vec.action_space.high, (B, 3)).astype(np.float32)
acts[:, 0] = np.minimum(acts[:, 0], 0.2)
for _ in range(2): vec.step(acts)
T.clear()
torch.cuda.synchronize(); t0 = time.perf_counter()
n = 5
for _ in range(n): vec.step(acts)
torch.cuda.synchronize(); tot = time.perf_counter() - t0
for k, v in sorted(T.items(), key=lambda kv: -kv[1]): print(f"{k:28s} {v / n * 1e3:7.2f} ms/step")
print(f"total {tot / n * 1e3:.2f} ms/step; unaccounted {(tot - sum(T.values())) / n * 1e3:.2f}")
