"""cProfile of VectorPreferenceCountRateMOSTED.reset (config 4: 256 environments at 96 x 96, per-env fluorophores)."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_sted_b200 import envs  # noqa: E402

B = int(os.environ.get("PROBE_B", 256))
vec = envs.make_vec("PreferenceCountRateMOSTED-hard-v0", B, seed=0)
vec.reset(seed=0)
torch.cuda.synchronize()
ts = []
for i in range(3):
    t0 = time.perf_counter(); vec.reset(seed=i + 1); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("reset ms:", [round(t * 1e3, 1) for t in ts])
pr = cProfile.Profile()
pr.enable()
vec.reset(seed=9)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
