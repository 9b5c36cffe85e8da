"""cProfile of VectorContextualMOSTED.step (host side), B = 256 at 64 x 64."""
import cProfile
import os
import pstats
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_sted_b200 import envs  # noqa: E402

B = int(os.environ.get("PROBE_B", 256))
vec = envs.make_vec("ContextualMOSTED-easy-hslb-v0", B, compute_f1=True, seed=0, max_episode_steps=64)
vec.reset(seed=0)
acts = np.random.default_rng(0).uniform(vec.action_space.low, vec.action_space.high, (B, 3)).astype(np.float32)
for _ in range(3):
    vec.step(acts)
pr = cProfile.Profile()
pr.enable()
for _ in range(30):
    vec.step(acts)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
