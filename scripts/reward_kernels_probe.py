"""Reward-side kernels alone at a given image size (PROBE_HW, default 256; PROBE_B images, default 64): STED images from
the simulator, then detection candidates, decorrelation resolution, objectives - device time of each by CUDA events."""
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_sted_b200 import envs, rewards  # noqa: E402

B = int(os.environ.get("PROBE_B", 64)); SIZE = int(os.environ.get("PROBE_HW", 256)); REPS = int(os.environ.get("PROBE_REPS", 5))
vec = envs.make_vec("ContextualMOSTED-easy-hslb-v0", B, compute_f1=True, seed=0, image_size=SIZE)
vec.reset(seed=0)
acts = np.random.default_rng(0).uniform(vec.action_space.low, vec.action_space.high, (B, 3)).astype(np.float32)
info = vec.step(acts)[-1]
sted, conf1, conf2 = info["sted_image"], info["conf1"], info["conf2"]


def timed(name, fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(REPS):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{name:40s} median {np.median(ts):8.3f} ms  min {min(ts):8.3f}  ({B} images {SIZE}x{SIZE})")


timed("nanodomain candidates (device)", lambda: rewards._candidates_device(sted))
timed("resolution chain (device)", lambda: rewards.resolution_device(sted))
timed("foreground + objectives (device)", lambda: rewards.foreground_objectives_device(conf1, sted, conf2))

# Gaussian-fit validation of the candidates of these images: device time, number of fits, and a digest of the fitted
# parameters (two builds of the solver that differ only in inlining must print the same digest)
import hashlib  # noqa: E402

sted32, peaks, npk, _ = rewards._candidates_device(sted)
timed("Gaussian fits (device)", lambda: rewards.gaussfit_validate(sted32, peaks, npk))
valid, params = rewards.gaussfit_validate(sted32, peaks, npk, want_params=True)
torch.cuda.synchronize()
n = npk.cpu().numpy()
v, p = valid.cpu().numpy(), params.cpu().numpy()
h = hashlib.sha256()
for b in range(B):
    h.update(v[b, :n[b]].tobytes()); h.update(p[b, :n[b]].tobytes())
print(f"fits {int(n.sum())} (max {int(n.max())} per image), kept {int(sum(v[b, :n[b]].sum() for b in range(B)))}, digest {h.hexdigest()[:16]}")
