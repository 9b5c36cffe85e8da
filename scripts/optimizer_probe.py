"""Times FluorescenceOptimizer.optimize_batch (sted_fluo_optimize) for 256 criteria sampled like BleachSampler("uniform")."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_sted_b200 import defaults, envs, optimizer  # noqa: E402

m = envs.make_microscope(detector={"noise": False, "background": defaults.DETECTOR["background"]})
opt = optimizer.FluorescenceOptimizer(m)
rng = np.random.default_rng(0)
B = int(os.environ.get("PROBE_B", 256))
crit = [{"bleach": {"p_ex": 10e-6, "p_sted": 150e-3, "pdt": 30e-6, "target": float(rng.uniform(0.2, 0.7))},
         "signal": {"p_ex": 10e-6, "p_sted": 0.0, "pdt": 10e-6, "target": float(rng.uniform(30, 200))}} for _ in range(B)]
opt.optimize_batch(crit[:8])
torch.cuda.synchronize()
t0 = time.perf_counter()
out = opt.optimize_batch(crit)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(json.dumps({"criteria": B, "seconds": dt, "ms_per_environment": dt / B * 1e3,
                  "max_rel_signal_miss": float(np.max(np.abs(out[:, 3] / [c["signal"]["target"] for c in crit] - 1)))}))
