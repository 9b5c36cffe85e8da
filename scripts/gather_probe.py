"""Gather-only timing probe (CUDA events): confocal gather of the bench workload, noise-free (no detector launch)."""
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_sted_b200 import defaults, synapse  # noqa: E402
from gym_sted_b200.batch import BatchedMicroscope  # noqa: E402
from gym_sted_b200.envs import make_microscope  # noqa: E402

B = int(os.environ.get("PROBE_B", 256)); H = W = int(os.environ.get("PROBE_HW", 256))
micro = make_microscope(defaults.load_routine("high-signal_low-bleach"))
bm = BatchedMicroscope(micro, B, H, W, seed=1)
maps, _ = synapse.generate_batch(min(B, 32), H, W, seed=1234)
if os.environ.get("PROBE_DENSE"):          # every pixel occupied: nothing to skip
    maps = np.maximum(maps, 1)
bm.set_datamaps(torch.from_numpy(np.tile(maps, ((B + 31) // 32, 1, 1))[:B]).float().cuda())
kw = dict(pdt=torch.full((B,), 10e-6, dtype=torch.float64).cuda(), p_ex=torch.full((B,), 10e-6, dtype=torch.float64).cuda(),
          p_sted=torch.full((B,), 0.0, dtype=torch.float64).cuda())
for _ in range(3):
    out = bm.get_signal_and_bleach(**kw, bleach=False, noise=False)
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = bm.get_signal_and_bleach(**kw, bleach=False, noise=False); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
import ctypes
ph = (ctypes.c_uint64 * 8)()
bm.lib.sted_debug_phase_cycles(ph, 1)
if ph[3]:
    print("phase cycles per CTA: load %.0f  main %.0f  epilogue %.0f  (%d CTAs)" % (ph[0] / ph[3], ph[1] / ph[3], ph[2] / ph[3], ph[3]))
print(f"{os.environ.get('STED_B200_LIB', 'default'):40s} gather+tables+detect(noise-free) median {np.median(ts):.3f} ms  min {min(ts):.3f}  checksum {float(out[0].double().sum()):.6e}")
