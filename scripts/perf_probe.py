"""Kernel-level timing probe (CUDA events): gather / sweep variants on the bench workload."""
import ctypes
import sys
import os
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_sted_b200 import _lib, defaults, synapse  # noqa: E402
from gym_sted_b200.batch import BatchedMicroscope  # noqa: E402
from tests import helpers  # noqa: E402

B = int(os.environ.get("PROBE_B", 256)); H = W = int(os.environ.get("PROBE_HW", 256))
micro = helpers.product_microscope(defaults.load_routine("high-signal_low-bleach"), noise=True,
                                   background=defaults.DETECTOR["background"])
bm = BatchedMicroscope(micro, B, H, W, seed=1)
maps, _ = synapse.generate_batch(min(B, 32), H, W, seed=1234)
if os.environ.get("PROBE_DENSE"):          # every pixel occupied: nothing to skip
    maps = np.maximum(maps, 1)
roi = torch.from_numpy(np.tile(maps, ((B + 31) // 32, 1, 1))[:B]).float().cuda()
g = torch.Generator().manual_seed(0)
lo = torch.tensor([0., 0., 1e-6], dtype=torch.float64); hi = torch.tensor([0.35, 20e-6, 100e-6], dtype=torch.float64)
act = (lo + (hi - lo) * torch.rand((B, 3), generator=g, dtype=torch.float64)).cuda()


def timeit(name, fn, n=5):
    for _ in range(2):
        bm.set_datamaps(roi); fn()
    ts = []
    for _ in range(n):
        bm.set_datamaps(roi)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ph = (ctypes.c_uint64 * 8)()
    bm.lib.sted_debug_phase_cycles(ph, 1)
    extra = ""
    if sum(ph):
        tot = sum(ph)
        extra = "  G/wait/replay/wait/cycle %: " + " ".join(f"{100 * x / tot:.0f}" for x in list(ph)[:8]) + f"  cyc/row/cta {tot / (n + 2) / B / max(1, (W + 127) // 128) / H:.0f}"
    print(f"{name:48s} {np.median(ts):8.3f} ms{extra}", flush=True)


full = dict(pdt=act[:, 2].contiguous(), p_ex=act[:, 1].contiguous(), p_sted=act[:, 0].contiguous())
weak = dict(pdt=torch.full((B,), 10e-6, dtype=torch.float64).cuda(), p_ex=torch.full((B,), 10e-6, dtype=torch.float64).cuda(),
            p_sted=torch.full((B,), 0.0, dtype=torch.float64).cuda())
mid = dict(pdt=torch.full((B,), 30e-6, dtype=torch.float64).cuda(), p_ex=torch.full((B,), 10e-6, dtype=torch.float64).cuda(),
           p_sted=torch.full((B,), 0.15, dtype=torch.float64).cuda())
timeit("confocal gather (+detect)", lambda: bm.get_signal_and_bleach(**weak, bleach=False))
timeit("confocal gather, noise-free", lambda: bm.get_signal_and_bleach(**weak, bleach=False, noise=False))
timeit("sweep expectation, uniform actions", lambda: bm.get_signal_and_bleach(**full, bleach=True, sample_bleach=False, update=False))
timeit("sweep stochastic, p_sted=0 (no deaths)", lambda: bm.get_signal_and_bleach(**weak, bleach=True, update=False))
timeit("sweep stochastic, 150 mW / 30 us", lambda: bm.get_signal_and_bleach(**mid, bleach=True, update=False))
timeit("sweep stochastic, uniform actions", lambda: bm.get_signal_and_bleach(**full, bleach=True, update=False))
pk = ctypes.c_double(); _lib.check(bm.lib.sted_fma_peak(ctypes.byref(pk), None)); print("fma peak TFLOP/s", pk.value)
