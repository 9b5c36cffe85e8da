"""Timing probe of one STED acquisition with in-scan stochastic bleaching (the north_star kernel) at the BASELINE
workload, device resident: CUDA-event times of the whole scan for a few schedule knobs, and a short run for ncu.

    python scripts/scan_probe.py            # event timings (v2 default row range, forced row ranges, round-1 schedule)
    PROBE_NCU=1 ncu ... python scripts/scan_probe.py     # 1 warm-up + 2 scans, nothing else
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_sted_b200 import defaults, envs, synapse  # noqa: E402
from gym_sted_b200.batch import BatchedMicroscope  # noqa: E402

B = int(os.environ.get("PROBE_B", 256))
H = W = int(os.environ.get("PROBE_HW", 256))
m = envs.make_microscope(defaults.load_routine("high-signal_low-bleach"))
maps, _ = synapse.generate_batch(B, H, W, seed=1234)
rng = np.random.default_rng(0)
lo = np.array([defaults.action_spaces[n]["low"] for n in ("pdt", "p_ex", "p_sted")])
hi = np.array([defaults.action_spaces[n]["high"] for n in ("pdt", "p_ex", "p_sted")])
act = rng.uniform(lo, hi, (B, 3))
pdt, p_ex, p_sted = (torch.from_numpy(np.ascontiguousarray(act[:, i])).cuda() for i in range(3))
bm = BatchedMicroscope(m, B, H, W, device="cuda:0", seed=1)
roi = torch.from_numpy(maps).cuda()
flops = 3.0 * B * H * W * bm.K ** 2


def scan(n, label):
    times = []
    for i in range(n):
        bm.set_datamaps(roi, max_molecules=int(maps.sum()))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        bm.get_signal_and_bleach(pdt, p_ex, p_sted, bleach=True)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    t = float(np.median(times[1:])) if n > 1 else times[0]
    import ctypes
    ph = (ctypes.c_uint64 * 8)()
    bm.lib.sted_debug_phase_cycles(ph, 1)
    if sum(ph) and os.environ.get("PROBE_S3"):           # SWEEP_PROBE == 5 build: sweep3_kernel, cycles per scan row and CTA
        g = max(1, ph[7])
        names = ("g.wait", "gather", "g.bar1", "g.decrement", "g.bar2+tma", "c.wait", "c.corr+write")
        print("    per row: " + "  ".join(f"{nm} {ph[i] / g:.0f}" for i, nm in enumerate(names)), flush=True)
    elif sum(ph):                                        # SWEEP_PROBE == 3 build: cycles per group of rows and CTA
        g = max(1, ph[6])
        names = ("gather", "g.wait1", "g.phaseB", "aux.stage+ring", "aux.corr", "aux.wait1")
        if os.environ.get("PROBE_SPLIT_B"):              # SWEEP_PROBE == 4 build: thread 0's view of the phase between the barriers
            names = ("gather", "wait1", "B.wait2", "B.ring", "B.corr", "-")
        print("    per group: " + "  ".join(f"{nm} {ph[i] / g:.0f}" for i, nm in enumerate(names)) + f"  loop {ph[7] / g:.0f}", flush=True)
    print(f"{label:40s} {t:7.3f} ms   {flops / t / 1e9:6.1f} TFLOP/s (3N)   deaths/env {int(maps.sum() - bm.datamaps_roi().sum().item()) // B}", flush=True)
    return t


if os.environ.get("PROBE_NCU"):
    scan(3, "ncu run")
    sys.exit(0)
scan(6, "v2 default")
for rr in os.environ.get("PROBE_RR", "32,64,128,256").split(","):
    os.environ["STED_B200_ROW_RANGE"] = rr
    scan(5, f"v2 row range {rr}")
os.environ.pop("STED_B200_ROW_RANGE")
os.environ["STED_B200_SCAN_V1"] = "1"
scan(5, "round-1 schedule")
