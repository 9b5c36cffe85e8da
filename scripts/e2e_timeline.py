"""Device timeline (CUPTI through torch.profiler) of two host-buffer steps of the bench's session schedule."""
import json
import os
import sys
import tempfile
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_sted_b200 import defaults, synapse  # noqa: E402
from gym_sted_b200.envs import make_microscope  # noqa: E402
from gym_sted_b200.session import StedSession, pinned_empty  # noqa: E402

B, H, W = 256, 256, 256
ORDER = int(os.environ.get("ORDER", 1))
micro = make_microscope(defaults.load_routine("high-signal_low-bleach"))
maps, _ = synapse.generate_batch(32, H, W, seed=1234)
maps = np.tile(maps, (8, 1, 1)).astype(np.uint16)
mol = int(maps.astype(np.int64).sum())
rng = np.random.default_rng(0)
lo = np.array([0., 0., 1e-6]); hi = np.array([0.35, 20e-6, 100e-6])
acts = lo + (hi - lo) * rng.random((16, B, 3))
conf = (np.full(B, defaults.PDT), np.full(B, defaults.P_EX), np.zeros(B))
main = StedSession(micro, B, H, W, seed=1)
nxt = StedSession(micro, B, H, W, seed=2, low_priority=True)
h_maps = [pinned_empty((B, H, W), np.uint16) for _ in range(2)]
for h in h_maps:
    h[...] = maps
h_out = [pinned_empty((B, H, W), np.uint16) for _ in range(4)]
h_bl = pinned_empty((B, H, W), np.uint16)


def step(i):
    a = acts[i % 16]
    nxt.upload(h_maps[i % 2], max_molecules=mol)
    nxt.acquire(*conf, bleach=False, out=h_out[3])
    main.acquire(*conf, bleach=False, out=h_out[0])
    main.acquire(np.ascontiguousarray(a[:, 2]), np.ascontiguousarray(a[:, 1]), np.ascontiguousarray(a[:, 0]), bleach=True, out=h_out[1])
    main.download(h_bl)
    if ORDER:
        main.after(nxt)
    main.acquire(*conf, bleach=False, out=h_out[2])
    main.adopt(nxt)
    main.sync(); nxt.sync()


main.upload(h_maps[0], max_molecules=mol)
for i in range(3):
    step(i)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(3); step(4)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "t.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
print(f"{'start ms':>9s} {'dur ms':>8s} stream  name")
for e in ev:
    nm = e["name"].replace("(anonymous namespace)::", "")[:70]
    print(f"{(e['ts'] - t0) / 1e3:9.3f} {e['dur'] / 1e3:8.3f} {e['args'].get('stream', '?'):>6}  {nm}")
