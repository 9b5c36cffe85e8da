"""Where the host-buffer (e2e) step spends its time: the bench's session schedule with parts switched off."""
import os
import sys
import time
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_sted_b200 import defaults, synapse  # noqa: E402
from gym_sted_b200.envs import make_microscope  # noqa: E402
from gym_sted_b200.session import StedSession, pinned_empty  # noqa: E402

B, H, W = 256, 256, 256
micro = make_microscope(defaults.load_routine("high-signal_low-bleach"))
maps, _ = synapse.generate_batch(32, H, W, seed=1234)
maps = np.tile(maps, (8, 1, 1)).astype(np.uint16)
mol = int(maps.astype(np.int64).sum())
rng = np.random.default_rng(0)
lo = np.array([0., 0., 1e-6]); hi = np.array([0.35, 20e-6, 100e-6])
acts = lo + (hi - lo) * rng.random((16, B, 3))
conf = (np.full(B, defaults.PDT), np.full(B, defaults.P_EX), np.zeros(B))


def run(name, upload=True, download=True, images=True, sync_every=True, low=True, steps=8, order=0, imgs=(1, 1, 1, 1)):
    main = StedSession(micro, B, H, W, seed=1)
    nxt = StedSession(micro, B, H, W, seed=2, low_priority=low)
    h_maps = [pinned_empty((B, H, W), np.uint16) for _ in range(2)]
    for h in h_maps:
        h[...] = maps
    h_out = [pinned_empty((B, H, W), np.uint16) for _ in range(4)]
    h_bl = pinned_empty((B, H, W), np.uint16)
    dmap = torch.from_numpy(maps.astype(np.float32)).cuda()
    o = (lambda k: h_out[k] if imgs[k] else None) if images else (lambda k: None)

    def step(i):
        a = acts[i % 16]
        if upload:
            nxt.upload(h_maps[i % 2], max_molecules=mol)
        else:
            nxt.upload_device(dmap, max_molecules=mol)
        nxt.acquire(*conf, bleach=False, out=o(3))
        main.acquire(*conf, bleach=False, out=o(0))
        if order == 2:
            main.after(nxt)
        main.acquire(np.ascontiguousarray(a[:, 2]), np.ascontiguousarray(a[:, 1]), np.ascontiguousarray(a[:, 0]), bleach=True, out=o(1))
        if download:
            main.download(h_bl)
        if order == 1:
            main.after(nxt)
        main.acquire(*conf, bleach=False, out=o(2))
        main.adopt(nxt)
        if sync_every:
            main.sync(); nxt.sync()

    main.upload(h_maps[0], max_molecules=mol)
    step(0); step(1)
    main.sync(); nxt.sync(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        step(i + 2)
    main.sync(); nxt.sync(); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    print(f"{name:58s} {dt * 1e3:7.3f} ms/step  {4 * B / dt / 1e3:7.1f} k acq/s", flush=True)
    main.close(); nxt.close()


run("full e2e step (bench schedule)")
run("  conf2 ordered after next's gather", order=1)
for k, nm in enumerate(("conf1", "sted", "conf2", "next")):
    m = [0, 0, 0, 0]; m[k] = 1
    run(f"  order=1, no map download, only the {nm} image", order=1, download=False, imgs=tuple(m))
run("  order=1, map download only", order=1, imgs=(0, 0, 0, 0))
run("  order=1, all images, no map", order=1, download=False)
run("  order=1, no upload", order=1, upload=False)
