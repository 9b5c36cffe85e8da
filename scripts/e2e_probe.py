import os
import sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_sted_b200 import _lib, defaults, synapse
from gym_sted_b200.base import fluo_struct
from tests import helpers
B, H, W = 256, 256, 256
lib = _lib.load()
micro = helpers.product_microscope(noise=True, background=defaults.DETECTOR["background"])
cache = micro._packed_cache(helpers.PX)
K = cache.shape[-1]
fl = (_lib.StedFluoParams * B)(*[fluo_struct(micro.fluo, micro.excitation, micro.sted)] * B)
det = micro.detector.params()
maps, _ = synapse.generate_batch(B, H, W, seed=0)
h_map0 = torch.from_numpy(maps.astype(np.float32)).pin_memory()
h_sig = torch.empty((B, H, W), dtype=torch.float32).pin_memory()
vp = ctypes.c_void_p
conf = [np.full(B, v) for v in (defaults.P_EX, 0.0, defaults.PDT)]
sted = [np.full(B, v) for v in (8e-6, 0.12, 15e-6)]
def call(h_map, p, flags, off):
    _lib.check(lib.sted_acquire_host(h_map.data_ptr(), cache.ctypes.data_as(vp), ctypes.cast(fl, vp), p[0].ctypes.data_as(vp),
               p[1].ctypes.data_as(vp), p[2].ctypes.data_as(vp), B, H, W, K, flags, ctypes.byref(det), 1234, off, 0, None, h_sig.data_ptr()))
# raw PCIe
d = torch.empty((B, H, W), dtype=torch.float32, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(h_map0, non_blocking=True)), ("D2H", lambda: h_sig.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"{name}: {dt*1e3:.2f} ms for 67 MB = {d.numel()*4/dt/1e9:.1f} GB/s")
hs = [h_map0.clone().pin_memory() for _ in range(8)]
for name, p, flags in (("conf", conf, _lib.STED_SAMPLE_PHOTONS), ("sted", sted, _lib.STED_BLEACH | _lib.STED_SAMPLE_BLEACH | _lib.STED_SAMPLE_PHOTONS)):
    call(hs[0], p, flags, 1); call(hs[1], p, flags, 2)
    t0 = time.perf_counter()
    for i in range(5): call(hs[2 + i], p, flags, 3 + i)
    dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {dt*1e3:.2f} ms per host call")
