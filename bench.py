#!/usr/bin/env python
"""bench.py — STED acquisitions/sec on the ContextualMOSTED acquisition schedule.

One *step* = the four acquisitions a ContextualMOSTED ``env.step`` performs for every environment of
the batch (gym_sted/envs/mosted_env.py:222-234 + :184): confocal, STED with in-scan bleaching
(stochastic, detector noise on), confocal on the bleached map, confocal on a fresh datamap.
Workload at N=1: BASELINE.json configs[1], ContextualMOSTED-easy-hslb-v0, 256 envs, 256x256 px,
K=59 PSF, actions uniform in the action box.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  ``--impl reference`` times the CPU restatement of pySTED's loop
(oracle/raster_oracle.c, kind "port": pySTED itself is not installable here) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ACQ_PER_STEP = 4
K = 59


def algorithmic_flops(H, W, bleach):
    n = H * W * K * K
    return (3 if bleach else 2) * n


def algorithmic_bytes(H, W, bleach):
    hp, wp = H + K - 1, W + K - 1
    return 4 * hp * wp + 4 * H * W + (4 * hp * wp if bleach else 0)


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi in the background during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        # NVML in-process (5 ms period) when nvidia_ml_py is importable, else the nvidia-smi loop of the profiling recipe
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = (pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.gpu))
            self.samples, self.stop_flag = [], threading.Event()
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv, h = self.nvml
        while not self.stop_flag.is_set():
            try:
                self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM),
                                     nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
            except Exception:
                pass
            time.sleep(0.005)

    def _stop_nvml(self):
        nv, _ = self.nvml
        self.stop_flag.set()
        self.thread.join(timeout=2)
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        reasons = sorted(n for n, b in bits.items() if any(s[2] & b for s in self.samples))
        sm = [s[0] for s in self.samples]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(s[1] for s in self.samples)) if sm else None,
                "samples": len(sm), "reasons": reasons, "source": "nvml"}

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            return self._stop_nvml()
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's C dwell loop on host threads (ctypes releases the GIL)
# ------------------------------------------------------------------------------------------------
def cpu_env_steps(n_envs, H, W, threads, seed=1234):
    """Runs `n_envs` full env-steps (4 acquisitions each) of the workload on `threads` host threads.
    Returns (seconds, acquisitions)."""
    from concurrent.futures import ThreadPoolExecutor
    from gym_sted_b200 import defaults, synapse
    from oracle import pysted_oracle as po
    from tests import helpers

    om = helpers.oracle_microscope(defaults.load_routine("high-signal_low-bleach"), noise=True,
                                   background=defaults.DETECTOR["background"])
    po.raster_lib()
    rng = np.random.default_rng(seed)
    lo = np.array([defaults.action_spaces[k]["low"] for k in ("p_sted", "p_ex", "pdt")])
    hi = np.array([defaults.action_spaces[k]["high"] for k in ("p_sted", "p_ex", "pdt")])
    actions = rng.uniform(lo, hi, (n_envs, 3))
    maps, _ = synapse.generate_batch(2 * n_envs, H, W, seed=seed)
    i_ex = om.cache(helpers.PX)[0]

    def one(i):
        dm = po.Datamap(maps[i], helpers.PX); dm.set_roi(i_ex, "max")
        conf = dict(pdt=defaults.PDT, p_ex=defaults.P_EX, p_sted=0.0)
        om.get_signal_and_bleach(dm, helpers.PX, **conf, bleach=False)
        om.get_signal_and_bleach(dm, helpers.PX, pdt=actions[i, 2], p_ex=actions[i, 1], p_sted=actions[i, 0],
                                 bleach=True, seed=i + 1)
        om.get_signal_and_bleach(dm, helpers.PX, **conf, bleach=False)
        dm2 = po.Datamap(maps[n_envs + i], helpers.PX); dm2.set_roi(i_ex, "max")
        om.get_signal_and_bleach(dm2, helpers.PX, **conf, bleach=False)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, range(n_envs)))
    return time.perf_counter() - t0, n_envs * ACQ_PER_STEP


def run_reference(args, rank, world):
    if rank != 0:
        return
    H = W = args.size
    cores = os.cpu_count() or 1
    n_envs = max(cores, 8) if args.cpu_envs is None else args.cpu_envs
    for _ in range(min(args.warmup, 1)):
        cpu_env_steps(min(cores, n_envs), H, W, cores)
    times, acqs = [], 0
    for _ in range(args.steps):
        t, a = cpu_env_steps(n_envs, H, W, cores)
        times.append(t); acqs += a
    total = sum(times)
    value = acqs / total
    sample = f"{n_envs} envs x 4 acquisitions per step ({H}x{W} px, K=59), {args.steps} steps, {cores} threads"
    line = {
        "impl": "reference", "metric": "STED acquisitions/sec", "value": value, "unit": "acquisitions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, n_envs, 1),
        "cpu_baseline": {"value": value, "unit": "acquisitions/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "acquisitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def ncu_traffic(kernel_substr):
    """dram__bytes_read + dram__bytes_write per launch (bytes) from the committed `ncu --set full` summary of the
    default workload (profiles/r01_ncu_summary.json); None when absent."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_summary.json")))
    except OSError:
        return None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = [float(r["dram__bytes_read.sum"]) * scale.get(r["dram__bytes_read.sum.unit"], 1.0)
            + float(r["dram__bytes_write.sum"]) * scale.get(r["dram__bytes_write.sum.unit"], 1.0)
            for r in table if kernel_substr in r["kernel"]]
    return float(np.mean(vals)) if vals else None


def l2_note(envs, size):
    """Bytes one step streams through the GPU (two padded map pools, their bleached copies and four images) against
    the 126 MB L2; nothing is flushed between timed iterations."""
    pad = size + K - 1
    mb = envs * (3 * pad * ((pad + 3) & ~3) + 4 * size * size) * 4 / 1e6
    if mb > 126:
        return f"inputs larger than L2: molecule maps + images of one step are {mb:.0f} MB (> 126 MB)"
    return f"inputs smaller than L2 ({mb:.0f} MB per step, not flushed): not a valid timing configuration, parity/contract runs only"


def workload_config(args, envs_per_gpu, world):
    return {"workload": f"ContextualMOSTED-easy-hslb-v0 acquisition schedule (conf, STED+bleach, conf, conf-on-new-map), "
                        f"{envs_per_gpu} envs/GPU, {args.size}x{args.size} px, K=59, stochastic bleaching + Poisson detection",
            "envs_per_gpu": envs_per_gpu, "global_envs": envs_per_gpu * world, "image": [args.size, args.size],
            "psf": [K, K], "fluorophore": "high-signal_low-bleach", "actions": "uniform in action box",
            "l2": l2_note(envs_per_gpu, args.size),
            "schedule": "sequential" if args.sequential else "next structure's confocal on a low-priority stream under the scan",
            "parallelism": f"env-parallel x{world}, no data-path collective"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        dist.barrier()
    from gym_sted_b200 import _lib, defaults, synapse
    from gym_sted_b200.batch import BatchedMicroscope
    from tests import helpers  # builders only (no oracle use on this path except cpu_baseline below)

    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    # several ranks streaming host buffers: keep each rank (and the pinned memory it allocates) on its GPU's NUMA node
    from gym_sted_b200 import dist as sted_dist
    numa_bound = sted_dist.bind_to_gpu_numa(local_rank) if world > 1 else False
    H = W = args.size
    B = args.envs
    env_base = rank * B
    micro = helpers.product_microscope(defaults.load_routine("high-signal_low-bleach"), noise=True,
                                       background=defaults.DETECTOR["background"])
    bm = BatchedMicroscope(micro, B, H, W, device=dev, env_base=env_base, seed=1234)
    # the next structure's confocal image does not depend on the current structure's acquisitions: it is acquired on
    # a low-priority stream with its own device buffers and fills the idle slots of the main stream's kernels (the
    # scan's second wave above all); --sequential issues the four acquisitions one after the other instead
    bm_next = None if args.sequential else BatchedMicroscope(micro, B, H, W, device=dev, env_base=env_base, seed=1234 ^ 0x5EED)
    main_stream = torch.cuda.Stream(dev, priority=-1)
    side_stream = torch.cuda.Stream(dev, priority=0)

    # synthetic inputs: two pools of B maps (current / fresh), resident in HBM and in pinned host memory
    n_unique = min(B, 32)
    maps_np, _ = synapse.generate_batch(2 * n_unique, H, W, seed=1234, env_base=env_base)
    reps = (B + n_unique - 1) // n_unique
    pool = [torch.from_numpy(np.tile(maps_np[i * n_unique:(i + 1) * n_unique], (reps, 1, 1))[:B]).to(torch.float32)
            for i in range(2)]
    pool_dev = [p.to(dev) for p in pool]
    pool_tot = [int(p.sum().item()) for p in pool]
    pool_pin = [p.pin_memory() for p in pool]
    g = torch.Generator().manual_seed(1234 + rank)
    lo = torch.tensor([defaults.action_spaces[k]["low"] for k in ("p_sted", "p_ex", "pdt")], dtype=torch.float64)
    hi = torch.tensor([defaults.action_spaces[k]["high"] for k in ("p_sted", "p_ex", "pdt")], dtype=torch.float64)
    n_act = args.steps + args.warmup + 2
    actions = (lo + (hi - lo) * torch.rand((n_act, B, 3), generator=g, dtype=torch.float64))
    actions_dev = actions.to(dev)
    conf = [torch.full((B,), v, dtype=torch.float64, device=dev) for v in (defaults.PDT, defaults.P_EX, 0.0)]
    sig = [torch.empty((B, H, W), dtype=torch.float32, device=dev) for _ in range(4)]
    launches = [0]
    kern_ev = []

    def step(i, timed):
        """The four acquisitions of one ContextualMOSTED env-step for the whole batch (mosted_env.py:184,222-234)."""
        a = actions_dev[i]
        bm.set_datamaps(pool_dev[0], max_molecules=pool_tot[0])
        ev = None
        if timed:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        bm.make_effective(conf[1], conf[2], conf[0])
        _gather_only(bm, conf[0], sig[0], ev)
        cur = torch.cuda.current_stream(dev)
        if bm_next is not None:
            side_stream.wait_stream(cur)                 # after the bracketed gather: its timing stays undisturbed
            with torch.cuda.stream(side_stream):
                bm_next.set_datamaps(pool_dev[1], max_molecules=pool_tot[1])
                bm_next.get_signal_and_bleach(conf[0], conf[1], conf[2], bleach=False, out_signal=sig[3])
        _scan_only(bm, a, sig[1], timed)
        bm.get_signal_and_bleach(conf[0], conf[1], conf[2], bleach=False, out_signal=sig[2])
        if bm_next is not None:
            cur.wait_stream(side_stream)
        else:
            bm.set_datamaps(pool_dev[1], max_molecules=pool_tot[1])
            bm.get_signal_and_bleach(conf[0], conf[1], conf[2], bleach=False, out_signal=sig[3])
        launches[0] += 15     # 3 x (make_effective, gather, detect) + (make_effective, presample x2, bucket scan, sweep, detect)

    scan_ev = []

    def _scan_only(bm, a, out, timed):
        """STED acquisition with stochastic bleaching; the scan (presample + bucket scan + sweep) is bracketed."""
        import ctypes
        pdt, p_ex, p_sted = a[:, 2].contiguous(), a[:, 1].contiguous(), a[:, 0].contiguous()
        bm.make_effective(p_ex, p_sted, pdt)
        bm.offset += 1
        st = torch.cuda.current_stream(dev).cuda_stream
        flags = _lib.STED_BLEACH | _lib.STED_SAMPLE_BLEACH | _lib.STED_SAMPLE_PHOTONS
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) if timed else None
        if ev: ev[0].record()
        _lib.check(bm.lib.sted_acquire(bm.dmap.data_ptr(), bm.dmap_alt.data_ptr(), bm.tables.data_ptr(), pdt.data_ptr(), bm.B,
                                       bm.H, bm.W, bm.K, flags | _lib.STED_NO_DETECT, ctypes.byref(bm.det), bm.lambda_fluo,
                                       bm.seed, bm.offset, bm.env_base, None, out.data_ptr(), bm.workspace.data_ptr(),
                                       bm.workspace_bytes, st))
        if ev: ev[1].record(); scan_ev.append(ev)
        _lib.check(bm.lib.sted_detect(out.data_ptr(), pdt.data_ptr(), bm.B, bm.H, bm.W, flags, ctypes.byref(bm.det),
                                      bm.lambda_fluo, bm.seed, bm.offset, bm.env_base, st))
        bm.dmap, bm.dmap_alt = bm.dmap_alt, bm.dmap

    def _gather_only(bm, pdt, out, ev):
        import ctypes
        bm.offset += 1
        st = torch.cuda.current_stream(dev).cuda_stream
        flags = _lib.STED_SAMPLE_PHOTONS
        if ev: ev[0].record()
        _lib.check(bm.lib.sted_acquire(bm.dmap.data_ptr(), None, bm.tables.data_ptr(), pdt.data_ptr(), bm.B, bm.H,
                                       bm.W, bm.K, flags | _lib.STED_NO_DETECT, ctypes.byref(bm.det), bm.lambda_fluo,
                                       bm.seed, bm.offset, bm.env_base, None, out.data_ptr(), None, 0, st))
        if ev: ev[1].record(); kern_ev.append(ev)
        _lib.check(bm.lib.sted_detect(out.data_ptr(), pdt.data_ptr(), bm.B, bm.H, bm.W, flags, ctypes.byref(bm.det),
                                      bm.lambda_fluo, bm.seed, bm.offset, bm.env_base, st))

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    sampler.start()
    torch.cuda.synchronize(dev)
    torch.cuda.set_stream(main_stream)       # everything below is issued on the high-priority stream
    for i in range(args.warmup):
        step(i, False)
    sync_all()
    t_wait = time.perf_counter()
    while sampler.proc is not None and not sampler.lines and time.perf_counter() - t_wait < 3.0:
        time.sleep(0.02)                     # nvidia-smi needs a moment before its first sample
    sampler.lines.clear()                    # keep only samples taken from here on (the timed region)
    if sampler.nvml is not None:
        sampler.samples.clear()
    launches[0] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i, True)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    n_launch = launches[0]                   # kernels launched inside the timed region
    if len(sampler.samples if sampler.nvml is not None else sampler.lines) < 3:      # very short timed region: keep the GPU under the same load a little longer
        t_more = time.perf_counter()
        while len(sampler.samples if sampler.nvml is not None else sampler.lines) < 3 and time.perf_counter() - t_more < 1.0:
            step(args.warmup, False)
            torch.cuda.synchronize(dev)
    clocks = sampler.stop()
    gather_ms = float(np.mean([a.elapsed_time(b) for a, b in kern_ev]))
    scan_ms = float(np.mean([a.elapsed_time(b) for a, b in scan_ev]))
    # the scan on its own (nothing co-running): three extra scans after the timed region
    scan_alone_ms = scan_ms
    if bm_next is not None:
        n_before = len(scan_ev)
        for i in range(3):
            bm.set_datamaps(pool_dev[0], max_molecules=pool_tot[0])
            _scan_only(bm, actions_dev[args.warmup + i % max(args.steps, 1)], sig[1], True)
            torch.cuda.synchronize(dev)
        scan_alone_ms = float(np.mean([a.elapsed_time(b) for a, b in scan_ev[n_before:]]))

    # ---- e2e: the reference-facing call with HOST buffers (pinned), copies inside the timed region
    import ctypes
    lib = bm.lib
    cache = micro._packed_cache(helpers.PX)
    from gym_sted_b200.base import fluo_struct
    fl = (_lib.StedFluoParams * B)(*[fluo_struct(micro.fluo, micro.excitation, micro.sted)] * B)
    det = micro.detector.params()
    # host buffers of the e2e leg: uint16 molecule maps / photon counts (STED_HOST_U16; the env's observation is uint16
    # too, contextual_sted_env.py:93), pinned
    h_sig = torch.empty((B, H, W), dtype=torch.uint16).pin_memory()
    conf_np = [np.full(B, v) for v in (defaults.P_EX, 0.0, defaults.PDT)]
    vp = ctypes.c_void_p
    e2e_steps = max(1, min(args.steps, 5))
    # the STED call overwrites its host map with the bleached one, so every e2e step gets its own pinned copy of the
    # current maps (prepared before the timed region); the fresh maps are only read
    h_cur = [pool[0].to(torch.uint16).pin_memory() for _ in range(e2e_steps + 1)]
    h_new = pool_pin[1].to(torch.uint16).pin_memory()

    def host_acquire(h_map, p_ex, p_sted, pdt, flags, off, out=None):
        _lib.check(lib.sted_acquire_host(h_map.data_ptr(), cache.ctypes.data_as(vp), ctypes.cast(fl, vp),
                                         p_ex.ctypes.data_as(vp), p_sted.ctypes.data_as(vp), pdt.ctypes.data_as(vp),
                                         B, H, W, K, flags | _lib.STED_HOST_U16, ctypes.byref(det), 1234, off, env_base, None,
                                         (h_sig if out is None else out).data_ptr()))

    # The next structure's confocal does not depend on the other three calls of the step: a second host thread issues it
    # on low-priority streams (sted_acquire_host keeps its scratch and streams per thread; ctypes drops the GIL).
    import queue
    import threading
    jobs, finished, failure = queue.Queue(), queue.Queue(), []
    h_sig_next = torch.empty((B, H, W), dtype=torch.uint16).pin_memory()

    def side_worker():
        torch.cuda.set_device(dev)
        try:
            _lib.check(lib.sted_host_thread_priority(1))
        except Exception as exc:                       # pragma: no cover - reported through the queue
            failure.append(exc)
        while True:
            off = jobs.get()
            if off is None:
                return
            try:
                host_acquire(h_new, *conf_np, _lib.STED_SAMPLE_PHOTONS, off, out=h_sig_next)
            except Exception as exc:                   # pragma: no cover
                failure.append(exc)
            finished.put(off)

    worker = None
    if not args.sequential:
        worker = threading.Thread(target=side_worker, daemon=True)
        worker.start()

    def e2e_step(i):
        a = actions[i].numpy()
        h_map = h_cur[i]
        if worker is not None:
            jobs.put(4 * i + 4)
        host_acquire(h_map, *conf_np, _lib.STED_SAMPLE_PHOTONS, 4 * i + 1)
        host_acquire(h_map, np.ascontiguousarray(a[:, 1]), np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 2]),
                     _lib.STED_BLEACH | _lib.STED_SAMPLE_BLEACH | _lib.STED_SAMPLE_PHOTONS, 4 * i + 2)
        host_acquire(h_map, *conf_np, _lib.STED_SAMPLE_PHOTONS, 4 * i + 3)
        if worker is not None:
            finished.get()
            if failure:
                raise failure[0]
        else:
            host_acquire(h_new, *conf_np, _lib.STED_SAMPLE_PHOTONS, 4 * i + 4)

    e2e_step(0)
    sync_all()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i + 1)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    if worker is not None:
        jobs.put(None)
        worker.join(timeout=10)
    img_bytes = B * H * W * 2          # uint16 transport
    h2d = 4 * img_bytes + 4 * (3 * K * K * 8 + B * (136 + 24))
    d2h = 4 * img_bytes + img_bytes

    # ---- max over ranks
    t = torch.tensor([ms, e2e_s, gather_ms, scan_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_s, gather_ms, scan_ms = [float(x) for x in t.tolist()]

    if rank == 0:
        acq = ACQ_PER_STEP * B * world * args.steps
        value = acq / (ms * 1e-3)
        e2e_value = ACQ_PER_STEP * B * world * e2e_steps / e2e_s
        # roofline of the dominant kernel (gather_kernel<59>: 3 of 4 acquisitions)
        peak = [0.0]
        pk = ctypes.c_double()
        _lib.check(lib.sted_fma_peak(ctypes.byref(pk), None))
        fl_launch = algorithmic_flops(H, W, False) * B
        achieved = fl_launch / (gather_ms * 1e-3) * 1e-12
        by_launch = algorithmic_bytes(H, W, False) * B
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        roofline = {"bound": "fp32", "kernel": "gather_kernel<59>", "achieved": achieved, "peak": pk.value,
                    "unit": "TFLOP/s", "frac": achieved / pk.value if pk.value else None,
                    "traffic": ncu_traffic("gather_kernel") if (H, B) == (256, 256) else None,
                    "peak_source": "FP32 FFMA micro-benchmark run in this process (sted_fma_peak); "
                                   "MEASURED_PEAKS.json has no FP32 entry",
                    "algorithmic_flops_per_launch": fl_launch, "launch_ms": gather_ms,
                    "hbm": {"achieved": by_launch / (gather_ms * 1e-3) * 1e-9, "peak": peaks.get("hbm_gbs", 6650.0),
                            "unit": "GB/s", "peak_source": "measured" if peaks else "fallback"}}
        fl_scan = algorithmic_flops(H, W, True) * B
        roofline_scan = {"bound": "fp32", "kernel": "presample_kernel x2 + bucket_scan_kernel + sweep_kernel<59,16,stoch> "
                                                    "(one STED acquisition with in-scan bleaching)",
                         "achieved": fl_scan / (scan_ms * 1e-3) * 1e-12, "peak": pk.value, "unit": "TFLOP/s",
                         "frac": fl_scan / (scan_ms * 1e-3) * 1e-12 / pk.value if pk.value else None,
                         "algorithmic_flops_per_launch": fl_scan, "launch_ms": scan_ms,
                         "alone": {"launch_ms": scan_alone_ms, "frac": fl_scan / (scan_alone_ms * 1e-3) * 1e-12 / pk.value if pk.value else None,
                                   "note": "same scan with nothing co-running, 3 launches after the timed region"},
                         "note": "3N algorithmic flops (SURVEY 8d); the kernels execute 2N" + ("" if args.sequential else
                                 "; timed with the next structure's low-priority gather co-running (run --sequential for the scan alone)")}
        cpu = None
        if not args.no_cpu:
            cores = os.cpu_count() or 1
            n_envs = args.cpu_envs or max(cores, 8)
            tcpu, a = cpu_env_steps(n_envs, H, W, cores)
            t1, a1 = cpu_env_steps(1, H, W, 1)                       # the same loop on one core
            cpu = {"value": a / tcpu, "unit": "acquisitions/s", "cores": cores, "kind": "port",
                   "single_core_value": a1 / t1,
                   "sample": f"{n_envs} envs x 4 acquisitions of the same workload ({H}x{W}, K=59) on {cores} host "
                             f"threads, {tcpu:.1f} s (+ 1 env on 1 thread, {t1:.1f} s); C restatement of pySTED's dwell "
                             f"loop (oracle/raster_oracle.c), not pySTED"}
        line = {
            "metric": "STED acquisitions/sec", "value": value, "unit": "acquisitions/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, B, world), "env_steps_per_s": value / ACQ_PER_STEP,
            "e2e": {"value": e2e_value, "unit": "acquisitions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "sted_acquire_host (C ABI, pinned host buffers, uint16 counts: STED_HOST_U16)" + ("" if args.sequential else "; the next structure's call comes from a second host thread on low-priority streams"), "numa_bound": numa_bound},
            "gpu_launches": n_launch, "clocks": clocks, "roofline": roofline, "roofline_scan": roofline_scan,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=256, help="environments per GPU")
    ap.add_argument("--size", type=int, default=256, help="image side in pixels")
    ap.add_argument("--cpu-envs", type=int, default=None, help="env-steps in the CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--sequential", action="store_true",
                    help="issue the four acquisitions of a step one after the other (no overlap of the next structure's confocal)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
