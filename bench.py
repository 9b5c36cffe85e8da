#!/usr/bin/env python
"""bench.py — STED acquisitions/sec on the ContextualMOSTED acquisition schedule.

One *step* = the four acquisitions a ContextualMOSTED ``env.step`` performs for every environment of
the batch (gym_sted/envs/mosted_env.py:222-234 + :184): confocal, STED with in-scan bleaching
(stochastic, detector noise on), confocal on the bleached map, confocal on a fresh datamap.
Workload at N=1: BASELINE.json configs[1], ContextualMOSTED-easy-hslb-v0, 256 envs, 256x256 px,
K=59 PSF, actions uniform in the action box.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload acquisitions|config5|env_step]

``--workload config5``: BASELINE.json configs[4], 512x512 px, 2048 environments in total, split over the N GPUs
(strong scaling: 2048 / N environments per GPU).  ``--workload env_step``: whole ``VectorContextualMOSTED.step`` calls
(4 acquisitions + foreground / objectives / nanodomain F1 + observation packing, results as numpy arrays on the host).

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


def gather_executed_fraction(maps, tile=64, rows_per_thread=4):
    """Share of gather_kernel's multiply-adds that are issued for these (n, H, W) molecule maps: a warp (two groups of
    `rows_per_thread` output rows, all `tile` columns) skips a band row when both datamap rows it would read hold no
    molecule inside the tile's window (G_SKIP in csrc/sted_kernels.cu).  Same geometry as the kernel (32-row tiles when
    H mod 64 is in 1..32, as launch_gather chooses), counted on the host."""
    maps = np.asarray(maps)
    n, H, W = maps.shape
    tile_h = 32 if -(-H // 32) * 32 < -(-H // 64) * 64 else 64
    warps = tile_h // (2 * rows_per_thread)
    p, kp = K // 2, ((K + 1) + 3) & ~3
    span_r, span_c = tile_h + K - 1, tile + kp
    ty, tx = -(-H // tile_h), -(-W // tile)
    nz = np.zeros((n, ty * tile_h + span_r, tx * tile + span_c), dtype=bool)
    nz[:, p:p + H, p:p + W] = maps != 0
    issued = total = 0
    for cx in range(tx):
        row_any = nz[:, :, cx * tile:cx * tile + span_c].any(axis=2)          # (n, rows): a molecule inside the window's columns
        for cy in range(ty):
            for w in range(warps):
                for yy in range(rows_per_thread + K - 1):
                    pairs = sum(0 <= yy - 2 * q <= K for q in range(rows_per_thread // 2))   # live row pairs of this band row
                    r = cy * tile_h + w * 2 * rows_per_thread + yy
                    live = row_any[:, r] | row_any[:, r + rows_per_thread]
                    issued += pairs * int(live.sum())
                    total += pairs * n
    return issued / total


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

    PX = defaults.PIXELSIZE
    om = po.default_microscope(defaults.load_routine("high-signal_low-bleach"),
                               {"noise": True, "background": defaults.DETECTOR["background"]},
                               defaults.LASER_EX, defaults.LASER_STED, defaults.OBJECTIVE)
    cache_file = os.path.join(ROOT, "tests", "golden", "oracle_psf_cache.npz")     # the oracle's quad beams, memoised
    if os.path.exists(cache_file):
        z = np.load(cache_file)
        om._cache[int(round(PX * 1e9))] = (z["i_ex"], z["i_sted"], z["psf_det"])
    po.raster_lib()
    rng = np.random.default_rng(seed)
    lo = np.array([defaults.action_spaces[k]["low"] for k in ("p_sted", "p_ex", "pdt")])
    hi = np.array([defaults.action_spaces[k]["high"] for k in ("p_sted", "p_ex", "pdt")])
    actions = rng.uniform(lo, hi, (n_envs, 3))
    maps, _ = synapse.generate_batch(2 * n_envs, H, W, seed=seed)
    i_ex = om.cache(PX)[0]

    def one(i):
        dm = po.Datamap(maps[i], PX); dm.set_roi(i_ex, "max")
        conf = dict(pdt=defaults.PDT, p_ex=defaults.P_EX, p_sted=0.0)
        om.get_signal_and_bleach(dm, PX, **conf, bleach=False)
        om.get_signal_and_bleach(dm, PX, pdt=actions[i, 2], p_ex=actions[i, 1], p_sted=actions[i, 0],
                                 bleach=True, seed=i + 1)
        om.get_signal_and_bleach(dm, PX, **conf, bleach=False)
        dm2 = po.Datamap(maps[n_envs + i], PX); dm2.set_roi(i_ex, "max")
        om.get_signal_and_bleach(dm2, PX, **conf, bleach=False)

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
        "config": workload_config(args, n_envs, 1, schedule=f"one environment per host thread ({cores} threads), the four "
                                  f"acquisitions of an env-step one after the other; {n_envs} environments per step — the per-"
                                  f"environment work is the GPU arm's, the batch is bounded so that a step takes seconds"),
        "cpu_baseline": {"value": value, "unit": "acquisitions/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "acquisitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line["config"]["l2"] = "n/a: CPU arm (host caches; nothing runs on the GPU)"
    print(json.dumps(line), flush=True)


def ncu_traffic(kernel_substr):
    """dram__bytes_read + dram__bytes_write per launch (bytes) from the committed `ncu --set full` summary of the
    default workload (the newest of profiles/r02*_ncu_gather.json); None when absent."""
    table = None
    for tag in ("r02d", "r02b", "r02"):
        try:
            table = json.load(open(os.path.join(ROOT, "profiles", f"{tag}_ncu_gather.json")))
            break
        except OSError:
            continue
    if table is None:
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


def workload_config(args, envs_per_gpu, world, schedule=None):
    name = {"acquisitions": "ContextualMOSTED-easy-hslb-v0 acquisition schedule",
            "config5": "BASELINE config 5 (512x512 synthetic spine datamaps, 2048 envs env-parallel), ContextualMOSTED acquisition schedule",
            "env_step": "ContextualMOSTED-easy-hslb-v0 whole env.step (acquisitions + objectives + F1 + observation)"}[args.workload]
    if schedule is None:
        schedule = ("sequential: one session, every kernel on one stream" if args.sequential else
                    "sessions: tables + death schedule of an acquisition under the one before it, detector under the one after "
                    "it; next structure's confocal on a low-priority session")
    return {"workload": f"{name} (conf, STED+bleach, conf, conf-on-new-map), "
                        f"{envs_per_gpu} envs/GPU, {args.size}x{args.size} px, K=59, stochastic bleaching + Poisson detection",
            "envs_per_gpu": envs_per_gpu, "global_envs": envs_per_gpu * world, "image": [args.size, args.size],
            "psf": [K, K], "fluorophore": "high-signal_low-bleach", "actions": "uniform in action box",
            "l2": l2_note(envs_per_gpu, args.size), "schedule": schedule,
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
    from gym_sted_b200 import _lib, defaults, envs, synapse
    from gym_sted_b200.batch import BatchedMicroscope
    from gym_sted_b200.session import StedSession, pinned_empty

    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    # several ranks streaming host buffers: keep each rank (and the pinned memory it allocates) on its GPU's NUMA node
    from gym_sted_b200 import dist as sted_dist
    numa_bound = sted_dist.bind_to_gpu_numa(local_rank) if world > 1 else False
    H = W = args.size
    B = args.envs
    env_base = rank * B
    micro = envs.make_microscope(defaults.load_routine("high-signal_low-bleach"))     # detector: defaults.DETECTOR (noise + background)
    if args.sequential:
        os.environ["STED_B200_SESSION_SERIAL"] = "1"        # sessions queue every kernel on one stream (A/B of the schedule)
    # The timed path drives resident SESSIONS (sted_session_*, include/sted_b200.h) with device-resident structures: a
    # session runs the tables and the death schedule of an acquisition under the gather / scan queued before it and
    # the detector under the one queued after it (own streams).  The next structure's confocal image does not depend on
    # the current structure's acquisitions: a second, low-priority session uploads and images it and the main session
    # then adopts the map, as gym-sted's step does (mosted_env.py:174-186).  --sequential: ONE session, one stream.
    main_sess = StedSession(micro, B, H, W, env_base=env_base, seed=1234)
    next_sess = None if args.sequential else StedSession(micro, B, H, W, env_base=env_base, seed=1234 ^ 0x5EED, low_priority=True)
    bm = BatchedMicroscope(micro, B, H, W, device=dev, env_base=env_base, seed=1234)     # single-stream brackets ("alone")

    # synthetic inputs: two pools of B maps (current / fresh), resident in HBM and in pinned host memory
    n_unique = min(B, 32)
    maps_np, _ = synapse.generate_batch(2 * n_unique, H, W, seed=1234, env_base=env_base)
    reps = (B + n_unique - 1) // n_unique
    pool = [torch.from_numpy(np.tile(maps_np[i * n_unique:(i + 1) * n_unique], (reps, 1, 1))[:B]).to(torch.float32)
            for i in range(2)]
    pool_dev = [p.to(dev).contiguous() for p in pool]
    pool_tot = [int(p.sum().item()) for p in pool]
    mol = max(pool_tot)
    g = torch.Generator().manual_seed(1234 + rank)
    lo = torch.tensor([defaults.action_spaces[k]["low"] for k in ("p_sted", "p_ex", "pdt")], dtype=torch.float64)
    hi = torch.tensor([defaults.action_spaces[k]["high"] for k in ("p_sted", "p_ex", "pdt")], dtype=torch.float64)
    n_act = args.steps + args.warmup + 2
    actions = (lo + (hi - lo) * torch.rand((n_act, B, 3), generator=g, dtype=torch.float64))
    actions_dev = actions.to(dev)
    actions_np = actions.numpy()
    conf_np = (defaults.PDT, defaults.P_EX, 0.0)
    conf = [torch.full((B,), v, dtype=torch.float64, device=dev) for v in (defaults.PDT, defaults.P_EX, 0.0)]
    sig = [torch.empty((B, H, W), dtype=torch.float32, device=dev) for _ in range(2)]
    launches = [0]

    def step(i):
        """The four acquisitions of one ContextualMOSTED env-step for the whole batch (mosted_env.py:184,222-234); the
        structures and the images stay on the device."""
        a = actions_np[i]
        if next_sess is not None:
            next_sess.upload_device(pool_dev[(i + 1) % 2], max_molecules=mol)
            next_sess.acquire(*conf_np, bleach=False)
        main_sess.acquire(*conf_np, bleach=False)
        main_sess.acquire(np.ascontiguousarray(a[:, 2]), np.ascontiguousarray(a[:, 1]), np.ascontiguousarray(a[:, 0]), bleach=True)
        main_sess.acquire(*conf_np, bleach=False)
        if next_sess is not None:
            main_sess.adopt(next_sess)
        else:
            main_sess.upload_device(pool_dev[(i + 1) % 2], max_molecules=mol)
            main_sess.acquire(*conf_np, bleach=False)
        # pad + 3 x (gather, detect) + (make_effective, presample, bucket scan, scatter, sweep2, detect); the confocal
        # acquisitions find their tables in the session (same imaging parameters at every call)
        launches[0] += 13

    def join_all():
        main_sess.join()
        if next_sess is not None:
            next_sess.join()

    def _scan_only(bm, a, out, evs):
        """STED acquisition with stochastic bleaching on ONE stream; the scan (presample + bucket scan + scatter + sweep)
        is bracketed."""
        import ctypes
        pdt, p_ex, p_sted = a[:, 2].contiguous(), a[:, 1].contiguous(), a[:, 0].contiguous()
        bm.make_effective(p_ex, p_sted, pdt)
        bm.offset += 1
        st = torch.cuda.current_stream(dev).cuda_stream
        flags = _lib.STED_BLEACH | _lib.STED_SAMPLE_BLEACH | _lib.STED_SAMPLE_PHOTONS
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
        _lib.check(bm.lib.sted_acquire(bm.dmap.data_ptr(), bm.dmap_alt.data_ptr(), bm.tables.data_ptr(), pdt.data_ptr(), bm.B,
                                       bm.H, bm.W, bm.K, flags | _lib.STED_NO_DETECT, ctypes.byref(bm.det), bm.lambda_fluo,
                                       bm.seed, bm.offset, bm.env_base, None, out.data_ptr(), bm.workspace.data_ptr(),
                                       bm.workspace_bytes, st))
        ev[1].record(); evs.append(ev)
        _lib.check(bm.lib.sted_detect(out.data_ptr(), pdt.data_ptr(), bm.B, bm.H, bm.W, flags, ctypes.byref(bm.det),
                                      bm.lambda_fluo, bm.seed, bm.offset, bm.env_base, st))
        bm.dmap, bm.dmap_alt = bm.dmap_alt, bm.dmap

    def _gather_only(bm, pdt, out, evs):
        import ctypes
        bm.offset += 1
        st = torch.cuda.current_stream(dev).cuda_stream
        flags = _lib.STED_SAMPLE_PHOTONS
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
        _lib.check(bm.lib.sted_acquire(bm.dmap.data_ptr(), None, bm.tables.data_ptr(), pdt.data_ptr(), bm.B, bm.H,
                                       bm.W, bm.K, flags | _lib.STED_NO_DETECT, ctypes.byref(bm.det), bm.lambda_fluo,
                                       bm.seed, bm.offset, bm.env_base, None, out.data_ptr(), None, 0, st))
        ev[1].record(); evs.append(ev)
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
    main_stream = torch.cuda.Stream(dev, priority=-1)
    torch.cuda.set_stream(main_stream)       # the stream the timing events live on; it joins the sessions' streams
    main_sess.upload_device(pool_dev[0], max_molecules=mol)
    for i in range(args.warmup):
        step(i)
    join_all()
    sync_all()
    t_wait = time.perf_counter()
    while sampler.proc is not None and not sampler.lines and time.perf_counter() - t_wait < 3.0:
        time.sleep(0.02)                     # nvidia-smi needs a moment before its first sample
    sampler.lines.clear()                    # keep only samples taken from here on (the timed region)
    if sampler.nvml is not None:
        sampler.samples.clear()
    launches[0] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()                              # the GPU is idle: this is the start of the region on the device
    for i in range(args.steps):
        step(args.warmup + i)
    join_all()                               # main_stream waits for every stream of both sessions ...
    e1.record()                              # ... so this fires when the last kernel of the region has finished
    sync_all()
    ms = e0.elapsed_time(e1)
    assert main_sess.sync() == 0 and (next_sess is None or next_sess.sync() == 0), "bleaching events were dropped"
    n_launch = launches[0]                   # kernels launched inside the timed region
    # device time of the gather / scan kernels inside the timed region (CUDA events on the session's compute stream)
    per_step = 3 if next_sess is not None else 4
    n_back = min(per_step * args.steps, 64)
    k_ms = [main_sess.kernel_ms(k) for k in range(n_back)][::-1]          # oldest first
    first = (per_step * args.steps - n_back) % per_step
    gather_in = [t for k, t in enumerate(k_ms) if (k + first) % per_step != 1]
    sweep_in = [t for k, t in enumerate(k_ms) if (k + first) % per_step == 1]
    if len(sampler.samples if sampler.nvml is not None else sampler.lines) < 3:      # very short timed region: keep the GPU under the same load a little longer
        t_more = time.perf_counter()
        while len(sampler.samples if sampler.nvml is not None else sampler.lines) < 3 and time.perf_counter() - t_more < 1.0:
            step(args.warmup)
            join_all()
            torch.cuda.synchronize(dev)
    clocks = sampler.stop()
    gather_ms = float(np.mean(gather_in))
    sweep_ms = float(np.mean(sweep_in))
    # the same kernels with nothing co-running: single-stream device API, five repetitions after the timed region
    kern_ev, scan_ev = [], []
    for i in range(5):
        bm.set_datamaps(pool_dev[0], max_molecules=pool_tot[0])
        bm.make_effective(conf[1], conf[2], conf[0])
        torch.cuda.synchronize(dev)
        _gather_only(bm, conf[0], sig[0], kern_ev)
        torch.cuda.synchronize(dev)
        _scan_only(bm, actions_dev[args.warmup + i % max(args.steps, 1)], sig[1], scan_ev)
        torch.cuda.synchronize(dev)
    gather_alone_ms = float(np.mean([a.elapsed_time(b) for a, b in kern_ev]))
    scan_alone_ms = float(np.mean([a.elapsed_time(b) for a, b in scan_ev]))
    main_sess.close()
    if next_sess is not None:
        next_sess.close()

    # ---- e2e: the reference-facing call with HOST buffers (pinned), copies inside the timed region.
    # Sessions (sted_session_*): a structure is uploaded ONCE, as the next structure of the step before (its first confocal
    # image is that step's observation, mosted_env.py:174-186), then the main session adopts it on the device and images
    # it three times (conf, STED + bleaching, conf) and returns the bleached map.  Per step: 1 map up, 4 images + 1 map
    # down, all uint16 (the env's observation dtype), host synchronisation at the end of every step.
    import ctypes
    lib = bm.lib
    e2e_steps = max(1, min(args.steps, 5))
    main_sess = StedSession(micro, B, H, W, env_base=env_base, seed=1234)
    next_sess = StedSession(micro, B, H, W, env_base=env_base, seed=1234 ^ 0x5EED, low_priority=not args.sequential)
    h_maps = [pinned_empty((B, H, W), np.uint16) for _ in range(2)]
    for i in range(2):
        h_maps[i][...] = pool[i].numpy().astype(np.uint16)
    h_out = [pinned_empty((B, H, W), np.uint16) for _ in range(4)]
    h_bleached = pinned_empty((B, H, W), np.uint16)

    def e2e_step(i):
        a = actions[i].numpy()
        next_sess.upload(h_maps[(i + 1) % 2], max_molecules=mol)
        next_sess.acquire(*conf_np, bleach=False, out=h_out[3])
        main_sess.acquire(*conf_np, bleach=False, out=h_out[0])
        main_sess.acquire(np.ascontiguousarray(a[:, 2]), np.ascontiguousarray(a[:, 1]), np.ascontiguousarray(a[:, 0]),
                          bleach=True, out=h_out[1])
        main_sess.download(h_bleached)
        main_sess.acquire(*conf_np, bleach=False, out=h_out[2])
        main_sess.adopt(next_sess)                     # the next step images the structure uploaded in this one
        dropped = main_sess.sync() + next_sess.sync()
        assert dropped == 0

    main_sess.upload(h_maps[0], max_molecules=mol)
    e2e_step(0)
    sync_all()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i + 1)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    e2e_check = int(h_out[1].astype(np.int64).sum())          # the step's result was read on the host
    main_sess.close(); next_sess.close()
    img_bytes = B * H * W * 2          # uint16 transport
    h2d = img_bytes + 4 * (3 * B * 8)
    d2h = 4 * img_bytes + img_bytes

    # ---- the host's transfer ceiling with every rank copying at once (no kernels): the step's five device->host buffers,
    # 4 rounds, then the map upload.  The e2e leg cannot be faster than d2h_bytes_per_step / this rate.
    d_buf = torch.empty((B, H, W), dtype=torch.uint16, device=dev)
    t_pins = [h._owner for h in h_out + [h_bleached]]        # the pinned torch tensors behind pinned_empty's arrays
    cs = torch.cuda.Stream(dev)

    def copy_rate(down):
        def burst(n):
            with torch.cuda.stream(cs):
                for _ in range(n):
                    for tp in t_pins:
                        tp.copy_(d_buf, non_blocking=True) if down else d_buf.copy_(tp, non_blocking=True)
            cs.synchronize()
        burst(1)
        sync_all()
        c0 = time.perf_counter()
        burst(4)
        return 4 * len(t_pins) * img_bytes / (time.perf_counter() - c0) * 1e-9

    d2h_gbs, h2d_gbs = copy_rate(True), copy_rate(False)

    # ---- the one collective of the design: per-step rewards gathered on rank 0 over NCCL (outside the timed region)
    gathered = None
    if world > 1:
        r_local = torch.full((B,), float(rank), dtype=torch.float32, device=dev)
        allr = sted_dist.gather_rewards(r_local, world * B)
        assert allr.shape == (world * B,) and all(float(allr[k * B]) == k for k in range(world)), "NCCL reward gather"
        gathered = int(allr.numel())

    # ---- max over ranks
    t = torch.tensor([ms, e2e_s, gather_ms, sweep_ms, scan_alone_ms, gather_alone_ms, -d2h_gbs, -h2d_gbs], dtype=torch.float64,
                     device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_s, gather_ms, sweep_ms, scan_alone_ms, gather_alone_ms, d2h_gbs, h2d_gbs = [float(x) for x in t.tolist()]
    d2h_gbs, h2d_gbs = -d2h_gbs, -h2d_gbs                   # slowest rank

    if rank == 0:
        acq = ACQ_PER_STEP * B * world * args.steps
        value = acq / (ms * 1e-3)
        e2e_value = ACQ_PER_STEP * B * world * e2e_steps / e2e_s
        # roofline of the dominant kernel (gather_kernel<59>: 3 of 4 acquisitions)
        peak = [0.0]
        pk = ctypes.c_double()
        _lib.check(lib.sted_fma_peak(ctypes.byref(pk), None))
        fl_launch = algorithmic_flops(H, W, False) * B
        exec_frac = gather_executed_fraction(maps_np[:min(len(maps_np), 8)])
        achieved = fl_launch / (gather_ms * 1e-3) * 1e-12
        by_launch = algorithmic_bytes(H, W, False) * B
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        roofline = {"bound": "fp32", "kernel": "gather_kernel<59, 64>" if -(-H // 32) * 32 >= -(-H // 64) * 64 else "gather_kernel<59, 32>", "achieved": achieved, "peak": pk.value,
                    "unit": "TFLOP/s", "frac": achieved / pk.value if pk.value else None,
                    "traffic": ncu_traffic("gather_kernel") if (H, B) == (256, 256) else None,
                    "peak_source": "FP32 FFMA micro-benchmark run in this process (sted_fma_peak); "
                                   "MEASURED_PEAKS.json has no FP32 entry",
                    "algorithmic_flops_per_launch": fl_launch, "launch_ms": gather_ms,
                    "executed": {"flop_fraction": exec_frac, "tflops_alone": exec_frac * fl_launch / (gather_alone_ms * 1e-3) * 1e-12,
                                 "frac_alone": exec_frac * fl_launch / (gather_alone_ms * 1e-3) * 1e-12 / pk.value if pk.value else None,
                                 "note": "the kernel skips band rows whose datamap rows hold no molecule (E * 0 terms; results are "
                                         "bit-identical), so `achieved` - ALGORITHMIC flops over time, the contract's definition - can "
                                         "exceed the FP32 peak; flop_fraction is the share of the multiply-adds that are issued for "
                                         "these maps (counted on the host with the kernel's geometry) and frac_alone the FP32-pipe "
                                         "fraction of the issued work"},
                    "launch_ms_note": "average over the gather kernels inside the timed region (CUDA events on the session's "
                                      "compute stream, sted_session_kernel_ms)" + ("" if args.sequential else
                                      "; the detector of the previous acquisition, the tables / death schedule of the next one "
                                      "and the low-priority session's kernels share the SMs with it"),
                    "alone": {"launch_ms": gather_alone_ms, "frac": fl_launch / (gather_alone_ms * 1e-3) * 1e-12 / pk.value if pk.value else None,
                              "note": "the same kernel with nothing else on the GPU (single-stream device API, 5 launches after the "
                                      "timed region)"},
                    "hbm": {"achieved": by_launch / (gather_ms * 1e-3) * 1e-9, "peak": peaks.get("hbm_gbs", 6650.0),
                            "unit": "GB/s", "peak_source": "measured" if peaks else "fallback"}}
        fl_scan = algorithmic_flops(H, W, True) * B
        roofline_scan = {"bound": "fp32", "kernel": "presample_kernel<2> + bucket_scan_kernel + scatter_events_kernel + "
                                                    "sweep2_kernel<59,16,1> (one STED acquisition with in-scan bleaching)",
                         "achieved": fl_scan / (scan_alone_ms * 1e-3) * 1e-12, "peak": pk.value, "unit": "TFLOP/s",
                         "frac": fl_scan / (scan_alone_ms * 1e-3) * 1e-12 / pk.value if pk.value else None,
                         "algorithmic_flops_per_launch": fl_scan, "launch_ms": scan_alone_ms,
                         "in_step": {"sweep2_ms": sweep_ms,
                                     "note": "sweep2_kernel alone inside the timed steps (session compute stream)" + (
                                             "" if args.sequential else "; its death schedule (presample + bucket scan + scatter) "
                                             "was drawn on the session's second stream under the confocal gather before it, and the "
                                             "low-priority session's gather shares the SMs with it")},
                         "note": "3N algorithmic flops (SURVEY 8d); the kernels execute at most 2N (the survival factors are pre-drawn deaths, and the "
                                 "gather visits only the tap rows whose datamap row holds a molecule inside the warp's window - on these "
                                 "maps about 0.6 of them - so the fraction can approach or exceed 1 like the gather's).  launch_ms is the whole chain on one "
                                 "stream with nothing else on the GPU (CUDA events on the launching stream, 5 scans after the timed "
                                 "region, each on freshly loaded maps)"}
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
            "scaling": "strong" if args.workload == "config5" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, B, world), "env_steps_per_s": value / ACQ_PER_STEP,
            "e2e": {"value": e2e_value, "unit": "acquisitions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "frac_of_device_resident": e2e_value / value,
                    "api": "sted_session_* (C ABI, pinned host buffers, uint16 maps and counts): per step the next structure is "
                           "uploaded once and imaged by a " + ("second" if args.sequential else "low-priority") + " session, the main "
                           "session adopts it on the device, images it 3x and returns the bleached map; host sync every step",
                    "numa_bound": numa_bound, "rewards_gathered_over_nccl": gathered,
                    "host_ceiling": {"d2h_gbs_per_rank": d2h_gbs, "h2d_gbs_per_rank": h2d_gbs, "ranks_copying": world,
                                     "d2h_bound_value": ACQ_PER_STEP * B * world / (d2h / (d2h_gbs * 1e9)),
                                     "note": "pinned-memory copy rate of the slowest rank with all ranks copying at once and no "
                                             "kernels running (measured right after the e2e leg); d2h_bound_value = acquisitions/s "
                                             "if a step cost only its device->host bytes at that rate"}},
            "gpu_launches": n_launch, "clocks": clocks, "roofline": roofline, "roofline_scan": roofline_scan,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# whole env.step (acquisitions + rewards + observation), through the vector environment's public API
# ------------------------------------------------------------------------------------------------
def run_env_step(args, rank, local_rank, world):
    """``VectorContextualMOSTED.step`` for B environments per GPU: every step returns its observation, rewards and
    objectives as numpy arrays on the host, so the device-resident and the end-to-end number are the same number."""
    import torch
    import torch.distributed as dist

    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        dist.barrier()
    from gym_sted_b200 import envs
    from gym_sted_b200 import dist as sted_dist
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    B = args.envs
    env = envs.make_vec("ContextualMOSTED-easy-hslb-v0", B, image_size=args.size, device=dev, seed=1234 + rank,
                        env_base=rank * B, max_episode_steps=args.steps + args.warmup + 2, zero_copy_observation=True)
    rng = np.random.default_rng(rank)
    env.reset(seed=1234 + rank)
    sampler = ClockSampler(local_rank)
    sampler.start()

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    acts = rng.uniform(env.action_space.low, env.action_space.high, (args.steps + args.warmup, B, len(env.actions)))
    for i in range(args.warmup):
        # results are held until the next step returns, as in the timed loop: the device tensors of `info` then keep
        # their blocks one step longer and the allocator's pools reach their steady size HERE (dropping the results
        # at once left that growth - cudaMalloc calls, 10-100 ms - to the second timed step)
        (img, vec), reward, done, _, info = env.step(acts[i])
    sync_all()
    if sampler.nvml is not None:
        sampler.samples.clear()
    sampler.lines.clear()
    d2h = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    step_ms = []
    for i in range(args.steps):
        ts = time.perf_counter()
        (img, vec), reward, done, _, info = env.step(acts[args.warmup + i])
        step_ms.append((time.perf_counter() - ts) * 1e3)
        d2h = img.nbytes + reward.nbytes + info["mo_objs"].nbytes
    e1.record()
    sync_all()
    wall = time.perf_counter() - t0
    ms = max(e0.elapsed_time(e1), wall * 1e3)            # the host work of a step is part of the step
    clocks = sampler.stop()
    if rank == 0:
        print("env_step: host ms of every timed step: " + " ".join(f"{m:.1f}" for m in step_ms), file=sys.stderr)
    rewards_all = sted_dist.gather_rewards(torch.as_tensor(reward, dtype=torch.float32, device=dev), world * B) if world > 1 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        steps_s = B * world * args.steps / (ms * 1e-3)
        value = ACQ_PER_STEP * steps_s
        line = {"metric": "STED acquisitions/sec", "value": value, "unit": "acquisitions/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args, B, world, schedule="VectorContextualMOSTED.step: conf, STED+bleach, conf, objectives + "
                                          "Otsu + decorrelation + nanodomain F1, new synapses generated on the device, confocal "
                                          "of the new structure, uint16 observation to the host (returned as a view of the pinned buffer it arrived in)"),
                "env_steps_per_s": steps_s, "step_ms_rank0": {"median": float(np.median(step_ms)), "min": float(np.min(step_ms)),
                                                                "max": float(np.max(step_ms))},
                "e2e": {"value": value, "unit": "acquisitions/s", "h2d_bytes_per_step": int(acts[0].nbytes), "d2h_bytes_per_step": int(d2h),
                        "api": "gym_sted_b200.envs.make_vec(...).step(actions) -> numpy observation / reward / info",
                        "note": "value IS the end-to-end number for this workload: every step ends on the host",
                        "rewards_gathered_over_nccl": None if rewards_all is None else int(rewards_all.numel())},
                # own kernels per step: 3 x (gather + detector), STED (tables, death schedule x3, scan, detector), synapses x2,
                # padding, objectives, apodise / spectrum gather / decorrelation, candidates, Gaussian fits, F1, observation
                "gpu_launches": 23 * args.steps, "clocks": clocks, "roofline": None, "cpu_baseline": None,
                "mean_reward": float(np.mean(reward))}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="acquisitions", choices=["acquisitions", "config5", "env_step"])
    ap.add_argument("--envs", type=int, default=None, help="environments per GPU (default 256; config5: 2048 / N)")
    ap.add_argument("--size", type=int, default=None, help="image side in pixels (default 256; config5: 512; env_step: 64)")
    ap.add_argument("--cpu-envs", type=int, default=None, help="env-steps in the CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--sequential", action="store_true",
                    help="issue the four acquisitions of a step one after the other (no overlap of the next structure's confocal)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.size is None:
        args.size = {"acquisitions": 256, "config5": 512, "env_step": 64}[args.workload]
    if args.envs is None:
        args.envs = 2048 // world if args.workload == "config5" else 256
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload == "env_step":
        run_env_step(args, rank, local_rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
