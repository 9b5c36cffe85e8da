"""ContextualMOSTED environments on the batched CUDA path.

``VectorContextualMOSTED`` steps B independent environments per call with the acquisition schedule of
``ContextualSTEDMultiObjectivesEnv.step`` (``gym_sted/envs/contextual_sted_env.py:49-93`` +
``mosted_env.py:168-246``): conf1 -> STED (+bleach, datamap mutated) -> conf2 -> foregrounds ->
[Resolution, Bleach, SNR] -> nanodomain F1 reward -> observation packing -> fresh datamap + confocal.
``ContextualSTEDMultiObjectivesEnv`` is the single-environment (B = 1) gym-API view of it and ``make``
resolves the reference's ids (``gym_sted/__init__.py:587-610,665-674``); ``gym`` itself is not installed in
this image, so ``Box``/``Tuple`` below are minimal stand-ins for ``gym.spaces``.

Deviations, all because of un-vendored dependencies of the reference: synapses come from the device generator
``synapse.DeviceSynapseGenerator`` (``sted_synapse_generate``; stand-in for ``pysted.exp_data_gen.Synapse``).
``bleach_sampling`` "uniform"/"choice" draw the targets of ``defaults.fluorescence_criterions`` and fit
``(sigma_abs, k1, b)`` with the device-side ``FluorescenceOptimizer`` (``optimizer.py``).

A step queues everything — three acquisitions, Otsu / Bleach / SNR, the decorrelation resolution, nanodomain detection +
Gaussian fits (side stream) + F1 matching, the next structures, their confocal image and the uint16 observation — without
reading anything back, then makes ONE transfer into pinned host memory and waits once.
"""
from __future__ import annotations

import os
import random
import time
from types import SimpleNamespace

import numpy as np
import torch

from . import base, defaults, rewards, synapse
from .batch import BatchedMicroscope

OBJ_NAMES = ["Resolution", "Bleach", "SNR"]


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if np.issubdtype(np.dtype(dtype), np.integer):      # e.g. Box(0, 2**16, dtype=uint16): clip to the dtype
            info = np.iinfo(dtype)
            low, high = np.clip(low, info.min, info.max), np.clip(high, info.min, info.max)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shape if shape is not None else np.shape(low)).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.low.shape).copy()
        self.shape, self.dtype = self.low.shape, np.dtype(dtype)

    def sample(self, rng=None):
        rng = rng or np.random.default_rng()
        return rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


class Tuple(tuple):
    def __new__(cls, spaces):
        return super().__new__(cls, spaces)


class Normalizer:
    """``gym_sted/utils.py:414-440``: (x - low) / (high - low), un-clipped."""

    def __init__(self, names, scales):
        self.low = np.array([scales[n]["low"] for n in names], dtype=np.float64)
        self.span = np.array([scales[n]["high"] - scales[n]["low"] for n in names], dtype=np.float64)

    def __call__(self, x):
        return (np.asarray(x, dtype=np.float64) - self.low) / self.span


def make_microscope(fluo=None, detector=None):
    """``MicroscopeGenerator.generate_microscope`` (``gym_sted/utils.py:151-166``)."""
    return base.Microscope(base.GaussianBeam(**defaults.LASER_EX), base.DonutBeam(**defaults.LASER_STED),
                           base.Detector(**(detector or defaults.DETECTOR)), base.Objective(**defaults.OBJECTIVE),
                           base.Fluorescence(**(fluo or defaults.FLUO)), load_cache=True)


class BleachSampler:
    """``gym_sted/utils.py:271-376`` (modes constant / uniform / choice; see the module docstring)."""

    def __init__(self, mode, value=None, routine=None, seed=None, criterions=None):
        self.rng = random.Random(seed)
        self.mode, self.value = mode, value
        if isinstance(routine, str):
            self.value = defaults.load_routine(routine)
        self.criterions = criterions or defaults.fluorescence_criterions
        if mode not in ("constant", "uniform", "choice"):
            raise ValueError(f"unknown bleach sampling mode {mode!r}")
        self._optimizer, self.criteria = None, None

    def _draw(self, v):
        if isinstance(v, (list, tuple)):
            if self.mode == "choice":
                return float(self.rng.choice(list(np.linspace(*v, 5))))
            return self.rng.uniform(*v)
        return v

    def draw_criteria(self):
        """One ``UniformCriterion`` / ``ChoiceCriterion`` (``utils.py:718-769``): dict order, one draw per range."""
        return {k: {kk: self._draw(vv) for kk, vv in v.items()} for k, v in self.criterions.items()}

    def sample_batch(self, n, microscope=None, samples=None):
        """``n`` fluorophore parameter dicts.  ``uniform`` / ``choice``: ``n`` criteria are drawn and fitted together by
        the device-side ``FluorescenceOptimizer`` (``optimizer.py``), which follows the reference's optimisation path.
        ``samples``: per environment the sample type whose correction factor sizes the test structure
        (``optimizer.set_correction_factor(self.group)``, ``preference_sted_env.py:154``); one launch per distinct factor."""
        return self.sample_batch_collect(self.sample_batch_launch(n, microscope, samples))

    def sample_batch_launch(self, n, microscope=None, samples=None):
        """First half of ``sample_batch``: draws the ``n`` criteria (the sampler's RNG advances here) and QUEUES the fits -
        one launch per distinct factor, each on its own stream (a launch lasts as long as ONE fit, 25 ms, however few
        environments it holds: six of them one after the other were 200 ms of a 256-environment reset).  Nothing is
        waited for: an environment queues the fits of its NEXT episode and steps through the current one meanwhile."""
        if self.mode == "constant":
            return ("constant", n)
        from .optimizer import FluorescenceOptimizer
        if self._optimizer is None:
            self._optimizer = FluorescenceOptimizer(microscope or make_microscope())
        opt = self._optimizer
        criteria = [self.draw_criteria() for _ in range(n)]
        names = ["clusters"] * n if samples is None else list(samples)
        cur = torch.cuda.current_stream(opt.device)
        pending = []
        for g, name in enumerate(sorted(set(names))):
            idx = [i for i, s_ in enumerate(names) if s_ == name]
            opt.set_correction_factor(name)
            side = rewards.side_stream(opt.device, 8 + g)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                out_d, keep = opt.optimize_batch_device([criteria[i] for i in idx])
            pending.append((idx, out_d, keep, side))
        return ("fits", n, criteria, pending)

    def sample_batch_collect(self, handle):
        """Second half of ``sample_batch``: waits for the queued fits and returns the ``n`` parameter dicts; ``criteria``
        then holds the criteria they were fitted to."""
        if handle[0] == "constant":
            return [self.value or defaults.FLUO] * handle[1]
        _, n, criteria, pending = handle
        fits = np.empty((n, 5))
        for idx, out_d, keep, side in pending:
            side.synchronize()
            fits[idx] = out_d.cpu().numpy()
        self.criteria = criteria
        lam_ex = int(round(self._optimizer.microscope.excitation.lambda_ * 1e9))
        out = []
        for sigma, k1, b, _, _ in fits:
            fluo = {k: (dict(v) if isinstance(v, dict) else v) for k, v in defaults.FLUO.items()}
            fluo["sigma_abs"][lam_ex], fluo["k1"], fluo["b"] = float(sigma), float(k1), float(b)
            out.append(fluo)
        return out

    def sample(self, microscope=None):
        return self.sample_batch(1, microscope)[0]


class VectorContextualMOSTED:
    """B ContextualMOSTED environments stepped together on one GPU."""

    obj_names = OBJ_NAMES
    prefetch_structures = True       # step() generates the next structures on a side stream while it images the current ones

    def __init__(self, num_envs=1, bleach_sampling="constant", actions=("p_sted", "p_ex", "pdt"), max_episode_steps=10,
                 scale_nanodomain_reward=1.0, normalize_observations=True, image_size=64, device="cuda", seed=None,
                 env_base=0, compute_f1=True, on_invalid="mask", zero_copy_observation=False):
        """``zero_copy_observation``: ``step`` returns the uint16 image stack as a VIEW of the pinned buffer the device
        wrote it to (two buffers, alternating: the array stays valid until the step after the next one) instead of a fresh
        copy - at 256 x 256 px and 256 environments the copy of the 100 MB stack is half of the step.

        ``on_invalid``: what a ``None`` objective (negative SNR, ``objectives.py:95-96``) does.  The reference's
        ``Normalizer`` raises ``TypeError`` on it; "raise" keeps that (the single-environment views use it), "mask"
        (default for batches) carries NaN in that environment's objective / observation entries and lists the
        environment in ``info["invalid"]`` instead of aborting the other B-1 environments mid-step."""
        if on_invalid not in ("mask", "raise"):
            raise ValueError("on_invalid must be 'mask' or 'raise'")
        self.on_invalid = on_invalid
        self.zero_copy_observation = bool(zero_copy_observation)
        self.B, self.H, self.W = int(num_envs), int(image_size), int(image_size)
        self.actions = list(actions)
        self.max_episode_steps = int(max_episode_steps)
        self.scale_nanodomain_reward = scale_nanodomain_reward
        self.normalize_observations = normalize_observations
        self.compute_f1 = compute_f1
        self.device = torch.device(device)
        self.env_base = int(env_base)
        self.np_rng = np.random.default_rng(seed)
        low = np.array([defaults.action_spaces[n]["low"] for n in self.actions], dtype=np.float32)
        high = np.array([defaults.action_spaces[n]["high"] for n in self.actions], dtype=np.float32)
        self.action_space = Box(low, high, dtype=np.float32)
        vec = (len(OBJ_NAMES) + len(self.actions)) * self.max_episode_steps
        self.observation_space = Tuple((Box(0, 2 ** 16, shape=(self.H, self.W, 3), dtype=np.uint16),
                                        Box(0, 1024, shape=(vec,), dtype=np.float32)))
        self.action_normalizer = Normalizer(self.actions, defaults.action_spaces)
        self.obj_normalizer = Normalizer(OBJ_NAMES, defaults.scales_dict)
        self.microscope = make_microscope()
        if isinstance(bleach_sampling, dict):
            self.bleach_sampler = BleachSampler(**bleach_sampling)
        else:
            self.bleach_sampler = BleachSampler(mode=bleach_sampling)
        if seed is not None:
            self.bleach_sampler.rng.seed(int(self.np_rng.integers(0, 2 ** 31 - 1)))
        self.bm = BatchedMicroscope(self.microscope, self.B, self.H, self.W, defaults.PIXELSIZE, device=self.device,
                                    env_base=self.env_base, seed=int(self.np_rng.integers(0, 2 ** 31 - 1)))
        self.conf = (defaults.PDT, defaults.P_EX, 0.0)            # generate_params(): pdt, p_ex, p_sted
        self.current_step = 0
        self.frames = None
        self.state = None
        self.episode_memory = {"actions": [], "mo_objs": [], "reward": []}
        self.synapse_gen = synapse.DeviceSynapseGenerator(self.B, self.H, self.W, self.device, env_base=self.env_base,
                                                          seed=int(self.np_rng.integers(0, 2 ** 62)))
        self._truths = None          # rewards.DeviceTruths of the structures being imaged
        self._prefetched = None      # ((frames, coords, counts), event) of structures generated ahead on a side stream
        self._fluo_pending = None    # BleachSampler.sample_batch_launch handle: the next episode's fluorophore fits
        self.prefetch_fluorophores = os.environ.get("STED_B200_PREFETCH_FLUO", "1") != "0"
        self._coords_host = None     # the same coordinates on the host (list of (n_i, 2) int64 arrays), read back lazily
        self._pin = {}               # pinned host buffers of the end-of-step transfer
        self._conf_d = None
        self._obs_vec = None
        self.step_timing = bool(os.environ.get("STED_B200_STEP_TIMING"))   # info["timing"]: host seconds per phase of step()
        # Two schedules of a step (same kernels, same results).  "wide": acquisitions on a high-priority stream, three
        # side streams (next structures / candidates + fits / decorrelation) - pays when the kernels are long enough to
        # fill the GPU (16.2 instead of 19.2 ms per 256-env step at 256 x 256).  "narrow": the acquisitions and the
        # decorrelation chain on the caller's stream, the fits on one side stream - at 64 x 64 the step is bound by the
        # host and the extra stream switches and event waits of the wide schedule cost 0.3 ms of 2.3.
        layout = os.environ.get("STED_B200_STEP_LAYOUT", "auto")
        if layout not in ("auto", "wide", "narrow"):
            raise ValueError("STED_B200_STEP_LAYOUT must be auto, wide or narrow")
        self.wide_step = layout == "wide" or (layout == "auto" and self.B * self.H * self.W >= (1 << 22))
        self._step_stream = torch.cuda.Stream(self.device, priority=-2) if self.wide_step else None

    # ------------------------------------------------------------------ helpers
    def _pinned(self, name, shape, dtype):
        buf = self._pin.get(name)
        if buf is None or tuple(buf.shape) != tuple(shape):
            buf = self._pin[name] = torch.empty(tuple(shape), dtype=dtype, pin_memory=True)
        return buf

    def _img_slot(self):
        """Name of the pinned buffer this step's observation image goes to (alternating when views are handed out)."""
        return f"img{self.current_step & 1}" if self.zero_copy_observation else "img"

    def _conf_dev(self):
        """The confocal imaging parameters (pdt, p_ex, p_sted) as (B,) float64 device tensors, uploaded once."""
        if self._conf_d is None:
            c = torch.tensor(self.conf, dtype=torch.float64).view(3, 1).expand(3, self.B).contiguous().to(self.device)
            self._conf_d = (c[0], c[1], c[2])
        return self._conf_d

    @property
    def synapses(self):
        """One ``Synapse`` container per environment with the nanodomain coordinates of the structure being imaged."""
        if self._coords_host is None:
            if self._truths is None:
                return None
            T, n = self._truths.coords.cpu().numpy(), self._truths.counts.cpu().numpy()
            self._coords_host = [T[b, :n[b]].astype(np.int64) for b in range(self.B)]
        return [synapse.Synapse(frame=np.zeros((0, 0)), coords=c) for c in self._coords_host]

    def _new_datamaps(self):
        """``_update_datamap`` (mosted_env.py:168-187): new synapse, new datamap, confocal image of it — all queued on
        the device (``sted_synapse_generate`` -> ``sted_pad_maps`` -> acquisition), nothing read back."""
        if self._prefetched is not None:                       # generated on a side stream since the start of the step
            (frames, coords, counts), ready = self._prefetched
            self._prefetched = None
            torch.cuda.current_stream(self.device).wait_event(ready)
        else:
            frames, coords, counts = self.synapse_gen.generate()
        self._truths = rewards.DeviceTruths.from_device(coords, counts)   # ground truth of the device F1 matching
        self._coords_host = None
        self.frames = frames                                   # molecule maps of the current synapses (device, int32)
        self.bm.set_datamaps(frames, max_molecules=self.synapse_gen.max_molecules)
        state, _ = self.bm.get_signal_and_bleach(*self._conf_dev(), bleach=False, tables_key="conf")
        return state

    def _prefetch_structures(self):
        """Queue the generation of the NEXT structures (``sted_synapse_generate``: independent of everything the step
        images) on a side stream, so that it runs under the step's first acquisitions instead of after the last one;
        ``_new_datamaps`` waits for it."""
        cur = torch.cuda.current_stream(self.device)
        side = rewards.side_stream(self.device, 2)
        side.wait_stream(cur)                                  # the generator's scratch is shared with earlier calls
        with torch.cuda.stream(side):
            out = self.synapse_gen.generate()
            ready = torch.cuda.Event()
            ready.record(side)
        for t in out:
            t.record_stream(cur)                               # consumed by the step's stream after `ready`
        self._prefetched = (out, ready)

    def _pack_observation(self, state, conf1=None, sted=None):
        """(state, conf1, sted) -> uint16 (B,H,W,3) device tensor (``sted_pack_observation``); missing channels are 0."""
        zeros = torch.zeros_like(state) if conf1 is None or sted is None else None
        conf1 = zeros if conf1 is None else conf1
        sted = zeros if sted is None else sted
        out = torch.empty((self.B, self.H, self.W, 3), dtype=torch.uint16, device=state.device)
        from . import _lib
        _lib.check(self.bm.lib.sted_pack_observation(state.data_ptr(), conf1.data_ptr(), sted.data_ptr(), self.B, self.H, self.W,
                                                     out.data_ptr(), torch.cuda.current_stream(state.device).cuda_stream))
        return out

    def _full_action(self, action):
        cols = {n: action[:, self.actions.index(n)] if n in self.actions else
                np.full(self.B, getattr(defaults, n.upper())) for n in ("pdt", "p_ex", "p_sted")}
        return cols["pdt"].astype(np.float64), cols["p_ex"].astype(np.float64), cols["p_sted"].astype(np.float64)

    # ------------------------------------------------------------------ gym API
    def step(self, action):
        """``_step_impl`` queued on the environment's own HIGH-priority stream.  The step's acquisitions then win the SMs
        from what runs beside them on the (default-priority) side streams - Gaussian fits, decorrelation, the next
        structures: long grids of light CTAs that, at equal priority, are dispatched in launch order and keep a gather
        launched after them waiting until their last CTA has started (`scripts/env_step_kernels.py`, PROBE_TIMELINE)."""
        st = self._step_stream
        if st is None:
            return self._step_impl(action)
        cur = torch.cuda.current_stream(self.device)
        st.wait_stream(cur)                                    # after whatever the caller queued (reset, its own kernels)
        with torch.cuda.stream(st):
            out = self._step_impl(action)                      # ends with a synchronize of `st`
        for v in out[-1].values():
            if torch.is_tensor(v) and v.is_cuda:
                v.record_stream(cur)                           # allocated on `st`, used by the caller on its stream
        return out

    def reset(self, seed=None, options=None):
        if seed is not None:
            self.np_rng = np.random.default_rng(seed)
            self.bleach_sampler.rng.seed(int(self.np_rng.integers(0, 2 ** 31 - 1)))
            self.synapse_gen.seed, self.synapse_gen.offset = int(self.np_rng.integers(0, 2 ** 62)), 0
        fluos = None
        if self.bleach_sampler.mode != "constant":
            # the fits of THIS episode were queued at the previous reset (they ran under that episode's steps) unless
            # this reset reseeds the sampler; the criteria sequence is the one of fitting at reset time either way
            handle, self._fluo_pending = self._fluo_pending, None
            if handle is None or seed is not None:
                handle = self.bleach_sampler.sample_batch_launch(self.B, self.microscope)
            fluos = [base.Fluorescence(**f) for f in self.bleach_sampler.sample_batch_collect(handle)]
            if self.prefetch_fluorophores:
                self._fluo_pending = self.bleach_sampler.sample_batch_launch(self.B, self.microscope)
        if fluos is None:
            self.microscope = make_microscope(self.bleach_sampler.sample())
            self.bm.microscope = self.microscope
        self.bm.set_fluorophores(fluos)
        self.current_step = 0
        self.episode_memory = {"actions": [], "mo_objs": [], "reward": []}
        self._obs_vec = np.zeros((self.B, self.observation_space[1].shape[0]), dtype=np.float32)
        state = self._new_datamaps()
        zeros = torch.zeros_like(state)
        self.state = (state, zeros, zeros)
        img = self._pack_observation(state).cpu().numpy()
        return (img, self._obs_vec.copy()), {}

    def _step_impl(self, action):
        action = np.clip(np.asarray(action, dtype=np.float32).reshape(self.B, len(self.actions)),
                         self.action_space.low, self.action_space.high)
        if self._obs_vec is None:
            raise RuntimeError("step() called before reset()")
        per_step = len(self.actions) + len(OBJ_NAMES)
        if (self.current_step + 1) * per_step > self._obs_vec.shape[1]:
            raise ValueError("episode longer than max_episode_steps: the observation vector is full (call reset())")
        marks = [("start", time.perf_counter())] if self.step_timing else None
        pdt, p_ex, p_sted = self._full_action(action)
        bm, dev, B = self.bm, self.device, self.B
        h_act = self._pinned("act", (3, B), torch.float64)
        h_act.numpy()[:] = (pdt, p_ex, p_sted)
        d_act = h_act.to(dev, non_blocking=True)
        conf = self._conf_dev()
        # ---- everything below only queues work.  Three side streams run beside the acquisitions: the next structures
        # (from the start), and - as soon as the STED image exists - candidates + Gaussian fits and the decorrelation
        # chain, both of which need nothing but that image and so overlap the second confocal and what follows it.
        wide = self.wide_step
        if wide and self.prefetch_structures:
            self._prefetch_structures()
        conf1, _ = bm.get_signal_and_bleach(*conf, bleach=False, tables_key="conf")
        sted, _ = bm.get_signal_and_bleach(d_act[0], d_act[1], d_act[2], bleach=True)
        if wide:
            pending = rewards.detect_nanodomains_launch(sted) if self.compute_f1 else None
            res_d, res_done = rewards.resolution_launch(sted, lane=1)
        conf2, _ = bm.get_signal_and_bleach(*conf, bleach=False, tables_key="conf")
        bleached = bm.datamaps_roi().clone()
        if not wide:
            pending = rewards.detect_nanodomains_launch(sted) if self.compute_f1 else None
        fg_s, fg_c, objs_d, oflags_d = rewards.foreground_objectives_device(conf1, sted, conf2)
        if not wide:
            res_d, res_done = rewards.resolution_device(sted), None
        truths, coords = self._truths, self._coords_host
        exhausted = self._exhausted()
        state = self._next_state()
        obs_img = self._pack_observation(state, conf1, sted)
        # ---- one transfer into pinned memory, one wait (the image first: it does not wait for the side streams)
        img_slot = self._img_slot()
        host = {img_slot: self._pinned(img_slot, obs_img.shape, obs_img.dtype)}
        host[img_slot].copy_(obs_img, non_blocking=True)
        if res_done is not None:
            torch.cuda.current_stream(dev).wait_event(res_done)
        pulls = [("res", res_d), ("objs", objs_d), ("oflags", oflags_d)]
        if self.compute_f1:
            f1_d, cnt_d, dfl_d, mfl_d = rewards.nanodomain_f1_launch(sted, truths, handle=pending)
            pulls += [("f1", f1_d), ("cnt", cnt_d), ("dfl", dfl_d), ("mfl", mfl_d)]
        if coords is None:
            pulls += [("tc", truths.coords), ("tn", truths.counts)]
        if self._truths is not truths:                       # the next structures' ground truth rides along
            pulls += [("nc", self._truths.coords), ("nn", self._truths.counts)]
        for name, t in pulls:
            host[name] = self._pinned(name, t.shape, t.dtype)
            host[name].copy_(t, non_blocking=True)
        if marks:
            marks.append(("queued", time.perf_counter()))
        torch.cuda.current_stream(dev).synchronize()
        if marks:
            marks.append(("device_done", time.perf_counter()))
        host = {k: v.numpy() for k, v in host.items()}
        if coords is None:
            tc, tn = host["tc"].astype(np.int64), host["tn"]      # one conversion, then per-environment views of the copy
            coords = [tc[b, :tn[b]] for b in range(B)]
        if self._truths is not truths:
            nc, nn = host["nc"].astype(np.int64), host["nn"].copy()   # the pinned buffers are rewritten by the next step
            self._coords_host = [nc[b, :nn[b]] for b in range(B)]
        else:
            self._coords_host = coords
        snr = np.where((host["oflags"] & 1) != 0, np.nan, host["objs"][:, 1])
        mo_objs = np.stack([host["res"], host["objs"][:, 0], snr], axis=1)      # (B,3)
        if self.compute_f1:
            rewards.check_f1_flags(host["dfl"], host["mfl"])
            f1 = host["f1"].copy()
        else:
            f1 = np.zeros(B)

        done = np.full(B, self.current_step >= self.max_episode_steps - 1) | exhausted
        self.episode_memory["mo_objs"].append(mo_objs)
        self.episode_memory["actions"].append(action)

        invalid = np.isnan(mo_objs).any(axis=1)
        if invalid.any() and self.on_invalid == "raise" and self.normalize_observations:
            # the reference's Normalizer raises TypeError on a None objective (SURVEY App. C "edge cases")
            raise TypeError("unsupported operand type(s) for -: 'NoneType' and 'int' (SNR objective is None)")
        # per past step [action || objectives] (contextual_sted_env.py:84-93): this step's slice of the running vector
        at = self.current_step * per_step
        self._obs_vec[:, at:at + len(self.actions)] = self.action_normalizer(action) if self.normalize_observations else action
        self._obs_vec[:, at + len(self.actions):at + per_step] = self.obj_normalizer(mo_objs) if self.normalize_observations else mo_objs
        self.current_step += 1
        self.state = (state, conf1, sted)
        img = host[img_slot] if self.zero_copy_observation else host[img_slot].copy()

        reward = f1 * self.scale_nanodomain_reward
        self.episode_memory["reward"].append(reward)
        info = {"action": action, "bleached": bleached, "sted_image": sted, "conf1": conf1, "conf2": conf2,
                "fg_c": fg_c.view(torch.bool), "fg_s": fg_s.view(torch.bool), "mo_objs": mo_objs, "reward": reward,
                "f1-score": f1, "nanodomains-coords": coords, "invalid": invalid}
        if marks:                                            # host seconds: queueing, waiting for the device, unpacking
            marks.append(("end", time.perf_counter()))
            info["timing"] = {b[0]: b[1] - a[1] for a, b in zip(marks, marks[1:])}
        return (img, self._obs_vec.copy()), reward, done, np.zeros(B, dtype=bool), info

    def _next_state(self):
        """Contextual bandit: every step ends on a fresh synapse (contextual_sted_env.py:90)."""
        return self._new_datamaps()

    def _exhausted(self):
        return np.zeros(self.B, dtype=bool)

    def close(self):
        return None


class VectorSequenceMOSTED(VectorContextualMOSTED):
    """B ``SequenceSTEDMultiObjectivesEnv`` environments (``gym_sted/envs/sequence_sted_env.py:19-140``): the same
    structure is imaged again at every step, so the datamap carries its photobleaching through the episode
    (the boundary's in-place update, ``:84-90``), and an environment is done early once its datamap is empty
    (``:109``)."""

    prefetch_structures = False      # no new structure inside an episode

    def _next_state(self):
        state, _ = self.bm.get_signal_and_bleach(*self._conf_dev(), bleach=False, tables_key="conf")
        return state

    def _exhausted(self):
        return (self.bm.datamaps_roi().amax(dim=(1, 2)) == 0).cpu().numpy()


class ContextualSTEDMultiObjectivesEnv:
    """Single-environment gym-API view (reset -> (obs, info); step -> (obs, reward, done, False, info))."""

    metadata = {"render.modes": ["human"]}
    obj_names = OBJ_NAMES

    vector_cls = VectorContextualMOSTED

    def __init__(self, bleach_sampling="constant", actions=("p_sted",), max_episode_steps=10, scale_nanodomain_reward=1.0,
                 normalize_observations=True, **kwargs):
        kwargs.setdefault("on_invalid", "raise")             # one environment: fail like the reference does
        self.vec = self.vector_cls(1, bleach_sampling=bleach_sampling, actions=actions, max_episode_steps=max_episode_steps,
                                   scale_nanodomain_reward=scale_nanodomain_reward,
                                   normalize_observations=normalize_observations, **kwargs)
        self.action_space, self.observation_space = self.vec.action_space, self.vec.observation_space
        self.actions = self.vec.actions
        self.spec = SimpleNamespace(id=None, max_episode_steps=max_episode_steps)

    @property
    def microscope(self):
        return self.vec.microscope

    def reset(self, seed=None, options=None):
        (img, vec), info = self.vec.reset(seed=seed, options=options)
        return (img[0], vec[0]), info

    def step(self, action):
        self.vec.max_episode_steps = self.spec.max_episode_steps
        (img, vec), reward, done, trunc, info = self.vec.step(np.asarray(action)[None])
        one = {k: (v[0] if isinstance(v, (np.ndarray, list)) else v[0].cpu().numpy()) for k, v in info.items()
               if k != "invalid"}                            # the reference's info keys only (contextual_sted_env.py:69-81)
        return (img[0], vec[0]), float(reward[0]), bool(done[0]), False, one

    def close(self):
        return None


class SequenceSTEDMultiObjectivesEnv(ContextualSTEDMultiObjectivesEnv):
    """Single-environment view of ``VectorSequenceMOSTED``."""

    vector_cls = VectorSequenceMOSTED


class SyntheticDatamapGenerator:
    """Stand-in for ``gym_sted/datamaps/datamap.py`` (``DatamapGenerator``): the reference crops 96x96 windows
    out of U-Net-predicted structure maps under ``./data/datamap/<group>`` which are not shipped
    (``defaults.py:6-13``).  Here every group gets a procedural structure map in [0, 1] — filaments for
    actin / tubulin / lifeact, clusters for psd95 / camkii / PSD95-Bassoon — scaled by ``int(N(40, 4))``
    molecules (``datamap.py:68-85``) and augmented with the same flips / rot90."""

    GROUPS = ["actin", "psd95", "lifeact", "PSD95-Bassoon", "tubulin", "camkii"]

    def __init__(self, shape=(96, 96), molecules=40, molecules_scale=0.1, seed=None):
        self.shape, self.molecules, self.molecules_scale = shape, molecules, molecules_scale
        self.rng = np.random.default_rng(seed)
        self._pin = None

    def sample_group(self):
        return str(self.rng.choice(self.GROUPS))

    def __call__(self, group):
        H, W = self.shape
        yy, xx = np.mgrid[:H, :W].astype(np.float64)
        img = np.zeros((H, W))
        if group in ("actin", "tubulin", "lifeact"):
            width = {"actin": 1.2, "tubulin": 1.0, "lifeact": 1.6}[group]
            for _ in range(int(self.rng.integers(3, 8))):
                th = self.rng.uniform(0, np.pi)
                off = self.rng.uniform(0.1, 0.9) * (H * abs(np.sin(th)) + W * abs(np.cos(th)))
                d = np.abs(xx * np.cos(th) + yy * np.sin(th) - off + 3 * np.sin(0.08 * (yy * np.cos(th) - xx * np.sin(th))))
                img = np.maximum(img, np.exp(-0.5 * (d / width) ** 2))
        else:
            for _ in range(int(self.rng.integers(8, 25))):
                cy, cx = self.rng.uniform(4, H - 4), self.rng.uniform(4, W - 4)
                sig = self.rng.uniform(1.0, 2.5)
                img = np.maximum(img, np.exp(-0.5 * ((yy - cy) ** 2 + (xx - cx) ** 2) / sig ** 2))
        img[img < 0.5] = 0.0
        scale = self.rng.normal(self.molecules, self.molecules_scale * self.molecules)
        while scale < 1 or scale > self.molecules:
            scale = self.rng.normal(self.molecules, self.molecules_scale * self.molecules)
        img = np.floor(img * int(scale))
        if self.rng.random() < 0.5:
            img = np.rot90(img, int(self.rng.integers(1, 4)))
        if self.rng.random() < 0.5:
            img = np.flip(img, axis=0)
        if self.rng.random() < 0.5:
            img = np.flip(img, axis=1)
        return np.ascontiguousarray(img)

    NOBJ = 24

    def draw_batch(self, groups):
        """The random parameters of one structure map per entry of ``groups`` (host RNG, a few numbers per map) in the
        layout ``sted_structure_maps`` takes: ``objects`` float64 (B, 24, 4), ``env`` float64 (B, 6)."""
        H, W = self.shape
        B = len(groups)
        rng = self.rng
        fil = np.array([g in ("actin", "tubulin", "lifeact") for g in groups])
        width = np.array([{"actin": 1.2, "tubulin": 1.0, "lifeact": 1.6}.get(g, 1.0) for g in groups])
        NF, NC = 7, self.NOBJ
        n_obj = np.where(fil, rng.integers(3, 8, B), rng.integers(8, 25, B))
        th = rng.uniform(0, np.pi, (B, NF))
        off = rng.uniform(0.1, 0.9, (B, NF)) * (H * np.abs(np.sin(th)) + W * np.abs(np.cos(th)))
        cy, cx = rng.uniform(4, H - 4, (B, NC)), rng.uniform(4, W - 4, (B, NC))
        sig = rng.uniform(1.0, 2.5, (B, NC))
        scale = rng.normal(self.molecules, self.molecules_scale * self.molecules, B)
        bad = (scale < 1) | (scale > self.molecules)
        while bad.any():
            scale[bad] = rng.normal(self.molecules, self.molecules_scale * self.molecules, int(bad.sum()))
            bad = (scale < 1) | (scale > self.molecules)
        rot = np.where(rng.random(B) < 0.5, rng.integers(1, 4, B), 0) if H == W else np.zeros(B, dtype=np.int64)
        flip0, flip1 = rng.random(B) < 0.5, rng.random(B) < 0.5
        objects = np.zeros((B, NC, 4))
        objects[~fil, :, 0], objects[~fil, :, 1], objects[~fil, :, 2] = cy[~fil], cx[~fil], sig[~fil]
        objects[fil, :NF, 0], objects[fil, :NF, 1], objects[fil, :NF, 2] = np.cos(th[fil]), np.sin(th[fil]), off[fil]
        objects[fil, :NF, 3] = width[fil, None]
        env = np.stack([n_obj, np.floor(scale), rot, flip0, flip1, fil], axis=1).astype(np.float64)
        return objects, env

    def render_numpy(self, objects, env):
        """Float64 numpy rendering of ``draw_batch`` parameters: the model ``sted_structure_maps`` evaluates (checker of
        the kernel, and the renderer when no GPU is asked for)."""
        H, W = self.shape
        yy, xx = np.mgrid[:H, :W].astype(np.float64)
        out = np.zeros((len(env), H, W), dtype=np.float32)
        for b, (e, ob) in enumerate(zip(env, objects)):
            img = np.zeros((H, W))
            for a0, a1, a2, a3 in ob[:int(e[0])]:
                if e[5]:
                    d = np.abs(xx * a0 + yy * a1 - a2 + 3.0 * np.sin(0.08 * (yy * a0 - xx * a1)))
                    v = np.exp(-0.5 * ((d / a3) * (d / a3)))
                else:
                    v = np.exp(-0.5 * ((yy - a0) * (yy - a0) + (xx - a1) * (xx - a1)) / (a2 * a2))
                img = np.maximum(img, v)
            img[img < 0.5] = 0.0
            img = np.floor(img * e[1])
            img = np.rot90(img, int(e[2]))
            if e[3]:
                img = np.flip(img, axis=0)
            if e[4]:
                img = np.flip(img, axis=1)
            out[b] = img
        return out

    def batch(self, groups, device) -> torch.Tensor:
        """One structure map per entry of ``groups`` as a (B,H,W) float32 tensor on ``device``: the same procedural model
        as ``__call__`` with the random parameters drawn on the host and the images rendered by ``sted_structure_maps``
        (one thread per pixel, float64) on a CUDA device, by ``render_numpy`` otherwise."""
        objects, env = self.draw_batch(groups)
        device = torch.device(device)
        if device.type != "cuda":
            return torch.from_numpy(self.render_numpy(objects, env)).to(device)
        from . import _lib
        lib = _lib.load()
        H, W = self.shape
        B = len(groups)
        no = B * self.NOBJ * 4
        if self._pin is None or self._pin.numel() != no + B * 6:   # one pinned block: objects, then env
            self._pin = torch.empty(no + B * 6, dtype=torch.float64, pin_memory=True)
        self._pin.numpy()[:no] = objects.ravel()
        self._pin.numpy()[no:] = env.ravel()
        d_par = self._pin.to(device, non_blocking=True)        # the caller's step waits for its stream before the next draw
        d_obj, d_env = d_par[:no], d_par[no:]
        out = torch.empty((B, H, W), dtype=torch.float32, device=device)
        _lib.check(lib.sted_structure_maps(d_obj.data_ptr(), d_env.data_ptr(), len(groups), H, W, out.data_ptr(),
                                           torch.cuda.current_stream(device).cuda_stream))
        return out


FACTORS = {"synthetic": 1.0, "clusters": 2.25, "camkii": 2.25, "psd95": 2.25, "PSD95-Bassoon": 2.25, "actin": 3.0,
           "tubulin": 3.75, "lifeact": 3.75}          # gym_sted/utils.py:447-456


class VectorPreferenceCountRateMOSTED(VectorContextualMOSTED):
    """B ``PreferenceCountRateScaleSTEDMultiObjectivesEnv`` environments
    (``gym_sted/envs/preference_sted_env.py:205-323``): 96x96 structure-map crops, confocal power lowered by
    x0.75 until the count rate is <= 20 MHz, reward = raw PrefNet score of [Resolution, Bleach, SNR], replaced
    by -10 when the STED image exceeds the count-rate limit; the observation vector starts with
    ``conf_params["p_ex"] / P_EX``."""

    def __init__(self, num_envs=1, bleach_sampling="constant", actions=("p_sted", "p_ex", "pdt"), max_episode_steps=30,
                 max_count_rate=20e6, negative_reward=-10.0, image_size=96, **kw):
        kw.pop("scale_nanodomain_reward", None)
        super().__init__(num_envs, bleach_sampling, actions, max_episode_steps, image_size=image_size, compute_f1=False, **kw)
        vec = 1 + (len(OBJ_NAMES) + len(self.actions)) * self.max_episode_steps
        self.observation_space = Tuple((Box(0, 2 ** 16, shape=(self.H, self.W, 3), dtype=np.uint16),
                                        Box(0, 1, shape=(vec,), dtype=np.float32)))
        self.max_count_rate, self.negative_reward = max_count_rate, negative_reward
        self.datamap_generator = SyntheticDatamapGenerator((self.H, self.W), seed=int(self.np_rng.integers(0, 2 ** 31)))
        self.articulator = rewards.PreferenceArticulator(device="cpu")
        self.groups = None
        self.conf_p_ex = None

    def reset(self, seed=None, options=None):
        if seed is not None:
            self.np_rng = np.random.default_rng(seed)
            self.bleach_sampler.rng.seed(int(self.np_rng.integers(0, 2 ** 31 - 1)))
            self.datamap_generator.rng = np.random.default_rng(seed + 1)
        self.groups = [self.datamap_generator.sample_group() for _ in range(self.B)]
        self.conf_p_ex = None
        if self.bleach_sampler.mode != "constant":
            fluos = [base.Fluorescence(**f) for f in self.bleach_sampler.sample_batch(self.B, self.microscope, samples=self.groups)]
            self.bm.set_fluorophores(fluos)
        else:
            self.microscope = make_microscope(self.bleach_sampler.sample())
            self.bm.microscope = self.microscope
            self.bm.set_fluorophores(None)
        self.current_step = 0
        self.episode_memory = {"actions": [], "mo_objs": [], "reward": []}
        state = self._new_datamaps()
        zeros = torch.zeros_like(state)
        self.state = (state, zeros, zeros)
        img = self._pack_observation(state).cpu().numpy()
        return (img, np.zeros((self.B, self.observation_space[1].shape[0]), dtype=np.float32)), {}

    def _new_datamaps(self):
        """``_update_datamap`` (preference_sted_env.py:173-203) with the per-env count-rate loop (first structures of an
        episode only; afterwards the state image is queued with the kept tables of the lowered confocal powers)."""
        frames = self.datamap_generator.batch(self.groups, self.device)
        self.bm.set_datamaps(frames, max_molecules=self.B * self.H * self.W * self.datamap_generator.molecules)
        pdt = defaults.PDT
        if self.conf_p_ex is None:
            p_ex = np.full(self.B, defaults.P_EX)
            for _ in range(64):
                state, _ = self.bm.get_signal_and_bleach(pdt, p_ex, 0.0, bleach=False)
                hot = (state.amax(dim=(1, 2)) / pdt > self.max_count_rate).cpu().numpy()
                if not hot.any():
                    break
                p_ex = np.where(hot, p_ex * 0.75, p_ex)
            self.conf_p_ex = p_ex
            c = torch.tensor(np.stack([np.full(self.B, pdt), p_ex, np.zeros(self.B)]), dtype=torch.float64).to(self.device)
            self._state_conf_d = (c[0], c[1], c[2])
        state, _ = self.bm.get_signal_and_bleach(*self._state_conf_d, bleach=False, tables_key="state")
        return state

    def _step_impl(self, action):
        """One step of B environments (preference_sted_env.py:262-323).  Like the contextual step it only QUEUES device
        work - three acquisitions with kept confocal tables, objectives, the decorrelation chain, the peak count of the
        STED image, the next structure maps and their state image, the uint16 observation - then makes one transfer into
        pinned memory and waits once; the PrefNet score of the (B, 3) objectives is a 3-10-10-1 MLP on the host."""
        action = np.clip(np.asarray(action, dtype=np.float32).reshape(self.B, len(self.actions)),
                         self.action_space.low, self.action_space.high)
        marks = [("start", time.perf_counter())] if self.step_timing else None
        pdt, p_ex, p_sted = self._full_action(action)
        bm, dev, B = self.bm, self.device, self.B
        h_act = self._pinned("act", (3, B), torch.float64)
        h_act.numpy()[:] = (pdt, p_ex, p_sted)
        d_act = h_act.to(dev, non_blocking=True)
        conf = self._conf_dev()                               # _acquire uses the DEFAULT confocal params
        conf1, _ = bm.get_signal_and_bleach(*conf, bleach=False, tables_key="conf")
        sted, _ = bm.get_signal_and_bleach(d_act[0], d_act[1], d_act[2], bleach=True)
        res_d, res_done = rewards.resolution_launch(sted)     # side stream: beside conf2, objectives, the next structures
        conf2, _ = bm.get_signal_and_bleach(*conf, bleach=False, tables_key="conf")
        bleached = bm.datamaps_roi().clone()
        fg_s, fg_c, objs_d, oflags_d = rewards.foreground_objectives_device(conf1, sted, conf2)
        peak_d = sted.amax(dim=(1, 2))
        exhausted = self._exhausted()
        state = self._new_datamaps()
        obs_img = self._pack_observation(state, conf1, sted)
        torch.cuda.current_stream(dev).wait_event(res_done)
        host = {}
        img_slot = self._img_slot()
        for name, t in (("res", res_d), ("objs", objs_d), ("oflags", oflags_d), ("peak", peak_d), (img_slot, obs_img)):
            host[name] = self._pinned(name, t.shape, t.dtype)
            host[name].copy_(t, non_blocking=True)
        if marks:
            marks.append(("queued", time.perf_counter()))
        torch.cuda.current_stream(dev).synchronize()
        if marks:
            marks.append(("device_done", time.perf_counter()))
        host = {k: v.numpy() for k, v in host.items()}
        snr = np.where((host["oflags"] & 1) != 0, np.nan, host["objs"][:, 1])
        mo_objs = np.stack([host["res"], host["objs"][:, 0], snr], axis=1)
        invalid = np.isnan(mo_objs).any(axis=1)
        if invalid.any() and self.on_invalid == "raise":
            raise TypeError("SNR objective is None (negative signal ratio): the reference fails here as well")
        scores, _, _ = self.articulator.articulate(np.nan_to_num(mo_objs), use_sigmoid=False)
        reward = np.where(invalid, np.nan, scores.ravel().astype(np.float64))
        count_rate = host["peak"].astype(np.float64) / pdt
        reward = np.where(count_rate > self.max_count_rate, self.negative_reward, reward)

        done = np.full(B, self.current_step >= self.max_episode_steps - 1) | exhausted
        self.current_step += 1
        self.episode_memory["mo_objs"].append(mo_objs)
        self.episode_memory["actions"].append(action)
        self.episode_memory["reward"].append(reward)
        info = {"action": action, "bleached": bleached, "sted_image": sted, "conf1": conf1, "conf2": conf2,
                "fg_c": fg_c.view(torch.bool), "fg_s": fg_s.view(torch.bool), "mo_objs": mo_objs, "reward": reward,
                "sample": list(self.groups),
                "conf_params": [{"pdt": defaults.PDT, "p_ex": float(p), "p_sted": 0.0} for p in self.conf_p_ex],
                "invalid": invalid}
        obs = [(self.conf_p_ex / defaults.P_EX)[:, None]]
        for a, mo in zip(self.episode_memory["actions"], self.episode_memory["mo_objs"]):
            obs.append(self.action_normalizer(a) if self.normalize_observations else a)
            obs.append(self.obj_normalizer(mo) if self.normalize_observations else mo)
        obs = np.concatenate(obs, axis=1)
        obs = np.pad(obs, ((0, 0), (0, self.observation_space[1].shape[0] - obs.shape[1])))
        self.state = (state, conf1, sted)
        img = host[img_slot] if self.zero_copy_observation else host[img_slot].copy()
        if marks:                                            # host seconds: queueing, waiting for the device, unpacking
            marks.append(("end", time.perf_counter()))
            info["timing"] = {b[0]: b[1] - a[1] for a, b in zip(marks, marks[1:])}
        return (img, obs.astype(np.float32)), reward, done, np.zeros(B, dtype=bool), info


# ids and kwargs of gym_sted/__init__.py:587-610,665-674 (Appendix C)
REGISTRY = {
    "ContextualMOSTED-easy-v0": dict(max_episode_steps=10, kwargs=dict(
        actions=["p_sted", "p_ex", "pdt"], bleach_sampling="constant", scale_nanodomain_reward=1.0)),
    "ContextualMOSTED-easy-hslb-v0": dict(max_episode_steps=10, kwargs=dict(
        actions=["p_sted", "p_ex", "pdt"], bleach_sampling={"mode": "constant", "routine": "high-signal_low-bleach"},
        scale_nanodomain_reward=1.0)),
    "ContextualMOSTED-hard-v0": dict(max_episode_steps=10, kwargs=dict(
        actions=["p_sted", "p_ex", "pdt"], bleach_sampling="uniform", scale_nanodomain_reward=1.0)),
    # gym_sted/__init__.py:822-909
    **{f"SequenceMOSTED-{name}-v0": dict(max_episode_steps=10, cls="sequence", kwargs=dict(
        actions=["p_sted", "p_ex", "pdt"], bleach_sampling=sampling, scale_nanodomain_reward=1.0))
       for name, sampling in (("easy", "constant"), ("mid", "choice"), ("hard", "uniform"),
                              *((f"easy-{short}", {"mode": "constant", "routine": routine}) for short, routine in (
                                  ("hslb", "high-signal_low-bleach"), ("hshb", "high-signal_high-bleach"),
                                  ("lslb", "low-signal_low-bleach"), ("lshb", "low-signal_high-bleach"))))},
    # gym_sted/__init__.py:522-531
    "PreferenceCountRateMOSTED-hard-v0": dict(max_episode_steps=30, cls="preference", kwargs=dict(
        actions=["p_sted", "p_ex", "pdt"], bleach_sampling="uniform", max_episode_steps=30)),
}


class PreferenceCountRateScaleSTEDMultiObjectivesEnv(ContextualSTEDMultiObjectivesEnv):
    """Single-environment view of ``VectorPreferenceCountRateMOSTED``
    (``gym_sted/envs/preference_sted_env.py:205-323``; id ``PreferenceCountRateMOSTED-hard-v0``)."""

    vector_cls = VectorPreferenceCountRateMOSTED

    def __init__(self, bleach_sampling="constant", actions=("p_sted", "p_ex", "pdt"), max_episode_steps=30, **kwargs):
        super().__init__(bleach_sampling=bleach_sampling, actions=actions, max_episode_steps=max_episode_steps, **kwargs)


_SINGLE = {"preference": PreferenceCountRateScaleSTEDMultiObjectivesEnv, "sequence": SequenceSTEDMultiObjectivesEnv}
_VECTOR = {"preference": VectorPreferenceCountRateMOSTED, "sequence": VectorSequenceMOSTED}


def make(env_id: str, **overrides):
    """``gym.make`` for the registered ids (accepts the ``gym_sted:`` prefix)."""
    key = env_id.split(":")[-1]
    if key not in REGISTRY:
        raise KeyError(f"unknown environment id {env_id!r}; known: {sorted(REGISTRY)}")
    entry = REGISTRY[key]
    cls = _SINGLE.get(entry.get("cls"), ContextualSTEDMultiObjectivesEnv)
    env = cls(**{"max_episode_steps": entry["max_episode_steps"], **entry["kwargs"], **overrides})
    env.spec = SimpleNamespace(id=key, max_episode_steps=entry["max_episode_steps"])
    return env


def make_vec(env_id: str, num_envs: int, **overrides):
    key = env_id.split(":")[-1]
    entry = REGISTRY[key]
    kwargs = {"max_episode_steps": entry["max_episode_steps"], **entry["kwargs"], **overrides}
    return _VECTOR.get(entry.get("cls"), VectorContextualMOSTED)(num_envs, **kwargs)


def register_with_gym(namespace=None):
    """``register()`` every id of ``REGISTRY`` with ``gymnasium`` or ``gym``, whichever imports
    (``gym_sted/__init__.py:522-531,587-674,822-909``), onto the single-environment classes of this module.
    Returns the list of registered ids (empty when neither package is installed).  Ids are registered as
    ``<id>`` under the entry point ``gym_sted_b200.envs:<Class>`` so ``gym.make("ContextualMOSTED-easy-v0")`` or
    ``gym.make("gym_sted_b200.envs:ContextualMOSTED-easy-v0")`` resolves to the CUDA path."""
    import importlib
    registration = None
    for name in ("gymnasium", "gym"):
        try:
            registration = importlib.import_module(name + ".envs.registration")
            break
        except ImportError:
            continue
    if registration is None:
        return []
    register, registry = registration.register, registration.registry
    done = []
    for key, entry in REGISTRY.items():
        cls = _SINGLE.get(entry.get("cls"), ContextualSTEDMultiObjectivesEnv)
        full = f"{namespace}/{key}" if namespace else key
        if full in registry:
            continue
        register(id=full, entry_point=f"{__name__}:{cls.__name__}", max_episode_steps=entry["max_episode_steps"],
                 kwargs=dict(entry["kwargs"]))
        done.append(full)
    return done


try:                                        # same side effect as ``import gym_sted``: the ids exist after the import
    register_with_gym()
except Exception:                           # an incompatible gym must not break the package
    pass
