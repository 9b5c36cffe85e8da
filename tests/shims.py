"""Minimal stand-ins for the reference's third-party imports that are absent from this image — TEST INFRASTRUCTURE.

They exist so the tests can import the UNMODIFIED reference (``baseline/_ref/gym_sted`` here and on the GPU box,
``/root/reference`` as a fallback in the build container) and run it against either the CPU oracle or
``gym_sted_b200`` presented as ``pysted``:

* ``gym`` (>= 0.26 API surface the reference touches: ``Env``, ``spaces.Box/Tuple/Discrete``, ``utils.seeding``,
  ``envs.registration.register/make/spec/registry``, ``error``);
* ``matplotlib.pyplot`` (imported at module level by every env file, used only by ``render``);
* ``skimage`` (``filters.threshold_otsu/gaussian``, ``measure.label``, ``morphology.remove_small_objects``,
  ``feature.peak_local_max``, ``transform.resize``, ``metrics.structural_similarity``) restated on scipy by
  ``oracle/rewards_oracle.py``;
* ``metrics`` (FLClab/metrics ``CentroidDetectionError``: Hungarian matching, F1);
* ``pysted`` -> the oracle classes (``install(pysted="oracle")``) or the product (``pysted="product"``).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CANDIDATES = [os.path.join(ROOT, "baseline", "_ref"), "/root/reference"]


def reference_root():
    for p in REF_CANDIDATES:
        if os.path.isdir(os.path.join(p, "gym_sted")):
            return p
    return None


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


# ---------------------------------------------------------------------------------------------- gym
class _Space:
    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)


class Box(_Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        dtype = np.dtype(dtype)
        if np.issubdtype(dtype, np.integer):
            info = np.iinfo(dtype)
            low, high = np.clip(low, info.min, info.max), np.clip(high, info.min, info.max)
        shape = tuple(shape) if shape is not None else np.shape(low)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), shape).copy()
        self.shape, self.dtype = shape, dtype
        self._rng = np.random.default_rng()

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


class Discrete(_Space):
    def __init__(self, n):
        self.n, self.shape, self.dtype = int(n), (), np.dtype(np.int64)
        self._rng = np.random.default_rng()

    def sample(self):
        return int(self._rng.integers(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n


class Tuple(_Space, tuple):
    def __new__(cls, spaces):
        return tuple.__new__(cls, spaces)

    def __init__(self, spaces):
        self.spaces = tuple(spaces)

    def sample(self):
        return tuple(s.sample() for s in self.spaces)


class Env:
    metadata, spec, np_random = {}, None, None

    def reset(self, seed=None, options=None):
        if seed is not None or self.np_random is None:
            self.np_random = np.random.default_rng(seed)

    def step(self, action):
        raise NotImplementedError

    def close(self):
        return None

    @property
    def unwrapped(self):
        return self


class EnvSpec:
    def __init__(self, id, entry_point, max_episode_steps=None, kwargs=None, **_):
        self.id, self.entry_point, self.max_episode_steps, self.kwargs = id, entry_point, max_episode_steps, dict(kwargs or {})


_registry: dict = {}


def register(id, entry_point=None, max_episode_steps=None, kwargs=None, **extra):
    _registry[id] = EnvSpec(id, entry_point, max_episode_steps, kwargs, **extra)


def spec(id):
    return _registry[id.split(":")[-1]]


def make(id, **kwargs):
    if ":" in id:
        importlib.import_module(id.split(":")[0])
    sp = spec(id)
    ep = sp.entry_point
    if isinstance(ep, str):
        mod, _, name = ep.partition(":")
        ep = getattr(importlib.import_module(mod), name)
    env = ep(**{**sp.kwargs, **kwargs})
    env.spec = sp
    return env


def _install_gym():
    seeding = _module("gym.utils.seeding", np_random=lambda seed=None: (np.random.default_rng(seed), seed))
    utils = _module("gym.utils", seeding=seeding)
    spaces = _module("gym.spaces", Box=Box, Tuple=Tuple, Discrete=Discrete, Space=_Space)
    error = _module("gym.error", Error=type("Error", (Exception,), {}))
    registration = _module("gym.envs.registration", registry=_registry, register=register, make=make, spec=spec,
                           EnvSpec=EnvSpec)
    envs = _module("gym.envs", registration=registration)
    _module("gym", Env=Env, spaces=spaces, utils=utils, error=error, envs=envs, make=make, register=register,
            spec=spec, __version__="0.26.2-shim")


# ---------------------------------------------------------------------------------------------- skimage / metrics
def _install_skimage():
    import scipy.ndimage as ndi
    from oracle import rewards_oracle as ro

    def gaussian(image, sigma=1, mode="nearest", cval=0, preserve_range=False, truncate=4.0, **_):
        image = np.asarray(image)
        if np.issubdtype(image.dtype, np.integer) and not preserve_range:
            # img_as_float of a signed integer image: (x + 0.5) * 2 / (imax - imin)
            info = np.iinfo(image.dtype)
            image = (image.astype(np.float64) + 0.5) * 2.0 / (float(info.max) - float(info.min))
        return ndi.gaussian_filter(image.astype(np.float64), sigma, mode=mode, cval=cval, truncate=truncate)

    def label(mask, connectivity=None, **_):
        structure = np.ones((3, 3), bool) if connectivity in (None, 2) else None
        return ndi.label(mask, structure=structure)[0]

    def remove_small_objects(mask, min_size=64, connectivity=1, **_):
        structure = np.ones((3, 3), bool) if connectivity == 2 else None
        lab, n = ndi.label(mask, structure=structure)
        if not n:
            return mask.copy()
        small = np.bincount(lab.ravel()) < min_size
        small[0] = False
        return mask & ~small[lab]

    def peak_local_max(image, min_distance=1, labels=None, exclude_border=True, threshold_rel=None, **_):
        if labels is None:
            raise NotImplementedError("shim: only the labelled form the reference uses (objectives.py:422)")
        assert exclude_border is False and threshold_rel is None
        return ro.peak_local_max(np.asarray(image), labels, min_distance=min_distance)

    def resize(image, output_shape, **_):
        zoom = [o / s for o, s in zip(output_shape, image.shape)]
        return ndi.zoom(np.asarray(image, dtype=np.float64), zoom, order=1)

    def not_available(*_, **__):
        raise NotImplementedError("not part of the shimmed surface")

    filters = _module("skimage.filters", threshold_otsu=ro.threshold_otsu, gaussian=gaussian)
    measure = _module("skimage.measure", label=label)
    morphology = _module("skimage.morphology", remove_small_objects=remove_small_objects)
    feature = _module("skimage.feature", peak_local_max=peak_local_max)
    transform = _module("skimage.transform", resize=resize)
    sk_metrics = _module("skimage.metrics", structural_similarity=not_available)
    _module("skimage", filters=filters, measure=measure, morphology=morphology, feature=feature, transform=transform,
            metrics=sk_metrics, __version__="0.22-shim")


class CentroidDetectionError:
    """FLClab/metrics: Hungarian matching of centroids; a pair counts when its distance is below ``threshold``."""

    def __init__(self, truth_coords, guess_coords, threshold, algorithm="hungarian"):
        from oracle import rewards_oracle as ro
        assert algorithm == "hungarian"
        self.tp, self.fp, self.fn = ro.match_counts(truth_coords, guess_coords, threshold)

    @property
    def f1_score(self):
        from oracle import rewards_oracle as ro
        return ro.f1_from_counts(self.tp, self.fp, self.fn)


def _install_misc():
    pyplot = _module("matplotlib.pyplot")
    _module("matplotlib", pyplot=pyplot)
    _module("metrics", CentroidDetectionError=CentroidDetectionError)


# ---------------------------------------------------------------------------------------------- pysted
def _install_pysted(which):
    if which == "product":
        import gym_sted_b200
        gym_sted_b200.install_as_pysted()
        return
    from oracle import pysted_oracle as po
    base = _module("pysted.base", **{k: getattr(po, k) for k in (
        "GaussianBeam", "DonutBeam", "Objective", "Detector", "Fluorescence", "Datamap", "Microscope")})
    utils = _module("pysted.utils")
    dg = _module("pysted.exp_data_gen")
    _module("pysted", base=base, utils=utils, exp_data_gen=dg)


def install(pysted="oracle"):
    """Register the stand-ins and put the reference on ``sys.path``; returns the reference root or None."""
    ref = reference_root()
    if ref is None:
        return None
    _install_gym()
    _install_skimage()
    _install_misc()
    for name in [m for m in sys.modules if m == "pysted" or m.startswith("pysted.")]:
        del sys.modules[name]
    _install_pysted(pysted)
    for name in [m for m in sys.modules if m == "gym_sted" or m.startswith("gym_sted.")]:
        del sys.modules[name]               # re-import against the pysted chosen now
    if ref not in sys.path:
        sys.path.insert(0, ref)
    return ref
