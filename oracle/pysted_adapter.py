"""Adapter to the REAL pySTED — TEST INFRASTRUCTURE, never imported by the product path.

pySTED (PyPI ``pysted`` / github FLClab/pySTED) is the un-vendored, un-pinned dependency of the reference that holds
the acquisition arithmetic (``/root/reference/setup.py:13``).  It is neither in this image nor in the wheelhouse, so
the acquisition oracle (``oracle/pysted_oracle.py``, ``oracle/raster_oracle.c``) is a restatement from the published
model and every parity claim about rows a1-a5 of SURVEY.md §8 is **unpinned**.  The day ``import pysted`` works (a
wheel dropped into ``baseline/_ref``, a checkout on ``PYTHONPATH`` or ``PYSTED_PATH``) this module finds it and
``tests/test_pysted_parity.py`` runs the oracle AND the CUDA path against it on the same seeded inputs; until then
those tests are skipped with the reason spelled out.

Only the call surface gym-sted itself uses is touched (``gym_sted/utils.py:157-180``, ``envs/mosted_env.py:184-234``).
"""
from __future__ import annotations

import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = [os.environ.get("PYSTED_PATH"), os.path.join(_ROOT, "baseline", "_ref")]
_cached = None


def load():
    """The real ``pysted`` package, or None.  A ``pysted`` entry in ``sys.modules`` that is one of this repo's stand-ins
    (``gym_sted_b200.install_as_pysted``, ``tests/shims.py``) does not count."""
    global _cached
    if _cached is not None:
        return _cached or None
    saved = {k: v for k, v in sys.modules.items() if k == "pysted" or k.startswith("pysted.")}
    for k in saved:
        del sys.modules[k]
    added = [p for p in _CANDIDATES if p and os.path.isdir(p) and p not in sys.path]
    sys.path[:0] = added
    mod = None
    try:
        cand = importlib.import_module("pysted")
        base = importlib.import_module("pysted.base")
        origin = os.path.abspath(getattr(base, "__file__", "") or "")
        if origin and not origin.startswith(os.path.join(_ROOT, "gym_sted_b200")) and hasattr(base, "Microscope"):
            mod = cand
    except Exception:
        mod = None
    finally:
        for p in added:
            sys.path.remove(p)
        if mod is None:
            for k in [k for k in sys.modules if k == "pysted" or k.startswith("pysted.")]:
                del sys.modules[k]
            sys.modules.update(saved)
    _cached = mod or False
    return mod


def available() -> bool:
    return load() is not None


def why_unavailable() -> str:
    return ("the real pySTED is not importable (not in the image, the wheelhouse, baseline/_ref, PYSTED_PATH): the "
            "acquisition oracle stays parity-unpinned (SURVEY.md §8c)")


def microscope(fluo, detector, laser_ex, laser_sted, objective):
    """``MicroscopeGenerator.generate_microscope`` (``gym_sted/utils.py:151-166``) on the real classes."""
    base = importlib.import_module("pysted.base")
    return base.Microscope(base.GaussianBeam(**laser_ex), base.DonutBeam(**laser_sted), base.Detector(**detector),
                           base.Objective(**objective), base.Fluorescence(**fluo), load_cache=False)


def datamap(roi, pixelsize, i_ex):
    base = importlib.import_module("pysted.base")
    dm = base.Datamap(roi, pixelsize)
    dm.set_roi(i_ex, "max")
    return dm
