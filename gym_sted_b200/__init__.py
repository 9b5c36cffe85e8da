"""gym_sted_b200 — B200-native acquisition hot path behind gym-sted's environments.

Public surface (see DESIGN.md / INTEGRATION.md):

* ``gym_sted_b200.base``  — pysted.base-compatible ``Microscope`` / ``Datamap`` facade (single env);
* ``gym_sted_b200.batch`` — ``BatchedMicroscope``: whole batches of environments per kernel launch;
* ``gym_sted_b200._lib``  — ctypes binding of ``libsted_b200.so`` (C ABI in ``include/sted_b200.h``).
"""
from . import _lib, defaults, physics  # noqa: F401
from .base import (Datamap, Detector, DonutBeam, Fluorescence, GaussianBeam, Microscope,  # noqa: F401
                   Objective)

__version__ = "0.2.0"


def install_as_pysted(force: bool = False):
    """Register ``gym_sted_b200.pysted`` as the top-level module ``pysted`` (and its ``base`` / ``utils`` /
    ``exp_data_gen`` sub-modules) so that the unmodified reference — ``from pysted import base, utils``,
    ``gym_sted/utils.py:11-12`` — runs on the CUDA path.  Call it before ``import gym_sted``.
    A real ``pysted`` that is already imported is left alone unless ``force`` is set."""
    import sys

    from . import pysted as tree
    if "pysted" in sys.modules and not force and getattr(sys.modules["pysted"], "__name__", "") != tree.__name__:
        existing = sys.modules["pysted"]
        if getattr(existing, "__file__", None):
            raise RuntimeError("a real pysted is already imported; pass force=True to shadow it")
    sys.modules["pysted"] = tree
    for sub in ("base", "utils", "exp_data_gen"):
        sys.modules["pysted." + sub] = getattr(tree, sub)
    return tree
