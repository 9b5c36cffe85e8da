"""gym_sted_b200 — B200-native acquisition hot path behind gym-sted's environments.

Public surface (see DESIGN.md / INTEGRATION.md):

* ``gym_sted_b200.base``  — pysted.base-compatible ``Microscope`` / ``Datamap`` facade (single env);
* ``gym_sted_b200.batch`` — ``BatchedMicroscope``: whole batches of environments per kernel launch;
* ``gym_sted_b200._lib``  — ctypes binding of ``libsted_b200.so`` (C ABI in ``include/sted_b200.h``).
"""
from . import _lib, defaults, physics  # noqa: F401
from .base import (Datamap, Detector, DonutBeam, Fluorescence, GaussianBeam, Microscope,  # noqa: F401
                   Objective)

__version__ = "0.1.0"
