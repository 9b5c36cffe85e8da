"""The constant tables of ``gym_sted_b200.defaults`` against the reference's own module, imported as is
(``gym_sted/defaults.py:15-42,102-162``; ``gym_sted/routines/routines.json``): every entry the hot path reads -
imaging defaults, laser / detector / objective / fluorophore parameters, action spaces, objective bounds and scales,
the fluorophore criteria, the four frozen routines - is equal, key for key.  Skipped when the reference is not installed
(``baseline/install_ref.sh``)."""
import json
import os

import numpy as np
import pytest

from gym_sted_b200 import defaults
from tests import shims

pytestmark = pytest.mark.skipif(shims.reference_root() is None, reason="reference not installed (run baseline/install_ref.sh)")


def _equal(a, b, path=""):
    if isinstance(a, dict):
        assert isinstance(b, dict) and set(a) == set(b), (path, sorted(map(str, a)), sorted(map(str, b)))
        for k in a:
            _equal(a[k], b[k], f"{path}/{k}")
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _equal(x, y, f"{path}[{i}]")
    elif isinstance(a, (int, float, np.floating, np.integer)) and not isinstance(a, bool):
        assert float(a) == float(b), (path, a, b)
    else:
        assert a == b, (path, a, b)


@pytest.fixture(scope="module")
def ref():
    shims.install("product")
    import gym_sted  # noqa: F401
    import gym_sted.defaults as ref_defaults
    return ref_defaults


@pytest.mark.parametrize("name", ["P_STED", "P_EX", "PDT", "LASER_EX", "LASER_STED", "DETECTOR", "OBJECTIVE", "FLUO",
                                  "action_spaces", "scales_dict", "fluorescence_criterions"])
def test_table_equals_the_reference(ref, name):
    _equal(getattr(ref, name), getattr(defaults, name), name)


def test_bounds_of_the_path_objectives(ref):
    for obj in ("SNR", "Bleach", "Resolution", "NbNanodomains"):
        _equal(ref.bounds_dict[obj], defaults.bounds_dict[obj], f"bounds_dict/{obj}")


def test_frozen_routines_equal_routines_json(ref):
    path = os.path.join(os.path.dirname(ref.__file__), "routines", "routines.json")
    routines = json.load(open(path))
    assert set(defaults.ROUTINE_FITS) <= set(routines)
    for name in defaults.ROUTINE_FITS:
        want, got = routines[name], defaults.load_routine(name)
        assert float(want["b"]) == got["b"] and float(want["k1"]) == got["k1"], name
        sig = {int(k): float(v) for k, v in want["sigma_abs"].items()}
        assert sig[635] == got["sigma_abs"][635], name
        for key, val in want.items():                       # everything else of a routine equals FLUO
            if key in ("b", "k1", "sigma_abs"):
                continue
            ours = got[key]
            if isinstance(val, dict):
                _equal({int(k): v for k, v in val.items()}, {int(k): v for k, v in ours.items()}, f"{name}/{key}")
            else:
                _equal(val, ours, f"{name}/{key}")
