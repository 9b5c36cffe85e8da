"""The C-ABI library loads and exports every symbol include/sted_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import numpy as np
import pytest

from gym_sted_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sted_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sted_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 10
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.StedFluoParams) == 16 * 8 + 8
    assert ctypes.sizeof(_lib.StedDetectorParams) == 24
    assert _lib.kpitch(59) == 60 and _lib.kpitch(7) == 8 and _lib.kpitch(63) == 64


# Random123 known-answer vectors for philox4x32-10 (kat_vectors of the Random123 distribution)
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize("ctr,key,expect", KAT)
def test_philox_known_answers(ctr, key, expect):
    lib = _lib.load()
    c = (ctypes.c_uint32 * 4)(*ctr)
    k = (ctypes.c_uint32 * 2)(*key)
    out = (ctypes.c_uint32 * 4)()
    lib.sted_philox4x32_10(c, k, out)
    assert tuple(out) == expect


def test_argument_validation_happens_before_device_use():
    lib = _lib.load()
    det = _lib.StedDetectorParams(0.1, 0.5, 0.0)
    buf = np.zeros(16, dtype=np.float32)
    p = buf.ctypes.data
    # even K is rejected as EINVAL, K > STED_MAX_K as EUNSUPPORTED, in-place bleaching as EINVAL
    assert lib.sted_acquire(p, p, p, p, 1, 4, 4, 4, 0, ctypes.byref(det), 690e-9, 0, 0, 0, None, p, None, 0, None) == -1
    assert lib.sted_acquire(p, p, p, p, 1, 4, 4, 65, 0, ctypes.byref(det), 690e-9, 0, 0, 0, None, p, None, 0, None) == -4
    assert lib.sted_acquire(p, p, p, p, 1, 4, 4, 5, 1, ctypes.byref(det), 690e-9, 0, 0, 0, None, p, None, 0, None) == -1
    assert b"distinct" in lib.sted_last_error()
    assert lib.sted_make_effective(None, p, p, p, p, 1, 5, p, None, None) == -1


@pytest.mark.skipif(_lib.device_count() > 0, reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_device():
    """Product path fails loudly (STED_ENODEVICE) instead of computing on the CPU."""
    lib = _lib.load()
    det = _lib.StedDetectorParams(0.1, 0.5, 0.0)
    buf = np.zeros(4096, dtype=np.float32)
    p = buf.ctypes.data
    rc = lib.sted_acquire(p, None, p, p, 1, 4, 4, 5, 0, ctypes.byref(det), 690e-9, 0, 0, 0, None, p, None, 0, None)
    assert rc == -2 and b"no CPU fallback" in lib.sted_last_error()
    with pytest.raises(_lib.StedError):
        _lib.check(rc)
    from tests import helpers
    m = helpers.product_microscope()
    from gym_sted_b200 import base
    dm = base.Datamap(np.ones((8, 8)), 20e-9)
    dm.set_roi(m.cache(20e-9)[0], "max")
    with pytest.raises(_lib.StedError):
        m.get_signal_and_bleach(dm, 20e-9, pdt=10e-6, p_ex=10e-6, p_sted=0.0, bleach=False)
