"""ctypes binding of ``libsted_b200.so`` (the C ABI declared in ``include/sted_b200.h``).

There is no CPU fallback: importing this module without the built library raises, and every
compute entry point returns ``STED_ENODEVICE`` (raised here as ``StedError``) without a GPU.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("STED_B200_LIB", os.path.join(_HERE, "libsted_b200.so"))

STED_BLEACH, STED_SAMPLE_PHOTONS, STED_SAMPLE_BLEACH, STED_NO_DETECT, STED_HOST_U16, STED_FRAME_MOLECULES = 1, 2, 4, 8, 16, 32
STED_PRESAMPLED = 64
STED_TAB_E, STED_TAB_ET, STED_TAB_CH, STED_TAB_RS, STED_TAB_E2, STED_TABLES = 0, 1, 2, 3, 4, 8
STED_MAX_K = 63


def kpitch(K: int) -> int:
    return ((K + 1) + 3) & ~3


class StedError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"sted_b200 error {code}: {message}")
        self.code = code


class StedFluoParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in (
        "lambda_ex", "lambda_sted", "lambda_fluo", "sigma_abs_ex", "sigma_abs_sted", "sigma_ste", "qy", "tau",
        "tau_vib", "tau_tri", "k0", "k1", "b", "triplet_dynamics_frac", "sted_tau", "sted_rate")] + [
        ("anti_stoke", ctypes.c_int32), ("_pad", ctypes.c_int32)]


class StedFluoCriteria(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in ("sig_p_ex", "sig_p_sted", "sig_pdt", "sig_target", "bl_p_ex", "bl_p_sted",
                                               "bl_pdt", "bl_target")]


class StedSynapseParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("molecules", "n_lo", "n_hi", "m_lo", "m_hi", "_pad")] + [
        (n, ctypes.c_double) for n in ("min_dist_px", "thick_lo", "thick_hi")]


class StedDetectorParams(ctypes.Structure):
    _fields_ = [("pcef", ctypes.c_double), ("pdef", ctypes.c_double), ("background", ctypes.c_double)]


_vp, _i, _u32, _u64, _i64, _d = (ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int64,
                                 ctypes.c_double)

# name -> (restype, argtypes); every symbol include/sted_b200.h declares
SIGNATURES = {
    "sted_version": (ctypes.c_char_p, []),
    "sted_last_error": (ctypes.c_char_p, []),
    "sted_device_count": (_i, []),
    "sted_make_effective": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "sted_acquire": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _u32, ctypes.POINTER(StedDetectorParams), _d, _u64,
                          _u64, _i64, _vp, _vp, _vp, _u64, _vp]),
    "sted_acquire_dwellmaps": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _u32,
                                    ctypes.POINTER(StedDetectorParams), _d, _u64, _u64, _i64, _vp, _vp, _vp]),
    "sted_bleach_presample": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _u32, _u64, _u64, _i64, _vp, _u64, _vp]),
    "sted_bleach_workspace_bytes": (_u64, [_i, _i, _u64]),
    "sted_bleach_dropped_events": (_i, [_vp, ctypes.POINTER(ctypes.c_int), _vp]),
    "sted_host_thread_priority": (_i, [_i]),
    "sted_detect": (_i, [_vp, _vp, _i, _i, _i, _u32, ctypes.POINTER(StedDetectorParams), _d, _u64, _u64, _i64, _vp]),
    "sted_acquire_host": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _u32,
                               ctypes.POINTER(StedDetectorParams), _u64, _u64, _i64, _vp, _vp]),
    "sted_session_create": (_i, [_i, _i, _i, _i, _vp, _i, ctypes.POINTER(_vp)]),
    "sted_session_destroy": (_i, [_vp]),
    "sted_session_set_fluorophores": (_i, [_vp, _vp]),
    "sted_session_upload": (_i, [_vp, _vp, _u32, _u64]),
    "sted_session_acquire": (_i, [_vp, _vp, _vp, _vp, _u32, ctypes.POINTER(StedDetectorParams), _d, _u64, _u64, _i64, _vp,
                                  _vp]),
    "sted_session_upload_device": (_i, [_vp, _vp, _u32, _u64, _vp]),
    "sted_session_result": (_i, [_vp, _i, ctypes.POINTER(_vp), ctypes.POINTER(_vp), _vp, _i]),
    "sted_session_kernel_ms": (_i, [_vp, _i, ctypes.POINTER(ctypes.c_float)]),
    "sted_session_join": (_i, [_vp, _vp]),
    "sted_session_adopt": (_i, [_vp, _vp]),
    "sted_session_after": (_i, [_vp, _vp]),
    "sted_session_download": (_i, [_vp, _vp, _u32]),
    "sted_session_sync": (_i, [_vp, ctypes.POINTER(_i64)]),
    "sted_session_device_map": (_i, [_vp, ctypes.POINTER(_vp)]),
    "sted_fluo_optimize": (_i, [_vp, _vp, _i, _i, _vp, ctypes.POINTER(StedDetectorParams), _vp, _i, _i, _vp, _vp]),
    "sted_sample_poisson": (_i, [_vp, _i64, _u64, _u64, _vp, _vp]),
    "sted_sample_binomial": (_i, [_vp, _vp, _i64, _u64, _u64, _vp, _vp]),
    "sted_philox4x32_10": (None, [ctypes.POINTER(_u32), ctypes.POINTER(_u32), ctypes.POINTER(_u32)]),
    "sted_foreground_objectives": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sted_objectives_masked": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sted_nanodomain_candidates": (_i, [_vp, _i, _i, _i, ctypes.POINTER(_d), _i64, _d, _vp, _vp, _vp, _vp, _vp]),
    "sted_decorrelation_resolution": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _d, _d, _d, _d, _d, _vp, _vp]),
    "sted_gaussfit_validate": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _d, _vp, _vp, _vp]),
    "sted_f1_match": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _d, _vp, _vp, _vp, _vp]),
    "sted_apodize_for_fft": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "sted_spectrum_gather": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "sted_synapse_scratch_bytes": (_u64, [_i, _i, _i]),
    "sted_synapse_generate": (_i, [ctypes.POINTER(StedSynapseParams), _i, _i, _i, _u64, _u64, _i64, _vp, _vp, _vp, _i, _vp, _vp]),
    "sted_pad_maps": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "sted_structure_maps": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "sted_pack_observation": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "sted_lmfit_test": (_i, [_vp, _i, _i, _d, _vp, _vp, _vp]),
    "sted_debug_phase_cycles": (_i, [ctypes.POINTER(ctypes.c_uint64), _i]),
    "sted_fma_peak": (_i, [ctypes.POINTER(_d), _vp]),
}

_lib = None


def load():
    """Load the shared library (raises ``OSError`` with a build hint when it is missing)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). gym_sted_b200 has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int):
    if rc != 0:
        raise StedError(rc, load().sted_last_error().decode())


def device_count() -> int:
    return load().sted_device_count()
