"""Pins gym_sted_b200/csrc/lm_minpack.cuh (the Levenberg-Marquardt solver of the Gaussian-fit validation) against
scipy.optimize.leastsq, i.e. the MINPACK LMDIF the reference calls (gym_sted/rewards/utils.py:25-49).  The header is
compiled for the host with g++; on a model without transcendentals the two must agree bit for bit, solution and info
code, at the reference's ftol and at a tight one.  With exp() in the model the only difference left is exp's last bit."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest
import scipy.optimize

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
X, Y = np.indices((7, 7)) - 3.0


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("lm") / "liblm.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tests", "lm_harness.cpp")])
    lib = ctypes.CDLL(out)
    lib.lm_harness_fit.restype = ctypes.c_int
    lib.lm_harness_fit.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_void_p]
    return lib


def windows(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        s, c = rng.uniform(0.6, 4, 2), rng.uniform(-1.5, 1.5, 2)
        out.append(rng.poisson(0.5 + rng.uniform(5, 200) * np.exp(-((X - c[0]) ** 2 / (2 * s[0] ** 2) + (Y - c[1]) ** 2 / (2 * s[1] ** 2)))).astype(np.float64))
    out.append(np.zeros((7, 7)))                     # all-zero window: singular Jacobian path
    out.append(np.full((7, 7), 3.0))                 # flat window
    return out


def rational(p, data):
    wx, wy = float(p[3]) + 1e-3, float(p[4]) + 1e-3
    return np.ravel(p[0] / (1.0 + ((p[1] - X) * (p[1] - X) / (2.0 * (wx * wx)) + (p[2] - Y) * (p[2] - Y) / (2.0 * (wy * wy)))) - data)


_exp = np.frompyfunc(math.exp, 1, 1)


def gauss_libm(p, data):
    wx, wy = float(p[3]) + 1e-3, float(p[4]) + 1e-3
    arg = -1 * ((p[1] - X) ** 2.0 / (2 * wx ** 2.0) + (p[2] - Y) ** 2.0 / (2 * wy ** 2.0))
    return np.ravel(p[0] * _exp(arg).astype(np.float64) - data)


def fit_harness(lib, data, model, ftol):
    x = np.array([data.max(), 0, 0, 1, 1.0])
    d = np.ascontiguousarray(data.ravel())
    info = lib.lm_harness_fit(d.ctypes.data, model, ftol, x.ctypes.data)
    return x, info


@pytest.mark.parametrize("ftol", [1e-3, 1e-10])
def test_rational_model_bit_exact_against_minpack(harness, ftol):
    import warnings
    for data in windows(150, 1):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            p, ier = scipy.optimize.leastsq(rational, [data.max(), 0, 0, 1, 1], args=(data,), ftol=ftol)
        x, info = fit_harness(harness, data, 1, ftol)
        assert info == ier
        assert np.array_equal(x, p), (x, p)


def test_gaussian_model_with_libm_exp(harness):
    exact = 0
    ws = windows(150, 2)
    for data in ws:
        p, ier = scipy.optimize.leastsq(gauss_libm, [data.max(), 0, 0, 1, 1], args=(data,), ftol=1e-3)
        x, info = fit_harness(harness, data, 0, 1e-3)
        exact += bool(np.array_equal(x, p) and info == ier)
        np.testing.assert_allclose(x, p, rtol=1e-6, atol=1e-9)
    assert exact >= 0.9 * len(ws), exact         # the rest: C pow(w, 2.0) vs w*w in the last bit
