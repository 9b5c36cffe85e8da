"""Pins gym_sted_b200/csrc/lbfgsb_small.cuh — the optimiser inside the device-side FluorescenceOptimizer — against
scipy's L-BFGS-B, called exactly as the reference calls it (gym_sted/utils.py:574-594):
``minimize(fun, x0, options={"eps": 0.01, "maxiter": 25}, tol=1e-3, bounds=...)``.  The header is compiled for the
host; objectives are transcendental-free so both sides evaluate them bit-identically.  Interior starts, starts next
to a bound (projected Cauchy path, subspace step), several iterations (BFGS pairs), immediate convergence."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest
import scipy.optimize

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INF = float("inf")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("lb") / "liblb.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tests", "lbfgsb_harness.cpp")])
    lib = ctypes.CDLL(out)
    d, p = ctypes.c_double, ctypes.c_void_p
    lib.lb_harness_1d.restype = ctypes.c_int
    lib.lb_harness_1d.argtypes = [d, d, p, d, d, d, ctypes.c_int, d, p]
    lib.lb_harness_2d.restype = ctypes.c_int
    lib.lb_harness_2d.argtypes = [d, d, d, p, p, p, d, ctypes.c_int, d, p]
    return lib


def _last_eval(fun, x0, **kw):
    """scipy result + the last point the objective was called with."""
    seen = []

    def wrapped(x):
        seen.append(np.array(x, dtype=float))
        return fun(x)
    res = scipy.optimize.minimize(wrapped, x0=x0, **kw)
    return res, seen[-1]


@pytest.mark.parametrize("tol,maxiter", [(1e-3, 25), (1e-9, 60)])
def test_one_variable_matches_scipy(harness, tol, maxiter):
    rng = np.random.default_rng(0)
    moved = 0
    for t in range(60):
        c = rng.uniform(50.0, 800.0)
        x0 = rng.uniform(0.02, 2.0) if t % 3 else rng.uniform(0.005, 0.05)          # every third start hugs the bound 0
        cur = math.floor(c * x0 * 1e5) / 1e5 + 0.1667
        wt = cur * rng.uniform(0.3, 3.0)

        def fun(x):
            s = math.floor(c * float(x[0]) * 1e5) / 1e5 + 0.1667
            e = (wt - s) / wt
            return e * e
        res, last_ref = _last_eval(fun, [x0], options={"eps": 0.01, "maxiter": maxiter}, tol=tol, bounds=[(0.0, np.inf)])
        x, last = np.array([x0]), np.zeros(1)
        harness.lb_harness_1d(c, wt, x.ctypes.data, 0.0, INF, 0.01, maxiter, tol, last.ctypes.data)
        # the reference's settings: equal to rounding; the long runs at a tight tolerance amplify the last-bit differences
        # between scipy's compact limited-memory matrices and the dense ones on this staircase objective
        rel = 1e-12 if tol == 1e-3 else 1e-6
        assert x[0] == pytest.approx(res.x[0], rel=rel, abs=1e-14), (t, res.message)
        assert last[0] == pytest.approx(last_ref[0], rel=rel, abs=1e-14)
        moved += abs(res.x[0] - x0) > 1e-6
    assert moved > 40


@pytest.mark.parametrize("tol,maxiter", [(1e-3, 25), (1e-8, 40)])
def test_two_variables_match_scipy(harness, tol, maxiter):
    rng = np.random.default_rng(1)
    iters = []
    for t in range(80):
        a, c = rng.uniform(0.2, 5.0), rng.uniform(0.0, 3.0)
        x0 = np.array([rng.uniform(0.05, 0.6), rng.uniform(0.2, 4.8)])
        if t % 4 == 0:
            x0[0] = rng.uniform(0.001, 0.02)                       # next to the lower bound of x
        if t % 5 == 0:
            x0[1] = rng.uniform(4.9, 4.999)                        # next to the upper bound of y
        h0 = a * x0[0] * (1.0 + x0[1] + c * x0[1] * x0[1])
        wt = min(max(h0 / (1.0 + h0) + rng.uniform(-0.3, 0.3), 0.02), 0.98)

        def fun(x):
            h = a * float(x[0]) * (1.0 + float(x[1]) + c * float(x[1]) * float(x[1]))
            e = wt - h / (1.0 + h)
            return e * e
        res, last_ref = _last_eval(fun, x0.copy(), options={"eps": 0.01, "maxiter": maxiter}, tol=tol,
                                   bounds=[(0.0, np.inf), (0.0, 5.0)])
        x, last = x0.copy(), np.zeros(2)
        lo, up = np.array([0.0, 0.0]), np.array([INF, 5.0])
        harness.lb_harness_2d(a, c, wt, x.ctypes.data, lo.ctypes.data, up.ctypes.data, 0.01, maxiter, tol, last.ctypes.data)
        if tol == 1e-3:                                             # the reference's settings: equal to rounding
            np.testing.assert_allclose(x, res.x, rtol=1e-9, atol=1e-12, err_msg=f"case {t}: {res.message}")
            np.testing.assert_allclose(last, last_ref, rtol=1e-9, atol=1e-12)
        else:               # long runs down a degenerate valley (one equation, two unknowns): same minimum, nearby point
            assert abs(fun(x) - res.fun) < 1e-9
            np.testing.assert_allclose(x, res.x, rtol=2e-2, atol=1e-6, err_msg=f"case {t}: {res.message}")
        iters.append(res.nit)
    assert max(iters) >= 3
