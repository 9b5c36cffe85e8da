"""the device-side ``FluorescenceOptimizer`` against the reference's own optimiser.

``tests/golden/fluo_optimizer.json`` holds what ``gym_sted/utils.py:517-597`` (unmodified, scipy's L-BFGS-B) returns for
40 sampled criteria on the CPU oracle with a noise-free detector (``tests/golden/make_optimizer_fixture.py``).  The
kernel follows the same optimisation path — including which of the infinitely many (k1, b) pairs that meet a bleach
target comes out — so the three fitted parameters must agree closely, not just the reached targets."""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from gym_sted_b200 import defaults, envs, optimizer  # noqa: E402
from tests import helpers  # noqa: E402

pytestmark = pytest.mark.gpu
CASES = json.load(open(os.path.join(helpers.GOLDEN, "fluo_optimizer.json")))["cases"]


def _opt(sample="clusters"):
    if not torch.cuda.is_available():
        pytest.fail("needs a CUDA device (no CPU fallback exists)")
    m = envs.make_microscope(detector={"noise": False, "background": defaults.DETECTOR["background"]})
    return optimizer.FluorescenceOptimizer(m, sample=sample)


def test_batch_matches_the_reference_optimiser():
    worst = np.zeros(3)
    for sample in sorted({c["sample"] for c in CASES}):
        cases = [c for c in CASES if c["sample"] == sample]
        got = _opt(sample).optimize_batch([c["criterions"] for c in cases])
        for g, c in zip(got, cases):
            err = np.array([abs(g[0] / c["sigma_abs"] - 1), abs(g[1] / c["k1"] - 1), abs(g[2] - c["b"])])
            worst = np.maximum(worst, err)
            # beams by Gauss-Legendre instead of quad, CUDA's pow/exp instead of numpy's, a different summation order:
            # the path is the same, the numbers agree far inside the optimiser's own tolerance (tol = 1e-3)
            assert err[0] < 1e-6 and err[1] < 1e-5 and err[2] < 1e-6, (sample, c["criterions"], g[:3], err)
    print("worst relative sigma_abs, k1 and absolute b errors:", worst)


def test_reached_targets_and_reference_api():
    opt = _opt()
    c = CASES[0]["criterions"]
    out = opt.optimize(c)
    assert set(out) == {"bleach", "signal"} and set(out["bleach"]) == {"k1", "b"}
    assert set(out["signal"]["sigma_abs"]) == {635, 750} and out["signal"]["sigma_abs"][750] == defaults.FLUO["sigma_abs"][750]
    sigma, k1, b, sig, ble = opt.optimize_batch([c])[0]
    assert sig == pytest.approx(c["signal"]["target"], rel=0.02)
    assert ble == pytest.approx(c["bleach"]["target"], abs=0.08)
    # the host twins of the closed forms see the fitted parameters the same way the kernel does
    f = opt.microscope.fluo
    f.sigma_abs, f.k1, f.b = {635: sigma, 750: defaults.FLUO["sigma_abs"][750]}, k1, b
    assert opt.expected_bleach(c["bleach"]["p_ex"], c["bleach"]["p_sted"], c["bleach"]["pdt"]) == pytest.approx(ble, rel=1e-9)
    assert opt.expected_confocal_signal(c["signal"]["p_ex"], 0.0, c["signal"]["pdt"]) == pytest.approx(sig, rel=1e-6)
    only_signal = opt.optimize({"signal": c["signal"]})
    assert set(only_signal) == {"signal"}


def test_hard_environment_resets_with_fitted_fluorophores():
    """ContextualMOSTED-hard-v0: every environment of the batch gets its own fit, spread like the criteria."""
    vec = envs.make_vec("ContextualMOSTED-hard-v0", 32, compute_f1=False, seed=5)
    vec.reset(seed=5)
    crit = vec.bleach_sampler.criteria
    assert len(crit) == 32
    fl = np.frombuffer(vec.bm.d_fluo.cpu().numpy().tobytes(), dtype=np.float64).reshape(32, -1)
    sigma, k1, b = fl[:, 3], fl[:, 11], fl[:, 12]
    assert len(np.unique(sigma)) == 32 and 1e-22 < sigma.min() and sigma.max() < 5e-20
    assert 1.3 < b.min() and b.max() < 2.1 and 1e-16 < k1.min() and k1.max() < 2e-15
    # larger signal targets need larger cross-sections (the signal is linear in sigma_abs * p_ex)
    want = np.array([c["signal"]["target"] / c["signal"]["p_ex"] for c in crit])
    assert np.corrcoef(want, sigma)[0, 1] > 0.95
