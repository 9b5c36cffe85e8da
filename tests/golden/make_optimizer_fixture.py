"""Generates ``fluo_optimizer.json``: the REFERENCE'S ``FluorescenceOptimizer.optimize`` (``gym_sted/utils.py:517-597``,
run unmodified from ``/root/reference`` / ``baseline/_ref`` with scipy's L-BFGS-B) on criteria drawn like
``UniformCriterion`` does (``utils.py:741-769``, ranges of ``defaults.fluorescence_criterions``), with the CPU oracle
registered as ``pysted`` and the detector's noise switched off (the reference averages 25 Poisson draws per objective
evaluation; the fixture pins the deterministic path, which is what ``sted_fluo_optimize`` follows).

    python tests/golden/make_optimizer_fixture.py        # about a second per case
"""
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests import helpers, shims  # noqa: E402

N_CASES = 40


def main():
    assert shims.install("oracle") is not None, "reference not found"
    import gym_sted.utils as gu
    random.seed(42)
    np.random.seed(42)
    cases = []
    for i in range(N_CASES):
        sample = ("clusters", "clusters", "actin", "tubulin", "synthetic")[i % 5]
        opt = gu.FluorescenceOptimizer(sample=sample)
        opt.microscope._cache[20] = helpers.oracle_psf_cache()        # the oracle's beams, memoised (quad per pixel)
        opt.microscope.detector.noise = False
        crit = gu.UniformCriterion(gu.defaults.fluorescence_criterions)
        if i % 8 == 7:                                                # the extremes of the ranges
            crit.criterions["signal"]["target"] = 10.0 if i % 16 == 7 else 400.0
            crit.criterions["bleach"]["target"] = 0.9 if i % 16 == 7 else 0.1
        out = opt.optimize(crit)
        cases.append({"sample": sample, "criterions": crit.criterions,
                      "sigma_abs": out["signal"]["sigma_abs"][635], "k1": out["bleach"]["k1"], "b": out["bleach"]["b"]})
        print(i, sample, cases[-1]["sigma_abs"], cases[-1]["k1"], cases[-1]["b"], flush=True)
    with open(os.path.join(HERE, "fluo_optimizer.json"), "w") as fh:
        json.dump({"source": "gym_sted/utils.py:517-597 (scipy L-BFGS-B) on oracle/pysted_oracle.py, detector.noise = False",
                   "cases": cases}, fh, indent=1)


if __name__ == "__main__":
    main()
