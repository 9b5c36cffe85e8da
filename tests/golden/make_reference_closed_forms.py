"""Generates ``reference_closed_forms.json``: outputs of the REFERENCE'S OWN restatement of the acquisition physics
(``FluorescenceOptimizer.expected_bleach`` / ``expected_confocal_signal`` / ``kb_map_to_im_bleach``,
``gym_sted/utils.py:599-661``), executed unmodified from ``/root/reference`` (or ``baseline/_ref``) with the CPU oracle
registered as ``pysted`` (``tests/shims.py``; pySTED itself is not installable here).

What this pins: the reference code's call contract on ``cache`` / ``get_photons`` / ``get_k_bleach`` /
``get_effective`` / ``detector.get_signal`` (argument order, units, the time-averaged STED flux of utils.py:627-630, the
``1 - exp(-sum(kb * pdt))`` bleach estimate) evaluated on the oracle's physics, for the four fitted parameter sets of
``gym_sted/routines/routines.json`` at the criteria of ``gym_sted/routines/create.py:6-63`` plus a grid of actions.
What it cannot pin: pySTED's own constants (SURVEY.md §8c — parity of the acquisition stays "unpinned").

    python tests/golden/make_reference_closed_forms.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests import shims  # noqa: E402

ACTIONS = [  # (p_ex, p_sted, pdt): the criteria of routines/create.py and a few points of the action box
    (10e-6, 150e-3, 30e-6), (2e-6, 150e-3, 25e-6), (10e-6, 0.0, 10e-6), (5e-6, 50e-3, 10e-6), (20e-6, 350e-3, 100e-6),
    (15e-6, 250e-3, 60e-6), (1e-6, 10e-3, 1e-6),
]


def evaluate(gu, fluo):
    """Runs the reference's FluorescenceOptimizer closed forms on a microscope built by the reference's generator."""
    mg = gu.MicroscopeGenerator(fluo=fluo)
    m = mg.generate_microscope()
    m.detector.noise = False                                  # the 25 Poisson repetitions of utils.py:656-659 become exact
    m.detector.background = 0.0
    opt = gu.FluorescenceOptimizer(microscope=m)              # its __init__ resets sigma_abs / k1 / b to defaults.FLUO
    m.fluo.sigma_abs, m.fluo.k1, m.fluo.b = fluo["sigma_abs"], fluo["k1"], fluo["b"]
    out = []
    for p_ex, p_sted, pdt in ACTIONS:
        out.append({"p_ex": p_ex, "p_sted": p_sted, "pdt": pdt,
                    "expected_bleach": float(opt.expected_bleach(p_ex, p_sted, pdt)),
                    "expected_confocal_signal": float(opt.expected_confocal_signal(p_ex, p_sted, pdt))})
    return out


def main():
    ref = shims.install("oracle")
    assert ref is not None, "reference not found"
    import gym_sted.utils as gu
    routines = json.load(open(os.path.join(ref, "gym_sted", "routines", "routines.json")))
    table = {"default": evaluate(gu, gu.defaults.FLUO)}
    for name, fluo in routines.items():
        fluo = {k: ({int(kk): vv for kk, vv in v.items()} if isinstance(v, dict) else v) for k, v in fluo.items()}
        table[name] = evaluate(gu, fluo)
    with open(os.path.join(HERE, "reference_closed_forms.json"), "w") as fh:
        json.dump({"source": "gym_sted/utils.py:599-661 run on oracle/pysted_oracle.py", "values": table}, fh, indent=1)
    print(json.dumps(table["high-signal_low-bleach"][:3], indent=1))


if __name__ == "__main__":
    main()
