"""Host-side checks of the reward-side drop-in (``gym_sted_b200.objectives``): the classes the reference environments
instantiate exist with the reference's constructor / ``evaluate`` parameter names (compared with the reference's source
when it is installed), and without a CUDA device every ``evaluate`` fails loudly - there is no CPU path behind them."""
import ast
import inspect
import os

import numpy as np
import pytest

from gym_sted_b200 import _lib, objectives
from tests import shims

NAMES = ["Signal_Ratio", "Bleach", "Resolution", "NumberNanodomains"]
CALCS = ["RewardCalculator", "MORewardCalculator", "NanodomainsRewardCalculator"]


def _ref_signatures(path, classes):
    tree = ast.parse(open(path).read())
    out = {}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in classes:
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name in ("__init__", "evaluate"):
                    out[(node.name, fn.name)] = [a.arg for a in fn.args.args]
    return out


def test_classes_and_attributes():
    for name in NAMES + CALCS + ["Objective"]:
        assert inspect.isclass(getattr(objectives, name)), name
    assert objectives.Signal_Ratio(75).label == "Signal Ratio" and objectives.Signal_Ratio(75).select_optimal is np.argmax
    assert objectives.Bleach().label == "Bleach" and objectives.Bleach().select_optimal is np.argmin
    res = objectives.Resolution(20e-9)
    assert res.label == "Resolution (nm)" and res.res_cap == 250 and res.pixelsize == 20e-9
    assert objectives.NumberNanodomains().threshold == 2 and objectives.NumberNanodomains().label == "NbNanodomains"
    calc = objectives.MORewardCalculator({"Bleach": objectives.Bleach()}, scales={"Bleach": {"low": 0.0, "high": 2.0}})
    assert calc.rescale(1.0, "Bleach") == 0.5 and list(calc.objectives) == ["Bleach"]
    with pytest.raises(ValueError):
        objectives.Signal_Ratio(50)


@pytest.mark.skipif(shims.reference_root() is None, reason="reference not installed (run baseline/install_ref.sh)")
def test_parameter_names_follow_the_reference():
    root = os.path.join(shims.reference_root(), "gym_sted", "rewards")
    ref = _ref_signatures(os.path.join(root, "objectives.py"), NAMES)
    ref.update(_ref_signatures(os.path.join(root, "rewards.py"), CALCS))
    assert len(ref) >= 12
    for (cls, fn), params in ref.items():
        ours = [p for p in inspect.signature(getattr(getattr(objectives, cls), fn)).parameters
                if p not in ("args", "kwargs")]
        assert ours == [p for p in params], (cls, fn, ours, params)


@pytest.mark.skipif(_lib.device_count() > 0, reason="a CUDA device is present: the failure path is for machines without one")
def test_no_cpu_path_behind_the_objectives():
    img = np.ones((64, 64), dtype=np.int64)
    fg = np.ones((64, 64), dtype=bool)
    for call in (lambda: objectives.Bleach().evaluate([img], img, img, fg, fg),
                 lambda: objectives.Signal_Ratio(75).evaluate([img], img, img, fg, fg),
                 lambda: objectives.Resolution(20e-9).evaluate([img], img, img, fg, fg),
                 lambda: objectives.get_foreground(img)):
        with pytest.raises(_lib.StedError):
            call()


@pytest.mark.skipif(shims.reference_root() is None, reason="reference not installed (run baseline/install_ref.sh)")
def test_install_into_gym_sted_replaces_the_shared_objective_table():
    shims.install("product")
    import gym_sted  # noqa: F401
    import gym_sted.defaults as ref_defaults
    from gym_sted.envs import mosted_env
    assert mosted_env.obj_dict is ref_defaults.obj_dict          # one dict, imported by name
    old = objectives.install_into_gym_sted()
    try:
        table = mosted_env.obj_dict
        assert type(table["SNR"]) is objectives.Signal_Ratio and type(table["Bleach"]) is objectives.Bleach
        assert type(table["Resolution"]) is objectives.Resolution and table["Resolution"].res_cap == old["Resolution"].res_cap
        assert table["Resolution"].pixelsize == old["Resolution"].pixelsize
        assert type(table["NbNanodomains"]) is objectives.NumberNanodomains
        assert type(table["Squirrel"]).__module__.startswith("gym_sted")     # untouched
    finally:
        ref_defaults.obj_dict.update(old)
    assert type(ref_defaults.obj_dict["Bleach"]).__module__.startswith("gym_sted")
