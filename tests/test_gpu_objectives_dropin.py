"""The reward-side seam: ``gym_sted_b200.objectives`` offers the classes of
``gym_sted/rewards/objectives.py`` / ``rewards.py`` with the reference's ``evaluate`` signatures on the CUDA kernels.

* on the reference's 50 recorded episodes the calculators return the recorded objectives bit for bit, with the
  recorded masks passed in as the reference environments pass them;
* with arbitrary caller-supplied masks (not Otsu's) they equal the numpy restatement of the reference formulas;
* the reference's None / 0 / NaN edge cases;
* the UNMODIFIED reference environment runs with its calculators swapped for these (GPU + reference present).
"""
from collections import OrderedDict
from types import SimpleNamespace

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from gym_sted_b200 import defaults, objectives  # noqa: E402
from oracle import rewards_oracle as ro  # noqa: E402
from tests import helpers, shims  # noqa: E402

pytestmark = pytest.mark.gpu


def _gpu():
    if not torch.cuda.is_available():
        pytest.fail("these tests need a CUDA device (no CPU fallback exists)")


def _calculators():
    objs = OrderedDict(Resolution=objectives.Resolution(defaults.PIXELSIZE, res_cap=300), Bleach=objectives.Bleach(),
                       SNR=objectives.Signal_Ratio(75))
    scales = {"NbNanodomains": {"low": 0.0, "high": 1.0}}
    return (objectives.MORewardCalculator(objs),
            objectives.NanodomainsRewardCalculator({"NbNanodomains": objectives.NumberNanodomains()}, scales=scales))


def test_recorded_episodes_through_the_reference_interface():
    _gpu()
    demo = helpers.load_demonstrations()
    mo, nb = _calculators()
    for i in range(0, 50, 3):
        got = mo.evaluate(demo["sted_image"][i], demo["conf1"][i], demo["conf2"][i], demo["fg_s"][i], demo["fg_c"][i])
        assert got == list(demo["mo_objs"][i]), (i, got, demo["mo_objs"][i])
        # the one-objective calls of the reference interface give the same numbers as the calculator's fused launch
        args = ([demo["sted_image"][i]], demo["conf1"][i], demo["conf2"][i], demo["fg_s"][i], demo["fg_c"][i])
        assert objectives.Bleach().evaluate(*args) == demo["mo_objs"][i, 1]
        assert objectives.Signal_Ratio(75).evaluate(*args) == demo["mo_objs"][i, 2]
        assert np.array_equal(objectives.get_foreground(demo["conf1"][i]), demo["fg_c"][i])
        n = int((demo["coords"][i, :, 0] >= 0).sum())                  # padded with -1
        synapse = SimpleNamespace(nanodomains_coords=demo["coords"][i, :n].astype(np.int64), datamap_pixelsize_nm=20)
        f1 = nb.evaluate(demo["sted_image"][i], demo["conf1"][i], demo["conf2"][i], demo["fg_s"][i], demo["fg_c"][i],
                         synapse=synapse)
        assert f1 == ro.nanodomain_f1(demo["sted_image"][i], synapse.nanodomains_coords)


def test_caller_supplied_masks_are_honoured():
    _gpu()
    rng = np.random.default_rng(4)
    H = W = 40
    mo, _ = _calculators()
    for k in range(6):
        conf1 = rng.poisson(4.0, (H, W)) + (rng.random((H, W)) < 0.1) * rng.poisson(50, (H, W))
        sted = rng.poisson(1.0, (H, W)) + (rng.random((H, W)) < 0.05) * rng.poisson(30, (H, W))
        conf2 = rng.poisson(3.0, (H, W))
        fg_c = rng.random((H, W)) < 0.3                      # NOT Otsu's masks
        fg_s = rng.random((H, W)) < 0.2
        _, bleach, snr = mo.evaluate(sted, conf1, conf2, fg_s, fg_c)
        assert bleach == ro.bleach(conf1, conf2, fg_c)
        assert snr == ro.signal_ratio(sted, conf1, fg_s, fg_c)


def test_reference_edge_cases():
    _gpu()
    rng = np.random.default_rng(5)
    H = W = 32
    conf1 = rng.poisson(5.0, (H, W)); conf2 = rng.poisson(4.0, (H, W)); sted = rng.poisson(2.0, (H, W))
    fg_c = conf1 > 5
    none = np.zeros((H, W), dtype=bool)
    snr_obj = objectives.Signal_Ratio(75)
    assert snr_obj.evaluate([sted], conf1, conf2, none, fg_c) == 0                     # no STED foreground
    bright_bg = sted.copy(); fg_s = np.zeros((H, W), dtype=bool); fg_s[:2] = True
    bright_bg[2:] += 100                                                               # background above foreground
    assert snr_obj.evaluate([bright_bg], conf1, conf2, fg_s, fg_c) is None             # the reference's error value
    assert np.isnan(objectives.Bleach().evaluate([sted], conf1, conf2, fg_s, none))     # mean of nothing
    with pytest.raises(ValueError):
        objectives.Signal_Ratio(50)
    noise = rng.poisson(3.0, (64, 64))                                                 # no structure: the cap applies
    for cap in (250, 120):
        res = objectives.Resolution(20e-9, res_cap=cap).evaluate([noise], conf1, conf2, fg_s, fg_c)
        assert res <= cap and abs(res - ro.resolution(noise, 20e-9, cap)) <= 1e-9 * cap


@pytest.mark.skipif(shims.reference_root() is None, reason="reference not installed (run baseline/install_ref.sh)")
def test_reference_env_with_the_calculators_swapped():
    """``ContextualMOSTED-easy-v0`` of the unmodified reference, its two reward calculators replaced by ours: every step's
    objectives equal what the reference's own numpy objectives give on the same images and masks."""
    _gpu()
    shims.install("product")
    import gym
    import gym_sted  # noqa: F401
    env = gym.make("gym_sted:ContextualMOSTED-easy-v0")
    ref_mo, ref_nb = env.mo_reward_calculator, env.nb_reward_calculator
    ours_mo = objectives.MORewardCalculator(OrderedDict(
        (name, {"Resolution": objectives.Resolution(ref_mo.objectives["Resolution"].pixelsize,
                                                    ref_mo.objectives["Resolution"].res_cap),
                "Bleach": objectives.Bleach(), "SNR": objectives.Signal_Ratio(75)}[name])
        for name in ref_mo.objectives), scales=ref_mo.scales)
    ours_nb = objectives.NanodomainsRewardCalculator({"NbNanodomains": objectives.NumberNanodomains()}, scales=ref_nb.scales)
    env.mo_reward_calculator, env.nb_reward_calculator = ours_mo, ours_nb
    np.random.seed(0)
    env.reset(seed=0)
    rng = np.random.default_rng(2)
    for _ in range(3):
        action = rng.uniform(env.action_space.low, env.action_space.high).astype(np.float32)
        action[0] = max(action[0], 0.05)
        synapse = env.synapse
        obs, reward, done, truncated, info = env.step(action)
        args = (info["sted_image"], info["conf1"], info["conf2"], info["fg_s"], info["fg_c"])
        want = ref_mo.evaluate(*args)
        np.testing.assert_allclose(info["mo_objs"][1:], want[1:], rtol=1e-12)
        np.testing.assert_allclose(info["mo_objs"][0], want[0], rtol=1e-9)
        assert abs(info["f1-score"] - ref_nb.evaluate(*args, synapse=synapse)) < 1e-12


@pytest.mark.skipif(shims.reference_root() is None, reason="reference not installed (run baseline/install_ref.sh)")
def test_install_into_gym_sted_runs_the_reference_env_on_the_cuda_objectives():
    """``objectives.install_into_gym_sted()``: no source change, the reference's own calculators call the CUDA objectives."""
    _gpu()
    shims.install("product")
    import gym
    import gym_sted  # noqa: F401
    import gym_sted.defaults as ref_defaults
    old = objectives.install_into_gym_sted()
    try:
        env = gym.make("gym_sted:ContextualMOSTED-easy-v0")
        assert type(env.mo_reward_calculator.objectives["Bleach"]) is objectives.Bleach
        assert type(env.nb_reward_calculator.objectives["NbNanodomains"]) is objectives.NumberNanodomains
        np.random.seed(3)
        env.reset(seed=3)
        for _ in range(2):
            synapse = env.synapse
            action = np.array([0.12, 8e-6, 30e-6], dtype=np.float32)
            obs, reward, done, truncated, info = env.step(action)
            args = ([info["sted_image"]], info["conf1"], info["conf2"], info["fg_s"], info["fg_c"])
            want = [old["Resolution"].evaluate(*args), old["Bleach"].evaluate(*args), old["SNR"].evaluate(*args)]
            np.testing.assert_allclose(info["mo_objs"][1:], want[1:], rtol=1e-12)
            np.testing.assert_allclose(info["mo_objs"][0], want[0], rtol=1e-9)
            f1_ref = old["NbNanodomains"].evaluate(info["sted_image"], info["conf1"], info["conf2"], info["fg_s"], info["fg_c"],
                                                   synapse=synapse)
            assert abs(info["f1-score"] - f1_ref) < 1e-12
    finally:
        ref_defaults.obj_dict.update(old)
