"""Registry / gym shim (gym_sted/__init__.py:522-531,587-674): ids, kwargs, registration with an importable gym."""
import sys

import numpy as np
import pytest


def test_registry_matches_reference_ids():
    from gym_sted_b200 import envs
    for key, steps in (("ContextualMOSTED-easy-v0", 10), ("ContextualMOSTED-easy-hslb-v0", 10),
                       ("ContextualMOSTED-hard-v0", 10), ("PreferenceCountRateMOSTED-hard-v0", 30),
                       ("SequenceMOSTED-easy-v0", 10)):
        assert envs.REGISTRY[key]["max_episode_steps"] == steps
        assert envs.REGISTRY[key]["kwargs"]["actions"] == ["p_sted", "p_ex", "pdt"]
    assert envs.REGISTRY["ContextualMOSTED-hard-v0"]["kwargs"]["bleach_sampling"] == "uniform"
    assert envs.REGISTRY["ContextualMOSTED-easy-hslb-v0"]["kwargs"]["bleach_sampling"] == {
        "mode": "constant", "routine": "high-signal_low-bleach"}
    with pytest.raises(KeyError):
        envs.make("NoSuchEnv-v0")


def test_ids_are_registered_with_an_importable_gym(monkeypatch):
    """With ``gym`` (here: the stand-in of tests/shims.py) importable, every id resolves through ``gym.spec``."""
    from tests import shims
    saved = {k: v for k, v in sys.modules.items() if k == "gym" or k.startswith("gym.")}
    try:
        shims._install_gym()
        import gym
        from gym_sted_b200 import envs
        envs.register_with_gym()
        for key, entry in envs.REGISTRY.items():
            sp = gym.spec(key)
            assert sp.entry_point.startswith("gym_sted_b200.envs:") and sp.max_episode_steps == entry["max_episode_steps"]
        assert gym.spec("PreferenceCountRateMOSTED-hard-v0").entry_point.endswith("PreferenceCountRateScaleSTEDMultiObjectivesEnv")
    finally:
        for k in [k for k in sys.modules if k == "gym" or k.startswith("gym.")]:
            del sys.modules[k]
        sys.modules.update(saved)


@pytest.mark.gpu
def test_make_preference_env_single_view():
    """``make("PreferenceCountRateMOSTED-hard-v0")`` is a B = 1 view (BASELINE config 4): 96x96 observation,
    1 + 6*30 vector starting with p_ex/P_EX, PrefNet reward or the -10 count-rate penalty."""
    from gym_sted_b200 import envs
    env = envs.make("gym_sted:PreferenceCountRateMOSTED-hard-v0", seed=3)
    (img, vec), _ = env.reset(seed=3)
    assert img.shape == (96, 96, 3) and img.dtype == np.uint16 and vec.shape == (181,)
    obs, reward, done, trunc, info = env.step(np.array([0.1, 5e-6, 20e-6], dtype=np.float32))
    assert obs[0].shape == (96, 96, 3) and obs[1].shape == (181,) and 0 < obs[1][0] <= 1.0
    assert isinstance(reward, float) and (reward == -10.0 or np.isfinite(reward))
    assert done is False and trunc is False and env.spec.max_episode_steps == 30
    for key in ("mo_objs", "sample", "conf_params", "sted_image"):
        assert key in info


@pytest.mark.gpu
def test_invalid_objective_masks_one_env_instead_of_aborting_the_batch():
    """A None objective (negative SNR) in one environment: the vector env reports it in ``info["invalid"]`` with
    NaN entries for that environment only; the single-environment view raises like the reference's Normalizer."""
    import torch
    from gym_sted_b200 import envs, rewards
    env = envs.make_vec("ContextualMOSTED-easy-v0", 4, seed=0, compute_f1=False)
    env.reset(seed=0)
    real = rewards.foreground_objectives_device

    def with_a_none(conf1, sted, conf2):                 # flag bit 0: SNR < 0, the reference's None
        fg_s, fg_c, objs, flags = real(conf1, sted, conf2)
        flags = flags.clone()
        flags[2] |= 1
        return fg_s, fg_c, objs, flags

    rewards.foreground_objectives_device = with_a_none
    try:
        obs, reward, done, trunc, info = env.step(np.tile([0.1, 10e-6, 10e-6], (4, 1)).astype(np.float32))
        assert info["invalid"].tolist() == [False, False, True, False]
        assert np.isnan(obs[1][2]).any() and not np.isnan(obs[1][[0, 1, 3]]).any()
        single = envs.make("ContextualMOSTED-easy-v0", seed=0, compute_f1=False)
        single.reset(seed=0)

        def always_none(conf1, sted, conf2):
            fg_s, fg_c, objs, flags = real(conf1, sted, conf2)
            return fg_s, fg_c, objs, flags | 1

        rewards.foreground_objectives_device = always_none
        with pytest.raises(TypeError):
            single.step(np.array([0.1, 10e-6, 10e-6], dtype=np.float32))
    finally:
        rewards.foreground_objectives_device = real
