"""The drop-in claim, tested: the UNMODIFIED reference (``gym_sted`` as installed under ``baseline/_ref`` by
``baseline/install_ref.sh``; ``/root/reference`` as a fallback in the build container) runs with
``gym_sted_b200`` registered as ``pysted`` (``gym_sted_b200.install_as_pysted``).

The reference's other third-party imports that are absent from this image (gym, skimage, matplotlib, metrics)
are replaced by the minimal stand-ins of ``tests/shims.py``; nothing of the reference itself is edited.
CPU part: imports, object construction, the host-side accessors ``FluorescenceOptimizer`` touches
(``gym_sted/utils.py:599-632``).  GPU part: ``gym.make`` of the BASELINE ids, ``reset`` + ``step`` through the
reference's own env classes, ``BleachSampler("uniform").sample()``.
"""
import numpy as np
import pytest

from tests import shims

REF = shims.reference_root()
needs_ref = pytest.mark.skipif(REF is None, reason="reference not installed (run baseline/install_ref.sh)")


@pytest.fixture()
def ref_on_product():
    shims.install("product")
    import gym_sted  # noqa: F401  (registers the ids with the gym stand-in)
    import gym_sted.utils as gu
    return gu


@needs_ref
def test_reference_imports_against_the_facade(ref_on_product):
    import pysted
    import pysted.base
    from pysted import exp_data_gen as dg
    import gym_sted_b200
    assert pysted.base.Microscope is gym_sted_b200.base.Microscope
    assert dg.Synapse is gym_sted_b200.synapse.Synapse
    import gym
    import gym_sted.envs  # noqa: F401  every env module imports (timed/debug ones included)
    for env_id in ("ContextualMOSTED-easy-v0", "ContextualMOSTED-easy-hslb-v0", "ContextualMOSTED-hard-v0",
                   "PreferenceCountRateMOSTED-hard-v0", "SequenceMOSTED-easy-v0"):
        assert gym.spec(env_id).max_episode_steps in (10, 30)


@needs_ref
def test_microscope_generator_and_host_accessors(ref_on_product):
    gu = ref_on_product
    from gym_sted_b200 import base
    mg = gu.MicroscopeGenerator()
    m = mg.generate_microscope()
    assert isinstance(m, base.Microscope)
    i_ex, i_sted, psf_det = m.cache(mg.pixelsize)
    assert i_ex.shape == i_sted.shape == psf_det.shape == (59, 59)
    dm = mg.generate_datamap()                               # 50x50 default disposition, ROI "max"
    assert dm.whole_datamap.shape == (108, 108) and dm.roi == (slice(29, 79), slice(29, 79))
    assert dm["base"] is dm.whole_datamap and dm.pixelsize == 20e-9
    # the physics accessors of utils.py:616-631 (FluorescenceOptimizer.expected_bleach runs on the host)
    opt = gu.FluorescenceOptimizer(microscope=m)
    b1 = opt.expected_bleach(10e-6, 150e-3, 30e-6)
    b2 = opt.expected_bleach(10e-6, 150e-3, 60e-6)
    assert 0 < b1 < b2 < 1
    np.testing.assert_allclose(1 - b2, (1 - b1) ** 2, rtol=1e-9)      # hazard is linear in the dwell time
    phi = m.fluo.get_photons(i_ex * 1e-6, lambda_=m.excitation.lambda_)
    assert phi.shape == (59, 59) and phi.max() > 1e20
    # the sampler modes that need no optimisation
    assert gu.BleachSampler("constant").sample() is gu.defaults.FLUO
    synapse = gu.SynapseGenerator(mode="mushroom", seed=3, n_nanodomains=(3, 15), n_molecs_in_domain=(100, 175))(rotate=True)
    assert synapse.frame.shape == (64, 64) and 3 <= len(synapse.nanodomains_coords) <= 15
    assert all(synapse.frame[y, x] >= 100 for y, x in synapse.nanodomains_coords)
    assert synapse.datamap_pixelsize_nm == 20


def _check_step(obs, reward, done, truncated, info, size=64, vec=60):
    img, v = obs
    assert img.shape == (size, size, 3) and img.dtype == np.uint16
    assert v.shape == (vec,) and v.dtype == np.float32
    assert truncated is False and isinstance(done, (bool, np.bool_))
    assert 0.0 <= float(reward) <= 1.0
    for key in ("action", "bleached", "sted_image", "conf1", "conf2", "fg_c", "fg_s", "mo_objs", "reward", "f1-score",
                "nanodomains-coords"):
        assert key in info, key
    assert info["sted_image"].shape == (size, size) and info["bleached"].shape == (size, size)
    assert info["conf1"].sum() > info["conf2"].sum() * 0.5
    assert info["bleached"].sum() > 0


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("env_id", ["ContextualMOSTED-easy-v0", "ContextualMOSTED-easy-hslb-v0"])
def test_reference_contextual_env_runs_on_the_cuda_path(ref_on_product, env_id):
    import gym
    from gym_sted_b200 import base
    env = gym.make("gym_sted:" + env_id)
    assert type(env).__module__ == "gym_sted.envs.contextual_sted_env"      # the reference's class, not ours
    assert isinstance(env.microscope, base.Microscope)
    np.random.seed(0)
    (img, vec), info = env.reset(seed=0)
    assert img.shape == (64, 64, 3) and img.dtype == np.uint16 and img[..., 0].max() > 0 and not img[..., 1:].any()
    assert vec.shape == (60,) and not vec.any()
    rng = np.random.default_rng(1)
    for step in range(3):
        action = rng.uniform(env.action_space.low, env.action_space.high).astype(np.float32)
        action[0] = max(action[0], 0.05)
        obs, reward, done, truncated, info = env.step(action)
        _check_step(obs, reward, done, truncated, info)
        assert np.count_nonzero(obs[1]) == 6 * (step + 1) or np.count_nonzero(obs[1]) >= 6 * (step + 1) - 2
        res, bleach, snr = info["mo_objs"]
        assert 0 < res <= 300 and -0.2 < bleach <= 1.0
        assert not done
    env.close()


@needs_ref
@pytest.mark.gpu
def test_reference_sequence_env_carries_bleaching(ref_on_product):
    """``SequenceSTEDMultiObjectivesEnv`` (gym_sted/envs/sequence_sted_env.py) re-images the same datamap: the
    in-place mutation of the boundary must make the molecule count fall from step to step."""
    import gym
    env = gym.make("gym_sted:SequenceMOSTED-easy-v0")
    env.reset(seed=1)
    totals = [int(env.datamap.whole_datamap.sum())]
    for _ in range(3):
        env.step(np.array([0.2, 10e-6, 40e-6], dtype=np.float32))
        totals.append(int(env.datamap.whole_datamap.sum()))
    assert totals[0] > totals[1] > totals[2] > totals[3] > 0


@needs_ref
@pytest.mark.gpu
def test_reference_bleach_sampler_uniform_runs(ref_on_product):
    """``BleachSampler("uniform")`` (gym_sted/utils.py:339-351): the reference's own 25 x L-BFGS-B loop, every
    objective evaluation going through ``fluo.get_k_bleach`` / ``microscope.get_effective`` of the facade."""
    import random
    gu = ref_on_product
    random.seed(42)
    np.random.seed(42)
    sampler = gu.BleachSampler("uniform")
    fluo = sampler.sample()
    crit = sampler.criterion.criterions
    # criteria are uniform in gym_sted/defaults.py:149-162: up to 400 photons at as little as 1 uW of excitation needs
    # sigma_abs up to ~8e-20 m^2 (3.2e-21 gives ~163 photons at 10 uW); the fit is run on a Poisson-noisy objective
    assert 1e-22 < fluo["sigma_abs"][635] < 1.2e-19 and 0 < fluo["k1"] < 1e-14 and 0 < fluo["b"] <= 5
    opt = sampler.optimizer
    opt.microscope.fluo.sigma_abs, opt.microscope.fluo.k1, opt.microscope.fluo.b = fluo["sigma_abs"], fluo["k1"], fluo["b"]
    reached = opt.expected_bleach(crit["bleach"]["p_ex"], crit["bleach"]["p_sted"], crit["bleach"]["pdt"])
    assert abs(reached - crit["bleach"]["target"]) < 0.12
    signal = opt.expected_confocal_signal(crit["signal"]["p_ex"], 0.0, crit["signal"]["pdt"])
    assert abs(signal / crit["signal"]["target"] - 1) < 0.25
