"""Host-side helpers of the path against the reference's own classes, imported as they are (``baseline/_ref``; gym,
skimage, metrics, pysted stood in by ``tests/shims.py``):

* ``Normalizer`` (``gym_sted/utils.py:414-440``) - bit-equal float64 on random action / objective vectors;
* ``PreferenceArticulator`` / ``PrefNet`` (``gym_sted/prefnet/articulation.py:51-75``, ``prefNet.py:11-27``) - the same
  scores and the same ranking from the same weights;
* ``UniformCriterion`` / ``ChoiceCriterion`` (``utils.py:718-769``) - ``BleachSampler.draw_criteria`` consumes the
  Mersenne-Twister stream in the reference's order, so a given seed gives the reference's criteria.
"""
import random

import numpy as np
import pytest

from gym_sted_b200 import defaults, envs, rewards
from tests import shims

pytestmark = pytest.mark.skipif(shims.reference_root() is None, reason="reference not installed (run baseline/install_ref.sh)")


@pytest.fixture(scope="module")
def gu():
    shims.install("product")
    import gym_sted  # noqa: F401
    import gym_sted.utils as ref_utils
    return ref_utils


def test_normalizer_equals_the_reference(gu):
    rng = np.random.default_rng(0)
    for names, scales in ((["p_sted", "p_ex", "pdt"], defaults.action_spaces), (envs.OBJ_NAMES, defaults.scales_dict)):
        ref, ours = gu.Normalizer(names, scales), envs.Normalizer(names, scales)
        for _ in range(50):
            x = rng.uniform(-1.0, 400.0, len(names)) * rng.choice([1e-6, 1e-3, 1.0], len(names))
            np.testing.assert_array_equal(ours(x), ref(x))
        batch = rng.uniform(0.0, 1.0, (7, len(names)))
        np.testing.assert_array_equal(ours(batch), np.stack([ref(row) for row in batch]))


def test_preference_articulator_equals_the_reference(gu):
    from gym_sted.prefnet import PreferenceArticulator as RefArticulator
    ref, ours = RefArticulator(), rewards.PreferenceArticulator(device="cpu")
    rng = np.random.default_rng(1)
    thetas = np.stack([rng.uniform(40, 300, 9), rng.uniform(0, 1, 9), rng.uniform(0, 2, 9)], axis=1)
    for use_sigmoid in (False, True):
        s_ref, best_ref, order_ref = ref.articulate(thetas, use_sigmoid=use_sigmoid)
        s_ours, best_ours, order_ours = ours.articulate(thetas, use_sigmoid=use_sigmoid)
        np.testing.assert_allclose(np.asarray(s_ours).ravel(), np.asarray(s_ref).ravel(), rtol=1e-6, atol=1e-7)
        assert int(best_ours) == int(best_ref) and list(order_ours) == list(order_ref)


@pytest.mark.parametrize("mode,cls", [("uniform", "UniformCriterion"), ("choice", "ChoiceCriterion")])
def test_criteria_follow_the_reference_random_stream(gu, mode, cls):
    for seed in (0, 7, 123):
        random.seed(seed)
        want = [getattr(gu, cls)(gu.defaults.fluorescence_criterions).criterions for _ in range(4)]
        sampler = envs.BleachSampler(mode, seed=seed)
        got = [sampler.draw_criteria() for _ in range(4)]
        assert got == want, (mode, seed)


def test_gaussian_validation_of_the_oracle_equals_the_reference_code(gu):
    """``gym_sted/rewards/utils.py:6-49`` (``fitgaussian`` / ``validate_positions``: numpy + scipy only) run as it is on
    the reference's recorded STED images: the oracle's restatement - the checker of ``gaussfit_kernel`` - returns the same
    fitted parameters bit for bit and keeps the same candidates, also for windows that overlap the image border."""
    from gym_sted.rewards import utils as ref_ru
    from oracle import rewards_oracle as ro
    from tests import helpers
    demo = helpers.load_demonstrations()
    fits = kept = 0
    for i in range(0, 50, 2):
        sted = demo["sted_image"][i].astype(np.int64)
        cand = ro.candidate_peaks(sted)
        extra = np.array([[0, 0], [0, 63], [63, 0], [63, 63], [1, 30]])          # windows padded with zeros
        guesses = np.concatenate([cand.reshape(-1, 2), extra]) if len(cand) else extra
        want = ref_ru.validate_positions(guesses, sted, 20e-9)
        got = ro.validate_positions(guesses, sted, 20e-9)
        assert np.array_equal(np.asarray(got).reshape(-1, 2), np.asarray(want).reshape(-1, 2)), i
        kept += len(want)
        padded = np.pad(sted, 3, constant_values=0)
        for row, col in guesses[:6]:
            data = padded[row:row + 7, col:col + 7]
            np.testing.assert_array_equal(ro.fit_gaussian(data), ref_ru.fitgaussian(data))
            fits += 1
    assert fits >= 100 and kept >= 20


def test_objective_oracle_equals_the_reference_classes_beyond_the_recorded_episodes(gu):
    """``Resolution`` / ``Bleach`` / ``Signal_Ratio`` of ``gym_sted/rewards/objectives.py`` are numpy code: run as they are
    on images the fixture does not hold (other sizes, noise only, bright background, arbitrary masks), they give what
    ``oracle/rewards_oracle.py`` - the checker of ``objectives_kernel`` and ``decorr_kernel`` - gives, bit for bit."""
    from gym_sted.rewards import objectives as ref_obj
    from oracle import rewards_oracle as ro
    rng = np.random.default_rng(11)
    res_ref, snr_ref, bl_ref = ref_obj.Resolution(pixelsize=20e-9, res_cap=300), ref_obj.Signal_Ratio(75), ref_obj.Bleach()
    for k, (H, W) in enumerate([(64, 64), (96, 96), (64, 80), (128, 128), (50, 50), (64, 64)]):
        yy, xx = np.mgrid[:H, :W]
        blobs = sum(rng.uniform(20, 120) * np.exp(-((yy - rng.uniform(8, H - 8)) ** 2 + (xx - rng.uniform(8, W - 8)) ** 2)
                                                  / (2 * rng.uniform(1.2, 3.0) ** 2)) for _ in range(6))
        sted = rng.poisson(blobs * (k != 5) + 0.4)                   # the last image is noise only
        conf1 = rng.poisson(3 * blobs + 2.0)
        conf2 = rng.poisson(2 * blobs + 2.0)
        fg_c, fg_s = rng.random((H, W)) < 0.3, rng.random((H, W)) < 0.15
        if k == 3:
            sted = sted + 50 * (~fg_s)                               # background above foreground: negative ratio
        args = ([sted], conf1, conf2, fg_s, fg_c)
        assert ro.resolution(sted, 20e-9, 300) == res_ref.evaluate(*args), (k, H, W)
        assert ro.bleach(conf1, conf2, fg_c) == bl_ref.evaluate(*args), k
        want = snr_ref.evaluate(*args)
        got = ro.signal_ratio(sted, conf1, fg_s, fg_c)
        assert (got is None and want is None) or got == want, (k, got, want)
        if k == 3:
            assert want is None
