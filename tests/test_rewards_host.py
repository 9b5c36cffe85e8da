"""CPU suite for the reward/observation host logic of the product package (no GPU needed)."""
import numpy as np
import pytest
import torch

from gym_sted_b200 import defaults, rewards
from oracle import rewards_oracle as ro
from tests import helpers


def test_prefnet_known_answer():
    """SURVEY App. B probe: model 2021-07-12 scores mo_objs = [80, 0.2, 0.9] at 0.9572."""
    art = rewards.PreferenceArticulator()
    scores, best, order = art.articulate([[80, 0.2, 0.9], [200, 0.9, 0.1]])
    assert scores.shape == (2, 1) and scores[0, 0] == pytest.approx(0.9572, abs=5e-4)
    assert best == 0 and list(order) == [1, 0]


def test_registry_and_spaces_without_gpu():
    from gym_sted_b200 import envs
    assert set(envs.REGISTRY) == {"ContextualMOSTED-easy-v0", "ContextualMOSTED-easy-hslb-v0", "ContextualMOSTED-hard-v0",
                                  "PreferenceCountRateMOSTED-hard-v0", "SequenceMOSTED-easy-v0", "SequenceMOSTED-mid-v0",
                                  "SequenceMOSTED-hard-v0", "SequenceMOSTED-easy-hslb-v0", "SequenceMOSTED-easy-hshb-v0",
                                  "SequenceMOSTED-easy-lslb-v0", "SequenceMOSTED-easy-lshb-v0"}
    for key, entry in envs.REGISTRY.items():
        assert entry["max_episode_steps"] == (30 if key.startswith("Preference") else 10)
        assert entry["kwargs"]["actions"] == ["p_sted", "p_ex", "pdt"]
    gen = envs.SyntheticDatamapGenerator(seed=0)
    for g in gen.GROUPS:
        d = gen(g)
        assert d.shape == (96, 96) and d.min() == 0 and 15 <= d.max() <= 40 and np.array_equal(d, np.floor(d))
    box = envs.Box(np.array([0, 0, 1e-6], dtype=np.float32), np.array([0.35, 20e-6, 100e-6], dtype=np.float32))
    assert box.shape == (3,) and box.contains(box.sample(np.random.default_rng(0)))
    norm = envs.Normalizer(["p_sted", "p_ex", "pdt"], defaults.action_spaces)
    np.testing.assert_allclose(norm([0.35, 10e-6, 1e-6]), [1.0, 0.5, 0.0])
    norm = envs.Normalizer(envs.OBJ_NAMES, defaults.scales_dict)
    np.testing.assert_allclose(norm([110, 0.5, 0.25]), [0.5, 0.5, 0.25])
    s = envs.BleachSampler("constant", routine="high-signal_low-bleach").sample()
    assert s["k1"] == pytest.approx(3.626623963067252e-16) and s["sigma_abs"][635] == pytest.approx(3.2482521692221336e-21)
    assert envs.BleachSampler("constant").sample() is defaults.FLUO
    with pytest.raises(KeyError):
        envs.make("NoSuchEnv-v0")


def test_synthetic_synapse_generator():
    from gym_sted_b200 import synapse
    s = synapse.generate(64, 64, seed=5)
    assert s.frame.shape == (64, 64) and s.datamap_pixelsize_nm == 20
    n = len(s.nanodomains_coords)
    assert 3 <= n <= 15
    vals = s.frame[tuple(s.nanodomains_coords.T)]
    assert vals.min() >= 100 and vals.max() <= 175
    d = np.linalg.norm(s.nanodomains_coords[:, None] - s.nanodomains_coords[None], axis=-1)
    assert d[np.triu_indices(n, 1)].min() >= 5
    assert set(np.unique(s.frame)) - set(vals.tolist()) <= {0, 5}
    big, coords = synapse.generate_batch(2, 256, 256, seed=1)
    assert big.shape == (2, 256, 256) and 0.05 < (big > 0).mean() < 0.3
    assert np.array_equal(synapse.generate_batch(2, 256, 256, seed=1)[0], big)


def test_batched_structure_maps_follow_the_single_map_model():
    """``SyntheticDatamapGenerator.batch`` (device rendering; here on the CPU device) draws from the same procedural
    model as ``__call__``: integer molecule counts in [0, 40], empty below the 0.5 structure threshold, filament
    groups cover more pixels than cluster groups."""
    from gym_sted_b200 import envs
    gen = envs.SyntheticDatamapGenerator((96, 96), seed=4)
    groups = [gen.GROUPS[i % 6] for i in range(48)]
    maps = gen.batch(groups, "cpu").numpy()
    assert maps.shape == (48, 96, 96) and np.array_equal(maps, np.floor(maps)) and maps.min() == 0 and 20 <= maps.max() <= 40
    single = np.stack([gen(g) for g in groups])
    for kind in (("actin", "tubulin", "lifeact"), ("psd95", "camkii", "PSD95-Bassoon")):
        sel = np.array([g in kind for g in groups])
        cover_b, cover_s = (maps[sel] > 0).mean(), (single[sel] > 0).mean()
        assert 0.5 < cover_b / cover_s < 2.0, (kind, cover_b, cover_s)
        assert abs(maps[sel][maps[sel] > 0].mean() - single[sel][single[sel] > 0].mean()) < 3


def test_batched_matching_equals_hungarian():
    """``match_counts_batch`` (vectorised; Hungarian only for contested pairs) against ``match_counts`` per image,
    including empty sets, duplicates and contested detections."""
    rng = np.random.default_rng(7)
    truths, guesses = [], []
    for i in range(200):
        t = rng.integers(5, 59, (rng.integers(0, 15), 2))
        g = [t[:len(t) // 2] + rng.integers(-1, 2, (len(t) // 2, 2)), rng.integers(5, 59, (rng.integers(0, 4), 2))]
        if i % 3 == 0 and len(t) >= 2:
            g.append(t[:2] + 1)                                  # two detections competing for the same truths
        truths.append(t)
        guesses.append(np.vstack(g) if i % 7 else np.zeros((0, 2), dtype=np.int64))
    ref = np.array([rewards.match_counts(t, g) for t, g in zip(truths, guesses)])
    assert np.array_equal(rewards.match_counts_batch(truths, guesses), ref)
    assert rewards.match_counts_batch([], []).shape == (0, 3)
