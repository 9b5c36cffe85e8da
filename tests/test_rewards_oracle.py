"""The reward-side oracle against the reference's own recorded episodes (golden fixture)."""
import numpy as np
import pytest

from oracle import rewards_oracle as ro
from tests import helpers


@pytest.fixture(scope="module")
def demo():
    return helpers.load_demonstrations()


def test_fixture_shape(demo):
    assert demo["sted_image"].shape == (50, 64, 64) and demo["mo_objs"].shape == (50, 3)
    assert demo["fg_c"].shape == (50, 64, 64) and demo["fg_c"].dtype == bool


def test_otsu_foregrounds_reproduce_recorded_masks(demo):
    for i in range(50):
        fg_s, fg_c = ro.foregrounds(demo["conf1"][i], demo["sted_image"][i])
        assert np.array_equal(fg_c, demo["fg_c"][i]), i
        assert np.array_equal(fg_s, demo["fg_s"][i]), i


def test_objectives_reproduce_recorded_values_exactly(demo):
    for i in range(50):
        res, bl, snr = ro.mo_objectives(demo["sted_image"][i], demo["conf1"][i], demo["conf2"][i], demo["fg_s"][i],
                                        demo["fg_c"][i])
        assert res == demo["mo_objs"][i, 0], (i, res, demo["mo_objs"][i, 0])
        assert bl == demo["mo_objs"][i, 1], i
        assert snr == demo["mo_objs"][i, 2], i


def test_f1_formula_and_matching_rule_against_recorded_scores(demo):
    """The recorded F1 came from the older detector (commented line objectives.py:426):
    peak_local_max(sted, min_distance=2, threshold_rel=0.5); the ground truth of record i is coords[i-1]
    (SURVEY App. B caveats).  This pins the F1 formula and the Hungarian rule, not today's detector."""
    import scipy.ndimage as ndi
    hits = total = 0
    for i in range(50):
        if demo["episode"][i] != demo["episode"][i - 1] or i == 0:
            continue
        img = demo["sted_image"][i].astype(np.float64)
        truth = demo["coords"][i - 1]
        truth = truth[truth[:, 0] >= 0]
        thr = max(img.min(), 0.5 * img.max())
        pm = (img == ndi.maximum_filter(img, size=5, mode="nearest")) & (img > thr)
        pm[:2] = pm[-2:] = False                       # exclude_border=True (the default) removes min_distance px
        pm[:, :2] = pm[:, -2:] = False
        cc = np.transpose(np.nonzero(pm))
        cc = cc[np.argsort(-img[tuple(cc.T)], kind="stable")]
        guess = ro._ensure_spacing(cc, 2)
        total += 1
        hits += abs(ro.f1_score(truth, guess) - demo["f1"][i]) < 1e-9
    assert total == 45 and hits >= 36, (hits, total)


def test_f1_edge_cases():
    assert ro.f1_score(np.zeros((0, 2)), np.zeros((0, 2))) == 0
    assert ro.match_counts([[5, 5]], [[5, 6], [20, 20]]) == (1, 1, 0)
    assert ro.match_counts([[5, 5], [9, 9]], [[5, 7]]) == (0, 1, 2)        # distance 2 is not < 2
    assert ro.f1_score([[1, 1], [8, 8]], [[1, 1], [8, 8]]) == pytest.approx(1.0, abs=2e-6)


def test_detection_finds_isolated_gaussian_spots():
    rng = np.random.default_rng(0)
    img = rng.poisson(0.2, (64, 64)).astype(np.int64)
    truth = [(12, 14), (30, 40), (50, 22)]
    yy, xx = np.mgrid[:64, :64]
    for (y, x) in truth:
        img += np.rint(120 * np.exp(-((yy - y) ** 2 + (xx - x) ** 2) / (2 * 2.0 ** 2))).astype(np.int64)
    found = ro.detect_nanodomains(img)
    assert ro.match_counts(truth, found) == (3, 0, 0)
    assert ro.nanodomain_f1(img, truth) == pytest.approx(1.0, abs=2e-6)
    assert ro.detect_nanodomains(np.zeros((64, 64), dtype=np.int64)).size == 0
