"""Edge cases of the acquisition path against the CPU oracle (``oracle/raster_oracle.c``): empty maps, fully occupied
maps (no row of the band can be skipped), single molecules in the corners of the ROI, one-pixel and one-line images,
and the largest molecule count per pixel the death schedule indexes.  The sparse paths of ``gather_kernel`` and
``sweep2_kernel`` (rows without molecules are not visited) must give the results of the dense loop on all of them.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import pysted_oracle as po  # noqa: E402
from tests import helpers  # noqa: E402

pytestmark = pytest.mark.gpu
PX = helpers.PX
RTOL = 1e-5


@pytest.fixture(scope="module")
def micro():
    if not torch.cuda.is_available():
        pytest.fail("these tests need a CUDA device (no CPU fallback exists)")
    return helpers.product_microscope()


@pytest.fixture(scope="module")
def omicro():
    return helpers.oracle_microscope()


def batched(micro, B, H, W, **kw):
    from gym_sted_b200.batch import BatchedMicroscope
    return BatchedMicroscope(micro, B, H, W, device="cuda:0", **kw)


def oracle_acquire(om, roi, pdt, p_ex, p_sted, mode):
    dm = po.Datamap(roi, PX)
    dm.set_roi(om.cache(PX)[0], "max")
    E = om.get_effective(PX, p_ex, p_sted)
    kb = om.get_k_bleach_map(PX, p_ex, p_sted, pdt)
    work = dm.whole_datamap.astype(np.float64).copy()
    inten = po.raster(work, roi.shape[0], roi.shape[1], E, np.exp(-kb * pdt), mode, 0)
    return inten, work[dm.roi]


def edge_maps(H, W):
    maps = np.zeros((6, H, W), dtype=np.int64)
    maps[1] = 3                                        # every pixel occupied: nothing to skip
    maps[2, 0, 0] = 1                                  # lone molecules in the four corners, one map each side
    maps[3, H - 1, W - 1] = 150
    maps[4, 0, W - 1] = 7
    maps[4, H - 1, 0] = 9
    maps[5, H // 2] = 2                                # one occupied row, the rest of the band empty
    return maps


ACT_CONF = (10e-6, 10e-6, 0.0)
ACT_STED = (30e-6, 10e-6, 0.15)


@pytest.mark.parametrize("H,W", [(72, 72), (40, 150)])
def test_gather_on_empty_full_and_corner_maps(micro, omicro, H, W):
    maps = edge_maps(H, W)
    bm = batched(micro, len(maps), H, W)
    bm.set_datamaps(torch.from_numpy(maps))
    for act in (ACT_CONF, ACT_STED):
        _, inten = bm.get_signal_and_bleach(*act, bleach=False, want_intensity=True, noise=False)
        inten = inten.cpu().numpy()
        assert not inten[0].any()                      # no molecules: exactly zero, not "almost"
        for b in range(1, len(maps)):
            ref, _ = oracle_acquire(omicro, maps[b], *act, 0)
            np.testing.assert_allclose(inten[b], ref, rtol=RTOL, atol=1e-7 * ref.max())


@pytest.mark.parametrize("H,W", [(72, 72), (40, 150)])
def test_expectation_scan_on_empty_full_and_corner_maps(micro, omicro, H, W):
    maps = edge_maps(H, W)
    bm = batched(micro, len(maps), H, W)
    bm.set_datamaps(torch.from_numpy(maps))
    _, inten = bm.get_signal_and_bleach(*ACT_STED, bleach=True, sample_bleach=False, want_intensity=True, noise=False)
    inten, bleached = inten.cpu().numpy(), bm.datamaps_roi().cpu().numpy()
    assert not inten[0].any() and not bleached[0].any()
    for b in range(1, len(maps)):
        ref_i, ref_d = oracle_acquire(omicro, maps[b], *ACT_STED, 1)
        np.testing.assert_allclose(inten[b], ref_i, rtol=RTOL, atol=1e-7 * ref_i.max())
        np.testing.assert_allclose(bleached[b], ref_d, rtol=RTOL, atol=1e-6)


@pytest.mark.parametrize("H,W", [(1, 1), (1, 200), (200, 1), (3, 131), (2, 257)])
def test_one_pixel_and_one_line_images(micro, omicro, H, W):
    rng = np.random.default_rng(H * 1000 + W)
    maps = (rng.integers(0, 6, (3, H, W)) * (rng.random((3, H, W)) < 0.6)).astype(np.int64)
    maps[0].flat[0] = 4                                # never all-empty
    bm = batched(micro, 3, H, W)
    bm.set_datamaps(torch.from_numpy(maps))
    _, inten = bm.get_signal_and_bleach(*ACT_CONF, bleach=False, want_intensity=True, noise=False)
    inten = inten.cpu().numpy()
    for b in range(3):
        ref, _ = oracle_acquire(omicro, maps[b], *ACT_CONF, 0)
        np.testing.assert_allclose(inten[b], ref, rtol=RTOL, atol=1e-7 * max(ref.max(), 1e-30))
    _, inten = bm.get_signal_and_bleach(*ACT_STED, bleach=True, sample_bleach=False, want_intensity=True, noise=False)
    inten, bleached = inten.cpu().numpy(), bm.datamaps_roi().cpu().numpy()
    for b in range(3):
        ref_i, ref_d = oracle_acquire(omicro, maps[b], *ACT_STED, 1)
        np.testing.assert_allclose(inten[b], ref_i, rtol=RTOL, atol=1e-7 * max(ref_i.max(), 1e-30))
        np.testing.assert_allclose(bleached[b], ref_d, rtol=RTOL, atol=1e-6)


def test_stochastic_scan_on_empty_and_corner_maps(micro, omicro):
    """Deaths drawn for maps with no molecule, with one molecule per corner and with the largest count the schedule
    indexes (32 767): nothing appears from nothing, nothing is lost outside the law Binomial(n, exp(-sum k_b pdt))."""
    H = W = 72
    maps = edge_maps(H, W)
    maps[3, H - 1, W - 1] = 32767
    R = 64
    act = (60e-6, 10e-6, 0.3)
    stack = np.repeat(maps[None], R, axis=0).reshape(-1, H, W)
    bm = batched(micro, len(stack), H, W, seed=77)
    bm.set_datamaps(torch.from_numpy(stack))
    _, inten = bm.get_signal_and_bleach(*act, bleach=True, sample_bleach=True, want_intensity=True, noise=False)
    assert bm.dropped_events() == 0
    inten = inten.cpu().numpy().reshape(R, len(maps), H, W)
    after = bm.datamaps_roi().cpu().numpy().reshape(R, len(maps), H, W)
    assert not inten[:, 0].any() and not after[:, 0].any()
    assert (after <= maps[None]).all() and (after >= 0).all()
    assert (after == np.rint(after)).all()
    for b in (1, 3, 5):
        _, exp_d = oracle_acquire(omicro, maps[b], *act, 1)            # n * survival of the whole scan, per pixel
        n = maps[b].astype(np.float64)
        occ = n > 0
        p = np.where(occ, exp_d / np.maximum(n, 1), 0.0)
        mean = after[:, b].mean(axis=0)
        se = np.sqrt(np.maximum(n * p * (1 - p), 1e-12) / R)
        z = (mean - exp_d)[occ] / se[occ]
        assert np.abs(z).max() < 6.0, (b, np.abs(z).max())
        if occ.sum() > 100:
            assert abs(z.mean()) < 0.2 and 0.8 < z.std() < 1.2
    # an image pixel sees a molecule only before the scan has bleached it: the first dwell of the lone corner
    # molecule of map 2 is the oracle's, whatever was drawn afterwards
    ref_i, _ = oracle_acquire(omicro, maps[2], *act, 0)
    np.testing.assert_allclose(inten[:, 2, 0, 0], ref_i[0, 0], rtol=RTOL)
