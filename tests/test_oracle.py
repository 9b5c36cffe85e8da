"""CPU suite: the oracle against the reference's in-repo anchors, and the row-sweep derivation
(the schedule the CUDA kernel implements) against the sequential oracle."""
import numpy as np
import pytest

from gym_sted_b200 import base, defaults, physics
from oracle import pysted_oracle as po
from tests import helpers, rowsweep_model as rm

PX = helpers.PX

# criteria of gym_sted/routines/create.py:6-63 -> (p_ex, p_sted, pdt, target)
BLEACH_CRIT = {"low-bleach": (10e-6, 150e-3, 30e-6, 0.2), "high-bleach": (2e-6, 150e-3, 25e-6, 0.7)}
SIGNAL_TARGET = {"high-signal": 200.0, "low-signal": 30.0}


def gaussian_blob(K, sigma=2.25, scale=40.0):
    """Datamap of FluorescenceOptimizer.expected_confocal_signal (gym_sted/utils.py:646-652):
    skimage.filters.gaussian of a delta == separable Gaussian, mode 'nearest', truncate 4."""
    from scipy.ndimage import gaussian_filter
    d = np.zeros((K, K))
    d[K // 2, K // 2] = 1
    d = gaussian_filter(d, sigma, mode="nearest", truncate=4.0)
    return d / d.max() * scale


@pytest.mark.parametrize("routine", sorted(defaults.ROUTINE_FITS))
def test_routines_json_anchors(routine):
    """Loose known answers: with each fitted parameter set the reference's own closed forms
    (utils.py:599-661) must return about the target it was fitted to.  Optimiser tolerance and the
    recalled pySTED constants allow ~30 %."""
    m = helpers.oracle_microscope(defaults.load_routine(routine))
    sig_name, bl_name = routine.split("_")
    E = m.get_effective(PX, 10e-6, 0.0)
    photons = m.fluo.get_photons((E * gaussian_blob(E.shape[0])).sum())
    signal = float(m.detector.get_signal(photons, 10e-6))
    assert signal == pytest.approx(SIGNAL_TARGET[sig_name], rel=0.30)
    p_ex, p_sted, pdt, target = BLEACH_CRIT[bl_name]
    kb = m.get_k_bleach_map(PX, p_ex, p_sted, pdt)
    bleach = 1 - np.exp(-(kb * pdt).sum())                 # kb_map_to_im_bleach, utils.py:599-604
    assert bleach == pytest.approx(target, rel=0.30)


def test_product_psf_cache_matches_oracle_quadrature():
    """physics.py (Gauss-Legendre, vectorised) vs oracle (scipy.integrate.quad per pixel)."""
    ref = helpers.oracle_psf_cache()
    got = helpers.product_microscope().cache(PX)
    for a, b in zip(ref, got):
        assert a.shape == b.shape == (59, 59)
        np.testing.assert_allclose(b, a, rtol=1e-10, atol=1e-12 * np.abs(a).max())
    i_ex, i_sted, psf_det = got
    assert i_sted[29, 29] < 0.02 * i_sted.max()            # zero_residual floor only
    assert np.unravel_index(i_ex.argmax(), i_ex.shape) == (29, 29)
    assert psf_det.max() == pytest.approx(0.82)


def test_psf_sides_follow_wavelength():
    assert physics.psf_side(635e-9, 1.4, PX) == 51
    assert physics.psf_side(690e-9, 1.4, PX) == 55
    assert physics.psf_side(750e-9, 1.4, PX) == 59


def test_effective_psf_limits():
    m = helpers.oracle_microscope()
    conf = m.get_effective(PX, 10e-6, 0.0)
    sted = m.get_effective(PX, 10e-6, 150e-3)
    assert np.all(sted <= conf * (1 + 1e-12))
    # depletion leaves the centre (zero_residual 1 %) and removes the rim
    assert sted[29, 29] / conf[29, 29] > 0.3
    assert sted[29, 40] / conf[29, 40] < 0.05
    # linear in excitation power
    np.testing.assert_allclose(m.get_effective(PX, 20e-6, 0.0), 2 * conf, rtol=1e-12)


def test_datamap_roi_padding():
    m = helpers.product_microscope()
    i_ex = m.cache(PX)[0]
    for cls in (base.Datamap, po.Datamap):
        dm = cls(np.arange(12).reshape(3, 4), PX)
        dm.set_roi(i_ex, "max")
        assert dm.whole_datamap.shape == (3 + 58, 4 + 58)
        assert dm.roi == (slice(29, 32), slice(29, 33))
        np.testing.assert_array_equal(dm["base"][dm.roi], np.arange(12).reshape(3, 4))
        assert dm.whole_datamap.sum() == 66


def test_c_raster_matches_numpy_twin():
    rng = np.random.default_rng(1)
    for H, W, K in [(5, 7, 3), (8, 8, 5), (4, 9, 7)]:
        E, S = rng.random((K, K)), rng.uniform(0.5, 1, (K, K))
        D = rng.integers(0, 9, (H + K - 1, W + K - 1)).astype(np.float64)
        for mode in (0, 1):
            a, b = D.copy(), D.copy()
            ia = po.raster(a, H, W, E, S, mode)
            ib = po.raster_numpy(b, H, W, E, S, mode)
            np.testing.assert_allclose(ia, ib, rtol=1e-13)
            np.testing.assert_allclose(a, b, rtol=1e-13)


SHAPES = [(9, 11, 5), (6, 4, 7), (3, 3, 7), (12, 12, 5), (20, 17, 9), (1, 1, 3), (16, 5, 11)]


@pytest.mark.parametrize("H,W,K", SHAPES)
def test_rowsweep_expectation_equals_sequential_scan(H, W, K):
    """The per-row schedule is algebraically the reference's dwell-by-dwell scan (mean field)."""
    rng = np.random.default_rng(H * 100 + W * 10 + K)
    E, A = rng.random((K, K)), rng.random((K, K)) * 0.3
    p = K // 2
    D = np.zeros((H + K - 1, W + K - 1))
    D[p:p + H, p:p + W] = rng.integers(0, 8, (H, W)) * (rng.random((H, W)) < 0.5)
    ref_D = D.copy()
    ref_I = po.raster(ref_D, H, W, E, np.exp(-A), 1)
    I, Db = rm.sweep_expectation(D, H, W, E, A)
    np.testing.assert_allclose(I, ref_I, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(Db, ref_D, rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("variant", ["binomial_per_pass", "scheduled_deaths", "presampled"])
def test_rowsweep_stochastic_matches_sequential_moments(variant):
    """Three exact re-orderings of the reference's per-dwell per-molecule thinning — one binomial per (pixel,
    scan row); a next-death schedule; and the one the CUDA kernels use, all deaths drawn before the scan and
    replayed row by row — give the same image / datamap moments as the sequential oracle."""
    rng = np.random.default_rng(7)
    H, W, K = 6, 7, 5
    E, A = rng.random((K, K)), rng.random((K, K)) * 0.2
    p = K // 2
    D = np.zeros((H + K - 1, W + K - 1))
    D[p:p + H, p:p + W] = rng.integers(0, 30, (H, W)) * (rng.random((H, W)) < 0.6)
    fn = {"binomial_per_pass": rm.sweep_stochastic, "scheduled_deaths": rm.sweep_stochastic_scheduled,
          "presampled": rm.sweep_stochastic_presampled}[variant]
    N = 1500
    a = np.zeros((N, H, W)); b = np.zeros((N, H, W))
    da = np.zeros((N,) + D.shape); db = np.zeros((N,) + D.shape)
    for i in range(N):
        d1 = D.copy()
        a[i] = po.raster(d1, H, W, E, np.exp(-A), 2, seed=i + 1)
        da[i] = d1
        b[i], db[i] = fn(D, H, W, E, A, rng)
    # mean field is exact for both
    exp_D = D.copy()
    exp_I = po.raster(exp_D, H, W, E, np.exp(-A), 1)
    for img, dm in ((a, da), (b, db)):
        np.testing.assert_allclose(img.mean(0), exp_I, rtol=0.02)
        np.testing.assert_allclose(dm.mean(0), exp_D, atol=0.35)
    sa, sb = a.std(0), b.std(0)
    np.testing.assert_allclose(sb, sa, rtol=0.12, atol=1e-6 * sa.max())
    np.testing.assert_allclose(db.std(0), da.std(0), atol=0.25)
    # image / datamap covariance (the thinning trajectory couples them)
    ca = ((a[:, 3, 3] - a[:, 3, 3].mean()) * (da.sum((1, 2)) - da.sum((1, 2)).mean())).mean()
    cb = ((b[:, 3, 3] - b[:, 3, 3].mean()) * (db.sum((1, 2)) - db.sum((1, 2)).mean())).mean()
    assert cb == pytest.approx(ca, rel=0.25)


def test_oracle_microscope_mutates_datamap_like_pysted():
    m = helpers.oracle_microscope()
    rng = np.random.default_rng(3)
    dm = po.Datamap(helpers.synthetic_datamap(rng, 16, 16), PX)
    dm.set_roi(m.cache(PX)[0], "max")
    before = dm.whole_datamap.sum()
    conf, bl, inten = m.get_signal_and_bleach(dm, PX, pdt=10e-6, p_ex=10e-6, p_sted=0.0, bleach=False)
    assert dm.whole_datamap.sum() == before and conf.shape == (16, 16)
    sted, bl, _ = m.get_signal_and_bleach(dm, PX, pdt=50e-6, p_ex=10e-6, p_sted=0.2, bleach=True, seed=5)
    assert dm.whole_datamap.sum() < before
    assert bl["base"] is dm.whole_datamap and bl["base"].dtype == np.int64
    with pytest.raises(ValueError):
        m.get_signal_and_bleach(po.Datamap(np.ones((4, 4)), PX), PX, pdt=1e-5, p_ex=1e-5, p_sted=0)


def test_raster_tables_reduces_to_raster_and_follows_the_dwell_order():
    """oracle_raster_tables: with one table it IS oracle_raster; with per-dwell tables it equals a plain numpy
    restatement of the dwell loop (look the table up, gather, thin the window, move on)."""
    from oracle import pysted_oracle as po
    rng = np.random.default_rng(0)
    H, W, K, n = 9, 11, 5, 4
    E = rng.uniform(0, 1, (n, K, K))
    S = rng.uniform(0.5, 1, (n, K, K))
    D0 = rng.integers(0, 20, (H + K - 1, W + K - 1)).astype(np.float64)
    a, b = D0.copy(), D0.copy()
    ia = po.raster(a, H, W, E[2], S[2], 1)
    ib = po.raster_tables(b, H, W, E, S, np.full((H, W), 2), 1)
    np.testing.assert_array_equal(ia, ib)
    np.testing.assert_array_equal(a, b)
    idx = rng.integers(0, n, (H, W))
    c, d = D0.copy(), D0.copy()
    ic = po.raster_tables(c, H, W, E, S, idx, 1)
    idd = np.empty((H, W))
    for r in range(H):
        for q in range(W):
            win = d[r:r + K, q:q + K]
            idd[r, q] = (E[idx[r, q]] * win).sum()
            win *= S[idx[r, q]]
    np.testing.assert_allclose(ic, idd, rtol=1e-13)
    np.testing.assert_allclose(c, d, rtol=1e-13)
    e = D0.copy()
    po.raster_tables(e, H, W, E, S, idx, 2, seed=3)
    assert (e <= D0).all() and (e == np.floor(e)).all() and e.sum() < D0.sum()
