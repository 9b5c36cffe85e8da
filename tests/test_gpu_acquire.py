"""GPU parity tests of the acquisition path — CUDA kernels (through the C ABI) vs the CPU oracle.

Tolerances: floating-point fields (effective PSF, expected photons, expected bleached maps) within
1e-5 relative as north_star states; stochastic mode by per-pixel moments and KS tests.
"""
import ctypes

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from gym_sted_b200 import _lib, base, defaults  # noqa: E402
from oracle import pysted_oracle as po  # noqa: E402
from tests import helpers, rowsweep_model as rm  # noqa: E402

pytestmark = pytest.mark.gpu
PX = helpers.PX
RTOL = 1e-5


def _require_gpu():
    if not torch.cuda.is_available():
        pytest.fail("these tests need a CUDA device (no CPU fallback exists)")


@pytest.fixture(scope="module")
def micro():
    _require_gpu()
    return helpers.product_microscope()


@pytest.fixture(scope="module")
def omicro():
    return helpers.oracle_microscope()


def batched(micro, B, H, W, **kw):
    from gym_sted_b200.batch import BatchedMicroscope
    return BatchedMicroscope(micro, B, H, W, device="cuda:0", **kw)


def oracle_acquire(om, roi, pdt, p_ex, p_sted, mode, seed=0):
    """(intensity, bleached padded map) from the sequential float64 oracle."""
    dm = po.Datamap(roi, PX)
    dm.set_roi(om.cache(PX)[0], "max")
    E = om.get_effective(PX, p_ex, p_sted)
    S = np.exp(-om.get_k_bleach_map(PX, p_ex, p_sted, pdt) * pdt)
    work = dm.whole_datamap.astype(np.float64).copy()
    inten = po.raster(work, roi.shape[0], roi.shape[1], E, S, mode, seed)
    return inten, work[dm.roi]


ACTIONS = [(10e-6, 10e-6, 0.0), (30e-6, 10e-6, 0.15), (100e-6, 20e-6, 0.35), (1e-6, 2e-6, 0.05)]  # pdt, p_ex, p_sted


# ------------------------------------------------------------------ K1: get_effective / get_k_bleach
def test_make_effective_tables_match_oracle(micro, omicro):
    B = len(ACTIONS)
    bm = batched(micro, B, 8, 8)
    pdt, p_ex, p_sted = (torch.tensor([a[i] for a in ACTIONS], dtype=torch.float64, device="cuda") for i in range(3))
    kb = bm.make_effective(p_ex, p_sted, pdt, want_kbleach=True).cpu().numpy()
    tab = bm.tables.cpu().numpy()
    K = bm.K
    assert tab.shape == (B, _lib.STED_TABLES, K, 60) and np.all(tab[:, :3, :, K:] == 0)
    for b, (dt, pe, ps) in enumerate(ACTIONS):
        E = omicro.get_effective(PX, pe, ps)
        kref = omicro.get_k_bleach_map(PX, pe, ps, dt)
        np.testing.assert_allclose(tab[b, _lib.STED_TAB_E, :, :K], E, rtol=2e-7, atol=0)
        np.testing.assert_allclose(kb[b], kref, rtol=1e-11, atol=0)
        ET, CH = rm.tables(E, kref * dt)
        np.testing.assert_allclose(tab[b, _lib.STED_TAB_ET, :, :K], ET, rtol=3e-7)
        np.testing.assert_allclose(tab[b, _lib.STED_TAB_CH, :, :K], CH[:, :K], rtol=3e-7)
        np.testing.assert_allclose(tab[b, _lib.STED_TAB_RS, :, :K], np.exp(-CH[:, :K]), rtol=3e-7)
        # row-interleaved copy for the FFMA2 gather: E2[u][v] = (E[u][v], E[u-1][v]), u = 0..K, zero outside
        e2 = tab[b, _lib.STED_TAB_E2:].reshape(-1)[:(K + 1) * 60 * 2].reshape(K + 1, 60, 2)
        e_pad = np.zeros((K + 2, 60), dtype=np.float32)
        e_pad[1:K + 1] = tab[b, _lib.STED_TAB_E]
        assert np.array_equal(e2[..., 0], e_pad[1:]) and np.array_equal(e2[..., 1], e_pad[:-1])


def test_get_effective_facade(micro, omicro):
    got = micro.get_effective(PX, 10e-6, 0.1)
    np.testing.assert_allclose(got, omicro.get_effective(PX, 10e-6, 0.1), rtol=2e-7)
    np.testing.assert_allclose(micro.get_k_bleach(PX, 10e-6, 0.1, 20e-6),
                               omicro.get_k_bleach_map(PX, 10e-6, 0.1, 20e-6), rtol=1e-11)


# ------------------------------------------------------------------ K2: gather (bleach=False)
@pytest.mark.parametrize("H,W", [(64, 64), (80, 72), (33, 130), (1, 1)])
def test_gather_matches_oracle(micro, omicro, H, W):
    rng = np.random.default_rng(H * 1000 + W)
    B = 3
    rois = np.stack([helpers.synthetic_datamap(rng, H, W) if min(H, W) > 8 else rng.integers(0, 9, (H, W))
                     for _ in range(B)])
    rois[1] = rng.integers(0, 200, (H, W))          # dense map
    bm = batched(micro, B, H, W)
    bm.set_datamaps(torch.from_numpy(rois))
    acts = ACTIONS[:B]
    pdt, p_ex, p_sted = (torch.tensor([a[i] for a in acts], dtype=torch.float64) for i in range(3))
    sig, inten = bm.get_signal_and_bleach(pdt, p_ex, p_sted, bleach=False, want_intensity=True, noise=False)
    sig, inten = sig.cpu().numpy(), inten.cpu().numpy()
    det = micro.detector
    for b, (dt, pe, ps) in enumerate(acts):
        ref, _ = oracle_acquire(omicro, rois[b], dt, pe, ps, 0)
        np.testing.assert_allclose(inten[b], ref, rtol=RTOL, atol=1e-7 * ref.max())
        mean = omicro.fluo.get_photons(ref) * det.pcef * det.pdef * dt + det.background * dt
        np.testing.assert_allclose(sig[b], mean, rtol=RTOL, atol=1e-6 * max(mean.max(), 1e-30))
    assert torch.equal(bm.datamaps_roi().cpu(), torch.from_numpy(rois).float())   # untouched


def test_gather_generic_psf_size(micro):
    """K != 59 goes through the runtime-K instantiation; random dense data, random tables."""
    _require_gpu()
    lib = _lib.load()
    rng = np.random.default_rng(5)
    for K, H, W in [(7, 20, 31), (11, 70, 66), (3, 5, 4)]:
        KP = _lib.kpitch(K)
        B = 2
        Hp, Wp = H + K - 1, W + K - 1
        Wpitch = (Wp + 3) & ~3
        E = rng.random((B, K, K))
        D = np.zeros((B, Hp, Wpitch), dtype=np.float32)
        D[:, :, :Wp] = rng.integers(0, 50, (B, Hp, Wp))
        tab = helpers.empty_tables(B, K)
        tab[:, 0, :, :K] = E
        helpers.fill_e2(tab)
        d_D, d_tab = torch.from_numpy(D).cuda(), torch.from_numpy(tab).cuda()
        d_pdt = torch.full((B,), 1e-5, dtype=torch.float64, device="cuda")
        d_int = torch.empty((B, H, W), dtype=torch.float32, device="cuda")
        d_sig = torch.empty_like(d_int)
        det = _lib.StedDetectorParams(0.1, 0.5, 0.0)
        _lib.check(lib.sted_acquire(d_D.data_ptr(), None, d_tab.data_ptr(), d_pdt.data_ptr(), B, H, W, K, 0,
                                    ctypes.byref(det), 690e-9, 1, 1, 0, d_int.data_ptr(), d_sig.data_ptr(), None, 0, None))
        got = d_int.cpu().numpy()
        for b in range(B):
            work = D[b, :, :Wp].astype(np.float64).copy()
            ref = po.raster(work, H, W, tab[b, 0, :, :K].astype(np.float64), np.ones((K, K)), 0)
            np.testing.assert_allclose(got[b], ref, rtol=RTOL)


# ------------------------------------------------------------------ K3: scan with bleaching, expectation mode
@pytest.mark.parametrize("H,W", [(64, 64), (40, 96), (48, 160), (70, 30)])
def test_sweep_expectation_matches_oracle(micro, omicro, H, W):
    rng = np.random.default_rng(H * 7 + W)
    acts = ACTIONS[1:4]
    B = len(acts)
    rois = np.stack([helpers.synthetic_datamap(rng, H, W) for _ in range(B)])
    rois[2] = rng.integers(0, 60, (H, W))
    bm = batched(micro, B, H, W)
    bm.set_datamaps(torch.from_numpy(rois))
    pdt, p_ex, p_sted = (torch.tensor([a[i] for a in acts], dtype=torch.float64) for i in range(3))
    _, inten = bm.get_signal_and_bleach(pdt, p_ex, p_sted, bleach=True, sample_bleach=False, want_intensity=True,
                                        noise=False)
    inten = inten.cpu().numpy()
    bleached = bm.datamaps_roi().cpu().numpy()
    whole = bm.dmap.cpu().numpy()
    for b, (dt, pe, ps) in enumerate(acts):
        ref_i, ref_d = oracle_acquire(omicro, rois[b], dt, pe, ps, 1)
        np.testing.assert_allclose(inten[b], ref_i, rtol=RTOL, atol=1e-7 * ref_i.max())
        np.testing.assert_allclose(bleached[b], ref_d, rtol=RTOL, atol=1e-6)
        if ps >= 0.1:
            assert ref_d.sum() < rois[b].sum() * 0.999        # the case does bleach
    p = bm.p
    frame = whole.copy()
    frame[:, p:p + H, p:p + W] = 0
    assert not frame.any()                                     # zero frame preserved


def test_sweep_generic_psf_edge_shapes(micro):
    """Runtime-K sweep on shapes smaller than the PSF, one-pixel scans, ragged widths."""
    _require_gpu()
    lib = _lib.load()
    rng = np.random.default_rng(11)
    for K, H, W in [(7, 9, 11), (7, 6, 4), (7, 3, 3), (5, 1, 1), (9, 20, 17), (11, 16, 70), (3, 1, 9)]:
        KP = _lib.kpitch(K)
        B = 2
        Hp, Wp = H + K - 1, W + K - 1
        Wpitch = (Wp + 3) & ~3
        p = K // 2
        D = np.zeros((B, Hp, Wpitch), dtype=np.float32)
        D[:, p:p + H, p:p + W] = rng.integers(0, 9, (B, H, W)) * (rng.random((B, H, W)) < 0.6)
        tab = helpers.empty_tables(B, K)
        tab[:, 3] = 1.0
        Es, As = rng.random((B, K, K)), rng.random((B, K, K)) * 0.3
        for b in range(B):
            ET, CH = rm.tables(Es[b], As[b])
            tab[b, 0, :, :K], tab[b, 1, :, :K] = Es[b], ET
            tab[b, 2, :, :K], tab[b, 3, :, :K] = CH[:, :K], np.exp(-CH[:, :K])
        helpers.fill_e2(tab)
        d_D, d_tab = torch.from_numpy(D).cuda(), torch.from_numpy(tab).cuda()
        d_out = torch.full_like(d_D, -1.0)
        d_pdt = torch.full((B,), 1e-5, dtype=torch.float64, device="cuda")
        d_int = torch.empty((B, H, W), dtype=torch.float32, device="cuda")
        d_sig = torch.empty_like(d_int)
        det = _lib.StedDetectorParams(0.1, 0.5, 0.0)
        _lib.check(lib.sted_acquire(d_D.data_ptr(), d_out.data_ptr(), d_tab.data_ptr(), d_pdt.data_ptr(), B, H, W, K,
                                    _lib.STED_BLEACH, ctypes.byref(det), 690e-9, 1, 1, 0, d_int.data_ptr(),
                                    d_sig.data_ptr(), None, 0, None))
        got_i, got_d = d_int.cpu().numpy(), d_out.cpu().numpy()
        for b in range(B):
            work = D[b, :, :Wp].astype(np.float64).copy()
            ref = po.raster(work, H, W, Es[b], np.exp(-As[b]), 1)
            np.testing.assert_allclose(got_i[b], ref, rtol=2e-5, atol=1e-6, err_msg=f"K={K} H={H} W={W}")
            np.testing.assert_allclose(got_d[b, :, :Wp], work, rtol=2e-5, atol=1e-6, err_msg=f"K={K} H={H} W={W}")


# ------------------------------------------------------------------ stochastic mode
def test_sweep_stochastic_moments_and_ks(micro, omicro):
    """256 replicas of one environment: per-pixel mean equals the exact mean field, spread and
    distributions match the reference-order per-molecule thinning (oracle mode 2)."""
    from scipy import stats
    H = W = 24
    rng = np.random.default_rng(2)
    roi = helpers.synthetic_datamap(rng, H, W, n_domains=4)
    R = 256
    dt, pe, ps = 30e-6, 10e-6, 0.15
    bm = batched(micro, R, H, W, seed=99)
    bm.set_datamaps(torch.from_numpy(np.broadcast_to(roi, (R, H, W)).copy()))
    _, inten = bm.get_signal_and_bleach(dt, pe, ps, bleach=True, sample_bleach=True, want_intensity=True, noise=False)
    gi = inten.cpu().numpy().astype(np.float64)
    gd = bm.datamaps_roi().cpu().numpy().astype(np.float64)
    assert np.all(gd == np.rint(gd)) and np.all(gd <= roi) and np.all(gd >= 0)
    exp_i, exp_d = oracle_acquire(omicro, roi, dt, pe, ps, 1)
    oi = np.zeros((R, H, W)); od = np.zeros((R, H, W))
    for r in range(R):
        oi[r], od[r] = oracle_acquire(omicro, roi, dt, pe, ps, 2, seed=1000 + r)
    assert 0.2 < exp_d.sum() / roi.sum() < 0.95               # substantial but partial bleaching
    # means against the exact expectation (standard error of the mean from the oracle's spread)
    se_i = oi.std(0) / np.sqrt(R) + 1e-9 * exp_i.max()
    assert np.all(np.abs(gi.mean(0) - exp_i) < 5 * se_i + 1e-5 * exp_i)
    se_d = od.std(0) / np.sqrt(R) + 1e-9
    assert np.all(np.abs(gd.mean(0) - exp_d) < 5 * se_d + 1e-5 * exp_d + 1e-6)
    # spreads
    bright = exp_i > 0.2 * exp_i.max()
    np.testing.assert_allclose(gi.std(0)[bright], oi.std(0)[bright], rtol=0.25)
    occupied = roi > 0
    np.testing.assert_allclose(gd.std(0)[occupied], od.std(0)[occupied], rtol=0.3, atol=0.15)
    # KS on the brightest image pixel, on the total surviving molecules and on every nanodomain pixel
    y, x = np.unravel_index(exp_i.argmax(), exp_i.shape)
    assert stats.ks_2samp(gi[:, y, x], oi[:, y, x]).pvalue > 1e-3
    assert stats.ks_2samp(gd.sum((1, 2)), od.sum((1, 2))).pvalue > 1e-3
    for (yy, xx) in zip(*np.nonzero(roi > 50)):
        assert stats.ks_2samp(gd[:, yy, xx], od[:, yy, xx]).pvalue > 1e-3


def test_stochastic_results_do_not_depend_on_sharding(micro):
    """Draws are keyed by (seed, offset, global env id, pixel): B=8 in one launch == two launches of 4."""
    H = W = 40
    rng = np.random.default_rng(4)
    rois = np.stack([helpers.synthetic_datamap(rng, H, W) for _ in range(8)])
    m = helpers.product_microscope(noise=True, background=defaults.DETECTOR["background"])
    full = batched(m, 8, H, W, seed=7)
    full.set_datamaps(torch.from_numpy(rois))
    s_full, _ = full.get_signal_and_bleach(40e-6, 10e-6, 0.2, bleach=True)
    d_full = full.datamaps_roi().clone()
    for half in range(2):
        part = batched(m, 4, H, W, seed=7, env_base=4 * half)
        part.set_datamaps(torch.from_numpy(rois[4 * half:4 * half + 4]))
        s, _ = part.get_signal_and_bleach(40e-6, 10e-6, 0.2, bleach=True)
        assert torch.equal(s, s_full[4 * half:4 * half + 4])
        assert torch.equal(part.datamaps_roi(), d_full[4 * half:4 * half + 4])
    assert (d_full.sum() < torch.from_numpy(rois).sum())


# ------------------------------------------------------------------ samplers
@pytest.mark.parametrize("lam", [0.05, 0.9, 3.0, 11.9, 12.1, 47.0, 1.0e3, 2.5e5])
def test_poisson_sampler_distribution(lam):
    from scipy import stats
    _require_gpu()
    n = 400_000
    lib = _lib.load()
    d_lam = torch.full((n,), lam, dtype=torch.float32, device="cuda")
    d_out = torch.empty_like(d_lam)
    _lib.check(lib.sted_sample_poisson(d_lam.data_ptr(), n, 12345, 3, d_out.data_ptr(), None))
    x = d_out.cpu().numpy().astype(np.float64)
    assert np.all(x == np.rint(x)) and x.min() >= 0
    assert abs(x.mean() - lam) < 5 * np.sqrt(lam / n)
    assert abs(x.var() - lam) < 6 * lam * np.sqrt(2.0 / n) + 6 * np.sqrt(lam / n)
    if lam < 100:
        ks = np.arange(0, int(lam + 12 * np.sqrt(lam) + 12))
        obs = np.bincount(x.astype(np.int64), minlength=len(ks) + 1)
        expc = stats.poisson.pmf(ks, lam) * n
        keep = expc > 5
        chi2 = ((obs[:len(ks)][keep] - expc[keep]) ** 2 / expc[keep]).sum()
        assert chi2 < stats.chi2.ppf(1 - 1e-6, keep.sum())
    else:
        assert stats.kstest(x + np.random.default_rng(0).uniform(-0.5, 0.5, n), stats.norm(lam, np.sqrt(lam)).cdf).statistic < 0.01


@pytest.mark.parametrize("nmol,q", [(5, 1e-4), (5, 0.02), (150, 0.03), (150, 0.7), (1000, 0.5), (3, 0.999), (64, 0.25),
                                    (65, 0.25), (1, 0.5)])
def test_binomial_sampler_distribution(nmol, q):
    from scipy import stats
    _require_gpu()
    n = 400_000
    lib = _lib.load()
    d_n = torch.full((n,), float(nmol), dtype=torch.float32, device="cuda")
    d_q = torch.full((n,), q, dtype=torch.float32, device="cuda")
    d_out = torch.empty_like(d_n)
    _lib.check(lib.sted_sample_binomial(d_n.data_ptr(), d_q.data_ptr(), n, 777, 1, d_out.data_ptr(), None))
    x = d_out.cpu().numpy().astype(np.int64)
    assert x.min() >= 0 and x.max() <= nmol
    ks = np.arange(nmol + 1)
    expc = stats.binom.pmf(ks, nmol, float(np.float32(q))) * n
    obs = np.bincount(x, minlength=nmol + 1)
    keep = expc > 5
    chi2 = ((obs[keep] - expc[keep]) ** 2 / expc[keep]).sum()
    assert chi2 < stats.chi2.ppf(1 - 1e-6, max(keep.sum(), 1))
    assert abs(x.mean() - nmol * q) < 5 * np.sqrt(nmol * q * (1 - q) / n) + 1e-7 * nmol


def test_detector_noise_moments(micro):
    """Poisson detection: per-pixel mean/variance over replicas equal the noise-free signal (+ background)."""
    m = helpers.product_microscope(noise=True, background=defaults.DETECTOR["background"])
    H = W = 32
    rng = np.random.default_rng(8)
    roi = helpers.synthetic_datamap(rng, H, W)
    R = 512
    bm = batched(m, R, H, W, seed=5)
    bm.set_datamaps(torch.from_numpy(np.broadcast_to(roi, (R, H, W)).copy()))
    noisy, _ = bm.get_signal_and_bleach(10e-6, 10e-6, 0.0, bleach=False)
    clean, _ = bm.get_signal_and_bleach(10e-6, 10e-6, 0.0, bleach=False, noise=False)
    noisy, clean = noisy.cpu().numpy().astype(np.float64), clean.cpu().numpy().astype(np.float64)[0]
    assert np.all(noisy == np.rint(noisy))
    assert clean.min() >= 0.5 / 3 * (1 - 1e-6)                    # background 0.5/(3*PDT) * PDT everywhere
    assert np.all(np.abs(noisy.mean(0) - clean) < 5.5 * np.sqrt(clean / R))
    ratio = noisy.var(0) / clean
    assert abs(ratio.mean() - 1) < 0.02 and ratio.min() > 0.6 and ratio.max() < 1.5


# ------------------------------------------------------------------ reference-facing call (host buffers)
def test_microscope_facade_matches_oracle_and_mutates_datamap(omicro):
    _require_gpu()
    m = helpers.product_microscope()
    rng = np.random.default_rng(12)
    roi = helpers.synthetic_datamap(rng, 64, 64)
    dm = base.Datamap(roi, PX)
    dm.set_roi(m.cache(PX)[0], "max")
    odm = po.Datamap(roi, PX)
    odm.set_roi(omicro.cache(PX)[0], "max")
    conf = dict(pdt=10e-6, p_ex=10e-6, p_sted=0.0)
    c1, bl, i1 = m.get_signal_and_bleach(dm, PX, **conf, bleach=False)
    oc1, _, oi1 = omicro.get_signal_and_bleach(odm, PX, **conf, bleach=False)
    np.testing.assert_allclose(i1, oi1, rtol=RTOL, atol=1e-7 * oi1.max())
    np.testing.assert_allclose(c1, oc1, rtol=RTOL, atol=1e-6)
    assert bl["base"].shape == (122, 122) and np.array_equal(bl["base"], dm.whole_datamap)
    sted = dict(pdt=30e-6, p_ex=10e-6, p_sted=0.15)
    s, bl, i2 = m.get_signal_and_bleach(dm, PX, **sted, bleach=True, sample_bleach=False)
    os_, obl, oi2 = omicro.get_signal_and_bleach(odm, PX, **sted, bleach=True, sample_bleach=False)
    np.testing.assert_allclose(i2, oi2, rtol=RTOL, atol=1e-7 * oi2.max())
    np.testing.assert_allclose(bl["base"], obl["base"], rtol=RTOL, atol=1e-6)
    assert bl["base"] is dm.whole_datamap                      # mutated in place like pySTED (update=True)
    c2, _, i3 = m.get_signal_and_bleach(dm, PX, **conf, bleach=False)
    _, _, oi3 = omicro.get_signal_and_bleach(odm, PX, **conf, bleach=False)
    np.testing.assert_allclose(i3, oi3, rtol=2 * RTOL, atol=1e-7 * oi3.max())
    assert i3.sum() < i1.sum()
    # stochastic call: integer outputs, fewer molecules, update=False leaves the datamap alone
    before = dm.whole_datamap.copy()
    m2 = helpers.product_microscope(noise=True, background=defaults.DETECTOR["background"])
    dm2 = base.Datamap(roi, PX)
    dm2.set_roi(m2.cache(PX)[0], "max")
    ph, bl2, _ = m2.get_signal_and_bleach(dm2, PX, **sted, bleach=True, update=False, seed=3)
    assert ph.dtype == np.int64 and bl2["base"].dtype == np.int64
    assert bl2["base"].sum() < roi.sum() and np.array_equal(dm2.whole_datamap[dm2.roi], roi)
    assert np.array_equal(dm.whole_datamap, before)
    with pytest.raises(ValueError):
        m.get_signal_and_bleach(base.Datamap(roi, PX), PX, **conf)


# ------------------------------------------------------------------ full-size properties (BASELINE sizes)
def test_full_size_properties(micro):
    """256x256, K=59: linearity of the gather, monotone bleaching, integer thinning — properties that
    need no CPU oracle at this size."""
    H = W = 256
    B = 4
    from gym_sted_b200 import synapse
    maps, _ = synapse.generate_batch(2 * B, H, W, seed=3)
    bm = batched(micro, B, H, W, seed=1)
    a, b = torch.from_numpy(maps[:B]).float(), torch.from_numpy(maps[B:]).float()
    out = []
    for roi in (a, b, a + b):
        bm.set_datamaps(roi)
        _, i = bm.get_signal_and_bleach(10e-6, 10e-6, 0.0, bleach=False, want_intensity=True, noise=False)
        out.append(i.clone())
    torch.testing.assert_close(out[0] + out[1], out[2], rtol=1e-5, atol=1e-7 * float(out[2].max()))
    bm.set_datamaps(a)
    _, i_exp = bm.get_signal_and_bleach(50e-6, 10e-6, 0.25, bleach=True, sample_bleach=False, want_intensity=True,
                                        noise=False)
    d_exp = bm.datamaps_roi().clone()
    assert torch.all(d_exp <= a.cuda() + 1e-4) and d_exp.sum() < 0.9 * a.sum()
    bm.set_datamaps(a)
    _, i_st = bm.get_signal_and_bleach(50e-6, 10e-6, 0.25, bleach=True, sample_bleach=True, want_intensity=True,
                                       noise=False)
    d_st = bm.datamaps_roi()
    assert torch.all(d_st == torch.round(d_st)) and torch.all(d_st <= a.cuda()) and torch.all(d_st >= 0)
    # totals agree with the mean field within a few sigma of binomial noise
    tot_exp, tot_st = float(d_exp.sum()), float(d_st.sum())
    assert abs(tot_exp - tot_st) < 6 * np.sqrt(float(a.sum()))
    assert abs(float(i_st.sum()) / float(i_exp.sum()) - 1) < 0.01


@pytest.mark.parametrize("B", [5, 40, 131])
def test_host_api_equals_device_api_for_every_chunking(B):
    """``sted_acquire_host`` cuts the batch in 1/2/4/8 chunks on separate streams; with the same seed, offset and
    env_base it must return exactly what the device-resident call returns (Philox is keyed by the global environment
    id, the scan workspace is per chunk): photons, intensity and bleached maps, stochastic mode included."""
    _require_gpu()
    import ctypes
    from gym_sted_b200.base import fluo_struct
    H, W = 48, 80
    m = helpers.product_microscope(defaults.load_routine("high-signal_low-bleach"), noise=True,
                                   background=defaults.DETECTOR["background"])
    rng = np.random.default_rng(B)
    maps = np.stack([helpers.synthetic_datamap(rng, H, W) for _ in range(B)]).astype(np.float32)
    pdt, p_ex, p_sted = rng.uniform(5e-6, 40e-6, B), rng.uniform(2e-6, 15e-6, B), rng.uniform(0.02, 0.25, B)
    lib = _lib.load()
    cache = m._packed_cache(PX)
    K = cache.shape[-1]
    fl = (_lib.StedFluoParams * B)(*[fluo_struct(m.fluo, m.excitation, m.sted)] * B)
    det = m.detector.params()
    vp = ctypes.c_void_p
    flags = _lib.STED_BLEACH | _lib.STED_SAMPLE_BLEACH | _lib.STED_SAMPLE_PHOTONS
    h_map = maps.copy()
    h_int, h_sig = np.empty_like(maps), np.empty_like(maps)
    _lib.check(lib.sted_acquire_host(h_map.ctypes.data_as(vp), cache.ctypes.data_as(vp), ctypes.cast(fl, vp),
                                     p_ex.ctypes.data_as(vp), p_sted.ctypes.data_as(vp), pdt.ctypes.data_as(vp), B, H, W, K,
                                     flags, ctypes.byref(det), 77, 5, 1000, h_int.ctypes.data_as(vp), h_sig.ctypes.data_as(vp)))
    bm = batched(m, B, H, W, pixelsize=PX, env_base=1000, seed=77)
    bm.set_datamaps(torch.from_numpy(maps))
    bm.offset = 4                                        # the call increments it to 5
    sig, inten = bm.get_signal_and_bleach(pdt, p_ex, p_sted, bleach=True, want_intensity=True)
    assert np.array_equal(h_sig, sig.cpu().numpy())
    assert np.array_equal(h_int, inten.cpu().numpy())
    assert np.array_equal(h_map, bm.datamaps_roi().cpu().numpy())
    assert h_map.sum() < maps.sum() and h_sig.max() > 5
    # uint16 transport (STED_HOST_U16): same counts, half the bytes over the bus
    m16, s16 = maps.astype(np.uint16), np.empty(maps.shape, dtype=np.uint16)
    _lib.check(lib.sted_acquire_host(m16.ctypes.data_as(vp), cache.ctypes.data_as(vp), ctypes.cast(fl, vp),
                                     p_ex.ctypes.data_as(vp), p_sted.ctypes.data_as(vp), pdt.ctypes.data_as(vp), B, H, W, K,
                                     flags | _lib.STED_HOST_U16, ctypes.byref(det), 77, 5, 1000, None, s16.ctypes.data_as(vp)))
    assert np.array_equal(m16, h_map.astype(np.uint16)) and np.array_equal(s16, h_sig.astype(np.uint16))
    # fractional outputs cannot travel as uint16
    rc = lib.sted_acquire_host(m16.ctypes.data_as(vp), cache.ctypes.data_as(vp), ctypes.cast(fl, vp), p_ex.ctypes.data_as(vp),
                               p_sted.ctypes.data_as(vp), pdt.ctypes.data_as(vp), B, H, W, K,
                               _lib.STED_BLEACH | _lib.STED_SAMPLE_PHOTONS | _lib.STED_HOST_U16, ctypes.byref(det), 77, 5, 1000,
                               None, s16.ctypes.data_as(vp))
    assert rc == -1 and b"STED_HOST_U16" in lib.sted_last_error()            # STED_EINVAL


def test_host_api_calls_from_two_threads_run_concurrently_and_agree():
    """``sted_acquire_host`` keeps its scratch and streams per host thread: two threads (one on low-priority streams)
    calling it at the same time get exactly the results of the same calls made one after the other."""
    _require_gpu()
    import ctypes
    import threading
    from gym_sted_b200.base import fluo_struct
    B, H, W = 48, 64, 64
    m = helpers.product_microscope(defaults.load_routine("high-signal_low-bleach"), noise=True,
                                   background=defaults.DETECTOR["background"])
    rng = np.random.default_rng(5)
    maps = [np.stack([helpers.synthetic_datamap(rng, H, W) for _ in range(B)]).astype(np.float32) for _ in range(2)]
    lib = _lib.load()
    cache = m._packed_cache(PX)
    K = cache.shape[-1]
    fl = (_lib.StedFluoParams * B)(*[fluo_struct(m.fluo, m.excitation, m.sted)] * B)
    det = m.detector.params()
    vp = ctypes.c_void_p
    pdt, p_ex, p_sted = np.full(B, 20e-6), np.full(B, 8e-6), np.full(B, 0.12)
    flags = _lib.STED_BLEACH | _lib.STED_SAMPLE_BLEACH | _lib.STED_SAMPLE_PHOTONS

    def call(h_map, h_sig, off):
        _lib.check(lib.sted_acquire_host(h_map.ctypes.data_as(vp), cache.ctypes.data_as(vp), ctypes.cast(fl, vp),
                                         p_ex.ctypes.data_as(vp), p_sted.ctypes.data_as(vp), pdt.ctypes.data_as(vp), B, H, W, K,
                                         flags, ctypes.byref(det), 7, off, 0, None, h_sig.ctypes.data_as(vp)))

    ref = []
    for k in range(2):
        hm, hs = maps[k].copy(), np.empty_like(maps[k])
        for rep in range(3):
            call(hm, hs, 10 * k + rep)
        ref.append((hm, hs))
    got = [None, None]
    errors = []

    def worker(k):
        try:
            if k == 1:
                _lib.check(lib.sted_host_thread_priority(1))
            hm, hs = maps[k].copy(), np.empty_like(maps[k])
            for rep in range(3):
                call(hm, hs, 10 * k + rep)
            got[k] = (hm, hs)
        except Exception as exc:                     # pragma: no cover
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k in range(2):
        assert np.array_equal(got[k][0], ref[k][0]) and np.array_equal(got[k][1], ref[k][1])


# ------------------------------------------------------------------ other pixel sizes: the PSF side K is not 59
@pytest.mark.parametrize("px,K", [(30e-9, 39), (25e-9, 47)])
def test_other_pixel_sizes_take_the_generic_kernels(px, K):
    """gym-sted images at 20 nm per pixel (K = 59, the compiled-in case); the facade accepts any pixel size, and those
    run ``gather_kernel<0>`` and the run-time-K scans.  Gather and expectation scan against the sequential oracle at
    1e-5 (same PSF cache on both sides: ``physics.py`` equals the oracle's quadrature, ``tests/test_oracle.py``), the
    stochastic scan by its totals."""
    _require_gpu()
    m = helpers.product_microscope()
    cache = m.cache(px)
    assert cache[0].shape == (K, K)
    om = helpers.oracle_microscope()
    om._cache[int(round(px * 1e9))] = tuple(np.asarray(c, dtype=np.float64) for c in cache)
    rng = np.random.default_rng(int(px * 1e10))
    B, H, W = 3, 40, 70
    maps = np.stack([helpers.synthetic_datamap(rng, H, W) for _ in range(B)])
    bm = batched(m, B, H, W, pixelsize=px, seed=5)
    assert bm.K == K
    bm.set_datamaps(torch.from_numpy(maps))
    _, conf_i = bm.get_signal_and_bleach(10e-6, 10e-6, 0.0, bleach=False, want_intensity=True, noise=False)
    _, sted_i = bm.get_signal_and_bleach(30e-6, 10e-6, 0.15, bleach=True, sample_bleach=False, want_intensity=True,
                                         noise=False, update=False)
    exp_map = bm.bleached_roi().cpu().numpy()
    _, stoch_i = bm.get_signal_and_bleach(30e-6, 10e-6, 0.15, bleach=True, want_intensity=True, noise=False)
    stoch_map = bm.datamaps_roi().cpu().numpy()
    for b in range(B):
        dm = po.Datamap(maps[b], px)
        dm.set_roi(om.cache(px)[0], "max")
        E0 = om.get_effective(px, 10e-6, 0.0)
        E1 = om.get_effective(px, 10e-6, 0.15)
        S1 = np.exp(-om.get_k_bleach_map(px, 10e-6, 0.15, 30e-6) * 30e-6)
        work = dm.whole_datamap.astype(np.float64).copy()
        i0 = po.raster(work, H, W, E0, np.ones_like(E0), 0)
        np.testing.assert_allclose(conf_i[b].cpu().numpy(), i0, rtol=RTOL, atol=1e-7 * i0.max())
        work = dm.whole_datamap.astype(np.float64).copy()
        i1 = po.raster(work, H, W, E1, S1, 1)
        np.testing.assert_allclose(sted_i[b].cpu().numpy(), i1, rtol=RTOL, atol=1e-7 * i1.max())
        np.testing.assert_allclose(exp_map[b], work[dm.roi], rtol=RTOL, atol=1e-6)
        # stochastic scan: integer counts that only fall, totals and image within a few sigma of the expectation
        assert np.all(stoch_map[b] == np.rint(stoch_map[b])) and np.all(stoch_map[b] <= maps[b])
        lost, lost_exp = maps[b].sum() - stoch_map[b].sum(), maps[b].sum() - work[dm.roi].sum()
        assert abs(lost - lost_exp) < 6 * np.sqrt(lost_exp + 1)
        assert abs(stoch_i[b].cpu().numpy().sum() / i1.sum() - 1) < 0.02


def test_psf_larger_than_the_compiled_limit_is_refused():
    _require_gpu()
    m = helpers.product_microscope()
    with pytest.raises(_lib.StedError):
        batched(m, 1, 16, 16, pixelsize=12e-9).get_signal_and_bleach(10e-6, 10e-6, 0.0, bleach=False)
