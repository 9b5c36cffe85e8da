"""GPU parity at the BASELINE shapes (256x256, 512x512, 96x96) and through the kernel instantiations the benchmark
actually runs (``sweep_kernel<59,16,stoch>``: 128-column strips, several strips per environment).

Everything is compared with the CPU oracle (``oracle/raster_oracle.c``, float64, reference dwell order):

* noise-free fields (gather, expectation-mode scan, bleached maps) within 1e-5 relative — the tolerance
  north_star states — at the full sizes;
* stochastic scan: per-pixel means against the oracle's exact mean field (mode 1) in standard errors, with an
  explicit assertion on the columns either side of the strip seam; final molecule counts against the exact
  Binomial(n, exp(-H_tot)) law the reference's chain of per-dwell binomials produces (z-scores with the exact
  variance, KS on every nanodomain pixel); image spreads against oracle mode 2 (per-molecule thinning in the
  reference's order);
* config 3: every environment with its own fluorophore (sigma_abs, k1, b).
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from gym_sted_b200 import _lib, defaults, synapse  # noqa: E402
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


def oracle_tables(om, pdt, p_ex, p_sted):
    E = om.get_effective(PX, p_ex, p_sted)
    kb = om.get_k_bleach_map(PX, p_ex, p_sted, pdt)
    return E, kb


def oracle_acquire(om, roi, pdt, p_ex, p_sted, mode, seed=0):
    dm = po.Datamap(roi, PX)
    dm.set_roi(om.cache(PX)[0], "max")
    E, kb = oracle_tables(om, pdt, p_ex, p_sted)
    work = dm.whole_datamap.astype(np.float64).copy()
    inten = po.raster(work, roi.shape[0], roi.shape[1], E, np.exp(-kb * pdt), mode, seed)
    return inten, work[dm.roi]


def spine_maps(B, H, W, seed):
    maps, _ = synapse.generate_batch(B, H, W, seed=seed)
    return maps.astype(np.int64)


# ------------------------------------------------------------------ noise-free fields at the full sizes
@pytest.mark.parametrize("H,W", [(256, 256), (512, 512), (96, 96)])
def test_gather_full_size_matches_oracle(micro, omicro, H, W):
    B = 2
    rois = spine_maps(B, H, W, seed=H)
    if H == 96:                                   # config 4: dense structure maps, up to 40 molecules per pixel
        rng = np.random.default_rng(96)
        rois[1] = rng.integers(0, 41, (H, W)) * (rng.random((H, W)) < 0.5)
    bm = batched(micro, B, H, W)
    bm.set_datamaps(torch.from_numpy(rois))
    acts = [(10e-6, 10e-6, 0.0), (30e-6, 12e-6, 0.15)]
    pdt, p_ex, p_sted = (torch.tensor([a[i] for a in acts], dtype=torch.float64) for i in range(3))
    _, inten = bm.get_signal_and_bleach(pdt, p_ex, p_sted, bleach=False, want_intensity=True, noise=False)
    inten = inten.cpu().numpy()
    for b, (dt, pe, ps) in enumerate(acts):
        ref, _ = oracle_acquire(omicro, rois[b], dt, pe, ps, 0)
        np.testing.assert_allclose(inten[b], ref, rtol=RTOL, atol=1e-7 * ref.max())


@pytest.mark.parametrize("H,W", [(256, 256), (512, 512), (96, 96)])
def test_expectation_scan_full_size_matches_oracle(micro, omicro, H, W):
    B = 2
    rois = spine_maps(B, H, W, seed=H + 1)
    bm = batched(micro, B, H, W)
    bm.set_datamaps(torch.from_numpy(rois))
    acts = [(30e-6, 10e-6, 0.15), (60e-6, 5e-6, 0.3)]
    pdt, p_ex, p_sted = (torch.tensor([a[i] for a in acts], dtype=torch.float64) for i in range(3))
    _, inten = bm.get_signal_and_bleach(pdt, p_ex, p_sted, bleach=True, sample_bleach=False, want_intensity=True,
                                        noise=False)
    inten, bleached = inten.cpu().numpy(), bm.datamaps_roi().cpu().numpy()
    for b, (dt, pe, ps) in enumerate(acts):
        ref_i, ref_d = oracle_acquire(omicro, rois[b], dt, pe, ps, 1)
        np.testing.assert_allclose(inten[b], ref_i, rtol=RTOL, atol=1e-7 * ref_i.max())
        np.testing.assert_allclose(bleached[b], ref_d, rtol=RTOL, atol=1e-6)
        assert ref_d.sum() < 0.95 * rois[b].sum()


# ------------------------------------------------------------------ stochastic scan through sweep_kernel<59,16,1>
def _stochastic_replicas(micro, roi, R, act, seed):
    H, W = roi.shape
    bm = batched(micro, R, H, W, seed=seed)
    bm.set_datamaps(torch.from_numpy(np.broadcast_to(roi, (R, H, W)).copy()))
    dt, pe, ps = act
    _, inten = bm.get_signal_and_bleach(dt, pe, ps, bleach=True, sample_bleach=True, want_intensity=True, noise=False)
    assert bm.dropped_events() == 0
    return inten.cpu().numpy().astype(np.float64), bm.datamaps_roi().cpu().numpy().astype(np.float64)


def discrete_ks_pvalue(sample, dist, n):
    """Kolmogorov-Smirnov against a DISCRETE law on {0..n}: D = max_k |F_R(k) - F(k)| evaluated at the integers
    (both functions right-continuous), p-value from the continuous-case distribution of D, which is conservative
    for discrete laws.  (scipy.stats.kstest assumes a continuous cdf: its D- term compares F(k) with the empirical
    cdf just BELOW k and is inflated by the probability mass at k.)"""
    from scipy import stats
    ks = np.arange(n + 1)
    ecdf = np.searchsorted(np.sort(sample), ks, side="right") / len(sample)
    d = np.abs(ecdf - dist.cdf(ks)).max()
    return float(stats.kstwo.sf(d, len(sample)))


def _check_counts_against_exact_law(gd, roi, exp_d, R):
    """Final counts are Binomial(n, s) with s = exp_d / n exactly (survival factors commute)."""
    assert np.all(gd == np.rint(gd)) and np.all(gd >= 0) and np.all(gd <= roi)
    occ = roi > 0
    n = roi[occ].astype(np.float64)
    s = exp_d[occ] / n
    var = n * s * (1 - s)
    ok = var > 1e-9
    z = (gd.mean(0)[occ][ok] - (n * s)[ok]) / np.sqrt(var[ok] / R)
    assert np.abs(z).max() < 5.5, f"max |z| = {np.abs(z).max():.2f}"
    assert 0.85 < np.mean(z * z) < 1.18, f"mean z^2 = {np.mean(z * z):.3f}"
    # sample variance against the exact one, pooled
    ratio = gd.var(0, ddof=1)[occ][ok] / var[ok]
    assert abs(ratio.mean() - 1) < 0.05, f"variance ratio {ratio.mean():.3f}"
    # the same for the nanodomain pixels alone (their deaths are drawn as several sub-binomials)
    big = (n > 50)[ok]
    if big.sum() >= 8:
        rb = ratio[big].mean()
        tol = 4.5 * np.sqrt(2.0 / (R - 1) / big.sum()) + 0.01
        assert abs(rb - 1) < tol, f"nanodomain variance ratio {rb:.3f} (tolerance {tol:.3f}, {big.sum()} pixels)"
        assert np.abs(z[big]).max() < 4.5
    return z


@pytest.mark.parametrize("H,W,R", [(64, 256, 256), (48, 160, 256)])
def test_stochastic_scan_two_strips_moments_seam_and_ks(micro, omicro, H, W, R):
    from scipy import stats
    rng = np.random.default_rng(W)
    roi = spine_maps(1, 64, 256, seed=W)[0][:H, :W]
    # nanodomains right on the seam columns (strips are 128 columns wide; K//2 = 29 columns of halo either side)
    for x in (120, 127, 128, 129, 136):
        roi[rng.integers(8, H - 8), x] = rng.integers(100, 176)
    act = (30e-6, 10e-6, 0.15)
    gi, gd = _stochastic_replicas(micro, roi, R, act, seed=11)
    exp_i, exp_d = oracle_acquire(omicro, roi, *act, 1)
    assert 0.15 < exp_d.sum() / roi.sum() < 0.95
    z_d = _check_counts_against_exact_law(gd, roi, exp_d, R)
    assert len(z_d) > 1000
    # image: per-pixel mean against the exact mean field, in standard errors of the replicas
    se = gi.std(0, ddof=1) / np.sqrt(R) + 1e-6 * exp_i.max()
    z_i = (gi.mean(0) - exp_i) / se
    assert np.abs(z_i).max() < 6.0, f"max |z| = {np.abs(z_i).max():.2f} at {np.unravel_index(np.abs(z_i).argmax(), z_i.shape)}"
    seam = slice(120, 137)
    assert np.abs(z_i[:, seam]).max() < 5.0 and abs(z_i[:, seam].mean()) < 0.5
    assert abs(z_i.mean()) < 0.25
    # relative bias of the seam columns (a missing cross-strip correction would show as a step at column 128)
    rel = gi.mean(0)[:, seam].sum(0) / exp_i[:, seam].sum(0) - 1
    assert np.abs(rel).max() < 0.01, rel
    # spreads and distributions against the reference-order per-molecule thinning (oracle mode 2)
    RO = 48
    oi = np.zeros((RO, H, W))
    od = np.zeros((RO, H, W))
    for r in range(RO):
        oi[r], od[r] = oracle_acquire(omicro, roi, *act, 2, seed=500 + r)
    bright = exp_i > 0.25 * exp_i.max()
    ratio = gi.std(0)[bright] / np.maximum(oi.std(0)[bright], 1e-30)
    assert abs(np.median(ratio) - 1) < 0.1 and abs(ratio.mean() - 1) < 0.1, (np.median(ratio), ratio.mean())
    ys, xs = np.nonzero(roi > 50)
    for (yy, xx) in zip(ys, xs):
        n, s = int(roi[yy, xx]), exp_d[yy, xx] / roi[yy, xx]
        # KS against the exact binomial law and against the oracle's draws
        assert discrete_ks_pvalue(gd[:, yy, xx], stats.binom(n, s), n) > 1e-3, (yy, xx)
        assert stats.ks_2samp(gd[:, yy, xx], od[:, yy, xx]).pvalue > 1e-4, (yy, xx)
    y, x = np.unravel_index(exp_i.argmax(), exp_i.shape)
    assert stats.ks_2samp(gi[:, y, x], oi[:, y, x]).pvalue > 1e-4
    assert stats.ks_2samp(gi[:, H // 2, 128], oi[:, H // 2, 128]).pvalue > 1e-4
    assert stats.ks_2samp(gd.sum((1, 2)), od.sum((1, 2))).pvalue > 1e-4


@pytest.mark.parametrize("H,R", [(256, 64), (512, 24), (96, 128)])
def test_stochastic_scan_full_size_means(micro, omicro, H, R):
    """BASELINE sizes: exact-law z-scores of every occupied pixel, image means in replica standard errors."""
    W = H
    roi = spine_maps(1, H, W, seed=7 * H)[0]
    act = (25e-6, 8e-6, 0.12)
    gi, gd = _stochastic_replicas(micro, roi, R, act, seed=H)
    exp_i, exp_d = oracle_acquire(omicro, roi, *act, 1)
    _check_counts_against_exact_law(gd, roi, exp_d, R)
    se = gi.std(0, ddof=1) / np.sqrt(R) + 1e-6 * exp_i.max()
    z = (gi.mean(0) - exp_i) / se
    # Student-t tails with R-1 degrees of freedom: bound the bulk and the bias, not the single worst pixel
    assert np.quantile(np.abs(z), 0.999) < 5.0 and abs(z.mean()) < 0.2
    for seam in range(128, W, 128):                  # every strip seam of sweep_kernel<59,16,1>
        cols = slice(seam - 8, seam + 9)
        rel = gi.mean(0)[:, cols].sum(0) / exp_i[:, cols].sum(0) - 1
        assert np.abs(rel).max() < 0.01, (seam, rel)
    np.testing.assert_allclose(gi.mean(0).sum(), exp_i.sum(), rtol=2e-3)


# ------------------------------------------------------------------ config 3: one fluorophore per environment
def test_per_env_fluorophores_match_oracle(micro):
    from gym_sted_b200 import base
    rng = np.random.default_rng(3)
    B, H, W = 6, 64, 64
    fluos = []
    for b in range(B):
        f = {k: (dict(v) if isinstance(v, dict) else v) for k, v in defaults.FLUO.items()}
        f["sigma_abs"][635] = float(rng.uniform(4.4e-22, 3.3e-21))       # ranges of routines.json
        f["k1"] = float(rng.uniform(3.1e-16, 3.6e-16))
        f["b"] = float(rng.uniform(1.62, 1.79))
        fluos.append(f)
    fluos[-1] = defaults.load_routine("low-signal_high-bleach")
    rois = spine_maps(B, H, W, seed=33)
    bm = batched(micro, B, H, W, fluos=[base.Fluorescence(**f) for f in fluos])
    bm.set_datamaps(torch.from_numpy(rois))
    pdt = rng.uniform(10e-6, 60e-6, B)
    p_ex = rng.uniform(2e-6, 18e-6, B)
    p_sted = rng.uniform(0.05, 0.3, B)
    t = lambda a: torch.from_numpy(a)
    kb = bm.make_effective(t(p_ex).cuda(), t(p_sted).cuda(), t(pdt).cuda(), want_kbleach=True).cpu().numpy()
    tab = bm.tables.cpu().numpy()
    _, inten = bm.get_signal_and_bleach(t(pdt), t(p_ex), t(p_sted), bleach=True, sample_bleach=False,
                                        want_intensity=True, noise=False)
    inten, bleached = inten.cpu().numpy(), bm.datamaps_roi().cpu().numpy()
    K = bm.K
    distinct = set()
    for b in range(B):
        om = helpers.oracle_microscope(fluos[b])
        E, kref = oracle_tables(om, pdt[b], p_ex[b], p_sted[b])
        np.testing.assert_allclose(tab[b, _lib.STED_TAB_E, :, :K], E, rtol=2e-7)
        np.testing.assert_allclose(kb[b], kref, rtol=1e-11)
        ref_i, ref_d = oracle_acquire(om, rois[b], pdt[b], p_ex[b], p_sted[b], 1)
        np.testing.assert_allclose(inten[b], ref_i, rtol=RTOL, atol=1e-7 * ref_i.max())
        np.testing.assert_allclose(bleached[b], ref_d, rtol=RTOL, atol=1e-6)
        distinct.add(float(np.float32(E.max() / p_ex[b])))
    assert len(distinct) == B                        # the environments really ran different fluorophores


# ------------------------------------------------------------------ f2: a structure imaged again and again
def test_sequence_trajectory_matches_oracle(micro, omicro):
    """SequenceMOSTED (gym_sted/envs/sequence_sted_env.py:84-109): the same datamap object is imaged step after step and
    carries its photobleaching.  Four steps of (confocal, STED + bleaching, confocal) with different actions per step and
    environment, resident maps on the GPU against the oracle's datamap object mutated in place: every image and the map
    after every step within the noise-free tolerance (float32 rounding accumulates over the steps: 3e-5 at the end)."""
    rng = np.random.default_rng(21)
    B, H, W, steps = 3, 64, 64, 4
    rois = spine_maps(B, H, W, seed=77)
    bm = batched(micro, B, H, W)
    bm.set_datamaps(torch.from_numpy(rois))
    dms = []
    for b in range(B):
        dm = po.Datamap(rois[b], PX)
        dm.set_roi(omicro.cache(PX)[0], "max")
        dms.append(dm)
    conf = (defaults.PDT, defaults.P_EX, 0.0)
    t = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64))
    for s in range(steps):
        pdt, p_ex, p_sted = rng.uniform(10e-6, 60e-6, B), rng.uniform(2e-6, 18e-6, B), rng.uniform(0.02, 0.3, B)
        _, c1 = bm.get_signal_and_bleach(*conf, bleach=False, want_intensity=True, noise=False)
        _, st = bm.get_signal_and_bleach(t(pdt), t(p_ex), t(p_sted), bleach=True, sample_bleach=False, want_intensity=True,
                                         noise=False)
        _, c2 = bm.get_signal_and_bleach(*conf, bleach=False, want_intensity=True, noise=False)
        maps = bm.datamaps_roi().cpu().numpy()
        tol = RTOL * (1 + s)
        for b in range(B):
            _, _, o1 = omicro.get_signal_and_bleach(dms[b], PX, *conf, bleach=False)
            _, _, os_ = omicro.get_signal_and_bleach(dms[b], PX, pdt[b], p_ex[b], p_sted[b], bleach=True, sample_bleach=False)
            _, _, o2 = omicro.get_signal_and_bleach(dms[b], PX, *conf, bleach=False)
            for got, ref in ((c1, o1), (st, os_), (c2, o2)):
                np.testing.assert_allclose(got[b].cpu().numpy(), ref, rtol=tol, atol=1e-7 * ref.max(), err_msg=f"step {s} env {b}")
            np.testing.assert_allclose(maps[b], dms[b]["base"][dms[b].roi], rtol=tol, atol=1e-6, err_msg=f"map, step {s} env {b}")
    assert maps.sum() < 0.9 * rois.sum()                 # the trajectory does bleach
