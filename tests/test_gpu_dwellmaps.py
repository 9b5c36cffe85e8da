"""The general scan: per-dwell ``pdt`` / ``p_ex`` / ``p_sted`` arrays as the timed environments
hand them to ``get_signal_and_bleach`` (``gym_sted/envs/timed_sted_env.py:141-148``), against the oracle's dwell loop
with one (E, S) table per distinct parameter triple (``oracle/raster_oracle.c: oracle_raster_tables``), which makes no
linearity assumption — so these tests also pin the kernel's decomposition (E and k_b linear in p_ex, hazard linear in
pdt) against the physics."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from gym_sted_b200 import base, defaults  # noqa: E402
from tests import helpers  # noqa: E402

pytestmark = pytest.mark.gpu

HSLB = defaults.load_routine("high-signal_low-bleach")


def _pair(maps, fluo=None, noise=False, background=0.0):
    from oracle import pysted_oracle as po
    pm, om = helpers.product_microscope(fluo, noise, background), helpers.oracle_microscope(fluo, noise, background)
    i_ex = pm.cache(helpers.PX)[0]
    dp, do = base.Datamap(maps, helpers.PX), po.Datamap(maps, helpers.PX)
    dp.set_roi(i_ex, "max")
    do.set_roi(i_ex, "max")
    return pm, om, dp, do


def _param_maps(rng, H, W, kind):
    pdt = np.full((H, W), 20e-6)
    pex = np.full((H, W), 8e-6)
    psted = np.full((H, W), 0.12)
    if kind == "clock":                 # the experiment's clock runs out mid-image: zeros from a raster position on
        flat = pdt.reshape(-1)
        flat[int(0.6 * H * W) + 3:] = 0.0
    elif kind == "pdt":
        pdt = rng.uniform(5e-6, 40e-6, (H, W))
    elif kind == "pdt_pex":
        pdt = rng.choice([5e-6, 10e-6, 30e-6], (H, W))
        pex = rng.choice([2e-6, 8e-6, 15e-6], (H, W))
    elif kind == "all":
        pdt = rng.choice([5e-6, 10e-6, 30e-6], (H, W))
        pex = rng.choice([2e-6, 8e-6], (H, W))
        psted = rng.choice([0.0, 0.05, 0.12, 0.3], (H, W))
    elif kind == "rows":                # depletion power changes from line to line (a line-wise adaptive scheme)
        psted = np.repeat(rng.uniform(0.0, 0.3, (H, 1)), W, axis=1)
    return pdt, pex, psted


@pytest.mark.parametrize("kind", ["clock", "pdt", "pdt_pex", "all", "rows"])
@pytest.mark.parametrize("H,W", [(24, 31), (40, 40)])
def test_expectation_scan_matches_oracle(kind, H, W):
    rng = np.random.default_rng(H * 7 + len(kind))
    maps = helpers.synthetic_datamap(rng, H, W)
    pm, om, dp, do = _pair(maps, HSLB)
    pdt, pex, psted = _param_maps(rng, H, W, kind)
    sig_p, bl_p, int_p = pm.get_signal_and_bleach(dp, helpers.PX, pdt, pex, psted, bleach=True, sample_bleach=False)
    sig_o, bl_o, int_o = om.get_signal_and_bleach(do, helpers.PX, pdt, pex, psted, bleach=True, sample_bleach=False)
    np.testing.assert_allclose(int_p, int_o, rtol=1e-5, atol=1e-7 * int_o.max())
    np.testing.assert_allclose(bl_p["base"], bl_o["base"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(sig_p, sig_o, rtol=2e-5, atol=2e-5 * max(sig_o.max(), 1.0))
    assert bl_o["base"].sum() < maps.sum() * 0.999          # the case does bleach
    np.testing.assert_array_equal(dp.whole_datamap, bl_p["base"])      # mutated in place like pySTED


def test_no_bleach_with_frame_molecules():
    rng = np.random.default_rng(3)
    H = W = 20
    pm, om, dp, do = _pair(helpers.synthetic_datamap(rng, H, W))
    frame = rng.integers(0, 4, dp.whole_datamap.shape)
    dp.whole_datamap = frame + dp.whole_datamap
    do.whole_datamap = dp.whole_datamap.copy()
    pdt, pex, psted = _param_maps(rng, H, W, "all")
    _, bl_p, int_p = pm.get_signal_and_bleach(dp, helpers.PX, pdt, pex, psted, bleach=False)
    _, bl_o, int_o = om.get_signal_and_bleach(do, helpers.PX, pdt, pex, psted, bleach=False)
    np.testing.assert_allclose(int_p, int_o, rtol=1e-5, atol=1e-7 * int_o.max())
    np.testing.assert_array_equal(bl_p["base"], do.whole_datamap)


def test_constant_arrays_take_the_row_sweep_and_agree_with_the_general_scan():
    rng = np.random.default_rng(11)
    H = W = 32
    maps = helpers.synthetic_datamap(rng, H, W)
    pm, _, d1, _ = _pair(maps, HSLB)
    _, _, d2, _ = _pair(maps, HSLB)
    ones = np.ones((H, W))
    a = pm.get_signal_and_bleach(d1, helpers.PX, 20e-6 * ones, 8e-6, 0.1 * ones, bleach=True, sample_bleach=False)
    wob = ones.copy()
    wob[0, 0] = 1.0 + 1e-12                                  # not constant: forces the general scan
    b = pm.get_signal_and_bleach(d2, helpers.PX, 20e-6 * wob, 8e-6, 0.1, bleach=True, sample_bleach=False)
    np.testing.assert_allclose(a[2], b[2], rtol=1e-5, atol=1e-7 * a[2].max())
    np.testing.assert_allclose(a[1]["base"], b[1]["base"], rtol=1e-5, atol=1e-6)


def test_anti_stoke_makes_tables_per_power_pair():
    rng = np.random.default_rng(5)
    H = W = 18
    maps = helpers.synthetic_datamap(rng, H, W)
    pm, om, dp, do = _pair(maps, HSLB)
    pm.sted.anti_stoke = om.sted.anti_stoke = True
    pdt, pex, psted = _param_maps(rng, H, W, "all")
    _, bl_p, int_p = pm.get_signal_and_bleach(dp, helpers.PX, pdt, pex, psted, bleach=True, sample_bleach=False)
    _, bl_o, int_o = om.get_signal_and_bleach(do, helpers.PX, pdt, pex, psted, bleach=True, sample_bleach=False)
    np.testing.assert_allclose(int_p, int_o, rtol=1e-5, atol=1e-7 * int_o.max())
    np.testing.assert_allclose(bl_p["base"], bl_o["base"], rtol=1e-5, atol=1e-6)


def test_too_many_depletion_settings_is_an_error():
    rng = np.random.default_rng(1)
    pm, _, dp, _ = _pair(helpers.synthetic_datamap(rng, 20, 20))
    with pytest.raises(NotImplementedError):
        pm.get_signal_and_bleach(dp, helpers.PX, 10e-6, 5e-6, rng.uniform(0, 0.3, (20, 20)), bleach=False)


def test_stochastic_scan_moments_and_reproducibility():
    """Binomial chain per dwell: integer maps that only lose molecules, identical for identical (seed, counter), and
    whose mean over replicas follows the oracle's expectation scan pixel by pixel (5 standard errors, binomial variance
    bound n*p*(1-p) <= mean)."""
    rng = np.random.default_rng(2)
    H, W = 20, 24
    maps = helpers.synthetic_datamap(rng, H, W, n_domains=4)
    fluo = defaults.load_routine("high-signal_high-bleach")
    pdt, pex, psted = _param_maps(rng, H, W, "all")
    pm, om, _, do = _pair(maps, fluo)
    _, bl_o, _ = om.get_signal_and_bleach(do, helpers.PX, pdt, pex, psted, bleach=True, sample_bleach=False)
    expect = bl_o["base"]
    R = 200
    acc = np.zeros_like(expect)
    first = None
    for rep in range(R):
        dp = base.Datamap(maps, helpers.PX)
        dp.set_roi(pm.cache(helpers.PX)[0], "max")
        pm._acq_counter = 0
        _, bl, _ = pm.get_signal_and_bleach(dp, helpers.PX, pdt, pex, psted, bleach=True, seed=1000 + rep)
        out = bl["base"]
        assert out.dtype == np.int64 and (out >= 0).all() and (out <= dp_initial(maps, out.shape)).all()
        acc += out
        if rep == 0:
            first = out
    dp = base.Datamap(maps, helpers.PX)
    dp.set_roi(pm.cache(helpers.PX)[0], "max")
    pm._acq_counter = 0
    _, again, _ = pm.get_signal_and_bleach(dp, helpers.PX, pdt, pex, psted, bleach=True, seed=1000)
    np.testing.assert_array_equal(again["base"], first)
    mean = acc / R
    n0 = dp_initial(maps, expect.shape)
    var = np.maximum(expect * (1.0 - expect / np.maximum(n0, 1)), 1e-9)
    z = (mean - expect) / np.sqrt(var / R + 1e-12)
    live = n0 > 0
    assert expect[live].sum() < 0.9 * n0.sum()               # substantial bleaching
    assert np.abs(z[live]).max() < 5.0, np.abs(z[live]).max()
    assert abs(z[live].mean()) < 0.3                          # no systematic bias across the map
    assert (mean[~live] == 0).all()


def dp_initial(maps, shape):
    p = (shape[0] - maps.shape[0]) // 2
    return np.pad(maps, p).astype(np.float64)


def test_detector_uses_each_pixels_dwell_time():
    rng = np.random.default_rng(9)
    H = W = 16
    maps = helpers.synthetic_datamap(rng, H, W)
    bg = defaults.DETECTOR["background"]
    pm, om, dp, do = _pair(maps, None, noise=False, background=bg)
    pdt = rng.uniform(0, 50e-6, (H, W))
    pdt[3, :] = 0.0
    sig_p, _, int_p = pm.get_signal_and_bleach(dp, helpers.PX, pdt, 10e-6, 0.0, bleach=False)
    sig_o, _, int_o = om.get_signal_and_bleach(do, helpers.PX, pdt, 10e-6, 0.0, bleach=False)
    np.testing.assert_allclose(sig_p, sig_o, rtol=2e-5, atol=1e-4)
    assert (sig_p[3] == 0).all()


def test_batched_c_abi_matches_single_calls():
    """B environments with their own weight maps through sted_acquire_dwellmaps in one launch == B single launches."""
    import ctypes
    from gym_sted_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(4)
    B, H, W = 3, 16, 20
    pm = helpers.product_microscope(HSLB)
    cache = np.ascontiguousarray(np.stack(pm.cache(helpers.PX)))
    K = cache.shape[-1]
    Hp, Wp = H + K - 1, W + K - 1
    Wpitch = (Wp + 3) & ~3
    dev = torch.device("cuda")
    dmaps = np.zeros((B, Hp, Wpitch), np.float32)
    for b in range(B):
        dmaps[b, K // 2:K // 2 + H, K // 2:K // 2 + W] = helpers.synthetic_datamap(rng, H, W)
    nt = 2
    fl = base.fluo_struct(pm.fluo, pm.excitation, pm.sted)
    d_fluo = torch.frombuffer(bytearray(bytes(fl)), dtype=torch.uint8).to(dev).repeat(B * nt)
    pex = torch.tensor(rng.uniform(5e-6, 10e-6, B).repeat(nt), dtype=torch.float64, device=dev)
    psted = torch.tensor(np.tile([0.05, 0.2], B), dtype=torch.float64, device=dev)
    pdt = torch.tensor(rng.uniform(10e-6, 30e-6, B), dtype=torch.float64, device=dev)
    pdt_t = pdt.repeat_interleave(nt)
    tables = torch.empty((B * nt, _lib.STED_TABLES, K, _lib.kpitch(K)), dtype=torch.float32, device=dev)
    kb = torch.empty((B * nt, K, K), dtype=torch.float64, device=dev)
    d_cache = torch.from_numpy(cache).to(dev)
    _lib.check(lib.sted_make_effective(d_cache.data_ptr(), d_fluo.data_ptr(), pex.data_ptr(), psted.data_ptr(), pdt_t.data_ptr(),
                                       B * nt, K, tables.data_ptr(), kb.data_ptr(), None))
    wdt = torch.tensor(rng.uniform(0.2, 1.0, (B, H, W)), dtype=torch.float32, device=dev)
    wex = torch.tensor(rng.uniform(0.2, 1.0, (B, H, W)), dtype=torch.float32, device=dev)
    tid = torch.tensor(rng.integers(0, nt, (B, H, W)), dtype=torch.uint8, device=dev)
    det = _lib.StedDetectorParams(0.1, 0.5, 0.0)
    d_in = torch.from_numpy(dmaps).to(dev)

    def run(sl, env_base):
        n = sl.stop - sl.start
        out = torch.zeros((n, Hp, Wpitch), dtype=torch.float32, device=dev)
        inten = torch.empty((n, H, W), dtype=torch.float32, device=dev)
        sig = torch.empty_like(inten)
        _lib.check(lib.sted_acquire_dwellmaps(
            d_in[sl].contiguous().data_ptr(), out.data_ptr(), tables[sl.start * nt:sl.stop * nt].contiguous().data_ptr(),
            kb[sl.start * nt:sl.stop * nt].contiguous().data_ptr(), pdt[sl].contiguous().data_ptr(),
            wdt[sl].contiguous().data_ptr(), wex[sl].contiguous().data_ptr(), tid[sl].contiguous().data_ptr(), nt, n, H, W, K,
            _lib.STED_BLEACH | _lib.STED_SAMPLE_BLEACH | _lib.STED_SAMPLE_PHOTONS, ctypes.byref(det), 690e-9, 77, 3, env_base,
            inten.data_ptr(), sig.data_ptr(), None))
        torch.cuda.synchronize()
        return out.cpu().numpy(), inten.cpu().numpy(), sig.cpu().numpy()

    whole = run(slice(0, B), 10)
    for b in range(B):
        one = run(slice(b, b + 1), 10 + b)
        for x, y in zip(whole, one):
            np.testing.assert_array_equal(x[b], y[0])        # env ids key the draws: sharding-invariant
    assert whole[0].sum() < dmaps.sum()
