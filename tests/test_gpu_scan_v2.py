"""The second-generation stochastic scan (fused presample + scatter + ``sweep2_kernel``: row-range work items, groups
of T scan rows between barriers, correction warps) against the round-1 schedule it replaces and against the session API.

Both schedules draw the same deaths (same Philox keys) and apply them with order-independent integer arithmetic: the
bleached maps must be IDENTICAL bit for bit for every split of the scan into row ranges and row groups.  The intensities
agree to float32 rounding (and the Poisson counts drawn from them with the same Philox numbers almost everywhere): the
second-generation gather sums even and odd taps in separate packed accumulators (FFMA2), and with T > 1 a later row of a
group is gathered against the counts at the start of the group and corrected for the deaths in between."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from gym_sted_b200 import defaults, session as sess, synapse  # noqa: E402
from tests import helpers  # noqa: E402

pytestmark = pytest.mark.gpu


def _run(m, maps, act, seed, env=None):
    from gym_sted_b200.batch import BatchedMicroscope
    saved = {k: os.environ.get(k) for k in ("STED_B200_SCAN_V1", "STED_B200_ROW_RANGE", "STED_B200_ROW_GROUP", "STED_B200_SCAN_V3")}
    try:
        for k, v in (env or {}).items():
            os.environ[k] = v
        B, H, W = maps.shape
        bm = BatchedMicroscope(m, B, H, W, device="cuda:0", seed=seed, env_base=100)
        bm.set_datamaps(torch.from_numpy(maps))
        sig, inten = bm.get_signal_and_bleach(*act, bleach=True, want_intensity=True)
        assert bm.dropped_events() == 0
        return sig.cpu().numpy(), inten.cpu().numpy(), bm.datamaps_roi().cpu().numpy()
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("H,W,B", [(64, 64, 7), (96, 96, 5), (100, 200, 3), (256, 256, 3), (40, 300, 2), (17, 23, 4)])
def test_scan_v2_equals_v1_for_every_row_range(H, W, B):
    if not torch.cuda.is_available():
        pytest.fail("needs a CUDA device")
    m = helpers.product_microscope(defaults.load_routine("high-signal_low-bleach"), noise=True,
                                   background=defaults.DETECTOR["background"])
    maps, _ = synapse.generate_batch(B, max(H, 64), max(W, 64), seed=H + W)
    maps = np.ascontiguousarray(maps[:, -H:, -W:]).astype(np.float32)
    rng = np.random.default_rng(H)
    if maps.sum() < 50 * B:                          # a crop that missed the structure: add a dense patch
        maps[:, H // 4:H // 2 + 2, W // 4:W // 2 + 2] = rng.integers(1, 30, (B, H // 2 + 2 - H // 4, W // 2 + 2 - W // 4))
    act = (rng.uniform(10e-6, 50e-6, B), rng.uniform(3e-6, 15e-6, B), rng.uniform(0.05, 0.3, B))
    ref = _run(m, maps, act, 5, {"STED_B200_SCAN_V1": "1"})
    assert ref[2].sum() < maps.sum()
    for rr in (None, "16", "1", "37", str(H)):
        for grp in (None, "1", "2", "3", "4", "v3"):     # "v3": the third-generation kernel (TMA ring, specialised warps)
            env = {} if rr is None else {"STED_B200_ROW_RANGE": rr}
            if grp == "v3":
                env["STED_B200_SCAN_V3"] = "1"
            elif grp is not None:
                env["STED_B200_ROW_GROUP"] = grp
            got = _run(m, maps, act, 5, env)
            what = f"(row range {rr}, row group {grp})"
            assert np.array_equal(got[2], ref[2]), f"bleached map differs from the round-1 schedule {what}"
            np.testing.assert_allclose(got[1], ref[1], rtol=2e-5, atol=2e-6 * ref[1].max(), err_msg=what)
            assert (got[0] != ref[0]).mean() < 1e-3 and np.abs(got[0] - ref[0]).max() <= 1, what


def test_frame_molecules_take_part_in_the_scan():
    """STED_FRAME_MOLECULES: molecules in the K//2 frame are imaged and bleached (expectation and stochastic modes
    agree on the totals); without the flag the stochastic scan leaves the frame alone."""
    from gym_sted_b200 import base
    m = helpers.product_microscope()
    rng = np.random.default_rng(0)
    H = W = 40
    dm = base.Datamap(helpers.synthetic_datamap(rng, H, W), helpers.PX)
    dm.set_roi(m.cache(helpers.PX)[0], "max")
    whole = dm.whole_datamap.copy()
    whole[24:29, 29:60] = 50                      # inside the frame (rows < 29), next to the first scan rows
    dm.whole_datamap = whole
    act = dict(pdt=40e-6, p_ex=10e-6, p_sted=0.2)
    _, bl_e, _ = m.get_signal_and_bleach(dm, helpers.PX, **act, bleach=True, update=False, sample_bleach=False)
    _, bl_s, _ = m.get_signal_and_bleach(dm, helpers.PX, **act, bleach=True, update=False, seed=1)
    fe, fs = bl_e["base"][24:29, 29:60].sum(), bl_s["base"][24:29, 29:60].sum()
    assert fe < 0.98 * whole[24:29, 29:60].sum()
    assert abs(fs - fe) < 6 * np.sqrt(whole[24:29, 29:60].sum())
    assert np.all(bl_s["base"] <= whole)


def test_session_equals_device_api():
    """upload once -> confocal, STED + bleaching, confocal, download: identical to the device-resident calls."""
    from gym_sted_b200.batch import BatchedMicroscope
    m = helpers.product_microscope(defaults.load_routine("high-signal_low-bleach"), noise=True,
                                   background=defaults.DETECTOR["background"])
    B, H, W = 12, 48, 80
    maps, _ = synapse.generate_batch(B, 64, 128, seed=2)
    maps = np.ascontiguousarray(maps[:, :H, :W])
    rng = np.random.default_rng(1)
    pdt, p_ex, p_sted = rng.uniform(10e-6, 40e-6, B), rng.uniform(3e-6, 15e-6, B), rng.uniform(0.05, 0.25, B)
    bm = BatchedMicroscope(m, B, H, W, device="cuda:0", seed=9, env_base=40)
    bm.set_datamaps(torch.from_numpy(maps))
    ref = []
    ref.append(bm.get_signal_and_bleach(10e-6, 10e-6, 0.0, bleach=False)[0].cpu().numpy())
    s, i = bm.get_signal_and_bleach(pdt, p_ex, p_sted, bleach=True, want_intensity=True)
    ref.append(s.cpu().numpy())
    ref_int = i.cpu().numpy()
    ref.append(bm.get_signal_and_bleach(10e-6, 10e-6, 0.0, bleach=False)[0].cpu().numpy())
    ref_map = bm.datamaps_roi().cpu().numpy()

    for u16 in (True, False):
        dt = np.uint16 if u16 else np.float32
        s = sess.StedSession(m, B, H, W, env_base=40, seed=9)
        s.upload(maps.astype(dt), max_molecules=int(maps.sum()))
        outs = [sess.pinned_empty((B, H, W), dt) for _ in range(3)]
        inten = np.empty((B, H, W), dtype=np.float32)
        s.acquire(10e-6, 10e-6, 0.0, bleach=False, out=outs[0])
        s.acquire(pdt, p_ex, p_sted, bleach=True, out=outs[1], intensity_out=inten)
        s.acquire(10e-6, 10e-6, 0.0, bleach=False, out=outs[2])
        got_map = s.download(np.empty((B, H, W), dtype=dt))
        assert s.sync() == 0
        for a, b in zip(outs, ref):
            assert np.array_equal(np.asarray(a).astype(np.float32), b)
        assert np.array_equal(inten, ref_int)
        assert np.array_equal(got_map.astype(np.float32), ref_map)
        # a second structure through the same session; results independent of what ran before
        s.upload(maps[::-1].copy().astype(dt), max_molecules=int(maps.sum()))
        out2 = s.acquire(10e-6, 10e-6, 0.0, bleach=False, noise=False, out=np.empty((B, H, W), dtype=np.float32))
        s.sync()
        bm2 = BatchedMicroscope(m, B, H, W, device="cuda:0")
        bm2.set_datamaps(torch.from_numpy(maps[::-1].copy()))
        exp2 = bm2.get_signal_and_bleach(10e-6, 10e-6, 0.0, bleach=False, noise=False)[0].cpu().numpy()
        assert np.array_equal(out2, exp2)
        s.close()


def test_session_reports_event_overflow():
    m = helpers.product_microscope()
    B, H, W = 2, 32, 32
    maps = np.full((B, H, W), 40, dtype=np.uint16)
    s = sess.StedSession(m, B, H, W)
    s.upload(maps, max_molecules=100)                 # far too small for 2 x 1024 x 40 molecules
    s.acquire(50e-6, 10e-6, 0.3, bleach=True, noise=False)
    from gym_sted_b200 import _lib
    with pytest.raises(_lib.StedError) as err:
        s.sync()
    assert err.value.code == -5
    s.upload(maps, max_molecules=int(maps.sum()))
    s.acquire(50e-6, 10e-6, 0.3, bleach=True, noise=False)
    assert s.sync() == 0
    s.close()


def test_session_device_resident_io():
    """upload_device / result / join / kernel_ms: structures and images stay on the GPU, same numbers as the host-buffer
    session calls and as the single-stream device API (the session spreads an acquisition over three streams)."""
    m = helpers.product_microscope(defaults.load_routine("high-signal_low-bleach"), noise=True,
                                   background=defaults.DETECTOR["background"])
    B, H, W = 6, 40, 72
    maps, _ = synapse.generate_batch(B, 64, 128, seed=4)
    maps = np.ascontiguousarray(maps[:, :H, :W]).astype(np.float32)
    rng = np.random.default_rng(2)
    pdt, p_ex, p_sted = rng.uniform(10e-6, 40e-6, B), rng.uniform(3e-6, 15e-6, B), rng.uniform(0.05, 0.25, B)

    host = sess.StedSession(m, B, H, W, env_base=7, seed=3)
    host.upload(maps, max_molecules=int(maps.sum()))
    ref = [np.empty((B, H, W), dtype=np.float32) for _ in range(3)]
    host.acquire(10e-6, 10e-6, 0.0, bleach=False, out=ref[0])
    host.acquire(pdt, p_ex, p_sted, bleach=True, out=ref[1])
    host.acquire(10e-6, 10e-6, 0.0, bleach=False, out=ref[2])
    ref_map = host.download(np.empty((B, H, W), dtype=np.float32))
    assert host.sync() == 0
    host.close()

    d = sess.StedSession(m, B, H, W, env_base=7, seed=3)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):                        # the maps are produced on a stream of the caller's
        dev_maps = torch.from_numpy(maps).cuda() * 1.0
        d.upload_device(dev_maps, max_molecules=int(maps.sum()))
    d.acquire(10e-6, 10e-6, 0.0, bleach=False)
    d.acquire(pdt, p_ex, p_sted, bleach=True)
    d.acquire(10e-6, 10e-6, 0.0, bleach=False)
    got = [d.result(back).clone() for back in (2, 1, 0)]          # current stream waits for each detector
    torch.cuda.synchronize()
    for a, b in zip(got, ref):
        assert np.array_equal(a.cpu().numpy(), b)
    assert all(d.kernel_ms(back) > 0 for back in range(3))
    with pytest.raises(Exception):
        d.kernel_ms(3)                                           # only three acquisitions were queued
    out = d.download(np.empty((B, H, W), dtype=np.float32))
    d.join()
    torch.cuda.current_stream().synchronize()                    # join: the download has landed
    assert np.array_equal(out, ref_map)
    # a second device upload (uint16) replaces the structure
    d.upload_device(torch.from_numpy(maps[::-1].copy()).cuda().to(torch.uint16), max_molecules=int(maps.sum()))
    d.acquire(10e-6, 10e-6, 0.0, bleach=False, noise=False)
    img = d.result(0).clone()
    from gym_sted_b200.batch import BatchedMicroscope
    bm = BatchedMicroscope(m, B, H, W, device="cuda:0")
    bm.set_datamaps(torch.from_numpy(maps[::-1].copy()))
    exp = bm.get_signal_and_bleach(10e-6, 10e-6, 0.0, bleach=False, noise=False)[0]
    torch.cuda.synchronize()
    assert torch.equal(img, exp)
    assert d.sync() == 0
    d.close()


def test_session_half_batch_gathers_and_ordering_do_not_change_results():
    """Batches of >= 64 environments: a confocal acquisition with a host destination runs as two half-batch gathers with
    the detector / conversion / copy in runs of environments (odd sizes included), a low-priority session is ordered in
    front with ``after`` and adopted.  Photon counts and maps must equal the single-launch device API bit for bit."""
    from gym_sted_b200.batch import BatchedMicroscope
    m = helpers.product_microscope(defaults.load_routine("high-signal_low-bleach"), noise=True,
                                   background=defaults.DETECTOR["background"])
    B, H, W = 71, 24, 40
    maps, _ = synapse.generate_batch(B, 64, 64, seed=9)
    maps = np.ascontiguousarray(maps[:, 10:10 + H, 8:8 + W]).astype(np.uint16)
    nxt_maps = np.ascontiguousarray(maps[::-1])
    rng = np.random.default_rng(5)
    pdt, p_ex, p_sted = rng.uniform(10e-6, 40e-6, B), rng.uniform(3e-6, 15e-6, B), rng.uniform(0.05, 0.25, B)
    mol = int(maps.astype(np.int64).sum())

    main = sess.StedSession(m, B, H, W, env_base=3, seed=21)
    low = sess.StedSession(m, B, H, W, env_base=3, seed=22, low_priority=True)
    out = [sess.pinned_empty((B, H, W), np.uint16) for _ in range(5)]
    inten = np.empty((B, H, W), dtype=np.float32)
    main.upload(maps, max_molecules=mol)
    low.upload(nxt_maps, max_molecules=mol)
    low.acquire(10e-6, 10e-6, 0.0, bleach=False, out=out[3])
    main.acquire(10e-6, 10e-6, 0.0, bleach=False, out=out[0], intensity_out=inten)
    main.acquire(pdt, p_ex, p_sted, bleach=True, out=out[1])
    bleached = main.download(sess.pinned_empty((B, H, W), np.uint16))
    main.after(low)
    main.acquire(10e-6, 10e-6, 0.0, bleach=False, out=out[2])
    main.adopt(low)
    main.acquire(10e-6, 10e-6, 0.0, bleach=False, out=out[4])
    assert main.sync() == 0 and low.sync() == 0
    with pytest.raises(_lib_error()):
        main.after(main)
    main.close(); low.close()

    bm = BatchedMicroscope(m, B, H, W, device="cuda:0", seed=21, env_base=3)
    bm.set_datamaps(torch.from_numpy(maps.astype(np.float32)))
    c1, i1 = bm.get_signal_and_bleach(10e-6, 10e-6, 0.0, bleach=False, want_intensity=True)
    st, _ = bm.get_signal_and_bleach(pdt, p_ex, p_sted, bleach=True)
    c2, _ = bm.get_signal_and_bleach(10e-6, 10e-6, 0.0, bleach=False)
    assert np.array_equal(out[0], c1.cpu().numpy().astype(np.uint16))
    assert np.array_equal(inten, i1.cpu().numpy())
    assert np.array_equal(out[1], st.cpu().numpy().astype(np.uint16))
    assert np.array_equal(out[2], c2.cpu().numpy().astype(np.uint16))
    assert np.array_equal(bleached, bm.datamaps_roi().cpu().numpy().astype(np.uint16))
    bm.set_datamaps(torch.from_numpy(nxt_maps.astype(np.float32)))
    c4, _ = bm.get_signal_and_bleach(10e-6, 10e-6, 0.0, bleach=False)       # 4th acquisition of the main session: offset 4
    assert np.array_equal(out[4], c4.cpu().numpy().astype(np.uint16))
    bl = BatchedMicroscope(m, B, H, W, device="cuda:0", seed=22, env_base=3)
    bl.set_datamaps(torch.from_numpy(nxt_maps.astype(np.float32)))
    c3, _ = bl.get_signal_and_bleach(10e-6, 10e-6, 0.0, bleach=False)
    assert np.array_equal(out[3], c3.cpu().numpy().astype(np.uint16))


def _lib_error():
    from gym_sted_b200 import _lib
    return _lib.StedError


def test_presample_then_acquire_equals_one_call():
    """sted_bleach_presample + sted_acquire(STED_PRESAMPLED) on two streams == sted_acquire alone."""
    import ctypes
    from gym_sted_b200 import _lib
    from gym_sted_b200.batch import BatchedMicroscope
    m = helpers.product_microscope(defaults.load_routine("high-signal_low-bleach"), noise=True,
                                   background=defaults.DETECTOR["background"])
    B, H, W = 5, 64, 96
    maps, _ = synapse.generate_batch(B, 64, 128, seed=6)
    maps = np.ascontiguousarray(maps[:, :H, :W]).astype(np.float32)
    act = (np.full(B, 30e-6), np.full(B, 8e-6), np.full(B, 0.2))
    ref = _run(m, maps, act, 11)
    bm = BatchedMicroscope(m, B, H, W, device="cuda:0", seed=11, env_base=100)
    bm.set_datamaps(torch.from_numpy(maps))
    pdt, p_ex, p_sted = (bm._vec(x) for x in act)
    bm.make_effective(p_ex, p_sted, pdt)
    bm.offset += 1
    flags = _lib.STED_BLEACH | _lib.STED_SAMPLE_BLEACH | _lib.STED_SAMPLE_PHOTONS
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    _lib.check(bm.lib.sted_bleach_presample(bm.dmap.data_ptr(), bm.tables.data_ptr(), pdt.data_ptr(), B, H, W, bm.K, flags,
                                            bm.seed, bm.offset, bm.env_base, bm.workspace.data_ptr(), bm.workspace_bytes,
                                            s1.cuda_stream))
    s2.wait_stream(s1)
    sig = torch.empty((B, H, W), dtype=torch.float32, device="cuda")
    inten = torch.empty_like(sig)
    _lib.check(bm.lib.sted_acquire(bm.dmap.data_ptr(), bm.dmap_alt.data_ptr(), bm.tables.data_ptr(), pdt.data_ptr(), B, H, W,
                                   bm.K, flags | _lib.STED_PRESAMPLED, ctypes.byref(bm.det), bm.lambda_fluo, bm.seed, bm.offset,
                                   bm.env_base, inten.data_ptr(), sig.data_ptr(), bm.workspace.data_ptr(), bm.workspace_bytes,
                                   s2.cuda_stream))
    torch.cuda.synchronize()
    p = bm.p
    assert np.array_equal(sig.cpu().numpy(), ref[0])
    assert np.array_equal(inten.cpu().numpy(), ref[1])
    assert np.array_equal(bm.dmap_alt[:, p:p + H, p:p + W].cpu().numpy(), ref[2])
    with pytest.raises(_lib.StedError):                  # the first half of a stochastic bleaching scan only
        _lib.check(bm.lib.sted_bleach_presample(bm.dmap.data_ptr(), bm.tables.data_ptr(), pdt.data_ptr(), B, H, W, bm.K,
                                                _lib.STED_BLEACH, bm.seed, bm.offset, bm.env_base, bm.workspace.data_ptr(),
                                                bm.workspace_bytes, None))
