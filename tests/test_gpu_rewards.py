"""GPU parity of the reward side (Otsu foregrounds, Bleach, SNR, Resolution) and env API conformance."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from gym_sted_b200 import defaults, envs, rewards  # noqa: E402
from oracle import rewards_oracle as ro  # noqa: E402
from tests import helpers  # noqa: E402

pytestmark = pytest.mark.gpu


def _gpu():
    if not torch.cuda.is_available():
        pytest.fail("these tests need a CUDA device (no CPU fallback exists)")


def _check_against_oracle(conf1, sted, conf2):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().cuda()
    fg_s, fg_c, bl, snr, flags = rewards.foreground_objectives(t(conf1), t(sted), t(conf2))
    fg_s, fg_c, bl, snr = fg_s.cpu().numpy(), fg_c.cpu().numpy(), bl.cpu().numpy(), snr.cpu().numpy()
    for i in range(len(conf1)):
        rs, rc = ro.foregrounds(conf1[i], sted[i])
        assert np.array_equal(fg_c[i], rc), i
        assert np.array_equal(fg_s[i], rs), i
        assert bl[i] == ro.bleach(conf1[i], conf2[i], rc), i                      # bit-exact float64
        ref = ro.signal_ratio(sted[i], conf1[i], rs, rc)
        if ref is None:
            assert np.isnan(snr[i]) and int(flags[i]) & 1
        else:
            assert snr[i] == ref, (i, snr[i], ref)


def test_foregrounds_and_objectives_reproduce_recorded_episodes():
    """Integer outputs (masks) bit-exact; Bleach/SNR bit-exact in float64 against the reference's records."""
    _gpu()
    demo = helpers.load_demonstrations()
    _check_against_oracle(demo["conf1"], demo["sted_image"], demo["conf2"])
    t = lambda a: torch.from_numpy(a).float().cuda()
    fg_s, fg_c, bl, snr, _ = rewards.foreground_objectives(t(demo["conf1"]), t(demo["sted_image"]), t(demo["conf2"]))
    assert np.array_equal(fg_c.cpu().numpy(), demo["fg_c"]) and np.array_equal(fg_s.cpu().numpy(), demo["fg_s"])
    assert np.array_equal(bl.cpu().numpy(), demo["mo_objs"][:, 1])
    assert np.array_equal(snr.cpu().numpy(), demo["mo_objs"][:, 2])
    res = rewards.resolution(t(demo["sted_image"]))
    assert np.array_equal(res, demo["mo_objs"][:, 0])


def test_foregrounds_edge_cases():
    _gpu()
    rng = np.random.default_rng(0)
    H = W = 48
    conf1 = rng.poisson(3.0, (6, H, W)) + (rng.random((6, H, W)) < 0.1) * rng.poisson(60, (6, H, W))
    sted = rng.poisson(1.0, (6, H, W)) + (rng.random((6, H, W)) < 0.05) * rng.poisson(25, (6, H, W))
    conf2 = rng.poisson(2.0, (6, H, W))
    sted[1] = 0                                   # empty STED image -> fg_s = fg_c, SNR from the ones-mask branch
    sted[2] = rng.poisson(0.02, (H, W))           # nearly empty
    conf1[3] = rng.integers(0, 20000, (H, W))     # wide value range -> global histogram path
    sted[4] = rng.poisson(40.0, (H, W)) * (conf1[4] <= np.median(conf1[4]))      # bright background -> SNR < 0 -> None
    conf1[5, :, : W // 2] = 7; conf1[5, :, W // 2:] = 9                           # two-valued image
    _check_against_oracle(conf1.astype(np.int64), sted.astype(np.int64), conf2.astype(np.int64))


def test_objectives_on_simulated_acquisitions():
    _gpu()
    from gym_sted_b200.batch import BatchedMicroscope
    from gym_sted_b200 import synapse
    m = helpers.product_microscope(noise=True, background=defaults.DETECTOR["background"])
    B = 8
    maps, _ = synapse.generate_batch(B, 64, 64, seed=9)
    bm = BatchedMicroscope(m, B, 64, 64, seed=3)
    bm.set_datamaps(torch.from_numpy(maps))
    conf1, _ = bm.get_signal_and_bleach(10e-6, 10e-6, 0.0, bleach=False)
    sted, _ = bm.get_signal_and_bleach(20e-6, 10e-6, 0.1, bleach=True)
    conf2, _ = bm.get_signal_and_bleach(10e-6, 10e-6, 0.0, bleach=False)
    assert bm.dropped_events() == 0
    c1, s, c2 = (np.rint(x.cpu().numpy()).astype(np.int64) for x in (conf1, sted, conf2))
    _check_against_oracle(c1, s, c2)
    res = rewards.resolution(sted)
    ref = np.array([ro.resolution(s[b]) for b in range(B)])
    np.testing.assert_allclose(res, ref, rtol=1e-12)
    assert np.all(res < 250) and np.all(res > 30)             # STED at 100 mW resolves well below the confocal limit


def test_env_api_conformance_single():
    """Spaces, dtypes, info keys, episode length and observation layout of SURVEY Appendix C."""
    _gpu()
    env = envs.make("gym_sted:ContextualMOSTED-easy-v0")
    assert env.spec.max_episode_steps == 10
    np.testing.assert_allclose(env.action_space.low, [0, 0, 1e-6])
    np.testing.assert_allclose(env.action_space.high, np.array([0.35, 20e-6, 100e-6], dtype=np.float32))
    assert env.observation_space[0].shape == (64, 64, 3) and env.observation_space[0].dtype == np.uint16
    assert env.observation_space[1].shape == (60,) and env.observation_space[1].dtype == np.float32
    (img, vec), info = env.reset(seed=0)
    assert img.shape == (64, 64, 3) and img.dtype == np.uint16 and vec.shape == (60,) and vec.dtype == np.float32
    assert img[..., 0].any() and not img[..., 1:].any() and not vec.any() and info == {}
    for t in range(10):
        action = np.array([0.1, 12e-6, 20e-6], dtype=np.float32) * (1 + 0.05 * t)
        (img, vec), reward, done, truncated, info = env.step(action)
        assert set(info) == {"action", "bleached", "sted_image", "conf1", "conf2", "fg_c", "fg_s", "mo_objs", "reward",
                             "f1-score", "nanodomains-coords"}
        assert truncated is False and done == (t == 9) and 0.0 <= reward <= 1.0 and reward == info["f1-score"]
        assert img.dtype == np.uint16 and np.array_equal(img[..., 1], info["conf1"].astype(np.uint16))
        assert np.array_equal(img[..., 2], info["sted_image"].astype(np.uint16))
        mo = info["mo_objs"]
        assert 0 < mo[0] <= 300 and mo[1] < 1
        exp = np.concatenate([(np.clip(action, env.action_space.low, env.action_space.high) - [0, 0, 1e-6])
                              / [0.35, 20e-6, 99e-6], [(mo[0] - 40) / 140, mo[1], mo[2]]])
        np.testing.assert_allclose(vec[6 * t:6 * t + 6], exp.astype(np.float32), rtol=1e-5, atol=1e-6)
        assert not vec[6 * (t + 1):].any()
        assert info["bleached"].shape == (64, 64) and info["fg_s"].dtype == bool
    # out-of-range actions are clipped, not rejected
    env.reset()
    _, _, _, _, info = env.step(np.array([5.0, 1.0, 1.0], dtype=np.float32))
    np.testing.assert_allclose(info["action"], env.action_space.high)


def test_zero_copy_observation_is_the_same_observation():
    """``zero_copy_observation=True`` hands out views of two alternating pinned buffers: same values as the copying
    environment, and the array of step t is untouched by step t + 1."""
    _gpu()
    a = envs.make_vec("ContextualMOSTED-easy-hslb-v0", 4, seed=3)
    b = envs.make_vec("ContextualMOSTED-easy-hslb-v0", 4, seed=3, zero_copy_observation=True)
    a.reset(seed=3); b.reset(seed=3)
    acts = np.random.default_rng(0).uniform(a.action_space.low, a.action_space.high, (3, 4, 3)).astype(np.float32)
    kept = []
    for t in range(3):
        (img_a, vec_a), r_a, *_ = a.step(acts[t])
        (img_b, vec_b), r_b, *_ = b.step(acts[t])
        assert np.array_equal(img_a, img_b) and np.array_equal(vec_a, vec_b) and np.array_equal(r_a, r_b)
        assert not img_b.flags.owndata
        kept.append((img_a, img_b))
        if t:
            assert np.array_equal(*kept[t - 1])                  # the previous step's view is still intact


def test_vector_env_routine_and_per_env_fluorophores():
    _gpu()
    vec = envs.make_vec("ContextualMOSTED-easy-hslb-v0", 6, compute_f1=False)
    assert vec.bleach_sampler.sample()["k1"] == pytest.approx(3.626623963067252e-16)
    (img, v), _ = vec.reset(seed=1)
    assert img.shape == (6, 64, 64, 3) and v.shape == (6, 60)
    acts = np.tile(np.array([[0.15, 10e-6, 30e-6]], dtype=np.float32), (6, 1))
    before = vec.frames.clone()
    (img, v), reward, done, trunc, info = vec.step(acts)
    assert reward.shape == (6,) and done.shape == (6,) and info["mo_objs"].shape == (6, 3)
    assert float(info["bleached"].sum()) < float(before.sum())                               # molecules were lost
    hard = envs.make_vec("ContextualMOSTED-hard-v0", 4, compute_f1=False, seed=3)
    (img, v), _ = hard.reset(seed=3)
    (img, v), reward, done, trunc, info = hard.step(acts[:4])
    assert np.isfinite(info["mo_objs"][:, :2]).all()
    # closed-form sampler: sigma_abs and k1 inside the ranges spanned by the reference's fitted routines x the criteria
    f = hard.bleach_sampler.sample(hard.microscope)
    assert 1e-23 < f["sigma_abs"][635] < 2e-20 and 1e-18 < f["k1"] < 1e-13


def test_preference_count_rate_env():
    """Config 4: 96x96 crops, count-rate loop on the confocal power, PrefNet reward with the -10 penalty."""
    _gpu()
    vec = envs.make_vec("PreferenceCountRateMOSTED-hard-v0", 4, seed=2)
    assert vec.observation_space[0].shape == (96, 96, 3) and vec.observation_space[1].shape == (1 + 6 * 30,)
    (img, v), _ = vec.reset(seed=2)
    assert img.shape == (4, 96, 96, 3) and v.shape == (4, 181) and not v.any()
    assert np.all(vec.conf_p_ex <= defaults.P_EX) and np.all(vec.conf_p_ex > 0)
    # the loop enforced the count rate on its last acquisition; the returned state is a fresh Poisson draw at that
    # power (preference_sted_env.py:200-203), so it may overshoot by shot noise only
    assert np.all(img[..., 0].reshape(4, -1).max(1) / defaults.PDT <= 1.25 * 20e6)
    acts = np.tile(np.array([[0.08, 5e-6, 15e-6]], dtype=np.float32), (4, 1))
    acts[3] = [0.0, 20e-6, 100e-6]                                                # confocal at full power, long dwell
    (img, v), reward, done, trunc, info = vec.step(acts)
    # the reference's keys (preference_sted_env.py:299-311) + the vector environments' per-env validity mask
    assert set(info) == {"action", "bleached", "sted_image", "conf1", "conf2", "fg_c", "fg_s", "mo_objs", "reward", "sample",
                         "conf_params", "invalid"}
    np.testing.assert_allclose(v[:, 0], vec.conf_p_ex / defaults.P_EX, rtol=1e-6)
    rate = info["sted_image"].amax(dim=(1, 2)).cpu().numpy() / acts[:, 2]
    assert np.array_equal(reward == -10.0, rate > 20e6)
    ok = reward != -10.0
    art = rewards.PreferenceArticulator()
    np.testing.assert_allclose(reward[ok], art.articulate(info["mo_objs"][ok])[0].ravel(), rtol=1e-6)
    assert not done.any() and vec.max_episode_steps == 30


def test_nanodomain_candidates_bit_exact():
    """Integer detection outputs (peak coordinates, in skimage's order) against the oracle on recorded STED images,
    simulated acquisitions of several sizes and synthetic spot images."""
    _gpu()
    demo = helpers.load_demonstrations()
    images = [demo["sted_image"].astype(np.int64)]
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[:64, :64]
    spots = []
    for _ in range(12):
        img = rng.poisson(0.3, (64, 64)).astype(np.int64)
        for (y, x) in rng.integers(5, 59, (7, 2)):
            img += rng.poisson(rng.uniform(30, 150) * np.exp(-((yy - y) ** 2 + (xx - x) ** 2) / (2 * rng.uniform(1.0, 2.5) ** 2)))
        spots.append(img)
    images.append(np.stack(spots))
    images.append(np.zeros((2, 64, 64), dtype=np.int64))                          # empty image: no candidates
    from gym_sted_b200.batch import BatchedMicroscope
    from gym_sted_b200 import synapse
    m = helpers.product_microscope(noise=True, background=defaults.DETECTOR["background"])
    for size, B in ((96, 4), (256, 3)):
        maps, _ = synapse.generate_batch(B, size, size, seed=size)
        bm = BatchedMicroscope(m, B, size, size, seed=1)
        bm.set_datamaps(torch.from_numpy(maps))
        sted, _ = bm.get_signal_and_bleach(15e-6, 8e-6, 0.12, bleach=True)
        images.append(np.rint(sted.cpu().numpy()).astype(np.int64))
    total = 0
    for stack in images:
        got = rewards.nanodomain_candidates(torch.from_numpy(stack).float().cuda())
        for b in range(len(stack)):
            ref = ro.candidate_peaks(stack[b])
            assert np.array_equal(got[b], ref.reshape(-1, 2)), (stack.shape, b, got[b], ref)
            total += len(ref)
    assert total > 100
    # end to end: batched F1 equals the oracle's per-image F1
    truth = [rng.integers(5, 59, (6, 2)) for _ in range(len(spots))]
    f1 = rewards.nanodomain_f1_batch(torch.from_numpy(np.stack(spots)).float().cuda(), truth)
    ref = np.array([ro.nanodomain_f1(spots[i], truth[i]) for i in range(len(spots))])
    assert np.array_equal(f1, ref)


def test_gaussfit_validation_matches_minpack():
    """The device Levenberg-Marquardt fit against scipy.optimize.leastsq (the reference's solver, through the
    oracle) on every candidate of many images: same keep/reject decision.  The fitted parameters differ only through
    the last bit of exp() (tests/test_lm_minpack.py pins the solver itself bit for bit), which the forward-difference
    Jacobian and the loose ftol = 1e-3 amplify to ~1e-6 typically, more on ill-conditioned windows."""
    _gpu()
    import scipy.optimize
    rng = np.random.default_rng(11)
    yy, xx = np.mgrid[:64, :64]
    imgs = []
    for i in range(256):
        img = rng.poisson(rng.uniform(0.1, 1.0), (64, 64)).astype(np.int64)
        for (y, x) in rng.integers(0, 64, (rng.integers(3, 12), 2)):       # includes spots cut by the border (zero padding)
            s = rng.uniform(0.6, 4.5, 2)
            img += rng.poisson(rng.uniform(10, 200) * np.exp(-((yy - y) ** 2 / (2 * s[0] ** 2) + (xx - x) ** 2 / (2 * s[1] ** 2))))
        imgs.append(img)
    stack = np.stack(imgs)
    dev = torch.from_numpy(stack).float().cuda()
    sted, peaks, npk, flags = rewards._candidates_device(dev)
    valid, params = rewards.gaussfit_validate(sted, peaks, npk, want_params=True)
    peaks, npk, flags, valid, params = (t.cpu().numpy() for t in (peaks, npk, flags, valid, params))
    X, Y = np.indices((7, 7)) - 3.0
    nfit = nagree = 0
    diffs = []
    for b in range(len(stack)):
        if flags[b]:
            continue
        padded = np.pad(stack[b], 3, mode="constant")
        for k in range(npk[b]):
            r, c = peaks[b, k]
            data = padded[r:r + 7, c:c + 7].astype(np.float64)

            def err(p):
                wx, wy = float(p[3]) + 1e-3, float(p[4]) + 1e-3
                return np.ravel(p[0] * np.exp(-1 * ((p[1] - X) ** 2.0 / (2 * wx ** 2.0) + (p[2] - Y) ** 2.0 / (2 * wy ** 2.0))) - data)

            p, _ = scipy.optimize.leastsq(err, [data.max(), 0, 0, 1, 1], ftol=1e-3)
            fw = 2.355 * p[3:5] * defaults.PIXELSIZE
            ok = not np.any((fw > 200e-9) | (fw < 40e-9))
            nfit += 1
            margin = np.min(np.abs(np.concatenate([fw - 200e-9, fw - 40e-9]))) / 40e-9
            if bool(valid[b, k]) == ok:
                nagree += 1
            else:
                assert margin < 1e-6, (b, k, p, params[b, k])            # only a width sitting on a limit may flip
            diffs.append(np.max(np.abs(params[b, k] - p) / (np.abs(p) + 1e-3)))
    assert nfit > 500 and nagree == nfit, (nfit, nagree)
    diffs = np.array(diffs)
    print(f'{nfit} fits; parameter difference quantiles 50/90/99/100 %: {np.quantile(diffs, [0.5, 0.9, 0.99, 1.0])}; >1e-6: {np.sum(diffs > 1e-6)}')
    assert np.median(diffs) < 1e-5 and np.mean(diffs > 1e-3) < 0.03, np.quantile(diffs, [0.5, 0.9, 0.99, 1.0])
    # the batched detector equals the oracle's detector (candidates + scipy fit) image by image
    got = rewards.detect_nanodomains_batch(dev)
    for b in range(0, len(stack), 7):
        assert np.array_equal(got[b], ro.detect_nanodomains(stack[b]).reshape(-1, 2)), b


def test_lm_solver_device_equals_host_build():
    """The solver on the GPU against the g++ build of the same header (itself bit-exact against scipy's MINPACK in
    tests/test_lm_minpack.py): identical bits on the transcendental-free model, solution and info code."""
    _gpu()
    import ctypes
    import subprocess
    import tempfile
    from gym_sted_b200 import _lib
    from tests import test_lm_minpack as tl
    out = os.path.join(tempfile.mkdtemp(), "liblm.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tests", "lm_harness.cpp")])
    host = ctypes.CDLL(out)
    host.lm_harness_fit.restype = ctypes.c_int
    host.lm_harness_fit.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_void_p]
    ws = tl.windows(500, 3)
    data = torch.from_numpy(np.stack([w.ravel() for w in ws])).cuda()
    for ftol in (1e-3, 1e-10):
        params = torch.from_numpy(np.stack([[w.max(), 0, 0, 1, 1.0] for w in ws])).cuda()
        info = torch.zeros(len(ws), dtype=torch.int32, device="cuda")
        _lib.check(_lib.load().sted_lmfit_test(data.data_ptr(), len(ws), 1, ftol, params.data_ptr(), info.data_ptr(), None))
        params, info = params.cpu().numpy(), info.cpu().numpy()
        for i, w in enumerate(ws):
            x, rc = tl.fit_harness(host, w, 1, ftol)
            assert rc == info[i] and np.array_equal(x, params[i]), (i, x, params[i])


def test_resolution_reproduces_recorded_values_exactly():
    """Decorrelation kernel against the resolutions the reference recorded for these images (bit-exact)."""
    _gpu()
    demo = helpers.load_demonstrations()
    res = rewards.resolution(torch.from_numpy(demo["sted_image"]).to(torch.float32).cuda())
    assert np.array_equal(res, demo["mo_objs"][:, 0])


def test_resolution_matches_oracle_on_other_sizes():
    _gpu()
    rng = np.random.default_rng(3)
    yy, xx = np.mgrid[:96, :96]
    imgs = []
    for s in (1.5, 3.0, 6.0):
        img = np.zeros((96, 96))
        for _ in range(25):
            y, x = rng.uniform(10, 86, 2)
            img += 80 * np.exp(-((yy - y) ** 2 + (xx - x) ** 2) / (2 * s ** 2))
        imgs.append(rng.poisson(img + 0.2))
    got = rewards.resolution(torch.from_numpy(np.stack(imgs)).to(torch.float32).cuda())
    ref = np.array([ro.resolution(i) for i in imgs])
    np.testing.assert_allclose(got, ref, rtol=1e-12)
    assert got[0] < got[1] < got[2] <= 300


def test_sequence_env_carries_bleaching_and_terminates_on_empty_datamap():
    """SequenceMOSTED (sequence_sted_env.py:53-140): one structure per episode, the datamap keeps its photobleaching
    from step to step, the next observation is a confocal of the bleached structure, done when it is empty."""
    _gpu()
    assert {"SequenceMOSTED-easy-v0", "SequenceMOSTED-easy-hshb-v0", "SequenceMOSTED-mid-v0", "SequenceMOSTED-hard-v0"} <= set(envs.REGISTRY)
    vec = envs.make_vec("SequenceMOSTED-easy-hslb-v0", 6, seed=3)
    (img0, v0), _ = vec.reset(seed=3)
    coords0 = [c.copy() for c in (s.nanodomains_coords for s in vec.synapses)]
    total = [float(vec.bm.datamaps_roi().sum())]
    acts = np.tile(np.array([[0.15, 10e-6, 20e-6]], dtype=np.float32), (6, 1))         # STED power that bleaches visibly
    acts[5] = [0.0, 1e-6, 1e-6]                                                        # nearly harmless confocal
    for t in range(4):
        before = vec.bm.datamaps_roi().clone()
        (img, v), reward, done, trunc, info = vec.step(acts)
        after = vec.bm.datamaps_roi()
        assert torch.equal(info["bleached"], after) and bool((after <= before).all())    # same maps, only ever bleached
        total.append(float(after.sum()))
        for c, s in zip(coords0, vec.synapses):
            assert np.array_equal(c, s.nanodomains_coords)                             # same structure all episode
        assert not done.any()
    assert total[4] < total[3] < total[2] < total[1] < total[0] and total[4] < 0.85 * total[0], total
    # channel 0 of the next observation is a confocal of the *bleached* structure: dimmer than the first one
    assert img[:5, ..., 0].astype(np.int64).sum() < 0.9 * img0[:5, ..., 0].astype(np.int64).sum()
    assert abs(int(img[5, ..., 0].astype(np.int64).sum()) - int(img0[5, ..., 0].astype(np.int64).sum())) < 0.2 * img0[5, ..., 0].sum()
    # an empty datamap ends the episode of that environment only
    roi = vec.bm.datamaps_roi().clone()
    roi[2] = 0
    vec.bm.set_datamaps(roi)
    _, _, done, _, _ = vec.step(acts)
    assert done.tolist() == [False, False, True, False, False, False]
    # single-environment view
    env = envs.make("SequenceMOSTED-easy-v0")
    (img, v), _ = env.reset(seed=1)
    for t in range(10):
        (img, v), r, d, tr, info = env.step(np.array([0.1, 10e-6, 20e-6], dtype=np.float32))
        assert d == (t == 9) and tr is False and info["bleached"].shape == (64, 64)


def test_detection_and_f1_match_oracle():
    _gpu()
    rng = np.random.default_rng(1)
    yy, xx = np.mgrid[:64, :64]
    for trial in range(4):
        img = rng.poisson(0.3, (64, 64)).astype(np.int64)
        truth = rng.integers(6, 58, (6, 2))
        for (y, x) in truth:
            img += rng.poisson(90 * np.exp(-((yy - y) ** 2 + (xx - x) ** 2) / (2 * 1.8 ** 2)))
        a = rewards.detect_nanodomains(img)
        b = ro.detect_nanodomains(img)
        assert np.array_equal(a, b)
        assert rewards.match_counts(truth, a) == ro.match_counts(truth, b)
        assert rewards.nanodomain_f1(img, truth) == ro.nanodomain_f1(img, truth)


def test_flat_components_follow_skimage_isolated_pixel_rule():
    """Constant plateaus make whole components 'all maxima'; skimage then keeps only the pixels a binary opening
    removes (mask xor opening).  Saturated blocks, bars and crosses, against the oracle (host restatement of skimage)."""
    _gpu()
    imgs = []
    base_img = np.zeros((64, 64), dtype=np.int64)
    a = base_img.copy(); a[10:16, 10:17] = 1000; a[40:42, 20:40] = 1000            # block + 2-pixel-thick bar
    b = base_img.copy(); b[30, 10:30] = 500; b[20:41, 20] = 500                    # thin cross
    c = base_img.copy(); c[5:9, 5:9] = 70; c[50:54, 50:60] = 70; c[20:24, 30:33] = 70
    d = np.full((64, 64), 7, dtype=np.int64)                                       # constant image: no candidates
    imgs = [a, b, c, d]
    got = rewards.nanodomain_candidates(torch.from_numpy(np.stack(imgs)).float().cuda())
    n = 0
    for g, img in zip(got, imgs):
        ref = ro.candidate_peaks(img).reshape(-1, 2)
        assert np.array_equal(g, ref), (g, ref)
        n += len(ref)
    assert n > 0


def test_long_and_interleaved_components_label_like_skimage():
    """The detector labels the mask by parallel minimum-label propagation: long winding components (many propagation
    rounds), components whose raster ranges interleave, diagonal-only (8-connected) links and more than 256 threads'
    worth of mask pixels, non-square images - peak coordinates and their order against the oracle."""
    _gpu()
    rng = np.random.default_rng(17)
    imgs = []
    H, W = 160, 224
    snake = rng.poisson(0.2, (H, W)).astype(np.int64)
    y, x = 10, 5
    for k in range(330):                                   # a one-pixel-wide staircase / zig-zag across the image
        snake[y, x] = 200 + (k * 37) % 90
        if k % 3 == 0:
            y += 1 if (k // 60) % 2 == 0 else -1
            y = min(max(y, 2), H - 3)
        x = min(x + (1 if k % 3 else 0) + (1 if k % 7 == 0 else 0), W - 2) if k % 3 else x
    imgs.append(snake)
    combs = rng.poisson(0.2, (H, W)).astype(np.int64)      # two interleaved combs: raster ranges of the components overlap
    combs[20, 20:200] = 300; combs[60, 20:200] = 280
    for c in range(24, 200, 8):
        combs[20:38, c] = 250 + c % 40
        combs[43:60, c + 4] = 240 + c % 30
    imgs.append(combs)
    diag = rng.poisson(0.2, (H, W)).astype(np.int64)       # diagonal chains: connected only through corners
    for k in range(120):
        diag[15 + k, 30 + k] = 150 + (k * 13) % 50
        diag[140 - k, 60 + k] = 170 + (k * 11) % 60
    imgs.append(diag)
    got = rewards.nanodomain_candidates(torch.from_numpy(np.stack(imgs)).float().cuda())
    n = 0
    for g, img in zip(got, imgs):
        ref = ro.candidate_peaks(img).reshape(-1, 2)
        assert np.array_equal(g, ref), (g, ref)
        n += len(ref)
    assert n >= 6


def test_resolution_at_the_metric_size_matches_oracle():
    """256 x 256 (BASELINE metric size): the three fused decorrelation sweeps against the oracle's 22 passes."""
    _gpu()
    rng = np.random.default_rng(8)
    yy, xx = np.mgrid[:256, :256]
    imgs = []
    for s in (1.2, 2.5, 5.0):
        img = np.zeros((256, 256))
        for _ in range(120):
            y, x = rng.uniform(10, 246, 2)
            img += 60 * np.exp(-((yy - y) ** 2 + (xx - x) ** 2) / (2 * s ** 2))
        imgs.append(rng.poisson(img + 0.2))
    imgs.append(np.zeros((256, 256), dtype=np.int64))              # empty image: capped resolution
    got = rewards.resolution(torch.from_numpy(np.stack(imgs)).to(torch.float32).cuda())
    ref = np.array([ro.resolution(i) for i in imgs])
    np.testing.assert_allclose(got, ref, rtol=1e-12)
    assert got[0] < got[1] < got[2] <= 300


def test_detection_capacity_overflow_is_an_error_not_a_fallback():
    _gpu()
    from gym_sted_b200 import _lib
    rng = np.random.default_rng(0)
    big = rng.poisson(5.0, (1, 700, 700)).astype(np.float32)        # 1 % of 490 000 pixels > 4096 mask pixels
    with pytest.raises(_lib.StedError):
        rewards.nanodomain_candidates(torch.from_numpy(big).cuda())


def test_warp_parallel_fit_equals_single_thread_fit():
    """``gaussfit_kernel`` runs one fit per warp (residuals spread over the lanes, solver state replicated); it must
    return exactly what the one-thread solver (``sted_lmfit_test``, model 0) returns for the same windows."""
    _gpu()
    from gym_sted_b200 import _lib
    rng = np.random.default_rng(21)
    yy, xx = np.mgrid[:64, :64]
    imgs = []
    for _ in range(48):
        img = rng.poisson(0.4, (64, 64)).astype(np.int64)
        for (y, x) in rng.integers(0, 64, (8, 2)):
            s = rng.uniform(0.7, 4.0, 2)
            img += rng.poisson(rng.uniform(15, 180) * np.exp(-((yy - y) ** 2 / (2 * s[0] ** 2) + (xx - x) ** 2 / (2 * s[1] ** 2))))
        imgs.append(img)
    stack = np.stack(imgs)
    sted, peaks, npk, flags = rewards._candidates_device(torch.from_numpy(stack).float().cuda())
    valid, params = rewards.gaussfit_validate(sted, peaks, npk, want_params=True)
    peaks_h, npk_h, params_h = peaks.cpu().numpy(), npk.cpu().numpy(), params.cpu().numpy()
    wins, starts, where = [], [], []
    for b in range(len(stack)):
        padded = np.pad(stack[b], 3, mode="constant").astype(np.float64)
        for k in range(npk_h[b]):
            r, c = peaks_h[b, k]
            w = padded[r:r + 7, c:c + 7]
            wins.append(w.ravel()); starts.append([w.max(), 0, 0, 1, 1.0]); where.append((b, k))
    assert len(wins) > 100
    d = torch.from_numpy(np.stack(wins)).cuda()
    x = torch.from_numpy(np.array(starts)).cuda()
    info = torch.zeros(len(wins), dtype=torch.int32, device="cuda")
    _lib.check(_lib.load().sted_lmfit_test(d.data_ptr(), len(wins), 0, 1e-3, x.data_ptr(), info.data_ptr(), None))
    x = x.cpu().numpy()
    for i, (b, k) in enumerate(where):
        assert np.array_equal(x[i], params_h[b, k]), (b, k, x[i], params_h[b, k])


def test_cooperative_solver_equals_one_thread_solver_bit_for_bit():
    """``lm_fit_warp`` (lm_minpack_warp.cuh: rows owned by lanes, reductions in MINPACK's order) against ``lm_fit`` on the
    device: same solutions and info codes, bit for bit, for both models, at the reference's ftol and at a tight one,
    including the all-zero and the flat window (singular Jacobians, the general ENORM path) and windows with huge and
    tiny counts."""
    _gpu()
    from gym_sted_b200 import _lib
    from tests import test_lm_minpack as tl
    ws = tl.windows(400, 5)
    rng = np.random.default_rng(6)
    ws += [w * s for w, s in zip(tl.windows(20, 7), rng.choice([1e-25, 1e-12, 1e9, 1e21], 22))]
    one = np.zeros((7, 7)); one[3, 3] = 50.0
    ws += [one, np.roll(one, 2, axis=0), rng.poisson(0.3, (7, 7)).astype(np.float64)]
    data = torch.from_numpy(np.stack([w.ravel() for w in ws])).cuda()
    for ftol in (1e-3, 1e-10):
        for serial, coop in ((0, 2), (1, 3)):
            out = []
            for model in (serial, coop):
                params = torch.from_numpy(np.stack([[w.max(), 0, 0, 1, 1.0] for w in ws])).cuda()
                info = torch.full((len(ws),), -7, dtype=torch.int32, device="cuda")
                _lib.check(_lib.load().sted_lmfit_test(data.data_ptr(), len(ws), model, ftol, params.data_ptr(), info.data_ptr(), None))
                out.append((params.cpu().numpy(), info.cpu().numpy()))
            assert np.array_equal(out[0][1], out[1][1]), np.flatnonzero(out[0][1] != out[1][1])
            same = (out[0][0] == out[1][0]) | (np.isnan(out[0][0]) & np.isnan(out[1][0]))
            assert same.all(), np.flatnonzero(~same.all(axis=1))


def test_device_f1_matching_equals_hungarian():
    """``sted_f1_match`` (maximum matching under the distance threshold, on the device) against the Hungarian matching
    on the priced-out cost matrix (``metrics.CentroidDetectionError(..., "hungarian")`` as restated by
    ``rewards.match_counts``): TP / FP / FN identical, F1 identical in float64 — with contested pairs (several detections
    within reach of one truth and vice versa, chains that need an augmenting path), invalid candidates, empty lists."""
    from gym_sted_b200 import _lib
    _gpu()
    lib = _lib.load()
    rng = np.random.default_rng(7)
    B, maxp = 64, 256
    peaks = np.zeros((B, maxp, 2), dtype=np.int32)
    npk = np.zeros(B, dtype=np.int32)
    valid = np.zeros((B, maxp), dtype=np.uint8)
    truths = []
    for b in range(B):
        nt = int(rng.integers(0, 16)) if b else 0
        t = rng.integers(4, 60, (nt, 2))
        if b % 3 == 1 and nt:                                        # dense clusters: many pairs closer than 2 px
            t = (t[0] + rng.integers(-2, 3, (nt, 2))).clip(0, 63)
        truths.append(t)
        ng = int(rng.integers(0, 40)) if b != 1 else 0
        g = rng.integers(0, 64, (ng, 2))
        k = min(nt, ng)
        if k:
            g[:k] = t[:k] + rng.integers(-1, 2, (k, 2))              # near misses of the truths
        if b % 5 == 2 and nt >= 2 and ng >= 2:                       # a chain: g0 reaches t0 and t1, g1 only t0
            t[0], t[1] = (20, 20), (20, 22)
            g[0], g[1] = (20, 21), (19, 20)
        peaks[b, :ng] = g
        npk[b] = ng
        valid[b, :ng] = rng.random(ng) < 0.8
    dt = rewards.DeviceTruths(truths, "cuda")
    d = lambda a: torch.from_numpy(a).cuda()
    dp, dn, dv = d(peaks), d(npk), d(valid)
    counts = torch.empty((B, 3), dtype=torch.int32, device="cuda")
    f1 = torch.empty((B,), dtype=torch.float64, device="cuda")
    flags = torch.empty((B,), dtype=torch.int32, device="cuda")
    _lib.check(lib.sted_f1_match(dp.data_ptr(), dn.data_ptr(), dv.data_ptr(), maxp, dt.coords.data_ptr(), dt.counts.data_ptr(),
                                 dt.coords.shape[1], B, 2.0, counts.data_ptr(), f1.data_ptr(), flags.data_ptr(), None))
    torch.cuda.synchronize()
    counts, f1 = counts.cpu().numpy(), f1.cpu().numpy()
    assert not flags.cpu().numpy().any()
    contested = 0
    for b in range(B):
        g = peaks[b, :npk[b]][valid[b, :npk[b]].astype(bool)]
        ref = rewards.match_counts(truths[b], g)
        assert tuple(counts[b]) == ref, (b, counts[b], ref)
        assert f1[b] == rewards.f1_from_counts(*ref)
        dist = np.sqrt(((np.asarray(truths[b])[:, None, :] - g[None, :, :]) ** 2).sum(-1)) if len(truths[b]) and len(g) else np.zeros((0, 0))
        contested += int(((dist < 2).sum(0) > 1).any() or ((dist < 2).sum(1) > 1).any()) if dist.size else 0
    assert contested >= 10                                           # the augmenting-path code was exercised
    # valid_dev = NULL keeps every candidate
    c2 = torch.empty((B, 3), dtype=torch.int32, device="cuda")
    f2 = torch.empty((B,), dtype=torch.float64, device="cuda")
    _lib.check(lib.sted_f1_match(dp.data_ptr(), dn.data_ptr(), None, maxp, dt.coords.data_ptr(), dt.counts.data_ptr(),
                                 dt.coords.shape[1], B, 2.0, c2.data_ptr(), f2.data_ptr(), None, None))
    torch.cuda.synchronize()
    for b in (2, 7, 30):
        assert tuple(c2[b].cpu().numpy()) == rewards.match_counts(truths[b], peaks[b, :npk[b]])


def test_f1_batch_runs_on_the_device_and_matches_the_host_path():
    _gpu()
    rng = np.random.default_rng(3)
    yy, xx = np.mgrid[:64, :64]
    B = 5
    imgs, truths = [], []
    for b in range(B):
        img = rng.poisson(0.3, (64, 64)).astype(np.float32)
        t = rng.integers(6, 58, (5, 2))
        for (y, x) in t:
            img += rng.poisson(90 * np.exp(-((yy - y) ** 2 + (xx - x) ** 2) / (2 * 1.8 ** 2)))
        imgs.append(img)
        truths.append(t)
    sted = torch.from_numpy(np.stack(imgs)).cuda()
    f1, counts = rewards.nanodomain_f1_counts(sted, truths)
    for b in range(B):
        ref = rewards.match_counts(truths[b], rewards.detect_nanodomains(imgs[b]))
        assert tuple(counts[b]) == ref
        assert f1[b] == rewards.f1_from_counts(*ref)
    assert np.array_equal(rewards.nanodomain_f1_batch(sted, rewards.pad_coords(truths)), f1)
    assert np.array_equal(rewards.nanodomain_f1_batch(sted, rewards.DeviceTruths(truths)), f1)
