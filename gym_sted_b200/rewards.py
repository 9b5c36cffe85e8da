"""Per-step foregrounds, objectives and rewards for batches of environments (SURVEY.md §8a rows a6–a10).

* ``foreground_objectives``  Otsu foregrounds + Bleach + SNR: one CUDA kernel (``rewards_kernels.cu``),
  bit-identical to ``gym_sted/utils.py:16-24``, ``mosted_env.py:237-244``, ``objectives.py:91-111``.
* ``resolution``             image-decorrelation resolution (``objectives.py:121-291``): ``apodize_kernel`` ->
  cuFFT (complex128) -> ``spectrum_gather_kernel`` -> ``decorr_kernel`` (all 22 decorrelation passes and their peak
  searches, one CTA per image, float64); nothing runs on the host.
* ``nanodomain_f1``          detection + F1 (``objectives.py:404-430``, ``rewards/utils.py:6-49``): candidates
  (``nanodomain_kernel``), Gaussian-fit validation (``gaussfit_kernel``, MINPACK restated) and the matching +
  F1 (``f1_match_kernel``: maximum matching of the pairs closer than the threshold, which is what the
  reference's Hungarian solution counts) all run on the device; one packed device->host read per step.
* ``PrefNet`` / ``PreferenceArticulator``  the 3-10-10-1 ReLU MLP of ``gym_sted/prefnet`` on the device.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import _lib

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
RES_CAP = 300.0          # gym_sted/defaults.py:117
PIXELSIZE = 20e-9


# ------------------------------------------------------------------------------------------------
# a6 + a7
# ------------------------------------------------------------------------------------------------
_HIST_SCRATCH: dict = {}


def foreground_objectives_device(conf1: torch.Tensor, sted: torch.Tensor, conf2: torch.Tensor):
    """``objectives_kernel`` with nothing read back: (B,H,W) float32 photon counts -> fg_s, fg_c (uint8 0/1), objs
    (B,2) float64 = {Bleach, SNR}, flags (B,) int32 (bit 0: SNR < 0, the reference's None), all device tensors.  The
    64 Ki-bin histogram scratch (only touched by images spanning more than 8192 values) is kept per (B, device)."""
    lib = _lib.load()
    B, H, W = conf1.shape
    dev = conf1.device
    conf1, sted, conf2 = (t.contiguous().to(torch.float32) for t in (conf1, sted, conf2))
    fg_c = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    fg_s = torch.empty_like(fg_c)
    objs = torch.empty((B, 2), dtype=torch.float64, device=dev)
    flags = torch.empty((B,), dtype=torch.int32, device=dev)
    key = (B, str(dev))
    if key not in _HIST_SCRATCH:
        _HIST_SCRATCH.clear()                                         # one batch shape at a time: 0.5 MB per environment
        _HIST_SCRATCH[key] = torch.empty((B, 2, 65536), dtype=torch.int32, device=dev)
    _lib.check(lib.sted_foreground_objectives(conf1.data_ptr(), sted.data_ptr(), conf2.data_ptr(), B, H, W,
                                              fg_c.data_ptr(), fg_s.data_ptr(), objs.data_ptr(), flags.data_ptr(),
                                              _HIST_SCRATCH[key].data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    return fg_s, fg_c, objs, flags


def foreground_objectives(conf1: torch.Tensor, sted: torch.Tensor, conf2: torch.Tensor):
    """(B,H,W) float32 photon counts -> fg_s, fg_c (bool), bleach, snr (float64, NaN where the reference
    returns None), flags (int32)."""
    fg_s, fg_c, objs, flags = foreground_objectives_device(conf1, sted, conf2)
    snr = torch.where((flags & 1) != 0, torch.full_like(objs[:, 1], float("nan")), objs[:, 1])
    return fg_s.view(torch.bool), fg_c.view(torch.bool), objs[:, 0], snr, flags


# ------------------------------------------------------------------------------------------------
# a8: resolution
# ------------------------------------------------------------------------------------------------
class _Decorrelator:
    """Geometry shared by every image of one shape: apodisation window, radius map, radius-sorted order."""

    def __init__(self, H, W, device, w=20):
        edge = (np.sin(np.linspace(-np.pi / 2, np.pi / 2, w)) + 1) / 2
        wins = [np.concatenate([edge, np.ones(l - 2 * w), edge[::-1]]) for l in (H, W)]
        self.window = torch.from_numpy(wins[0][:, None] * wins[1][None, :]).to(device)
        self.Hc, self.Wc = H - 1 + H % 2, W - 1 + W % 2
        ys = np.linspace(-1, 1, self.Hc)[:, None]
        xs = np.linspace(-1, 1, self.Wc)[None, :]
        R = np.sqrt(xs ** 2 + ys ** 2).ravel()
        order = np.argsort(R, kind="stable")
        self.R_sorted_np = R[order]
        self.order = torch.from_numpy(order).to(device)
        self.R_sorted = torch.from_numpy(self.R_sorted_np).to(device)
        self.R2_sorted = self.R_sorted ** 2
        self.inside = self.R_sorted < 1
        self.n_inside = int(np.sum(self.R_sorted_np < 1))
        self._r = {}

    def r_coarse(self, Nr):
        if Nr not in self._r:
            self._r[Nr] = torch.from_numpy(np.linspace(0, 1, Nr)).to(self.R_sorted.device)
        return self._r[Nr]


_DECORR_CACHE: dict = {}


def resolution_device(sted: torch.Tensor, pixelsize=PIXELSIZE, res_cap=RES_CAP, Nr=50, Ng=10, tol=0.0005) -> torch.Tensor:
    """(B,H,W) photon counts on the GPU -> (B,) float64 DEVICE tensor of resolutions in nm (capped); nothing is read
    back.  ``apodize_kernel`` (mean, window, float32 rounding, crop, ifftshift) -> cuFFT (through torch, complex128)
    -> ``spectrum_gather_kernel`` (centring, radius order, R < 1 mask, phase-only copy) -> ``decorr_kernel`` (the 22
    decorrelation passes and their peak searches, one CTA per image)."""
    assert Nr == 50 and Ng == 10, "decorr_kernel is compiled for the reference's Nr = 50, Ng = 10"
    B, H, W = sted.shape
    dev = sted.device
    key = (H, W, str(dev))
    if key not in _DECORR_CACHE:
        _DECORR_CACHE[key] = _Decorrelator(H, W, dev)
    geo = _DECORR_CACHE[key]
    lib = _lib.load()
    st = torch.cuda.current_stream(dev).cuda_stream
    img = sted.contiguous().to(torch.float32)
    N = geo.Hc * geo.Wc
    pre = torch.empty((B, geo.Hc, geo.Wc), dtype=torch.complex128, device=dev)
    _lib.check(lib.sted_apodize_for_fft(img.data_ptr(), geo.window.data_ptr(), B, H, W, geo.Hc, geo.Wc, pre.data_ptr(), st))
    spec = torch.fft.fftn(pre, dim=(1, 2))
    Ik = torch.empty((B, N), dtype=torch.complex128, device=dev)
    Ikn = torch.empty_like(Ik)
    _lib.check(lib.sted_spectrum_gather(spec.data_ptr(), geo.order.data_ptr(), B, geo.Hc, geo.Wc, geo.n_inside,
                                        Ik.data_ptr(), Ikn.data_ptr(), st))
    res = torch.empty((B,), dtype=torch.float64, device=dev)
    _lib.check(lib.sted_decorrelation_resolution(
        Ik.data_ptr(), Ikn.data_ptr(), B, N, geo.n_inside, geo.R_sorted.data_ptr(), geo.R2_sorted.data_ptr(),
        geo.r_coarse(Nr).data_ptr(), geo.Hc / 4, max(geo.Hc, geo.Wc) / 2, float(pixelsize), float(res_cap), float(tol),
        res.data_ptr(), st))
    return res


def side_stream(dev, lane=0):
    """One of the package's side streams of ``dev`` (created on first use).  Lane 0 carries the nanodomain detection,
    lane 1 the decorrelation chain of a step that also detects nanodomains, lane 2 the next structures."""
    key = (str(dev), int(lane))
    if key not in _SIDE_STREAMS:
        # the short decorrelation chain goes ahead of the long grids of the fits and of the structure generator when
        # they compete for the SMs a step's acquisitions (one more priority step up, envs.py) leave free
        _SIDE_STREAMS[key] = torch.cuda.Stream(dev, priority=-1 if lane == 1 else 0)
    return _SIDE_STREAMS[key]


def resolution_launch(sted: torch.Tensor, lane=0):
    """``resolution_device`` queued on a side stream (it then runs beside whatever the caller queues next on its own
    stream, e.g. ``objectives_kernel``); returns ``(res, done)``: wait for the event ``done`` before using ``res`` on
    another stream, and keep ``sted`` alive until then."""
    dev = sted.device
    side = side_stream(dev, lane)
    cur = torch.cuda.current_stream(dev)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        res = resolution_device(sted)
        done = torch.cuda.Event()
        done.record(side)
    res.record_stream(cur)                    # read by the caller's stream after `done`
    return res, done


def resolution(sted: torch.Tensor, pixelsize=PIXELSIZE, res_cap=RES_CAP, Nr=50, Ng=10, tol=0.0005) -> np.ndarray:
    """``resolution_device`` read back as a (B,) float64 numpy array."""
    return resolution_device(sted, pixelsize, res_cap, Nr, Ng, tol).cpu().numpy()


# ------------------------------------------------------------------------------------------------
# a9: nanodomain detection + F1
# ------------------------------------------------------------------------------------------------
def _gauss_taps(sigma=0.5, truncate=4.0):
    """scipy.ndimage._filters._gaussian_kernel1d (order 0): the three distinct taps of the 5-tap kernel."""
    radius = int(truncate * sigma + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    phi = phi / phi.sum()
    return phi[radius], phi[radius + 1], phi[radius + 2]


def _quantile_index(n, q=0.99):
    """numpy.quantile's 'linear' virtual index (_compute_virtual_index with alpha = beta = 1) -> (k0, gamma)."""
    vi = np.float64(n) * np.float64(q) + (np.float64(1) + np.float64(q) * (np.float64(1) - np.float64(1) - np.float64(1))) - np.float64(1)
    prev = np.floor(vi)
    return int(prev), float(vi - prev)


def _candidates_device(sted: torch.Tensor):
    """Launch ``nanodomain_kernel``; returns the device tensors (sted_f32, peaks (B,256,2), npeaks (B,), flags (B,))."""
    lib = _lib.load()
    B, H, W = sted.shape
    dev = sted.device
    sted = sted.contiguous().to(torch.float32)
    scratch = torch.empty((2, B, H, W), dtype=torch.float64, device=dev)
    peaks = torch.empty((B, 256, 2), dtype=torch.int32, device=dev)
    npk = torch.empty((B,), dtype=torch.int32, device=dev)
    flags = torch.empty((B,), dtype=torch.int32, device=dev)
    import ctypes
    w = (ctypes.c_double * 3)(*_gauss_taps())
    k0, gamma = _quantile_index(H * W)
    _lib.check(lib.sted_nanodomain_candidates(sted.data_ptr(), B, H, W, w, k0, gamma, scratch.data_ptr(), peaks.data_ptr(),
                                              npk.data_ptr(), flags.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    return sted, peaks, npk, flags


def _check_capacity(flags: np.ndarray):
    if flags.any():
        b = int(np.flatnonzero(flags)[0])
        raise _lib.StedError(-4, f"nanodomain_kernel capacity exceeded for image {b} (flags={int(flags[b])}: bit0 >4096 mask "
                                 f"pixels, bit1 >64 maxima in one component, bit3 >256 peaks)")


def nanodomain_candidates(sted: torch.Tensor):
    """(B,H,W) photon counts on the GPU -> list of (n_i, 2) int arrays: peak_local_max candidates per image
    (gaussian -> q99 mask -> remove_small_objects -> label -> per-label maxima), computed by one CUDA kernel."""
    sted, peaks, npk, flags = _candidates_device(sted)
    peaks, npk, flags = peaks.cpu().numpy(), npk.cpu().numpy(), flags.cpu().numpy()
    _check_capacity(flags)
    return [peaks[b, :npk[b]].astype(np.int64) for b in range(sted.shape[0])]


def gaussfit_validate(sted: torch.Tensor, peaks: torch.Tensor, npeaks: torch.Tensor, pixelsize=PIXELSIZE, want_params=False):
    """Levenberg-Marquardt validation of device-resident candidates (``sted_gaussfit_validate``): returns the
    uint8 (B, max_peaks) validity tensor (and the fitted parameters when asked)."""
    lib = _lib.load()
    B, H, W = sted.shape
    sted = sted.contiguous().to(torch.float32)
    maxp = peaks.shape[1]
    valid = torch.zeros((B, maxp), dtype=torch.uint8, device=sted.device)
    params = torch.zeros((B, maxp, 5), dtype=torch.float64, device=sted.device) if want_params else None
    _lib.check(lib.sted_gaussfit_validate(sted.data_ptr(), B, H, W, peaks.data_ptr(), npeaks.data_ptr(), maxp, float(pixelsize),
                                          valid.data_ptr(), params.data_ptr() if want_params else None,
                                          torch.cuda.current_stream(sted.device).cuda_stream))
    return (valid, params) if want_params else valid


_SIDE_STREAMS: dict = {}


def detect_nanodomains_launch(sted: torch.Tensor, pixelsize=PIXELSIZE):
    """Enqueue candidates + Gaussian-fit validation of B device images on a side stream and return a handle for
    ``detect_nanodomains_collect``; the fits (few warps, long latency) then overlap whatever the caller runs next."""
    dev = sted.device
    side = side_stream(dev, 0)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        sted32, peaks, npk, flags = _candidates_device(sted)
        valid = gaussfit_validate(sted32, peaks, npk, pixelsize)
        done = torch.cuda.Event()
        done.record(side)
    return sted, sted32, peaks, npk, flags, valid, done


def detect_nanodomains_collect(handle):
    """Device->host read of the surviving coordinates of a ``detect_nanodomains_launch`` handle."""
    _, sted32, peaks, npk, flags, valid, done = handle
    done.synchronize()
    peaks, npk, flags, valid = peaks.cpu().numpy(), npk.cpu().numpy(), flags.cpu().numpy(), valid.cpu().numpy().astype(bool)
    _check_capacity(flags)
    return [peaks[b, :npk[b]][valid[b, :npk[b]]].astype(np.int64) for b in range(sted32.shape[0])]


def detect_nanodomains_batch(sted: torch.Tensor, pixelsize=PIXELSIZE):
    """``NumberNanodomains`` detector for B device images: candidates and Gaussian-fit validation on the GPU, one
    device->host read of the surviving coordinates."""
    return detect_nanodomains_collect(detect_nanodomains_launch(sted, pixelsize))


def detect_nanodomains(sted, pixelsize=PIXELSIZE) -> np.ndarray:
    """Single image (array or tensor) through the same device path."""
    t = torch.as_tensor(np.asarray(sted, dtype=np.float32) if not torch.is_tensor(sted) else sted)
    return detect_nanodomains_batch(t.to("cuda", torch.float32)[None], pixelsize)[0]


def pad_coords(coords, fill=1e9):
    """List of (n_i, 2) coordinate arrays -> ((B, max n, 2) float64 padded with ``fill``, (B,) counts)."""
    n = np.array([len(c) for c in coords], dtype=np.int64)
    out = np.full((len(coords), max(int(n.max()) if len(n) else 0, 1), 2), fill, dtype=np.float64)
    for b, c in enumerate(coords):
        if n[b]:
            out[b, :n[b]] = np.asarray(c, dtype=np.float64).reshape(-1, 2)
    return out, n


def _match_padded(T, nt, G, keep, threshold=2):
    """(B,3) int array of (TP, FP, FN) from padded truths ``T`` (rows >= nt are far away) and candidate detections ``G``
    with their ``keep`` mask.  Pairs closer than ``threshold`` form a bipartite graph and TP is its maximum matching
    (that is what the Hungarian solution on the masked costs counts); when no detection and no truth has more than one
    partner the matching is the edge set itself, so only images with a contested pair go through the Hungarian solver."""
    ng = keep.sum(axis=1)
    d = T[:, :, None, :] - G[:, None, :, :]
    close = (np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) < threshold) & keep[:, None, :]
    close &= (np.arange(T.shape[1])[None, :] < nt[:, None])[:, :, None]
    simple = (close.sum(axis=1).max(axis=1) <= 1) & (close.sum(axis=2).max(axis=1) <= 1)
    tp = close.sum(axis=(1, 2))
    for b in np.flatnonzero(~simple):
        tp[b] = match_counts(T[b, :nt[b]], G[b][keep[b]], threshold)[0]
    return np.stack([tp, ng - tp, nt - tp], axis=1).astype(np.int64)


def match_counts_batch(truths, guesses, threshold=2):
    """``match_counts`` for many images (lists of coordinate arrays): (B,3) int array of (TP, FP, FN)."""
    if len(truths) == 0:
        return np.zeros((0, 3), dtype=np.int64)
    T, nt = pad_coords(truths)
    G, ng = pad_coords(guesses, fill=-1e9)
    return _match_padded(T, nt, G, np.arange(G.shape[1])[None, :] < ng[:, None], threshold)


class DeviceTruths:
    """Ground-truth nanodomain coordinates of B structures on the device: int32 (B, max_truth, 2) + int32 (B,) counts
    (built once per structure; ``sted_f1_match`` reads them at every step)."""

    def __init__(self, coords, device="cuda"):
        n = np.array([len(c) for c in coords], dtype=np.int32)
        T = np.zeros((len(coords), max(int(n.max()) if len(n) else 0, 1), 2), dtype=np.int32)
        for b, c in enumerate(coords):
            if n[b]:
                T[b, :n[b]] = np.asarray(c, dtype=np.int64).reshape(-1, 2)
        self.coords, self.counts = torch.from_numpy(T).to(device), torch.from_numpy(n).to(device)

    @classmethod
    def from_device(cls, coords: torch.Tensor, counts: torch.Tensor):
        """Wrap tensors that already live on the device (``DeviceSynapseGenerator.generate``)."""
        self = cls.__new__(cls)
        self.coords, self.counts = coords.contiguous(), counts.contiguous()
        return self


def nanodomain_f1_launch(sted: torch.Tensor, truths, pixelsize=PIXELSIZE, threshold=2, handle=None):
    """Queue the F1 of B images on the current stream and return DEVICE tensors ``(f1 (B,) float64, counts (B,3) int32 =
    TP FP FN, detector flags (B,) int32, matcher flags (B,) int32)`` — nothing is read back, so a caller can fold them
    into its own end-of-step transfer.  ``truths``: ``DeviceTruths`` (or anything ``nanodomain_f1_counts`` accepts)."""
    lib = _lib.load()
    _, sted32, peaks, npk, flags, valid, done = handle if handle is not None else detect_nanodomains_launch(sted, pixelsize)
    dev = sted32.device
    if isinstance(truths, tuple):                                     # pad_coords output
        T, nt = truths
        truths = [T[b, :nt[b]] for b in range(len(nt))]
    if not isinstance(truths, DeviceTruths):
        truths = DeviceTruths(truths, dev)
    B = sted32.shape[0]
    cur = torch.cuda.current_stream(dev)
    cur.wait_event(done)                                              # the detection ran on the side stream
    counts = torch.empty((B, 3), dtype=torch.int32, device=dev)
    mflags = torch.empty((B,), dtype=torch.int32, device=dev)
    f1 = torch.empty((B,), dtype=torch.float64, device=dev)
    _lib.check(lib.sted_f1_match(peaks.data_ptr(), npk.data_ptr(), valid.data_ptr(), peaks.shape[1], truths.coords.data_ptr(),
                                 truths.counts.data_ptr(), truths.coords.shape[1], B, float(threshold), counts.data_ptr(),
                                 f1.data_ptr(), mflags.data_ptr(), cur.cuda_stream))
    for t in (peaks, npk, valid, truths.coords, truths.counts):       # read on `cur`, allocated on the side stream
        t.record_stream(cur)
    return f1, counts, flags, mflags


def check_f1_flags(dflags: np.ndarray, mflags: np.ndarray):
    """Raise for the capacity flags of the detector / matcher (host arrays)."""
    _check_capacity(np.asarray(dflags).astype(np.int64))
    if np.asarray(mflags).any():
        raise _lib.StedError(-4, f"f1_match_kernel capacity exceeded (more than 512 truths or detections) for image "
                                 f"{int(np.flatnonzero(mflags)[0])}")


def nanodomain_f1_counts(sted: torch.Tensor, truths, pixelsize=PIXELSIZE, threshold=2, handle=None):
    """F1 of B images entirely on the device (north_star: "nanocluster detection plus the F1 reward run on the device"):
    candidates, Gaussian-fit validation (already enqueued when ``handle`` is given), then ``f1_match_kernel`` — TP = size
    of a maximum matching of the (truth, detection) pairs closer than ``threshold``, which is what the reference's
    Hungarian solution on the priced-out cost matrix counts (``objectives.py:425-430``).  One packed device->host read.
    ``truths``: ``DeviceTruths``, a list of (n_i, 2) arrays, or the ``pad_coords`` tuple of such a list.
    Returns (f1 (B,) float64, counts (B, 3) int64 = TP, FP, FN)."""
    f1, counts, flags, mflags = nanodomain_f1_launch(sted, truths, pixelsize, threshold, handle)
    B = f1.shape[0]
    out = torch.empty((B, 6), dtype=torch.float64, device=f1.device)  # f1 | TP FP FN | detector flags | matcher flags
    out[:, 0] = f1
    out[:, 1:4] = counts
    out[:, 4] = flags
    out[:, 5] = mflags
    host = out.cpu().numpy()
    check_f1_flags(host[:, 4], host[:, 5])
    return host[:, 0].copy(), host[:, 1:4].astype(np.int64)


def nanodomain_f1_batch(sted: torch.Tensor, truths, pixelsize=PIXELSIZE, threshold=2, handle=None) -> np.ndarray:
    """(B,) F1 scores; see ``nanodomain_f1_counts``."""
    return nanodomain_f1_counts(sted, truths, pixelsize, threshold, handle)[0]


def match_counts(truth, guess, threshold=2):
    """(TP, FP, FN): Hungarian matching on Euclidean distances, pairs at distance >= threshold forbidden
    (``metrics.CentroidDetectionError(gt, guess, 2, algorithm="hungarian")``)."""
    import scipy.optimize
    import scipy.spatial

    truth = np.asarray(truth, dtype=np.float64).reshape(-1, 2)
    guess = np.asarray(guess, dtype=np.float64).reshape(-1, 2)
    if len(truth) == 0 or len(guess) == 0:
        return 0, len(guess), len(truth)
    cost = scipy.spatial.distance.cdist(truth, guess)
    masked = np.where(cost < threshold, cost, 1e6)
    ri, ci = scipy.optimize.linear_sum_assignment(masked)
    tp = int(np.sum(masked[ri, ci] < 1e6))
    return tp, len(guess) - tp, len(truth) - tp


def f1_from_counts(tp, fp, fn):
    prec = tp / (tp + fp + 1e-6)
    rec = tp / (tp + fn + 1e-6)
    return 2 * prec * rec / (prec + rec + 1e-6)


def nanodomain_f1(sted, truth_coords, pixelsize=PIXELSIZE, threshold=2):
    return f1_from_counts(*match_counts(truth_coords, detect_nanodomains(sted, pixelsize), threshold))


# ------------------------------------------------------------------------------------------------
# a10: preference network
# ------------------------------------------------------------------------------------------------
class PrefNet(torch.nn.Module):
    """``gym_sted/prefnet/prefNet.py:11-31``: Linear(nb_obj,10)-ReLU-Linear(10,10)-ReLU-Linear(10,1)."""

    def __init__(self, nb_obj=3, middle_size=10):
        super().__init__()
        self.f1 = torch.nn.Linear(nb_obj, middle_size)
        self.f2 = torch.nn.Linear(middle_size, middle_size)
        self.out = torch.nn.Linear(middle_size, 1)

    def forward(self, X):
        return self.out(torch.relu(self.f2(torch.relu(self.f1(X)))))


class PreferenceArticulator:
    """``gym_sted/prefnet/articulation.py:36-75`` with the weights of model ``2021-07-12-06-40-29_ResolutionBleachSNR``."""

    def __init__(self, nb_obj=3, model_name="2021-07-12-06-40-29_ResolutionBleachSNR", device="cpu"):
        path = os.path.join(_DATA, "prefnet_" + model_name)
        self.model = PrefNet(nb_obj=nb_obj)
        self.model.load_state_dict(torch.load(os.path.join(path, "weights.t7"), map_location="cpu"))
        self.model.eval().to(device)
        self.device = torch.device(device)
        with open(os.path.join(path, "config.json")) as fh:
            self.config = json.load(fh)

    def articulate(self, thetas, use_sigmoid=False):
        thetas = (np.array(thetas) - self.config["train_mean"]) / self.config["train_std"]
        with torch.no_grad():
            scores = self.model(torch.from_numpy(thetas).float().to(self.device)).cpu().numpy()
        if use_sigmoid:
            scores = torch.sigmoid(torch.tensor(scores)).numpy()
        order = np.argsort(scores.ravel())
        return scores, order[-1], order
