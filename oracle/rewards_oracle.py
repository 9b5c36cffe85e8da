"""CPU oracle for the per-step foreground / objective / detection computations — TEST INFRASTRUCTURE ONLY.

Restates, in float64 numpy + scipy, the reference's reward-side functions on the hot path
(SURVEY.md §8a rows a6–a9):

=================================================  ==========================================
reference                                          here
=================================================  ==========================================
``gym_sted/utils.py:16-24`` ``get_foreground``     ``threshold_otsu``, ``get_foreground``
``gym_sted/envs/mosted_env.py:237-244``            ``foregrounds``
``gym_sted/rewards/objectives.py:91-100``          ``signal_ratio``
``gym_sted/rewards/objectives.py:107-111``         ``bleach``
``gym_sted/rewards/objectives.py:121-291``         ``resolution`` / ``decorrelation``
``gym_sted/rewards/objectives.py:404-430``         ``detect_nanodomains``, ``f1_score``
``gym_sted/rewards/utils.py:6-49``                 ``fit_gaussian``, ``validate_positions``
=================================================  ==========================================

PINNED: ``tests/golden/demonstrations.npz`` (made from the reference's own recorded episodes by
``tests/golden/make_golden.py``) fixes the Otsu masks and the Resolution / Bleach / SNR objectives of
50 records exactly, and the F1 formula + matching rule through the recorded scores.
UNPINNED: the current detection routine (gaussian -> q99 -> remove_small_objects -> label ->
peak_local_max -> Gaussian-fit validation) depends on scikit-image, which is not installed here; its
semantics are restated from the scikit-image 0.19–0.22 sources as recalled, on top of scipy.ndimage.
The un-vendored ``metrics`` package (``setup.py:15``) is restated as Hungarian matching on Euclidean
distances with pairs at distance >= threshold forbidden.
"""
from __future__ import annotations

import numpy as np
import scipy.ndimage as ndi
import scipy.optimize
import scipy.spatial


# ------------------------------------------------------------------------------------------------
# a6: Otsu foreground (skimage.filters.threshold_otsu on integer images = exact bincount histogram)
# ------------------------------------------------------------------------------------------------
def threshold_otsu(image: np.ndarray):
    image = np.asarray(image)
    lo, hi = int(image.min()), int(image.max())
    if lo == hi:
        return lo
    hist = np.bincount((image.ravel() - lo).astype(np.int64), minlength=hi - lo + 1).astype(np.float64)
    centers = np.arange(lo, hi + 1, dtype=np.float64)
    weight1 = np.cumsum(hist)
    weight2 = np.cumsum(hist[::-1])[::-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        mean1 = np.cumsum(hist * centers) / weight1
        mean2 = (np.cumsum((hist * centers)[::-1]) / weight2[::-1])[::-1]
    variance12 = weight1[:-1] * weight2[1:] * (mean1[:-1] - mean2[1:]) ** 2
    return centers[int(np.argmax(variance12))]


def get_foreground(image):
    return image > threshold_otsu(image)


def foregrounds(conf1, sted):
    """``_acquire`` tail: fg_c from the first confocal, fg_s from the STED image (all ones if it is empty),
    restricted to fg_c."""
    fg_c = get_foreground(conf1)
    fg_s = get_foreground(sted) if np.any(sted) else np.ones_like(fg_c)
    return fg_s * fg_c, fg_c


# ------------------------------------------------------------------------------------------------
# a7: signal ratio and bleach
# ------------------------------------------------------------------------------------------------
def signal_ratio(sted, conf1, fg_s, fg_c, percentile=75):
    if not np.any(fg_s):
        return 0
    foreground = np.percentile(sted[fg_s], percentile)
    background = np.mean(sted[np.invert(fg_s)])
    ratio = (foreground - background) / np.percentile(conf1[fg_c], percentile)
    return None if ratio < 0 else ratio


def bleach(conf1, conf2, fg_c):
    signal_i = np.mean(conf1[fg_c])
    signal_e = np.mean(conf2[fg_c])
    return (signal_i - signal_e) / signal_i


# ------------------------------------------------------------------------------------------------
# a8: image-decorrelation resolution
# ------------------------------------------------------------------------------------------------
def _ring_sums(values, ring, n_rings):
    return np.bincount(ring.ravel(), weights=values.ravel(), minlength=n_rings + 1)[:n_rings]


def _decorr(Ik, Ikn_conj, Ikn_abs2, R, radii, sigma=None):
    if sigma:
        Ik = Ik * (1 - np.exp(-2 * sigma ** 2 * R ** 2))
    ring = np.searchsorted(radii, R, side="right")          # 0: R < r0 ; i: r[i-1] <= R < r[i] ; len: outside
    n = len(radii)
    nom = np.cumsum(_ring_sums(np.real(Ik * Ikn_conj), ring, n))
    denom1 = np.cumsum(_ring_sums(Ikn_abs2, ring, n))
    denom2 = np.sum(np.abs(Ik) ** 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        d = nom / np.sqrt(denom1 * denom2)
    d = np.floor(1000 * d) / 1000
    d[np.isnan(d)] = 0
    return d


def _decorr_peak(d, tol):
    d = np.asarray(d)
    idx = int(np.argmax(d))
    A = d[idx]
    while len(d) > 1:
        if (A - np.min(d[idx:])) >= tol or idx == 0:
            break
        d = d[:-1]
        idx = int(np.argmax(d))
        A = d[idx]
    return idx, A


def decorrelation(image, Nr=50, Ng=10, w=20, tol=0.0005):
    image = np.asarray(image, dtype=np.float64)
    edge = (np.sin(np.linspace(-np.pi / 2, np.pi / 2, w)) + 1) / 2
    wins = [np.concatenate([edge, np.ones(l - 2 * w), edge[::-1]]) for l in image.shape]
    W = wins[0][:, None] * wins[1][None, :]
    mean = image.mean()
    image = (W * (image - mean) + mean).astype(np.float32)
    h, wd = image.shape
    image = image[:h - 1 + h % 2, :wd - 1 + wd % 2]
    ys = np.linspace(-1, 1, image.shape[0])[:, None]
    xs = np.linspace(-1, 1, image.shape[1])[None, :]
    R = np.sqrt(xs ** 2 + ys ** 2)
    Ik = np.fft.fftshift(np.fft.fftn(np.fft.fftshift(image)))
    Ik = Ik * (R < 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        Ikn = Ik / np.abs(Ik)
    Ikn[~np.isfinite(Ikn)] = 0
    Ikn_conj, Ikn_abs2 = np.conj(Ikn), np.abs(Ikn) ** 2

    r = np.linspace(0, 1, Nr)
    idx0, A0 = _decorr_peak(_decorr(Ik, Ikn_conj, Ikn_abs2, R, r), tol)
    r0 = r[idx0]
    with np.errstate(divide="ignore"):
        g_max = 2 / r0
    if np.isinf(g_max):
        g_max = max(image.shape) / 2
    sigmas = np.array([image.shape[0] / 4] + list(np.exp(np.linspace(np.log(g_max), np.log(0.15), Ng))))
    As, rs = [A0], [r0]
    for sig in sigmas:
        idx, A = _decorr_peak(_decorr(Ik, Ikn_conj, Ikn_abs2, R, r, sig), tol)
        As.append(A)
        rs.append(r[idx])
    rs_a = np.array(rs)
    k = int(np.where(rs_a == rs_a.max())[0][-1])
    if k == 0:
        ind1 = 0
    elif k >= len(sigmas) - 1:
        ind1 = k - 2
    else:
        ind1 = k - 1
    ind2 = ind1 + 1
    r_fine = np.linspace(rs_a[k] - (r[1] - r[0]), min(rs_a[k] + 0.4, r[-1]), Nr)
    sig_fine = np.exp(np.linspace(np.log(sigmas[ind1]), np.log(sigmas[ind2]), Ng))
    As, rs = As[:-1], rs[:-1]
    for sig in sig_fine:
        idx, A = _decorr_peak(_decorr(Ik, Ikn_conj, Ikn_abs2, R, r_fine, sig), tol)
        As.append(A)
        rs.append(r_fine[idx])
    rs.append(r0)
    As.append(A0)
    rs, As = np.array(rs), np.array(As)
    rs[As < 0.05] = 0
    with np.errstate(divide="ignore"):
        return 2 / np.max(rs)


def resolution(sted, pixelsize=20e-9, res_cap=300):
    res = decorrelation(sted) * pixelsize / 1e-9
    return res_cap if res > res_cap else res


def mo_objectives(sted, conf1, conf2, fg_s, fg_c):
    """``MORewardCalculator.evaluate`` with obj_names [Resolution, Bleach, SNR] (gym_sted/rewards/rewards.py:55-70)."""
    return [resolution(sted), bleach(conf1, conf2, fg_c), signal_ratio(sted, conf1, fg_s, fg_c)]


# ------------------------------------------------------------------------------------------------
# a9: nanodomain detection and F1
# ------------------------------------------------------------------------------------------------
_EIGHT = np.ones((3, 3), dtype=bool)


def _ensure_spacing(coords, spacing):
    """skimage's ensure_spacing with p_norm=inf: walk the (already intensity-sorted) points, drop every later
    point closer than `spacing` (Chebyshev) to a kept one."""
    if len(coords) == 0:
        return coords
    tree = scipy.spatial.cKDTree(coords)
    near = tree.query_ball_point(coords, r=spacing, p=np.inf)
    rejected = set()
    for i, cand in enumerate(near):
        if i in rejected:
            continue
        for c in cand:
            if c != i and np.max(np.abs(coords[c] - coords[i])) < spacing:
                rejected.add(c)
    keep = [i for i in range(len(coords)) if i not in rejected]
    return coords[keep]


def peak_local_max(image, labels, min_distance=2):
    """skimage.feature.peak_local_max(image, min_distance, labels=labels, exclude_border=False)."""
    threshold = image.min()
    footprint = np.ones((2 * min_distance + 1,) * 2, dtype=bool)
    bg = np.finfo(image.dtype).min
    out = []
    for li, roi in enumerate(ndi.find_objects(labels)):
        if roi is None:
            continue
        lmask = labels[roi] == li + 1
        obj = image[roi].copy()
        obj[~lmask] = bg
        if obj.size == 1:
            pm = obj > threshold
        else:
            pm = obj == ndi.maximum_filter(obj, footprint=footprint, mode="nearest")
            if np.all(pm[lmask]):
                pm[:] = False
                pm[np.logical_xor(lmask, ndi.binary_opening(lmask))] = True
            pm &= obj > threshold
        cc = np.nonzero(pm)
        order = np.argsort(-obj[cc], kind="stable")
        coords = np.transpose(cc)[order]
        coords = _ensure_spacing(coords, min_distance)
        coords = coords + np.array([s.start for s in roi])
        out.append(coords)
    return np.vstack(out) if out else np.zeros((0, 2), dtype=np.int64)


def fit_gaussian(data):
    """``rewards/utils.fitgaussian``: Levenberg–Marquardt fit of amplitude*exp(-((cx-x)^2/(2sx^2)+(cy-y)^2/(2sy^2)))."""
    X, Y = np.indices(data.shape)
    X = X - np.max(X) / 2
    Y = Y - np.max(Y) / 2

    def err(p):
        amp, cx, cy, wx, wy = p
        wx, wy = float(wx) + 1e-3, float(wy) + 1e-3
        return np.ravel(amp * np.exp(-(((cx - X) ** 2.0 / (2 * wx ** 2.0)) + ((cy - Y) ** 2.0 / (2 * wy ** 2.0)))) - data)

    p, _ = scipy.optimize.leastsq(err, [data.max(), 0, 0, 1, 1], ftol=1e-3)
    return p


def validate_positions(guess, signal, pixelsize, half=3):
    padded = np.pad(signal, half, constant_values=0)
    valid = []
    for row, col in guess:
        data = padded[row:row + 2 * half + 1, col:col + 2 * half + 1]
        p = fit_gaussian(data)
        fx, fy = 2.355 * p[3] * pixelsize, 2.355 * p[4] * pixelsize
        if not (fx > 200e-9 or fx < 40e-9 or fy > 200e-9 or fy < 40e-9):
            valid.append((row, col))
    return np.asarray(valid)


def candidate_peaks(sted, threshold=2):
    """Everything of the detector before the Gaussian-fit validation (integer outputs).  skimage.filters.gaussian
    converts the int64 image with img_as_float first: x -> ((x + 0.5) * 2) / (imax - imin) = (2x + 1) * 2**-64; the
    power of two commutes with rounding, so 2x + 1 is filtered here."""
    sted = np.asarray(sted)
    filtered = ndi.gaussian_filter(2.0 * sted.astype(np.float64) + 1.0, 0.5, mode="nearest", truncate=4.0)
    mask = filtered > np.quantile(filtered, 0.99)
    lab, n = ndi.label(mask, structure=_EIGHT)
    if n:
        sizes = np.bincount(lab.ravel())
        small = sizes < 4
        small[0] = False
        mask = mask & ~small[lab]
    labels, _ = ndi.label(mask, structure=_EIGHT)
    return peak_local_max(filtered, labels, min_distance=threshold)


def detect_nanodomains(sted, pixelsize=20e-9, threshold=2):
    return validate_positions(candidate_peaks(sted, threshold), np.asarray(sted), pixelsize)


def match_counts(truth, guess, threshold=2):
    """(TP, FP, FN) of Hungarian matching on Euclidean distances; pairs at distance >= threshold are forbidden."""
    truth = np.asarray(truth, dtype=np.float64).reshape(-1, 2)
    guess = np.asarray(guess, dtype=np.float64).reshape(-1, 2)
    if len(truth) == 0 or len(guess) == 0:
        return 0, len(guess), len(truth)
    cost = scipy.spatial.distance.cdist(truth, guess)
    big = 1e6
    masked = np.where(cost < threshold, cost, big)
    ri, ci = scipy.optimize.linear_sum_assignment(masked)
    tp = int(np.sum(masked[ri, ci] < big))
    return tp, len(guess) - tp, len(truth) - tp


def f1_from_counts(tp, fp, fn):
    prec = tp / (tp + fp + 1e-6)
    rec = tp / (tp + fn + 1e-6)
    return 2 * prec * rec / (prec + rec + 1e-6)


def f1_score(truth, guess, threshold=2):
    return f1_from_counts(*match_counts(truth, guess, threshold))


def nanodomain_f1(sted, truth_coords, pixelsize=20e-9, threshold=2):
    """``NumberNanodomains.evaluate`` (objectives.py:404-430)."""
    return f1_score(truth_coords, detect_nanodomains(sted, pixelsize, threshold), threshold)
