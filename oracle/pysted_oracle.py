"""CPU oracle for the acquisition path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; nothing under ``gym_sted_b200/`` does.

It restates, in float64 numpy (+ ``raster_oracle.c`` for the dwell loop), what
``pysted.base`` does on the path gym-sted exercises:

======================================  =====================================================
reference call site                     restated here
======================================  =====================================================
``gym_sted/utils.py:157-164``           ``Microscope(...)``, ``Microscope.cache``
``gym_sted/utils.py:179-180``           ``Datamap``, ``Datamap.set_roi(i_ex, "max")``
``gym_sted/utils.py:616-631``           ``fluo.get_photons``, ``fluo.get_k_bleach``
``gym_sted/utils.py:646-657``           ``Microscope.get_effective``, ``detector.get_signal``
``gym_sted/envs/mosted_env.py:184-234`` ``Microscope.get_signal_and_bleach``
======================================  =====================================================

**PARITY UNPINNED.**  pySTED is an un-vendored, un-pinned dependency (``setup.py:13``) that is
not present in ``/root/reference`` nor installable here, and the reference's own tests pin no
number for this path.  The formulas follow the published model (Richards-Wolf
foci, Leutenegger 2010 depletion efficiency, Oracz 2017 ``k1*I^b`` bleaching) as recalled in
SURVEY.md Appendix A; the only in-repo anchors are the fitted parameter sets of
``gym_sted/routines/routines.json`` with the criteria of ``gym_sted/routines/create.py:6-63``
(expected signal 200 / 30 photons, expected bleach 0.2 / 0.7), reproduced to within about 25 %
by ``tests/test_oracle_anchors.py``.  The prefactor of the bleaching rate (``tau_sted``) was chosen
among dimensionally admissible candidates because it is the one that reproduces the four bleach
anchors.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import scipy.constants
import scipy.integrate
import scipy.signal
import scipy.special

HERE = os.path.dirname(os.path.abspath(__file__))
_H = scipy.constants.h
_C = scipy.constants.c


# --------------------------------------------------------------------------------------------
# beams — pixel-by-pixel scipy.integrate.quad like pySTED; the aperture
# integrals only depend on the radius, so they are memoised per distinct radius.
# --------------------------------------------------------------------------------------------
def _n_pixels(lambda_, na, px):
    return int(2.233 * lambda_ / (na * px) / 2) * 2 + 1


def _fwhm(values):
    hm = values.max() / 2
    i_max = int(np.argmax(values))
    left = int(np.abs(values[:i_max] - hm).argmin())
    right = int(np.abs(values[i_max:] - hm).argmin()) + i_max
    return right - left


def _fwhm_donut(values):
    hm = values.max() / 2
    c = len(values) // 2
    i_max = int(np.argmax(values[:c + 1]))
    ol = int(np.abs(values[:i_max] - hm).argmin())
    il = int(np.abs(values[i_max:c + 1] - hm).argmin()) + i_max
    return 2 * (c - ol), 2 * (c - il)


class GaussianBeam:
    def __init__(self, lambda_, polarization=np.pi / 2, beta=np.pi / 4):
        self.lambda_, self.polarization, self.beta = lambda_, polarization, beta

    def get_intensity(self, power, f, n, na, transmission, px):
        npx = _n_pixels(self.lambda_, na, px)
        ctr = npx // 2
        alpha = np.arcsin(na / n)
        k = 2 * np.pi * n / self.lambda_
        f1 = lambda t, kr: np.sqrt(np.cos(t)) * np.sin(t) * scipy.special.jv(0, kr * np.sin(t)) * (1 + np.cos(t))
        f2 = lambda t, kr: np.sqrt(np.cos(t)) * np.sin(t) ** 2 * scipy.special.jv(1, kr * np.sin(t))
        f3 = lambda t, kr: np.sqrt(np.cos(t)) * np.sin(t) * scipy.special.jv(2, kr * np.sin(t)) * (1 - np.cos(t))
        memo = {}
        out = np.zeros((npx, npx))
        ax = np.sin(self.polarization)
        ay = np.cos(self.polarization) * np.exp(1j * self.beta)
        for y in range(npx):
            h_rel = ctr - y
            for x in range(npx):
                w_rel = x - ctr
                ang, rad = np.arctan2(h_rel, w_rel), np.hypot(w_rel, h_rel)
                key = w_rel * w_rel + h_rel * h_rel
                if key not in memo:
                    kr = k * rad * px
                    memo[key] = tuple(scipy.integrate.quad(fn, 0, alpha, (kr,), epsabs=1e-13, epsrel=1e-12)[0]
                                      for fn in (f1, f2, f3))
                i1, i2, i3 = memo[key]
                ex = -1j * (ax * (i1 + i3 * np.cos(2 * ang)) + ay * i3 * np.sin(2 * ang))
                ey = -1j * (ax * i3 * np.sin(2 * ang) + ay * (i1 - i3 * np.cos(2 * ang)))
                ez = -2 * i2 * (ax * np.cos(ang) + ay * np.sin(ang))
                out[y, x] = abs(ex) ** 2 + abs(ey) ** 2 + abs(ez) ** 2
        out /= out.max()
        r = _fwhm(out[ctr])
        area = np.pi * (r * px) ** 2 / 2
        return out * transmission * power / area


class DonutBeam:
    def __init__(self, lambda_, polarization=np.pi / 4, beta=np.pi / 2, tau=400e-12, rate=40e6,
                 zero_residual=0.0, anti_stoke=True):
        self.lambda_, self.polarization, self.beta = lambda_, polarization, beta
        self.tau, self.rate, self.zero_residual, self.anti_stoke = tau, rate, zero_residual, anti_stoke

    def get_intensity(self, power, f, n, na, transmission, px):
        """Instantaneous (in-pulse) intensity per watt of average power."""
        npx = _n_pixels(self.lambda_, na, px)
        ctr = npx // 2
        alpha = np.arcsin(na / n)
        k = 2 * np.pi * n / self.lambda_
        sq = lambda t: np.sqrt(np.cos(t))
        fns = (
            lambda t, kr: sq(t) * np.sin(t) * (1 + np.cos(t)) * scipy.special.jv(1, kr * np.sin(t)),
            lambda t, kr: sq(t) * np.sin(t) * (1 - np.cos(t)) * scipy.special.jv(1, kr * np.sin(t)),
            lambda t, kr: sq(t) * np.sin(t) * (1 - np.cos(t)) * scipy.special.jv(3, kr * np.sin(t)),
            lambda t, kr: sq(t) * np.sin(t) ** 2 * scipy.special.jv(0, kr * np.sin(t)),
            lambda t, kr: sq(t) * np.sin(t) ** 2 * scipy.special.jv(2, kr * np.sin(t)),
        )
        memo = {}
        out = np.zeros((npx, npx))
        ax = np.sin(self.polarization)
        ay = np.cos(self.polarization) * np.exp(1j * self.beta)
        for y in range(npx):
            h_rel = ctr - y
            for x in range(npx):
                w_rel = x - ctr
                ang, rad = np.arctan2(h_rel, w_rel), np.hypot(w_rel, h_rel)
                key = w_rel * w_rel + h_rel * h_rel
                if key not in memo:
                    kr = k * rad * px
                    memo[key] = tuple(scipy.integrate.quad(fn, 0, alpha, (kr,), epsabs=1e-13, epsrel=1e-12)[0]
                                      for fn in fns)
                i1, i2, i3, i4, i5 = memo[key]
                e1, em, e2, e3 = np.exp(1j * ang), np.exp(-1j * ang), np.exp(2j * ang), np.exp(3j * ang)
                ex = ax * (i1 * e1 - i2 / 2 * em + i3 / 2 * e3) - ay * 0.5j * (i2 * em + i3 * e3)
                ey = -ax * 0.5j * (i2 * em + i3 * e3) + ay * (i1 * e1 + i2 / 2 * em - i3 / 2 * e3)
                ez = ax * 1j * (i4 - i5 * e2) - ay * (i4 + i5 * e2)
                out[y, x] = abs(ex) ** 2 + abs(ey) ** 2 + abs(ez) ** 2
        out /= out.max()
        r_out, r_in = _fwhm_donut(out[ctr])
        area = np.pi * (r_out * px) ** 2 / 2 - np.pi * (r_in * px) ** 2 / 2
        out = out * transmission * power / area
        if power > 0 and self.zero_residual > 0:
            old_max = out.max()
            out = out + self.zero_residual * old_max
            out = out / out.max() * old_max
        return out / (self.tau * self.rate)


class Objective:
    def __init__(self, f=2e-3, n=1.5, na=1.4, transmission=None):
        self.f, self.n, self.na = f, n, na
        self.transmission = transmission or {}

    def get_transmission(self, lambda_):
        return self.transmission[int(round(lambda_ * 1e9))]


class Detector:
    def __init__(self, n_airy=0.7, noise=False, background=0.0, pcef=0.1, pdef=0.5):
        self.n_airy, self.noise, self.background, self.pcef, self.pdef = n_airy, noise, background, pcef, pdef

    def get_detection_psf(self, lambda_, psf, na, transmission, px):
        radius = self.n_airy * 0.61 * lambda_ / na
        n = psf.shape[0]
        c = n // 2
        yy, xx = np.mgrid[:n, :n]
        pinhole = (((xx - c) ** 2 + (yy - c) ** 2) <= (radius / px) ** 2).astype(float)
        det = scipy.signal.convolve2d(psf, pinhole, "same")
        return det / det.max() * transmission

    def get_signal(self, photons, dwelltime, rate=None, rng=None):
        """mean = photons*pcef*pdef*dwelltime (+ background*dwelltime); Poisson when ``noise``."""
        mean = np.asarray(photons, dtype=np.float64) * self.pcef * self.pdef * dwelltime
        if not self.noise:
            return mean + self.background * dwelltime
        rng = rng or np.random.default_rng()
        signal = rng.poisson(mean)
        if self.background > 0:
            signal = signal + rng.poisson(self.background * dwelltime, np.shape(mean))
        return signal


class Fluorescence:
    def __init__(self, lambda_, qy, sigma_abs, sigma_ste, tau, tau_vib, tau_tri, k0, k1, b,
                 triplet_dynamics_frac=0, **_):
        self.lambda_, self.qy, self.sigma_abs, self.sigma_ste = lambda_, qy, dict(sigma_abs), dict(sigma_ste)
        self.tau, self.tau_vib, self.tau_tri = tau, tau_vib, tau_tri
        self.k0, self.k1, self.b, self.triplet_dynamics_frac = k0, k1, b, triplet_dynamics_frac

    def get_sigma_abs(self, lambda_):
        return self.sigma_abs[int(round(lambda_ * 1e9))]

    def get_sigma_ste(self, lambda_):
        return self.sigma_ste[int(round(lambda_ * 1e9))]

    def get_photons(self, intensity, lambda_=None):
        lambda_ = self.lambda_ if lambda_ is None else lambda_
        return np.asarray(intensity, dtype=np.float64) // (_H * _C / lambda_)

    def get_psf(self, na, px):
        n = _n_pixels(self.lambda_, na, px)
        c = n // 2
        sigma = self.lambda_ / (2 * na) / (2 * np.sqrt(2 * np.log(2)))
        out = np.zeros((n, n))
        cdf = lambda z: 0.5 * (1 + scipy.special.erf(z / (sigma * np.sqrt(2))))
        for y in range(n):
            gy = cdf((y - c + 0.5) * px) - cdf((y - c - 0.5) * px)
            for x in range(n):
                out[y, x] = gy * (cdf((x - c + 0.5) * px) - cdf((x - c - 0.5) * px))
        return out

    def get_k_bleach(self, lambda_ex, lambda_sted, phi_ex, phi_sted, tau_sted, tau_rep, dwelltime):
        """k_b = sigma_abs*phi_ex * tau_sted * (k0*I + k1*I**b), I = in-pulse STED intensity (W/m^2)
        recovered from the time-averaged photon flux ``phi_sted`` (gym_sted/utils.py:627-630)."""
        if self.triplet_dynamics_frac:
            raise NotImplementedError("triplet dynamics are not part of the restated model")
        i_inst = np.asarray(phi_sted, dtype=np.float64) * (_H * _C / lambda_sted) * tau_rep / tau_sted
        k_ex = self.get_sigma_abs(lambda_ex) * np.asarray(phi_ex, dtype=np.float64)
        return k_ex * tau_sted * (self.k0 * i_inst + self.k1 * i_inst ** self.b)


class Datamap:
    def __init__(self, whole_datamap, datamap_pixelsize):
        self.whole_datamap = np.array(whole_datamap)
        self.pixelsize = datamap_pixelsize
        self.roi = None
        self.sub_datamaps_dict = {}

    def set_roi(self, laser, intervals="max"):
        assert intervals == "max"
        p = laser.shape[0] // 2
        h, w = self.whole_datamap.shape
        self.whole_datamap = np.pad(self.whole_datamap, p)
        self.roi = (slice(p, p + h), slice(p, p + w))
        self.sub_datamaps_dict["base"] = self.whole_datamap

    def __getitem__(self, key):
        return self.sub_datamaps_dict[key]


def _resize(*images):
    side = max(im.shape[0] for im in images)
    return tuple(np.pad(im, (side - im.shape[0]) // 2) for im in images)


_lib = None


def raster_lib():
    """Compile (once) and load the C dwell loop."""
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "_build", "liboracle.so")
        src = os.path.join(HERE, "raster_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            os.makedirs(os.path.dirname(so), exist_ok=True)
            subprocess.check_call(["gcc", "-O3", "-march=native", "-fPIC", "-shared", "-o", so, src, "-lm"])
        _lib = ctypes.CDLL(so)
        dp = ctypes.POINTER(ctypes.c_double)
        _lib.oracle_raster.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp, ctypes.c_int,
                                       ctypes.c_uint64, dp]
        _lib.oracle_raster.restype = None
        _lib.oracle_raster_batch.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp,
                                             ctypes.c_int, ctypes.c_uint64, dp]
        _lib.oracle_raster_batch.restype = None
        _lib.oracle_raster_tables.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp,
                                              ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.c_uint64, dp]
        _lib.oracle_raster_tables.restype = None
    return _lib


def raster(padded, H, W, E, S, mode, seed=0):
    """Row-major dwell loop; ``padded`` (Hp,Wp) float64 is updated in place when mode > 0."""
    K = E.shape[0]
    assert padded.dtype == np.float64 and padded.flags.c_contiguous
    assert padded.shape == (H + K - 1, W + K - 1)
    E = np.ascontiguousarray(E, dtype=np.float64)
    S = np.ascontiguousarray(S, dtype=np.float64)
    out = np.empty((H, W), dtype=np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    raster_lib().oracle_raster(padded.ctypes.data_as(dp), H, W, K, E.ctypes.data_as(dp), S.ctypes.data_as(dp),
                               int(mode), int(seed), out.ctypes.data_as(dp))
    return out


def raster_tables(padded, H, W, E, S, idx, mode, seed=0):
    """Dwell loop with one (E, S) table per dwell: ``E``, ``S`` (n,K,K), ``idx`` (H,W) table index of every dwell."""
    K = E.shape[-1]
    assert padded.dtype == np.float64 and padded.flags.c_contiguous and padded.shape == (H + K - 1, W + K - 1)
    E = np.ascontiguousarray(E, dtype=np.float64)
    S = np.ascontiguousarray(S, dtype=np.float64)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    assert idx.shape == (H, W) and idx.min() >= 0 and idx.max() < E.shape[0] == S.shape[0]
    out = np.empty((H, W), dtype=np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    raster_lib().oracle_raster_tables(padded.ctypes.data_as(dp), H, W, K, E.ctypes.data_as(dp), S.ctypes.data_as(dp),
                                      idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), int(mode), int(seed),
                                      out.ctypes.data_as(dp))
    return out


def raster_numpy(padded, H, W, E, S, mode, rng=None):
    """Pure-numpy twin of the C loop (small cases only); mode 2 uses numpy's binomial."""
    K = E.shape[0]
    out = np.empty((H, W))
    for r in range(H):
        for c in range(W):
            win = padded[r:r + K, c:c + K]
            out[r, c] = (E * win).sum()
            if mode == 1:
                win *= S
            elif mode == 2:
                win[...] = rng.binomial(win.astype(np.int64), S)
    return out


class Microscope:
    def __init__(self, excitation, sted, detector, objective, fluo, load_cache=False):
        self.excitation, self.sted, self.detector, self.objective, self.fluo = excitation, sted, detector, objective, fluo
        self._cache = {}

    def cache(self, px, save_cache=False):
        key = int(round(px * 1e9))
        if key not in self._cache:
            o = self.objective
            i_ex = self.excitation.get_intensity(1, o.f, o.n, o.na, o.get_transmission(self.excitation.lambda_), px)
            i_sted = self.sted.get_intensity(1, o.f, o.n, o.na, o.get_transmission(self.sted.lambda_), px)
            psf = self.fluo.get_psf(o.na, px)
            psf_det = self.detector.get_detection_psf(self.fluo.lambda_, psf, o.na,
                                                      o.get_transmission(self.fluo.lambda_), px)
            self._cache[key] = _resize(i_ex, i_sted, psf_det)
        return self._cache[key]

    def get_effective(self, px, p_ex, p_sted):
        i_ex_u, i_sted_u, psf_det = self.cache(px)
        i_ex = i_ex_u * p_ex          # time averaged
        i_sted = i_sted_u * p_sted    # instantaneous
        fl, st = self.fluo, self.sted
        exc = fl.get_sigma_abs(self.excitation.lambda_) * i_ex * fl.qy
        if st.anti_stoke:
            exc = exc + fl.get_sigma_abs(st.lambda_) * i_sted * fl.qy
        i_sat = (_H * _C / st.lambda_) / (fl.get_sigma_ste(st.lambda_) * fl.tau)
        zeta = i_sted / i_sat
        k_vib, k_s1 = 1 / fl.tau_vib, 1 / fl.tau
        gamma = zeta * k_vib / (zeta * k_s1 + k_vib)
        T = 1 / st.rate
        eta = (((1 + gamma * np.exp(-k_s1 * st.tau * (1 + gamma))) / (1 + gamma))
               - np.exp(-k_s1 * (gamma * st.tau + T))) / (1 - np.exp(-k_s1 * T))
        return exc * eta * psf_det

    def get_k_bleach_map(self, px, p_ex, p_sted, pdt):
        """The map gym-sted's own restatement builds at utils.py:616-631."""
        i_ex_u, i_sted_u, _ = self.cache(px)
        st = self.sted
        phi_ex = self.fluo.get_photons(i_ex_u * p_ex, self.excitation.lambda_)
        phi_sted = self.fluo.get_photons(i_sted_u * p_sted * st.tau * st.rate, st.lambda_)
        return self.fluo.get_k_bleach(self.excitation.lambda_, st.lambda_, phi_ex, phi_sted, st.tau, 1 / st.rate, pdt)

    def get_signal_and_bleach(self, datamap, pixelsize, pdt, p_ex, p_sted, bleach=True, update=True, seed=None,
                              sample_bleach=True):
        if datamap.roi is None:
            raise ValueError("datamap ROI is not set")
        K = self.cache(pixelsize)[0].shape[0]
        rs, cs = datamap.roi
        H, W = rs.stop - rs.start, cs.stop - cs.start
        work = np.ascontiguousarray(datamap.whole_datamap, dtype=np.float64).copy()
        assert work.shape == (H + K - 1, W + K - 1)
        mode = 0 if not bleach else (2 if sample_bleach else 1)
        if np.ndim(pdt) or np.ndim(p_ex) or np.ndim(p_sted):
            # per-dwell parameters (timed_sted_env.py:141-148): one table per distinct (pdt, p_ex, p_sted) triple,
            # each from the physics above
            trip = np.stack([np.broadcast_to(np.asarray(x, dtype=np.float64), (H, W)).ravel() for x in (pdt, p_ex, p_sted)], 1)
            keys, idx = np.unique(trip, axis=0, return_inverse=True)
            E = np.stack([self.get_effective(pixelsize, k[1], k[2]) for k in keys])
            S = np.stack([np.exp(-self.get_k_bleach_map(pixelsize, k[1], k[2], k[0]) * k[0]) for k in keys])
            intensity = raster_tables(work, H, W, E, S, idx.reshape(H, W), mode, 0 if seed is None else seed)
            pdt = np.broadcast_to(np.asarray(pdt, dtype=np.float64), (H, W))
        else:
            E = self.get_effective(pixelsize, p_ex, p_sted)
            S = np.exp(-self.get_k_bleach_map(pixelsize, p_ex, p_sted, pdt) * pdt)
            intensity = raster(work, H, W, E, S, mode, 0 if seed is None else seed)
        photons = self.fluo.get_photons(intensity)
        rng = np.random.default_rng(seed)
        signal = self.detector.get_signal(photons, pdt, self.sted.rate, rng=rng)
        bleached = {"base": work if sample_bleach is False or not bleach else work.astype(np.int64)}
        if bleach and update:
            datamap.whole_datamap = bleached["base"]
            datamap.sub_datamaps_dict["base"] = bleached["base"]
        return signal, bleached, intensity


def default_microscope(fluo_params, detector_params, laser_ex, laser_sted, objective_params):
    """``MicroscopeGenerator.generate_microscope`` (gym_sted/utils.py:151-166) on the oracle classes."""
    return Microscope(GaussianBeam(**laser_ex), DonutBeam(**laser_sted), Detector(**detector_params),
                      Objective(**objective_params), Fluorescence(**fluo_params))
