"""``pysted.base``-compatible facade over the CUDA hot path (Python level).

Same names, argument meaning and mutation semantics as the pySTED objects gym-sted builds in
``gym_sted/utils.py:151-182`` and calls at ``gym_sted/envs/mosted_env.py:184,222,227,232``:

* ``Microscope.cache(px)`` -> ``(i_ex, i_sted, psf_det)`` (host set-up, ``physics.py``);
* ``Microscope.get_effective(px, p_ex, p_sted)`` -> ``(K,K)`` array, formed on the GPU;
* ``Microscope.get_signal_and_bleach(datamap, pixelsize, pdt, p_ex, p_sted, bleach=True,
  update=True, seed=None)`` -> ``(photons, {"base": bleached}, intensity)``; with ``bleach`` and
  ``update`` the datamap is mutated in place like pySTED does;
* ``Datamap(whole_datamap, datamap_pixelsize)``, ``.set_roi(i_ex, "max")``, ``.roi``,
  ``.whole_datamap``, ``["base"]``.

Everything numeric goes through ``libsted_b200.so``; there is no numpy fallback for the scan.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib, physics

_PSF_CACHE: dict = {}


class GaussianBeam:
    def __init__(self, lambda_, **kwargs):
        self.lambda_ = lambda_
        self.polarization = kwargs.get("polarization", np.pi / 2)
        self.beta = kwargs.get("beta", np.pi / 4)

    def get_intensity(self, power, f, n, na, transmission, datamap_pixelsize):
        return physics.gaussian_beam_intensity(self.lambda_, power, n, na, transmission, datamap_pixelsize,
                                               self.polarization, self.beta)


class DonutBeam:
    def __init__(self, lambda_, **kwargs):
        self.lambda_ = lambda_
        self.polarization = kwargs.get("polarization", np.pi / 4)
        self.beta = kwargs.get("beta", np.pi / 2)
        self.tau = kwargs.get("tau", 400e-12)
        self.rate = kwargs.get("rate", 40e6)
        self.zero_residual = kwargs.get("zero_residual", 0)
        self.anti_stoke = kwargs.get("anti_stoke", True)

    def get_intensity(self, power, f, n, na, transmission, datamap_pixelsize):
        return physics.donut_beam_intensity(self.lambda_, power, n, na, transmission, datamap_pixelsize, self.tau,
                                            self.rate, self.zero_residual, self.polarization, self.beta)


class Objective:
    def __init__(self, f=2e-3, n=1.5, na=1.4, transmission=None):
        self.f, self.n, self.na = f, n, na
        self.transmission = dict(transmission or {})

    def get_transmission(self, lambda_):
        return self.transmission[int(round(lambda_ * 1e9))]


class Detector:
    def __init__(self, **kwargs):
        self.n_airy = kwargs.get("n_airy", 0.7)
        self.noise = kwargs.get("noise", False)
        self.background = kwargs.get("background", 0.0)
        self.pcef = kwargs.get("pcef", 0.1)
        self.pdef = kwargs.get("pdef", 0.5)

    def get_detection_psf(self, lambda_, na, transmission, datamap_pixelsize):
        return physics.detection_psf(lambda_, na, transmission, datamap_pixelsize, self.n_airy)

    def get_signal(self, photons, dwelltime, rate=None, seed=None):
        """Host twin of the fused detector epilogue, kept for ``FluorescenceOptimizer``
        (``gym_sted/utils.py:654-657``), which calls it on a scalar."""
        mean = np.asarray(photons, dtype=np.float64) * self.pcef * self.pdef * dwelltime
        if not self.noise:
            return mean + self.background * dwelltime
        rng = np.random.default_rng(seed)
        signal = rng.poisson(mean)
        if self.background > 0:
            signal = signal + rng.poisson(self.background * dwelltime, np.shape(mean))
        return signal

    def params(self):
        return _lib.StedDetectorParams(self.pcef, self.pdef, self.background)


class Fluorescence:
    def __init__(self, lambda_, **kwargs):
        self.lambda_ = lambda_
        self.qy = kwargs.get("qy", 0.65)
        self.sigma_abs = dict(kwargs.get("sigma_abs", {}))
        self.sigma_ste = dict(kwargs.get("sigma_ste", {}))
        self.tau = kwargs.get("tau", 3.5e-9)
        self.tau_vib = kwargs.get("tau_vib", 1e-12)
        self.tau_tri = kwargs.get("tau_tri", 1.2e-6)
        self.k0 = kwargs.get("k0", 0)
        self.k1 = kwargs.get("k1", 2.9e-16)
        self.b = kwargs.get("b", 1.66)
        self.triplet_dynamics_frac = kwargs.get("triplet_dynamics_frac", 0)

    def get_sigma_abs(self, lambda_):
        return self.sigma_abs[int(round(lambda_ * 1e9))]

    def get_sigma_ste(self, lambda_):
        return self.sigma_ste[int(round(lambda_ * 1e9))]

    def get_photons(self, intensity, lambda_=None):
        """Photon flux of an intensity (W/m^2) -> photons/(m^2 s), floor division like pySTED
        (called at ``gym_sted/utils.py:624-625,653``)."""
        lambda_ = self.lambda_ if lambda_ is None else lambda_
        return np.asarray(intensity, dtype=np.float64) // (physics.H_PLANCK * physics.C_LIGHT / lambda_)

    def get_k_bleach(self, lambda_ex, lambda_sted, phi_ex, phi_sted, tau_sted, tau_rep, dwelltime):
        """Bleaching-rate map (1/s) from photon-flux maps: the 7-argument accessor the reference calls at
        ``gym_sted/utils.py:627-630`` (``phi_sted`` is the TIME-AVERAGED STED flux, ``phi*tau_sted/tau_rep``).
        Host float64 twin of the formula ``make_effective_kernel`` evaluates per environment on the device
        (``csrc/sted_kernels.cu``); the scan itself never calls this."""
        if self.triplet_dynamics_frac:
            raise NotImplementedError("triplet_dynamics_frac != 0 is outside the modelled path")
        e_sted = physics.H_PLANCK * physics.C_LIGHT / lambda_sted
        i_inst = np.asarray(phi_sted, dtype=np.float64) * e_sted * tau_rep / tau_sted
        k_ex = self.get_sigma_abs(lambda_ex) * np.asarray(phi_ex, dtype=np.float64)
        return k_ex * tau_sted * (self.k0 * i_inst + self.k1 * i_inst ** self.b)


def fluo_struct(fluo: Fluorescence, excitation: GaussianBeam, sted: DonutBeam) -> _lib.StedFluoParams:
    if fluo.triplet_dynamics_frac:
        raise NotImplementedError("triplet_dynamics_frac != 0 is outside the modelled path "
                                  "(every gym-sted parameter set uses 0: defaults.py:41, routines.json)")
    num = lambda v: float(np.asarray(v, dtype=np.float64).reshape(-1)[0])     # optimisers hand in 1-element arrays
    return _lib.StedFluoParams(
        excitation.lambda_, sted.lambda_, fluo.lambda_, num(fluo.get_sigma_abs(excitation.lambda_)),
        num(fluo.sigma_abs.get(int(round(sted.lambda_ * 1e9)), 0.0)), num(fluo.get_sigma_ste(sted.lambda_)), fluo.qy, fluo.tau,
        fluo.tau_vib, fluo.tau_tri, num(fluo.k0), num(fluo.k1), num(fluo.b), fluo.triplet_dynamics_frac, sted.tau, sted.rate,
        int(bool(sted.anti_stoke)), 0)


class Datamap:
    """``pysted.base.Datamap``: zero-pads the molecule map by K//2 and remembers the ROI slices."""

    def __init__(self, whole_datamap, datamap_pixelsize):
        self.whole_datamap = np.array(whole_datamap)
        self.whole_shape = self.whole_datamap.shape
        self.pixelsize = datamap_pixelsize
        self.roi = None
        self.roi_corners = None
        self.sub_datamaps_dict = {}

    def set_roi(self, laser, intervals="max"):
        if intervals != "max":
            raise NotImplementedError("only the 'max' ROI used by gym-sted (utils.py:180) is supported")
        p = laser.shape[0] // 2
        h, w = self.whole_datamap.shape
        self.whole_datamap = np.pad(self.whole_datamap, p)
        self.roi = (slice(p, p + h), slice(p, p + w))
        self.roi_corners = {"tl": (p, p), "tr": (p, p + w - 1), "bl": (p + h - 1, p), "br": (p + h - 1, p + w - 1)}
        self.sub_datamaps_dict = {"base": self.whole_datamap}

    def __getitem__(self, key):
        return self.sub_datamaps_dict[key]


class Microscope:
    def __init__(self, excitation, sted, detector, objective, fluo, load_cache=False, verbose=False):
        self.excitation, self.sted, self.detector, self.objective, self.fluo = excitation, sted, detector, objective, fluo
        self._acq_counter = 0

    # ---- a1: host set-up ------------------------------------------------------------------
    def _cache_key(self, px):
        o, e, s, d = self.objective, self.excitation, self.sted, self.detector
        return (int(round(px * 1e11)), e.lambda_, e.polarization, e.beta, s.lambda_, s.polarization, s.beta, s.tau,
                s.rate, s.zero_residual, self.fluo.lambda_, d.n_airy, o.n, o.na,
                o.get_transmission(e.lambda_), o.get_transmission(s.lambda_), o.get_transmission(self.fluo.lambda_))

    def is_cached(self, datamap_pixelsize):
        return self._cache_key(datamap_pixelsize) in _PSF_CACHE

    def cache(self, datamap_pixelsize, save_cache=False):
        key = self._cache_key(datamap_pixelsize)
        if key not in _PSF_CACHE:
            o = self.objective
            i_ex = self.excitation.get_intensity(1, o.f, o.n, o.na, o.get_transmission(self.excitation.lambda_),
                                                 datamap_pixelsize)
            i_sted = self.sted.get_intensity(1, o.f, o.n, o.na, o.get_transmission(self.sted.lambda_),
                                             datamap_pixelsize)
            psf_det = self.detector.get_detection_psf(self.fluo.lambda_, o.na, o.get_transmission(self.fluo.lambda_),
                                                      datamap_pixelsize)
            _PSF_CACHE[key] = physics.resize_common(i_ex, i_sted, psf_det)
        return _PSF_CACHE[key]

    def clear_cache(self):
        _PSF_CACHE.clear()

    def _packed_cache(self, px):
        return np.ascontiguousarray(np.stack(self.cache(px)), dtype=np.float64)

    # ---- a2 / a3 on the device -------------------------------------------------------------
    def _tables(self, px, p_ex, p_sted, pdt):
        import torch  # device memory only
        lib = _lib.load()
        cache = self._packed_cache(px)
        K = cache.shape[-1]
        dev = torch.device("cuda")
        d_cache = torch.from_numpy(cache).to(dev)
        fl = fluo_struct(self.fluo, self.excitation, self.sted)
        d_fluo = torch.frombuffer(bytearray(bytes(fl)), dtype=torch.uint8).to(dev)
        scal = torch.tensor([[p_ex], [p_sted], [pdt]], dtype=torch.float64, device=dev)
        tables = torch.empty((1, _lib.STED_TABLES, K, _lib.kpitch(K)), dtype=torch.float32, device=dev)
        kb = torch.empty((1, K, K), dtype=torch.float64, device=dev)
        _lib.check(lib.sted_make_effective(d_cache.data_ptr(), d_fluo.data_ptr(), scal[0].data_ptr(),
                                           scal[1].data_ptr(), scal[2].data_ptr(), 1, K, tables.data_ptr(),
                                           kb.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return tables[0].cpu().numpy()[:, :, :K].astype(np.float64), kb[0].cpu().numpy()

    def get_effective(self, datamap_pixelsize, p_ex, p_sted):
        tables, _ = self._tables(datamap_pixelsize, p_ex, p_sted, 0.0)
        return tables[_lib.STED_TAB_E]

    def get_k_bleach(self, datamap_pixelsize, p_ex, p_sted, pdt):
        """Bleaching-rate map (1/s): what ``fluo.get_k_bleach`` returns at ``gym_sted/utils.py:627-630``."""
        return self._tables(datamap_pixelsize, p_ex, p_sted, pdt)[1]

    # ---- a5 --------------------------------------------------------------------------------
    def get_signal_and_bleach(self, datamap, pixelsize, pdt, p_ex, p_sted, indices=None, acquired_intensity=None,
                              pixel_list=None, bleach=True, update=True, seed=None, filter_bypass=False,
                              bleach_func=None, sample_func=None, steps=None, sample_bleach=True):
        """One raster-scan acquisition of ``datamap`` (``gym_sted/envs/mosted_env.py:184,222,227,232``).

        ``pdt``, ``p_ex`` and ``p_sted`` may be scalars or (H, W) arrays with one value per dwell, as the timed
        environments pass them (``gym_sted/envs/timed_sted_env.py:141-148``).  Constant arrays take the scalar path;
        anything else runs the general dwell-by-dwell scan (``sted_acquire_dwellmaps``), which allows up to 256
        distinct ``p_sted`` values per image.  The whole padded map takes part — molecules a caller placed in the
        K//2 frame are imaged and bleached like pySTED does — and with ``bleach`` and ``update`` the datamap is
        mutated in place."""
        import torch  # device memory only

        if datamap.roi is None:
            raise ValueError("Datamap ROI is not set; call datamap.set_roi(i_ex, 'max') first")
        if pixel_list is not None or bleach_func is not None or sample_func is not None:
            raise NotImplementedError("custom pixel lists / bleach functions are outside the rebuilt path")
        lib = _lib.load()
        if _lib.device_count() == 0:
            raise _lib.StedError(-2, "no CUDA device: gym_sted_b200 has no CPU fallback")
        cache = self._packed_cache(pixelsize)
        K = cache.shape[-1]
        rs, cs = datamap.roi
        H, W = rs.stop - rs.start, cs.stop - cs.start
        whole = np.asarray(datamap.whole_datamap)
        Hp, Wp = H + K - 1, W + K - 1
        if whole.shape != (Hp, Wp):
            raise ValueError(f"datamap shape {whole.shape} does not match ROI {H}x{W} padded for K={K}")
        maps = [np.asarray(x, dtype=np.float64) for x in (pdt, p_ex, p_sted)]
        per_dwell = any(m.ndim and np.ptp(m) != 0 for m in maps)
        n_tables = 1
        if per_dwell:
            pdt_map, pex_map, psted_map = (np.ascontiguousarray(np.broadcast_to(m, (H, W))) for m in maps)
            if pdt_map.min() < 0 or pex_map.min() < 0 or psted_map.min() < 0:
                raise ValueError("pdt, p_ex and p_sted must be non-negative")
            pdt_ref = float(pdt_map.max()) or 1e-6
            # E and k_b are linear in p_ex unless the depletion beam also excites (anti_stoke): one table per distinct
            # p_sted, or per distinct (p_ex, p_sted) pair in that case
            if self.sted.anti_stoke and np.ptp(pex_map) != 0:
                keys, table_id = np.unique(np.stack([pex_map.ravel(), psted_map.ravel()], 1), axis=0, return_inverse=True)
                tab_pex, tab_psted, wex_map = keys[:, 0].copy(), keys[:, 1].copy(), np.ones((H, W))
            else:
                tab_psted, table_id = np.unique(psted_map.ravel(), return_inverse=True)
                pex_ref = float(pex_map.max()) or 1e-6
                tab_pex, wex_map = np.full(tab_psted.shape, pex_ref), pex_map / pex_ref
            n_tables = int(tab_psted.size)
            if n_tables > 256:
                raise NotImplementedError(f"{n_tables} distinct depletion settings in one image; the general scan "
                                          "holds at most 256")
        else:
            pdt_ref, pex_ref, p_sted = (float(m.reshape(-1)[0]) for m in maps)
            tab_pex, tab_psted = np.array([pex_ref]), np.array([p_sted])
        flags = 0
        if bleach:
            flags |= _lib.STED_BLEACH
            if sample_bleach:
                flags |= _lib.STED_SAMPLE_BLEACH
        if self.detector.noise:
            flags |= _lib.STED_SAMPLE_PHOTONS
        if float(np.sum(whole)) != float(np.sum(whole[rs, cs])):
            flags |= _lib.STED_FRAME_MOLECULES          # the caller placed molecules in the K//2 frame
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1))
        self._acq_counter += 1
        dev = torch.device("cuda")
        st = torch.cuda.current_stream(dev).cuda_stream
        Wpitch = (Wp + 3) & ~3
        d_in = torch.zeros((1, Hp, Wpitch), dtype=torch.float32, device=dev)
        d_in[0, :, :Wp] = torch.from_numpy(np.ascontiguousarray(whole, dtype=np.float32)).to(dev)
        d_out = torch.zeros_like(d_in) if bleach else None
        fl = fluo_struct(self.fluo, self.excitation, self.sted)
        det = self.detector.params()
        d_cache = torch.from_numpy(cache).to(dev)
        d_fluo = torch.frombuffer(bytearray(bytes(fl)), dtype=torch.uint8).to(dev)
        d_fluo = d_fluo.repeat(n_tables)
        scal = torch.from_numpy(np.stack([tab_pex, tab_psted, np.full(n_tables, pdt_ref)])).to(dev)
        tables = torch.empty((n_tables, _lib.STED_TABLES, K, _lib.kpitch(K)), dtype=torch.float32, device=dev)
        kb = torch.empty((n_tables, K, K), dtype=torch.float64, device=dev) if per_dwell else None
        _lib.check(lib.sted_make_effective(d_cache.data_ptr(), d_fluo.data_ptr(), scal[0].data_ptr(), scal[1].data_ptr(),
                                           scal[2].data_ptr(), n_tables, K, tables.data_ptr(),
                                           kb.data_ptr() if per_dwell else None, st))
        intensity = torch.empty((1, H, W), dtype=torch.float32, device=dev)
        signal = torch.empty_like(intensity)
        if per_dwell:
            d_wdt = torch.from_numpy(pdt_map / pdt_ref).to(dev, torch.float32).reshape(1, H, W).contiguous()
            d_wex = torch.from_numpy(wex_map).to(dev, torch.float32).reshape(1, H, W).contiguous()
            d_tid = torch.from_numpy(table_id.astype(np.uint8)).to(dev).reshape(1, H, W).contiguous()
            _lib.check(lib.sted_acquire_dwellmaps(
                d_in.data_ptr(), d_out.data_ptr() if bleach else None, tables.data_ptr(), kb.data_ptr(), scal[2].data_ptr(),
                d_wdt.data_ptr(), d_wex.data_ptr(), d_tid.data_ptr(), n_tables, 1, H, W, K, flags, ctypes.byref(det),
                float(self.fluo.lambda_), int(seed), self._acq_counter, 0, intensity.data_ptr(), signal.data_ptr(), st))
        else:
            ws, ws_bytes = None, 0
            if bleach and sample_bleach:
                ws_bytes = int(lib.sted_bleach_workspace_bytes(1, H, int(max(whole.sum(), 0)) + 16))
                ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
            _lib.check(lib.sted_acquire(
                d_in.data_ptr(), d_out.data_ptr() if bleach else None, tables.data_ptr(), scal[2].data_ptr(), 1, H, W, K,
                flags, ctypes.byref(det), float(self.fluo.lambda_), int(seed), self._acq_counter, 0,
                intensity.data_ptr(), signal.data_ptr(), ws.data_ptr() if ws is not None else None, ws_bytes, st))
        if bleach:
            out = d_out[0, :, :Wp].cpu().numpy()
            out = np.rint(out).astype(np.int64) if sample_bleach else out.astype(np.float64)
        else:
            out = whole.copy()
        bleached = {"base": out}
        if bleach and update:
            datamap.whole_datamap = out
            datamap.sub_datamaps_dict = {"base": out}
        sig = signal[0].cpu().numpy()
        photons = np.rint(sig).astype(np.int64) if self.detector.noise else sig.astype(np.float64)
        return photons, bleached, intensity[0].cpu().numpy().astype(np.float64)
