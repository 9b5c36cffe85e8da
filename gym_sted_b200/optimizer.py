"""``FluorescenceOptimizer`` on the device (``gym_sted/utils.py:442-686``).

The reference fits ``(sigma_abs, k1, b)`` to sampled signal / bleach criteria at every reset of its "hard" environments
with 25 rounds of two scipy L-BFGS-B runs (about one second per environment).  ``sted_fluo_optimize`` follows the same
optimisation path for a whole batch of criteria in one kernel launch (``csrc/optimizer_kernels.cu``,
``csrc/lbfgsb_small.cuh``); this module is the host-side mirror of the reference class: same constructor arguments,
``FACTORS``, ``optimize(criterions)`` with the reference's nested result, plus ``optimize_batch``.
The detector's Poisson noise inside ``expected_confocal_signal`` is replaced by its expectation.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib, defaults
from .base import Microscope, fluo_struct

_VP = ctypes.c_void_p


def gaussian_blob(K: int, sigma: float, scale: float = 40.0, truncate: float = 4.0):
    """The test structure of ``expected_confocal_signal`` (``utils.py:646-650``): ``skimage.filters.gaussian`` of a
    unit pixel (mode "nearest", truncate 4 — away from the border that is the outer product of scipy's 1-D kernel with
    itself, exactly), scaled to a peak of ``scale`` molecules.  Returns (blob (K,K) float64, radius)."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    w = np.exp(-0.5 / (float(sigma) * float(sigma)) * x ** 2)
    w = w / w.sum()
    c = K // 2
    if radius > c:
        raise ValueError("the blob does not fit the PSF window")
    blob = np.zeros((K, K))
    blob[c - radius:c + radius + 1, c - radius:c + radius + 1] = w[:, None] * w[None, :]
    return blob / blob.max() * scale, radius


def _criteria_struct(crit) -> _lib.StedFluoCriteria:
    get = crit.get if hasattr(crit, "get") else (lambda k, d=None: crit[k] if k in crit else d)
    s, b = get("signal", None), get("bleach", None)
    z = {"p_ex": 0.0, "p_sted": 0.0, "pdt": 0.0, "target": 0.0}
    s, b = s or z, b or z
    return _lib.StedFluoCriteria(s["p_ex"], s["p_sted"], s["pdt"], s["target"], b["p_ex"], b["p_sted"], b["pdt"], b["target"])


class FluorescenceOptimizer:
    FACTORS = {"synthetic": 1.0, "clusters": 2.25, "camkii": 2.25, "psd95": 2.25, "PSD95-Bassoon": 2.25, "actin": 3.0,
               "tubulin": 3.75, "lifeact": 3.75}                       # utils.py:447-456

    def __init__(self, microscope: Microscope = None, sample="clusters", iterations=25, pixelsize=20e-9, device="cuda"):
        import torch
        if _lib.device_count() == 0:
            raise _lib.StedError(-2, "no CUDA device: gym_sted_b200 has no CPU fallback")
        self.lib = _lib.load()
        if microscope is None:
            from .envs import make_microscope
            microscope = make_microscope()
        self.microscope, self.iterations, self.pixelsize = microscope, int(iterations), pixelsize
        assert sample in self.FACTORS
        self.correction_factor, self.scale_factor = self.FACTORS[sample], 40.0
        self.device = torch.device(device)
        cache = microscope._packed_cache(pixelsize)
        self.K = cache.shape[-1]
        self._cache = torch.from_numpy(cache).to(self.device)
        self._blob_for = {}
        self.default_parameters()

    def set_correction_factor(self, sample):
        assert sample in self.FACTORS
        self.correction_factor = self.FACTORS[sample]

    def default_parameters(self):
        f = self.microscope.fluo
        f.sigma_abs, f.k1, f.b = dict(defaults.FLUO["sigma_abs"]), defaults.FLUO["k1"], defaults.FLUO["b"]
        return defaults.FLUO["k1"], defaults.FLUO["b"], defaults.FLUO["sigma_abs"]

    def _blob(self):
        import torch
        key = float(self.correction_factor)
        if key not in self._blob_for:
            blob, radius = gaussian_blob(self.K, key, self.scale_factor)
            self._blob_for[key] = (torch.from_numpy(blob).to(self.device), radius)
        return self._blob_for[key]

    def optimize_batch(self, criteria) -> np.ndarray:
        """``criteria``: sequence of criterion dicts / ``Criterion`` objects.  Returns a (B,5) float64 array:
        ``sigma_abs`` (m^2, at the excitation wavelength), ``k1``, ``b``, signal reached, bleach reached."""
        return self.optimize_batch_device(criteria)[0].cpu().numpy()

    def optimize_batch_device(self, criteria):
        """``optimize_batch`` queued on the current stream with nothing read back: returns ``(out, keep)``, the (B,5)
        device tensor and the input buffers the caller must keep alive until the stream has run.  A fit is one CTA per
        criterion and takes tens of milliseconds whatever the batch, so callers with several independent batches (one per
        correction factor, ``BleachSampler.sample_batch``) queue them on different streams and wait once."""
        import torch
        B = len(criteria)
        m = self.microscope
        self.default_parameters()
        arr = (_lib.StedFluoCriteria * B)(*[_criteria_struct(c) for c in criteria])
        crit = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device)
        out = torch.empty((B, 5), dtype=torch.float64, device=self.device)
        blob, radius = self._blob()
        fl = fluo_struct(m.fluo, m.excitation, m.sted)
        det = m.detector.params()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sted_fluo_optimize(self._cache.data_ptr(), blob.data_ptr(), self.K, radius, ctypes.byref(fl),
                                                   ctypes.byref(det), crit.data_ptr(), B, self.iterations, out.data_ptr(),
                                                   torch.cuda.current_stream(self.device).cuda_stream))
        return out, (crit, blob)

    def optimize(self, criterions) -> dict:
        """The reference's call and result layout (``utils.py:517-597`` / ``aggregate``, ``:493-515``)."""
        sigma, k1, b, _, _ = self.optimize_batch([criterions])[0]
        m, out = self.microscope, {}
        if "bleach" in criterions:
            out["bleach"] = {"k1": float(k1), "b": float(b)}
        if "signal" in criterions:
            lam_st = int(m.sted.lambda_ * 1e9)
            out["signal"] = {"sigma_abs": {int(m.excitation.lambda_ * 1e9): float(sigma), lam_st: m.fluo.sigma_abs[lam_st]}}
        return out

    # the closed forms themselves, for callers that use them directly (utils.py:599-661)
    def kb_map_to_im_bleach(self, kb_map, dwelltime, linestep):
        return 1 - np.exp((-kb_map * dwelltime * linestep).sum())

    def expected_bleach(self, p_ex, p_sted, pdt):
        m = self.microscope
        i_ex, i_sted, _ = m.cache(self.pixelsize)
        tau, trep = m.sted.tau, 1 / m.sted.rate
        phi_ex = m.fluo.get_photons(i_ex * p_ex, lambda_=m.excitation.lambda_)
        phi_sted = m.fluo.get_photons(i_sted * p_sted, lambda_=m.sted.lambda_)
        kb = m.fluo.get_k_bleach(m.excitation.lambda_, m.sted.lambda_, phi_ex, phi_sted * tau / trep, tau, trep, pdt)
        return self.kb_map_to_im_bleach(kb, pdt, 1)

    def expected_confocal_signal(self, p_ex, p_sted, pdt):
        m = self.microscope
        effective = m.get_effective(self.pixelsize, p_ex, p_sted)
        blob, _ = gaussian_blob(effective.shape[0], self.correction_factor, self.scale_factor)
        photons = m.fluo.get_photons(np.sum(effective * blob))
        return float(np.mean([m.detector.get_signal(photons, pdt, m.sted.rate) for _ in range(25)]))
