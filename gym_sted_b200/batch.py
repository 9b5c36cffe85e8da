"""Batched acquisition: B environments per kernel launch (SURVEY.md §8b "batched entry").

``BatchedMicroscope`` owns the device-resident state of B independent environments — padded
molecule maps (ping-pong pair), per-env fluorophore structs and PSF tables — and exposes the
same call as the reference, vectorised:

    signal, intensity = bm.get_signal_and_bleach(pdt, p_ex, p_sted, bleach=True)

with ``pdt/p_ex/p_sted`` float64 tensors of shape (B,).  PyTorch is used for device memory and
streams only; all arithmetic runs in ``libsted_b200.so``.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .base import Microscope, fluo_struct


class BatchedMicroscope:
    def __init__(self, microscope: Microscope, B: int, H: int, W: int, pixelsize: float = 20e-9,
                 device="cuda", fluos=None, env_base: int = 0, seed: int = 0):
        self.lib = _lib.load()
        if _lib.device_count() == 0:
            raise _lib.StedError(-2, "no CUDA device: gym_sted_b200 has no CPU fallback")
        self.microscope = microscope
        self.B, self.H, self.W = int(B), int(H), int(W)
        self.device = torch.device(device)
        cache = microscope._packed_cache(pixelsize)
        self.K = K = cache.shape[-1]
        self.KP = _lib.kpitch(K)
        self.p = K // 2
        self.Hp, self.Wp = H + K - 1, W + K - 1
        self.Wpitch = (self.Wp + 3) & ~3
        self.env_base, self.seed, self.offset = int(env_base), int(seed), 0
        self.d_cache = torch.from_numpy(cache).to(self.device)
        self.set_fluorophores(fluos)
        self.tables = torch.empty((B, _lib.STED_TABLES, K, self.KP), dtype=torch.float32, device=self.device)
        self.dmap = torch.zeros((B, self.Hp, self.Wpitch), dtype=torch.float32, device=self.device)
        self.dmap_alt = torch.zeros_like(self.dmap)
        self.det = microscope.detector.params()
        self.lambda_fluo = float(microscope.fluo.lambda_)
        self.workspace, self.workspace_bytes = None, 0

    # per-env fluorophores (config 3: one parameter set per environment)
    def set_fluorophores(self, fluos=None):
        m = self.microscope
        if fluos is None:
            fluos = [m.fluo] * self.B
        assert len(fluos) == self.B
        arr = (_lib.StedFluoParams * self.B)(*[fluo_struct(f, m.excitation, m.sted) for f in fluos])
        self.d_fluo = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device)
        self._kept_tables = {}                # tables depend on the fluorophores

    def set_datamaps(self, roi: torch.Tensor, max_molecules=None):
        """roi: (B,H,W) molecule counts (any real dtype); pads with the K//2 zero frame.

        ``max_molecules`` bounds the total number of molecules of the batch (it sizes the scratch of the
        stochastic scan); when omitted it is read back from the device once (one synchronisation)."""
        assert roi.shape == (self.B, self.H, self.W)
        p = self.p
        if roi.is_cuda and roi.device == self.dmap.device and roi.dtype in (torch.int32, torch.float32) and roi.is_contiguous():
            _lib.check(self.lib.sted_pad_maps(roi.data_ptr(), int(roi.dtype == torch.int32), self.B, self.H, self.W, self.K,
                                              self.dmap.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream))
        else:
            self.dmap.zero_()
            self.dmap[:, p:p + self.H, p:p + self.W] = roi.to(self.device, torch.float32)
        if max_molecules is None:
            max_molecules = int(torch.clamp(roi, min=0).sum().item())
        self._reserve(int(max_molecules) + 16)

    def _reserve(self, max_events: int):
        need = int(self.lib.sted_bleach_workspace_bytes(self.B, self.H, max_events))
        if self.workspace is None or self.workspace.numel() < need:
            self.workspace = torch.empty((need,), dtype=torch.uint8, device=self.device)
        self.workspace_bytes = need

    def dropped_events(self) -> int:
        """Deaths the last stochastic scan could not store (0 unless ``max_molecules`` was too small)."""
        out = ctypes.c_int(0)
        _lib.check(self.lib.sted_bleach_dropped_events(self.workspace.data_ptr(), ctypes.byref(out),
                                                       torch.cuda.current_stream(self.device).cuda_stream))
        return out.value

    def datamaps_roi(self) -> torch.Tensor:
        p = self.p
        return self.dmap[:, p:p + self.H, p:p + self.W]

    def make_effective(self, p_ex, p_sted, pdt, want_kbleach=False):
        st = torch.cuda.current_stream(self.device).cuda_stream
        kb = torch.empty((self.B, self.K, self.K), dtype=torch.float64, device=self.device) if want_kbleach else None
        _lib.check(self.lib.sted_make_effective(
            self.d_cache.data_ptr(), self.d_fluo.data_ptr(), p_ex.data_ptr(), p_sted.data_ptr(), pdt.data_ptr(),
            self.B, self.K, self.tables.data_ptr(), kb.data_ptr() if want_kbleach else None, st))
        return kb

    def _vec(self, x):
        if not torch.is_tensor(x):
            x = torch.as_tensor(np.broadcast_to(np.asarray(x, dtype=np.float64), (self.B,)).copy())
        return x.to(self.device, torch.float64).contiguous()

    def get_signal_and_bleach(self, pdt, p_ex, p_sted, bleach=True, update=True, sample_bleach=True,
                              noise=None, want_intensity=False, seed=None, out_signal=None, tables_key=None):
        """Vectorised ``Microscope.get_signal_and_bleach``; returns ``(signal, intensity|None)`` as
        (B,H,W) float32 device tensors.  With ``bleach`` and ``update`` the resident maps advance.

        ``tables_key``: a name under which the per-environment PSF tables of THESE imaging parameters are kept, so later
        calls with the same key skip ``sted_make_effective`` (gym-sted's confocal acquisitions always use the same
        parameters, ``mosted_env.py:184,222,232``); the caller vouches that the parameters of a key do not change, and
        ``set_fluorophores`` drops every kept set."""
        pdt, p_ex, p_sted = self._vec(pdt), self._vec(p_ex), self._vec(p_sted)
        tables = self.tables
        if tables_key is None:
            self.make_effective(p_ex, p_sted, pdt)
        elif tables_key in self._kept_tables:
            tables = self._kept_tables[tables_key]
        else:
            self.make_effective(p_ex, p_sted, pdt)
            tables = self._kept_tables[tables_key] = self.tables.clone()
        flags = 0
        if bleach:
            flags |= _lib.STED_BLEACH | (_lib.STED_SAMPLE_BLEACH if sample_bleach else 0)
        if self.microscope.detector.noise if noise is None else noise:
            flags |= _lib.STED_SAMPLE_PHOTONS
        self.offset += 1
        signal = out_signal if out_signal is not None else torch.empty((self.B, self.H, self.W), dtype=torch.float32,
                                                                         device=self.device)
        intensity = torch.empty_like(signal) if want_intensity else None
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.sted_acquire(
            self.dmap.data_ptr(), self.dmap_alt.data_ptr() if bleach else None, tables.data_ptr(),
            pdt.data_ptr(), self.B, self.H, self.W, self.K, flags, ctypes.byref(self.det), self.lambda_fluo,
            self.seed if seed is None else int(seed), self.offset, self.env_base,
            intensity.data_ptr() if want_intensity else None, signal.data_ptr(),
            self.workspace.data_ptr() if self.workspace is not None else None, self.workspace_bytes, st))
        if bleach and update:
            self.dmap, self.dmap_alt = self.dmap_alt, self.dmap
        return signal, intensity

    def bleached_roi(self) -> torch.Tensor:
        """ROI of the last bleach output when ``update=False`` was used."""
        p = self.p
        return self.dmap_alt[:, p:p + self.H, p:p + self.W]
