"""Resident structures for host-buffer callers: ``StedSession`` wraps the ``sted_session_*`` entry points
(``include/sted_b200.h``).

The reference keeps ONE ``pysted.base.Datamap`` per structure and images it three or four times
(``gym_sted/envs/mosted_env.py:222-234``); ``sted_acquire_host`` re-uploads the map on every call.  A session uploads it
once, advances it on the device and queues uploads / acquisitions / downloads asynchronously; the host only waits in
``sync``.  numpy arrays in, numpy arrays out (pinned arrays from ``pinned_empty`` make the copies truly asynchronous).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .base import Microscope, fluo_struct

_VP = ctypes.c_void_p


def pinned_empty(shape, dtype):
    """Page-locked numpy array (through torch's pinned allocator; torch is device plumbing only)."""
    import torch
    t = torch.empty(tuple(shape), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
    return _Pinned(t.numpy(), t)


class _Pinned(np.ndarray):
    """ndarray view that keeps its pinned torch storage alive."""

    def __new__(cls, array, owner):
        obj = np.asarray(array).view(cls)
        obj._owner = owner
        return obj

    def __array_finalize__(self, obj):
        self._owner = getattr(obj, "_owner", None)


class StedSession:
    def __init__(self, microscope: Microscope, B: int, H: int, W: int, pixelsize: float = 20e-9, fluos=None,
                 low_priority: bool = False, env_base: int = 0, seed: int = 0, u16: bool = True):
        self.lib = _lib.load()
        if _lib.device_count() == 0:
            raise _lib.StedError(-2, "no CUDA device: gym_sted_b200 has no CPU fallback")
        self.microscope, self.B, self.H, self.W = microscope, int(B), int(H), int(W)
        cache = microscope._packed_cache(pixelsize)
        self.K = cache.shape[-1]
        self.env_base, self.seed, self.offset, self.u16 = int(env_base), int(seed), 0, bool(u16)
        self.det = microscope.detector.params()
        self.lambda_fluo = float(microscope.fluo.lambda_)
        handle = _VP()
        _lib.check(self.lib.sted_session_create(self.B, self.H, self.W, self.K, cache.ctypes.data_as(_VP),
                                                int(bool(low_priority)), ctypes.byref(handle)))
        self._h = handle
        self.set_fluorophores(fluos)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.sted_session_destroy(self._h)
            self._h = None

    __del__ = close

    def set_fluorophores(self, fluos=None):
        m = self.microscope
        fluos = [m.fluo] * self.B if fluos is None else list(fluos)
        assert len(fluos) == self.B
        arr = (_lib.StedFluoParams * self.B)(*[fluo_struct(f, m.excitation, m.sted) for f in fluos])
        _lib.check(self.lib.sted_session_set_fluorophores(self._h, ctypes.cast(arr, _VP)))

    def _map_flags(self, a):
        if a.dtype == np.uint16:
            return _lib.STED_HOST_U16
        if a.dtype == np.float32:
            return 0
        raise TypeError("molecule maps must be uint16 or float32 arrays")

    def upload(self, maps: np.ndarray, max_molecules: int = 0):
        """Queue the upload of (B,H,W) un-padded molecule maps (uint16 or float32, C-contiguous)."""
        assert maps.shape == (self.B, self.H, self.W) and maps.flags.c_contiguous
        self._keep = maps
        _lib.check(self.lib.sted_session_upload(self._h, maps.ctypes.data_as(_VP), self._map_flags(maps), int(max_molecules)))

    def upload_device(self, maps, max_molecules: int = 0):
        """Same from a (B,H,W) float32 / uint16 torch tensor on the session's device, ready on torch's current stream."""
        import torch
        assert tuple(maps.shape) == (self.B, self.H, self.W) and maps.is_contiguous() and maps.is_cuda
        flags = {torch.float32: 0, torch.uint16: _lib.STED_HOST_U16}[maps.dtype]
        self._keep = maps
        _lib.check(self.lib.sted_session_upload_device(self._h, maps.data_ptr(), flags, int(max_molecules),
                                                       torch.cuda.current_stream(maps.device).cuda_stream))

    def result(self, back: int = 0, intensity: bool = False, wait: bool = True):
        """Device view (torch, float32 (B,H,W), no copy) of the photon counts of the acquisition queued ``back`` calls
        ago; torch's current stream is made to wait for it.  Overwritten by the fourth acquisition after its own."""
        import torch
        sig, inten = _VP(), _VP()
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(self.lib.sted_session_result(self._h, int(back), ctypes.byref(sig), ctypes.byref(inten), st, int(wait)))

        def view(ptr):
            arr = {"data": (ptr.value, False), "shape": (self.B, self.H, self.W), "typestr": "<f4", "version": 3}
            return torch.as_tensor(type("_Dev", (), {"__cuda_array_interface__": arr})(), device="cuda")
        return (view(sig), view(inten)) if intensity else view(sig)

    def kernel_ms(self, back: int = 0) -> float:
        """Device time of the gather / scan kernel of the acquisition queued ``back`` calls ago (waits for it)."""
        ms = ctypes.c_float(0)
        _lib.check(self.lib.sted_session_kernel_ms(self._h, int(back), ctypes.byref(ms)))
        return float(ms.value)

    def join(self):
        """torch's current stream waits for everything queued on the session so far."""
        import torch
        _lib.check(self.lib.sted_session_join(self._h, torch.cuda.current_stream().cuda_stream))

    def _vec(self, x):
        return np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), (self.B,)))

    def acquire(self, pdt, p_ex, p_sted, bleach=True, sample_bleach=True, noise=None, out=None, intensity_out=None,
                seed=None):
        """Queue one acquisition (``Microscope.get_signal_and_bleach`` for the B resident structures).  ``out``:
        (B,H,W) uint16 or float32 array receiving the photon counts after ``sync`` (None: keep them on the device)."""
        flags = 0
        if bleach:
            flags |= _lib.STED_BLEACH | (_lib.STED_SAMPLE_BLEACH if sample_bleach else 0)
        if self.microscope.detector.noise if noise is None else noise:
            flags |= _lib.STED_SAMPLE_PHOTONS
        if out is not None:
            assert out.shape == (self.B, self.H, self.W) and out.flags.c_contiguous
            flags |= self._map_flags(out)
        self.offset += 1
        pdt, p_ex, p_sted = self._vec(pdt), self._vec(p_ex), self._vec(p_sted)
        _lib.check(self.lib.sted_session_acquire(
            self._h, p_ex.ctypes.data_as(_VP), p_sted.ctypes.data_as(_VP), pdt.ctypes.data_as(_VP), flags,
            ctypes.byref(self.det), self.lambda_fluo, self.seed if seed is None else int(seed), self.offset, self.env_base,
            out.ctypes.data_as(_VP) if out is not None else None,
            intensity_out.ctypes.data_as(_VP) if intensity_out is not None else None))
        return out

    def adopt(self, other: "StedSession"):
        """Take over ``other``'s current maps (queued device-to-device copy): the structure ``other`` uploaded and imaged
        as the NEXT structure becomes this session's current one."""
        _lib.check(self.lib.sted_session_adopt(self._h, other._h))

    def after(self, other: "StedSession"):
        """The gather / scan kernels queued on this session from now on start after those queued on ``other`` so far."""
        _lib.check(self.lib.sted_session_after(self._h, other._h))

    def download(self, out: np.ndarray):
        """Queue the download of the current (bleached) maps into ``out`` ((B,H,W) uint16 or float32)."""
        assert out.shape == (self.B, self.H, self.W) and out.flags.c_contiguous
        _lib.check(self.lib.sted_session_download(self._h, out.ctypes.data_as(_VP), self._map_flags(out)))
        return out

    def sync(self) -> int:
        dropped = ctypes.c_int64(0)
        _lib.check(self.lib.sted_session_sync(self._h, ctypes.byref(dropped)))
        return int(dropped.value)
