"""Synthetic spine / nanocluster molecule maps (BASELINE.md §3 "Inputs").

Stand-in for ``pysted.exp_data_gen.Synapse`` as gym-sted configures it
(``gym_sted/envs/mosted_env.py:80-83``: mode "mushroom", 5 molecules per outline pixel, 3-15
nanodomains of 100-175 molecules, >= 100 nm = 5 px apart, ``valid_thickness=(3, 10)``;
``gym_sted/utils.py:30-31``): per 64x64 cell one elliptical spine-head outline band 3-10 px thick
and single-pixel nanodomains inside the band, randomly rotated and translated.  Larger maps tile
cells (256^2 = 4x4, 512^2 = 8x8).  Input generation only — not part of the timed path.
"""
from __future__ import annotations

import numpy as np

CELL = 64


class Synapse:
    """``pysted.exp_data_gen.Synapse`` as gym-sted drives it (``gym_sted/utils.py:57-65``):

        synapse = Synapse(5, mode="mushroom", seed=None)
        synapse.add_nanodomains((3, 15), min_dist_nm=100, seed=None, n_molecs_in_domain=(100, 175),
                                valid_thickness=(3, 10))
        synapse.rotate_and_translate()
        synapse.frame, synapse.nanodomains_coords, synapse.datamap_pixelsize_nm

    pySTED is not vendored in the reference, so the geometry is restated from what the reference's recorded
    episodes show (``tests/golden/demonstrations.npz``: a FILLED elliptical spine head of 25-50 x 15-25 px on a
    neck, attached to a dendrite bar along one image border, arbitrarily rotated; about 14 % of the frame
    occupied): ``n_molecs`` molecules in every pixel of the structure, nanodomains as single hot pixels inside the
    ``valid_thickness`` rim of the head, at least ``min_dist_nm`` apart.  Integer arguments may be given as
    ``(low, high)`` ranges (inclusive), drawn uniformly, like the reference passes them."""

    MODES = ("mushroom", "bump")

    def __init__(self, n_molecs=5, datamap_pixelsize_nm=20, width_nm=(500, 1000), height_nm=(300, 500),
                 img_shape=(64, 64), dendrite_thickness=(1, 10), mode="rand", seed=None, frame=None, coords=None):
        self.rng = np.random.default_rng(seed)
        self.n_molecs_base = int(n_molecs)
        self.datamap_pixelsize_nm = datamap_pixelsize_nm
        self.img_shape = tuple(img_shape)
        self.nanodomains = []
        self.nanodomains_coords = np.zeros((0, 2), dtype=np.int64) if coords is None else np.asarray(coords)
        if frame is not None:                       # container use: an already generated structure
            self.frame = np.asarray(frame)
            self.head = self.frame > 0
            self.mode = mode
            return
        if mode == "rand":
            mode = self.MODES[int(self.rng.integers(len(self.MODES)))]
        if mode not in self.MODES:
            raise ValueError(f"unknown synapse mode {mode!r}; expected one of {self.MODES + ('rand',)}")
        self.mode = mode
        H, W = self.img_shape
        px = float(datamap_pixelsize_nm)
        g = self.geometry = {"a": self._draw(width_nm) / px / 2.0, "b": self._draw(height_nm) / px / 2.0,   # head half axes (px)
                             "thick": int(round(self._draw(dendrite_thickness))), "cx": (W - 1) / 2.0}
        if mode == "bump":
            g["cy"], g["neck_w"] = H - g["thick"] - 0.5, 0.0        # half of the ellipse emerges from the dendrite
        else:
            neck_len = self.rng.uniform(0.2, 0.45) * H / 2
            g["neck_w"] = max(1.5, self.rng.uniform(0.15, 0.4) * g["a"])
            g["cy"] = H - g["thick"] - neck_len - g["b"]
        yy, xx = np.mgrid[:H, :W].astype(np.float64)
        self._render(yy, xx)

    def _render(self, ys, xs):
        """Evaluate the analytic structure at (possibly rotated) source coordinates ``ys, xs`` of every frame pixel."""
        g, H = self.geometry, self.img_shape[0]
        self.head = ((xs - g["cx"]) / g["a"]) ** 2 + ((ys - g["cy"]) / g["b"]) ** 2 <= 1.0
        dendrite = (ys >= H - g["thick"] - 0.5) & (ys <= H - 0.5)     # a band, unbounded along the bar
        neck = (np.abs(xs - g["cx"]) <= g["neck_w"]) & (ys >= g["cy"]) & (ys < H - g["thick"]) if g["neck_w"] > 0 \
            else np.zeros_like(self.head)
        self._rim_exclude = (neck | dendrite) & ~self.head
        self._shape = self.head | neck | dendrite
        self.frame = np.where(self._shape, self.n_molecs_base, 0).astype(np.int64)

    def _draw(self, v, integer=False):
        if isinstance(v, (tuple, list)):
            return int(self.rng.integers(int(v[0]), int(v[1]) + 1)) if integer else float(self.rng.uniform(v[0], v[1]))
        return int(v) if integer else float(v)

    def add_nanodomains(self, n_nanodomains=40, min_dist_nm=200, n_molecs_in_domain=5, seed=None, valid_thickness=3):
        """Single-pixel nanodomains in the rim (``valid_thickness`` pixels wide) of the spine head."""
        import scipy.ndimage as ndi
        if seed is not None:
            self.rng = np.random.default_rng(seed)
        n = self._draw(n_nanodomains, integer=True)
        thick = max(1, self._draw(valid_thickness, integer=True))
        min_d2 = (float(min_dist_nm) / float(self.datamap_pixelsize_nm)) ** 2
        rim = self.head & ~ndi.binary_erosion(self.head, iterations=thick)
        rim &= ~getattr(self, "_rim_exclude", np.zeros_like(rim))
        ys, xs = np.nonzero(rim)
        placed = [tuple(c) for c in np.asarray(self.nanodomains_coords).reshape(-1, 2)]
        target = len(placed) + n
        for i in self.rng.permutation(len(ys)):
            if len(placed) >= target:
                break
            p = (int(ys[i]), int(xs[i]))
            if all((p[0] - q[0]) ** 2 + (p[1] - q[1]) ** 2 >= min_d2 for q in placed):
                placed.append(p)
                m = self._draw(n_molecs_in_domain, integer=True)
                self.nanodomains.append(m)
                self.frame[p] += m
        self.nanodomains_coords = np.asarray(placed, dtype=np.int64).reshape(-1, 2)
        return self

    def rotate_and_translate(self, rot_angle=None, translate=True):
        """Random rigid motion of the structure; nanodomains stay single pixels (their coordinates are moved
        analytically and re-stamped), the base structure is resampled with nearest-neighbour interpolation."""
        H, W = self.img_shape
        ang = self.rng.uniform(0, 2 * np.pi) if rot_angle is None else float(rot_angle)
        c, s = np.cos(ang), np.sin(ang)
        cy, cx = (H - 1) / 2.0, (W - 1) / 2.0
        ty, tx = (self.rng.uniform(-0.12, 0.12, 2) * (H, W)) if translate else (0.0, 0.0)
        yy, xx = np.mgrid[:H, :W].astype(np.float64)
        # destination -> source (inverse map); the analytic structure is evaluated there, so nothing is resampled
        ys = c * (yy - cy - ty) + s * (xx - cx - tx) + cy
        xs = -s * (yy - cy - ty) + c * (xx - cx - tx) + cx
        if not hasattr(self, "geometry"):
            raise ValueError("rotate_and_translate needs a generated synapse (not a frame container)")
        self._render(ys, xs)
        frame = self.frame
        coords, kept = [], []
        for (y, x), m in zip(np.asarray(self.nanodomains_coords).reshape(-1, 2), self.nanodomains):
            ny = int(round(c * (y - cy) - s * (x - cx) + cy + ty))
            nx = int(round(s * (y - cy) + c * (x - cx) + cx + tx))
            if 0 <= ny < H and 0 <= nx < W and (ny, nx) not in coords:
                coords.append((ny, nx))
                kept.append(m)
                frame[ny, nx] = self.n_molecs_base + m
        self.frame, self.nanodomains = frame, kept
        self.nanodomains_coords = np.asarray(coords, dtype=np.int64).reshape(-1, 2)
        return self


def _cell(rng, molecules, n_nanodomains, n_molecs_in_domain, min_dist_px, valid_thickness):
    yy, xx = np.mgrid[:CELL, :CELL].astype(np.float64)
    cy, cx = CELL / 2 + rng.uniform(-6, 6, 2)
    a, b = rng.uniform(12, 20), rng.uniform(9, 16)
    th = rng.uniform(0, np.pi)
    thick = rng.uniform(*valid_thickness)
    dy, dx = yy - cy, xx - cx
    ru = (dx * np.cos(th) + dy * np.sin(th)) / a
    rv = (-dx * np.sin(th) + dy * np.cos(th)) / b
    rad = np.hypot(ru, rv)
    mean_r = 0.5 * (a + b)
    band = np.abs(rad - 1.0) * mean_r <= thick / 2
    frame = np.zeros((CELL, CELL), dtype=np.int32)
    frame[band] = molecules
    ys, xs = np.nonzero(band)
    n = int(rng.integers(n_nanodomains[0], n_nanodomains[1] + 1))
    coords = []
    order = rng.permutation(len(ys))
    for i in order:
        p = (int(ys[i]), int(xs[i]))
        if all((p[0] - q[0]) ** 2 + (p[1] - q[1]) ** 2 >= min_dist_px ** 2 for q in coords):
            coords.append(p)
            if len(coords) == n:
                break
    for (y, x) in coords:
        frame[y, x] = int(rng.integers(n_molecs_in_domain[0], n_molecs_in_domain[1] + 1))
    return frame, coords


def generate(H=64, W=64, seed=None, molecules=5, n_nanodomains=(3, 15), n_molecs_in_domain=(100, 175),
             min_dist_nm=100, valid_thickness=(3, 10), pixelsize_nm=20) -> Synapse:
    rng = np.random.default_rng(seed)
    frame = np.zeros((H, W), dtype=np.int32)
    coords = []
    for cy in range(0, H, CELL):
        for cx in range(0, W, CELL):
            cell, cc = _cell(rng, molecules, n_nanodomains, n_molecs_in_domain, min_dist_nm / pixelsize_nm,
                             valid_thickness)
            h, w = min(CELL, H - cy), min(CELL, W - cx)
            frame[cy:cy + h, cx:cx + w] = cell[:h, :w]
            coords += [(cy + y, cx + x) for (y, x) in cc if y < h and x < w]
    return Synapse(molecules, pixelsize_nm, frame=frame, coords=np.asarray(coords, dtype=np.int64).reshape(-1, 2))


def generate_batch(B, H, W, seed=1234, env_base=0, **kw):
    """(B,H,W) int32 maps + list of coordinate arrays; env i uses the key (seed, env_base+i)."""
    maps = np.empty((B, H, W), dtype=np.int32)
    coords = []
    for i in range(B):
        s = generate(H, W, seed=[seed, env_base + i], **kw)
        maps[i] = s.frame
        coords.append(s.nanodomains_coords)
    return maps, coords


class DeviceSynapseGenerator:
    """Structures for B environments generated ON the device (``sted_synapse_generate``, ``csrc/env_kernels.cu``):
    the same construction as ``generate`` (per 64 x 64 cell one elliptical outline band, single-pixel nanodomains inside
    the band at least ``min_dist_px`` apart), one CTA per cell, Philox keyed by (seed, env_base + b, cell, structure
    number) so the maps do not depend on the batch split.  Nothing is read back: frames, ground-truth coordinates and
    their counts stay on the device for the acquisition and the F1 kernel.  Checked bit for bit against
    ``oracle/synapse_oracle.py`` (``tests/test_gpu_synapse.py``)."""

    def __init__(self, B, H, W, device="cuda", seed=0, env_base=0, molecules=5, n_nanodomains=(3, 15),
                 n_molecs_in_domain=(100, 175), min_dist_px=5.0, valid_thickness=(3, 10)):
        import torch
        from . import _lib
        self.lib = _lib.load()
        if _lib.device_count() == 0:
            raise _lib.StedError(-2, "no CUDA device: gym_sted_b200 has no CPU fallback")
        self._check = _lib.check
        self.B, self.H, self.W = int(B), int(H), int(W)
        self.device = torch.device(device)
        self.seed, self.env_base, self.offset = int(seed), int(env_base), 0
        self.cells = -(-self.H // CELL) * -(-self.W // CELL)
        self.max_truth = self.cells * int(n_nanodomains[1])
        self.max_molecules = self.B * self.cells * (int(molecules) * CELL * CELL + int(n_nanodomains[1]) * int(n_molecs_in_domain[1]))
        self.params = _lib.StedSynapseParams(int(molecules), int(n_nanodomains[0]), int(n_nanodomains[1]),
                                             int(n_molecs_in_domain[0]), int(n_molecs_in_domain[1]), 0,
                                             float(min_dist_px), float(valid_thickness[0]), float(valid_thickness[1]))
        self.scratch = torch.empty((int(self.lib.sted_synapse_scratch_bytes(self.B, self.H, self.W)),), dtype=torch.uint8,
                                   device=self.device)

    def generate(self, offset=None):
        """-> (frames int32 (B,H,W), coords int32 (B,max_truth,2), counts int32 (B,)), all on the device; queued on the
        current stream.  ``offset``: structure number (defaults to a counter that advances with every call)."""
        import ctypes
        import torch
        if offset is None:
            offset = self.offset
            self.offset += 1
        frames = torch.empty((self.B, self.H, self.W), dtype=torch.int32, device=self.device)
        coords = torch.empty((self.B, self.max_truth, 2), dtype=torch.int32, device=self.device)
        counts = torch.empty((self.B,), dtype=torch.int32, device=self.device)
        self._check(self.lib.sted_synapse_generate(
            ctypes.byref(self.params), self.B, self.H, self.W, self.seed, int(offset), self.env_base, frames.data_ptr(),
            coords.data_ptr(), counts.data_ptr(), self.max_truth, self.scratch.data_ptr(),
            torch.cuda.current_stream(self.device).cuda_stream))
        return frames, coords, counts
