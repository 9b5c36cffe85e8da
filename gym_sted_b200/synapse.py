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
    """Minimal ``exp_data_gen.Synapse`` surface used by the reference: ``frame``,
    ``nanodomains_coords`` and ``datamap_pixelsize_nm`` (``gym_sted/rewards/objectives.py:421-425``)."""

    def __init__(self, frame, coords, pixelsize_nm=20):
        self.frame = frame
        self.nanodomains_coords = coords
        self.datamap_pixelsize_nm = pixelsize_nm


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
    return Synapse(frame, np.asarray(coords, dtype=np.int64).reshape(-1, 2), pixelsize_nm)


def generate_batch(B, H, W, seed=1234, env_base=0, **kw):
    """(B,H,W) int32 maps + list of coordinate arrays; env i uses the key (seed, env_base+i)."""
    maps = np.empty((B, H, W), dtype=np.int32)
    coords = []
    for i in range(B):
        s = generate(H, W, seed=[seed, env_base + i], **kw)
        maps[i] = s.frame
        coords.append(s.nanodomains_coords)
    return maps, coords
