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


def generate_batch_torch(B, H, W, device="cuda", generator=None, molecules=5, n_nanodomains=(3, 15),
                         n_molecs_in_domain=(100, 175), min_dist_px=5, valid_thickness=(3, 10)):
    """Batched generator on the device: same construction as ``generate`` (elliptical outline band per 64x64 cell,
    single-pixel nanodomains inside the band, at least ``min_dist_px`` apart), vectorised over B x cells.
    The spacing comes from keeping only local maxima of a random priority map inside a (2*min_dist-1)^2 window.
    Returns ``(frames int32 (B,H,W) on the device, [coords (n_i,2) int64 numpy arrays])``."""
    import torch
    import torch.nn.functional as F

    dev = torch.device(device)
    ny, nx = -(-H // CELL), -(-W // CELL)
    C = B * ny * nx
    g = generator
    u = lambda lo, hi: lo + (hi - lo) * torch.rand((C, 1, 1), device=dev, generator=g)
    yy = torch.arange(CELL, device=dev, dtype=torch.float32).view(1, CELL, 1)
    xx = torch.arange(CELL, device=dev, dtype=torch.float32).view(1, 1, CELL)
    cy, cx, a, b = u(CELL / 2 - 6, CELL / 2 + 6), u(CELL / 2 - 6, CELL / 2 + 6), u(12, 20), u(9, 16)
    th, thick = u(0, float(np.pi)), u(*valid_thickness)
    dy, dx = yy - cy, xx - cx
    rad = torch.hypot((dx * torch.cos(th) + dy * torch.sin(th)) / a, (-dx * torch.sin(th) + dy * torch.cos(th)) / b)
    band = (rad - 1.0).abs() * (0.5 * (a + b)) <= thick / 2
    prio = torch.where(band, torch.rand((C, CELL, CELL), device=dev, generator=g) + 1e-6, torch.zeros((), device=dev))
    win = 2 * int(min_dist_px) - 1
    local_max = prio == F.max_pool2d(prio[:, None], win, stride=1, padding=win // 2)[:, 0]
    score = torch.where(band & local_max, prio, torch.zeros((), device=dev)).view(C, -1)
    kmax = n_nanodomains[1]
    top, pos = torch.topk(score, kmax, dim=1)
    n = torch.randint(n_nanodomains[0], n_nanodomains[1] + 1, (C, 1), device=dev, generator=g)
    valid = (top > 0) & (torch.arange(kmax, device=dev)[None, :] < n)
    counts = torch.randint(n_molecs_in_domain[0], n_molecs_in_domain[1] + 1, (C, kmax), device=dev, generator=g)
    cells = (band.to(torch.int32) * molecules).view(C, -1)
    cells.scatter_(1, pos, torch.where(valid, counts.to(torch.int32), cells.gather(1, pos)))
    cells = cells.view(B, ny, nx, CELL, CELL).permute(0, 1, 3, 2, 4).reshape(B, ny * CELL, nx * CELL)
    frames = cells[:, :H, :W].contiguous()
    # coordinates on the host (ground truth of the F1 reward)
    cyx = torch.stack((pos // CELL, pos % CELL), dim=-1).view(B, ny, nx, kmax, 2)
    off = torch.stack(torch.meshgrid(torch.arange(ny, device=dev) * CELL, torch.arange(nx, device=dev) * CELL, indexing="ij"), dim=-1)
    cyx = (cyx + off[None, :, :, None, :]).reshape(B, -1, 2)
    ok = valid.view(B, -1) & (cyx[..., 0] < H) & (cyx[..., 1] < W)
    cyx_h, ok_h = cyx.cpu().numpy().astype(np.int64), ok.cpu().numpy()
    return frames, [cyx_h[i][ok_h[i]] for i in range(B)]
