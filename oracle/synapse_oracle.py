"""TEST INFRASTRUCTURE — CPU restatement of the device synapse generator (``sted_synapse_generate``).

Only ``tests/`` may import this module; the product path never does.

What it restates: the synthetic spine datamap of SURVEY.md §8d that stands in for
``pysted.exp_data_gen.Synapse`` as gym-sted configures it (``gym_sted/envs/mosted_env.py:80-83,168-187``,
``gym_sted/utils.py:26-78``: 5 molecules per structure pixel, 3-15 nanodomains of 100-175 molecules at least
100 nm = 5 px apart, ``valid_thickness=(3, 10)``).  pySTED is not vendored in the reference, so there is nothing
upstream to pin the geometry against (parity unpinned for this row); what IS pinned is that the kernel and this
independent numpy restatement produce the same integers from the same Philox4x32-10 words: per 64 x 64 cell one
elliptical outline band, nanodomains by dart throwing in the order of a random priority.
"""
from __future__ import annotations

import numpy as np

CELL = 64
RNG_SYNAPSE = 7
M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    """Salmon et al. SC'11.  ``ctr``: (..., 4) uint32-valued array, ``key``: (k0, k1) -> (..., 4) uint64 array of words."""
    c = [np.asarray(ctr[..., i], dtype=np.uint64) for i in range(4)]
    k0, k1 = int(key[0]) & MASK, int(key[1]) & MASK
    for _ in range(10):
        p0 = np.uint64(M0) * c[0]
        p1 = np.uint64(M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & np.uint64(MASK)
        hi1, lo1 = p1 >> np.uint64(32), p1 & np.uint64(MASK)
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return np.stack(c, axis=-1)


def _u01(w):
    return (np.asarray(w, dtype=np.float64) + 0.5) * 2.0 ** -32


def _pick(w, lo, hi):
    return int(lo) + ((int(w) * (int(hi) - int(lo) + 1)) >> 32)


def _blocks(idx, env, cell, offset, seed):
    idx = np.atleast_1d(np.asarray(idx, dtype=np.uint64))
    ctr = np.stack([idx, np.full_like(idx, env), np.full_like(idx, (RNG_SYNAPSE << 28) | cell), np.full_like(idx, offset & MASK)], axis=-1)
    return philox4x32_10(ctr, (seed & MASK, (seed >> 32) & MASK))


def cell(env, cell_index, seed, offset, molecules=5, n_nanodomains=(3, 15), n_molecs_in_domain=(100, 175), min_dist_px=5.0,
         valid_thickness=(3, 10)):
    """One 64 x 64 cell -> (frame int32 (64, 64), [(y, x), ...] in acceptance order)."""
    g = _blocks([0, 1], env, cell_index, offset, seed)
    cy = CELL / 2 + (12.0 * _u01(g[0, 0]) - 6.0)
    cx = CELL / 2 + (12.0 * _u01(g[0, 1]) - 6.0)
    a = 12.0 + 8.0 * _u01(g[0, 2])
    b = 9.0 + 7.0 * _u01(g[0, 3])
    thick = valid_thickness[0] + (valid_thickness[1] - valid_thickness[0]) * _u01(g[1, 0])
    n = _pick(g[1, 1], *n_nanodomains)
    c, s = 1.0, 0.0
    for d in _blocks(np.arange(2, 32), env, cell_index, offset, seed):
        found = False
        for x, y in ((d[0], d[1]), (d[2], d[3])):
            x, y = 2.0 * _u01(x) - 1.0, 2.0 * _u01(y) - 1.0
            r2 = x * x + y * y
            if 1e-4 < r2 <= 1.0:
                r = np.sqrt(r2)
                c, s, found = x / r, y / r, True
                break
        if found:
            break
    yy, xx = np.mgrid[:CELL, :CELL].astype(np.float64)
    dy, dx = yy - cy, xx - cx
    ru = (dx * c + dy * s) / a
    rv = (dy * c - dx * s) / b
    rad = np.sqrt(ru * ru + rv * rv)
    band = np.abs(rad - 1.0) * (0.5 * (a + b)) <= thick / 2.0
    prio = _blocks(64 + np.arange(CELL * CELL // 4), env, cell_index, offset, seed).reshape(-1)     # word pix % 4 of block 64 + pix / 4
    pix = np.flatnonzero(band.ravel())
    order = pix[np.lexsort((pix, -prio[pix].astype(np.int64)))]           # priority descending, lower pixel index first
    accepted = []
    for p in order:
        if len(accepted) >= n:
            break
        y, x = divmod(int(p), CELL)
        if all(float((y - ay) ** 2 + (x - ax) ** 2) >= min_dist_px * min_dist_px for ay, ax in accepted):
            accepted.append((y, x))
    frame = np.where(band, molecules, 0).astype(np.int32)
    if accepted:
        mol = _blocks(32 + np.arange((len(accepted) + 3) // 4), env, cell_index, offset, seed).reshape(-1)
        for k, (y, x) in enumerate(accepted):
            frame[y, x] = _pick(mol[k], *n_molecs_in_domain)
    return frame, accepted


def generate_batch(B, H, W, seed, offset=0, env_base=0, **kw):
    """(B, H, W) int32 frames and a list of (n_i, 2) int64 coordinate arrays (cell order, clipped to the image)."""
    ny, nx = -(-H // CELL), -(-W // CELL)
    frames = np.zeros((B, H, W), dtype=np.int32)
    coords = []
    for b in range(B):
        cc = []
        for ci in range(ny * nx):
            y0, x0 = (ci // nx) * CELL, (ci % nx) * CELL
            f, acc = cell(env_base + b, ci, seed, offset, **kw)
            h, w = min(CELL, H - y0), min(CELL, W - x0)
            frames[b, y0:y0 + h, x0:x0 + w] = f[:h, :w]
            cc += [(y0 + y, x0 + x) for y, x in acc if y0 + y < H and x0 + x < W]
        coords.append(np.asarray(cc, dtype=np.int64).reshape(-1, 2))
    return frames, coords
