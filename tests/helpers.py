"""Shared builders for the test-suite (oracle + product objects on identical parameters)."""
import functools
import os

import numpy as np

from gym_sted_b200 import base, defaults

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PX = defaults.PIXELSIZE


def _det(noise=False, background=0.0):
    return {"noise": noise, "background": background}


@functools.lru_cache(maxsize=None)
def oracle_psf_cache():
    """(i_ex, i_sted, psf_det) from the oracle's scipy.integrate.quad beams; memoised on disk under
    tests/golden/ (regenerate with tests/golden/make_golden.py)."""
    path = os.path.join(GOLDEN, "oracle_psf_cache.npz")
    if os.path.exists(path):
        z = np.load(path)
        return z["i_ex"], z["i_sted"], z["psf_det"]
    from oracle import pysted_oracle as po        # lazy: bench.py's GPU arm imports this module for its builders only
    m = po.default_microscope(defaults.FLUO, _det(), defaults.LASER_EX, defaults.LASER_STED, defaults.OBJECTIVE)
    return m.cache(PX)


def oracle_microscope(fluo=None, noise=False, background=0.0):
    from oracle import pysted_oracle as po
    m = po.default_microscope(fluo or defaults.FLUO, _det(noise, background), defaults.LASER_EX, defaults.LASER_STED,
                              defaults.OBJECTIVE)
    m._cache[int(round(PX * 1e9))] = oracle_psf_cache()
    return m


def product_microscope(fluo=None, noise=False, background=0.0):
    return base.Microscope(base.GaussianBeam(**defaults.LASER_EX), base.DonutBeam(**defaults.LASER_STED),
                           base.Detector(**_det(noise, background)), base.Objective(**defaults.OBJECTIVE),
                           base.Fluorescence(**(fluo or defaults.FLUO)), load_cache=True)


def synthetic_datamap(rng, H, W, density=0.14, molecules=5, n_domains=6, domain=(100, 175)):
    """Sparse spine-like molecule map: a band of `molecules` per pixel plus single-pixel nanodomains."""
    d = np.zeros((H, W), dtype=np.int64)
    yy, xx = np.mgrid[:H, :W]
    cy, cx = H / 2 + rng.uniform(-H / 8, H / 8), W / 2 + rng.uniform(-W / 8, W / 8)
    ry, rx = H * rng.uniform(0.2, 0.35), W * rng.uniform(0.2, 0.35)
    rr = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2
    band = (rr < 1.0) & (rr > 0.45)
    d[band] = molecules
    ys, xs = np.nonzero(band)
    if len(ys):
        for i in rng.choice(len(ys), size=min(n_domains, len(ys)), replace=False):
            d[ys[i], xs[i]] = rng.integers(domain[0], domain[1] + 1)
    return d


def load_demonstrations():
    """The reference's recorded episodes (tests/golden/make_golden.py): dict of arrays, 50 records."""
    z = np.load(os.path.join(GOLDEN, "demonstrations.npz"))
    out = {k: z[k] for k in z.files}
    out["fg_c"] = np.unpackbits(out["fg_c"], axis=-1).astype(bool)
    out["fg_s"] = np.unpackbits(out["fg_s"], axis=-1).astype(bool)
    return out


def empty_tables(B, K):
    """Zeroed (B, STED_TABLES, K, KP) float32 table block as sted_acquire expects it (include/sted_b200.h)."""
    from gym_sted_b200 import _lib
    return np.zeros((B, _lib.STED_TABLES, K, _lib.kpitch(K)), dtype=np.float32)


def fill_e2(tab):
    """Derive the row-interleaved copy E2[u][v] = (E[u][v], E[u-1][v]) (slots STED_TAB_E2..) from slot STED_TAB_E,
    as sted_make_effective does."""
    from gym_sted_b200 import _lib
    B, _, K, KP = tab.shape
    for b in range(B):
        pad = np.zeros((K + 2, KP), dtype=np.float32)
        pad[1:K + 1] = tab[b, _lib.STED_TAB_E]
        e2 = np.stack([pad[1:], pad[:-1]], axis=-1)                    # (K+1, KP, 2)
        flat = tab[b, _lib.STED_TAB_E2:].reshape(-1)
        flat[:e2.size] = e2.reshape(-1)
    return tab
