"""Regenerates tests/golden/demonstrations.npz from the reference's recorded episodes.

Source: /root/reference/gym_sted/prefnet/demonstrations/demonstrations.pbz2 (loader:
gym_sted/prefnet/articulation.py:13-34; structure: SURVEY.md Appendix B) — 5 episodes x 10 steps of
{sted_image, conf1, conf2, bleached, fg_c, fg_s (64x64), mo_objs [Resolution, Bleach, SNR], f1-score,
nanodomains-coords, action}.  These records pin the reward side of the path (Otsu foregrounds and the
three objectives); the GPU box has no /root/reference, so only the .npz travels.

    python tests/golden/make_golden.py
"""
import bz2
import os
import pickle

import numpy as np

SRC = "/root/reference/gym_sted/prefnet/demonstrations/demonstrations.pbz2"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "demonstrations.npz")


def main():
    with bz2.open(SRC, "rb") as fh:
        episodes = pickle.load(fh)
    recs = [r for ep in episodes for r in ep]
    out = {k: np.stack([r[k] for r in recs]).astype(np.int32) for k in ("sted_image", "conf1", "conf2", "bleached")}
    out["fg_c"] = np.packbits(np.stack([r["fg_c"] for r in recs]), axis=-1)
    out["fg_s"] = np.packbits(np.stack([r["fg_s"] for r in recs]), axis=-1)
    out["mo_objs"] = np.array([[np.nan if v is None else v for v in r["mo_objs"]] for r in recs], dtype=np.float64)
    out["f1"] = np.array([r["f1-score"] for r in recs], dtype=np.float64)
    out["action"] = np.stack([r["action"] for r in recs]).astype(np.float32)
    nmax = max(len(r["nanodomains-coords"]) for r in recs)
    coords = np.full((len(recs), nmax, 2), -1, dtype=np.int16)
    for i, r in enumerate(recs):
        c = np.asarray(r["nanodomains-coords"]).reshape(-1, 2)
        coords[i, :len(c)] = c
    out["coords"] = coords
    out["episode"] = np.repeat(np.arange(len(episodes)), [len(ep) for ep in episodes]).astype(np.int16)
    np.savez_compressed(DST, **out)
    print(DST, os.path.getsize(DST), "bytes;", len(recs), "records")


if __name__ == "__main__":
    main()
