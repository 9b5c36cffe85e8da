"""float64 numpy model of the row-sweep schedule used by ``sweep_kernel`` (csrc/sted_kernels.cu).

Test-side only.  It lets the CPU suite check the *derivation*  against the
sequential oracle without a GPU: per scan row, gather against the row-start state with in-row
survival folded into the weights, then thin the K-row band once.
"""
import numpy as np


def tables(E, A):
    """E, A=(k_b*pdt): (K,K).  Returns ET, CH (K,K+1), as sted_make_effective builds them."""
    K = E.shape[0]
    CH = np.zeros((K, K + 1))
    CH[:, :K] = np.cumsum(A[:, ::-1], axis=1)[:, ::-1]      # CH[u][v] = sum_{v'>=v} A[u][v']
    ET = E * np.exp(-CH[:, 1:])
    return ET, CH


def sweep_expectation(D, H, W, E, A):
    """D: (Hp,Wp) float64, returns (intensity (H,W), bleached D)."""
    K = E.shape[0]
    ET, CH = tables(E, A)
    Wp = W + K - 1
    D = D.copy()
    out = np.zeros((H, W))
    x = np.arange(Wp)
    lo = np.maximum(0, x - W + 1)
    hi = np.minimum(K - 1, x)
    for r in range(H):
        band = D[r:r + K]                                   # view, band[u] = row r+u
        # in-row clip: pixel (u,x) was only thinned by taps v' <= hi(x)
        scaled = band * np.exp(CH[:, hi + 1])               # (K,Wp) * (K,Wp)
        for c in range(W):
            out[r, c] = (ET * scaled[:, c:c + K]).sum()
        band *= np.exp(-(CH[np.arange(K)[:, None], lo[None, :]] - CH[np.arange(K)[:, None], hi[None, :] + 1]))
    return out, D


def sweep_stochastic(D, H, W, E, A, rng):
    """Integer thinning with one binomial per (pixel, scan row) + hazard-inverted death taps."""
    K = E.shape[0]
    _, CH = tables(E, A)
    Wp = W + K - 1
    D = D.astype(np.int64).copy()
    out = np.zeros((H, W))
    for r in range(H):
        band = D[r:r + K]
        row = np.zeros(W)
        for c in range(W):
            row[c] = (E * band[:, c:c + K]).sum()
        ys, xs = np.nonzero(band)
        for u, x in zip(ys, xs):
            lo, hi = max(0, x - W + 1), min(K - 1, x)
            hz = CH[u, lo] - CH[u, hi + 1]
            if hz <= 0:
                continue
            n = band[u, x]
            dead = rng.binomial(n, -np.expm1(-hz))
            if dead == 0:
                continue
            band[u, x] = n - dead
            pd = -np.expm1(-hz)
            for _ in range(dead):
                t = -np.log1p(-rng.random() * pd)
                thr = t + CH[u, hi + 1]
                cand = np.nonzero(CH[u, lo:hi + 1] >= thr)[0]
                a = lo + cand.max()                          # largest v with CH[u][v] >= thr
                for v in range(lo, a):
                    c = x - v
                    if 0 <= c < W:
                        row[c] -= E[u, v]
        out[r] = row
    return out, D


def sweep_stochastic_scheduled(D, H, W, E, A, rng):
    """Twin of sweep_kernel<STOCH=true>: instead of one binomial per (pixel, pass), every pixel carries
    the pass of its next death.  The first death among n molecules needs a per-molecule hazard budget
    Exp(1)/n; by memorylessness the budget is re-drawn (truncated to the pass) when the pass arrives."""
    K = E.shape[0]
    _, CH = tables(E, A)
    D = D.astype(np.int64).copy()
    Hp, Wp = D.shape
    out = np.zeros((H, W))
    nd = np.full((Hp, Wp), 255, dtype=np.int64)

    def clip(x):
        return max(0, x - W + 1), min(K - 1, x)

    def hz(u, lo, hi):
        return CH[u, lo] - CH[u, hi + 1]

    def find_pass(u_from, tau, lo, hi):
        cum = 0.0
        for u in range(u_from, -1, -1):
            cum += hz(u, lo, hi)
            if cum >= tau:
                return u
        return 255

    def enter(y, u_entry):
        for x in range(Wp):
            if D[y, x] > 0:
                lo, hi = clip(x)
                nd[y, x] = find_pass(u_entry, rng.exponential() / D[y, x], lo, hi)

    for y in range(K):
        enter(y, y)
    for r in range(H):
        band = D[r:r + K]
        row = np.array([(E * band[:, c:c + K]).sum() for c in range(W)])
        for u in range(K):
            y = r + u
            for x in np.nonzero(nd[y] == u)[0]:
                lo, hi = clip(x)
                hzp = hz(u, lo, hi)
                n = D[y, x]
                pos = -np.log1p(-rng.random() * -np.expm1(-n * hzp)) / n
                nxt = 255
                while True:
                    cand = np.nonzero(CH[u, lo:hi + 1] >= pos + CH[u, hi + 1])[0]
                    a = lo + (cand.max() if len(cand) else 0)
                    for v in range(lo, a):
                        c = x - v
                        if 0 <= c < W:
                            row[c] -= E[u, v]
                    n -= 1
                    if n == 0:
                        break
                    tau = rng.exponential() / n
                    if tau < hzp - pos:
                        pos += tau
                        continue
                    nxt = find_pass(u - 1, tau - (hzp - pos), lo, hi)
                    break
                D[y, x] = n
                nd[y, x] = nxt
        out[r] = row
        if r + K < Hp:
            enter(r + K, K - 1)
    return out, D


def presample_events(D, H, W, A, rng):
    """Twin of presample_kernel: for every occupied pixel draw Binomial(n, 1-exp(-H_tot)) deaths and place
    each on its (scan row, tap) by inverting the cumulative hazard along the passes (tap rows u from
    min(K-1,y) down to max(0,y-H+1)) and, inside the pass, along the taps hi..lo.  Returns a list of
    (r, y, x, u, a): the molecule of pixel (y,x) dies at tap a of tap row u during scan row r."""
    K = A.shape[0]
    CH = np.zeros((K, K + 1))
    CH[:, :K] = np.cumsum(A[:, ::-1], axis=1)[:, ::-1]
    events = []
    ys, xs = np.nonzero(D)
    for y, x in zip(ys, xs):
        n = int(D[y, x])
        lo, hi = max(0, x - W + 1), min(K - 1, x)
        u_hi, u_lo = min(K - 1, y), max(0, y - (H - 1))
        if lo > hi or u_lo > u_hi:
            continue
        hz = np.array([CH[u, lo] - CH[u, hi + 1] for u in range(u_hi, u_lo - 1, -1)])   # in visiting order
        htot = hz.sum()
        if htot <= 0:
            continue
        dead = rng.binomial(n, -np.expm1(-htot))
        cum = np.cumsum(hz)
        for _ in range(dead):
            t = -np.log1p(-rng.random() * -np.expm1(-htot))
            k = min(int(np.searchsorted(cum, t)), len(hz) - 1)
            u = u_hi - k
            pos = t - (cum[k] - hz[k])
            cand = np.nonzero(CH[u, lo:hi + 1] >= pos + CH[u, hi + 1])[0]
            a = lo + (cand.max() if len(cand) else 0)
            events.append((y - u, y, x, u, a))
    return events


def sweep_stochastic_presampled(D, H, W, E, A, rng):
    """Twin of the stochastic sweep_kernel: dense gather of each scan row against the row-start counts, then
    replay of that row's pre-drawn deaths (count decrement + removal of the taps lo..a-1)."""
    K = E.shape[0]
    D = D.astype(np.int64).copy()
    by_row = [[] for _ in range(H)]
    for ev in presample_events(D, H, W, A, rng):
        by_row[ev[0]].append(ev)
    out = np.zeros((H, W))
    for r in range(H):
        band = D[r:r + K]
        row = np.array([(E * band[:, c:c + K]).sum() for c in range(W)])
        for (_, y, x, u, a) in by_row[r]:
            D[y, x] -= 1
            for v in range(max(0, x - W + 1), a):
                c = x - v
                if 0 <= c < W:
                    row[c] -= E[u, v]
        out[r] = row
    return out, D
