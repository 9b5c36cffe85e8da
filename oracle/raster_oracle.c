/*
 * raster_oracle.c — TEST INFRASTRUCTURE ONLY (never imported by gym_sted_b200/).
 *
 * CPU restatement, in the reference's own order and in float64, of the raster-scan loop inside
 * pysted.base.Microscope.get_signal_and_bleach that gym-sted calls at
 * gym_sted/envs/mosted_env.py:184,222,227,232.  pySTED itself is an un-vendored, unpinned
 * dependency of the reference (setup.py:13) and is absent from /root/reference and from this
 * image, so this follows the published algorithm as recalled in SURVEY.md Appendix A.3:
 *
 *   for every dwell (r, c) of the region of interest, row-major:
 *       I[r][c] = sum_{u,v} E[u][v] * D[r+u][c+v]                 (pysted raster loop)
 *       if bleach:  every pixel of the window is thinned with survival
 *                   S[u][v] = exp(-k_b[u][v] * pdt)
 *                   (default_update_survival_probabilities + sample_molecules)
 *
 * PARITY UNPINNED: the reference's tests hold no golden vector for this path;
 * the only anchors are the fitted parameter sets of gym_sted/routines/routines.json (checked in
 * tests/test_oracle_anchors.py).
 *
 * mode 0: no bleaching; mode 1: expected counts (D *= S, the mean field of the binomial chain);
 * mode 2: binomial thinning molecule by molecule, one uniform per molecule per dwell, like
 * sample_molecules' `rand() <= prob` loop (xorshift128+ here instead of libc rand()).
 *
 * The same loop, timed, is bench.py's cpu_baseline ("port").
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t s0, s1; } rng_t;

static inline uint64_t rng_next(rng_t* g) {
    uint64_t x = g->s0, y = g->s1;
    g->s0 = y;
    x ^= x << 23;
    g->s1 = x ^ y ^ (x >> 17) ^ (y >> 26);
    return g->s1 + y;
}

static inline double rng_uniform(rng_t* g) { return (double)(rng_next(g) >> 11) * (1.0 / 9007199254740992.0); }

static void rng_seed(rng_t* g, uint64_t seed) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; g->s0 = z ^ (z >> 31);
    z = seed + 2 * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; g->s1 = z ^ (z >> 31);
    if (!g->s0 && !g->s1) g->s1 = 1;
}

/*
 * D       : [Hp][Wp] float64 molecule map, Hp = H+K-1, Wp = W+K-1, updated IN PLACE when mode > 0
 * E, S    : [K][K] effective PSF and per-dwell survival probability
 * out     : [H][W] detected power per pixel
 */
void oracle_raster(double* D, int H, int W, int K, const double* E, const double* S, int mode,
                   uint64_t seed, double* out) {
    const int Wp = W + K - 1;
    rng_t g;
    rng_seed(&g, seed);
    for (int r = 0; r < H; ++r) {
        for (int c = 0; c < W; ++c) {
            double acc = 0.0;
            for (int u = 0; u < K; ++u) {
                const double* d = D + (size_t)(r + u) * Wp + c;
                const double* e = E + (size_t)u * K;
                for (int v = 0; v < K; ++v) acc += e[v] * d[v];
            }
            out[(size_t)r * W + c] = acc;
            if (mode == 1) {
                for (int u = 0; u < K; ++u) {
                    double* d = D + (size_t)(r + u) * Wp + c;
                    const double* s = S + (size_t)u * K;
                    for (int v = 0; v < K; ++v) d[v] *= s[v];
                }
            } else if (mode == 2) {
                for (int u = 0; u < K; ++u) {
                    double* d = D + (size_t)(r + u) * Wp + c;
                    const double* s = S + (size_t)u * K;
                    for (int v = 0; v < K; ++v) {
                        const long n = (long)d[v];
                        if (n <= 0) continue;
                        long kept = 0;
                        for (long m = 0; m < n; ++m) kept += (rng_uniform(&g) <= s[v]);
                        d[v] = (double)kept;
                    }
                }
            }
        }
    }
}

/*
 * Per-dwell imaging parameters (pySTED accepts (H, W) arrays for pdt, p_ex and p_sted; the timed environments use it,
 * gym_sted/envs/timed_sted_env.py:141-148): dwell (r, c) uses the idx[r][c]-th effective PSF / survival table, i.e.
 * the same loop as oracle_raster with E and S looked up per dwell.  The caller builds one table per distinct
 * (pdt, p_ex, p_sted) triple directly from the physics — no linearity assumption is made here.
 */
void oracle_raster_tables(double* D, int H, int W, int K, const double* E, const double* S, const int32_t* idx,
                          int mode, uint64_t seed, double* out) {
    const int Wp = W + K - 1;
    const size_t t = (size_t)K * K;
    rng_t g;
    rng_seed(&g, seed);
    for (int r = 0; r < H; ++r) {
        for (int c = 0; c < W; ++c) {
            const double* Ej = E + t * (size_t)idx[(size_t)r * W + c];
            const double* Sj = S + t * (size_t)idx[(size_t)r * W + c];
            double acc = 0.0;
            for (int u = 0; u < K; ++u) {
                const double* d = D + (size_t)(r + u) * Wp + c;
                for (int v = 0; v < K; ++v) acc += Ej[u * K + v] * d[v];
            }
            out[(size_t)r * W + c] = acc;
            if (mode == 0) continue;
            for (int u = 0; u < K; ++u) {
                double* d = D + (size_t)(r + u) * Wp + c;
                for (int v = 0; v < K; ++v) {
                    const double s = Sj[u * K + v];
                    if (mode == 1) { d[v] *= s; continue; }
                    const long n = (long)d[v];
                    if (n <= 0) continue;
                    long kept = 0;
                    for (long m = 0; m < n; ++m) kept += (rng_uniform(&g) <= s);
                    d[v] = (double)kept;
                }
            }
        }
    }
}

/* Batched driver used by the CPU baseline: B independent maps, one after the other on this thread. */
void oracle_raster_batch(double* D, int B, int H, int W, int K, const double* E, const double* S, int mode,
                         uint64_t seed, double* out) {
    const size_t dm = (size_t)(H + K - 1) * (W + K - 1), im = (size_t)H * W, t = (size_t)K * K;
    for (int b = 0; b < B; ++b)
        oracle_raster(D + b * dm, H, W, K, E + b * t, S + b * t, mode, seed + (uint64_t)b, out + b * im);
}
