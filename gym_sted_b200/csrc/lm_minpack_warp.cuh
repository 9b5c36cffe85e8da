// Warp-cooperative form of lm_minpack.cuh's lm_fit (MINPACK-1 LMDIF for m = 49, n = 5), device only.
//
// One warp runs one fit.  The three long vectors and the Jacobian live ONCE per warp in shared memory (the one-thread
// form keeps 343 doubles per thread in local memory; run redundantly by 32 lanes that was 86 KB of local memory per fit).
// Work is split so that every floating-point result is the one lm_fit computes, bit for bit:
//   * element-wise passes over the 49 rows (residuals, forward differences, column scaling, Householder updates,
//     copies) are done by the lane that owns the row (rows `lane` and `lane + 32`);
//   * every reduction over rows (ENORM, the dot products of QRFAC and of Q^T f) keeps MINPACK's order: the owners
//     write the 49 terms, then every lane adds them up sequentially, first to last — the same additions in the same
//     order as the one-thread loop.  ENORM's scaled accumulators only matter when a component is tiny or huge; the
//     common case (every component inside (rdwarf, agiant)) reduces to sqrt(sum of squares), which is what the
//     general code computes then (s1 = x3max = 0), and anything else takes the general code on the shared values;
//   * the 5 x 5 part (LMPAR / QRSOLV, norms of 5-vectors, the trust-region logic) runs redundantly on every lane from
//     a packed copy of R, with the one-thread routines themselves.
// All lanes hold bitwise identical scalars, so control flow never diverges and the __syncwarp()s are unconditional.
// tests/test_lm_minpack.py checks that solutions and info codes equal the one-thread device fit bit for bit on both
// models (and through it scipy.optimize.leastsq on the rational model).
#pragma once
// Code size.  With everything inlined and unrolled the Gaussian-fit kernel is ~20 000 SASS instructions (320 KB): warps
// of different fits sit in different parts of it and the kernel stalls on instruction fetch (`no_instruction` is the top
// stall in profiles/r02d_ncu_env256.csv; more resident warps change nothing).  LMW_COMPACT keeps ONE copy of the residual
// evaluation, of ENORM and of QRFAC (noinline, loops over the 5 parameters not unrolled); level 2 also keeps one copy of
// the scalar ENORM, level 3 of LMPAR / QRSOLV, level 4 rolls every loop over the 5 parameters.  The arithmetic - and so
// every result - is the same at every level (same digest of 15 552 fitted parameter sets); measured for those fits of
// 256 images of 256 x 256: 19 952 instructions 4.50 ms, level 1 13 808 / 4.25, level 2 7 392 / 3.99, level 3 6 344 / 3.86,
// level 4 3 840 / 3.52 ms.
#ifndef LMW_COMPACT
#define LMW_COMPACT 4
#endif
#if LMW_COMPACT >= 1
#define LMW_NOINLINE __noinline__
#define LMW_UNROLL1 _Pragma("unroll 1")
#else
#define LMW_NOINLINE inline
#define LMW_UNROLL1
#endif
#include "lm_minpack.cuh"

namespace lm {

constexpr int LMW_LD = 52;
// LMW_LANES: independent reductions of QRFAC (the five column norms it starts with; the dot products of the pivot
// column with every later column) are summed AT THE SAME TIME by different lanes - each still first term to last, so
// every sum is the one MINPACK computes - and broadcast, instead of one after the other by all lanes: 16 dependent
// 49-term chains per iteration of the fit instead of 26.
#ifndef LMW_LANES
#define LMW_LANES 1
#endif
struct LmWarpWork {
    double fvec[LMW_LD], wa4[LMW_LD], tmp[LMW_LD], data[LMW_LD];
    double fjac[LM_N][LMW_LD];
#if LMW_LANES
    double tmpn[LM_N][LMW_LD];       // the terms of up to five concurrent reductions
#endif
};

__device__ __forceinline__ double lmw_seq_sum(const double* v, int from, int to) {
    double s = 0.0;
    for (int i = from; i < to; ++i) s += v[i];
    return s;
}

__device__ LMW_NOINLINE double lmw_enorm(LmWarpWork& w, const double* x, int n) {
    const int lane = threadIdx.x & 31;
    const double rdwarf = 3.834e-20, rgiant = 1.304e19;
    const double agiant = rgiant / (double)n;
    bool mid = true;
    for (int i = lane; i < n; i += 32) {
        const double xabs = fabs(x[i]);
        mid = mid && xabs > rdwarf && xabs < agiant;
        w.tmp[i] = xabs * xabs;
    }
    const bool all_mid = __all_sync(0xffffffffu, mid);
    __syncwarp();
    const double r = all_mid ? sqrt(lmw_seq_sum(w.tmp, 0, n)) : lm_enorm(n, x);
    __syncwarp();                                   // tmp is free again
    return r;
}

__device__ LMW_NOINLINE void lmw_qrfac(LmWarpWork& w, int* ipvt, double* rdiag, double* acnorm, double* wa) {
    const int m = LM_M, n = LM_N, lane = threadIdx.x & 31;
#if LMW_LANES
    {   // the five column norms at once: lane j sums column j (lmw_enorm's common case; anything else falls back to it)
        const double rdwarf = 3.834e-20, agiant = 1.304e19 / (double)m;
        bool mid = true;
        LMW_UNROLL1
        for (int j = 0; j < n; ++j)
            for (int i = lane; i < m; i += 32) {
                const double xabs = fabs(w.fjac[j][i]);
                mid = mid && xabs > rdwarf && xabs < agiant;
                w.tmpn[j][i] = xabs * xabs;
            }
        const bool all_mid = __all_sync(0xffffffffu, mid);
        __syncwarp();
        double mine = 0.0;
        if (all_mid && lane < n) mine = sqrt(lmw_seq_sum(w.tmpn[lane], 0, m));
        LMW_UNROLL1
        for (int j = 0; j < n; ++j) {
            acnorm[j] = all_mid ? __shfl_sync(0xffffffffu, mine, j) : lmw_enorm(w, w.fjac[j], m);
            rdiag[j] = acnorm[j];
            wa[j] = rdiag[j];
            ipvt[j] = j;
        }
        __syncwarp();
    }
#else
    LMW_UNROLL1
    for (int j = 0; j < n; ++j) {
        acnorm[j] = lmw_enorm(w, w.fjac[j], m);
        rdiag[j] = acnorm[j];
        wa[j] = rdiag[j];
        ipvt[j] = j;
    }
#endif
    LMW_UNROLL1
    for (int j = 0; j < n; ++j) {
        int kmax = j;
        for (int k = j; k < n; ++k) if (rdiag[k] > rdiag[kmax]) kmax = k;
        if (kmax != j) {
            for (int i = lane; i < m; i += 32) { const double t = w.fjac[j][i]; w.fjac[j][i] = w.fjac[kmax][i]; w.fjac[kmax][i] = t; }
            __syncwarp();
            rdiag[kmax] = rdiag[j];
            wa[kmax] = wa[j];
            const int k = ipvt[j]; ipvt[j] = ipvt[kmax]; ipvt[kmax] = k;
        }
        double ajnorm = lmw_enorm(w, w.fjac[j] + j, m - j);
        if (ajnorm != 0.0) {
            if (w.fjac[j][j] < 0.0) ajnorm = -ajnorm;
            __syncwarp();
            for (int i = lane; i < m; i += 32) if (i >= j) w.fjac[j][i] /= ajnorm;
            __syncwarp();
            if (lane == 0) w.fjac[j][j] += 1.0;
            __syncwarp();
#if LMW_LANES
            // the later columns are independent of one another: lane k sums the dot product of column k with the pivot
            LMW_UNROLL1
            for (int k = j + 1; k < n; ++k)
                for (int i = lane; i < m; i += 32) if (i >= j) w.tmpn[k][i] = w.fjac[j][i] * w.fjac[k][i];
            __syncwarp();
            double mine = 0.0;
            if (lane > j && lane < n) mine = lmw_seq_sum(w.tmpn[lane], j, m) / w.fjac[j][j];
            __syncwarp();
            LMW_UNROLL1
            for (int k = j + 1; k < n; ++k) {
                const double temp = __shfl_sync(0xffffffffu, mine, k);
                for (int i = lane; i < m; i += 32) if (i >= j) w.fjac[k][i] -= temp * w.fjac[j][i];
            }
            __syncwarp();
#endif
            LMW_UNROLL1
            for (int k = j + 1; k < n; ++k) {
#if !LMW_LANES
                for (int i = lane; i < m; i += 32) if (i >= j) w.tmp[i] = w.fjac[j][i] * w.fjac[k][i];
                __syncwarp();
                const double sum = lmw_seq_sum(w.tmp, j, m);
                const double temp = sum / w.fjac[j][j];
                __syncwarp();
                for (int i = lane; i < m; i += 32) if (i >= j) w.fjac[k][i] -= temp * w.fjac[j][i];
                __syncwarp();
#endif
                if (rdiag[k] != 0.0) {
                    double t = w.fjac[k][j] / rdiag[k];
                    rdiag[k] *= sqrt(fmax(0.0, 1.0 - t * t));
                    t = rdiag[k] / wa[k];
                    if (0.05 * (t * t) <= LM_EPS) {
                        rdiag[k] = lmw_enorm(w, w.fjac[k] + j + 1, m - j - 1);
                        wa[k] = rdiag[k];
                    }
                }
            }
        }
        rdiag[j] = -ajnorm;
    }
}

// residuals(p, f): every lane calls it with the same p; on return f[0..48] (shared memory) holds the residuals.  It must
// start with a __syncwarp() (earlier readers of f are done) and end with one (the values are visible to every lane).
template <class WarpResiduals>
__device__ inline int lm_fit_warp(const WarpResiduals& residuals, LmWarpWork& w, double* x, double ftol, double xtol, double gtol,
                                  int maxfev, double factor) {
    const int m = LM_M, n = LM_N, lane = threadIdx.x & 31;
    double diag[LM_N], qtf[LM_N], wa1[LM_N], wa2[LM_N], wa3[LM_N], wa5[LM_N], r[LM_N * LM_N];
    int ipvt[LM_N];
    LM_U1
    for (int i = 0; i < n * n; ++i) r[i] = 0.0;
    residuals(x, w.fvec);
    int nfev = 1, info = 0, iter = 1;
    double fnorm = lmw_enorm(w, w.fvec, m), par = 0.0, delta = 0.0, xnorm = 0.0, gnorm = 0.0;
    const double eps = sqrt(LM_EPS);
    while (info == 0) {
        LMW_UNROLL1
        LM_U1
        for (int j = 0; j < n; ++j) {
            const double temp = x[j];
            double h = eps * fabs(temp);
            if (h == 0.0) h = eps;
            x[j] = temp + h;
            residuals(x, w.wa4);
            x[j] = temp;
            for (int i = lane; i < m; i += 32) w.fjac[j][i] = (w.wa4[i] - w.fvec[i]) / h;
        }
        __syncwarp();
        nfev += n;
        lmw_qrfac(w, ipvt, wa1, wa2, wa3);
        if (iter == 1) {
            LM_U1
            for (int j = 0; j < n; ++j) { diag[j] = wa2[j]; if (wa2[j] == 0.0) diag[j] = 1.0; }
            LM_U1
            for (int j = 0; j < n; ++j) wa3[j] = diag[j] * x[j];
            xnorm = lm_enorm(n, wa3);
            delta = factor * xnorm;
            if (delta == 0.0) delta = factor;
        }
        for (int i = lane; i < m; i += 32) w.wa4[i] = w.fvec[i];
        __syncwarp();
        LMW_UNROLL1
        LM_U1
        for (int j = 0; j < n; ++j) {
            const double ajj = w.fjac[j][j];
            if (ajj != 0.0) {
                for (int i = lane; i < m; i += 32) if (i >= j) w.tmp[i] = w.fjac[j][i] * w.wa4[i];
                __syncwarp();
                const double sum = lmw_seq_sum(w.tmp, j, m);
                const double temp = -sum / ajj;
                __syncwarp();
                for (int i = lane; i < m; i += 32) if (i >= j) w.wa4[i] += w.fjac[j][i] * temp;
                __syncwarp();
            }
            LM_U1
            for (int i = 0; i < j; ++i) r[i + n * j] = w.fjac[j][i];      // R, packed: column j above the diagonal ...
            r[j + n * j] = wa1[j];                                        // ... and its diagonal (rdiag)
            qtf[j] = w.wa4[j];
        }
        gnorm = 0.0;
        if (fnorm != 0.0) {
            double qn[LM_N];                                  // qtf[i] / fnorm: the same five quotients MINPACK forms 15 times
            LM_U1
            for (int i = 0; i < n; ++i) qn[i] = qtf[i] / fnorm;
            LM_U1
            for (int j = 0; j < n; ++j) {
                const int l = ipvt[j];
                if (wa2[l] != 0.0) {
                    double sum = 0.0;
                    LM_U1
                    for (int i = 0; i <= j; ++i) sum += r[i + n * j] * qn[i];
                    gnorm = fmax(gnorm, fabs(sum / wa2[l]));
                }
            }
        }
        if (gnorm <= gtol) { info = 4; break; }
        LM_U1
        for (int j = 0; j < n; ++j) diag[j] = fmax(diag[j], wa2[j]);
        double ratio = 0.0;
        do {
            lm_lmpar<LM_N>(r, ipvt, diag, qtf, delta, &par, wa1, wa2, wa3, wa5);
            LM_U1
            for (int j = 0; j < n; ++j) { wa1[j] = -wa1[j]; wa2[j] = x[j] + wa1[j]; wa3[j] = diag[j] * wa1[j]; }
            const double pnorm = lm_enorm(n, wa3);
            if (iter == 1) delta = fmin(delta, pnorm);
            residuals(wa2, w.wa4);
            ++nfev;
            const double fnorm1 = lmw_enorm(w, w.wa4, m);
            double actred = -1.0;
            if (0.1 * fnorm1 < fnorm) { const double t = fnorm1 / fnorm; actred = 1.0 - t * t; }
            LM_U1
            for (int j = 0; j < n; ++j) {
                wa3[j] = 0.0;
                const double temp = wa1[ipvt[j]];
                LM_U1
                for (int i = 0; i <= j; ++i) wa3[i] += r[i + n * j] * temp;
            }
            const double temp1 = lm_enorm(n, wa3) / fnorm, temp2 = (sqrt(par) * pnorm) / fnorm;
            const double prered = temp1 * temp1 + temp2 * temp2 / 0.5;
            const double dirder = -(temp1 * temp1 + temp2 * temp2);
            ratio = 0.0;
            if (prered != 0.0) ratio = actred / prered;
            if (ratio <= 0.25) {
                double temp = 0.5;
                if (actred < 0.0) temp = 0.5 * dirder / (dirder + 0.5 * actred);
                if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
                delta = temp * fmin(delta, pnorm / 0.1);
                par /= temp;
            } else if (par == 0.0 || ratio >= 0.75) {
                delta = pnorm / 0.5;
                par *= 0.5;
            }
            if (ratio >= 1e-4) {
                LM_U1
                for (int j = 0; j < n; ++j) { x[j] = wa2[j]; wa2[j] = diag[j] * x[j]; }
                for (int i = lane; i < m; i += 32) w.fvec[i] = w.wa4[i];
                __syncwarp();
                xnorm = lm_enorm(n, wa2);
                fnorm = fnorm1;
                ++iter;
            }
            const bool small = fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0;
            if (small) info = 1;
            if (delta <= xtol * xnorm) info = 2;
            if (small && info == 2) info = 3;
            if (info != 0) break;
            if (nfev >= maxfev) info = 5;
            if (fabs(actred) <= LM_EPS && prered <= LM_EPS && 0.5 * ratio <= 1.0) info = 6;
            if (delta <= LM_EPS * xnorm) info = 7;
            if (gnorm <= LM_EPS) info = 8;
            if (info != 0) break;
        } while (ratio < 1e-4);
    }
    return info;
}

}  // namespace lm
