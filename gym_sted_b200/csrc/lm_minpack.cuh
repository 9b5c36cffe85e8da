// Levenberg-Marquardt least squares for m = 49 residuals and n = 5 parameters: MINPACK-1's LMDIF / FDJAC2 / QRFAC /
// LMPAR / QRSOLV / ENORM (More, Garbow, Hillstrom) restated, one fit per thread, float64.  This is the solver behind
// scipy.optimize.leastsq, which the reference's Gaussian-fit validation calls (gym_sted/rewards/utils.py:6-49).
// Host-compilable (tests/lm_harness.cpp builds it with g++ to pin it bit-for-bit against scipy on a model without
// transcendentals); the device build of the including file must use -fmad=false so both execute the same IEEE operations.
#pragma once
#include <math.h>
#ifdef __CUDACC__
#define LM_HD __host__ __device__
#else
#define LM_HD
#endif

// see lm_minpack_warp.cuh: one copy of the scalar helpers in device code at the higher LMW_COMPACT levels
#if defined(__CUDA_ARCH__) && defined(LMW_COMPACT) && LMW_COMPACT >= 2
#define LM_INL2 __noinline__
#else
#define LM_INL2 inline
#endif
#if defined(__CUDA_ARCH__) && defined(LMW_COMPACT) && LMW_COMPACT >= 3
#define LM_INL3 __noinline__
#else
#define LM_INL3 inline
#endif
#if defined(__CUDA_ARCH__) && defined(LMW_COMPACT) && LMW_COMPACT >= 4      // rolled loops over the 5 parameters
#define LM_U1 _Pragma("unroll 1")
#else
#define LM_U1
#endif

namespace lm {

constexpr int LM_M = 49, LM_N = 5;
constexpr double LM_EPS = 2.220446049250313e-16, LM_DWARF = 2.2250738585072014e-308;

LM_HD LM_INL2 double lm_enorm(int n, const double* x) {
    const double rdwarf = 3.834e-20, rgiant = 1.304e19;
    double s1 = 0, s2 = 0, s3 = 0, x1max = 0, x3max = 0;
    const double agiant = rgiant / (double)n;
    {   // every component in (rdwarf, agiant) - the case of all but degenerate fits: the general code below then ends in
        // sqrt(s2 * (1.0 + (0 / s2) * (0 * 0))) = sqrt(s2) with the same s2 (same additions, same order), so this
        // shortcut returns the identical value without the scaled accumulators' branches and the division
        double q = 0.0;
        bool mid = n > 0;
        for (int i = 0; i < n; ++i) {
            const double xabs = fabs(x[i]);
            mid = mid && xabs > rdwarf && xabs < agiant;
            q += xabs * xabs;
        }
        if (mid) return sqrt(q);
    }
    LM_U1
    for (int i = 0; i < n; ++i) {
        const double xabs = fabs(x[i]);
        if (xabs > rdwarf && xabs < agiant) {
            s2 += xabs * xabs;
        } else if (xabs <= rdwarf) {
            if (xabs > x3max) { const double t = x3max / xabs; s3 = 1.0 + s3 * (t * t); x3max = xabs; }
            else if (xabs != 0.0) { const double t = xabs / x3max; s3 += t * t; }
        } else {
            if (xabs > x1max) { const double t = x1max / xabs; s1 = 1.0 + s1 * (t * t); x1max = xabs; }
            else { const double t = xabs / x1max; s1 += t * t; }
        }
    }
    if (s1 != 0.0) return x1max * sqrt(s1 + (s2 / x1max) / x1max);
    if (s2 != 0.0) {
        if (s2 >= x3max) return sqrt(s2 * (1.0 + (x3max / s2) * (x3max * s3)));
        return sqrt(x3max * ((s2 / x3max) + (x3max * s3)));
    }
    return x3max * sqrt(s3);
}

// The reference's model: amplitude*exp(-((cx-X)^2/(2 wx^2) + (cy-Y)^2/(2 wy^2))) - data on a 7x7 window, X = row-3, Y = col-3,
// widths offset by 1e-3 (gym_sted/rewards/utils.py:6-16).
struct GaussResiduals {
    const double* data;
    LM_HD void operator()(const double* p, double* f) const {
        const double wx = p[3] + 1e-3, wy = p[4] + 1e-3;
        const double dx2 = 2.0 * (wx * wx), dy2 = 2.0 * (wy * wy);
        LM_U1
        for (int i = 0; i < LM_M; ++i) {
            const double X = (double)(i / 7) - 3.0, Y = (double)(i % 7) - 3.0;
            const double a = (p[1] - X) * (p[1] - X) / dx2, b = (p[2] - Y) * (p[2] - Y) / dy2;
            f[i] = p[0] * exp(-1.0 * (a + b)) - data[i];
        }
    }
};

// Test model without transcendentals (every operation is IEEE-exact on CPU and GPU): a Lorentzian-like peak.
struct RationalResiduals {
    const double* data;
    LM_HD void operator()(const double* p, double* f) const {
        const double wx = p[3] + 1e-3, wy = p[4] + 1e-3;
        const double dx2 = 2.0 * (wx * wx), dy2 = 2.0 * (wy * wy);
        LM_U1
        for (int i = 0; i < LM_M; ++i) {
            const double X = (double)(i / 7) - 3.0, Y = (double)(i % 7) - 3.0;
            const double a = (p[1] - X) * (p[1] - X) / dx2, b = (p[2] - Y) * (p[2] - Y) / dy2;
            f[i] = p[0] / (1.0 + (a + b)) - data[i];
        }
    }
};

// LD: leading dimension of r (LM_M inside lm_fit, where r is the top of the Jacobian's storage; LM_N for a packed copy)
template <int LD = LM_M>
LM_HD LM_INL3 void lm_qrsolv(double* r /*[LD*LM_N], column major*/, const int* ipvt, const double* diag, const double* qtb,
                          double* x, double* sdiag, double* wa) {
    const int n = LM_N, ld = LD;
    LM_U1
    for (int j = 0; j < n; ++j) {
        LM_U1
        for (int i = j; i < n; ++i) r[i + ld * j] = r[j + ld * i];
        x[j] = r[j + ld * j];
        wa[j] = qtb[j];
    }
    LM_U1
    for (int j = 0; j < n; ++j) {
        const int l = ipvt[j];
        if (diag[l] != 0.0) {
            LM_U1
            for (int k = j; k < n; ++k) sdiag[k] = 0.0;
            sdiag[j] = diag[l];
            double qtbpj = 0.0;
            LM_U1
            for (int k = j; k < n; ++k) {
                if (sdiag[k] == 0.0) continue;
                double c, s;
                if (fabs(r[k + ld * k]) < fabs(sdiag[k])) {
                    const double cotan = r[k + ld * k] / sdiag[k];
                    s = 0.5 / sqrt(0.25 + 0.25 * (cotan * cotan));
                    c = s * cotan;
                } else {
                    const double tn = sdiag[k] / r[k + ld * k];
                    c = 0.5 / sqrt(0.25 + 0.25 * (tn * tn));
                    s = c * tn;
                }
                r[k + ld * k] = c * r[k + ld * k] + s * sdiag[k];
                const double temp = c * wa[k] + s * qtbpj;
                qtbpj = -s * wa[k] + c * qtbpj;
                wa[k] = temp;
                LM_U1
                for (int i = k + 1; i < n; ++i) {
                    const double t2 = c * r[i + ld * k] + s * sdiag[i];
                    sdiag[i] = -s * r[i + ld * k] + c * sdiag[i];
                    r[i + ld * k] = t2;
                }
            }
        }
        sdiag[j] = r[j + ld * j];
        r[j + ld * j] = x[j];
    }
    int nsing = n;
    LM_U1
    for (int j = 0; j < n; ++j) {
        if (sdiag[j] == 0.0 && nsing == n) nsing = j;
        if (nsing < n) wa[j] = 0.0;
    }
    LM_U1
    for (int k = 0; k < nsing; ++k) {
        const int j = nsing - k - 1;
        double sum = 0.0;
        LM_U1
        for (int i = j + 1; i < nsing; ++i) sum += r[i + ld * j] * wa[i];
        wa[j] = (wa[j] - sum) / sdiag[j];
    }
    LM_U1
    for (int j = 0; j < n; ++j) x[ipvt[j]] = wa[j];
}

template <int LD = LM_M>
LM_HD LM_INL3 void lm_lmpar(double* r, const int* ipvt, const double* diag, const double* qtb, double delta, double* par, double* x,
                         double* sdiag, double* wa1, double* wa2) {
    const int n = LM_N, ld = LD;
    int nsing = n;
    LM_U1
    for (int j = 0; j < n; ++j) {
        wa1[j] = qtb[j];
        if (r[j + ld * j] == 0.0 && nsing == n) nsing = j;
        if (nsing < n) wa1[j] = 0.0;
    }
    LM_U1
    for (int k = 0; k < nsing; ++k) {
        const int j = nsing - k - 1;
        wa1[j] /= r[j + ld * j];
        const double temp = wa1[j];
        LM_U1
        for (int i = 0; i < j; ++i) wa1[i] -= r[i + ld * j] * temp;
    }
    LM_U1
    for (int j = 0; j < n; ++j) x[ipvt[j]] = wa1[j];
    int iter = 0;
    LM_U1
    for (int j = 0; j < n; ++j) wa2[j] = diag[j] * x[j];
    double dxnorm = lm_enorm(n, wa2);
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) { *par = 0.0; return; }
    double parl = 0.0;
    if (nsing >= n) {
        LM_U1
        for (int j = 0; j < n; ++j) { const int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
        LM_U1
        for (int j = 0; j < n; ++j) {
            double sum = 0.0;
            LM_U1
            for (int i = 0; i < j; ++i) sum += r[i + ld * j] * wa1[i];
            wa1[j] = (wa1[j] - sum) / r[j + ld * j];
        }
        const double temp = lm_enorm(n, wa1);
        parl = ((fp / delta) / temp) / temp;
    }
    LM_U1
    for (int j = 0; j < n; ++j) {
        double sum = 0.0;
        LM_U1
        for (int i = 0; i <= j; ++i) sum += r[i + ld * j] * qtb[i];
        wa1[j] = sum / diag[ipvt[j]];
    }
    const double gnorm = lm_enorm(n, wa1);
    double paru = gnorm / delta;
    if (paru == 0.0) paru = LM_DWARF / fmin(delta, 0.1);
    *par = fmax(*par, parl);
    *par = fmin(*par, paru);
    if (*par == 0.0) *par = gnorm / dxnorm;
    for (;;) {
        ++iter;
        if (*par == 0.0) *par = fmax(LM_DWARF, 0.001 * paru);
        double temp = sqrt(*par);
        LM_U1
        for (int j = 0; j < n; ++j) wa1[j] = temp * diag[j];
        lm_qrsolv<LD>(r, ipvt, wa1, qtb, x, sdiag, wa2);
        LM_U1
        for (int j = 0; j < n; ++j) wa2[j] = diag[j] * x[j];
        dxnorm = lm_enorm(n, wa2);
        temp = fp;
        fp = dxnorm - delta;
        if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
        LM_U1
        for (int j = 0; j < n; ++j) { const int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
        LM_U1
        for (int j = 0; j < n; ++j) {
            wa1[j] /= sdiag[j];
            const double t2 = wa1[j];
            LM_U1
            for (int i = j + 1; i < n; ++i) wa1[i] -= r[i + ld * j] * t2;
        }
        temp = lm_enorm(n, wa1);
        const double parc = ((fp / delta) / temp) / temp;
        if (fp > 0.0) parl = fmax(parl, *par);
        if (fp < 0.0) paru = fmin(paru, *par);
        *par = fmax(parl, *par + parc);
    }
    if (iter == 0) *par = 0.0;
}

LM_HD inline void lm_qrfac(double* a /*[LM_M*LM_N]*/, int* ipvt, double* rdiag, double* acnorm, double* wa) {
    const int m = LM_M, n = LM_N;
    for (int j = 0; j < n; ++j) {
        acnorm[j] = lm_enorm(m, a + m * j);
        rdiag[j] = acnorm[j];
        wa[j] = rdiag[j];
        ipvt[j] = j;
    }
    for (int j = 0; j < n; ++j) {
        int kmax = j;
        for (int k = j; k < n; ++k) if (rdiag[k] > rdiag[kmax]) kmax = k;
        if (kmax != j) {
            for (int i = 0; i < m; ++i) { const double t = a[i + m * j]; a[i + m * j] = a[i + m * kmax]; a[i + m * kmax] = t; }
            rdiag[kmax] = rdiag[j];
            wa[kmax] = wa[j];
            const int k = ipvt[j]; ipvt[j] = ipvt[kmax]; ipvt[kmax] = k;
        }
        double ajnorm = lm_enorm(m - j, a + j + m * j);
        if (ajnorm != 0.0) {
            if (a[j + m * j] < 0.0) ajnorm = -ajnorm;
            for (int i = j; i < m; ++i) a[i + m * j] /= ajnorm;
            a[j + m * j] += 1.0;
            for (int k = j + 1; k < n; ++k) {
                double sum = 0.0;
                for (int i = j; i < m; ++i) sum += a[i + m * j] * a[i + m * k];
                const double temp = sum / a[j + m * j];
                for (int i = j; i < m; ++i) a[i + m * k] -= temp * a[i + m * j];
                if (rdiag[k] != 0.0) {
                    double t = a[j + m * k] / rdiag[k];
                    rdiag[k] *= sqrt(fmax(0.0, 1.0 - t * t));
                    t = rdiag[k] / wa[k];
                    if (0.05 * (t * t) <= LM_EPS) {
                        rdiag[k] = lm_enorm(m - j - 1, a + j + 1 + m * k);
                        wa[k] = rdiag[k];
                    }
                }
            }
        }
        rdiag[j] = -ajnorm;
    }
}

// LMDIF for the 7x7 Gaussian model; x holds the start on entry and the solution on return.
template <class Residuals>
LM_HD inline int lm_fit(const Residuals& residuals, double* x, double ftol, double xtol, double gtol, int maxfev, double factor) {
    const int m = LM_M, n = LM_N;
    double fvec[LM_M], wa4[LM_M], fjac[LM_M * LM_N];
    double diag[LM_N], qtf[LM_N], wa1[LM_N], wa2[LM_N], wa3[LM_N];
    int ipvt[LM_N];
    residuals(x, fvec);
    int nfev = 1, info = 0, iter = 1;
    double fnorm = lm_enorm(m, fvec), par = 0.0, delta = 0.0, xnorm = 0.0, gnorm = 0.0;
    const double eps = sqrt(LM_EPS);
    while (info == 0) {
        // fdjac2: forward differences
        for (int j = 0; j < n; ++j) {
            const double temp = x[j];
            double h = eps * fabs(temp);
            if (h == 0.0) h = eps;
            x[j] = temp + h;
            residuals(x, wa4);
            x[j] = temp;
            for (int i = 0; i < m; ++i) fjac[i + m * j] = (wa4[i] - fvec[i]) / h;
        }
        nfev += n;
        lm_qrfac(fjac, ipvt, wa1, wa2, wa3);
        if (iter == 1) {
            for (int j = 0; j < n; ++j) { diag[j] = wa2[j]; if (wa2[j] == 0.0) diag[j] = 1.0; }
            for (int j = 0; j < n; ++j) wa3[j] = diag[j] * x[j];
            xnorm = lm_enorm(n, wa3);
            delta = factor * xnorm;
            if (delta == 0.0) delta = factor;
        }
        for (int i = 0; i < m; ++i) wa4[i] = fvec[i];
        for (int j = 0; j < n; ++j) {
            if (fjac[j + m * j] != 0.0) {
                double sum = 0.0;
                for (int i = j; i < m; ++i) sum += fjac[i + m * j] * wa4[i];
                const double temp = -sum / fjac[j + m * j];
                for (int i = j; i < m; ++i) wa4[i] += fjac[i + m * j] * temp;
            }
            fjac[j + m * j] = wa1[j];
            qtf[j] = wa4[j];
        }
        gnorm = 0.0;
        if (fnorm != 0.0) {
            for (int j = 0; j < n; ++j) {
                const int l = ipvt[j];
                if (wa2[l] != 0.0) {
                    double sum = 0.0;
                    for (int i = 0; i <= j; ++i) sum += fjac[i + m * j] * (qtf[i] / fnorm);
                    gnorm = fmax(gnorm, fabs(sum / wa2[l]));
                }
            }
        }
        if (gnorm <= gtol) { info = 4; break; }
        for (int j = 0; j < n; ++j) diag[j] = fmax(diag[j], wa2[j]);
        double ratio = 0.0;
        do {
            lm_lmpar<LM_M>(fjac, ipvt, diag, qtf, delta, &par, wa1, wa2, wa3, wa4);
            for (int j = 0; j < n; ++j) { wa1[j] = -wa1[j]; wa2[j] = x[j] + wa1[j]; wa3[j] = diag[j] * wa1[j]; }
            const double pnorm = lm_enorm(n, wa3);
            if (iter == 1) delta = fmin(delta, pnorm);
            residuals(wa2, wa4);
            ++nfev;
            const double fnorm1 = lm_enorm(m, wa4);
            double actred = -1.0;
            if (0.1 * fnorm1 < fnorm) { const double t = fnorm1 / fnorm; actred = 1.0 - t * t; }
            for (int j = 0; j < n; ++j) {
                wa3[j] = 0.0;
                const double temp = wa1[ipvt[j]];
                for (int i = 0; i <= j; ++i) wa3[i] += fjac[i + m * j] * temp;
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
                for (int j = 0; j < n; ++j) { x[j] = wa2[j]; wa2[j] = diag[j] * x[j]; }
                for (int i = 0; i < m; ++i) fvec[i] = wa4[i];
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
