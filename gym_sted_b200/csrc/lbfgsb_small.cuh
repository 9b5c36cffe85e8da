// L-BFGS-B for one or two bounded variables, as scipy runs it for
//     scipy.optimize.minimize(fun, x0, bounds=..., options={"eps": eps, "maxiter": maxiter}, tol=tol)
// i.e. method L-BFGS-B, forward-difference gradients with the ABSOLUTE step `eps`, ftol = gtol = tol, 10 correction
// pairs, More'-Thuente line search (ftol 1e-3, gtol 0.9, xtol 0.1, at most 20 evaluations).  That call is how the
// reference fits fluorophore parameters (gym_sted/utils.py:574-594); scipy's implementation is a translation of
// L-BFGS-B 3.0 (Byrd, Lu, Nocedal, Zhu 1995; Morales, Nocedal 2011) with MINPACK-2's dcsrch/dcstep.  This header
// restates that algorithm for N <= 2, where the limited-memory matrix can be held densely:
//   * generalized Cauchy point along the projected steepest-descent path of the quadratic model,
//   * subspace minimisation over the variables free at that point, with the 3.0 projection / backtracking step,
//   * direction d = x_hat - x, first step min(1/|d|, 1) (bounded, not boxed) then 1, capped by the feasible step,
//   * dcsrch line search, stop on projected gradient <= tol or (f_old - f) <= tol * max(|f_old|, |f|, 1),
//   * BFGS pair (s, y) kept when s'y > eps_mach * |g_old'd| * step.
// Compiles for host and device; tests/test_lbfgsb_small.py pins the host build against scipy iterate by iterate
// (final x equal to ~1e-15).  The objective is called with all threads of a CTA holding the same x (it may reduce
// cooperatively); every scalar of the state machine is thread-uniform.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define LB_HD __host__ __device__
#else
#define LB_HD
#endif

namespace lbfgsb_small {

constexpr double kEpsMach = 2.220446049250313e-16;
constexpr int kPairs = 10;

LB_HD inline double lb_max3(double a, double b, double c) { return fmax(fmax(a, b), c); }

// one safeguarded interpolation step of the More'-Thuente search (MINPACK-2 dcstep)
LB_HD inline void dcstep(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double& stp, double fp,
                         double dp, bool& brackt, double stpmin, double stpmax) {
    const double sgnd = dp * (dx / fabs(dx));
    double stpf;
    if (fp > fx) {
        const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        const double s = lb_max3(fabs(theta), fabs(dx), fabs(dp));
        double gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp < stx) gamma = -gamma;
        const double p = (gamma - dx) + theta, q = ((gamma - dx) + gamma) + dp, r = p / q;
        const double stpc = stx + r * (stp - stx);
        const double stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
        stpf = (fabs(stpc - stx) < fabs(stpq - stx)) ? stpc : stpc + (stpq - stpc) / 2.0;
        brackt = true;
    } else if (sgnd < 0.0) {
        const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        const double s = lb_max3(fabs(theta), fabs(dx), fabs(dp));
        double gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp > stx) gamma = -gamma;
        const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dx, r = p / q;
        const double stpc = stp + r * (stx - stp);
        const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
        stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
        brackt = true;
    } else if (fabs(dp) < fabs(dx)) {
        const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        const double s = lb_max3(fabs(theta), fabs(dx), fabs(dp));
        double gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
        if (stp > stx) gamma = -gamma;
        const double p = (gamma - dp) + theta, q = (gamma + (dx - dp)) + gamma, r = p / q;
        double stpc;
        if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
        else if (stp > stx) stpc = stpmax;
        else stpc = stpmin;
        const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (brackt) {
            stpf = (fabs(stpc - stp) < fabs(stpq - stp)) ? stpc : stpq;
            if (stp > stx) stpf = fmin(stp + 0.66 * (sty - stp), stpf);
            else stpf = fmax(stp + 0.66 * (sty - stp), stpf);
        } else {
            stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
            stpf = fmin(stpmax, stpf);
            stpf = fmax(stpmin, stpf);
        }
    } else {
        if (brackt) {
            const double theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
            const double s = lb_max3(fabs(theta), fabs(dy), fabs(dp));
            double gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
            if (stp > sty) gamma = -gamma;
            const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dy, r = p / q;
            stpf = stp + r * (sty - stp);
        } else if (stp > stx) {
            stpf = stpmax;
        } else {
            stpf = stpmin;
        }
    }
    if (fp > fx) {
        sty = stp; fy = fp; dy = dp;
    } else {
        if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
        stx = stp; fx = fp; dx = dp;
    }
    stp = stpf;
}

// MINPACK-2 dcsrch as a resumable state machine: construct with the start point, then feed (f, g) at `stp`
struct LineSearch {
    double ftol, gtol, xtol, stpmin, stpmax;
    bool brackt;
    int stage;
    double finit, ginit, gtest, width, width1, stx, fx, gx, sty, fy, gy, stmin, stmax, stp;

    LB_HD LineSearch(double stp0, double f, double g, double ftol_, double gtol_, double xtol_, double stpmin_, double stpmax_)
        : ftol(ftol_), gtol(gtol_), xtol(xtol_), stpmin(stpmin_), stpmax(stpmax_), brackt(false), stage(1), finit(f), ginit(g),
          gtest(ftol_ * g), width(stpmax_ - stpmin_), width1((stpmax_ - stpmin_) / 0.5), stx(0.0), fx(f), gx(g), sty(0.0), fy(f),
          gy(g), stmin(0.0), stmax(stp0 + 4.0 * stp0), stp(stp0) {}

    // true when the search is over (stp is the accepted step)
    LB_HD bool step(double f, double g) {
        const double ftest = finit + stp * gtest;
        if (stage == 1 && f <= ftest && g >= 0.0) stage = 2;
        bool done = false;
        if (brackt && (stp <= stmin || stp >= stmax)) done = true;
        if (brackt && stmax - stmin <= xtol * stmax) done = true;
        if (stp == stpmax && f <= ftest && g <= gtest) done = true;
        if (stp == stpmin && (f > ftest || g >= gtest)) done = true;
        if (f <= ftest && fabs(g) <= gtol * (-ginit)) done = true;
        if (done) return true;
        if (stage == 1 && f <= fx && f > ftest) {
            const double fm = f - stp * gtest, gm = g - gtest;
            double fxm = fx - stx * gtest, fym = fy - sty * gtest, gxm = gx - gtest, gym = gy - gtest;
            dcstep(stx, fxm, gxm, sty, fym, gym, stp, fm, gm, brackt, stmin, stmax);
            fx = fxm + stx * gtest;
            fy = fym + sty * gtest;
            gx = gxm + gtest;
            gy = gym + gtest;
        } else {
            dcstep(stx, fx, gx, sty, fy, gy, stp, f, g, brackt, stmin, stmax);
        }
        if (brackt) {
            if (fabs(sty - stx) >= 0.66 * width1) stp = stx + 0.5 * (sty - stx);
            width1 = width;
            width = fabs(sty - stx);
        }
        if (brackt) {
            stmin = fmin(stx, sty);
            stmax = fmax(stx, sty);
        } else {
            stmin = stp + 1.1 * (stp - stx);
            stmax = stp + 4.0 * (stp - stx);
        }
        stp = fmin(fmax(stp, stpmin), stpmax);
        if ((brackt && (stp <= stmin || stp >= stmax)) || (brackt && stmax - stmin <= xtol * stmax)) stp = stx;
        return false;
    }
};

// Minimises fun over x (in place).  lower / upper: +-INFINITY for a missing bound.  last[] receives the LAST point the
// objective was evaluated at (the reference's objective functions mutate the microscope, gym_sted/utils.py:661-686,
// so that point is what the next stage of its optimiser sees).  Returns the number of completed iterations.
template <int N, class F>
LB_HD int minimize(F& fun, double* x, const double* lower, const double* upper, double eps, int maxiter, double tol,
                   double* last, int maxls = 20) {
    static_assert(N == 1 || N == 2, "dense limited-memory matrix: one or two variables");
    double f, g[N];
    double S[kPairs][N], Y[kPairs][N];
    int npairs = 0;
    for (int i = 0; i < N; ++i) last[i] = x[i];

    auto fg = [&](const double* xx, double& fo, double* go) {
        fo = fun(xx);
        for (int i = 0; i < N; ++i) {
            double x1[N];
            for (int j = 0; j < N; ++j) x1[j] = xx[j];
            double h = eps;
            if (xx[i] + h > upper[i]) h = -eps;
            x1[i] += h;
            go[i] = (fun(x1) - fo) / (x1[i] - xx[i]);
            for (int j = 0; j < N; ++j) last[j] = x1[j];
        }
    };
    auto projgr = [&](const double* xx, const double* gg) {
        double out = 0.0;
        for (int i = 0; i < N; ++i) {
            double gi = gg[i];
            if (gi < 0.0) gi = fmax(xx[i] - upper[i], gi);
            else gi = fmin(xx[i] - lower[i], gi);
            out = fmax(out, fabs(gi));
        }
        return out;
    };
    auto dot = [&](const double* a, const double* b) {
        double s = 0.0;
        for (int i = 0; i < N; ++i) s += a[i] * b[i];
        return s;
    };
    // dense B = H^-1, H applied to the unit vectors by the two-loop recursion over the stored pairs
    auto hessian = [&](double (*B)[N]) {
        if (npairs == 0) {
            for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) B[i][j] = (i == j) ? 1.0 : 0.0;
            return;
        }
        double Hm[N][N];
        for (int c = 0; c < N; ++c) {
            double q[N], alpha[kPairs];
            for (int i = 0; i < N; ++i) q[i] = (i == c) ? 1.0 : 0.0;
            for (int k = npairs - 1; k >= 0; --k) {
                alpha[k] = dot(S[k], q) / dot(Y[k], S[k]);
                for (int i = 0; i < N; ++i) q[i] = q[i] - alpha[k] * Y[k][i];
            }
            const double scale = dot(S[npairs - 1], Y[npairs - 1]) / dot(Y[npairs - 1], Y[npairs - 1]);
            for (int i = 0; i < N; ++i) q[i] = q[i] * scale;
            for (int k = 0; k < npairs; ++k) {
                const double beta = dot(Y[k], q) / dot(Y[k], S[k]);
                for (int i = 0; i < N; ++i) q[i] = q[i] + (alpha[k] - beta) * S[k][i];
            }
            for (int i = 0; i < N; ++i) Hm[i][c] = q[i];
        }
        if (N == 1) {
            B[0][0] = 1.0 / Hm[0][0];
        } else {
            const double det = Hm[0][0] * Hm[N - 1][N - 1] - Hm[0][N - 1] * Hm[N - 1][0];
            B[0][0] = Hm[N - 1][N - 1] / det;
            B[0][N - 1] = -Hm[0][N - 1] / det;
            B[N - 1][0] = -Hm[N - 1][0] / det;
            B[N - 1][N - 1] = Hm[0][0] / det;
        }
    };
    auto quad = [&](const double (*B)[N], const double* a, const double* b) {      // a' B b
        double s = 0.0;
        for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) s += a[i] * B[i][j] * b[j];
        return s;
    };

    fg(x, f, g);
    if (projgr(x, g) <= tol) return 0;
    bool boxed = true, cnstnd = false;
    for (int i = 0; i < N; ++i) {
        const bool lo = isfinite(lower[i]), up = isfinite(upper[i]);
        boxed = boxed && lo && up;
        cnstnd = cnstnd || lo || up;
    }
    int it = 0;
    for (; it < maxiter; ++it) {
        double B[N][N];
        hessian(B);
        // ---- generalized Cauchy point
        double tb[N], dd[N], z[N];
        bool fixed[N];
        bool any = false;
        for (int i = 0; i < N; ++i) {
            tb[i] = INFINITY;
            dd[i] = 0.0;
            z[i] = 0.0;
            fixed[i] = false;
            if (g[i] < 0.0 && isfinite(upper[i])) tb[i] = (x[i] - upper[i]) / g[i];
            else if (g[i] > 0.0 && isfinite(lower[i])) tb[i] = (x[i] - lower[i]) / g[i];
            if (tb[i] > 0.0 && g[i] != 0.0) { dd[i] = -g[i]; any = true; }
        }
        if (any) {
            double f1 = dot(g, dd), f2 = quad(B, dd, dd);
            const double f2org = f2;
            double dtm = -f1 / f2, told = 0.0;
            // breakpoints in increasing order (N <= 2)
            int order[N];
            for (int i = 0; i < N; ++i) order[i] = i;
            if (N == 2 && tb[N - 1] < tb[0]) { order[0] = N - 1; order[N - 1] = 0; }
            bool moving = true;
            for (int k = 0; k < N; ++k) {
                const int i = order[k];
                if (!(tb[i] > 0.0) || !isfinite(tb[i])) continue;
                const double dt = tb[i] - told;
                if (dtm < dt) break;
                for (int j = 0; j < N; ++j) z[j] = z[j] + dt * dd[j];
                z[i] = ((g[i] < 0.0) ? upper[i] : lower[i]) - x[i];
                dd[i] = 0.0;
                fixed[i] = true;
                told = tb[i];
                moving = false;
                for (int j = 0; j < N; ++j) moving = moving || dd[j] != 0.0;
                if (!moving) { dtm = 0.0; break; }
                f1 = dot(g, dd) + quad(B, z, dd);
                f2 = fmax(quad(B, dd, dd), kEpsMach * f2org);
                dtm = -f1 / f2;
            }
            dtm = moving ? fmax(dtm, 0.0) : 0.0;
            for (int j = 0; j < N; ++j) z[j] = z[j] + dtm * dd[j];
        }
        double xcp[N], xhat[N];
        int nfree = 0, freei[N];
        for (int i = 0; i < N; ++i) {
            xcp[i] = x[i] + z[i];
            xhat[i] = xcp[i];
            if (!fixed[i] && lower[i] < xcp[i] && xcp[i] < upper[i]) freei[nfree++] = i;
        }
        // ---- subspace minimisation over the free variables (skipped while there is no curvature pair)
        if (nfree > 0 && npairs > 0) {
            double r[N], dF[N];
            for (int i = 0; i < N; ++i) {
                double bz = 0.0;
                for (int j = 0; j < N; ++j) bz += B[i][j] * (xcp[j] - x[j]);
                r[i] = -(g[i] + bz);
            }
            if (nfree == 1) {
                const int i = freei[0];
                dF[0] = r[i] / B[i][i];
            } else {
                const double det = B[0][0] * B[N - 1][N - 1] - B[0][N - 1] * B[N - 1][0];
                dF[0] = (B[N - 1][N - 1] * r[0] - B[0][N - 1] * r[N - 1]) / det;
                dF[N - 1] = (B[0][0] * r[N - 1] - B[N - 1][0] * r[0]) / det;
            }
            double trial[N], proj[N];
            bool clipped = false;
            for (int i = 0; i < N; ++i) trial[i] = xcp[i];
            for (int k = 0; k < nfree; ++k) trial[freei[k]] = xcp[freei[k]] + dF[k];
            for (int i = 0; i < N; ++i) {
                proj[i] = fmin(fmax(trial[i], lower[i]), upper[i]);
                clipped = clipped || proj[i] != trial[i];
            }
            if (clipped) {
                double dd_p = 0.0;
                for (int i = 0; i < N; ++i) dd_p += (proj[i] - x[i]) * g[i];
                if (dd_p > 0.0) {
                    double alpha = 1.0;
                    for (int k = 0; k < nfree; ++k) {
                        const int i = freei[k];
                        if (dF[k] < 0.0 && isfinite(lower[i])) alpha = fmin(alpha, (lower[i] - xcp[i]) / dF[k]);
                        else if (dF[k] > 0.0 && isfinite(upper[i])) alpha = fmin(alpha, (upper[i] - xcp[i]) / dF[k]);
                    }
                    for (int i = 0; i < N; ++i) proj[i] = xcp[i];
                    for (int k = 0; k < nfree; ++k) proj[freei[k]] = xcp[freei[k]] + alpha * dF[k];
                }
                for (int i = 0; i < N; ++i) xhat[i] = proj[i];
            } else {
                for (int i = 0; i < N; ++i) xhat[i] = trial[i];
            }
        }
        double d[N];
        for (int i = 0; i < N; ++i) d[i] = xhat[i] - x[i];
        // ---- line search
        const double dnorm = sqrt(dot(d, d));
        double stpmx = 1e10;
        if (cnstnd) {
            if (it == 0) {
                stpmx = 1.0;
            } else {
                for (int i = 0; i < N; ++i) {
                    if (d[i] < 0.0 && isfinite(lower[i])) {
                        const double a2 = lower[i] - x[i];
                        if (a2 >= 0.0) stpmx = 0.0;
                        else if (d[i] * stpmx < a2) stpmx = a2 / d[i];
                    } else if (d[i] > 0.0 && isfinite(upper[i])) {
                        const double a2 = upper[i] - x[i];
                        if (a2 <= 0.0) stpmx = 0.0;
                        else if (d[i] * stpmx > a2) stpmx = a2 / d[i];
                    }
                }
            }
        }
        double stp = (it == 0 && !boxed) ? fmin(1.0 / dnorm, stpmx) : 1.0;
        double gd = dot(g, d);
        const double gdold = gd, fold = f;
        double xold[N], gold[N];
        for (int i = 0; i < N; ++i) { xold[i] = x[i]; gold[i] = g[i]; }
        bool failed = !(gd < 0.0);                               // not a descent direction
        LineSearch ls(fmin(stp, stpmx), f, gd, 1e-3, 0.9, 0.1, 0.0, stpmx);
        int nls = 0;
        while (!failed) {
            for (int i = 0; i < N; ++i) x[i] = (ls.stp != 1.0) ? xold[i] + ls.stp * d[i] : xold[i] + d[i];
            fg(x, f, g);
            ++nls;
            gd = dot(g, d);
            if (ls.step(f, gd)) break;
            if (nls - 1 >= maxls) failed = true;
        }
        if (failed) {
            // the line search gave up: back to the previous iterate; with curvature pairs the memory is dropped and the
            // iteration restarts from steepest descent, without them the run ends (L-BFGS-B's "abnormal termination")
            for (int i = 0; i < N; ++i) { x[i] = xold[i]; g[i] = gold[i]; }
            f = fold;
            if (npairs == 0) break;
            npairs = 0;
            --it;
            continue;
        }
        stp = ls.stp;
        if (projgr(x, g) <= tol) { ++it; break; }
        if ((fold - f) <= tol * lb_max3(fabs(fold), fabs(f), 1.0)) { ++it; break; }
        double dr, ddum, s[N], y[N];
        for (int i = 0; i < N; ++i) y[i] = g[i] - gold[i];
        if (stp == 1.0) {
            dr = gd - gdold; ddum = -gdold;
            for (int i = 0; i < N; ++i) s[i] = d[i];
        } else {
            dr = (gd - gdold) * stp; ddum = -gdold * stp;
            for (int i = 0; i < N; ++i) s[i] = d[i] * stp;
        }
        if (dr > kEpsMach * ddum) {
            if (npairs == kPairs) {
                for (int k = 1; k < kPairs; ++k) for (int i = 0; i < N; ++i) { S[k - 1][i] = S[k][i]; Y[k - 1][i] = Y[k][i]; }
                --npairs;
            }
            for (int i = 0; i < N; ++i) { S[npairs][i] = s[i]; Y[npairs][i] = y[i]; }
            ++npairs;
        }
    }
    return it;
}

}  // namespace lbfgsb_small
