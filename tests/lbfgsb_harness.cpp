// Host build of gym_sted_b200/csrc/lbfgsb_small.cuh for tests/test_lbfgsb_small.py (g++ -ffp-contract=off).
// Objectives without transcendentals, so that Python floats and C doubles evaluate them bit-identically.
#include "../gym_sted_b200/csrc/lbfgsb_small.cuh"

namespace {
struct Signal1D {            // ((wt - (floor(c * x * 1e5) / 1e5 + 0.1667)) / wt)^2  — a floored, scaled line like the photon count
    double c, wt;
    double operator()(const double* x) const {
        const double s = floor(c * x[0] * 1e5) / 1e5 + 0.1667;
        const double e = (wt - s) / wt;
        return e * e;
    }
};
struct Saturating2D {        // (wt - h / (1 + h))^2,  h = a * x * (1 + y + c * y * y): one equation, two unknowns, like the bleach fit
    double a, c, wt;
    double operator()(const double* x) const {
        const double h = a * x[0] * (1.0 + x[1] + c * x[1] * x[1]);
        const double e = wt - h / (1.0 + h);
        return e * e;
    }
};
}  // namespace

extern "C" int lb_harness_1d(double c, double wt, double* x, double lo, double up, double eps, int maxiter, double tol, double* last) {
    Signal1D f{c, wt};
    return lbfgsb_small::minimize<1>(f, x, &lo, &up, eps, maxiter, tol, last);
}

extern "C" int lb_harness_2d(double a, double c, double wt, double* x, const double* lo, const double* up, double eps, int maxiter,
                             double tol, double* last) {
    Saturating2D f{a, c, wt};
    return lbfgsb_small::minimize<2>(f, x, lo, up, eps, maxiter, tol, last);
}
