// Host build of gym_sted_b200/csrc/lm_minpack.cuh for the CPU test that pins the solver against scipy's MINPACK.
// Test infrastructure only: the product runs the same header inside rewards_kernels.cu on the GPU.
#include "../gym_sted_b200/csrc/lm_minpack.cuh"

extern "C" int lm_harness_fit(const double* data, int model, double ftol, double* x) {
    if (model == 0) return lm::lm_fit(lm::GaussResiduals{data}, x, ftol, 1.49012e-8, 0.0, 200 * (lm::LM_N + 1), 100.0);
    return lm::lm_fit(lm::RationalResiduals{data}, x, ftol, 1.49012e-8, 0.0, 200 * (lm::LM_N + 1), 100.0);
}
