// gym_sted_b200 — fluorophore-parameter fitting on the device.
//
// Reference: gym_sted/utils.py:442-686, FluorescenceOptimizer.optimize — called by BleachSampler("uniform" / "choice")
// at every reset of the "hard" environments (utils.py:339-376, ContextualMOSTED-hard-v0, PreferenceCountRateMOSTED-hard-
// v0).  For sampled criteria (a confocal signal of T photons at (p_ex, pdt); a bleached fraction of T at (p_ex, p_sted,
// pdt)) it runs 25 rounds of two scipy L-BFGS-B minimisations on closed forms of the acquisition physics:
//     sigma_abs      <- argmin ((wt - expected_confocal_signal(sigma_abs)) / wt)^2          wt: target blended in, (i+1)/25
//     (k1, b)        <- argmin (wt - expected_bleach(k1, b))^2
// with forward-difference gradients (eps = 0.01 in the scaled variables sigma*1e20, k1*1e15, b), maxiter 25, tol 1e-3.
// The (k1, b) problem has one equation for two unknowns, so WHICH solution comes out is a property of that optimiser's
// path; the parameter distribution of the hard environments is therefore reproduced only by following the same path.
// This file does that: one CTA per environment runs the reference's loop with lbfgsb_small.cuh (pinned against scipy
// by tests/test_lbfgsb_small.py), evaluating the closed forms cooperatively over the K x K taps in float64.
// Two quirks of the reference are part of the path and are kept (utils.py:661-686): its objective functions write the
// trial value into the microscope's fluorophore, so (a) the `current` value of a round and (b) sigma_abs inside the
// bleach fit are those of the LAST EVALUATED trial (the final forward-difference point, x + 0.01), not of the optimum.
// The reference averages 25 noisy detector draws in expected_confocal_signal; here the expectation is used (noise-free
// detector) — tests/test_gpu_optimizer.py compares with the reference's own code run with detector.noise = False.
#include <cstdint>

#include "common.cuh"
#include "lbfgsb_small.cuh"

namespace {

constexpr double kPlanckO = 6.62607015e-34, kLightO = 299792458.0;
#ifndef OPT_THREADS_N
#define OPT_THREADS_N 512
#endif
// One fit is latency: the bleach objective evaluates pow(I, b) in float64 on all K x K = 3 481 taps, ~1 500 times per
// environment, and nothing else of the CTA runs meanwhile.  512 threads = 7 taps each (128 threads, 27 taps each: 34 ms
// per launch whatever the batch; the order of the partial sums - and with it the last bits of an objective value -
// depends on this number, the path the optimiser follows does not: tests/test_gpu_optimizer.py).
constexpr int OPT_THREADS = OPT_THREADS_N;

// CTA-wide sum; every thread returns the same value (fixed reduction order: bit-reproducible)
__device__ double cta_sum(double v, double* scratch) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();                                  // scratch free (previous broadcast consumed)
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < OPT_THREADS / 32; ++w) s += scratch[w];
    return s;
}

struct Forms {
    // per-tap constants in shared memory
    // signal criterion, on the (2r+1)^2 taps where the test structure (a Gaussian blob) is not zero:
    const double *iex_sig, *eta_cur, *psf_det, *blob;  // I_ex = i_ex * p_ex, eta at the criterion's p_sted, detection PSF, blob
    const double *phi_ex, *i_inst;                     // bleach criterion, all taps: excitation photon flux, in-pulse STED intensity
    double* scratch;
    int KK, NB;
    StedFluoParams f;
    StedDetectorParams det;
    StedFluoCriteria c;
    double e_fluo;

    // FluorescenceOptimizer.expected_confocal_signal (utils.py:634-661) with sigma_abs = sig * 1e-20; `with_sted`: the
    // criterion's own p_sted (the `current` value of a round), else p_sted = 0 (optimize_sigma_abs, utils.py:683)
    __device__ double signal(double sig, bool with_sted) const {
        const double sigma = sig * 1e-20;
        double part = 0.0;
        for (int t = threadIdx.x; t < NB; t += OPT_THREADS) {
            const double exc = sigma * iex_sig[t] * f.qy;
            part += exc * (with_sted ? eta_cur[t] : 1.0) * psf_det[t] * blob[t];
        }
        const double intensity = cta_sum(part, scratch);
        const double photons = floor(intensity / e_fluo);
        return photons * det.pcef * det.pdef * c.sig_pdt + det.background * c.sig_pdt;
    }
    // FluorescenceOptimizer.expected_bleach (utils.py:606-632) with k1 = k1p * 1e-15
    __device__ double bleach(double k1p, double b, double sig) const {
        const double sigma = sig * 1e-20, k1 = k1p * 1e-15;
        double part = 0.0;
        for (int t = threadIdx.x; t < KK; t += OPT_THREADS) {
            const double I = i_inst[t];
            const double kb = sigma * phi_ex[t] * f.sted_tau * (f.k0 * I + k1 * pow(I, b));
            part += kb * c.bl_pdt;
        }
        return 1.0 - exp(-cta_sum(part, scratch));
    }
};

struct SigmaObjective {
    const Forms& F;
    double wt;
    __device__ double operator()(const double* x) const {
        const double e = (wt - F.signal(x[0], false)) / wt;
        return e * e;
    }
};
struct BleachObjective {
    const Forms& F;
    double wt, sig;
    __device__ double operator()(const double* x) const {
        const double e = wt - F.bleach(x[0], x[1], sig);
        return e * e;
    }
};

__global__ void __launch_bounds__(OPT_THREADS) fluo_optimize_kernel(const double* __restrict__ cache, const double* __restrict__ blob,
                                                                    const StedFluoParams f0, const StedDetectorParams det,
                                                                    const StedFluoCriteria* __restrict__ crit, int K, int radius,
                                                                    int iterations, double* __restrict__ out) {
    extern __shared__ double sm[];
    const int KK = K * K, env = blockIdx.x, side = 2 * radius + 1, NB = side * side, ctr = K / 2;
    double* s_iex = sm;
    double* s_eta = s_iex + NB;
    double* s_psf = s_eta + NB;
    double* s_blob = s_psf + NB;
    double* s_phi = s_blob + NB;
    double* s_ii = s_phi + KK;
    double* s_scratch = s_ii + KK;
    const StedFluoCriteria c = crit[env];
    const double* i_ex = cache;
    const double* i_sted = cache + (size_t)KK;
    const double* psf_det = cache + 2 * (size_t)KK;

    const double e_ex = kPlanckO * kLightO / f0.lambda_ex, e_st = kPlanckO * kLightO / f0.lambda_sted;
    const double k_vib = 1.0 / f0.tau_vib, k_s1 = 1.0 / f0.tau, T = 1.0 / f0.sted_rate;
    const double i_sat = e_st / (f0.sigma_ste * f0.tau);
    const double denom = 1.0 - exp(-k_s1 * T);
    for (int q = threadIdx.x; q < NB; q += OPT_THREADS) {
        const int t = (ctr - radius + q / side) * K + (ctr - radius + q % side);
        s_iex[q] = i_ex[t] * c.sig_p_ex;
        const double I_st = i_sted[t] * c.sig_p_sted;
        const double zeta = I_st / i_sat;
        const double gamma = zeta * k_vib / (zeta * k_s1 + k_vib);
        s_eta[q] = ((1.0 + gamma * exp(-k_s1 * f0.sted_tau * (1.0 + gamma))) / (1.0 + gamma) - exp(-k_s1 * (gamma * f0.sted_tau + T))) / denom;
        s_psf[q] = psf_det[t];
        s_blob[q] = blob[t];
    }
    for (int t = threadIdx.x; t < KK; t += OPT_THREADS) {
        // expected_bleach: photon fluxes with get_photons' floor division, STED flux time averaged, then back to the
        // in-pulse intensity inside get_k_bleach (utils.py:616-630)
        s_phi[t] = floor(i_ex[t] * c.bl_p_ex / e_ex);
        const double phi_st = floor(i_sted[t] * c.bl_p_sted / e_st) * f0.sted_tau / T;
        s_ii[t] = phi_st * e_st * T / f0.sted_tau;
    }
    __syncthreads();
    Forms F{s_iex, s_eta, s_psf, s_blob, s_phi, s_ii, s_scratch, KK, NB, f0, det, c, kPlanckO * kLightO / f0.lambda_fluo};

    // default_parameters(): the starting point, in the scaled variables (utils.py:566-571)
    double k1p = f0.k1 * 1e15, b = f0.b, sig = f0.sigma_abs_ex * 1e20;
    double sig_state = sig, k1_state = k1p, b_state = b;       // what the microscope's fluorophore holds
    const bool has_signal = c.sig_target > 0.0, has_bleach = c.bl_target > 0.0;
    const double lo1[1] = {0.0}, up1[1] = {INFINITY}, lo2[2] = {0.0, 0.0}, up2[2] = {INFINITY, 5.0};
    for (int i = 0; i < iterations; ++i) {
        const double w = (double)(i + 1) / (double)iterations;
        if (has_signal) {
            const double cur = F.signal(sig_state, true);
            SigmaObjective obj{F, cur + w * (c.sig_target - cur)};
            double x[1] = {sig}, last[1];
            lbfgsb_small::minimize<1>(obj, x, lo1, up1, 0.01, 25, 1e-3, last);
            sig = x[0];
            sig_state = last[0];
        }
        if (has_bleach) {
            const double cur = F.bleach(k1_state, b_state, sig_state);
            BleachObjective obj{F, cur + w * (c.bl_target - cur), sig_state};
            double x[2] = {k1p, b}, last[2];
            lbfgsb_small::minimize<2>(obj, x, lo2, up2, 0.01, 25, 1e-3, last);
            k1p = x[0]; b = x[1];
            k1_state = last[0]; b_state = last[1];
        }
    }
    const double reached_s = F.signal(sig, false), reached_b = F.bleach(k1p, b, sig);
    if (threadIdx.x == 0) {
        double* o = out + (size_t)env * 5;
        o[0] = sig * 1e-20; o[1] = k1p * 1e-15; o[2] = b; o[3] = reached_s; o[4] = reached_b;
    }
}

}  // namespace

extern "C" int sted_fluo_optimize(const double* psf_cache_dev, const double* blob_dev, int K, int blob_radius, const StedFluoParams* fluo_host,
                                  const StedDetectorParams* det_host, const StedFluoCriteria* criteria_dev, int B, int iterations,
                                  double* out_dev, void* stream) {
    if (!psf_cache_dev || !blob_dev || !fluo_host || !det_host || !criteria_dev || !out_dev) return fail(STED_EINVAL, "null pointer argument");
    if (B <= 0 || iterations <= 0) return fail(STED_EINVAL, "B and iterations must be positive");
    if (K < 1 || (K & 1) == 0 || K > STED_MAX_K) return fail(STED_EUNSUPPORTED, "K must be odd and <= %d (got %d)", STED_MAX_K, K);
    if (blob_radius < 0 || blob_radius > K / 2) return fail(STED_EINVAL, "blob_radius must lie in [0, K/2]");
    if (fluo_host->triplet_dynamics_frac != 0.0) return fail(STED_EUNSUPPORTED, "triplet dynamics are not part of the restated model");
    int rc = check_device();
    if (rc) return rc;
    const size_t nb = (size_t)(2 * blob_radius + 1) * (2 * blob_radius + 1);
    const size_t smem = (4 * nb + (size_t)2 * K * K + OPT_THREADS / 32 + 8) * sizeof(double);
    CUDA_TRY(cudaFuncSetAttribute(fluo_optimize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fluo_optimize_kernel<<<B, OPT_THREADS, smem, (cudaStream_t)stream>>>(psf_cache_dev, blob_dev, *fluo_host, *det_host, criteria_dev, K,
                                                                       blob_radius, iterations, out_dev);
    CUDA_TRY(cudaGetLastError());
    return STED_OK;
}
