// Counter-based Philox4x32-10 (Salmon et al., SC'11) and the exact discrete samplers built on it.
//
// pySTED draws bleaching with libc rand() and photons with numpy's global RandomState
// (SURVEY.md §8b), so draw-for-draw equality with the reference is not defined; what is
// required is equality in distribution (moments + KS, north_star).  Every draw here is a pure
// function of (seed, offset, env, pixel, purpose, sub-counter), which makes results independent
// of the launch geometry and of how environments are sharded over GPUs.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define STED_HD __host__ __device__ __forceinline__
#else
#define STED_HD inline
#endif

struct Philox4 { uint32_t x, y, z, w; };

STED_HD void philox_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * (uint64_t)b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
}

STED_HD Philox4 philox4x32_10(Philox4 c, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        philox_mulhilo(M0, c.x, hi0, lo0);
        philox_mulhilo(M1, c.z, hi1, lo1);
        Philox4 n;
        n.x = hi1 ^ c.y ^ k0;
        n.y = lo1;
        n.z = hi0 ^ c.w ^ k1;
        n.w = lo0;
        c = n;
        k0 += W0;
        k1 += W1;
    }
    return c;
}

// purposes (top 4 bits of counter word 2)
#define STED_RNG_PHOTON 1u
#define STED_RNG_BLEACH 2u
#define STED_RNG_TEST   3u
#define STED_RNG_BLEACH2 4u
#define STED_RNG_PHOTON2 5u
#define STED_RNG_BLEACH3 6u   /* per-dwell binomials of the general scan: site = pixel, row field = tap index */

// Stream of 32-bit words for one (pixel, env, purpose, row) site; refills by bumping a sub-counter.
struct PhiloxStream {
    Philox4 ctr, buf;
    uint32_t k0, k1;
    int have;
    STED_HD void init(uint64_t seed, uint32_t offset, uint32_t env, uint32_t site,
                      uint32_t purpose, uint32_t row) {
        k0 = (uint32_t)seed;
        k1 = (uint32_t)(seed >> 32);
        ctr.x = site;
        ctr.y = env;
        ctr.z = (purpose << 28) | ((row & 0xFFFFu) << 12);   // low 12 bits: sub-counter
        ctr.w = offset;
        have = 0;
    }
    // stream whose first word is given (drawn elsewhere, e.g. one Philox block shared by four pixels); the
    // following words come from this site's own counter
    STED_HD void init_with_first(uint32_t first, uint64_t seed, uint32_t offset, uint32_t env, uint32_t site,
                                 uint32_t purpose, uint32_t row) {
        init(seed, offset, env, site, purpose, row);
        buf.w = first;
        have = 1;
    }
    // same with two given words (one Philox block shared by two pixels)
    STED_HD void init_with_two(uint32_t first, uint32_t second, uint64_t seed, uint32_t offset, uint32_t env,
                               uint32_t site, uint32_t purpose, uint32_t row) {
        init(seed, offset, env, site, purpose, row);
        buf.z = first;
        buf.w = second;
        have = 2;
    }
    STED_HD uint32_t next() {
        if (have == 0) {
            buf = philox4x32_10(ctr, k0, k1);
            ctr.z += 1u;
            have = 4;
        }
        uint32_t v = (have == 4) ? buf.x : (have == 3) ? buf.y : (have == 2) ? buf.z : buf.w;
        --have;
        return v;
    }
    // uniform on (0,1): (x + 0.5) * 2^-32 — full relative precision near 0
    STED_HD float uniform() { return ((float)next() + 0.5f) * 2.3283064365386963e-10f; }
};

#if defined(__CUDACC__)

// ---------------------------------------------------------------------------------------------
// Binomial(n, q): number of molecules that bleach out of n, each with probability q.
// Upper-tail inversion: all compared quantities are small numbers held at full float relative
// precision, so rare events (n*q ~ 1e-8) keep their probability.  n is split in chunks of 64 so
// (1-q)^chunk never underflows; q > 1/2 is sampled through the complement.
// Restates the per-molecule rand() <= prob loop of pySTED's sample_molecules
// in distribution.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int binomial_chunk(int n, float q, float log1mq, float u) {
    // q <= 0.5, n <= 64.  p0 = (1-q)^n ; tail_1 = 1 - p0
    const float p0 = __expf((float)n * log1mq);
    float tail = -expm1f((float)n * log1mq);          // P(k >= 1)
    if (u >= tail) return 0;
    const float odds = q / (1.0f - q);
    float pmf = p0;
    int k = 0;
    while (k < n) {
        pmf *= odds * (float)(n - k) / (float)(k + 1);   // pmf(k+1)
        ++k;
        tail -= pmf;                                     // P(K >= k+1)
        if (u >= tail) break;
    }
    return k;
}

// hz = per-molecule hazard over the pass (q = 1 - exp(-hz)); returns number of deaths.
__device__ __forceinline__ int binomial_deaths(int n, float hz, PhiloxStream& rng) {
    if (n <= 0 || hz <= 0.0f) return 0;
    // cheap reject: P(any death) = 1 - exp(-n*hz) <= n*hz ; below 2^-33 no uniform can hit it
    if ((float)n * hz < 1.1e-10f) return 0;
    float q = -expm1f(-hz);
    bool flip = q > 0.5f;
    float log1mq;
    if (flip) { log1mq = log1pf(-__expf(-hz)); q = __expf(-hz); }   // sample survivors instead
    else      { log1mq = -hz; }
    int total = 0, left = n;
    while (left > 0) {
        int m = left < 64 ? left : 64;
        total += binomial_chunk(m, q, log1mq, rng.uniform());
        left -= m;
    }
    return flip ? n - total : total;
}

// ---------------------------------------------------------------------------------------------
// Poisson(lam): inversion below 12, PTRS (Hörmann 1993, the algorithm numpy's legacy
// RandomState.poisson uses) above.  Restates Detector.get_signal's numpy.random.poisson draws.
// ---------------------------------------------------------------------------------------------
static __constant__ float STED_LOGFACT[10] = {0.f, 0.f, 0.6931471806f, 1.7917594692f, 3.1780538303f, 4.7874917428f,
                                              6.5792512120f, 8.5251613611f, 10.6046029027f, 12.8018274801f};

__device__ __forceinline__ float poisson_sample(float lam, PhiloxStream& rng) {
    if (!(lam > 0.0f)) return 0.0f;
    if (lam < 12.0f) {
        // upper-tail inversion (same scheme as binomial_chunk): K = k iff P(K>=k+1) <= w < P(K>=k)
        const float w = rng.uniform();
        float tail = -expm1f(-lam);
        if (w >= tail) return 0.0f;
        float pmf = __expf(-lam);
        int k = 0;
        while (k < 200) {
            ++k;
            pmf *= __fdividef(lam, (float)k);
            tail -= pmf;
            if (w >= tail) break;
        }
        return (float)k;
    }
    const float slam = sqrtf(lam);
    const float b = 0.931f + 2.53f * slam;
    const float a = -0.059f + 0.02483f * b;
    const float invalpha = 1.1239f + 1.1328f / (b - 3.4f);
    const float vr = 0.9277f - 3.6224f / (b - 2.0f);
    for (int it = 0; it < 64; ++it) {
        float U = rng.uniform() - 0.5f;
        float V = rng.uniform();
        float us = 0.5f - fabsf(U);
        float kf = floorf((2.0f * a / us + b) * U + lam + 0.43f);
        if (us >= 0.07f && V <= vr) return kf;
        if (kf < 0.0f || (us < 0.013f && V > us)) continue;
        // acceptance bound log(V*invalpha/(a/us^2+b)) <= -lam + k*log(lam) - lgamma(k+1), all in float:
        // for k >= 10 the right side is rewritten with Stirling's series around k = lam + delta,
        //   -lam*((1+x)*log1p(x) - x) - 0.5*log(2*pi*k) - 1/(12k) + 1/(360k^3),  x = delta/lam,
        // which has no large cancelling terms (absolute error ~1e-4 at lam = 1e6).
        const float lhs = __logf(V * invalpha / (a / (us * us) + b));
        float rhs;
        if (kf < 10.0f) {
            rhs = -lam + kf * logf(lam) - STED_LOGFACT[(int)kf];
        } else {
            const float x = (kf - lam) / lam;
            const float g = (1.0f + x) * log1pf(x) - x;
            const float ik = 1.0f / kf;
            rhs = -lam * g - 0.5f * logf(6.2831853072f * kf) - ik * (0.0833333333f - 0.0027777778f * ik * ik);
        }
        if (lhs <= rhs) return kf;
    }
    return floorf(lam + 0.5f);
}

// First, branch-light stage of poisson_sample for the detector kernel: resolves the common exits with the two given
// words (inversion: K <= 2; PTRS: the squeeze acceptance of the first proposal).  Returns false when the pixel needs
// the full sampler; poisson_sample started from the same two words retraces these steps and continues, so a caller
// may defer such pixels and run them together (same result as calling poisson_sample directly).
__device__ __forceinline__ bool poisson_fast(float lam, uint32_t w0, uint32_t w1, float* out) {
    if (!(lam > 0.0f)) { *out = 0.0f; return true; }
    const float u0 = ((float)w0 + 0.5f) * 2.3283064365386963e-10f;
    if (lam < 12.0f) {
        float tail = -expm1f(-lam);
        if (u0 >= tail) { *out = 0.0f; return true; }
        float pmf = __expf(-lam) * lam;
        tail -= pmf;
        if (u0 >= tail) { *out = 1.0f; return true; }
        pmf *= lam * 0.5f;
        tail -= pmf;
        if (u0 >= tail) { *out = 2.0f; return true; }
        return false;
    }
    const float slam = sqrtf(lam);
    const float b = 0.931f + 2.53f * slam;
    const float a = -0.059f + 0.02483f * b;
    const float vr = 0.9277f - 3.6224f / (b - 2.0f);
    const float U = u0 - 0.5f;
    const float V = ((float)w1 + 0.5f) * 2.3283064365386963e-10f;
    const float us = 0.5f - fabsf(U);
    *out = floorf((2.0f * a / us + b) * U + lam + 0.43f);
    return us >= 0.07f && V <= vr;
}

#endif  // __CUDACC__
