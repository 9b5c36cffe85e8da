// sted_kernels.cu — B200 (sm_100a) kernels + C ABI for the STED acquisition hot path.
//
// Path rebuilt here: pysted.base.Microscope.get_effective / fluo.get_k_bleach
// (a2, a3) -> make_effective_kernel; Microscope.get_signal_and_bleach (a5) ->
// gather_kernel (bleach=False) and presample_kernel + bucket_scan_kernel + sweep_kernel (bleach=True);
// Detector.get_signal -> detect_kernel.  Reference call sites: gym_sted/envs/mosted_env.py:184,222,227,232 and
// gym_sted/utils.py:616-631,646-657.  See DESIGN.md for the derivations.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 (see __graft_entry__.build).
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <mutex>

#include "../../include/sted_b200.h"
#include "philox.cuh"

#include "common.cuh"

thread_local char sted_err_buf[512] = "";

namespace {

constexpr double kPlanck = 6.62607015e-34;
constexpr double kLight = 299792458.0;

// ------------------------------------------------------------------------------------------
// K1: effective PSF + bleaching tables, one CTA per environment: taps in parallel, suffix sums by one thread per tap row.
// All physics in float64; tables are rounded to float32 once.
//   E[u][v]  = sigma_abs*I_ex*qy * eta(I_sted) * psf_det            (Microscope.get_effective)
//   k_b[u][v]= sigma_abs*phi_ex * tau_sted * (k0*I + k1*I^b),  I = instantaneous STED W/m^2
//                                                                     (fluo.get_k_bleach)
//   A = k_b*pdt;  CH[u][v] = sum_{v'>=v} A[u][v'];  ET = E*exp(-CH[u][v+1]);  RS = exp(-CH)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) make_effective_kernel(const double* __restrict__ cache, const StedFluoParams* __restrict__ fluo,
                                      const double* __restrict__ p_ex, const double* __restrict__ p_sted,
                                      const double* __restrict__ pdt, int K, int KP,
                                      float* __restrict__ tables, double* __restrict__ kbleach) {
    extern __shared__ double sm_d[];                     // [K*K] E, then [K*K] A = k_b * pdt
    double* sE = sm_d;
    double* sA = sm_d + K * K;
    const int b = blockIdx.x;
    const StedFluoParams f = fluo[b];
    const double pex = p_ex[b], psted = p_sted[b], dwell = pdt[b];
    const double* i_ex = cache;
    const double* i_sted = cache + (size_t)K * K;
    const double* psf_det = cache + 2 * (size_t)K * K;

    const double e_ex = kPlanck * kLight / f.lambda_ex;
    const double e_st = kPlanck * kLight / f.lambda_sted;
    const double k_vib = 1.0 / f.tau_vib, k_s1 = 1.0 / f.tau;
    const double T = 1.0 / f.sted_rate;
    const double i_sat = e_st / (f.sigma_ste * f.tau);
    const double duty = f.sted_tau * f.sted_rate;
    const double denom = 1.0 - exp(-k_s1 * T);

    // every tap independently (all threads), then one thread per tap row for the suffix sums
    for (int t = threadIdx.x; t < K * K; t += blockDim.x) {
        const double I_ex = i_ex[t] * pex;
        const double I_st = i_sted[t] * psted;                 // instantaneous, in-pulse
        double exc = f.sigma_abs_ex * I_ex * f.qy;
        if (f.anti_stoke) exc += f.sigma_abs_sted * I_st * f.qy;
        const double zeta = I_st / i_sat;
        const double gamma = zeta * k_vib / (zeta * k_s1 + k_vib);
        const double eta = ((1.0 + gamma * exp(-k_s1 * f.sted_tau * (1.0 + gamma))) / (1.0 + gamma)
                            - exp(-k_s1 * (gamma * f.sted_tau + T))) / denom;
        // photon fluxes with get_photons' floor division
        const double phi_ex = floor(I_ex / e_ex);
        const double phi_st = floor(I_st * duty / e_st);        // time averaged
        const double I_inst = phi_st * e_st / duty;
        const double kb = f.sigma_abs_ex * phi_ex * f.sted_tau * (f.k0 * I_inst + f.k1 * pow(I_inst, f.b));
        if (kbleach) kbleach[(size_t)b * K * K + t] = kb;
        sE[t] = exc * eta * psf_det[t];
        sA[t] = kb * dwell;
    }
    __syncthreads();
    // row-interleaved copy of E for the FFMA2 gather: E2[u][v] = (E[u][v], E[u-1][v]), u = 0..K, zero outside
    {
        float2* tE2 = reinterpret_cast<float2*>(tables + ((size_t)b * STED_TABLES + STED_TAB_E2) * K * KP);
        for (int i = threadIdx.x; i < (K + 1) * KP; i += blockDim.x) {
            const int uu = i / KP, v = i - uu * KP;
            tE2[i] = make_float2((uu < K && v < K) ? (float)sE[uu * K + v] : 0.f, (uu > 0 && v < K) ? (float)sE[(uu - 1) * K + v] : 0.f);
        }
    }
    const int u = threadIdx.x;
    if (u >= K) return;
    float* tE = tables + ((size_t)b * STED_TABLES + STED_TAB_E) * K * KP + (size_t)u * KP;
    float* tET = tables + ((size_t)b * STED_TABLES + STED_TAB_ET) * K * KP + (size_t)u * KP;
    float* tCH = tables + ((size_t)b * STED_TABLES + STED_TAB_CH) * K * KP + (size_t)u * KP;
    float* tRS = tables + ((size_t)b * STED_TABLES + STED_TAB_RS) * K * KP + (size_t)u * KP;
    for (int v = K; v < KP; ++v) { tE[v] = 0.f; tET[v] = 0.f; tCH[v] = 0.f; tRS[v] = 1.f; }
    double ch = 0.0;   // suffix hazard, exclusive of the tap being visited on entry
    for (int v = K - 1; v >= 0; --v) {
        const double E = sE[u * K + v];
        tE[v] = (float)E;
        tET[v] = (float)(E * exp(-ch));
        ch += sA[u * K + v];
        tCH[v] = (float)ch;
        tRS[v] = (float)exp(-ch);
    }
}

// ------------------------------------------------------------------------------------------
// shared launch parameters
// ------------------------------------------------------------------------------------------
struct AcqParams {
    const float* din;
    float* dout;
    const float* tables;
    const double* pdt;
    float* intensity;
    float* signal;
    int B, H, W, K, KP, Hp, Wp, Wpitch;
    uint32_t flags;
    uint32_t offset;
    uint64_t seed;
    uint32_t env_base;
    float inv_e_photon;   // lambda_fluo/(h c)
    float det_eff;        // pcef*pdef
    float background;     // counts per second
    const float* wdt;     // [B][H][W] dwell time of every pixel relative to pdt[b], or nullptr (general scan only)
};

// Detector.get_signal: photons = floor(I/e_photon); mean = photons*pcef*pdef*pdt + bg*pdt; Poisson when noisy.
// `first` is this pixel's first random word (one Philox block serves four neighbouring pixels).
__device__ __forceinline__ float detect(const AcqParams& P, float I, float gain, float bg_mean,
                                        uint32_t env, uint32_t site, uint32_t first, uint32_t second) {
    const float mean = floorf(I * P.inv_e_photon) * gain + bg_mean;
    if (P.flags & STED_SAMPLE_PHOTONS) {
        PhiloxStream rng;
        rng.init_with_two(first, second, P.seed, P.offset, env, site, STED_RNG_PHOTON2, 0);
        return poisson_sample(mean, rng);
    }
    return mean;
}

// ------------------------------------------------------------------------------------------
// K2: dense PSF-window gather, no bleaching.
//   I[r][c] = sum_{u,v} E[u][v] * D[r+u][c+v]
// CTA = 64x64 output tile of one environment, 256 threads, thread = 4 rows x 4 columns, 3 CTAs (24 warps) per SM.
// smem: datamap tile (64+K-1) x (64+KP) floats + E table, both staged by TMA.  Per 4 taps and 4 output rows a
// thread issues 1 LDS.128 of datamap (sliding window) + 4 broadcast LDS.128 of E for 64 FFMA.
// ------------------------------------------------------------------------------------------
#ifndef G_UNROLL
#define G_UNROLL 15  // tap chunks (of 4) unrolled per iteration of the gather's inner loop (15 = the whole tap row)
#endif
#ifndef G_TILE_H
#define G_TILE_H 64
#endif
#ifndef G_ROWS_PER_THREAD
#define G_ROWS_PER_THREAD 4   // measured on B200: 8 -> 2.46 ms, 4 -> 2.0 ms, 2 -> 2.05 ms per 256x256^2 launch
#endif
constexpr int G_TH = G_TILE_H, G_TW = 64, G_RT = G_ROWS_PER_THREAD, kGatherUnroll = G_UNROLL;   // threads per CTA: (TH / G_RT) * 16
constexpr int G_MINB = 2;       // tile (60.5 KB) + row-interleaved PSF (28.8 KB) per CTA

#ifndef G_TRIM_COLS
#define G_TRIM_COLS 1  // 1: the padding taps of a tap row (columns K..KP-1) are not issued
#endif
#ifndef G_TRIM_ROWS
#define G_TRIM_ROWS 0  // 1/2: the dead half of the pairs at u = 0 and u = K becomes a scalar FFMA (1: compact edge rows, 2: unrolled).
                       // Measured on B200 (256 x 256^2): 2.053 ms off, 2.135 (1), 2.186 (2) - the extra code costs more than the
                       // 1.7 % of pipe time it saves; G_TRIM_COLS alone: 2.034 ms
#endif
#ifndef G_SKIP
#define G_SKIP 1       // 1: band rows whose datamap rows are all zero are skipped (bit-identical results); 2: also tap chunks whose
                       // windows are empty.  Measured on B200, 256 x 256^2 bench maps (68 % of the band rows live): 2.038 ms off,
                       // 1.491 ms (1), 1.523 ms (2: 61 % of the chunks live, but the per-chunk branch breaks the load pipelining)
#endif
#ifndef GATHER_TMA
#define GATHER_TMA 1   // 0: stage the tile with coalesced float4 loads instead of TMA (A/B timing only)
#endif

__device__ unsigned long long g_phase_cycles[8];   // timing experiments only (SWEEP_PROBE==2 / GATHER_PROBE builds)

// packed FP32 pairs for FFMA2 (fma.rn.f32x2, sm_100): lo = first element.  A pair built from one scalar becomes the
// instruction's scalar-broadcast operand (R.F32), so pack2(w, w) costs nothing.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }


__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (long long spin = 0; !ok; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1ll << 28)) __trap();           // a lost TMA must not hang the GPU
    }
}

// TH: output rows of a tile.  64 everywhere except for image heights whose last 64-row tile would be at most half full
// (config 4's 96 rows: three 32-row tiles instead of two 64-row ones, a quarter of the work less); a 32-row tile is a
// 128-thread CTA, three of them per SM.
template <int KT, int TH>
__global__ void __launch_bounds__((TH / G_RT) * 16, TH >= 64 ? G_MINB : 3) gather_kernel(const AcqParams P, const __grid_constant__ CUtensorMap tmap) {
    constexpr int THREADS = (TH / G_RT) * 16;
    const int K = KT ? KT : P.K;
    const int KP = KT ? STED_KPITCH(KT) : P.KP;
    const int PITCH = G_TW + KP;
    const int rows = TH + K - 1;
    extern __shared__ __align__(128) float smem[];
    float* sD = smem;
    // PSF rows interleaved in pairs for FFMA2: sE2[u][v] = (E[u][v], E[u-1][v]) for u = 0..K, E[-1] = E[K] = 0
    float2* sE2 = reinterpret_cast<float2*>(smem + rows * PITCH);
    __shared__ __align__(8) unsigned long long mbar;

#ifdef GATHER_PROBE
    const long long t_start = clock64();
#endif
    const int b = blockIdx.z;
    const int tr0 = blockIdx.y * TH, tc0 = blockIdx.x * G_TW;
    const int tid = threadIdx.x;
    const float* gE2 = P.tables + ((size_t)b * STED_TABLES + STED_TAB_E2) * K * KP;      // written by make_effective_kernel

#if GATHER_TMA
    // ---- stage tile + PSF with TMA: one 3-D tensor box (zero filled outside the padded map) and one bulk copy,
    // both completing on the same mbarrier
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes_d = (uint32_t)(rows * PITCH * sizeof(float)), bytes_e = (uint32_t)((K + 1) * KP * sizeof(float2));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&mbar)), "r"(bytes_d + bytes_e) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     :: "r"(smem_u32(sD)), "l"(&tmap), "r"(tc0), "r"(tr0), "r"(b), "r"(smem_u32(&mbar)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(sE2)), "l"(gE2), "r"(bytes_e), "r"(smem_u32(&mbar)) : "memory");
    }
    mbar_wait(smem_u32(&mbar), 0);
#else
    // ---- stage tile + PSF (float4, zero fill outside the padded map)
    {
        const float* src = P.din + (size_t)b * P.Hp * P.Wpitch;
        const int q = PITCH / 4;
        for (int i = tid; i < rows * q; i += THREADS) {
            const int ry = i / q, cx = (i - ry * q) * 4;
            const int y = tr0 + ry, x = tc0 + cx;
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y < P.Hp && x < P.Wpitch) val = __ldg(reinterpret_cast<const float4*>(src + (size_t)y * P.Wpitch + x));
            *reinterpret_cast<float4*>(sD + ry * PITCH + cx) = val;
        }
        const float4* e4 = reinterpret_cast<const float4*>(gE2);
        for (int i = tid; i < (K + 1) * KP / 2; i += THREADS) reinterpret_cast<float4*>(sE2)[i] = __ldg(e4 + i);
    }
    __syncthreads();
#endif

#if G_SKIP
    // Molecule maps are sparse (about 14 % of the pixels of a structure hold molecules, and the K/2 frame none): a band
    // row whose tile rows are all zero adds nothing to the outputs of the warp that would read it.  zrow[y] marks the
    // 4-column segments of tile row y that hold a molecule; a warp (two groups of G_RT output rows, all 64 columns)
    // skips band row yy when both tile rows it would read are empty.  Skipped terms are E * 0 added to a non-negative
    // sum: the results are bit-identical.
    __shared__ uint32_t zrow[128];
    {
        const int nseg = PITCH / 4;                       // <= 32
        for (int ry = tid >> 5; ry < rows; ry += THREADS / 32) {
            bool nz = false;
            if ((tid & 31) < nseg) {
                const float4 v = *reinterpret_cast<const float4*>(sD + ry * PITCH + 4 * (tid & 31));
                nz = v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f;
            }
            const unsigned m = __ballot_sync(0xffffffffu, nz);
            if ((tid & 31) == 0) zrow[ry] = m;
        }
        __syncthreads();
    }
    const int zr0 = (tid >> 5) * 2 * G_RT;                 // first tile row of the warp's first row group
    auto empty_row = [&](const int yy) { return (zrow[zr0 + yy] | zrow[zr0 + G_RT + yy]) == 0u; };
#else
    auto empty_row = [&](const int) { return false; };
#endif
#ifdef GATHER_PROBE
    const long long t_ready = clock64();
#endif
    const int lane16 = tid & 15, rg = tid >> 4;
    const int r0 = rg * G_RT, c0 = lane16 * 4;
    // Accumulators are packed pairs of two output ROWS (2p, 2p+1) of one column, and the inner loop issues FFMA2
    // (fma.rn.f32x2): the FP32 pipe delivers the same flops as with scalar FFMA but one instruction covers two, which
    // takes the kernel from issue bound (85 % of the issue slots for 77 % of the FMA pipe) to FMA-pipe bound.  Both
    // rows read the same datamap value (the instruction's scalar-broadcast operand) against the PSF taps of rows u and
    // u-1 (one 64-bit shared-memory word of the interleaved table).  Each half performs exactly the fused
    // multiply-adds of the scalar form, in the same order: results are bit-identical to it.
    static_assert(G_RT % 2 == 0, "row pairs");
    constexpr int NP = G_RT / 2;
    u64 acc[NP][4];
#pragma unroll
    for (int p = 0; p < NP; ++p) { acc[p][0] = acc[p][1] = acc[p][2] = acc[p][3] = 0ull; }

    const int nchunk = KP / 4;
    // taps of the last chunk of a tap row that exist (3 for K = 59, pitch 60): the padding taps are not issued
    constexpr int kTailTaps = KT ? KT - 4 * (STED_KPITCH(KT) / 4 - 1) : 4;
    static_assert(kTailTaps >= 0 && kTailTaps <= 4, "tap pitch");
    auto band_row = [&](const int yy, const bool check) {
        const float* drow = sD + (r0 + yy) * PITCH + c0;
        float4 wa = *reinterpret_cast<const float4*>(drow);
#if G_SKIP >= 2
        const uint32_t zbits = zrow[zr0 + yy] | zrow[zr0 + G_RT + yy];
#endif
#pragma unroll kGatherUnroll
        for (int vc = 0; vc < nchunk; ++vc) {
            const float4 wb = *reinterpret_cast<const float4*>(drow + 4 * (vc + 1));
#if G_SKIP >= 2
            // the warp's windows of this chunk cover tile columns 4 vc .. 4 vc + 66: segments vc .. vc + 16
            if (((zbits >> vc) & 0x1FFFFu) == 0u) { wa = wb; continue; }
#endif
            const u64 w[7] = {pack2(wa.x, wa.x), pack2(wa.y, wa.y), pack2(wa.z, wa.z), pack2(wa.w, wa.w),
                              pack2(wb.x, wb.x), pack2(wb.y, wb.y), pack2(wb.z, wb.z)};
            const bool tail = G_TRIM_COLS && KT && vc == nchunk - 1;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const int u = yy - 2 * p;                   // tap row of output row 2p; row 2p+1 uses u-1
                if (!check || (u >= 0 && u <= K)) {
                    const ulonglong2* e2 = reinterpret_cast<const ulonglong2*>(sE2 + u * KP + 4 * vc);
                    const ulonglong2 ea = e2[0], eb = e2[1];              // taps v, v+1 | v+2, v+3
                    const u64 e[4] = {ea.x, ea.y, eb.x, eb.y};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
#pragma unroll
                        for (int t = 0; t < 4; ++t)
                            if (!tail || t < kTailTaps) acc[p][j] = fma2(e[t], w[j + t], acc[p][j]);
                    }
                }
            }
            wa = wb;
        }
    };
    // Rows of the ramps where a pair has only ONE live half - tap row u = 0 (the odd output row would use E[-1] = 0) or
    // u = K (the even row would use E[K] = 0): that half is a scalar FFMA instead of an FFMA2 whose other half adds 0.
    // Compact code (not unrolled over the tap chunks): six of the 62 band rows of a thread go through here.
    auto band_row_edge = [&](const int yy) {
        const float* drow = sD + (r0 + yy) * PITCH + c0;
        float4 wa = *reinterpret_cast<const float4*>(drow);
#if G_TRIM_ROWS == 2
#pragma unroll kGatherUnroll
#else
#pragma unroll 1
#endif
        for (int vc = 0; vc < nchunk; ++vc) {
            const float4 wb = *reinterpret_cast<const float4*>(drow + 4 * (vc + 1));
            const float wf[7] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z};
            const int nt = (G_TRIM_COLS && KT && vc == nchunk - 1) ? kTailTaps : 4;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const int u = yy - 2 * p;
                if (u < 0 || u > K) continue;
                const float2* ef = sE2 + u * KP + 4 * vc;
                if (u == 0 || u == K) {
                    const bool low = u == 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float lo, hi;
                        unpack2(acc[p][j], lo, hi);
                        float a = low ? lo : hi;
#pragma unroll
                        for (int t = 0; t < 4; ++t)
                            if (t < nt) a = fmaf(low ? ef[t].x : ef[t].y, wf[j + t], a);
                        acc[p][j] = low ? pack2(a, hi) : pack2(lo, a);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
#pragma unroll
                        for (int t = 0; t < 4; ++t)
                            if (t < nt) acc[p][j] = fma2(*reinterpret_cast<const u64*>(ef + t), pack2(wf[j + t], wf[j + t]), acc[p][j]);
                    }
                }
            }
            wa = wb;
        }
    };
    int yy = 0;
#if G_TRIM_ROWS
    // band rows 0 .. 2(NP-1) and K .. K+2(NP-1)+1 hold the half-live pairs and the ramps; in between every pair is full
    for (; yy <= 2 * (NP - 1) && yy < K; ++yy) band_row_edge(yy);
    for (; yy < K; ++yy) band_row(yy, false);
    for (; yy < G_RT + K - 1; ++yy) band_row_edge(yy);
#else
    for (; yy < 2 * (NP - 1) && yy <= K; ++yy) if (!empty_row(yy)) band_row(yy, true);  // ramp-up
    for (; yy <= K; ++yy) if (!empty_row(yy)) band_row(yy, false);                      // steady: every row pair has a tap row
    for (; yy < G_RT + K - 1; ++yy) if (!empty_row(yy)) band_row(yy, true);             // ramp-down
#endif

#ifdef GATHER_PROBE
    const long long t_main = clock64();
#endif
    // ---- epilogue: stage the tile in shared memory (the datamap tile is dead), then run the detector
    // in a compact, non-unrolled loop with coalesced stores.  (Unrolling the Poisson sampler over the
    // 32 register accumulators made the kernel instruction-fetch bound: ncu stall_no_instruction.)
    __syncthreads();
    float* sOut = sD;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        float4 lo, hi;
        unpack2(acc[p][0], lo.x, hi.x); unpack2(acc[p][1], lo.y, hi.y); unpack2(acc[p][2], lo.z, hi.z); unpack2(acc[p][3], lo.w, hi.w);
        *reinterpret_cast<float4*>(sOut + (r0 + 2 * p) * G_TW + c0) = lo;
        *reinterpret_cast<float4*>(sOut + (r0 + 2 * p + 1) * G_TW + c0) = hi;
    }
    __syncthreads();
    // detected power goes to `signal`; detect_kernel turns it into photon counts in place
    for (int idx = tid; idx < TH * G_TW; idx += THREADS) {
        const int r = tr0 + (idx / G_TW), c = tc0 + (idx % G_TW);
        if (r >= P.H || c >= P.W) continue;
        const float I = sOut[idx];
        const size_t o = ((size_t)b * P.H + r) * P.W + c;
        if (P.intensity) P.intensity[o] = I;
        P.signal[o] = I;
    }
#ifdef GATHER_PROBE
    if (tid == 0) {
        const long long t_end = clock64();
        atomicAdd(&g_phase_cycles[0], (unsigned long long)(t_ready - t_start));
        atomicAdd(&g_phase_cycles[1], (unsigned long long)(t_main - t_ready));
        atomicAdd(&g_phase_cycles[2], (unsigned long long)(t_end - t_main));
        atomicAdd(&g_phase_cycles[3], 1ull);
    }
#endif
}

// Detector.get_signal for every pixel of the batch, in place on `signal` (power -> photon counts).
// A separate, light kernel: the Poisson sampler is latency bound and runs at full occupancy here
// instead of at the 12 warps/SM of the gather kernel.
constexpr int DET_THREADS = 256, DET_QCAP = 1024;     // ring capacity >= (DET_THREADS - 1) + 2 * DET_THREADS

__global__ void __launch_bounds__(DET_THREADS, 5) detect_kernel(const AcqParams P) {
    // thread = two neighbouring pixels sharing one Philox block (two words each).  Stage 1 computes the Poisson mean
    // and resolves the common sampler exits in straight-line code (poisson_fast).  Every other pixel is queued in a
    // shared-memory ring by sampler class - inversion (mean < 12) or PTRS - and the rings are drained 256 entries at
    // a time by whole warps, so the lanes of a warp run the same slow branch instead of the union of both at 3-9
    // active lanes.
    __shared__ uint32_t q_pix[2][DET_QCAP], q_w0[2][DET_QCAP], q_w1[2][DET_QCAP];
    __shared__ unsigned q_head[2], q_tail[2];
    const size_t hw = (size_t)P.H * P.W;
    const size_t groups_per_env = (hw + 1) / 2, n_groups = groups_per_env * P.B;
    const bool sample = (P.flags & STED_SAMPLE_PHOTONS) != 0;
    if (threadIdx.x < 2) { q_head[threadIdx.x] = 0; q_tail[threadIdx.x] = 0; }
    __syncthreads();
    auto drain = [&](bool all) {            // all threads; whole batches of DET_THREADS entries unless `all`
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
            const unsigned head = q_head[q], pending = q_tail[q] - head;
            const unsigned n = all ? pending : pending / DET_THREADS * DET_THREADS;
#pragma unroll 1
            for (unsigned i = threadIdx.x; i < n; i += DET_THREADS) {
                const unsigned slot = (head + i) % DET_QCAP;
                const uint32_t pix = q_pix[q][slot];              // pixel index within this launch (< 2^31)
                const int b = (int)(pix / hw);
                const uint32_t site = (uint32_t)(pix - (size_t)b * hw);
                const float dwell = (float)P.pdt[b] * (P.wdt ? P.wdt[pix] : 1.0f);
                float* px = P.signal + pix;
                *px = detect(P, *px, P.det_eff * dwell, P.background * dwell, P.env_base + (uint32_t)b, site, q_w0[q][slot],
                             q_w1[q][slot]);
            }
            __syncthreads();
            if (threadIdx.x == 0) q_head[q] = head + n;
        }
        __syncthreads();
    };
    for (size_t base = (size_t)blockIdx.x * DET_THREADS; base < n_groups; base += (size_t)gridDim.x * DET_THREADS) {
        const size_t g = base + threadIdx.x;
        bool full = false;                  // this thread's push completed a batch of DET_THREADS pending entries
        if (g < n_groups) {
            const int b = (int)(g / groups_per_env);
            const uint32_t grp = (uint32_t)(g - (size_t)b * groups_per_env);
            const uint32_t env = P.env_base + (uint32_t)b;
            const float dwell = (float)P.pdt[b];
            const uint32_t site = 2u * grp;
            const size_t pix = (size_t)b * hw + site;
            float* px = P.signal + pix;
            const bool two = site + 1 < hw;
            const float d0 = dwell * (P.wdt ? P.wdt[pix] : 1.0f), d1 = dwell * ((P.wdt && two) ? P.wdt[pix + 1] : 1.0f);
            const float i0 = px[0], i1 = two ? px[1] : 0.0f;
            const float m0 = floorf(i0 * P.inv_e_photon) * (P.det_eff * d0) + P.background * d0;
            const float m1 = floorf(i1 * P.inv_e_photon) * (P.det_eff * d1) + P.background * d1;
            if (sample) {
                Philox4 c;
                c.x = grp; c.y = env; c.z = STED_RNG_PHOTON << 28; c.w = P.offset;
                const Philox4 r4 = philox4x32_10(c, (uint32_t)P.seed, (uint32_t)(P.seed >> 32));
                auto stage1 = [&](float mean, uint32_t w0, uint32_t w1, uint32_t at_pix, float* dst) {
                    float k;
                    if (poisson_fast(mean, w0, w1, &k)) { *dst = k; return; }
                    const int q = mean < 12.0f ? 0 : 1;
                    const unsigned at = atomicAdd(&q_tail[q], 1u);
                    const unsigned slot = at % DET_QCAP;
                    q_pix[q][slot] = at_pix; q_w0[q][slot] = w0; q_w1[q][slot] = w1;   // the intensity stays in *dst for the drain
                    full |= at + 1u - q_head[q] >= (unsigned)DET_THREADS;
                };
                stage1(m0, r4.x, r4.y, (uint32_t)pix, px);
                if (two) stage1(m1, r4.z, r4.w, (uint32_t)pix + 1u, px + 1);
            } else {
                px[0] = m0;
                if (two) px[1] = m1;
            }
        }
        if (__syncthreads_or(full)) drain(false);     // uniform decision; all pushes are visible after the barrier
    }
    drain(true);
}

// ------------------------------------------------------------------------------------------
// K3: raster scan with in-scan bleaching.
//
// pySTED visits dwells in row-major order; at each dwell it gathers the window, then thins every
// molecule of the window with survival exp(-k_b[u][v]*pdt).  Two facts make this
// parallel:
//  (1) survival factors commute, so within scan row r the count a dwell sees at tap (u,v) is the count
//      at the start of the row thinned by the taps v' > v of the same tap row ("row sweep");
//  (2) the thinning of a pixel depends only on its own count and on the hazard tables, never on the
//      image, so in stochastic mode every death can be drawn BEFORE the scan: each molecule lives an
//      Exp(1) amount of cumulative hazard along the (pass, tap) sequence that visits its pixel.
//
//   expectation mode : state s = D*exp(CH[u][hi+1]); gather with ET; s *= exp(CHm1 - CH[u][lo])
//   stochastic mode  : presample_kernel draws, per occupied pixel, Binomial(n, 1-exp(-H_tot)) deaths and
//                      places each one on its (scan row, tap) by inverting the cumulative hazard; a
//                      count / prefix-sum / fill pass buckets the events by scan row; sweep_kernel
//                      replays them: after the dense gather of row r, every event of row r decrements its
//                      pixel and removes the taps the dead molecule no longer contributes to.
//                      Identical in distribution to the per-dwell chain of binomials.
// sweep_kernel: CTA = (environment, tile of TC = 8*CT output columns), 256 threads; scan rows are
// sequential, the K-row band lives in a shared-memory ring; warp w owns CT columns, lane l the tap rows
// l and l+32; partial sums meet through a shuffle reduce-scatter.
// ------------------------------------------------------------------------------------------
constexpr int S_THREADS = 256;
#ifndef SWEEP_PROBE
#define SWEEP_PROBE 0   // 2: per-phase cycle counters (timing experiments only)
#endif

__host__ __device__ inline int sweep_pitch(int TC, int KP) {
    int p = TC + KP;
    while ((p & 31) != 28 && (p & 31) != 4 && (p & 31) != 12 && (p & 31) != 20) p += 4;
    return p;
}

template <int CT>
__device__ __forceinline__ float reduce_scatter(float (&a)[CT], int lane) {
    // after the loop lane holds the full sum of column (lane >> (5 - log2 CT))
    int n = CT;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        if (n > 1) {
            n >>= 1;
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int j = 0; j < CT / 2; ++j) {
                if (j < n) {
                    const float send = upper ? a[j] : a[j + n];
                    const float keep = upper ? a[j + n] : a[j];
                    a[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
        } else {
            a[0] += __shfl_xor_sync(0xffffffffu, a[0], off);
        }
    }
    return a[0];
}

// workspace of the stochastic scan:
//   [4 ints status][B*H ints cursor][B*H+1 (+3 pad) ints offsets][cap events][cap staged (key, event) pairs]
struct BleachWs {
    int* status;         // [0] = events dropped because the workspace was too small, [1] = staging cursor
    int* cnt;            // [B*H] events per (env, scan row); reused as fill cursor
    int* off;            // [B*H + 1] exclusive prefix sums
    uint32_t* ev;        // [cap] events  x | u<<12 | a<<18 | (scan row & 255)<<24, bucketed by (env, scan row)
    uint2* stage;        // [cap] (env*H + scan row, event) in generation order (fused presample)
    unsigned long long cap;
};

// Death schedule of every occupied pixel.
//   FILL = false: count the events of each (env, scan row);  FILL = true: regenerate them (same Philox
//   streams, same arithmetic) and store them in their row bucket.
// A CTA takes a chunk of PS_CHUNK pixels of one environment, compacts the occupied ones into work items of
// at most 8 molecules (independent sub-binomials add up to the pixel's binomial; a 150-molecule nanodomain
// becomes 19 items) and then gives one item to each thread, so lanes stay busy and evenly loaded.
// Cumulative pass hazards come from QQ[v][u] = sum_{u'>=u} CH[u'][v]: for a pixel whose taps are clipped to
// [lo,hi] the hazard of the passes u..u_hi is G(u) - G(u_hi+1) with G(u) = QQ[lo][u] - QQ[hi+1][u].
constexpr int PS_THREADS = 256, PS_CHUNK = 4096, PS_ITEMS = 2048, PS_PER = PS_CHUNK / PS_THREADS;
// MODE 0: count the events of each (env, scan row).  MODE 1: regenerate them and store them in their row bucket.
// MODE 2 (the default path): ONE pass — every death is drawn once, written with its bucket key to a staging area
// (the chunk reserves room for all its molecules with one atomic) and counted; scatter_events_kernel then moves the
// staged events into their buckets.  No second Philox pass.
template <int MODE>
__global__ void __launch_bounds__(PS_THREADS, 4) presample_kernel(const AcqParams P, const BleachWs ws) {
    const int K = P.K, KP = P.KP, H = P.H, W = P.W, Wp = P.Wp;
    extern __shared__ __align__(128) float smem[];
    float* sCH = smem;                                   // [K][KP]
    float* sQQ = sCH + K * KP;                           // [K+1][KP]
    uint32_t* items = reinterpret_cast<uint32_t*>(sQQ + (K + 1) * KP);      // [PS_ITEMS] pixel | sub<<12 | m<<24
    __shared__ int s_warp[PS_THREADS / 32];
    __shared__ int s_total;
    __shared__ unsigned long long s_stage_base;
    __shared__ unsigned s_stage_n, s_stage_cap;
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const float4* c4 = reinterpret_cast<const float4*>(P.tables + ((size_t)b * STED_TABLES + STED_TAB_CH) * K * KP);
        for (int i = tid; i < K * KP / 4; i += PS_THREADS) reinterpret_cast<float4*>(sCH)[i] = __ldg(c4 + i);
    }
    __syncthreads();
    if (tid <= K) {                                      // column v = tid of CH, suffix-summed over tap rows
        float q = 0.f;
        sQQ[tid * KP + K] = 0.f;
        for (int u = K - 1; u >= 0; --u) { q += (tid < K) ? sCH[u * KP + tid] : 0.f; sQQ[tid * KP + u] = q; }
    }
    __syncthreads();
    if (sQQ[0] <= 0.f) return;                           // no bleaching at all for this environment
    const int p = K / 2;
    const float* din = P.din + (size_t)b * P.Hp * P.Wpitch;
    const uint32_t env = P.env_base + (uint32_t)b;
    const int base = blockIdx.x * PS_CHUNK;

    // region holding molecules: the H x W region of interest, or the whole padded map with STED_FRAME_MOLECULES
    const bool whole = (P.flags & STED_FRAME_MOLECULES) != 0;
    const int RW = whole ? Wp : W, RH = whole ? P.Hp : H, ro = whole ? 0 : p;
    int cnt[PS_PER];
    int mine = 0, molecules = 0;
#pragma unroll
    for (int j = 0; j < PS_PER; ++j) {
        const int idx = base + j * PS_THREADS + tid;
        int n = 0;
        if (idx < RH * RW) n = (int)__ldg(din + (size_t)(ro + idx / RW) * P.Wpitch + ro + idx % RW);
        if (n > 32767) { n = 32767; atomicAdd(&ws.status[0], 1); }   // 12-bit sub-stream index x 8 molecules: reported as dropped
        cnt[j] = n > 0 ? n : 0;
        mine += (cnt[j] + 7) >> 3;
        molecules += cnt[j];
    }
    if (MODE == 2) {                                     // room for every molecule of this chunk, one atomic
        for (int o = 16; o; o >>= 1) molecules += __shfl_xor_sync(0xffffffffu, molecules, o);
        if (tid == 0) s_stage_n = 0;
        __syncthreads();
        if (lane == 0 && molecules) atomicAdd(&s_stage_n, (unsigned)molecules);
        __syncthreads();
        if (tid == 0) {
            const unsigned want = s_stage_n;
            const unsigned long long at = want ? atomicAdd(reinterpret_cast<unsigned long long*>(ws.status + 2), (unsigned long long)want) : 0ull;
            s_stage_base = at;
            s_stage_cap = at >= ws.cap ? 0u : (unsigned)min((unsigned long long)want, ws.cap - at);
            s_stage_n = 0;
        }
        __syncthreads();
    }
    // exclusive scan of the per-thread item counts over the CTA
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int w = 0; w < PS_THREADS / 32; ++w) { const int t = s_warp[w]; s_warp[w] = run; run += t; }
        s_total = run;
    }
    __syncthreads();
    const int first = s_warp[warp] + incl - mine;        // index of this thread's first item
    const int total = s_total;

    for (int win = 0; win < total; win += PS_ITEMS) {    // rounds of at most PS_ITEMS items
        int g = first;
#pragma unroll
        for (int j = 0; j < PS_PER; ++j) {
            const int subs = (cnt[j] + 7) >> 3;
            for (int s = 0; s < subs; ++s, ++g) {
                if (g >= win && g < win + PS_ITEMS) {
                    const int m = min(8, cnt[j] - 8 * s);
                    items[g - win] = (uint32_t)(j * PS_THREADS + tid) | ((uint32_t)min(s, 4095) << 12) | ((uint32_t)m << 24);
                }
            }
        }
        __syncthreads();
        const int n_items = min(PS_ITEMS, total - win);
        for (int q = tid; q < n_items; q += PS_THREADS) {
            const uint32_t it = items[q];
            const int idx = base + (int)(it & 0xFFFu), sub = (it >> 12) & 0xFFF, m = it >> 24;
            const int y = ro + idx / RW, x = ro + idx % RW;
            const int lo = max(0, x - W + 1), hi = min(K - 1, x);        // taps that ever visit this column
            const int u_hi = min(K - 1, y), u_lo = max(0, y - (H - 1));   // passes, visited from u_hi down to u_lo
            const float* qa = sQQ + lo * KP;
            const float* qb = sQQ + (hi + 1) * KP;
            const float g_end = qa[u_hi + 1] - qb[u_hi + 1];
            const float htot = (qa[u_lo] - qb[u_lo]) - g_end;
            PhiloxStream rng;
            rng.init(P.seed, P.offset, env, (uint32_t)(y * Wp + x), STED_RNG_BLEACH, (uint32_t)sub);
            const int dead = binomial_deaths(m, htot, rng);
            const float pd = -expm1f(-htot);
            for (int d = 0; d < dead; ++d) {
                // hazard this molecule had consumed when it died: truncated Exp(1) on (0, htot)
                const float t = fminf(-log1pf(-rng.uniform() * pd), htot);
                const float thr = t + g_end;
                int a = u_lo, z = u_hi;                  // largest u in [u_lo,u_hi] with G(u) >= thr
                while (a < z) {
                    const int mid = (a + z + 1) >> 1;
                    if (qa[mid] - qb[mid] >= thr) a = mid; else z = mid - 1;
                }
                const int u = a, r = y - u;
                if (MODE == 0) {
                    atomicAdd(&ws.cnt[(size_t)b * H + r], 1);
                } else {
                    // tap: largest v in [lo,hi] whose suffix hazard still covers the in-pass position
                    const float before = (qa[u + 1] - qb[u + 1]) - g_end;
                    const float* ch = sCH + u * KP;
                    const float thr2 = fmaxf(t - before, 0.f) + ch[hi + 1];
                    int a2 = lo, z2 = hi;
                    while (a2 < z2) {
                        const int mid = (a2 + z2 + 1) >> 1;
                        if (ch[mid] >= thr2) a2 = mid; else z2 = mid - 1;
                    }
                    const uint32_t ev = (uint32_t)x | ((uint32_t)u << 12) | ((uint32_t)a2 << 18) | ((uint32_t)(r & 255) << 24);
                    if (MODE == 1) {
                        const unsigned long long slot = (unsigned long long)ws.off[(size_t)b * H + r] +
                                                        (unsigned long long)atomicAdd(&ws.cnt[(size_t)b * H + r], 1);
                        if (slot < ws.cap) ws.ev[slot] = ev;
                        else atomicAdd(&ws.status[0], 1);
                    } else {
                        const unsigned at = atomicAdd(&s_stage_n, 1u);
                        if (at < s_stage_cap) {
                            ws.stage[s_stage_base + at] = make_uint2((uint32_t)(b * H + r), ev);
                            atomicAdd(&ws.cnt[(size_t)b * H + r], 1);
                        } else {
                            atomicAdd(&ws.status[0], 1);
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    if (MODE == 2) {                                     // unused tail of this chunk's reservation: mark it empty
        const unsigned used = min(s_stage_n, s_stage_cap);
        for (unsigned i = used + tid; i < s_stage_cap; i += PS_THREADS) ws.stage[s_stage_base + i] = make_uint2(0xFFFFFFFFu, 0u);
    }
}

// staged (key, event) pairs -> row buckets (counting sort, third step; off[] holds the bucket starts, cnt[] the cursors)
__global__ void __launch_bounds__(256) scatter_events_kernel(const BleachWs ws) {
    unsigned long long n = *reinterpret_cast<const unsigned long long*>(ws.status + 2);
    if (n > ws.cap) n = ws.cap;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint2 e = ws.stage[i];
        if (e.x == 0xFFFFFFFFu) continue;
        const unsigned long long slot = (unsigned long long)ws.off[e.x] + (unsigned long long)atomicAdd(&ws.cnt[e.x], 1);
        ws.ev[slot] = e.y;
    }
}

// exclusive prefix sum of cnt[n] into off[n+1] (single CTA), then cnt is cleared to serve as fill cursor
__global__ void __launch_bounds__(1024) bucket_scan_kernel(const BleachWs ws, const int n) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        const int v = i < n ? ws.cnt[i] : 0;
        int s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
        if (lane == 31) warp_tot[warp] = s;
        __syncthreads();
        if (warp == 0) {
            int w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
            warp_tot[lane] = w;
        }
        __syncthreads();
        const int excl = carry + (warp ? warp_tot[warp - 1] : 0) + s - v;
        if (i < n) { ws.off[i] = excl; ws.cnt[i] = 0; }
        __syncthreads();
        if (tid == 1023) carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) ws.off[n] = carry;
}

template <int KT, int CT, bool STOCH>
__global__ void __launch_bounds__(S_THREADS, STOCH ? 3 : 2) sweep_kernel(const AcqParams P, const BleachWs ws) {
    constexpr int TC = 8 * CT;
    const int K = KT ? KT : P.K;
    const int KP = KT ? STED_KPITCH(KT) : P.KP;
    const int PR = sweep_pitch(TC, KP);
    const int span = TC + K - 1;

    extern __shared__ __align__(128) float smem[];
    float* ring = smem;                       // [K][PR]  band of the molecule map (counts / scaled expectations)
    float* sWt = ring + K * PR;               // [K][KP]  gather weights: E (stochastic) or ET (expectation)
    float* sI = sWt + K * KP;                 // [2][TC]  dense row intensities (even / odd scan rows)
    float* sScale = sI + 2 * TC;              // [2][TC]  2^30 / sI  (fixed-point scale of the corrections)
    int* sCorr = reinterpret_cast<int*>(sScale + 2 * TC);    // [2][TC] sum of removed taps, fixed point
    float* sCH = reinterpret_cast<float*>(sCorr + 2 * TC);   // [K][KP]  suffix hazards     (expectation mode only)
    float* sRS = sCH + K * KP;                // [K][KP]  exp(-CH)                           (expectation mode only)

    const int b = blockIdx.y;
    const int c0 = blockIdx.x * TC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = P.W, H = P.H, Wp = P.Wp;
    const float* din = P.din + (size_t)b * P.Hp * P.Wpitch;
    float* dout = P.dout + (size_t)b * P.Hp * P.Wpitch;

    {   // tables
        const size_t tsz = (size_t)K * KP;
        const float* tb = P.tables + (size_t)b * STED_TABLES * tsz;
        const float4* w4 = reinterpret_cast<const float4*>(tb + (STOCH ? STED_TAB_E : STED_TAB_ET) * tsz);
        const float4* c4 = reinterpret_cast<const float4*>(tb + STED_TAB_CH * tsz);
        const float4* r4 = reinterpret_cast<const float4*>(tb + STED_TAB_RS * tsz);
        for (int i = tid; i < (int)(tsz / 4); i += S_THREADS) {
            reinterpret_cast<float4*>(sWt)[i] = __ldg(w4 + i);
            if (!STOCH) {
                reinterpret_cast<float4*>(sCH)[i] = __ldg(c4 + i);
                reinterpret_cast<float4*>(sRS)[i] = __ldg(r4 + i);
            }
        }
    }
    if (tid < 2 * TC) sCorr[tid] = 0;
    __syncthreads();

    // columns of the molecule map this tile writes back (each column has exactly one owner)
    const int p = K / 2;
    const int xs = (c0 == 0) ? 0 : c0 + p;
    const int xe = (c0 + TC >= W) ? P.Wpitch : c0 + TC + p;

    // Retire band row y_old to global memory (expectation mode undoes the s = D*exp(CH[u][hi+1]) scaling of
    // the left-clipped columns) and install the prefetched row y_new in the same ring slot.  The same
    // thread owns a group of four columns in both halves, so no barrier is needed in between.
    auto cycle_row = [&](const int y_old, const int u_scale_old, const int y_new, const int u_entry, const float4 pre) {
        if (tid < PR / 4) {
            const int xl = 4 * tid, x = c0 + xl;
            if (y_old >= 0) {
                const float4 old = *reinterpret_cast<const float4*>(ring + (y_old % K) * PR + xl);
                const float* po_ = reinterpret_cast<const float*>(&old);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int xx = x + j;
                    if (xx >= xs && xx < xe) {
                        float val = po_[j];
                        if (!STOCH && u_scale_old >= 0 && xx < K - 1 && val != 0.f)
                            val *= expf(-sCH[u_scale_old * KP + xx + 1]);
                        dout[(size_t)y_old * P.Wpitch + xx] = val;
                    }
                }
            }
            if (y_new >= 0) {
                float4 val = pre;
                if (!STOCH) {
                    float* pv = reinterpret_cast<float*>(&val);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (x + j < K - 1 && pv[j] != 0.f) pv[j] *= expf(sCH[u_entry * KP + x + j + 1]);
                }
                *reinterpret_cast<float4*>(ring + (y_new % K) * PR + xl) = val;
            }
        }
    };
    auto fetch_row = [&](const int y) -> float4 {
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        const int x = c0 + 4 * tid;
        if (tid < PR / 4 && y >= 0 && y < P.Hp && x < P.Wpitch)
            val = __ldg(reinterpret_cast<const float4*>(din + (size_t)y * P.Wpitch + x));
        return val;
    };
    // writes scan row r (dense gather minus the taps removed by the deaths of that row)
    auto write_row = [&](const int r) {
        if (tid < TC && c0 + tid < W) {
            const int par = r & 1;
            float I = sI[par * TC + tid];
            if (STOCH) {
                I = fmaxf(I * (1.0f - (float)sCorr[par * TC + tid] * 9.313225746154785e-10f), 0.f);
                sCorr[par * TC + tid] = 0;
            }
            const size_t o = ((size_t)b * H + r) * W + c0 + tid;
            if (P.intensity) P.intensity[o] = I;
            P.signal[o] = I;                              // detect_kernel converts to photon counts
        }
    };

    for (int y = 0; y < K; ++y) cycle_row(-1, -1, y, y, fetch_row(y));
    __syncthreads();

#if SWEEP_PROBE == 2
    long long t_ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t_last = clock64();
#define PHASE_MARK(i) do { const long long t_now = clock64(); t_ph[i] += t_now - t_last; t_last = t_now; } while (0)
#else
#define PHASE_MARK(i) do { } while (0)
#endif
    const int* row_off = STOCH ? ws.off + (size_t)b * H : nullptr;
    for (int r = 0; r < H; ++r) {
        const int par = r & 1;
        const float4 pre = fetch_row(r + K < P.Hp ? r + K : -1);      // lands while the row is gathered
        // events of this scan row: the first one of each thread is prefetched too
        constexpr int EV_LANES = 4, EV_SLOTS = S_THREADS / EV_LANES, EV_PRE = 4;
        int ev_lo = 0, ev_hi = 0;
        uint32_t evp[EV_PRE] = {0u, 0u, 0u, 0u};
        if (STOCH) {
            ev_lo = __ldg(row_off + r); ev_hi = __ldg(row_off + r + 1);
            if ((unsigned long long)ev_hi > ws.cap) ev_hi = (int)ws.cap;
#pragma unroll
            for (int i = 0; i < EV_PRE; ++i) {
                const int q = ev_lo + tid / EV_LANES + i * EV_SLOTS;
                if (q < ev_hi) evp[i] = __ldg(ws.ev + q);
            }
            if (r > 0) write_row(r - 1);
        }
        // ---------------- phase A: dense gather of scan row r against the row-start state
        {
            float acc[CT];
#pragma unroll
            for (int j = 0; j < CT; ++j) acc[j] = 0.f;
            for (int uu = lane; uu < K; uu += 32) {
                const float* drow = ring + ((r + uu) % K) * PR + CT * warp;
                const float* erow = sWt + uu * KP;
                float win[CT + 4];
#pragma unroll
                for (int j = 0; j < CT; j += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(drow + j);
                    win[j] = t.x; win[j + 1] = t.y; win[j + 2] = t.z; win[j + 3] = t.w;
                }
                const int nchunk = KP / 4;
#pragma unroll 5
                for (int vc = 0; vc < nchunk; ++vc) {
                    const float4 nw = *reinterpret_cast<const float4*>(drow + CT + 4 * vc);
                    const float4 e = *reinterpret_cast<const float4*>(erow + 4 * vc);
                    win[CT] = nw.x; win[CT + 1] = nw.y; win[CT + 2] = nw.z; win[CT + 3] = nw.w;
#pragma unroll
                    for (int j = 0; j < CT; ++j) {
                        acc[j] = fmaf(e.x, win[j], acc[j]);
                        acc[j] = fmaf(e.y, win[j + 1], acc[j]);
                        acc[j] = fmaf(e.z, win[j + 2], acc[j]);
                        acc[j] = fmaf(e.w, win[j + 3], acc[j]);
                    }
#pragma unroll
                    for (int j = 0; j < CT; ++j) win[j] = win[j + 4];
                }
            }
            const float tot = reduce_scatter<CT>(acc, lane);
            constexpr int SH = (CT == 32) ? 0 : (CT == 16) ? 1 : (CT == 8) ? 2 : (CT == 4) ? 3 : 4;
            if ((lane & ((1 << SH) - 1)) == 0) {
                sI[par * TC + CT * warp + (lane >> SH)] = tot;
                if (STOCH) sScale[par * TC + CT * warp + (lane >> SH)] = 1073741824.0f / fmaxf(tot, 1e-30f);
            }
        }
        PHASE_MARK(0);
        __syncthreads();
        PHASE_MARK(1);

        // ---------------- phase B: thin the band with the taps of this scan row
        if (!STOCH) {
            for (int u = warp; u < K; u += S_THREADS / 32) {
                float* row = ring + ((r + u) % K) * PR;
                for (int xl = lane; xl < span; xl += 32) {
                    const float n = row[xl];
                    if (n == 0.f) continue;
                    const int x = c0 + xl;
                    if (x >= Wp) continue;
                    const int lo = max(0, x - W + 1), hi = min(K - 1, x);
                    float m;
                    if (hi == K - 1) m = sRS[u * KP + lo];
                    else m = expf((u > 0 ? sCH[(u - 1) * KP + hi + 1] : 0.f) - sCH[u * KP + lo]);
                    row[xl] = n * m;
                }
            }
            __syncthreads();
            write_row(r);
        } else {
            // replay the pre-drawn deaths of this scan row, one event per thread: the pixel loses a molecule
            // and the dwells at taps lo .. a-1 of tap row u no longer see it.  Corrections are fixed-point
            // integer atomics (native in shared memory; the sum does not depend on the order).
            // Four lanes share an event and split its taps.
            const float* sc = sScale + par * TC;
            int* corr = sCorr + par * TC;
            const int sub = tid % EV_LANES;
            int it = 0;
            for (int q = ev_lo + tid / EV_LANES; q < ev_hi; q += EV_SLOTS, ++it) {
                uint32_t ev;
                if (it == 0) ev = evp[0]; else if (it == 1) ev = evp[1]; else if (it == 2) ev = evp[2];
                else if (it == 3) ev = evp[3]; else ev = __ldg(ws.ev + q);
                const int xl = (int)(ev & 0xFFFu) - c0, u = (ev >> 12) & 0x3F, a = (ev >> 18) & 0x3F;
                if (xl < 0 || xl >= span) continue;                       // another tile's pixel
                if (sub == 0) atomicAdd(&ring[((r + u) % K) * PR + xl], -1.0f);
                const float* e = sWt + u * KP;
                const int v0 = max(max(0, c0 + xl - W + 1), xl - (TC - 1)), v1 = min(a, xl + 1);   // 0 <= xl - v < TC
#pragma unroll 4
                for (int v = v0 + sub; v < v1; v += EV_LANES) atomicAdd(&corr[xl - v], __float2int_rn(e[v] * sc[xl - v]));
            }
            PHASE_MARK(2);
            __syncthreads();
        }
        PHASE_MARK(3);
        // ---------------- retire band row y = r, install row r + K
        cycle_row(r, -1, (r + K < P.Hp) ? r + K : -1, K - 1, pre);
        __syncthreads();
        PHASE_MARK(4);
    }
#if SWEEP_PROBE == 2
    if (tid == 0) for (int i = 0; i < 8; ++i) atomicAdd(&g_phase_cycles[i], (unsigned long long)t_ph[i]);
#endif
    if (STOCH) write_row(H - 1);
    // flush the K-1 band rows below the last scan row (tap row they would have had next: y - H)
    for (int y = H; y < P.Hp; ++y) cycle_row(y, y - H, -1, 0, make_float4(0.f, 0.f, 0.f, 0.f));
}

// ------------------------------------------------------------------------------------------
// K3b: the stochastic scan, second generation (sweep2_kernel).
//
// Same arithmetic as sweep_kernel<.., true> — dense gather of scan rows against known counts, replay of the pre-drawn
// deaths — reorganised around the measurements of round 1 (profiles/r01_summary.md) and of round 2
// (profiles/r02_summary.md):
//  * the grid was 512 CTAs on 444 slots (the second wave 15 % full): work items are now (environment, column strip,
//    RANGE OF SCAN ROWS).  A range can start anywhere because the state of the map at scan row r0 is known in advance:
//    the initial counts minus the deaths of the K-1 scan rows before r0 that hit band rows >= r0 (events carry the low
//    byte of their scan row for this).  Several times as many, shorter CTAs leave a tail of a few percent;
//  * the row loop was latency bound (three CTA-wide barriers and a replay phase per scan row; barrier stalls lead the
//    ncu stall list): scan rows are processed in GROUPS of T.  The T gathers of a group run back to back, without a
//    barrier, against the counts at the START of the group; the deaths of the group then leave the ring in one phase
//    (two barriers per group, not per row).  What a later row of the group saw too much — the molecules that died in
//    the earlier rows of the group — is removed with the intensity corrections, which already existed for the taps a
//    death removes from its own row: a death costs (T-1)/2 * K more fixed-point atomics on average;
//  * those corrections touch nothing a gather reads, so two auxiliary warps apply them WHILE the eight gather warps
//    work on the next group; the auxiliary warps also move the ring: the rows retired by the previous group leave for
//    global memory and the rows the next group needs take their slots (ring of K + 2T - 1 rows).
// ------------------------------------------------------------------------------------------
#ifndef S2_FFMA2
#define S2_FFMA2 0   // 1: packed FFMA2 gather loop (measured equal at best while the row loop is latency bound, see profiles/r02_summary.md)
#endif
#ifndef S2_AUX_THREADS
#define S2_AUX_THREADS 64
#endif
#ifndef S2_WRITE_BY_GATHER
#define S2_WRITE_BY_GATHER 1   // 1: the gather warps write the previous group's intensities out (0: the auxiliary warps)
#endif
#ifndef S2_PARTS
#define S2_PARTS 3     // work units of the compacted gather per live tap row (runs of tap chunks)
#endif
#ifndef S2_COMPACT
#define S2_COMPACT 1   // 1: a gather warp only visits the tap rows whose ring row holds a molecule inside its window (see the kernel)
#endif
constexpr int S2_GATHER = 256, S2_AUX = S2_AUX_THREADS, S2_THREADS = S2_GATHER + S2_AUX;
__host__ __device__ constexpr int s2_evcap(int T) { return T <= 2 ? 1024 : T == 3 ? 512 : 256; }

template <int KT, int CT, int T>
__global__ void __launch_bounds__(S2_THREADS, 3) sweep2_kernel(const AcqParams P, const BleachWs ws, const int RR) {
    constexpr int TC = 8 * CT;
    const int K = KT ? KT : P.K;
    const int KP = KT ? STED_KPITCH(KT) : P.KP;
    const int NS = K + 2 * T - 1;             // ring slots
    const int PR = sweep_pitch(TC, KP);
    const int span = TC + K - 1;

    extern __shared__ __align__(128) float smem[];
    float* ring = smem;                       // [K+2T-1][PR] band of the molecule map (integer counts)
    float* sWt = ring + NS * PR;              // [K][KP]    effective PSF E
    float* sI = sWt + K * KP;                 // [2][T][TC] dense row intensities (even / odd groups)
    float* sScale = sI + 2 * T * TC;          // [T][TC]    2^30 / sI  (fixed-point scale of the corrections)
    int* sCorr = reinterpret_cast<int*>(sScale + T * TC);    // [T][TC] sum of removed taps, fixed point
    constexpr int EVCAP = s2_evcap(T);        // staged events per group (the rest is read from global memory)
    uint32_t* sEv = reinterpret_cast<uint32_t*>(sCorr + T * TC);            // [2][EVCAP] this / the previous group's events
    int* sMeta = reinterpret_cast<int*>(sEv + 2 * EVCAP);                   // [2][2] their range in ws.ev
    // [NS][2] per ring slot: which 4-column chunks of the row held a molecule BEFORE the scan (bit g = chunk g).  Molecule
    // maps are sparse; a gather warp (CT columns, K tap rows spread over its lanes) visits only the tap rows whose ring
    // row has a molecule inside the warp's window: the live rows are dealt to the lanes in order, and when there are at
    // most 32 of them (47 % of the (warp, scan row) pairs of the bench maps) one pass over the taps replaces two.  The
    // marks come from the map as it was before this acquisition - never from the counts of the moment - so the rows a warp
    // visits, and with them the order of its sums, do not depend on how the scan is cut in row ranges.
    uint32_t* sOcc = reinterpret_cast<uint32_t*>(sMeta + 4);
    unsigned char* sRows = reinterpret_cast<unsigned char*>(sOcc + 2 * NS);      // [8 gather warps][64] live tap rows, in order

    const int b = blockIdx.z;
    const int c0 = blockIdx.x * TC;
    const int r0 = blockIdx.y * RR;
    const int r1 = min(P.H, r0 + RR);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool aux = tid >= S2_GATHER;
    const int at = tid - S2_GATHER;           // index among the auxiliary threads
    const int W = P.W, H = P.H;
    const float* din = P.din + (size_t)b * P.Hp * P.Wpitch;
    float* dout = P.dout + (size_t)b * P.Hp * P.Wpitch;
    const int* row_off = ws.off + (size_t)b * H;

    {
        const float4* w4 = reinterpret_cast<const float4*>(P.tables + ((size_t)b * STED_TABLES + STED_TAB_E) * K * KP);
        for (int i = tid; i < K * KP / 4; i += S2_THREADS) reinterpret_cast<float4*>(sWt)[i] = __ldg(w4 + i);
    }
    if (tid < T * TC) sCorr[tid] = 0;
    if (T * TC > S2_THREADS)
        for (int i = tid + S2_THREADS; i < T * TC; i += S2_THREADS) sCorr[i] = 0;

    // columns of the molecule map this strip writes back (each column has exactly one owner)
    const int p = K / 2;
    const int xs = (c0 == 0) ? 0 : c0 + p;
    const int xe = (c0 + TC >= W) ? P.Wpitch : c0 + TC + p;
    const int q4 = PR / 4;                    // float4 groups of a ring row

    auto slot_of = [&](const int y) { return (KT ? y % (KT + 2 * T - 1) : y % NS) * PR; };
    auto slot_idx = [&](const int y) { return KT ? y % (KT + 2 * T - 1) : y % NS; };
    auto mark_row = [&](const int y, const int g, const float4 v) {
        if (S2_COMPACT && (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)) atomicOr(&sOcc[2 * slot_idx(y) + (g >> 5)], 1u << (g & 31));
    };
    auto fetch_row = [&](const int y, const int g) -> float4 {
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        const int x = c0 + 4 * g;
        if (y >= 0 && y < P.Hp && x < P.Wpitch) val = __ldg(reinterpret_cast<const float4*>(din + (size_t)y * P.Wpitch + x));
        return val;
    };
    auto retire_row = [&](const int y, const int g) {
        const float4 old = *reinterpret_cast<const float4*>(ring + slot_of(y) + 4 * g);
        const float* po_ = reinterpret_cast<const float*>(&old);
        const int x = c0 + 4 * g;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (x + j >= xs && x + j < xe) dout[(size_t)y * P.Wpitch + x + j] = po_[j];
    };
    auto clamp_hi = [&](const int v) { return (unsigned long long)v > ws.cap ? (int)ws.cap : v; };

    // band rows r0 .. r0+K+T-2 (what the first group reads) as they were BEFORE the scan ...
    if (S2_COMPACT) {
        for (int i = tid; i < 2 * NS; i += S2_THREADS) sOcc[i] = 0u;
        __syncthreads();
    }
    for (int i = tid; i < (K + T - 1) * q4; i += S2_THREADS) {
        const int yy = i / q4, g = i - yy * q4;
        const float4 v = fetch_row(r0 + yy, g);
        *reinterpret_cast<float4*>(ring + slot_of(r0 + yy) + 4 * g) = v;
        mark_row(r0 + yy, g, v);
    }
    __syncthreads();
    // ... minus the deaths of the scan rows before r0 that hit them (scan row r' reaches band rows r' .. r'+K-1)
    if (r0 > 0) {
        const int rp = max(0, r0 - (K - 1));
        const int lo = __ldg(row_off + rp), hi = clamp_hi(__ldg(row_off + r0));
        for (int q = lo + tid; q < hi; q += S2_THREADS) {
            const uint32_t ev = __ldg(ws.ev + q);
            const int xl = (int)(ev & 0xFFFu) - c0, u = (ev >> 12) & 0x3F;
            const int rr = r0 - 1 - (int)(((uint32_t)(r0 - 1) - (ev >> 24)) & 255u);     // scan row of the event
            const int y = rr + u;
            if (y >= r0 && xl >= 0 && xl < span) atomicAdd(&ring[slot_of(y) + xl], -1.0f);
        }
        __syncthreads();
    }

    // intensities of the group of scan rows R0 .. R0+Tg-1: dense gathers against the counts at the start of the group,
    // minus (a) the taps a death removes from the rest of its own row and (b) everything the molecules that died in
    // an earlier row of the group would have contributed to the later rows of the group: fixed-point sums in sCorr
    // (filled between the two barriers of the group), applied here when the auxiliary warps write the rows out
    auto write_rows = [&](const int R0, const int Tg, const int par, const int first, const int stride) {
        const float* dense = sI + par * T * TC;
        for (int t = first; t < Tg * TC; t += stride) {
            const int j = t / TC, c = t - j * TC;
            if (c0 + c < W) {
                const float I = fmaxf(dense[t] * (1.0f - (float)sCorr[t] * 9.313225746154785e-10f), 0.f);
                const size_t o = ((size_t)b * H + R0 + j) * W + c0 + c;
                if (P.intensity) P.intensity[o] = I;
                P.signal[o] = I;                          // detect_kernel converts to photon counts
            }
            sCorr[t] = 0;
        }
    };

#if SWEEP_PROBE >= 3
    long long s2_passes = 0, s2_rows = 0;
    long long t_ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t_a = clock64(), t_b;
    const long long t_begin = t_a;
#define S2_TICK(i) do { t_b = clock64(); t_ph[i] += t_b - t_a; t_a = t_b; } while (0)
#else
#define S2_TICK(i) do { } while (0)
#endif
    int grp = 0;
    for (int R0 = r0; R0 < r1; R0 += T, ++grp) {
        const int Tg = min(T, r1 - R0);
        const int par = grp & 1;
        S2_TICK(7);
        if (aux) {
            // ---- auxiliary warps: the rows the previous group finished leave the ring, the rows the NEXT group needs
            // take their slots (nothing of this group touches them); then the previous group's intensities
            const int ev_lo = __ldg(row_off + R0), ev_hi = clamp_hi(__ldg(row_off + R0 + Tg));
            {   // this group's events -> shared memory (read by every thread between the two barriers and by these warps
                // during the next group): loads first, stores after, so that their latencies overlap
                uint32_t* sev = sEv + par * EVCAP;
                const int ns = min(ev_hi - ev_lo, EVCAP);
                for (int i0 = 0; i0 < ns; i0 += 4 * S2_AUX) {
                    uint32_t t4[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) t4[k] = (i0 + k * S2_AUX + at < ns) ? __ldg(ws.ev + ev_lo + i0 + k * S2_AUX + at) : 0u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) if (i0 + k * S2_AUX + at < ns) sev[i0 + k * S2_AUX + at] = t4[k];
                }
                if (at == 0) { sMeta[2 * par] = ev_lo; sMeta[2 * par + 1] = ev_hi; }
            }
#ifdef S2_ABLATE_RING                        // timing ablation only: the ring never moves
            if (false)
#endif
            if (S2_COMPACT) {                    // the marks of the rows that leave (their slots are not read by this group)
                if (at < T) { const int sl = slot_idx(R0 - T + at + NS); sOcc[2 * sl] = 0u; sOcc[2 * sl + 1] = 0u; }
                asm volatile("bar.sync 1, %0;" :: "n"(S2_AUX) : "memory");
            }
            for (int i = at; i < T * q4; i += S2_AUX) {
                const int j = i / q4, g = i - j * q4;
                const int y_old = R0 - T + j, y_new = y_old + NS;
                if (grp > 0) retire_row(y_old, g);
                const float4 v = fetch_row(y_new, g);                                                 // zeros below the map
                *reinterpret_cast<float4*>(ring + slot_of(y_new) + 4 * g) = v;
                mark_row(y_new, g, v);
            }
            S2_TICK(3);
#if !S2_WRITE_BY_GATHER
            if (grp > 0) write_rows(R0 - T, T, par ^ 1, at, S2_AUX);
#endif
            S2_TICK(4);
        } else {
#if S2_WRITE_BY_GATHER
            // the previous group's intensities leave first (the auxiliary warps' chain of dependent loads - event list, ring
            // rows - is as long as a gather since the gather skips empty tap rows: the write-out moved over here)
            if (grp > 0) write_rows(R0 - T, T, par ^ 1, tid, S2_GATHER);
#endif
            // ---- gather warps: dense gathers of scan rows R0 .. R0+Tg-1 against the counts at the start of the group
            for (int j = 0; j < Tg; ++j) {
                float acc[CT];
#if S2_COMPACT
                // the tap rows with a molecule inside this warp's window (CT + KP columns from column CT * warp), dealt to
                // the lanes in order
                const unsigned long long wmask = ((1ull << ((CT + KP) >> 2)) - 1ull) << ((CT * warp) >> 2);
                auto live = [&](const int uu) {
                    const uint32_t* o = sOcc + 2 * slot_idx(R0 + j + uu);
                    return uu < K && ((((unsigned long long)o[1] << 32) | o[0]) & wmask) != 0ull;
                };
                const bool mine0 = live(lane), mine1 = live(lane + 32);
                const unsigned live0 = __ballot_sync(0xffffffffu, mine0), live1 = __ballot_sync(0xffffffffu, mine1);
                const int n0 = __popc(live0), nlive = n0 + __popc(live1);
                // each live row writes itself to its rank in the warp's list (__fns, the direct "n-th set bit", is a
                // software loop that cost as much as the skipped rows saved)
                unsigned char* my_rows = sRows + 64 * warp;
                if (mine0) my_rows[__popc(live0 & ((1u << lane) - 1u))] = (unsigned char)lane;
                if (mine1) my_rows[n0 + __popc(live1 & ((1u << lane) - 1u))] = (unsigned char)(lane + 32);
                __syncwarp();
#if SWEEP_PROBE == 3 && defined(S2_COUNT_PASSES)
                if (tid == 0) { s2_passes += (nlive + 31) / 32; s2_rows += 1; }      // warp 0's view
#endif
                // work units = (live tap row, one of S2_PARTS runs of tap chunks): with whole rows as units a warp with 33
                // live rows paid two full passes; the finer deal wastes at most one part per lane (1.17 passes per scan row
                // on the bench maps with three parts, 1.53 with one).  Testing every (row, run) unit against the row's marks
                // (0.93 pass-equivalents) measured SLOWER, 3.149 vs 3.085 ms for the whole scan: six ballots and rank stores
                // per scan row instead of two, and warps of one CTA whose unit counts differ more wait longer for each other.
                const int nchunk = KP / 4, per = (nchunk + S2_PARTS - 1) / S2_PARTS;
#endif
#if S2_FFMA2
                // FFMA2 (fma.rn.f32x2), one issue slot per two FMAs: the ncu profile of the scalar loop was issue bound
                // (83 % of the issue slots for 65 % of the FMA pipe).  A tap E[u][v] — the instruction's scalar-broadcast
                // operand — meets the datamap values of two neighbouring columns, which must be an even-aligned register
                // pair of the sliding window: true for the column pairs (2p, 2p+1) when v is even and for the pairs
                // (2q-1, 2q) when v is odd.  Hence two sets of packed accumulators, one per tap parity (the outer halves
                // of the odd set, columns -1 and CT, are discarded: 1/16 more FMA work); a column's sum is even + odd.
                u64 accA[CT / 2], accB[CT / 2 + 1];
#pragma unroll
                for (int pp = 0; pp < CT / 2; ++pp) accA[pp] = 0ull;
#pragma unroll
                for (int pp = 0; pp <= CT / 2; ++pp) accB[pp] = 0ull;
#if S2_COMPACT
                for (int iu = lane; iu < S2_PARTS * nlive; iu += 32) {
                    const int uu = my_rows[iu / S2_PARTS], vc0 = (iu % S2_PARTS) * per, nvc = min(nchunk, vc0 + per) - vc0;
#else
                for (int uu = lane; uu < K; uu += 32) {
                    const int vc0 = 0, nvc = KP / 4;
#endif
                    const float* drow = ring + slot_of(R0 + j + uu) + CT * warp + 4 * vc0;
                    const float* erow = sWt + uu * KP + 4 * vc0;
                    u64 wp[CT / 2 + 2];                               // window pairs (w[2k], w[2k+1])
#pragma unroll
                    for (int jj = 0; jj < CT / 4; ++jj) {
                        const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(drow + 4 * jj);
                        wp[2 * jj] = t.x; wp[2 * jj + 1] = t.y;
                    }
#pragma unroll 5
                    for (int vc = 0; vc < nvc; ++vc) {
                        const ulonglong2 nw = *reinterpret_cast<const ulonglong2*>(drow + CT + 4 * vc);
                        const float4 e = *reinterpret_cast<const float4*>(erow + 4 * vc);
                        wp[CT / 2] = nw.x; wp[CT / 2 + 1] = nw.y;
                        const u64 e0 = pack2(e.x, e.x), e1 = pack2(e.y, e.y), e2 = pack2(e.z, e.z), e3 = pack2(e.w, e.w);
#pragma unroll
                        for (int pp = 0; pp < CT / 2; ++pp) {
                            accA[pp] = fma2(e0, wp[pp], accA[pp]);
                            accA[pp] = fma2(e2, wp[pp + 1], accA[pp]);
                        }
#pragma unroll
                        for (int pp = 0; pp <= CT / 2; ++pp) {
                            accB[pp] = fma2(e1, wp[pp], accB[pp]);
                            accB[pp] = fma2(e3, wp[pp + 1], accB[pp]);
                        }
#pragma unroll
                        for (int jj = 0; jj < CT / 2; ++jj) wp[jj] = wp[jj + 2];
                    }
                }
#pragma unroll
                for (int pp = 0; pp < CT / 2; ++pp) {
                    float a0, a1, b0, b1, c0_, c1;
                    unpack2(accA[pp], a0, a1);
                    unpack2(accB[pp], b0, b1);            // (column 2pp-1, column 2pp)
                    unpack2(accB[pp + 1], c0_, c1);        // (column 2pp+1, column 2pp+2)
                    acc[2 * pp] = a0 + b1;
                    acc[2 * pp + 1] = a1 + c0_;
                }
#else
#pragma unroll
                for (int jj = 0; jj < CT; ++jj) acc[jj] = 0.f;
#if S2_COMPACT
                for (int iu = lane; iu < S2_PARTS * nlive; iu += 32) {
                    const int uu = my_rows[iu / S2_PARTS], vc0 = (iu % S2_PARTS) * per, vc1 = min(nchunk, vc0 + per);
                    const float* drow = ring + slot_of(R0 + j + uu) + CT * warp + 4 * vc0;
                    const float* erow = sWt + uu * KP + 4 * vc0;
                    float win[CT + 4];
#pragma unroll
                    for (int jj = 0; jj < CT; jj += 4) {
                        const float4 t = *reinterpret_cast<const float4*>(drow + jj);
                        win[jj] = t.x; win[jj + 1] = t.y; win[jj + 2] = t.z; win[jj + 3] = t.w;
                    }
#pragma unroll 5
                    for (int vc = 0; vc < vc1 - vc0; ++vc) {
                        const float4 nw = *reinterpret_cast<const float4*>(drow + CT + 4 * vc);
                        const float4 e = *reinterpret_cast<const float4*>(erow + 4 * vc);
                        win[CT] = nw.x; win[CT + 1] = nw.y; win[CT + 2] = nw.z; win[CT + 3] = nw.w;
#pragma unroll
                        for (int jj = 0; jj < CT; ++jj) {
                            acc[jj] = fmaf(e.x, win[jj], acc[jj]);
                            acc[jj] = fmaf(e.y, win[jj + 1], acc[jj]);
                            acc[jj] = fmaf(e.z, win[jj + 2], acc[jj]);
                            acc[jj] = fmaf(e.w, win[jj + 3], acc[jj]);
                        }
#pragma unroll
                        for (int jj = 0; jj < CT; ++jj) win[jj] = win[jj + 4];
                    }
                }
#else
                for (int uu = lane; uu < K; uu += 32) {
                    const float* drow = ring + slot_of(R0 + j + uu) + CT * warp;
                    const float* erow = sWt + uu * KP;
                    float win[CT + 4];
#pragma unroll
                    for (int jj = 0; jj < CT; jj += 4) {
                        const float4 t = *reinterpret_cast<const float4*>(drow + jj);
                        win[jj] = t.x; win[jj + 1] = t.y; win[jj + 2] = t.z; win[jj + 3] = t.w;
                    }
                    const int nchunk = KP / 4;
#pragma unroll 5
                    for (int vc = 0; vc < nchunk; ++vc) {
                        const float4 nw = *reinterpret_cast<const float4*>(drow + CT + 4 * vc);
                        const float4 e = *reinterpret_cast<const float4*>(erow + 4 * vc);
                        win[CT] = nw.x; win[CT + 1] = nw.y; win[CT + 2] = nw.z; win[CT + 3] = nw.w;
#pragma unroll
                        for (int jj = 0; jj < CT; ++jj) {
                            acc[jj] = fmaf(e.x, win[jj], acc[jj]);
                            acc[jj] = fmaf(e.y, win[jj + 1], acc[jj]);
                            acc[jj] = fmaf(e.z, win[jj + 2], acc[jj]);
                            acc[jj] = fmaf(e.w, win[jj + 3], acc[jj]);
                        }
#pragma unroll
                        for (int jj = 0; jj < CT; ++jj) win[jj] = win[jj + 4];
                    }
                }
#endif
#endif
                const float tot = reduce_scatter<CT>(acc, lane);
                constexpr int SH = (CT == 32) ? 0 : (CT == 16) ? 1 : (CT == 8) ? 2 : (CT == 4) ? 3 : 4;
                if ((lane & ((1 << SH) - 1)) == 0) {
                    sI[(par * T + j) * TC + CT * warp + (lane >> SH)] = tot;
                    sScale[j * TC + CT * warp + (lane >> SH)] = 1073741824.0f / fmaxf(tot, 1e-30f);
                }
            }
            S2_TICK(0);
        }
        __syncthreads();
        S2_TICK(aux ? 5 : 1);
        // ---- between two groups, every thread: the pixels of this group's deaths lose their molecules, and the dwells
        // that saw those molecules although they were gone lose the corresponding taps (eight lanes per death, all
        // loads first so that every thread keeps eight independent chains in flight)
        {
            const uint32_t* sev = sEv + par * EVCAP;
            const int ev_lo = sMeta[2 * par], n_ev = sMeta[2 * par + 1] - ev_lo;
#ifdef S2_ABLATE_DEC                         // timing ablation only: nothing bleaches
            if (false)
#endif
            for (int q = tid; q < n_ev; q += S2_THREADS) {
                const uint32_t ev = q < EVCAP ? sev[q] : __ldg(ws.ev + ev_lo + q);
                const int xl = (int)(ev & 0xFFFu) - c0, u = (ev >> 12) & 0x3F;
                const int rd = R0 + (int)(((ev >> 24) - (uint32_t)R0) & 255u);
                if (xl >= 0 && xl < span) atomicAdd(&ring[slot_of(rd + u) + xl], -1.0f);
            }
#if SWEEP_PROBE == 4
            if (!aux) S2_TICK(3);            // ring decrements of this thread done
#endif
            constexpr int EV_LANES = 8, EV_SLOTS = S2_THREADS / EV_LANES;
            const int sub = tid % EV_LANES;
#ifdef S2_ABLATE_CORR                        // timing ablation only: wrong intensities
            if (false)
#endif
            for (int q = tid / EV_LANES; q < n_ev; q += EV_SLOTS) {
                const uint32_t ev = q < EVCAP ? sev[q] : __ldg(ws.ev + ev_lo + q);
                const int xl = (int)(ev & 0xFFFu) - c0, u = (ev >> 12) & 0x3F, a = (ev >> 18) & 0x3F;
                if (xl < 0 || xl >= span) continue;                       // another strip's pixel
                const int jd = (int)(((ev >> 24) - (uint32_t)R0) & 255u);       // row of the death inside the group
                const int v0 = max(max(0, c0 + xl - W + 1), xl - (TC - 1));     // dwell column t = xl - v, 0 <= t < TC
                const int vall = min(K, xl + 1);
                const int jend = min(Tg, jd + u + 1);                           // tap row u - (j - jd) >= 0
                for (int j = jd; j < jend; ++j) {
                    const float* e = sWt + (u - (j - jd)) * KP;
                    const float* sc = sScale + j * TC + xl;
                    int* corr = sCorr + j * TC + xl;
                    const int v1 = (j == jd) ? min(a, vall) : vall;
                    float te[8], ts[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int v = v0 + sub + EV_LANES * k;
                        const bool ok = v < v1;
                        te[k] = ok ? e[v] : 0.f;
                        ts[k] = ok ? sc[-v] : 0.f;
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int v = v0 + sub + EV_LANES * k;
                        if (v < v1) atomicAdd(corr - v, __float2int_rn(te[k] * ts[k]));
                    }
                }
            }
        }
#if SWEEP_PROBE == 4
        if (!aux) S2_TICK(4);                // corrections of this thread done
#endif
        __syncthreads();
        S2_TICK(2);
    }
#if SWEEP_PROBE == 4
    if (tid == 0) {
        t_ph[6] = grp; t_ph[7] = t_a - t_begin; t_ph[5] = 0;
        for (int i = 0; i < 8; ++i) atomicAdd(&g_phase_cycles[i], (unsigned long long)t_ph[i]);
    }
#endif
#if SWEEP_PROBE == 3
    if (tid == 0 || tid == S2_GATHER) {
        t_ph[6] = (tid == 0) ? grp : 0;
        t_ph[7] = (tid == 0) ? (t_a - t_begin) : 0;     // whole loop, gather thread
        if (tid == 0) { t_ph[3] = t_ph[4] = t_ph[5] = 0; } else { t_ph[0] = t_ph[1] = t_ph[2] = 0; }
#ifdef S2_COUNT_PASSES
        if (tid == 0) { t_ph[4] = s2_passes * 1000; t_ph[5] = s2_rows * 1000; }          // read as passes / rows
#endif
        for (int i = 0; i < 8; ++i) atomicAdd(&g_phase_cycles[i], (unsigned long long)t_ph[i]);
    }
#endif
    // ---- epilogue: the last group's intensities, the band rows this range owns and has not written yet
    const int Rl = r0 + (grp - 1) * T;        // first row of the last group
    if (aux) write_rows(Rl, r1 - Rl, (grp - 1) & 1, at, S2_AUX);
    const int yend = (r1 == H) ? P.Hp : r1;
    for (int i = tid; i < (yend - Rl) * q4; i += S2_THREADS) {
        const int yy = i / q4;
        retire_row(Rl + yy, i - yy * q4);
    }
}

// ------------------------------------------------------------------------------------------
// K3c: the stochastic scan, third generation (sweep3_kernel).
//
// Same work items and the same arithmetic as sweep2_kernel with one scan row per group, reorganised around the round-2
// measurements (profiles/r02_summary.md): per scan row a sweep2 CTA spent as long between its two CTA-wide barriers
// (ring decrements + intensity corrections, every warp) and in the auxiliary warps' ring traffic as in the gather, and
// three CTAs per SM could not hide that.  Here
//  * the eight GATHER warps do nothing but gather: after the dense gather of scan row r they apply the row's deaths to
//    the ring (one compare-and-swap per death — the only thing the next row's gather depends on), meet at a barrier of
//    their own, and start on row r + 1;
//  * four CORRECTION warps take the intensity corrections of row r and its write-out WHILE row r + 1 is gathered
//    (named barriers hand the row over and back; dense intensities, scales and corrections are double buffered);
//  * the ring moves by TMA: the finished band row leaves with cp.async.bulk (shared -> global) and the row the gather
//    needs two scan rows later arrives with cp.async.bulk (global -> shared) on an mbarrier, issued by one gather
//    thread in a handful of instructions; the per-row event lists arrive the same way, three rows ahead.
// MEASURED (B200, 256 envs x 256 x 256, profiles/r02_summary.md): identical results, but 4.21 ms against sweep2's 3.87 ms
// for the whole scan.  Per scan row the correction warps need 15 k cycles for ~1.5 k instructions each: next to six
// always-ready FFMA warps per scheduler a latency-bound helper warp gets an issue slot every ~7 cycles, so the gather
// warps end up waiting for it (4.8 k cycles per row).  An owner-computes variant without atomics was no faster (17 k).
// sweep2's shape — every warp does the side work between two barriers — wastes fewer issue slots, so it stays the
// default; this kernel runs with STED_B200_SCAN_V3=1 and is covered by the same equality tests.
// ------------------------------------------------------------------------------------------
constexpr int S3_GATHER = 256, S3_CORR = 128, S3_THREADS = S3_GATHER + S3_CORR, S3_EVCAP = 512;
constexpr int S3_BAR_GATHER = 1, S3_BAR_READY = 2 /* +parity */, S3_BAR_FREE = 4 /* +parity */, S3_BAR_CORR = 6;

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// named barriers with immediate ids (a register id makes ptxas reserve all sixteen)
template <int ID, int N> __device__ __forceinline__ void named_sync() { asm volatile("bar.sync %0, %1;" :: "n"(ID), "n"(N) : "memory"); }
template <int ID, int N> __device__ __forceinline__ void named_arrive() { asm volatile("bar.arrive %0, %1;" :: "n"(ID), "n"(N) : "memory"); }
template <int ID, int N> __device__ __forceinline__ void named_sync2(int par) { if (par) named_sync<ID + 1, N>(); else named_sync<ID, N>(); }
template <int ID, int N> __device__ __forceinline__ void named_arrive2(int par) { if (par) named_arrive<ID + 1, N>(); else named_arrive<ID, N>(); }

__host__ __device__ inline size_t sweep3_smem_floats(int K, int KP, int TC) {
    return (size_t)(K + 2) * sweep_pitch(TC, KP) + (size_t)K * KP + 6 * TC + 3 * S3_EVCAP + 12;
}

template <int KT, int CT>
__global__ void __launch_bounds__(S3_THREADS, 3) sweep3_kernel(const AcqParams P, const BleachWs ws, const int RR) {
    constexpr int TC = 8 * CT;
    const int K = KT ? KT : P.K;
    const int KP = KT ? STED_KPITCH(KT) : P.KP;
    const int NS = K + 2;                     // ring slots: the K rows of a gather, the next one, and one in flight
    const int PR = sweep_pitch(TC, KP);
    const int span = TC + K - 1;

    extern __shared__ __align__(128) float smem[];
    float* ring = smem;                       // [K+2][PR] band of the molecule map (integer counts)
    float* sWt = ring + NS * PR;              // [K][KP]   effective PSF E
    float* sI = sWt + K * KP;                 // [2][TC]   dense row intensities (even / odd scan rows)
    float* sScale = sI + 2 * TC;              // [2][TC]   2^30 / sI  (fixed-point scale of the corrections)
    int* sCorr = reinterpret_cast<int*>(sScale + 2 * TC);               // [2][TC] sum of removed taps, fixed point
    uint32_t* sEv = reinterpret_cast<uint32_t*>(sCorr + 2 * TC);        // [3][S3_EVCAP] events of three scan rows
    int* sMeta = reinterpret_cast<int*>(sEv + 3 * S3_EVCAP);            // [3][4] lo, hi in ws.ev; offset and count staged
    __shared__ __align__(8) unsigned long long mbar[5];                 // [0..1] ring rows in flight, [2..4] event lists

    const int b = blockIdx.z;
    const int c0 = blockIdx.x * TC;
    const int r0 = blockIdx.y * RR;
    const int r1 = min(P.H, r0 + RR);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool corr = tid >= S3_GATHER;
    const int ct = tid - S3_GATHER;           // index among the correction threads
    const int W = P.W, H = P.H;
    const float* din = P.din + (size_t)b * P.Hp * P.Wpitch;
    float* dout = P.dout + (size_t)b * P.Hp * P.Wpitch;
    const int* row_off = ws.off + (size_t)b * H;
    const uint32_t bar_full = smem_u32(&mbar[0]), bar_ev = smem_u32(&mbar[2]);

    {
        const float4* w4 = reinterpret_cast<const float4*>(P.tables + ((size_t)b * STED_TABLES + STED_TAB_E) * K * KP);
        for (int i = tid; i < K * KP / 4; i += S3_THREADS) reinterpret_cast<float4*>(sWt)[i] = __ldg(w4 + i);
    }
    for (int i = tid; i < 2 * TC; i += S3_THREADS) sCorr[i] = 0;
    if (tid == 0) {
        for (int i = 0; i < 5; ++i) mbar_init(smem_u32(&mbar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // columns of the molecule map this strip writes back (each column has exactly one owner; the border lies in the
    // overlap of two strips, where both hold the same counts, on a 16-byte boundary for the bulk copies)
    const int xs = (c0 == 0) ? 0 : c0 + 32;
    const int xe = (c0 + TC >= W) ? P.Wpitch : c0 + TC + 32;
    const int q4 = PR / 4;                    // float4 groups of a ring row
    const uint32_t row_bytes = (uint32_t)(min(PR, P.Wpitch - c0) * 4);   // what a bulk load brings (the rest of the slot stays zero)

    auto slot_of = [&](const int y) { return (KT ? y % (KT + 2) : y % NS) * PR; };
    auto clamp_hi = [&](const int v) { return (unsigned long long)v > ws.cap ? (int)ws.cap : v; };

    // band rows r0 .. r0+K+1 as they were BEFORE the scan ...
    for (int i = tid; i < NS * q4; i += S3_THREADS) {
        const int yy = i / q4, g = i - yy * q4;
        const int y = r0 + yy, x = c0 + 4 * g;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y < P.Hp && x < P.Wpitch) val = __ldg(reinterpret_cast<const float4*>(din + (size_t)y * P.Wpitch + x));
        *reinterpret_cast<float4*>(ring + slot_of(y) + 4 * g) = val;
    }
    __syncthreads();
    // ... minus the deaths of the scan rows before r0 that hit them (scan row r' reaches band rows r' .. r'+K-1)
    if (r0 > 0) {
        const int rp = max(0, r0 - (K - 1));
        const int lo = __ldg(row_off + rp), hi = clamp_hi(__ldg(row_off + r0));
        for (int q = lo + tid; q < hi; q += S3_THREADS) {
            const uint32_t ev = __ldg(ws.ev + q);
            const int xl = (int)(ev & 0xFFFu) - c0, u = (ev >> 12) & 0x3F;
            const int rr = r0 - 1 - (int)(((uint32_t)(r0 - 1) - (ev >> 24)) & 255u);     // scan row of the event
            const int y = rr + u;
            if (y >= r0 && xl >= 0 && xl < span) atomicAdd(&ring[slot_of(y) + xl], -1.0f);
        }
    }

    // event list of scan row r -> shared memory (bulk copy of the 16-byte blocks that hold it, at most S3_EVCAP words;
    // longer rows read their tail from global memory)
    auto issue_events = [&](const int r) {
        const int k = (r - r0) % 3;
        const int lo = __ldg(row_off + r), hi = clamp_hi(__ldg(row_off + r + 1));
        const int n = max(hi - lo, 0);
        const uintptr_t a0 = reinterpret_cast<uintptr_t>(ws.ev + lo);
        const uintptr_t base = a0 & ~(uintptr_t)15;
        const int off = (int)((a0 - base) >> 2);
        const int cnt = n ? min((off + n + 3) & ~3, S3_EVCAP) : 0;
        sMeta[4 * k] = lo; sMeta[4 * k + 1] = lo + n; sMeta[4 * k + 2] = off; sMeta[4 * k + 3] = max(cnt - off, 0);
        mbar_expect_tx(bar_ev + 8 * k, (uint32_t)cnt * 4u);
        if (cnt) bulk_g2s(smem_u32(sEv + k * S3_EVCAP), reinterpret_cast<const void*>(base), (uint32_t)cnt * 4u, bar_ev + 8 * k);
    };
    __syncthreads();                          // ring, tables and mbarriers are ready
    if (tid == S3_GATHER)
        for (int r = r0; r < min(r1, r0 + 3); ++r) issue_events(r);

#if SWEEP_PROBE == 5
    long long t_ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t_a = clock64(), t_b;
#define S3_TICK(i) do { t_b = clock64(); t_ph[i] += t_b - t_a; t_a = t_b; } while (0)
#else
#define S3_TICK(i) do { } while (0)
#endif
    if (!corr) {
        // =================================== gather warps ===================================
        for (int r = r0; r < r1; ++r) {
            const int it = r - r0, par = it & 1;
            if (it >= 3) mbar_wait(bar_full + 8 * ((it - 2) & 1), (uint32_t)(((it - 3) >> 1) & 1));   // band row r+K-1 has landed
            if (it >= 2) named_sync2<S3_BAR_FREE, S3_THREADS>(par);       // row r-2 has been written out: sI / sScale[par] are free
            S3_TICK(0);                                                   // waits: ring row + free buffers
            {
                float acc[CT];
#pragma unroll
                for (int jj = 0; jj < CT; ++jj) acc[jj] = 0.f;
                for (int uu = lane; uu < K; uu += 32) {
                    const float* drow = ring + slot_of(r + uu) + CT * warp;
                    const float* erow = sWt + uu * KP;
                    float win[CT + 4];
#pragma unroll
                    for (int jj = 0; jj < CT; jj += 4) {
                        const float4 t = *reinterpret_cast<const float4*>(drow + jj);
                        win[jj] = t.x; win[jj + 1] = t.y; win[jj + 2] = t.z; win[jj + 3] = t.w;
                    }
                    const int nchunk = KP / 4;
#pragma unroll 5
                    for (int vc = 0; vc < nchunk; ++vc) {
                        const float4 nw = *reinterpret_cast<const float4*>(drow + CT + 4 * vc);
                        const float4 e = *reinterpret_cast<const float4*>(erow + 4 * vc);
                        win[CT] = nw.x; win[CT + 1] = nw.y; win[CT + 2] = nw.z; win[CT + 3] = nw.w;
#pragma unroll
                        for (int jj = 0; jj < CT; ++jj) {
                            acc[jj] = fmaf(e.x, win[jj], acc[jj]);
                            acc[jj] = fmaf(e.y, win[jj + 1], acc[jj]);
                            acc[jj] = fmaf(e.z, win[jj + 2], acc[jj]);
                            acc[jj] = fmaf(e.w, win[jj + 3], acc[jj]);
                        }
#pragma unroll
                        for (int jj = 0; jj < CT; ++jj) win[jj] = win[jj + 4];
                    }
                }
                const float tot = reduce_scatter<CT>(acc, lane);
                constexpr int SH = (CT == 32) ? 0 : (CT == 16) ? 1 : (CT == 8) ? 2 : (CT == 4) ? 3 : 4;
                if ((lane & ((1 << SH) - 1)) == 0) {
                    sI[par * TC + CT * warp + (lane >> SH)] = tot;
                    sScale[par * TC + CT * warp + (lane >> SH)] = 1073741824.0f / fmaxf(tot, 1e-30f);
                }
            }
            S3_TICK(1);                                                   // gather
            __syncwarp();
            named_sync<S3_BAR_GATHER, S3_GATHER>();                         // every gather of row r is done: the ring may change
            S3_TICK(2);
            {   // the pixels of this row's deaths lose their molecules
                const int k = it % 3;
                mbar_wait(bar_ev + 8 * k, (uint32_t)((it / 3) & 1));
                const uint32_t* sev = sEv + k * S3_EVCAP + sMeta[4 * k + 2];
                const int ev_lo = sMeta[4 * k], n_ev = sMeta[4 * k + 1] - ev_lo, staged = sMeta[4 * k + 3];
                for (int q = tid; q < n_ev; q += S3_GATHER) {
                    const uint32_t ev = q < staged ? sev[q] : __ldg(ws.ev + ev_lo + q);
                    const int xl = (int)(ev & 0xFFFu) - c0, u = (ev >> 12) & 0x3F;
                    if (xl >= 0 && xl < span) atomicAdd(&ring[slot_of(r + u) + xl], -1.0f);
                }
            }
            fence_async_smem();                                           // the bulk store below reads what was just written
            S3_TICK(3);                                                   // decrements
            __syncwarp();
            named_sync<S3_BAR_GATHER, S3_GATHER>();
            if (tid == 0) {
                if (it >= 1) {
                    bulk_wait_read0();                                    // row r-1 has left its slot (stored one scan row ago)
                    if (r + 2 < r1) {                                     // band row r+K+1, first needed by the gather of row r+2
                        const int y = r - 1 + NS;
                        mbar_expect_tx(bar_full + 8 * (it & 1), row_bytes);
                        bulk_g2s(smem_u32(ring + slot_of(y)), din + (size_t)y * P.Wpitch + c0, row_bytes, bar_full + 8 * (it & 1));
                    }
                }
                // band row r is final: the columns this strip owns go back to global memory
                bulk_s2g(dout + (size_t)r * P.Wpitch + xs, smem_u32(ring + slot_of(r) + (xs - c0)), (uint32_t)(xe - xs) * 4u);
                bulk_commit();
            }
            __syncwarp();
            named_arrive2<S3_BAR_READY, S3_THREADS>(par);                 // row r: dense intensities, scales and ring are settled
            S3_TICK(4);                                                   // second barrier + bulk copies
        }
#if SWEEP_PROBE == 5
        if (tid == 0) { t_ph[7] = r1 - r0; for (int i = 0; i < 8; ++i) if (i != 5 && i != 6) atomicAdd(&g_phase_cycles[i], (unsigned long long)t_ph[i]); }
#endif
    } else {
        // ================================= correction warps =================================
        constexpr int EV_LANES = 8, EV_SLOTS = S3_CORR / EV_LANES;
        const int sub = ct % EV_LANES;
        for (int r = r0; r < r1; ++r) {
            const int it = r - r0, par = it & 1, k = it % 3;
            named_sync2<S3_BAR_READY, S3_THREADS>(par);
            mbar_wait(bar_ev + 8 * k, (uint32_t)((it / 3) & 1));
            S3_TICK(5);                                                   // correction warps: waiting for a row
            const uint32_t* sev = sEv + k * S3_EVCAP + sMeta[4 * k + 2];
            const int ev_lo = sMeta[4 * k], n_ev = sMeta[4 * k + 1] - ev_lo, staged = sMeta[4 * k + 3];
            // a death removes the taps its molecule no longer contributes to from the rest of its own scan row
            // (eight lanes per death, all loads first so that every thread keeps eight independent chains in flight)
            for (int q = ct / EV_LANES; q < n_ev; q += EV_SLOTS) {
                const uint32_t ev = q < staged ? sev[q] : __ldg(ws.ev + ev_lo + q);
                const int xl = (int)(ev & 0xFFFu) - c0, u = (ev >> 12) & 0x3F, a = (ev >> 18) & 0x3F;
                if (xl < 0 || xl >= span) continue;                       // another strip's pixel
                const int v0 = max(max(0, c0 + xl - W + 1), xl - (TC - 1));     // dwell column t = xl - v, 0 <= t < TC
                const int v1 = min(a, min(K, xl + 1));
                const float* e = sWt + u * KP;
                const float* sc = sScale + par * TC + xl;
                int* cr = sCorr + par * TC + xl;
                float te[8], ts[8];
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const int v = v0 + sub + EV_LANES * kk;
                    const bool ok = v < v1;
                    te[kk] = ok ? e[v] : 0.f;
                    ts[kk] = ok ? sc[-v] : 0.f;
                }
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const int v = v0 + sub + EV_LANES * kk;
                    if (v < v1) atomicAdd(cr - v, __float2int_rn(te[kk] * ts[kk]));
                }
            }
            __syncwarp();
            named_sync<S3_BAR_CORR, S3_CORR>();
            for (int t = ct; t < TC; t += S3_CORR) {
                if (c0 + t < W) {
                    const float I = fmaxf(sI[par * TC + t] * (1.0f - (float)sCorr[par * TC + t] * 9.313225746154785e-10f), 0.f);
                    const size_t o = ((size_t)b * H + r) * W + c0 + t;
                    if (P.intensity) P.intensity[o] = I;
                    P.signal[o] = I;                          // detect_kernel converts to photon counts
                }
                sCorr[par * TC + t] = 0;
            }
            if (ct == 0 && r + 3 < r1) issue_events(r + 3);   // this row's list has been read by everybody: its buffer is free
            __syncwarp();
            if (r + 2 < r1) named_arrive2<S3_BAR_FREE, S3_THREADS>(par);
            S3_TICK(6);                                                   // corrections + write-out
        }
#if SWEEP_PROBE == 5
        if (tid == S3_GATHER) { atomicAdd(&g_phase_cycles[5], (unsigned long long)t_ph[5]); atomicAdd(&g_phase_cycles[6], (unsigned long long)t_ph[6]); }
#endif
    }
    __syncthreads();
    if (tid == 0) bulk_wait0();               // the bulk stores have left shared memory and reached global memory
    // the band rows below the image belong to the last range of the strip
    if (r1 == H) {
        for (int i = tid; i < (P.Hp - H) * q4; i += S3_THREADS) {
            const int yy = i / q4, g = i - yy * q4;
            const int y = H + yy, x = c0 + 4 * g;
            const float4 old = *reinterpret_cast<const float4*>(ring + slot_of(y) + 4 * g);
            const float* po_ = reinterpret_cast<const float*>(&old);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (x + j >= xs && x + j < xe) dout[(size_t)y * P.Wpitch + x + j] = po_[j];
        }
    }
}

// ------------------------------------------------------------------------------------------
// K4: the general scan — per-dwell imaging parameters.
//
// The timed environments hand get_signal_and_bleach one dwell time PER PIXEL (gym_sted/envs/timed_sted_env.py:141-148:
// `pdt * numpy.ones(dmap_shape)`, zeros where the experiment's clock runs out), and pySTED accepts the same for p_ex
// and p_sted.  With per-dwell hazards the suffix-sum tables of the row sweep no longer describe a scan row, so this
// path keeps the reference's order instead: dwells are visited one after the other, the K x K window of every dwell is
// gathered and thinned cooperatively by the whole CTA (512 threads, band of K map rows in shared memory, float64 when
// it fits), and the parallelism comes from the window and from the batch — one CTA per environment, which is what the
// north star's "wavefront" amounts to once the footprints of consecutive dwells overlap (K = 59 against 64-pixel
// rows: the dwells of a row never have disjoint footprints).
//   I[r][c]      = w_ex[r][c] * sum E_j[u][v] * D[r+u][c+v]                      j = table_id[r][c]
//   D[r+u][c+v] <- Binomial(D, exp(-A_j[u][v] * w_dt[r][c] * w_ex[r][c]))   or  D * exp(...) in expectation mode
// with E_j, A_j = k_b * pdt_ref the tables of the j-th distinct (p_sted) value at the reference p_ex and pdt (E and
// k_b are linear in p_ex; the photon-flux floor of get_photons moves k_b by < 1e-20 relative).  The per-dwell
// binomials are the reference's own chain (sample_molecules), drawn with Philox keyed by (pixel, tap).
// ------------------------------------------------------------------------------------------
constexpr int GS_THREADS = 512;

template <typename BandT>
__global__ void __launch_bounds__(GS_THREADS) general_scan_kernel(const AcqParams P, const double* __restrict__ kbleach,
                                                                  const float* __restrict__ w_dt, const float* __restrict__ w_ex,
                                                                  const unsigned char* __restrict__ table_id, const int n_tables) {
    const int K = P.K, KP = P.KP, KK = K * K, H = P.H, W = P.W, Wp = P.Wp, Hp = P.Hp;
    extern __shared__ __align__(16) unsigned char gs_smem[];
    double* sS = reinterpret_cast<double*>(gs_smem);                 // [KK] survival of the current dwell
    BandT* band = reinterpret_cast<BandT*>(sS + KK);                 // [K][Wp] ring of map rows, slot = row % K
    float* sE = reinterpret_cast<float*>(band + (size_t)K * Wp);     // [KK] effective PSF of the current table
    float* sA = sE + KK;                                             // [KK] hazard k_b * pdt_ref of the current table
    __shared__ double s_red[2][GS_THREADS / 32];     // double-buffered by dwell parity: one barrier per dwell
    const int b = blockIdx.x, tid = threadIdx.x;
    const bool bleach = (P.flags & STED_BLEACH) != 0, stoch = bleach && (P.flags & STED_SAMPLE_BLEACH);
    const float* din = P.din + (size_t)b * Hp * P.Wpitch;
    float* dout = bleach ? P.dout + (size_t)b * Hp * P.Wpitch : nullptr;
    const uint32_t env = P.env_base + (uint32_t)b;
    const double pdt_ref = P.pdt[b];

    auto load_table = [&](const int j) {
        const float* tE = P.tables + ((size_t)(b * n_tables + j) * STED_TABLES + STED_TAB_E) * K * KP;
        const double* kb = kbleach + (size_t)(b * n_tables + j) * KK;
        for (int t = tid; t < KK; t += GS_THREADS) {
            sE[t] = tE[(t / K) * KP + (t % K)];
            sA[t] = (float)(kb[t] * pdt_ref);
        }
    };
    for (int i = tid; i < K * Wp; i += GS_THREADS) {
        const int y = i / Wp, x = i - y * Wp;
        band[(size_t)(y % K) * Wp + x] = (y < Hp) ? (BandT)din[(size_t)y * P.Wpitch + x] : (BandT)0;
    }
    int cur_table = 0;
    float cur_w = -1.0f;
    load_table(0);
    __syncthreads();

    // taps of this thread: t = tid + i * GS_THREADS, fixed for the whole scan
    constexpr int GS_TAPS = (STED_MAX_K * STED_MAX_K + GS_THREADS - 1) / GS_THREADS;
    int tap_u[GS_TAPS], tap_v[GS_TAPS];
#pragma unroll
    for (int i = 0; i < GS_TAPS; ++i) {
        const int t = tid + i * GS_THREADS;
        tap_u[i] = t < KK ? t / K : -1;
        tap_v[i] = t < KK ? t % K : 0;
    }
    int parity = 0, r0 = 0;                                            // r0 = r % K
    for (int r = 0; r < H; ++r, r0 = (r0 + 1 == K ? 0 : r0 + 1)) {
        for (int c = 0; c < W; ++c) {
            const size_t o = ((size_t)b * H + r) * W + c;
            const float wx = w_ex ? w_ex[o] : 1.0f;
            const float w = (w_dt ? w_dt[o] : 1.0f) * wx;              // hazard scale of this dwell
            const int j = table_id ? (int)table_id[o] : 0;
            if (j != cur_table) {                                      // uniform: every thread reads the same dwell record
                __syncthreads();
                load_table(j);
                cur_table = j;
                cur_w = -1.0f;
                __syncthreads();
            }
            if (bleach && !stoch && w != cur_w) {
                for (int t = tid; t < KK; t += GS_THREADS) sS[t] = exp(-(double)sA[t] * (double)w);
                cur_w = w;
                __syncthreads();
            }
            double acc = 0.0;
            if (wx != 0.0f || (bleach && w != 0.0f)) {
#pragma unroll
                for (int i = 0; i < GS_TAPS; ++i) {
                    const int u = tap_u[i], v = tap_v[i];
                    if (u < 0) continue;
                    const int t = tid + i * GS_THREADS;
                    int slot = r0 + u;
                    slot -= slot >= K ? K : 0;
                    BandT* cell = band + (size_t)slot * Wp + c + v;
                    const BandT d = *cell;
                    if (d != (BandT)0) {
                        acc += (double)sE[t] * (double)d;
                        if (stoch) {
                            PhiloxStream rng;
                            rng.init(P.seed, P.offset, env, (uint32_t)((r + u) * Wp + c + v), STED_RNG_BLEACH3, (uint32_t)t);
                            *cell = d - (BandT)binomial_deaths((int)d, sA[t] * w, rng);
                        } else if (bleach) {
                            *cell = (BandT)((double)d * sS[t]);
                        }
                    }
                }
            }
            for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            if ((tid & 31) == 0) s_red[parity][tid >> 5] = acc;
            __syncthreads();                                           // the only barrier of a dwell: orders the band
            if (tid < 32) {                                            // updates and publishes the partial sums
                double tot = tid < GS_THREADS / 32 ? s_red[parity][tid] : 0.0;
                for (int off = 16; off; off >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, off);
                if (tid == 0) {
                    const float I = (float)(tot * (double)wx);
                    if (P.intensity) P.intensity[o] = I;
                    P.signal[o] = I;
                }
            }
            parity ^= 1;
        }
        __syncthreads();
        // map row r is final: it leaves the band and row r + K takes its slot
        BandT* slot = band + (size_t)(r % K) * Wp;
        for (int x = tid; x < Wp; x += GS_THREADS) {
            if (bleach) dout[(size_t)r * P.Wpitch + x] = (float)slot[x];
            slot[x] = (r + K < Hp) ? (BandT)din[(size_t)(r + K) * P.Wpitch + x] : (BandT)0;
        }
        __syncthreads();
    }
    if (bleach) {
        for (int i = tid; i < (K - 1) * Wp; i += GS_THREADS) {
            const int y = H + i / Wp, x = i % Wp;
            dout[(size_t)y * P.Wpitch + x] = (float)band[(size_t)(y % K) * Wp + x];
        }
        for (int i = tid; i < Hp * (P.Wpitch - Wp); i += GS_THREADS)           // pitch padding stays zero
            dout[(size_t)(i / (P.Wpitch - Wp)) * P.Wpitch + Wp + i % (P.Wpitch - Wp)] = 0.f;
    }
}

// ------------------------------------------------------------------------------------------
// sampler test kernels, FMA peak
// ------------------------------------------------------------------------------------------
__global__ void poisson_test_kernel(const float* lam, int64_t n, uint64_t seed, uint32_t offset, float* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PhiloxStream rng;
    rng.init(seed, offset, (uint32_t)(i >> 32), (uint32_t)i, STED_RNG_TEST, 0);
    out[i] = poisson_sample(lam[i], rng);
}

__global__ void binomial_test_kernel(const float* nn, const float* q, int64_t n, uint64_t seed, uint32_t offset, float* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PhiloxStream rng;
    rng.init(seed, offset, (uint32_t)(i >> 32), (uint32_t)i, STED_RNG_TEST, 1);
    const float hz = -log1pf(-q[i]);
    out[i] = (float)binomial_deaths((int)nn[i], hz, rng);
}

__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, int iters) {
    float a[16];
    const float x = 1.0f + 1e-7f * threadIdx.x, y = 1e-9f * blockIdx.x;
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = (float)j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = fmaf(a[j], x, y);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += a[j];
    if (s == 12345.678f) out[0] = s;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
int make_params(AcqParams& P, const float* din, float* dout, const float* tables, const double* pdt, int B, int H,
                int W, int K, uint32_t flags, const StedDetectorParams* det, double lambda_fluo, uint64_t seed,
                uint64_t offset, int64_t env_base, float* intensity, float* signal) {
    if (!din || !tables || !pdt || !signal || !det) return fail(STED_EINVAL, "null pointer argument");
    if (B <= 0 || H <= 0 || W <= 0) return fail(STED_EINVAL, "B, H, W must be positive (got %d, %d, %d)", B, H, W);
    if (K < 1 || (K & 1) == 0) return fail(STED_EINVAL, "K must be odd (got %d)", K);
    if (K > STED_MAX_K) return fail(STED_EUNSUPPORTED, "K=%d exceeds STED_MAX_K=%d", K, STED_MAX_K);
    if (H > 65535) return fail(STED_EUNSUPPORTED, "H=%d exceeds 65535 scan rows", H);
    if (B > 65535) return fail(STED_EUNSUPPORTED, "B=%d exceeds 65535 environments per launch (grid limit): split the batch", B);
    if (W + K - 1 > 4095) return fail(STED_EUNSUPPORTED, "W=%d exceeds the 4095-column limit of the event encoding", W);
    if ((flags & STED_BLEACH) && (!dout || dout == din))
        return fail(STED_EINVAL, "bleaching needs dmap_out_dev distinct from dmap_in_dev");
    if (!(lambda_fluo > 0)) return fail(STED_EINVAL, "lambda_fluo must be positive");
    P.din = din; P.dout = dout; P.tables = tables; P.pdt = pdt; P.intensity = intensity; P.signal = signal;
    P.B = B; P.H = H; P.W = W; P.K = K; P.KP = STED_KPITCH(K);
    P.Hp = H + K - 1; P.Wp = W + K - 1; P.Wpitch = (P.Wp + 3) & ~3;
    P.flags = flags; P.offset = (uint32_t)offset; P.seed = seed ^ (offset >> 32 << 17); P.env_base = (uint32_t)env_base;
    P.inv_e_photon = (float)(lambda_fluo / (kPlanck * kLight));
    P.det_eff = (float)(det->pcef * det->pdef);
    P.background = (float)det->background;
    P.wdt = nullptr;
    return STED_OK;
}

template <typename Kern>
int set_smem(Kern kern, size_t bytes) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return STED_OK;
}

// 3-D tensor map of the padded molecule maps [B][Hp][Wpitch] with a (tile + halo) box; elements outside the
// map read as zero, so edge tiles need no special case.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tile_map(const AcqParams& P, CUtensorMap* map, int tile_h) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) return fail(STED_ECUDA, "cuTensorMapEncodeTiled is not available in this driver");
        encode = (EncodeTiledFn)fn;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)P.Wpitch, (cuuint64_t)P.Hp, (cuuint64_t)P.B};
    const cuuint64_t strides[2] = {(cuuint64_t)P.Wpitch * sizeof(float), (cuuint64_t)P.Hp * P.Wpitch * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)(G_TW + P.KP), (cuuint32_t)(tile_h + P.K - 1), 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    if (((uintptr_t)P.din & 15u) != 0) return fail(STED_EINVAL, "dmap_in_dev must be 16-byte aligned");
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(P.din), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(STED_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return STED_OK;
}

template <int KT, int TH>
int launch_gather_t(const AcqParams& P, const CUtensorMap& map, cudaStream_t st) {
    const size_t smem = ((size_t)(TH + P.K - 1) * (G_TW + P.KP) + 2 * (size_t)(P.K + 1) * P.KP) * sizeof(float);
    const dim3 grid((P.W + G_TW - 1) / G_TW, (P.H + TH - 1) / TH, P.B);
    int rc;
    if ((rc = set_smem(gather_kernel<KT, TH>, smem))) return rc;
    gather_kernel<KT, TH><<<grid, (TH / G_RT) * 16, smem, st>>>(P, map);
    return STED_OK;
}

int launch_gather(const AcqParams& P, cudaStream_t st) {
    // 32-row tiles when they cover the image with fewer computed rows than 64-row tiles (H mod 64 in 1..32)
    const bool half = G_TH == 64 && ((P.H + 31) / 32) * 32 < ((P.H + 63) / 64) * 64;
    CUtensorMap map;
    int rc;
    if ((rc = make_tile_map(P, &map, half ? 32 : G_TH))) return rc;
    if (P.K == 59) rc = half ? launch_gather_t<59, 32>(P, map, st) : launch_gather_t<59, G_TH>(P, map, st);
    else rc = half ? launch_gather_t<0, 32>(P, map, st) : launch_gather_t<0, G_TH>(P, map, st);
    if (rc) return rc;
    CUDA_TRY(cudaGetLastError());
    return STED_OK;
}

template <int KT, int CT, bool STOCH>
int launch_sweep_t(const AcqParams& P, const BleachWs& ws, cudaStream_t st) {
    constexpr int TC = 8 * CT;
    const int PR = sweep_pitch(TC, P.KP);
    size_t words = (size_t)P.K * PR + (size_t)P.K * P.KP + 6 * TC;
    if (!STOCH) words += 2 * (size_t)P.K * P.KP;
    const size_t smem = words * sizeof(float);
    dim3 grid((P.W + TC - 1) / TC, P.B);
    int rc;
    if ((rc = set_smem(sweep_kernel<KT, CT, STOCH>, smem))) return rc;
    sweep_kernel<KT, CT, STOCH><<<grid, S_THREADS, smem, st>>>(P, ws);
    CUDA_TRY(cudaGetLastError());
    return STED_OK;
}

template <bool STOCH>
int launch_sweep_m(const AcqParams& P, const BleachWs& ws, cudaStream_t st) {
#ifndef SWEEP_WIDE_MIN
#define SWEEP_WIDE_MIN 64      // widest image still scanned in 64-column strips
#endif
    const bool wide = P.W > SWEEP_WIDE_MIN;
    if (P.K == 59) return wide ? launch_sweep_t<59, 16, STOCH>(P, ws, st) : launch_sweep_t<59, 8, STOCH>(P, ws, st);
    return wide ? launch_sweep_t<0, 16, STOCH>(P, ws, st) : launch_sweep_t<0, 8, STOCH>(P, ws, st);
}

// bytes of workspace the stochastic scan needs for B environments of H rows and at most max_events deaths:
// status + per-row counters and offsets + per event 8 bytes of staging and 4 bytes in its row bucket
size_t bleach_ws_bytes(int B, int H, unsigned long long max_events) {
    // (+ 16 bytes: the bulk copies of the event lists move whole 16-byte blocks)
    return 16 + ((size_t)B * H * 2 + 4) * sizeof(int) + (size_t)max_events * (sizeof(uint2) + sizeof(uint32_t)) + 16;
}

// rows per work item of sweep2_kernel: few enough waves of `slots` concurrent CTAs, long enough ranges that the
// K-row preload of a range stays small next to its gathers (cost model in units of one gathered row)
int pick_row_range(int B, int strips, int H, int slots) {
    int best_rr = H;
    double best = 1e300;
    for (int parts = 1; parts <= 16 && (H + parts - 1) / parts >= 16; ++parts) {
        const int rr = (H + parts - 1) / parts;
        const long long items = (long long)B * strips * ((H + rr - 1) / rr);
        const double waves = (double)((items + slots - 1) / slots);
        const double cost = waves * (rr + 8.0);
        if (cost < best - 1e-9) { best = cost; best_rr = rr; }
    }
    if (const char* e = getenv("STED_B200_ROW_RANGE")) {           // tuning knob
        const int v = atoi(e);
        if (v >= 1) best_rr = v < H ? v : H;
    }
    return best_rr;
}

template <int KT, int CT, int T>
int launch_sweep2_t(const AcqParams& P, const BleachWs& ws, cudaStream_t st) {
    constexpr int TC = 8 * CT;
    const int PR = sweep_pitch(TC, P.KP);
    size_t smem = ((size_t)(P.K + 2 * T - 1) * PR + (size_t)P.K * P.KP + 4 * T * TC + 2 * s2_evcap(T) + 4 +
                   2 * (size_t)(P.K + 2 * T - 1) + 8 * 64 / 4) * sizeof(float);
    if (const char* e = getenv("STED_B200_S2_PAD_SMEM")) smem += (size_t)atoi(e);     // occupancy experiments only
    int rc;
    if ((rc = set_smem(sweep2_kernel<KT, CT, T>, smem))) return rc;
    static thread_local int slots = 0;                             // concurrent CTAs on this device
    if (!slots) {
        int dev = 0, sms = 0, per_sm = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sweep2_kernel<KT, CT, T>, S2_THREADS, smem));
        slots = sms * (per_sm > 0 ? per_sm : 1);
    }
    const int strips = (P.W + TC - 1) / TC;
    int RR = pick_row_range(P.B, strips, P.H, slots);
    RR = (RR + T - 1) / T * T;                                     // whole groups, except at the end of a range
    dim3 grid(strips, (P.H + RR - 1) / RR, P.B);
    sweep2_kernel<KT, CT, T><<<grid, S2_THREADS, smem, st>>>(P, ws, RR);
    CUDA_TRY(cudaGetLastError());
    return STED_OK;
}

template <int T>
int launch_sweep2_g(const AcqParams& P, const BleachWs& ws, cudaStream_t st) {
    const bool wide = P.W > SWEEP_WIDE_MIN;
    if (P.K == 59) return wide ? launch_sweep2_t<59, 16, T>(P, ws, st) : launch_sweep2_t<59, 8, T>(P, ws, st);
    return wide ? launch_sweep2_t<0, 16, T>(P, ws, st) : launch_sweep2_t<0, 8, T>(P, ws, st);
}

template <int KT, int CT>
int launch_sweep3_t(const AcqParams& P, const BleachWs& ws, cudaStream_t st) {
    constexpr int TC = 8 * CT;
    size_t smem = sweep3_smem_floats(P.K, P.KP, TC) * sizeof(float);
    if (const char* e = getenv("STED_B200_S2_PAD_SMEM")) smem += (size_t)atoi(e);     // occupancy experiments only
    int rc;
    if ((rc = set_smem(sweep3_kernel<KT, CT>, smem))) return rc;
    static thread_local int slots = 0;                             // concurrent CTAs on this device
    if (!slots) {
        int dev = 0, sms = 0, per_sm = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sweep3_kernel<KT, CT>, S3_THREADS, smem));
        slots = sms * (per_sm > 0 ? per_sm : 1);
    }
    const int strips = (P.W + TC - 1) / TC;
    const int RR = pick_row_range(P.B, strips, P.H, slots);
    dim3 grid(strips, (P.H + RR - 1) / RR, P.B);
    sweep3_kernel<KT, CT><<<grid, S3_THREADS, smem, st>>>(P, ws, RR);
    CUDA_TRY(cudaGetLastError());
    return STED_OK;
}

int launch_sweep3(const AcqParams& P, const BleachWs& ws, cudaStream_t st) {
    const bool wide = P.W > SWEEP_WIDE_MIN;
    if (P.K == 59) return wide ? launch_sweep3_t<59, 16>(P, ws, st) : launch_sweep3_t<59, 8>(P, ws, st);
    return wide ? launch_sweep3_t<0, 16>(P, ws, st) : launch_sweep3_t<0, 8>(P, ws, st);
}

#ifndef SWEEP_GROUP
#define SWEEP_GROUP 1          // scan rows gathered between two barriers (measured: 1 -> 3.89 ms, 2 -> 4.16 ms, 4 -> 4.58 ms)
#endif
int launch_sweep2(const AcqParams& P, const BleachWs& ws, cudaStream_t st) {
    int T = SWEEP_GROUP;
    const char* grp = getenv("STED_B200_ROW_GROUP");
    if (!grp && getenv("STED_B200_SCAN_V3")) return launch_sweep3(P, ws, st);    // third generation: opt-in (8 % slower, see K3c)
    if (const char* e = grp) T = atoi(e);                                      // tuning knob
    switch (T) {
        case 2: return launch_sweep2_g<2>(P, ws, st);
        case 3: return launch_sweep2_g<3>(P, ws, st);
        case 4: return launch_sweep2_g<4>(P, ws, st);
        default: return launch_sweep2_g<1>(P, ws, st);
    }
}

// workspace pointers of the stochastic scan; fails when the buffer cannot even hold the fixed part
int bind_workspace(const AcqParams& P, void* workspace, size_t workspace_bytes, BleachWs& ws) {
    const size_t fixed = bleach_ws_bytes(P.B, P.H, 0);
    if (!workspace || workspace_bytes < fixed + 12)
        return fail(STED_EINVAL, "stochastic bleaching needs a workspace of at least %zu bytes (see sted_bleach_workspace_bytes)",
                    fixed + 12);
    if (((uintptr_t)workspace & 7u) != 0) return fail(STED_EINVAL, "workspace_dev must be 8-byte aligned");
    const int n = P.B * P.H;
    ws.status = reinterpret_cast<int*>(workspace);
    ws.cnt = ws.status + 4;
    ws.off = ws.cnt + n;
    ws.cap = (workspace_bytes - fixed) / (sizeof(uint2) + sizeof(uint32_t));       // `fixed` includes the 16 bytes of slack
    if (ws.cap > 0x7FFFFFFFull) ws.cap = 0x7FFFFFFFull;
    ws.stage = reinterpret_cast<uint2*>(ws.off + n + 4);           // 16 + (2n + 4) * 4 bytes in: 8-byte aligned
    ws.ev = reinterpret_cast<uint32_t*>(ws.stage + ws.cap);
    return STED_OK;
}

// the death schedule of a stochastic scan: drawn, counted per (environment, scan row) and bucketed (second-generation
// schedule; the round-1 schedule draws twice and is only reachable through launch_sweep)
int launch_presample(const AcqParams& P, const BleachWs& ws, void* workspace, cudaStream_t st) {
    const int n = P.B * P.H;
    CUDA_TRY(cudaMemsetAsync(workspace, 0, 16 + (size_t)n * sizeof(int), st));
    const size_t psmem = ((size_t)(2 * P.K + 1) * P.KP + PS_ITEMS) * sizeof(float);
    const bool whole = (P.flags & STED_FRAME_MOLECULES) != 0;
    const size_t region = whole ? (size_t)P.Hp * P.Wp : (size_t)P.H * P.W;
    dim3 pgrid((unsigned)((region + PS_CHUNK - 1) / PS_CHUNK), P.B);
    presample_kernel<2><<<pgrid, PS_THREADS, psmem, st>>>(P, ws);
    CUDA_TRY(cudaGetLastError());
    bucket_scan_kernel<<<1, 1024, 0, st>>>(ws, n);
    CUDA_TRY(cudaGetLastError());
    scatter_events_kernel<<<148 * 8, 256, 0, st>>>(ws);
    CUDA_TRY(cudaGetLastError());
    return STED_OK;
}

int launch_sweep(const AcqParams& P, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    BleachWs ws{};
    if (!(P.flags & STED_SAMPLE_BLEACH)) return launch_sweep_m<false>(P, ws, st);
    int rc = bind_workspace(P, workspace, workspace_bytes, ws);
    if (rc) return rc;
    if (P.flags & STED_PRESAMPLED) return launch_sweep2(P, ws, st);   // sted_bleach_presample filled the workspace
    const int n = P.B * P.H;
    const bool v1 = getenv("STED_B200_SCAN_V1") != nullptr;         // round-1 schedule, kept for A/B timing and tests
    if (v1) {
        CUDA_TRY(cudaMemsetAsync(workspace, 0, 16 + (size_t)n * sizeof(int), st));
        const size_t psmem = ((size_t)(2 * P.K + 1) * P.KP + PS_ITEMS) * sizeof(float);
        const bool whole = (P.flags & STED_FRAME_MOLECULES) != 0;
        const size_t region = whole ? (size_t)P.Hp * P.Wp : (size_t)P.H * P.W;
        dim3 pgrid((unsigned)((region + PS_CHUNK - 1) / PS_CHUNK), P.B);
        presample_kernel<0><<<pgrid, PS_THREADS, psmem, st>>>(P, ws);
        CUDA_TRY(cudaGetLastError());
        bucket_scan_kernel<<<1, 1024, 0, st>>>(ws, n);
        CUDA_TRY(cudaGetLastError());
        presample_kernel<1><<<pgrid, PS_THREADS, psmem, st>>>(P, ws);
        CUDA_TRY(cudaGetLastError());
        return launch_sweep_m<true>(P, ws, st);
    }
    if ((rc = launch_presample(P, ws, workspace, st))) return rc;
    return launch_sweep2(P, ws, st);
}

int launch_detect(const AcqParams& P, cudaStream_t st) {
    const size_t hw = (size_t)P.H * P.W;
    const int per_launch = (int)std::max<size_t>(1, std::min<size_t>((size_t)P.B, 0x7fffffffull / hw));   // 32-bit pixel indices
    for (int b0 = 0; b0 < P.B; b0 += per_launch) {
        AcqParams Q = P;
        Q.B = std::min(per_launch, P.B - b0);
        Q.signal = P.signal + (size_t)b0 * hw;
        Q.pdt = P.pdt + b0;
        Q.env_base = P.env_base + (uint32_t)b0;
        const size_t npx = ((hw + 1) / 2) * Q.B;                 // one thread per pair of pixels
        const size_t want = (npx + DET_THREADS - 1) / DET_THREADS;
        const unsigned blocks = (unsigned)(want < 148u * 32u ? want : 148u * 32u);
        detect_kernel<<<blocks, DET_THREADS, 0, st>>>(Q);
        CUDA_TRY(cudaGetLastError());
    }
    return STED_OK;
}

// growable device workspace for the *_host entry points
struct Workspace {
    void* dev[10] = {nullptr};
    size_t cap[10] = {0};
    int device = -1;                         // the scratch lives on the device that was current at the first call
    ~Workspace() {                           // thread exit: give the scratch back (errors at process teardown are moot)
        for (void* p : dev) if (p) cudaFree(p);
    }
    int bind() {                             // one device per host thread: a later cudaSetDevice(other) is refused, not mis-launched
        int cur = 0;
        CUDA_TRY(cudaGetDevice(&cur));
        if (device < 0) device = cur;
        if (device != cur)
            return fail(STED_EINVAL, "this host thread first called the host API on device %d and now runs on device %d: "
                                     "use one host thread per device", device, cur);
        return STED_OK;
    }
    int ensure(int slot, size_t bytes) {
        if (cap[slot] >= bytes) return STED_OK;
        if (dev[slot]) cudaFree(dev[slot]);
        dev[slot] = nullptr; cap[slot] = 0;
        CUDA_TRY(cudaMalloc(&dev[slot], bytes));
        cap[slot] = bytes;
        return STED_OK;
    }
};
// one workspace and one stream set per host thread: concurrent sted_acquire_host calls from different threads do not
// serialise on each other (each thread pays for its own device scratch)
thread_local Workspace g_ws;

__device__ __forceinline__ float host_to_count(float v) { return v; }
__device__ __forceinline__ float host_to_count(uint16_t v) { return (float)v; }
__device__ __forceinline__ void count_to_host(float v, float* o) { *o = v; }
__device__ __forceinline__ void count_to_host(float v, uint16_t* o) { *o = (uint16_t)fminf(fmaxf(rintf(v), 0.f), 65535.f); }   // saturating

// host-format ROI (float32 or uint16 counts) -> zero-framed float32 map, and back
template <class T>
__global__ void pad_kernel(const T* __restrict__ roi, float* __restrict__ padded, int B, int H, int W, int p,
                           int Hp, int Wpitch) {
    const size_t n = (size_t)B * Hp * Wpitch;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % Wpitch);
        const int y = (int)((i / Wpitch) % Hp);
        const int b = (int)(i / ((size_t)Wpitch * Hp));
        const int r = y - p, c = x - p;
        padded[i] = (r >= 0 && r < H && c >= 0 && c < W) ? host_to_count(roi[((size_t)b * H + r) * W + c]) : 0.f;
    }
}

template <class T>
__global__ void unpad_kernel(const float* __restrict__ padded, T* __restrict__ roi, int B, int H, int W, int p,
                             int Hp, int Wpitch) {
    const size_t n = (size_t)B * H * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % W);
        const int r = (int)((i / W) % H);
        const int b = (int)(i / ((size_t)W * H));
        count_to_host(padded[((size_t)b * Hp + r + p) * Wpitch + c + p], roi + i);
    }
}

// photon counts float32 -> uint16 (saturating) for the STED_HOST_U16 transport
__global__ void counts_to_u16_kernel(const float* __restrict__ in, uint16_t* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        count_to_host(in[i], out + i);
}

}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

const char* sted_version(void) { return "sted_b200 0.1 (sm_100a)"; }
const char* sted_last_error(void) { return sted_err_buf; }

int sted_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int sted_make_effective(const double* psf_cache_dev, const StedFluoParams* fluo_dev, const double* p_ex_dev,
                        const double* p_sted_dev, const double* pdt_dev, int B, int K, float* tables_dev,
                        double* kbleach_dev, void* stream) {
    if (!psf_cache_dev || !fluo_dev || !p_ex_dev || !p_sted_dev || !pdt_dev || !tables_dev)
        return fail(STED_EINVAL, "null pointer argument");
    if (B <= 0) return fail(STED_EINVAL, "B must be positive");
    if (K < 1 || (K & 1) == 0) return fail(STED_EINVAL, "K must be odd (got %d)", K);
    if (K > STED_MAX_K) return fail(STED_EUNSUPPORTED, "K=%d exceeds STED_MAX_K=%d", K, STED_MAX_K);
    int rc = check_device();
    if (rc) return rc;
    const size_t me_smem = 2 * (size_t)K * K * sizeof(double);
    CUDA_TRY(cudaFuncSetAttribute(make_effective_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)me_smem));
    make_effective_kernel<<<B, 256, me_smem, (cudaStream_t)stream>>>(psf_cache_dev, fluo_dev, p_ex_dev, p_sted_dev, pdt_dev,
                                                                      K, STED_KPITCH(K), tables_dev, kbleach_dev);
    CUDA_TRY(cudaGetLastError());
    return STED_OK;
}

int sted_acquire(const float* dmap_in_dev, float* dmap_out_dev, const float* tables_dev, const double* pdt_dev,
                 int B, int H, int W, int K, uint32_t flags, const StedDetectorParams* det, double lambda_fluo,
                 uint64_t seed, uint64_t offset, int64_t env_base, float* intensity_dev, float* signal_dev,
                 void* workspace_dev, uint64_t workspace_bytes, void* stream) {
    AcqParams P;
    int rc = make_params(P, dmap_in_dev, dmap_out_dev, tables_dev, pdt_dev, B, H, W, K, flags, det, lambda_fluo, seed,
                         offset, env_base, intensity_dev, signal_dev);
    if (rc) return rc;
    if ((rc = check_device())) return rc;
    rc = (flags & STED_BLEACH) ? launch_sweep(P, workspace_dev, (size_t)workspace_bytes, (cudaStream_t)stream)
                               : launch_gather(P, (cudaStream_t)stream);
    if (rc) return rc;
    if (!(flags & STED_NO_DETECT) && (rc = launch_detect(P, (cudaStream_t)stream))) return rc;
    return STED_OK;
}

int sted_bleach_presample(const float* dmap_in_dev, const float* tables_dev, const double* pdt_dev, int B, int H, int W, int K,
                          uint32_t flags, uint64_t seed, uint64_t offset, int64_t env_base, void* workspace_dev,
                          uint64_t workspace_bytes, void* stream) {
    if ((flags & (STED_BLEACH | STED_SAMPLE_BLEACH)) != (STED_BLEACH | STED_SAMPLE_BLEACH))
        return fail(STED_EINVAL, "sted_bleach_presample is the first half of a STED_BLEACH | STED_SAMPLE_BLEACH acquisition");
    AcqParams P;
    const StedDetectorParams det{1.0, 1.0, 0.0};                    // the detector plays no part here
    float dummy = 0.f;
    int rc = make_params(P, dmap_in_dev, &dummy, tables_dev, pdt_dev, B, H, W, K, flags, &det, 1.0, seed, offset, env_base,
                         nullptr, &dummy);
    if (rc) return rc;
    if ((rc = check_device())) return rc;
    BleachWs ws{};
    if ((rc = bind_workspace(P, workspace_dev, (size_t)workspace_bytes, ws))) return rc;
    return launch_presample(P, ws, workspace_dev, (cudaStream_t)stream);
}

int sted_acquire_dwellmaps(const float* dmap_in_dev, float* dmap_out_dev, const float* tables_dev, const double* kbleach_dev,
                           const double* pdt_dev, const float* w_dwell_dev, const float* w_ex_dev, const uint8_t* table_id_dev,
                           int n_tables, int B, int H, int W, int K, uint32_t flags, const StedDetectorParams* det, double lambda_fluo,
                           uint64_t seed, uint64_t offset, int64_t env_base, float* intensity_dev, float* signal_dev, void* stream) {
    AcqParams P;
    int rc = make_params(P, dmap_in_dev, dmap_out_dev, tables_dev, pdt_dev, B, H, W, K, flags, det, lambda_fluo, seed, offset,
                         env_base, intensity_dev, signal_dev);
    if (rc) return rc;
    if (n_tables < 1 || n_tables > 256) return fail(STED_EINVAL, "n_tables must lie in [1, 256]");
    if (n_tables > 1 && !table_id_dev) return fail(STED_EINVAL, "several tables need table_id_dev");
    if ((flags & STED_BLEACH) && !kbleach_dev) return fail(STED_EINVAL, "bleaching needs kbleach_dev (sted_make_effective's second output)");
    if ((rc = check_device())) return rc;
    P.wdt = w_dwell_dev;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t fixed = (size_t)K * K * (sizeof(double) + 2 * sizeof(float));
    const size_t band64 = (size_t)K * P.Wp * sizeof(double), band32 = (size_t)K * P.Wp * sizeof(float);
    const size_t limit = 200 * 1024;
    const double* kb = kbleach_dev;
    if (!kb) {                               // no bleaching: the hazard table is never used, any valid pointer will do
        kb = reinterpret_cast<const double*>(tables_dev);
    }
    if (fixed + band64 <= limit) {
        if ((rc = set_smem(general_scan_kernel<double>, fixed + band64))) return rc;
        general_scan_kernel<double><<<B, GS_THREADS, fixed + band64, st>>>(P, kb, w_dwell_dev, w_ex_dev, table_id_dev, n_tables);
    } else if (fixed + band32 <= limit) {
        if ((rc = set_smem(general_scan_kernel<float>, fixed + band32))) return rc;
        general_scan_kernel<float><<<B, GS_THREADS, fixed + band32, st>>>(P, kb, w_dwell_dev, w_ex_dev, table_id_dev, n_tables);
    } else {
        return fail(STED_EUNSUPPORTED, "image too wide for the general scan (W + K - 1 = %d columns)", P.Wp);
    }
    CUDA_TRY(cudaGetLastError());
    if (!(flags & STED_NO_DETECT) && (rc = launch_detect(P, st))) return rc;
    return STED_OK;
}

uint64_t sted_bleach_workspace_bytes(int B, int H, uint64_t max_events) {
    return (uint64_t)bleach_ws_bytes(B, H, max_events);
}

int sted_bleach_dropped_events(const void* workspace_dev, int* dropped_host, void* stream) {
    if (!workspace_dev || !dropped_host) return fail(STED_EINVAL, "null pointer argument");
    int rc = check_device();
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(dropped_host, workspace_dev, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return STED_OK;
}

int sted_detect(float* signal_dev, const double* pdt_dev, int B, int H, int W, uint32_t flags,
                const StedDetectorParams* det, double lambda_fluo, uint64_t seed, uint64_t offset, int64_t env_base,
                void* stream) {
    if (!signal_dev || !pdt_dev || !det) return fail(STED_EINVAL, "null pointer argument");
    if (B <= 0 || H <= 0 || W <= 0) return fail(STED_EINVAL, "B, H, W must be positive");
    if (!(lambda_fluo > 0)) return fail(STED_EINVAL, "lambda_fluo must be positive");
    int rc = check_device();
    if (rc) return rc;
    AcqParams P{};
    P.signal = signal_dev; P.pdt = pdt_dev; P.B = B; P.H = H; P.W = W;
    P.flags = flags; P.offset = (uint32_t)offset; P.seed = seed ^ (offset >> 32 << 17); P.env_base = (uint32_t)env_base;
    P.inv_e_photon = (float)(lambda_fluo / (kPlanck * kLight));
    P.det_eff = (float)(det->pcef * det->pdef);
    P.background = (float)det->background;
    return launch_detect(P, (cudaStream_t)stream);
}

}  // extern "C" (re-opened below)

namespace {

// Streams / events of the host-buffer entry point: upload, compute and download of different chunks overlap.
struct HostPipe {
    static constexpr int MAXC = 16;
    cudaStream_t s_in = nullptr, s_out = nullptr, s_comp[MAXC];   // one compute stream per chunk: their kernels co-run
    cudaEvent_t e_in[MAXC], e_comp[MAXC];
    int* h_drop = nullptr;               // pinned: dropped-event counter of each chunk
    bool ready = false;
    ~HostPipe() {
        if (!ready) return;
        cudaStreamDestroy(s_in); cudaStreamDestroy(s_out);
        for (int i = 0; i < MAXC; ++i) { cudaStreamDestroy(s_comp[i]); cudaEventDestroy(e_in[i]); cudaEventDestroy(e_comp[i]); }
        cudaFreeHost(h_drop);
    }
    int init(bool low_priority) {
        if (ready) return STED_OK;
        int lo = 0, hi = 0;                  // numerically lower = higher priority
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        const int prio = low_priority ? lo : hi;
        CUDA_TRY(cudaStreamCreateWithPriority(&s_in, cudaStreamNonBlocking, prio));
        CUDA_TRY(cudaStreamCreateWithPriority(&s_out, cudaStreamNonBlocking, prio));
        for (int i = 0; i < MAXC; ++i) {
            CUDA_TRY(cudaStreamCreateWithPriority(&s_comp[i], cudaStreamNonBlocking, prio));
            CUDA_TRY(cudaEventCreateWithFlags(&e_in[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&e_comp[i], cudaEventDisableTiming));
        }
        CUDA_TRY(cudaMallocHost(&h_drop, MAXC * sizeof(int)));
        for (int i = 0; i < MAXC; ++i) h_drop[i] = 0;
        ready = true;
        return STED_OK;
    }
};
thread_local HostPipe g_pipe;
thread_local bool g_host_low_priority = false;

}  // namespace

extern "C" int sted_host_thread_priority(int low) {
    if (g_pipe.ready) return fail(STED_EINVAL, "sted_host_thread_priority must be called before this thread's first sted_acquire_host");
    g_host_low_priority = low != 0;
    return STED_OK;
}

extern "C" int sted_acquire_host(void* dmap_host, const double* psf_cache_host, const StedFluoParams* fluo_host,
                                 const double* p_ex_host, const double* p_sted_host, const double* pdt_host, int B, int H,
                                 int W, int K, uint32_t flags, const StedDetectorParams* det, uint64_t seed, uint64_t offset,
                                 int64_t env_base, float* intensity_host, void* signal_host) {
    if (!dmap_host || !psf_cache_host || !fluo_host || !p_ex_host || !p_sted_host || !pdt_host || !signal_host || !det)
        return fail(STED_EINVAL, "null pointer argument");
    if (B <= 0 || H <= 0 || W <= 0) return fail(STED_EINVAL, "B, H, W must be positive");
    if (K < 1 || (K & 1) == 0) return fail(STED_EINVAL, "K must be odd (got %d)", K);
    if (K > STED_MAX_K) return fail(STED_EUNSUPPORTED, "K=%d exceeds STED_MAX_K=%d", K, STED_MAX_K);
    int rc = check_device();
    if (rc) return rc;
    if ((rc = g_ws.bind())) return rc;
    if ((rc = g_pipe.init(g_host_low_priority))) return rc;
    const int KP = STED_KPITCH(K), Hp = H + K - 1, Wp = W + K - 1, Wpitch = (Wp + 3) & ~3, p = K / 2;
    const size_t px_roi = (size_t)H * W, px_pad = (size_t)Hp * Wpitch;
    const size_t n_roi = (size_t)B * px_roi, n_pad = (size_t)B * px_pad;
    enum { ROI, PADA, PADB, TAB, CACHE, FLUO, SCAL, OUT, WS, SIG16 };
    const bool bleach = (flags & STED_BLEACH) != 0;
    const bool stoch = bleach && (flags & STED_SAMPLE_BLEACH);
    // STED_HOST_U16: maps and photon counts cross the bus as uint16 (half the bytes); counts must stay integers
    const bool u16 = (flags & STED_HOST_U16) != 0;
    if (u16 && ((bleach && !stoch) || !(flags & STED_SAMPLE_PHOTONS)))
        return fail(STED_EINVAL, "STED_HOST_U16 needs integer outputs: STED_SAMPLE_PHOTONS, and STED_SAMPLE_BLEACH when bleaching");
    const size_t esz = u16 ? sizeof(uint16_t) : sizeof(float);
    flags &= ~STED_HOST_U16;
    // chunks: upload of chunk i+1, compute of chunk i and download of chunk i-1 run concurrently
    int nch = B >= 128 ? 8 : (B >= 32 ? 4 : (B >= 8 ? 2 : 1));      // measured on B200, 256 envs x 256^2: 2 -> 3.9, 4 -> 3.4, 8..16 -> 3.2 ms/call
    if (const char* e = getenv("STED_B200_HOST_CHUNKS")) {        // tuning knob: 1..16 chunks
        const int v = atoi(e);
        if (v >= 1 && v <= HostPipe::MAXC && v <= B) nch = v;
    }
    const int cb = (B + nch - 1) / nch;
    nch = (B + cb - 1) / cb;                                      // chunks actually issued (e.g. B = 17, 16 asked: 9 chunks of 2)
    // scratch of the stochastic scan: starts at 4 deaths per pixel and grows when a scan reports dropped events
    static thread_local unsigned long long s_max_events = 0;
    if (stoch && s_max_events < 4ull * cb * px_roi) s_max_events = 4ull * cb * px_roi;
    if ((rc = g_ws.ensure(ROI, n_roi * 4)) || (rc = g_ws.ensure(PADA, n_pad * 4)) ||
        (rc = g_ws.ensure(PADB, bleach ? n_pad * 4 : 16)) ||
        (rc = g_ws.ensure(TAB, (size_t)B * STED_TABLES * K * KP * 4)) || (rc = g_ws.ensure(CACHE, 3ull * K * K * 8)) ||
        (rc = g_ws.ensure(FLUO, (size_t)B * sizeof(StedFluoParams))) || (rc = g_ws.ensure(SCAL, 3ull * B * 8)) ||
        (rc = g_ws.ensure(OUT, 2 * n_roi * 4)) || (rc = g_ws.ensure(SIG16, u16 ? n_roi * 2 : 16)))
        return rc;
    uint16_t* d_sig16 = (uint16_t*)g_ws.dev[SIG16];
    char* h_map = (char*)dmap_host;
    char* h_sig = (char*)signal_host;
    char* d_roi_b = (char*)g_ws.dev[ROI];
    float* d_roi = (float*)g_ws.dev[ROI];
    float* d_a = (float*)g_ws.dev[PADA];
    float* d_b = (float*)g_ws.dev[PADB];
    float* d_tab = (float*)g_ws.dev[TAB];
    double* d_cache = (double*)g_ws.dev[CACHE];
    StedFluoParams* d_fluo = (StedFluoParams*)g_ws.dev[FLUO];
    double* d_scal = (double*)g_ws.dev[SCAL];
    float* d_int = (float*)g_ws.dev[OUT];
    float* d_sig = d_int + n_roi;
    cudaStream_t s_in = g_pipe.s_in, s_out = g_pipe.s_out;

    CUDA_TRY(cudaMemcpyAsync(d_cache, psf_cache_host, 3ull * K * K * 8, cudaMemcpyHostToDevice, s_in));
    CUDA_TRY(cudaMemcpyAsync(d_fluo, fluo_host, (size_t)B * sizeof(StedFluoParams), cudaMemcpyHostToDevice, s_in));
    CUDA_TRY(cudaMemcpyAsync(d_scal, p_ex_host, (size_t)B * 8, cudaMemcpyHostToDevice, s_in));
    CUDA_TRY(cudaMemcpyAsync(d_scal + B, p_sted_host, (size_t)B * 8, cudaMemcpyHostToDevice, s_in));
    CUDA_TRY(cudaMemcpyAsync(d_scal + 2 * B, pdt_host, (size_t)B * 8, cudaMemcpyHostToDevice, s_in));

    bool uploaded = false;
    for (int attempt = 0;; ++attempt) {
        size_t ws_bytes = 0;
        if (stoch) {
            ws_bytes = (bleach_ws_bytes(cb, H, s_max_events) + 255) & ~(size_t)255;      // one scratch per chunk
            if ((rc = g_ws.ensure(WS, ws_bytes * nch))) return rc;
        }
        for (int c = 0; c < nch; ++c) {
            const int b0 = c * cb, nb = (B - b0 < cb) ? B - b0 : cb;
            if (nb <= 0) break;
            const size_t o_roi = (size_t)b0 * px_roi, o_pad = (size_t)b0 * px_pad;
            cudaStream_t s_comp = g_pipe.s_comp[c];
            char* ws_c = stoch ? (char*)g_ws.dev[WS] + (size_t)c * ws_bytes : nullptr;
            if (!uploaded) {
                CUDA_TRY(cudaMemcpyAsync(d_roi_b + o_roi * esz, h_map + o_roi * esz, (size_t)nb * px_roi * esz, cudaMemcpyHostToDevice, s_in));
                CUDA_TRY(cudaEventRecord(g_pipe.e_in[c], s_in));
                CUDA_TRY(cudaStreamWaitEvent(s_comp, g_pipe.e_in[c], 0));
                if (u16) pad_kernel<<<296, 256, 0, s_comp>>>((const uint16_t*)d_roi_b + o_roi, d_a + o_pad, nb, H, W, p, Hp, Wpitch);
                else pad_kernel<<<296, 256, 0, s_comp>>>(d_roi + o_roi, d_a + o_pad, nb, H, W, p, Hp, Wpitch);
                CUDA_TRY(cudaGetLastError());
                if ((rc = sted_make_effective(d_cache, d_fluo + b0, d_scal + b0, d_scal + B + b0, d_scal + 2 * B + b0, nb, K,
                                              d_tab + (size_t)b0 * STED_TABLES * K * KP, nullptr, s_comp)))
                    return rc;
            }
            if ((rc = sted_acquire(d_a + o_pad, bleach ? d_b + o_pad : nullptr, d_tab + (size_t)b0 * STED_TABLES * K * KP,
                                   d_scal + 2 * B + b0, nb, H, W, K, flags, det, fluo_host[0].lambda_fluo, seed, offset,
                                   env_base + b0, intensity_host ? d_int + o_roi : nullptr, d_sig + o_roi,
                                   ws_c, ws_bytes, s_comp)))
                return rc;
            g_pipe.h_drop[c] = 0;
            if (stoch) CUDA_TRY(cudaMemcpyAsync(&g_pipe.h_drop[c], ws_c, sizeof(int), cudaMemcpyDeviceToHost, s_comp));
            if (bleach) {
                if (u16) unpad_kernel<<<296, 256, 0, s_comp>>>(d_b + o_pad, (uint16_t*)d_roi_b + o_roi, nb, H, W, p, Hp, Wpitch);
                else unpad_kernel<<<296, 256, 0, s_comp>>>(d_b + o_pad, d_roi + o_roi, nb, H, W, p, Hp, Wpitch);
                CUDA_TRY(cudaGetLastError());
            }
            if (u16) {
                counts_to_u16_kernel<<<296, 256, 0, s_comp>>>(d_sig + o_roi, d_sig16 + o_roi, (size_t)nb * px_roi);
                CUDA_TRY(cudaGetLastError());
            }
            CUDA_TRY(cudaEventRecord(g_pipe.e_comp[c], s_comp));
            CUDA_TRY(cudaStreamWaitEvent(s_out, g_pipe.e_comp[c], 0));
            if (bleach) CUDA_TRY(cudaMemcpyAsync(h_map + o_roi * esz, d_roi_b + o_roi * esz, (size_t)nb * px_roi * esz, cudaMemcpyDeviceToHost, s_out));
            if (intensity_host)
                CUDA_TRY(cudaMemcpyAsync(intensity_host + o_roi, d_int + o_roi, (size_t)nb * px_roi * 4, cudaMemcpyDeviceToHost, s_out));
            if (u16) CUDA_TRY(cudaMemcpyAsync(h_sig + o_roi * 2, d_sig16 + o_roi, (size_t)nb * px_roi * 2, cudaMemcpyDeviceToHost, s_out));
            else CUDA_TRY(cudaMemcpyAsync(h_sig + o_roi * 4, d_sig + o_roi, (size_t)nb * px_roi * 4, cudaMemcpyDeviceToHost, s_out));
        }
        for (int c = 0; c < nch; ++c) CUDA_TRY(cudaStreamSynchronize(g_pipe.s_comp[c]));
        CUDA_TRY(cudaStreamSynchronize(s_out));
        uploaded = true;                 // the padded inputs stay on the device: a re-run needs no new upload
        long long dropped = 0;
        for (int c = 0; c < nch; ++c) dropped += g_pipe.h_drop[c];
        if (!stoch || dropped == 0) break;
        if (attempt >= 8) return fail(STED_ECUDA, "stochastic scan kept overflowing its event scratch");
        s_max_events = s_max_events * 2 + (unsigned long long)dropped;     // same seed => same deaths on the re-run
    }
    return STED_OK;
}

extern "C" {

int sted_sample_poisson(const float* lambda_dev, int64_t n, uint64_t seed, uint64_t offset, float* out_dev, void* stream) {
    if (!lambda_dev || !out_dev || n <= 0) return fail(STED_EINVAL, "bad argument");
    int rc = check_device();
    if (rc) return rc;
    poisson_test_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(lambda_dev, n, seed, (uint32_t)offset, out_dev);
    CUDA_TRY(cudaGetLastError());
    return STED_OK;
}

int sted_sample_binomial(const float* n_dev, const float* q_dev, int64_t n, uint64_t seed, uint64_t offset,
                         float* out_dev, void* stream) {
    if (!n_dev || !q_dev || !out_dev || n <= 0) return fail(STED_EINVAL, "bad argument");
    int rc = check_device();
    if (rc) return rc;
    binomial_test_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n_dev, q_dev, n, seed, (uint32_t)offset, out_dev);
    CUDA_TRY(cudaGetLastError());
    return STED_OK;
}

void sted_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out_host[4]) {
    Philox4 c{ctr[0], ctr[1], ctr[2], ctr[3]};
    Philox4 r = philox4x32_10(c, key[0], key[1]);
    out_host[0] = r.x; out_host[1] = r.y; out_host[2] = r.z; out_host[3] = r.w;
}

int sted_debug_phase_cycles(unsigned long long out_host[8], int reset) {
    int rc = check_device();
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyFromSymbol(out_host, g_phase_cycles, sizeof(unsigned long long) * 8));
    if (reset) { unsigned long long z[8] = {0}; CUDA_TRY(cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z))); }
    return STED_OK;
}

int sted_fma_peak(double* tflops, void* stream) {
    if (!tflops) return fail(STED_EINVAL, "null pointer argument");
    int rc = check_device();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    float* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 4));
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, iters = 1 << 15;
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    fma_peak_kernel<<<blocks, 256, 0, st>>>(d, 1 << 10);
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(cudaEventRecord(e0, st));
        fma_peak_kernel<<<blocks, 256, 0, st>>>(d, iters);
        CUDA_TRY(cudaEventRecord(e1, st));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 16 * iters * 256.0 * blocks / (ms * 1e-3) * 1e-12;
        if (fl > best) best = fl;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *tflops = best;
    return STED_OK;
}

}  // extern "C"

// ==========================================================================================
// Sessions: a structure stays on the device across the acquisitions made of it
// ==========================================================================================
// gym-sted images every structure three or four times (confocal, STED + bleaching, confocal; mosted_env.py:222-234)
// and pySTED keeps the datamap object alive in between.  sted_acquire_host re-uploads the map on every call; a session
// uploads it once, keeps the padded map (and its bleached successor) resident, and queues uploads, acquisitions and
// downloads asynchronously on its own streams: host code only waits in sted_session_sync.
struct StedSession {
    static constexpr int NSIG = 4;           // result buffers in flight
    int device = -1, B = 0, H = 0, W = 0, K = 0, KP = 0, Hp = 0, Wp = 0, Wpitch = 0;
    size_t n_roi = 0, n_pad = 0;
    // s_comp carries the gathers and scans (and everything that changes the maps); the work around them runs beside them:
    // s_prep the tables and the death schedule of an acquisition (under the acquisition queued before it), s_post the
    // detector and the uint16 conversion (under the acquisition queued after it)
    cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr, s_prep = nullptr, s_post = nullptr;
    cudaEvent_t e_up = nullptr, e_pad = nullptr, e_comp = nullptr, e_unpad = nullptr, e_out[NSIG] = {nullptr}, e_mapout = nullptr,
                e_lent = nullptr;            // another session has finished copying this session's map (sted_session_adopt)
    cudaEvent_t e_prep[NSIG] = {nullptr}, e_main[NSIG] = {nullptr}, e_det[NSIG] = {nullptr};
    static constexpr int NT = 64;            // timing brackets kept (ring)
    cudaEvent_t e_t0[NT] = {nullptr}, e_t1[NT] = {nullptr};        // around the gather / scan kernel of an acquisition, on s_comp
    cudaEvent_t e_join[5] = {nullptr};
    // tables of the most recent non-bleaching parameter set: gym-sted's confocal acquisitions always use the same
    // imaging parameters (generate_params() defaults, mosted_env.py:219), so three of the four acquisitions of a step
    // find their tables here instead of rebuilding them
    float* d_tab_keep = nullptr;
    double* d_scal_keep = nullptr;
    double* h_scal_keep = nullptr;           // host copy the next call's parameters are compared with
    cudaEvent_t e_keep = nullptr, e_keep_used = nullptr;
    bool keep_valid = false;
    long long n_acq = 0;
    bool serial = false;                     // STED_B200_SESSION_SERIAL: everything on s_comp (A/B timing)
    cudaEvent_t e_map = nullptr;             // the current map is final (upload / adopt / bleaching acquisition, on s_comp)
    cudaEvent_t e_wsfree = nullptr;          // the last scan has consumed the event scratch
    char* d_stage = nullptr;                 // un-padded upload / download staging [B][H][W] (4 bytes per pixel)
    float* d_map[2] = {nullptr, nullptr};    // padded maps, ping-pong
    float* d_tab[NSIG] = {nullptr};          // tables of each queued acquisition
    double* d_cache = nullptr;
    StedFluoParams* d_fluo = nullptr;
    double* d_scal[NSIG] = {nullptr};        // p_ex | p_sted | pdt of each queued acquisition
    double* h_scal[NSIG] = {nullptr};        // pinned copies (the caller's arrays may die after the call returns)
    float* d_sig[NSIG] = {nullptr};
    uint16_t* d_sig16[NSIG] = {nullptr};
    float* d_int[NSIG] = {nullptr};
    void* d_ws = nullptr;
    size_t ws_bytes = 0;
    unsigned long long max_events = 0;
    int* h_drop = nullptr;                   // pinned [NSIG]
    int cur = 0, next_sig = 0;
    bool fluo_set = false, map_set = false, stage_busy = false;
};

namespace {

int session_bind(StedSession* s) {
    if (!s) return fail(STED_EINVAL, "null session");
    int cur = 0;
    CUDA_TRY(cudaGetDevice(&cur));
    if (cur != s->device)
        return fail(STED_EINVAL, "session was created on device %d but the calling thread runs on device %d", s->device, cur);
    return STED_OK;
}

int session_reserve(StedSession* s, unsigned long long max_events) {
    if (max_events <= s->max_events) return STED_OK;
    s->max_events = max_events;
    const size_t need = (bleach_ws_bytes(s->B, s->H, s->max_events) + 255) & ~(size_t)255;
    if (need > s->ws_bytes) {
        CUDA_TRY(cudaStreamSynchronize(s->s_prep));
        CUDA_TRY(cudaStreamSynchronize(s->s_comp));
        if (s->d_ws) cudaFree(s->d_ws);
        s->d_ws = nullptr; s->ws_bytes = 0;
        CUDA_TRY(cudaMalloc(&s->d_ws, need));
        s->ws_bytes = need;
    }
    return STED_OK;
}

void session_free(StedSession* s) {
    if (!s) return;
    for (cudaStream_t st : {s->s_in, s->s_prep, s->s_comp, s->s_post, s->s_out}) if (st) cudaStreamSynchronize(st);
    cudaFree(s->d_stage); cudaFree(s->d_map[0]); cudaFree(s->d_map[1]); cudaFree(s->d_cache);
    cudaFree(s->d_tab_keep); cudaFree(s->d_scal_keep);
    if (s->h_scal_keep) free(s->h_scal_keep);
    for (cudaEvent_t e : {s->e_keep, s->e_keep_used}) if (e) cudaEventDestroy(e);
    cudaFree(s->d_fluo); cudaFree(s->d_ws);
    for (int i = 0; i < StedSession::NSIG; ++i) {
        cudaFree(s->d_scal[i]); cudaFree(s->d_sig[i]); cudaFree(s->d_sig16[i]); cudaFree(s->d_int[i]); cudaFree(s->d_tab[i]);
        if (s->h_scal[i]) cudaFreeHost(s->h_scal[i]);
        for (cudaEvent_t e : {s->e_out[i], s->e_prep[i], s->e_main[i], s->e_det[i]}) if (e) cudaEventDestroy(e);
    }
    if (s->h_drop) cudaFreeHost(s->h_drop);
    for (cudaEvent_t e : {s->e_up, s->e_pad, s->e_comp, s->e_unpad, s->e_mapout, s->e_lent, s->e_map, s->e_wsfree}) if (e) cudaEventDestroy(e);
    for (int i = 0; i < StedSession::NT; ++i) { if (s->e_t0[i]) cudaEventDestroy(s->e_t0[i]); if (s->e_t1[i]) cudaEventDestroy(s->e_t1[i]); }
    for (cudaEvent_t e : s->e_join) if (e) cudaEventDestroy(e);
    if (s->serial) s->s_prep = s->s_post = nullptr;           // aliases of s_comp
    for (cudaStream_t st : {s->s_in, s->s_prep, s->s_comp, s->s_post, s->s_out}) if (st) cudaStreamDestroy(st);
    delete s;
}

}  // namespace

extern "C" {

int sted_session_create(int B, int H, int W, int K, const double* psf_cache_host, int low_priority, StedSession** out) {
    if (!out || !psf_cache_host) return fail(STED_EINVAL, "null pointer argument");
    *out = nullptr;
    if (B <= 0 || H <= 0 || W <= 0) return fail(STED_EINVAL, "B, H, W must be positive");
    if (K < 1 || (K & 1) == 0) return fail(STED_EINVAL, "K must be odd (got %d)", K);
    if (K > STED_MAX_K) return fail(STED_EUNSUPPORTED, "K=%d exceeds STED_MAX_K=%d", K, STED_MAX_K);
    if (B > 65535) return fail(STED_EUNSUPPORTED, "B=%d exceeds 65535 environments per session", B);
    int rc = check_device();
    if (rc) return rc;
    StedSession* s = new StedSession();
    auto bail = [&](int code) { session_free(s); return code; };
#define S_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return bail(fail(STED_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e))); } while (0)
    S_TRY(cudaGetDevice(&s->device));
    s->B = B; s->H = H; s->W = W; s->K = K; s->KP = STED_KPITCH(K);
    s->Hp = H + K - 1; s->Wp = W + K - 1; s->Wpitch = (s->Wp + 3) & ~3;
    s->n_roi = (size_t)B * H * W; s->n_pad = (size_t)B * s->Hp * s->Wpitch;
    int lo = 0, hi = 0;
    S_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    // Priorities (numerically lower = higher).  The gather / scan stream sits one step BELOW the streams of the light ends
    // of an acquisition - upload + padding, tables + death schedule, detector, downloads: the block scheduler hands free
    // slots to the oldest grid of the highest priority, so at equal priority the tables and the death schedule of the next
    // acquisition only started in the tail wave of the running gather instead of under it.  A low-priority session puts
    // its gather / scan at the bottom of the range; its light ends stay on top, or its detector would wait for the end of
    // another session's saturating kernel and its download for the end of the step.
    const int prio = low_priority ? lo : (hi + 1 <= lo ? hi + 1 : hi);
    S_TRY(cudaStreamCreateWithPriority(&s->s_in, cudaStreamNonBlocking, hi));
    S_TRY(cudaStreamCreateWithPriority(&s->s_comp, cudaStreamNonBlocking, prio));
    S_TRY(cudaStreamCreateWithPriority(&s->s_out, cudaStreamNonBlocking, hi));
    s->serial = getenv("STED_B200_SESSION_SERIAL") != nullptr;
    if (s->serial) {
        s->s_prep = s->s_post = s->s_comp;
    } else {
        S_TRY(cudaStreamCreateWithPriority(&s->s_prep, cudaStreamNonBlocking, hi));
        S_TRY(cudaStreamCreateWithPriority(&s->s_post, cudaStreamNonBlocking, hi));
    }
    for (int i = 0; i < StedSession::NT; ++i) { S_TRY(cudaEventCreate(&s->e_t0[i])); S_TRY(cudaEventCreate(&s->e_t1[i])); }
    for (cudaEvent_t& e : s->e_join) S_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (cudaEvent_t* e : {&s->e_keep, &s->e_keep_used}) S_TRY(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    S_TRY(cudaMalloc(&s->d_tab_keep, (size_t)B * STED_TABLES * K * s->KP * 4));
    S_TRY(cudaMalloc(&s->d_scal_keep, 3ull * B * 8));
    s->h_scal_keep = (double*)malloc(3ull * B * 8);
    if (!s->h_scal_keep) return bail(fail(STED_ECUDA, "out of host memory"));
    for (cudaEvent_t* e : {&s->e_up, &s->e_pad, &s->e_comp, &s->e_unpad, &s->e_mapout, &s->e_lent, &s->e_map, &s->e_wsfree})
        S_TRY(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    S_TRY(cudaMalloc(&s->d_stage, s->n_roi * 4));
    S_TRY(cudaMalloc(&s->d_map[0], s->n_pad * 4));
    S_TRY(cudaMalloc(&s->d_map[1], s->n_pad * 4));
    S_TRY(cudaMemsetAsync(s->d_map[0], 0, s->n_pad * 4, s->s_comp));
    S_TRY(cudaMemsetAsync(s->d_map[1], 0, s->n_pad * 4, s->s_comp));
    S_TRY(cudaMalloc(&s->d_cache, 3ull * K * K * 8));
    S_TRY(cudaMalloc(&s->d_fluo, (size_t)B * sizeof(StedFluoParams)));
    S_TRY(cudaMallocHost(&s->h_drop, StedSession::NSIG * sizeof(int)));
    for (int i = 0; i < StedSession::NSIG; ++i) {
        s->h_drop[i] = 0;
        for (cudaEvent_t* e : {&s->e_out[i], &s->e_prep[i], &s->e_main[i], &s->e_det[i]}) S_TRY(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        S_TRY(cudaMalloc(&s->d_tab[i], (size_t)B * STED_TABLES * K * s->KP * 4));
        S_TRY(cudaMalloc(&s->d_scal[i], 3ull * B * 8));
        S_TRY(cudaMallocHost(&s->h_scal[i], 3ull * B * 8));
        S_TRY(cudaMalloc(&s->d_sig[i], s->n_roi * 4));
        S_TRY(cudaMalloc(&s->d_sig16[i], s->n_roi * 2));
        S_TRY(cudaMalloc(&s->d_int[i], s->n_roi * 4));
    }
    S_TRY(cudaMemcpyAsync(s->d_cache, psf_cache_host, 3ull * K * K * 8, cudaMemcpyHostToDevice, s->s_comp));
    S_TRY(cudaStreamSynchronize(s->s_comp));
#undef S_TRY
    *out = s;
    return STED_OK;
}

int sted_session_destroy(StedSession* s) {
    session_free(s);
    return STED_OK;
}

int sted_session_set_fluorophores(StedSession* s, const StedFluoParams* fluo_host) {
    int rc = session_bind(s);
    if (rc) return rc;
    if (!fluo_host) return fail(STED_EINVAL, "null pointer argument");
    // synchronous: called once per episode, and the caller's array need not outlive the call
    CUDA_TRY(cudaStreamSynchronize(s->s_prep));
    CUDA_TRY(cudaStreamSynchronize(s->s_comp));
    CUDA_TRY(cudaMemcpy(s->d_fluo, fluo_host, (size_t)s->B * sizeof(StedFluoParams), cudaMemcpyHostToDevice));
    s->fluo_set = true;
    s->keep_valid = false;                   // tables depend on the fluorophores
    return STED_OK;
}

int sted_session_upload(StedSession* s, const void* dmap_host, uint32_t flags, uint64_t max_molecules) {
    int rc = session_bind(s);
    if (rc) return rc;
    if (!dmap_host) return fail(STED_EINVAL, "null pointer argument");
    const bool u16 = (flags & STED_HOST_U16) != 0;
    const size_t esz = u16 ? 2 : 4;
    // the staging buffer is free once the previous pad / unpad + copy that used it has run
    CUDA_TRY(cudaStreamWaitEvent(s->s_in, s->e_pad, 0));
    CUDA_TRY(cudaStreamWaitEvent(s->s_in, s->e_mapout, 0));
    CUDA_TRY(cudaMemcpyAsync(s->d_stage, dmap_host, s->n_roi * esz, cudaMemcpyHostToDevice, s->s_in));
    CUDA_TRY(cudaEventRecord(s->e_up, s->s_in));
    CUDA_TRY(cudaStreamWaitEvent(s->s_comp, s->e_up, 0));
    CUDA_TRY(cudaStreamWaitEvent(s->s_comp, s->e_lent, 0));           // a session that adopted the previous map has copied it
    const int p = s->K / 2;
    if (u16) pad_kernel<<<296, 256, 0, s->s_comp>>>((const uint16_t*)s->d_stage, s->d_map[s->cur], s->B, s->H, s->W, p, s->Hp, s->Wpitch);
    else pad_kernel<<<296, 256, 0, s->s_comp>>>((const float*)s->d_stage, s->d_map[s->cur], s->B, s->H, s->W, p, s->Hp, s->Wpitch);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(s->e_pad, s->s_comp));
    CUDA_TRY(cudaEventRecord(s->e_map, s->s_comp));
    s->map_set = true;
    if (max_molecules == 0) max_molecules = 4ull * s->n_roi;           // too small: sted_session_sync reports STED_EOVERFLOW
    return session_reserve(s, max_molecules + 16);
}

int sted_session_upload_device(StedSession* s, const void* roi_dev, uint32_t flags, uint64_t max_molecules, void* stream) {
    int rc = session_bind(s);
    if (rc) return rc;
    if (!roi_dev) return fail(STED_EINVAL, "null pointer argument");
    const bool u16 = (flags & STED_HOST_U16) != 0;
    // roi_dev is ready once what `stream` holds now has run
    CUDA_TRY(cudaEventRecord(s->e_up, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamWaitEvent(s->s_comp, s->e_up, 0));
    CUDA_TRY(cudaStreamWaitEvent(s->s_comp, s->e_lent, 0));
    const int p = s->K / 2;
    if (u16) pad_kernel<<<296, 256, 0, s->s_comp>>>((const uint16_t*)roi_dev, s->d_map[s->cur], s->B, s->H, s->W, p, s->Hp, s->Wpitch);
    else pad_kernel<<<296, 256, 0, s->s_comp>>>((const float*)roi_dev, s->d_map[s->cur], s->B, s->H, s->W, p, s->Hp, s->Wpitch);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(s->e_map, s->s_comp));
    s->map_set = true;
    if (max_molecules == 0) max_molecules = 4ull * s->n_roi;
    return session_reserve(s, max_molecules + 16);
}

int sted_session_result(StedSession* s, int back, float** signal_dev, float** intensity_dev, void* stream, int wait) {
    int rc = session_bind(s);
    if (rc) return rc;
    if (back < 0 || back >= StedSession::NSIG) return fail(STED_EINVAL, "back must lie in [0, %d)", StedSession::NSIG);
    const int i = (s->next_sig + 2 * StedSession::NSIG - 1 - back) % StedSession::NSIG;
    if (signal_dev) *signal_dev = s->d_sig[i];
    if (intensity_dev) *intensity_dev = s->d_int[i];
    if (wait) CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, s->e_det[i], 0));
    return STED_OK;
}

int sted_session_kernel_ms(StedSession* s, int back, float* ms) {
    int rc = session_bind(s);
    if (rc) return rc;
    if (!ms) return fail(STED_EINVAL, "null pointer argument");
    if (back < 0 || back >= StedSession::NT || back >= s->n_acq)
        return fail(STED_EINVAL, "back must lie in [0, min(%d, acquisitions queued so far))", StedSession::NT);
    const int it = (int)((s->n_acq - 1 - back) % StedSession::NT);
    CUDA_TRY(cudaEventSynchronize(s->e_t1[it]));
    CUDA_TRY(cudaEventElapsedTime(ms, s->e_t0[it], s->e_t1[it]));
    return STED_OK;
}

int sted_session_join(StedSession* s, void* stream) {
    int rc = session_bind(s);
    if (rc) return rc;
    const cudaStream_t all[5] = {s->s_in, s->s_prep, s->s_comp, s->s_post, s->s_out};
    for (int k = 0; k < 5; ++k) {
        CUDA_TRY(cudaEventRecord(s->e_join[k], all[k]));
        CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, s->e_join[k], 0));
    }
    return STED_OK;
}

int sted_session_adopt(StedSession* dst, StedSession* src) {
    int rc = session_bind(dst);
    if (rc) return rc;
    if ((rc = session_bind(src))) return rc;
    if (dst == src) return fail(STED_EINVAL, "a session cannot adopt its own map");
    if (dst->B != src->B || dst->H != src->H || dst->W != src->W || dst->K != src->K)
        return fail(STED_EINVAL, "sessions of different shapes");
    if (!src->map_set) return fail(STED_EINVAL, "the source session has no map");
    if ((rc = session_reserve(dst, src->max_events))) return rc;
    // after everything queued on both sessions' compute streams; the source's next upload waits for the copy
    CUDA_TRY(cudaEventRecord(src->e_comp, src->s_comp));
    CUDA_TRY(cudaStreamWaitEvent(dst->s_comp, src->e_comp, 0));
    CUDA_TRY(cudaMemcpyAsync(dst->d_map[dst->cur], src->d_map[src->cur], dst->n_pad * 4, cudaMemcpyDeviceToDevice, dst->s_comp));
    CUDA_TRY(cudaEventRecord(src->e_lent, dst->s_comp));
    CUDA_TRY(cudaEventRecord(dst->e_map, dst->s_comp));
    dst->map_set = true;
    return STED_OK;
}

int sted_session_after(StedSession* s, StedSession* other) {
    int rc = session_bind(s);
    if (rc) return rc;
    if ((rc = session_bind(other))) return rc;
    if (s == other) return fail(STED_EINVAL, "a session cannot wait for itself");
    CUDA_TRY(cudaEventRecord(other->e_comp, other->s_comp));
    CUDA_TRY(cudaStreamWaitEvent(s->s_comp, other->e_comp, 0));
    return STED_OK;
}

int sted_session_acquire(StedSession* s, const double* p_ex_host, const double* p_sted_host, const double* pdt_host,
                         uint32_t flags, const StedDetectorParams* det, double lambda_fluo, uint64_t seed, uint64_t offset,
                         int64_t env_base, void* signal_host, float* intensity_host) {
    int rc = session_bind(s);
    if (rc) return rc;
    if (!p_ex_host || !p_sted_host || !pdt_host || !det) return fail(STED_EINVAL, "null pointer argument");
    if (!s->fluo_set || !s->map_set) return fail(STED_EINVAL, "session needs sted_session_set_fluorophores and sted_session_upload first");
    const bool bleach = (flags & STED_BLEACH) != 0;
    const bool stoch = bleach && (flags & STED_SAMPLE_BLEACH);
    const bool u16 = (flags & STED_HOST_U16) != 0;
    if (u16 && ((bleach && !stoch) || !(flags & STED_SAMPLE_PHOTONS)))
        return fail(STED_EINVAL, "STED_HOST_U16 needs integer outputs: STED_SAMPLE_PHOTONS, and STED_SAMPLE_BLEACH when bleaching");
    flags &= ~STED_HOST_U16;
    const int i = s->next_sig;
    s->next_sig = (i + 1) % StedSession::NSIG;
    const int B = s->B;
    // result slot i is reusable once its previous download has finished (also covers h_scal[i] / d_scal[i])
    CUDA_TRY(cudaEventSynchronize(s->e_out[i]));
    memcpy(s->h_scal[i], p_ex_host, (size_t)B * 8);
    memcpy(s->h_scal[i] + B, p_sted_host, (size_t)B * 8);
    memcpy(s->h_scal[i] + 2 * B, pdt_host, (size_t)B * 8);
    // ---- s_prep: parameters, tables and (stochastic scan) the death schedule.  Neither depends on the acquisition queued
    // before this one, so they run under it; the schedule needs the final map and a free event scratch.
    const double* d_pdt = s->d_scal[i] + 2 * B;
    const float* d_tab = s->d_tab[i];
    if (!bleach) {
        // confocal-type acquisition: the kept tables, rebuilt only when the parameters differ from the kept ones
        if (!s->keep_valid || memcmp(s->h_scal_keep, s->h_scal[i], 3ull * B * 8) != 0) {
            memcpy(s->h_scal_keep, s->h_scal[i], 3ull * B * 8);
            CUDA_TRY(cudaStreamWaitEvent(s->s_prep, s->e_keep_used, 0));      // acquisitions still reading the old ones
            CUDA_TRY(cudaMemcpyAsync(s->d_scal_keep, s->h_scal[i], 3ull * B * 8, cudaMemcpyHostToDevice, s->s_prep));
            if ((rc = sted_make_effective(s->d_cache, s->d_fluo, s->d_scal_keep, s->d_scal_keep + B, s->d_scal_keep + 2 * B, B, s->K,
                                          s->d_tab_keep, nullptr, s->s_prep)))
                return rc;
            CUDA_TRY(cudaEventRecord(s->e_keep, s->s_prep));
            s->keep_valid = true;
        }
        CUDA_TRY(cudaStreamWaitEvent(s->s_prep, s->e_keep, 0));
        d_pdt = s->d_scal_keep + 2 * B;
        d_tab = s->d_tab_keep;
    } else {
        CUDA_TRY(cudaMemcpyAsync(s->d_scal[i], s->h_scal[i], 3ull * B * 8, cudaMemcpyHostToDevice, s->s_prep));
        if ((rc = sted_make_effective(s->d_cache, s->d_fluo, s->d_scal[i], s->d_scal[i] + B, d_pdt, B, s->K, s->d_tab[i], nullptr,
                                      s->s_prep)))
            return rc;
    }
    float* d_in = s->d_map[s->cur];
    float* d_out = s->d_map[s->cur ^ 1];
    if (stoch) {
        CUDA_TRY(cudaStreamWaitEvent(s->s_prep, s->e_map, 0));
        CUDA_TRY(cudaStreamWaitEvent(s->s_prep, s->e_wsfree, 0));
        if ((rc = sted_bleach_presample(d_in, d_tab, d_pdt, B, s->H, s->W, s->K, flags, seed, offset, env_base, s->d_ws,
                                        s->ws_bytes, s->s_prep)))
            return rc;
    }
    CUDA_TRY(cudaEventRecord(s->e_prep[i], s->s_prep));
    // ---- s_comp: the gather or the scan
    CUDA_TRY(cudaStreamWaitEvent(s->s_comp, s->e_prep[i], 0));
    const int it = (int)(s->n_acq++ % StedSession::NT);
    CUDA_TRY(cudaEventRecord(s->e_t0[it], s->s_comp));
    // A confocal acquisition with a host destination runs as two gathers over halves of the batch: the detector and the
    // download of the first half then run under the gather of the second, and when the last kernel of a step has
    // finished only the tail of half an image is left to copy (a caller that synchronises every step waits for exactly
    // that).  Two launches of >= 2048 CTAs cost about 1 % in partial waves; the scan is not cut (3 CTAs per SM: half a
    // batch would leave its last wave half empty).  Results do not depend on the cut: Philox counters carry the global
    // environment id.
    const size_t hw = (size_t)s->H * s->W;
    const int nhalf = (!bleach && signal_host && B >= 64) ? 2 : 1;
    const int nrun = (signal_host && B >= 8) ? 4 / nhalf : 1;         // detector / conversion / copy runs per half
    for (int h = 0; h < nhalf; ++h) {
        const int hb0 = (int)((long long)B * h / nhalf), hnb = (int)((long long)B * (h + 1) / nhalf) - hb0;
        if ((rc = sted_acquire(d_in + (size_t)hb0 * s->Hp * s->Wpitch, bleach ? d_out : nullptr,
                               d_tab + (size_t)hb0 * STED_TABLES * s->K * s->KP, d_pdt + hb0, hnb, s->H, s->W, s->K,
                               flags | STED_NO_DETECT | (stoch ? STED_PRESAMPLED : 0u), det, lambda_fluo, seed, offset, env_base + hb0,
                               intensity_host ? s->d_int[i] + hb0 * hw : nullptr, s->d_sig[i] + hb0 * hw, stoch ? s->d_ws : nullptr,
                               stoch ? s->ws_bytes : 0, s->s_comp)))
            return rc;
        if (h == nhalf - 1) {
            CUDA_TRY(cudaEventRecord(s->e_t1[it], s->s_comp));
            s->h_drop[i] = 0;
            if (stoch) {
                CUDA_TRY(cudaMemcpyAsync(&s->h_drop[i], s->d_ws, sizeof(int), cudaMemcpyDeviceToHost, s->s_comp));
                CUDA_TRY(cudaEventRecord(s->e_wsfree, s->s_comp));
            }
            if (bleach) {
                s->cur ^= 1;
                CUDA_TRY(cudaEventRecord(s->e_map, s->s_comp));
            }
        }
        CUDA_TRY(cudaEventRecord(s->e_main[i], s->s_comp));
        // ---- s_post: detector (+ uint16 conversion) in runs of environments, under whatever s_comp runs next; the copy of
        // one run overlaps the detector of the next
        CUDA_TRY(cudaStreamWaitEvent(s->s_post, s->e_main[i], 0));
        for (int c = 0; c < nrun; ++c) {
            const int b0 = hb0 + (int)((long long)hnb * c / nrun), nb = hb0 + (int)((long long)hnb * (c + 1) / nrun) - b0;
            if ((rc = sted_detect(s->d_sig[i] + b0 * hw, d_pdt + b0, nb, s->H, s->W, flags, det, lambda_fluo, seed, offset, env_base + b0,
                                  s->s_post)))
                return rc;
            if (h == nhalf - 1 && c == nrun - 1 && !bleach) CUDA_TRY(cudaEventRecord(s->e_keep_used, s->s_post));
            if (u16 && signal_host) {
                counts_to_u16_kernel<<<296, 256, 0, s->s_post>>>(s->d_sig[i] + b0 * hw, s->d_sig16[i] + b0 * hw, nb * hw);
                CUDA_TRY(cudaGetLastError());
            }
            CUDA_TRY(cudaEventRecord(s->e_det[i], s->s_post));
            CUDA_TRY(cudaStreamWaitEvent(s->s_out, s->e_det[i], 0));      // e_out[i] below then covers the slot's kernels too
            if (signal_host) {
                if (u16) CUDA_TRY(cudaMemcpyAsync((uint16_t*)signal_host + b0 * hw, s->d_sig16[i] + b0 * hw, nb * hw * 2, cudaMemcpyDeviceToHost, s->s_out));
                else CUDA_TRY(cudaMemcpyAsync((float*)signal_host + b0 * hw, s->d_sig[i] + b0 * hw, nb * hw * 4, cudaMemcpyDeviceToHost, s->s_out));
            }
        }
    }
    if (intensity_host) CUDA_TRY(cudaMemcpyAsync(intensity_host, s->d_int[i], s->n_roi * 4, cudaMemcpyDeviceToHost, s->s_out));
    CUDA_TRY(cudaEventRecord(s->e_out[i], s->s_out));
    // the next kernels on s_comp must not overwrite slot i ... they use slot i+1; slot i is guarded by e_out[i] above
    return STED_OK;
}

int sted_session_download(StedSession* s, void* dmap_host, uint32_t flags) {
    int rc = session_bind(s);
    if (rc) return rc;
    if (!dmap_host) return fail(STED_EINVAL, "null pointer argument");
    if (!s->map_set) return fail(STED_EINVAL, "no map uploaded");
    const bool u16 = (flags & STED_HOST_U16) != 0;
    const int p = s->K / 2;
    // the staging buffer may still feed a pad kernel or an earlier download
    CUDA_TRY(cudaStreamWaitEvent(s->s_comp, s->e_mapout, 0));
    CUDA_TRY(cudaStreamWaitEvent(s->s_comp, s->e_up, 0));
    if (u16) unpad_kernel<<<296, 256, 0, s->s_comp>>>(s->d_map[s->cur], (uint16_t*)s->d_stage, s->B, s->H, s->W, p, s->Hp, s->Wpitch);
    else unpad_kernel<<<296, 256, 0, s->s_comp>>>(s->d_map[s->cur], (float*)s->d_stage, s->B, s->H, s->W, p, s->Hp, s->Wpitch);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(s->e_unpad, s->s_comp));
    CUDA_TRY(cudaEventRecord(s->e_pad, s->s_comp));                     // later uploads wait for the unpad as well
    CUDA_TRY(cudaStreamWaitEvent(s->s_out, s->e_unpad, 0));
    CUDA_TRY(cudaMemcpyAsync(dmap_host, s->d_stage, s->n_roi * (u16 ? 2 : 4), cudaMemcpyDeviceToHost, s->s_out));
    CUDA_TRY(cudaEventRecord(s->e_mapout, s->s_out));
    return STED_OK;
}

int sted_session_sync(StedSession* s, int64_t* dropped_events) {
    int rc = session_bind(s);
    if (rc) return rc;
    for (cudaStream_t st : {s->s_in, s->s_prep, s->s_comp, s->s_post, s->s_out}) CUDA_TRY(cudaStreamSynchronize(st));
    long long dropped = 0;
    for (int i = 0; i < StedSession::NSIG; ++i) { dropped += s->h_drop[i]; s->h_drop[i] = 0; }
    if (dropped_events) *dropped_events = dropped;
    if (dropped > 0)
        return fail(STED_EOVERFLOW, "%lld bleaching events did not fit the session's event scratch: the results since the last "
                                    "sync are invalid; upload again with a larger max_molecules", dropped);
    return STED_OK;
}

int sted_session_device_map(StedSession* s, float** map_dev) {
    int rc = session_bind(s);
    if (rc) return rc;
    if (!map_dev) return fail(STED_EINVAL, "null pointer argument");
    *map_dev = s->d_map[s->cur];
    return STED_OK;
}

}  // extern "C"
