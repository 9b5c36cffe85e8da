// rewards_kernels.cu — per-step foreground masks and objectives on the device.
//
//   gym_sted/utils.py:16-24            get_foreground   = skimage.filters.threshold_otsu on an integer image
//   gym_sted/envs/mosted_env.py:237-244 fg_c, fg_s (all ones when the STED image is empty), fg_s *= fg_c
//   gym_sted/rewards/objectives.py:91-100   Signal_Ratio(75)
//   gym_sted/rewards/objectives.py:107-111  Bleach
//
// Photon counts are integers, so everything the reference computes in float64 from them can be made
// bit-identical: histograms, cumulative sums and masked sums are exact integers (any summation order),
// and the few float64 operations (means, between-class variance, percentile interpolation) are done in
// the same order as numpy does them.  One CTA per environment.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace {

constexpr int R_THREADS = 256;
constexpr int R_SMEM_BINS = 8192;       // per histogram held in shared memory
constexpr int R_MAX_BINS = 65536;       // with the global scratch

struct BlockStats {
    long long a, b;
};

__device__ __forceinline__ long long block_sum(long long v, long long* scratch) {
    // scratch: [R_THREADS/32 + 1]
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    long long t = 0;
    for (int i = 0; i < R_THREADS / 32; ++i) t += scratch[i];
    return t;
}

__device__ __forceinline__ int block_min(int v, int* scratch) {
    for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = scratch[0];
    for (int i = 1; i < R_THREADS / 32; ++i) t = min(t, scratch[i]);
    return t;
}

__device__ __forceinline__ int block_max(int v, int* scratch) { return -block_min(-v, scratch); }

// Otsu threshold of the histogram hist[0..nbins) whose bin i sits at value lo+i (hist[0], hist[nbins-1] > 0).
// Warp 0 does it: each lane owns a contiguous run of bins.  Returns the threshold value (all threads).
__device__ int otsu_threshold(const int* hist, int nbins, int lo, long long total_n, long long total_s, int* s_out) {
    if (nbins <= 1) return lo;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const int per = (nbins + 31) / 32;
        const int b0 = min(lane * per, nbins), b1 = min(b0 + per, nbins);
        long long cn = 0, cs = 0;
        for (int i = b0; i < b1; ++i) { cn += hist[i]; cs += (long long)hist[i] * (lo + i); }
        long long pn = cn, ps = cs;                       // inclusive scans over lanes
        for (int o = 1; o < 32; o <<= 1) {
            const long long tn = __shfl_up_sync(0xffffffffu, pn, o), ts = __shfl_up_sync(0xffffffffu, ps, o);
            if (lane >= o) { pn += tn; ps += ts; }
        }
        long long w1 = pn - cn, s1 = ps - cs;             // exclusive prefix: bins before b0
        double best = -1.0;
        int best_i = 0x7fffffff;
        for (int i = b0; i < b1 && i < nbins - 1; ++i) {
            w1 += hist[i];
            s1 += (long long)hist[i] * (lo + i);
            const long long w2 = total_n - w1, s2 = total_s - s1;        // bins i+1 ..
            const double mean1 = (double)s1 / (double)w1;
            const double mean2 = (double)s2 / (double)w2;
            const double diff = mean1 - mean2;
            const double var = ((double)w1 * (double)w2) * (diff * diff);
            if (var > best) { best = var; best_i = i; }     // first maximum inside the run
        }
        for (int o = 16; o; o >>= 1) {                       // first maximum over the lanes
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
        }
        if (lane == 0) *s_out = lo + best_i;
    }
    __syncthreads();
    return *s_out;
}

// Value at sorted rank k (0-based) of the multiset described by hist (bin i = value lo+i).  Thread 0 walks.
__device__ void ranks_from_hist(const int* hist, int nbins, int lo, long long k0, long long k1, int* out2) {
    if (threadIdx.x == 0) {
        long long c = 0;
        int v0 = lo, v1 = lo;
        bool f0 = false;
        for (int i = 0; i < nbins; ++i) {
            c += hist[i];
            if (!f0 && c > k0) { v0 = lo + i; f0 = true; }
            if (c > k1) { v1 = lo + i; break; }
        }
        out2[0] = v0; out2[1] = v1;
    }
    __syncthreads();
}

// numpy.percentile(x, 75) with the default 'linear' method, from the two neighbouring order statistics
__device__ __forceinline__ double percentile75(long long n, const int* hist, int nbins, int lo, int* tmp2) {
    const double vi = (double)n * 0.75 + (1.0 + 0.75 * (1.0 - 1.0 - 1.0)) - 1.0;     // _compute_virtual_index(alpha=beta=1)
    const double prev = floor(vi);
    const double gamma = vi - prev;
    long long k0 = (long long)prev, k1 = k0 + 1;
    if (k1 > n - 1) k1 = n - 1;
    if (k0 > n - 1) k0 = n - 1;
    ranks_from_hist(hist, nbins, lo, k0, k1, tmp2);
    const double a = (double)tmp2[0], b = (double)tmp2[1];
    const double diff = b - a;
    double r = a + diff * gamma;                          // numpy _lerp
    if (gamma >= 0.5) r = b - diff * (1.0 - gamma);
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(R_THREADS) objectives_kernel(const float* __restrict__ conf1, const float* __restrict__ sted,
                                                             const float* __restrict__ conf2, int HW,
                                                             uint8_t* __restrict__ fg_c, uint8_t* __restrict__ fg_s,
                                                             double* __restrict__ objs, int* __restrict__ flags,
                                                             int* __restrict__ hist_scratch) {
    extern __shared__ int sh[];
    __shared__ long long red[R_THREADS / 32 + 1];
    __shared__ int redi[R_THREADS / 32 + 1];
    __shared__ int s_tmp[4];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* c1 = conf1 + (size_t)b * HW;
    const float* st = sted + (size_t)b * HW;
    const float* c2 = conf2 + (size_t)b * HW;
    uint8_t* mc = fg_c + (size_t)b * HW;
    uint8_t* ms = fg_s + (size_t)b * HW;
    int flag = 0;

    // ---- ranges
    int c_lo = 0x7fffffff, c_hi = -0x7fffffff, s_lo = 0x7fffffff, s_hi = -0x7fffffff;
    for (int i = tid; i < HW; i += R_THREADS) {
        const int a = __float2int_rn(c1[i]), s = __float2int_rn(st[i]);
        c_lo = min(c_lo, a); c_hi = max(c_hi, a); s_lo = min(s_lo, s); s_hi = max(s_hi, s);
    }
    c_lo = block_min(c_lo, redi); c_hi = block_max(c_hi, redi);
    s_lo = block_min(s_lo, redi); s_hi = block_max(s_hi, redi);
    const int nb_c = c_hi - c_lo + 1, nb_s = s_hi - s_lo + 1;
    int* hist_c = sh;
    int* hist_s = sh + R_SMEM_BINS;
    if (nb_c > R_SMEM_BINS || nb_s > R_SMEM_BINS) {
        if (hist_scratch && nb_c <= R_MAX_BINS && nb_s <= R_MAX_BINS) {
            hist_c = hist_scratch + (size_t)b * 2 * R_MAX_BINS;
            hist_s = hist_c + R_MAX_BINS;
        } else {
            if (tid == 0) { flags[b] = 2; objs[2 * b] = objs[2 * b + 1] = __longlong_as_double(0x7ff8000000000000LL); }
            return;                                       // value range beyond the histogram capacity
        }
    }
    auto clear = [&](int* h, int n) { for (int i = tid; i < n; i += R_THREADS) h[i] = 0; };
    clear(hist_c, nb_c); clear(hist_s, nb_s);
    __syncthreads();
    long long sum_c = 0, sum_s = 0;
    for (int i = tid; i < HW; i += R_THREADS) {
        const int a = __float2int_rn(c1[i]), s = __float2int_rn(st[i]);
        atomicAdd(&hist_c[a - c_lo], 1); atomicAdd(&hist_s[s - s_lo], 1);
        sum_c += a; sum_s += s;
    }
    sum_c = block_sum(sum_c, red);
    sum_s = block_sum(sum_s, red);
    __syncthreads();

    // ---- Otsu thresholds and masks
    const int thr_c = otsu_threshold(hist_c, nb_c, c_lo, HW, sum_c, &s_tmp[0]);
    const bool sted_any = (s_hi != 0 || s_lo != 0);       // numpy.any(sted_image)
    const int thr_s = sted_any ? otsu_threshold(hist_s, nb_s, s_lo, HW, sum_s, &s_tmp[1]) : 0;
    __syncthreads();
    clear(hist_c, nb_c); clear(hist_s, nb_s);
    __syncthreads();
    // masked histograms / sums: conf1 over fg_c, sted over fg_s; sums of conf1, conf2 over fg_c, sted outside fg_s
    long long n_c = 0, n_s = 0, sc1 = 0, sc2 = 0, sbg = 0;
    for (int i = tid; i < HW; i += R_THREADS) {
        const int a = __float2int_rn(c1[i]), s = __float2int_rn(st[i]);
        const bool fc = a > thr_c;
        const bool fs = (sted_any ? (s > thr_s) : true) && fc;
        mc[i] = fc; ms[i] = fs;
        if (fc) { ++n_c; sc1 += a; sc2 += __float2int_rn(c2[i]); atomicAdd(&hist_c[a - c_lo], 1); }
        if (fs) { ++n_s; atomicAdd(&hist_s[s - s_lo], 1); } else sbg += s;
    }
    n_c = block_sum(n_c, red); n_s = block_sum(n_s, red);
    sc1 = block_sum(sc1, red); sc2 = block_sum(sc2, red); sbg = block_sum(sbg, red);
    __syncthreads();

    // ---- Bleach = (mean(conf1[fg_c]) - mean(conf2[fg_c])) / mean(conf1[fg_c])
    const double m1 = (double)sc1 / (double)n_c, m2 = (double)sc2 / (double)n_c;
    const double bleach = (m1 - m2) / m1;
    // ---- SNR = (P75(sted[fg_s]) - mean(sted[~fg_s])) / P75(conf1[fg_c]); 0 without STED foreground; None if < 0
    double snr = 0.0;
    if (n_s > 0) {
        const double fg = percentile75(n_s, hist_s, nb_s, s_lo, &s_tmp[2]);
        const double bg = (double)sbg / (double)((long long)HW - n_s);
        const double den = percentile75(n_c, hist_c, nb_c, c_lo, &s_tmp[2]);
        snr = (fg - bg) / den;
        if (snr < 0.0) flag |= 1;                         // the reference returns None here
    }
    if (tid == 0) { objs[2 * b] = bleach; objs[2 * b + 1] = snr; flags[b] = flag; }
}

}  // namespace

extern "C" int sted_foreground_objectives(const float* conf1_dev, const float* sted_dev, const float* conf2_dev, int B,
                                          int H, int W, uint8_t* fg_c_dev, uint8_t* fg_s_dev, double* objs_dev,
                                          int32_t* flags_dev, int32_t* hist_scratch_dev, void* stream) {
    if (!conf1_dev || !sted_dev || !conf2_dev || !fg_c_dev || !fg_s_dev || !objs_dev || !flags_dev)
        return fail(STED_EINVAL, "null pointer argument");
    if (B <= 0 || H <= 0 || W <= 0) return fail(STED_EINVAL, "B, H, W must be positive");
    int rc = check_device();
    if (rc) return rc;
    const size_t smem = 2 * R_SMEM_BINS * sizeof(int);
    CUDA_TRY(cudaFuncSetAttribute(objectives_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    objectives_kernel<<<B, R_THREADS, smem, (cudaStream_t)stream>>>(conf1_dev, sted_dev, conf2_dev, H * W, fg_c_dev, fg_s_dev,
                                                                    objs_dev, flags_dev, hist_scratch_dev);
    CUDA_TRY(cudaGetLastError());
    return STED_OK;
}
