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
#include "lm_minpack_warp.cuh"

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
                                                             uint8_t* fg_c, uint8_t* fg_s,
                                                             double* __restrict__ objs, int* __restrict__ flags,
                                                             int* __restrict__ hist_scratch, const int masks_given) {
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

    // ---- Otsu thresholds and masks (masks_given: the caller's masks are read instead, objectives.py's evaluate()
    // takes them as arguments)
    const int thr_c = masks_given ? 0 : otsu_threshold(hist_c, nb_c, c_lo, HW, sum_c, &s_tmp[0]);
    const bool sted_any = (s_hi != 0 || s_lo != 0);       // numpy.any(sted_image)
    const int thr_s = (sted_any && !masks_given) ? otsu_threshold(hist_s, nb_s, s_lo, HW, sum_s, &s_tmp[1]) : 0;
    __syncthreads();
    clear(hist_c, nb_c); clear(hist_s, nb_s);
    __syncthreads();
    // masked histograms / sums: conf1 over fg_c, sted over fg_s; sums of conf1, conf2 over fg_c, sted outside fg_s
    long long n_c = 0, n_s = 0, sc1 = 0, sc2 = 0, sbg = 0;
    for (int i = tid; i < HW; i += R_THREADS) {
        const int a = __float2int_rn(c1[i]), s = __float2int_rn(st[i]);
        const bool fc = masks_given ? mc[i] != 0 : a > thr_c;
        const bool fs = masks_given ? ms[i] != 0 : (sted_any ? (s > thr_s) : true) && fc;
        if (!masks_given) { mc[i] = fc; ms[i] = fs; }
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
        // numpy.percentile of an empty selection is NaN (only possible with caller-supplied masks; n_c is block-uniform)
        const double den = n_c > 0 ? percentile75(n_c, hist_c, nb_c, c_lo, &s_tmp[2]) : __longlong_as_double(0x7ff8000000000000LL);
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
                                                                    objs_dev, flags_dev, hist_scratch_dev, 0);
    CUDA_TRY(cudaGetLastError());
    return STED_OK;
}

extern "C" int sted_objectives_masked(const float* conf1_dev, const float* sted_dev, const float* conf2_dev, int B, int H, int W,
                                      const uint8_t* fg_c_dev, const uint8_t* fg_s_dev, double* objs_dev, int32_t* flags_dev,
                                      int32_t* hist_scratch_dev, void* stream) {
    if (!conf1_dev || !sted_dev || !conf2_dev || !fg_c_dev || !fg_s_dev || !objs_dev || !flags_dev)
        return fail(STED_EINVAL, "null pointer argument");
    if (B <= 0 || H <= 0 || W <= 0) return fail(STED_EINVAL, "B, H, W must be positive");
    int rc = check_device();
    if (rc) return rc;
    const size_t smem = 2 * R_SMEM_BINS * sizeof(int);
    CUDA_TRY(cudaFuncSetAttribute(objectives_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    objectives_kernel<<<B, R_THREADS, smem, (cudaStream_t)stream>>>(conf1_dev, sted_dev, conf2_dev, H * W,
                                                                    const_cast<uint8_t*>(fg_c_dev), const_cast<uint8_t*>(fg_s_dev),
                                                                    objs_dev, flags_dev, hist_scratch_dev, 1);
    CUDA_TRY(cudaGetLastError());
    return STED_OK;
}

// =============================================================================================================
// Nanodomain candidate detection (the integer part of NumberNanodomains.evaluate,
// gym_sted/rewards/objectives.py:416-422):
//     filtered = skimage.filters.gaussian(sted, sigma=0.5)      (scipy.ndimage.gaussian_filter, mode nearest, truncate 4)
//     mask     = filtered > numpy.quantile(filtered, 0.99)
//     mask     = remove_small_objects(mask, min_size=4, connectivity=2);  labels = measure.label(mask)
//     peaks    = peak_local_max(filtered, min_distance=2, labels=labels, exclude_border=False)
// Everything is float64 in the reference; the float operations are replayed in the same order with explicitly
// rounded multiplies/adds (no FMA contraction), so the filtered image, the quantile and therefore the integer
// outputs (mask, labels, peak coordinates) are bit-identical.  One CTA per image; the mask holds at most 1 % of
// the pixels, so labelling and peak picking run on a short list.
// =============================================================================================================
namespace {

constexpr int D_THREADS = 512;         // 16 warps: the passes over the image are latency bound
constexpr int D_MAXFG = 4096;        // foreground pixels per image (1 % of up to 640 x 640)
constexpr int D_MAXPEAKS = 256;
constexpr int D_SMEM_BYTES = D_MAXFG * (3 * (int)sizeof(int) + 2 * (int)sizeof(short));   // 64 KB

__device__ __forceinline__ double gauss5(const double w0, const double w1, const double w2, const double xm2, const double xm1,
                                         const double x0, const double xp1, const double xp2) {
    // NI_Correlate1D, symmetric branch: tmp = x0*fw[0]; tmp += fw[-2]*(x[-2]+x[2]); tmp += fw[-1]*(x[-1]+x[1])
    double t = __dmul_rn(x0, w0);
    t = __dadd_rn(t, __dmul_rn(w2, __dadd_rn(xm2, xp2)));
    t = __dadd_rn(t, __dmul_rn(w1, __dadd_rn(xm1, xp1)));
    return t;
}

__global__ void __launch_bounds__(D_THREADS) nanodomain_kernel(const float* __restrict__ sted, int H, int W, double w0, double w1,
                                                             double w2, long long k0, double gamma, double* __restrict__ tmp,
                                                             double* __restrict__ filt, int* __restrict__ peaks,
                                                             int* __restrict__ npeaks, int* __restrict__ flags) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long s_prefix, s_v0, s_v1;
    __shared__ long long s_rank;
    __shared__ int s_count, s_warp_off[D_THREADS / 32 + 1];
    __shared__ double s_min[D_THREADS / 32];
    // short-list state of the labelling / peak picking (dynamic shared memory, D_SMEM_BYTES)
    extern __shared__ __align__(16) unsigned char d_smem[];
    int* fg_idx = reinterpret_cast<int*>(d_smem);                 // [D_MAXFG] raster-ordered pixel indices of the mask
    int* csize = fg_idx + D_MAXFG;                                // [D_MAXFG] per root: size -> kept peaks -> output offset
    int* clast = csize + D_MAXFG;                                 // [D_MAXFG] per root: list position of its last pixel
    short* lab = reinterpret_cast<short*>(clast + D_MAXFG);       // [D_MAXFG] component label (= list position of its first pixel)
    short* korder = lab + D_MAXFG;                                // [D_MAXFG] rank of a kept peak inside its component, else -1
    __shared__ int s_nfg, s_flag, s_total;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = H * W;
    const float* img = sted + (size_t)b * N;
    double* t1 = tmp + (size_t)b * N;
    double* f = filt + (size_t)b * N;

    // ---- gaussian along axis 0, then along axis 1 (scipy's order), edges replicated
    for (int i = tid; i < N; i += D_THREADS) {
        const int y = i / W, x = i - y * W;
        // skimage.filters.gaussian first maps the int64 image through img_as_float: x -> ((x + 0.5) * 2) / 2^64; the
        // power of two commutes with every rounding, so filtering 2x + 1 gives the same bits up to that factor
        auto at = [&](int yy) { return 2.0 * (double)img[min(max(yy, 0), H - 1) * W + x] + 1.0; };
        t1[i] = gauss5(w0, w1, w2, at(y - 2), at(y - 1), at(y), at(y + 1), at(y + 2));
    }
    __syncthreads();
    double vmin = 1.0e300;
    for (int i = tid; i < N; i += D_THREADS) {
        const int y = i / W, x = i - y * W;
        auto at = [&](int xx) { return t1[y * W + min(max(xx, 0), W - 1)]; };
        const double v = gauss5(w0, w1, w2, at(x - 2), at(x - 1), at(x), at(x + 1), at(x + 2));
        f[i] = v;
        vmin = fmin(vmin, v);
    }
    for (int o = 16; o; o >>= 1) vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    if (lane == 0) s_min[warp] = vmin;
    __syncthreads();
    vmin = s_min[0];
    for (int i = 1; i < D_THREADS / 32; ++i) vmin = fmin(vmin, s_min[i]);

    // ---- order statistic k0 by MSB-first radix select on the (non-negative) doubles' bit patterns
    if (tid == 0) { s_prefix = 0ull; s_rank = k0; }
    for (int shift = 56; shift >= 0; shift -= 8) {
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        const unsigned long long himask = (shift == 56) ? 0ull : (~0ull << (shift + 8));
        // lanes of a warp that hit the same bin add once (the high bytes of neighbouring pixels are nearly always equal:
        // 32-way conflicts on one shared-memory word otherwise)
        for (int base = 0; base < N; base += D_THREADS) {
            const int i = base + tid;
            unsigned int bin = 0xFFFFFFFFu;
            if (i < N) {
                const unsigned long long key = (unsigned long long)__double_as_longlong(f[i]);
                if ((key & himask) == prefix) bin = (unsigned int)(key >> shift) & 0xFFu;
            }
            const unsigned int peers = __match_any_sync(0xffffffffu, bin);
            if (bin != 0xFFFFFFFFu && lane == __ffs(peers) - 1) atomicAdd(&hist[bin], (unsigned int)__popc(peers));
        }
        __syncthreads();
        if (tid == 0) {
            long long r = s_rank;
            int bin = 0;
            for (; bin < 256; ++bin) { if (r < (long long)hist[bin]) break; r -= hist[bin]; }
            s_prefix = prefix | ((unsigned long long)bin << shift);
            s_rank = r;
        }
        __syncthreads();
    }
    // next order statistic: v0 again if enough copies of it, else the smallest value above it
    {
        const unsigned long long v0 = s_prefix;
        unsigned long long above = ~0ull;
        int le = 0;
        for (int i = tid; i < N; i += D_THREADS) {
            const unsigned long long key = (unsigned long long)__double_as_longlong(f[i]);
            if (key <= v0) ++le; else above = min(above, key);
        }
        for (int o = 16; o; o >>= 1) { le += __shfl_xor_sync(0xffffffffu, le, o); above = min(above, __shfl_xor_sync(0xffffffffu, above, o)); }
        if (tid == 0) { s_count = 0; s_v1 = ~0ull; }
        __syncthreads();
        if (lane == 0) { atomicAdd(&s_count, le); atomicMin(&s_v1, above); }
        __syncthreads();
        if (tid == 0) { s_v0 = v0; if ((long long)s_count > k0 + 1 || k0 + 1 > N - 1) s_v1 = v0; }
        __syncthreads();
    }
    const double a = __longlong_as_double((long long)s_v0), bq = __longlong_as_double((long long)s_v1);
    const double diff = __dadd_rn(bq, -a);
    double thr = __dadd_rn(a, __dmul_rn(diff, gamma));                       // numpy _lerp
    if (gamma >= 0.5) thr = __dadd_rn(bq, -__dmul_rn(diff, __dadd_rn(1.0, -gamma)));

    // ---- mask pixels in raster order
    if (tid == 0) s_nfg = 0;
    __syncthreads();
    for (int base = 0; base < N; base += D_THREADS) {
        const int i = base + tid;
        const bool on = (i < N) && (f[i] > thr);
        const unsigned int bal = __ballot_sync(0xffffffffu, on);
        if (lane == 0) s_warp_off[warp] = __popc(bal);
        __syncthreads();
        if (tid == 0) {
            int run = s_nfg;
            for (int w = 0; w < D_THREADS / 32; ++w) { const int c = s_warp_off[w]; s_warp_off[w] = run; run += c; }
            s_nfg = run;
        }
        __syncthreads();
        if (on) {
            const int pos = s_warp_off[warp] + __popc(bal & ((1u << lane) - 1u));
            if (pos < D_MAXFG) fg_idx[pos] = i;
        }
        __syncthreads();
    }
    // ---- the rest runs on the short list of mask pixels (at most 1 % of the image), all threads
    int n = s_nfg;
    if (tid == 0) { s_flag = n > D_MAXFG ? 1 : 0; s_total = 0; }
    n = min(n, D_MAXFG);
    auto lower = [&](int pix) {                // first list position whose pixel index is >= pix
        int lo = 0, hi = n;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (fg_idx[mid] < pix) lo = mid + 1; else hi = mid; }
        return lo;
    };
    auto find = [&](int pix) {                 // position of pixel in the sorted list, -1 if absent
        const int k = lower(pix);
        return (k < n && fg_idx[k] == pix) ? k : -1;
    };
    // 8-connected components (measure.label / remove_small_objects, connectivity 2): every pixel pulls the smallest
    // label of its 3x3 neighbourhood and then jumps to that label's label, until nothing changes.  The fixed point is
    // the list position of the component's first pixel in raster order - the root the sequential union-find (smaller
    // root wins) arrives at - whatever the interleaving of the threads: labels only decrease and stay inside the component.
    for (int i = tid; i < n; i += D_THREADS) { lab[i] = (short)i; csize[i] = 0; clast[i] = 0; korder[i] = -1; }
    __syncthreads();
    for (;;) {
        bool changed = false;
        for (int i = tid; i < n; i += D_THREADS) {
            const int y = fg_idx[i] / W, x = fg_idx[i] - y * W;
            const int xa = max(x - 1, 0), xb = min(x + 1, W - 1);
            int m = lab[i];
            for (int yy = max(y - 1, 0); yy <= min(y + 1, H - 1); ++yy) {
                const int pb = yy * W + xb;
                for (int k = lower(yy * W + xa); k < n && fg_idx[k] <= pb; ++k) m = min(m, (int)lab[k]);
            }
            m = min(m, (int)lab[m]);
            if (m < lab[i]) { lab[i] = (short)m; changed = true; }
        }
        if (!__syncthreads_or(changed)) break;
    }
    for (int i = tid; i < n; i += D_THREADS) { atomicAdd(&csize[lab[i]], 1); atomicMax(&clast[lab[i]], i); }
    __syncthreads();
    // remove_small_objects(min_size=4): pixels of small components get the label -1
    for (int i = tid; i < n; i += D_THREADS) if ((csize[lab[i]] & 0xFFFF) < 4) korder[i] = -2;
    __syncthreads();
    for (int i = tid; i < n; i += D_THREADS) if (korder[i] == -2) { lab[i] = -1; korder[i] = -1; }
    __syncthreads();
    // local maxima (peak_local_max inside one label): a pixel is a maximum when no pixel of ITS component within the 5x5
    // window is brighter.  One thread per mask pixel; the window's rows are short runs of the raster-ordered list.
    // csize[root]: size in the low 16 bits, number of maxima in the high 16; korder = -3 marks a candidate
    // (a maximum above the image minimum).
    for (int i = tid; i < n; i += D_THREADS) {
        const int c = lab[i];
        if (c < 0) continue;
        const int y = fg_idx[i] / W, x = fg_idx[i] - y * W;
        const int xa = max(x - 2, 0), xb = min(x + 2, W - 1);
        const double v = f[fg_idx[i]];
        bool is_max = true;
        for (int yy = max(y - 2, 0); yy <= min(y + 2, H - 1) && is_max; ++yy) {
            const int pb = yy * W + xb;
            for (int k = lower(yy * W + xa); k < n && fg_idx[k] <= pb; ++k)
                if (lab[k] == c && f[fg_idx[k]] > v) { is_max = false; break; }
        }
        if (is_max) { atomicAdd(&csize[c], 1 << 16); if (v > vmin) korder[i] = -3; }
    }
    __syncthreads();
    // per component: candidates in raster order, spacing.  One thread per component; its pixels lie between its first and
    // its last list position, a short range because the list is in raster order.
    for (int c = tid; c < n; c += D_THREADS) {
        if (lab[c] != c) continue;
        const int last = clast[c];
        const int size = csize[c] & 0xFFFF, npk = csize[c] >> 16;
        int flag = 0;
        int cand[64], nc = 0;
        for (int i = c; i <= last; ++i) {
            if (lab[i] != c || korder[i] != -3) continue;
            korder[i] = -1;
            if (nc < 64) cand[nc++] = i; else flag |= 2;
        }
        if (npk == size && size > 1) {
            // "trivial image": every pixel of the component is a maximum (all values equal).  skimage then keeps only
            // the pixels the binary opening (3x3 cross, zero border) removes: mask xor opening(mask).
            nc = 0;
            auto member = [&](int y, int x) {
                if (y < 0 || x < 0 || y >= H || x >= W) return false;
                const int j = find(y * W + x);
                return j >= 0 && lab[j] == c;
            };
            auto eroded = [&](int y, int x) {
                return member(y, x) && member(y - 1, x) && member(y + 1, x) && member(y, x - 1) && member(y, x + 1);
            };
            for (int i = c; i <= last; ++i) {
                if (lab[i] != c) continue;
                const int y = fg_idx[i] / W, x = fg_idx[i] - y * W;
                const bool opened = eroded(y, x) || eroded(y - 1, x) || eroded(y + 1, x) || eroded(y, x - 1) || eroded(y, x + 1);
                if (!opened && f[fg_idx[i]] > vmin) { if (nc < 64) cand[nc++] = i; else flag |= 2; }
            }
        }
        // sort by value, descending, stable in raster order (insertion sort on a handful of entries)
        for (int i = 1; i < nc; ++i) {
            const int ci = cand[i];
            const double vi = f[fg_idx[ci]];
            int j = i - 1;
            while (j >= 0 && f[fg_idx[cand[j]]] < vi) { cand[j + 1] = cand[j]; --j; }
            cand[j + 1] = ci;
        }
        // ensure_spacing: drop points closer than 2 (Chebyshev) to an already kept, brighter one
        int kept[64], nk = 0;
        for (int i = 0; i < nc; ++i) {
            const int y = fg_idx[cand[i]] / W, x = fg_idx[cand[i]] - y * W;
            bool ok = true;
            for (int k = 0; k < nk && ok; ++k) {
                const int yk = fg_idx[kept[k]] / W, xk = fg_idx[kept[k]] - yk * W;
                if (max(abs(yk - y), abs(xk - x)) < 2) ok = false;
            }
            if (ok) { korder[cand[i]] = (short)nk; kept[nk++] = cand[i]; }
        }
        csize[c] = nk;
        if (flag) atomicOr(&s_flag, flag);
    }
    __syncthreads();
    // output offsets: components in order of their first pixel (= measure.label numbering); exclusive scan by warp 0
    if (warp == 0) {
        int carry = 0;
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            const bool is_root = i < n && lab[i] == i;
            const int v = is_root ? csize[i] : 0;
            int incl = v;
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            if (is_root) csize[i] = carry + incl - v;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) s_total = carry;
    }
    __syncthreads();
    int* out = peaks + (size_t)b * D_MAXPEAKS * 2;
    for (int i = tid; i < n; i += D_THREADS) {
        if (korder[i] < 0) continue;
        const int pos = csize[lab[i]] + korder[i];
        if (pos < D_MAXPEAKS) { out[2 * pos] = fg_idx[i] / W; out[2 * pos + 1] = fg_idx[i] % W; }
    }
    if (tid == 0) {
        npeaks[b] = min(s_total, D_MAXPEAKS);
        flags[b] = s_flag | (s_total > D_MAXPEAKS ? 8 : 0);
    }
}

}  // namespace

extern "C" int sted_nanodomain_candidates(const float* sted_dev, int B, int H, int W, const double weights[3], int64_t k0,
                                          double gamma, double* scratch_dev, int32_t* peaks_dev, int32_t* npeaks_dev,
                                          int32_t* flags_dev, void* stream) {
    if (!sted_dev || !weights || !scratch_dev || !peaks_dev || !npeaks_dev || !flags_dev)
        return fail(STED_EINVAL, "null pointer argument");
    if (B <= 0 || H <= 0 || W <= 0) return fail(STED_EINVAL, "B, H, W must be positive");
    if ((long long)H * W > 100ll * D_MAXFG) return fail(STED_EUNSUPPORTED, "image larger than %d pixels", 100 * D_MAXFG);
    int rc = check_device();
    if (rc) return rc;
    const size_t n = (size_t)B * H * W;
    CUDA_TRY(cudaFuncSetAttribute(nanodomain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, D_SMEM_BYTES));
    nanodomain_kernel<<<B, D_THREADS, D_SMEM_BYTES, (cudaStream_t)stream>>>(sted_dev, H, W, weights[0], weights[1], weights[2], k0, gamma,
                                                                 scratch_dev, scratch_dev + n, peaks_dev, npeaks_dev, flags_dev);
    CUDA_TRY(cudaGetLastError());
    return STED_OK;
}

// =============================================================================================================
// Resolution objective: image-decorrelation analysis (gym_sted/rewards/objectives.py:146-335, Descloux et al. 2019).
// The host prepares the apodised, Fourier-transformed image (cuFFT through torch) with its pixels sorted by radius;
// this kernel runs all 22 decorrelation passes and the three dependent peak searches of one image in one CTA:
//   pass 0        : d(r_i) on r = linspace(0, 1, 50)                              -> r0, A0
//   passes 1..11  : high-pass sigma_0 = Hc/4, sigma_1..10 = exp(linspace(log g_max, log 0.15, 10))
//   passes 12..21 : refined radii around the best peak and 10 refined sigmas
//   resolution    = 2 / max(r_k with A_k >= 0.05) * pixelsize / 1 nm, capped.
// d(r) = sum_{R<r} Re(Ik' conj(Ikn)) / sqrt(sum_{R<r} |Ikn|^2 * sum |Ik'|^2), Ik' = Ik (1 - exp(-2 sigma^2 R^2)),
// then floor(1000 d)/1000.  Pixels are radius-sorted, so every ring [r_{i-1}, r_i) is a contiguous range: one warp
// per ring, lanes striding the range, fixed-order shuffles (deterministic), float64 throughout.
// =============================================================================================================
namespace {

constexpr int DC_NR = 50, DC_NG = 10, DC_THREADS = 512;

constexpr int DC_NS = DC_NG + 1;          // high-pass sigmas handled by one sweep over the spectrum
constexpr int DC_MAXU = 152, DC_CUTS = 96, DC_MIN_CHUNK = 1024;   // units <= rings + cuts

struct DecorrShared {
    int bounds[DC_NR];
    double ring_num[DC_NS][DC_NR + 1], ring_pow[DC_NS][DC_NR + 1], ring_abs[DC_NR + 1];
    double d[DC_NS][DC_NR], r[DC_NR], r_fine[DC_NR], sigmas[DC_NG + 1], sig_fine[DC_NG], rs[2 * DC_NG + 3], As[2 * DC_NG + 3];
    int idx[DC_NS];
    double A[DC_NS];
    // work units of a sweep: (ring, run of at most `chunk` radius-sorted pixels); a ring longer than a chunk is cut so that
    // the warps finish together (the first and the last ring of the refined radii hold most of the spectrum)
    int n_units, next_unit;
    int unit_lo[DC_MAXU], unit_hi[DC_MAXU];
    short ring_u0[DC_NR + 2];                  // first unit of every ring (ring_u0[ring + 1] = one past its last)
    double part_num[DC_NS][DC_MAXU], part_pow[DC_NS][DC_MAXU], part_abs[DC_MAXU];
};

__device__ void dc_linspace(double a, double b, int n, double* out) {            // numpy.linspace(a, b, n)
    const double step = (b - a) / (double)(n - 1);
    for (int i = 0; i < n; ++i) out[i] = (double)i * step + a;
    out[n - 1] = b;
}

// objectives.py:217-227: shorten the curve from the right until its maximum stands tol above the minimum to its right:
//     len = n; idx = argmax(d[:len]); while len > 1 and not (d[idx] - min(d[idx:len]) >= tol or idx == 0): len -= 1; idx = argmax(d[:len])
// done by one warp: lane l evaluates the candidate lengths l + 1 and l + 33 on their own (first maximum of the
// prefix, minimum from there to the end of the prefix, the reference's test); the answer is the longest prefix that
// passes, which is where the sequential loop stops (length 1 always passes: its maximum is element 0).
__device__ void dc_peak_warp(const double* d, double tol, int* idx_out, double* A_out) {
    const int lane = threadIdx.x & 31;
    auto eval = [&](int len, int* idx_o) {
        int k = 0;
        for (int i = 1; i < len; ++i) if (d[i] > d[k]) k = i;
        double mn = d[k];
        for (int i = k + 1; i < len; ++i) mn = fmin(mn, d[i]);
        *idx_o = k;
        return (d[k] - mn) >= tol || k == 0;
    };
    int idx_hi = 0, idx_lo = 0;
    const bool ok_hi = lane + 33 <= DC_NR && eval(lane + 33, &idx_hi);
    const bool ok_lo = eval(lane + 1, &idx_lo);
    const unsigned m_hi = __ballot_sync(0xffffffffu, ok_hi), m_lo = __ballot_sync(0xffffffffu, ok_lo);
    int idx;
    if (m_hi) idx = __shfl_sync(0xffffffffu, idx_hi, 31 - __clz(m_hi));
    else idx = __shfl_sync(0xffffffffu, idx_lo, 31 - __clz(m_lo));
    if (lane == 0) { *idx_out = idx; *A_out = d[idx]; }
}

__device__ void dc_bounds(DecorrShared& sh, const double* __restrict__ R, int N, int n_inside, const double* radii) {
    if (threadIdx.x < DC_NR) {                 // numpy.searchsorted(R_sorted, r, 'left') = number of pixels with R < r
        const double r = radii[threadIdx.x];
        int lo = 0, hi = N;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (R[mid] < r) lo = mid + 1; else hi = mid; }
        sh.bounds[threadIdx.x] = lo;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int chunk = (n_inside + DC_CUTS - 1) / DC_CUTS;
        chunk = max(DC_MIN_CHUNK, (chunk + 31) & ~31);
        int u = 0;
        for (int ring = 0; ring <= DC_NR; ++ring) {
            const int lo = ring == 0 ? 0 : min(sh.bounds[ring - 1], n_inside);
            const int hi = ring < DC_NR ? min(sh.bounds[ring], n_inside) : n_inside;
            sh.ring_u0[ring] = (short)u;
            for (int s0 = lo; s0 < hi; s0 += chunk) { sh.unit_lo[u] = s0; sh.unit_hi[u] = min(s0 + chunk, hi); ++u; }
        }
        sh.ring_u0[DC_NR + 1] = (short)u;
        sh.n_units = u;
    }
    __syncthreads();
}

// One sweep over the radius-sorted spectrum for NS high-pass sigmas at once (sigma == 0: no filter).  The decorrelation
// curves of the sigmas of one stage do not depend on each other, so the 1 + 11 + 10 passes of the reference become three
// sweeps: a pixel is loaded once and weighted NS times.  Per sigma the additions are those of a pass of its own, in the
// same order - lanes striding the ring, the same shuffle tree, the running sums in numpy's order - so the curves are
// bit-identical to the pass-per-sigma form (a ring cut into several units is summed unit by unit, in order).
template <int NS>
__device__ void dc_sweep(DecorrShared& sh, const double2* __restrict__ Ik, const double2* __restrict__ Ikn,
                         const double* __restrict__ R2, int n_inside, const double* sig, double tol) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double c[NS];
    bool any = false;
#pragma unroll
    for (int j = 0; j < NS; ++j) { c[j] = -2.0 * (sig[j] * sig[j]); any |= sig[j] != 0.0; }
    if (threadIdx.x == 0) sh.next_unit = 0;
    __syncthreads();
    for (;;) {
        int u = 0;
        if (lane == 0) u = atomicAdd(&sh.next_unit, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= sh.n_units) break;
        const int lo = sh.unit_lo[u], hi = sh.unit_hi[u];
        double sn[NS], sp[NS], sa = 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j) sn[j] = sp[j] = 0.0;
        for (int p = lo + lane; p < hi; p += 32) {
            const double2 a0 = Ik[p];
            const double2 b = Ikn[p];
            const double r2 = any ? R2[p] : 0.0;
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                double2 a = a0;
                if (sig[j] != 0.0) { const double f = 1.0 - exp(c[j] * r2); a.x *= f; a.y *= f; }
                sn[j] += a.x * b.x + a.y * b.y;
                sp[j] += a.x * a.x + a.y * a.y;
            }
            sa += b.x * b.x + b.y * b.y;
        }
        for (int o = 16; o; o >>= 1) {
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                sn[j] += __shfl_xor_sync(0xffffffffu, sn[j], o);
                sp[j] += __shfl_xor_sync(0xffffffffu, sp[j], o);
            }
            sa += __shfl_xor_sync(0xffffffffu, sa, o);
        }
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < NS; ++j) { sh.part_num[j][u] = sn[j]; sh.part_pow[j][u] = sp[j]; }
            sh.part_abs[u] = sa;
        }
    }
    __syncthreads();
    // ring sums: the units of a ring in order (which warp computed a unit does not matter: the sums are deterministic)
    for (int t = threadIdx.x; t < (2 * NS + 1) * (DC_NR + 1); t += DC_THREADS) {
        const int q = t / (DC_NR + 1), ring = t - q * (DC_NR + 1);
        const double* part = q < NS ? sh.part_num[q] : (q < 2 * NS ? sh.part_pow[q - NS] : sh.part_abs);
        double v = 0.0;
        for (int u = sh.ring_u0[ring]; u < sh.ring_u0[ring + 1]; ++u) v += part[u];
        if (q < NS) sh.ring_num[q][ring] = v; else if (q < 2 * NS) sh.ring_pow[q - NS][ring] = v; else sh.ring_abs[ring] = v;
    }
    __syncthreads();
    // the curves d(r_i) and their peaks: the running sums keep numpy's order (one thread per sigma, 150 additions); the
    // divisions, square roots and the peak searches - O(n^2) comparisons in the reference's loop - are spread over the CTA
    if (threadIdx.x < NS) {
        const int j = threadIdx.x;
        double denom2 = 0.0;
        for (int i = 0; i <= DC_NR; ++i) denom2 += sh.ring_pow[j][i];
        double cn = 0.0, ca = 0.0;
        for (int i = 0; i < DC_NR; ++i) {
            cn += sh.ring_num[j][i];
            ca += sh.ring_abs[i];
            sh.ring_num[j][i] = cn;                      // in place: cumulative numerator, cumulative |Ikn|^2 * total power
            sh.ring_pow[j][i] = ca * denom2;
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < NS * DC_NR; t += DC_THREADS) {
        const int j = t / DC_NR, i = t - j * DC_NR;
        double v = sh.ring_num[j][i] / sqrt(sh.ring_pow[j][i]);
        v = floor(1000.0 * v) / 1000.0;
        sh.d[j][i] = isnan(v) ? 0.0 : v;
    }
    __syncthreads();
    static_assert(NS <= DC_THREADS / 32, "one warp per sigma for the peak search");
    if (warp < NS) dc_peak_warp(sh.d[warp], tol, &sh.idx[warp], &sh.A[warp]);
    __syncthreads();
}

__global__ void __launch_bounds__(DC_THREADS) decorr_kernel(const double2* __restrict__ Ik_all, const double2* __restrict__ Ikn_all,
                                                            const double* __restrict__ R, const double* __restrict__ R2, int N,
                                                            int n_inside, const double* __restrict__ r_coarse, double sigma0,
                                                            double gmax_flat, double pixelsize, double res_cap, double tol,
                                                            double* __restrict__ res_out) {
    __shared__ DecorrShared sh;
    const double2* Ik = Ik_all + (size_t)blockIdx.x * N;
    const double2* Ikn = Ikn_all + (size_t)blockIdx.x * N;
    if (threadIdx.x < DC_NR) sh.r[threadIdx.x] = r_coarse[threadIdx.x];
    __syncthreads();
    dc_bounds(sh, R, N, n_inside, sh.r);
    if (threadIdx.x == 0) sh.sigmas[0] = 0.0;
    __syncthreads();
    dc_sweep<1>(sh, Ik, Ikn, R2, n_inside, sh.sigmas, tol);
    if (threadIdx.x == 0) {
        const double r0 = sh.r[sh.idx[0]];
        sh.rs[0] = r0;
        sh.As[0] = sh.A[0];
        double g_max = 2.0 / r0;
        if (isinf(g_max)) g_max = gmax_flat;
        sh.sigmas[0] = sigma0;
        dc_linspace(log(g_max), log(0.15), DC_NG, sh.sigmas + 1);
        for (int j = 1; j <= DC_NG; ++j) sh.sigmas[j] = exp(sh.sigmas[j]);
    }
    __syncthreads();
    dc_sweep<DC_NG + 1>(sh, Ik, Ikn, R2, n_inside, sh.sigmas, tol);
    if (threadIdx.x == 0) {
        for (int j = 0; j <= DC_NG; ++j) { sh.rs[1 + j] = sh.r[sh.idx[j]]; sh.As[1 + j] = sh.A[j]; }
        const int n = DC_NG + 2;
        double mx = sh.rs[0];
        for (int i = 1; i < n; ++i) mx = fmax(mx, sh.rs[i]);
        int k = 0;
        for (int i = 0; i < n; ++i) if (sh.rs[i] == mx) k = i;                     // last index of the maximum
        const int ind1 = k == 0 ? 0 : (k >= DC_NG ? k - 2 : k - 1);
        dc_linspace(sh.rs[k] - (sh.r[1] - sh.r[0]), fmin(sh.rs[k] + 0.4, sh.r[DC_NR - 1]), DC_NR, sh.r_fine);
        dc_linspace(log(sh.sigmas[ind1]), log(sh.sigmas[ind1 + 1]), DC_NG, sh.sig_fine);
        for (int j = 0; j < DC_NG; ++j) sh.sig_fine[j] = exp(sh.sig_fine[j]);
    }
    __syncthreads();
    dc_bounds(sh, R, N, n_inside, sh.r_fine);
    dc_sweep<DC_NG>(sh, Ik, Ikn, R2, n_inside, sh.sig_fine, tol);
    // the last coarse entry (index DC_NG+1) is dropped by the reference before the refined ones are appended
    if (threadIdx.x == 0)
        for (int j = 0; j < DC_NG; ++j) { sh.rs[DC_NG + 1 + j] = sh.r_fine[sh.idx[j]]; sh.As[DC_NG + 1 + j] = sh.A[j]; }
    if (threadIdx.x == 0) {
        const int n = 2 * DC_NG + 2;
        sh.rs[n - 1] = sh.rs[0];
        sh.As[n - 1] = sh.As[0];
        double mx = -1.0e300;
        for (int i = 0; i < n; ++i) mx = fmax(mx, sh.As[i] < 0.05 ? 0.0 : sh.rs[i]);
        const double res = 2.0 / mx * pixelsize / 1e-9;
        res_out[blockIdx.x] = res > res_cap ? res_cap : res;
    }
}

}  // namespace

extern "C" int sted_decorrelation_resolution(const void* ik_dev, const void* ikn_dev, int B, int N, int n_inside,
                                             const double* radius_dev, const double* radius2_dev, const double* r_coarse_dev,
                                             double sigma0, double gmax_flat, double pixelsize, double res_cap, double tol,
                                             double* res_dev, void* stream) {
    if (!ik_dev || !ikn_dev || !radius_dev || !radius2_dev || !r_coarse_dev || !res_dev) return fail(STED_EINVAL, "null pointer argument");
    if (B <= 0 || N <= 0 || n_inside < 0 || n_inside > N) return fail(STED_EINVAL, "bad decorrelation sizes");
    int rc = check_device();
    if (rc) return rc;
    decorr_kernel<<<B, DC_THREADS, 0, (cudaStream_t)stream>>>((const double2*)ik_dev, (const double2*)ikn_dev, radius_dev, radius2_dev,
                                                             N, n_inside, r_coarse_dev, sigma0, gmax_flat, pixelsize, res_cap, tol,
                                                             res_dev);
    CUDA_TRY(cudaGetLastError());
    return STED_OK;
}

// =============================================================================================================
// Gaussian-fit validation of the candidates (gym_sted/rewards/utils.py:6-49): for every candidate a 7x7 window of
// the (zero-padded) STED image is fitted with amplitude*exp(-((cx-x)^2/(2 wx^2) + (cy-y)^2/(2 wy^2))) by
// Levenberg-Marquardt, and the candidate is kept when 40 nm <= 2.355*w*pixelsize <= 200 nm for both widths.
// The reference calls scipy.optimize.leastsq(errorfunction, [data.max(), 0, 0, 1, 1], ftol=1e-3), i.e. MINPACK's
// LMDIF with xtol = 1.49012e-8, gtol = 0, maxfev = 200*(n+1), epsfcn = eps, factor = 100, mode 1.  lm_minpack.cuh
// restates LMDIF / FDJAC2 / QRFAC / LMPAR / QRSOLV / ENORM (Moré, Garbow, Hillstrom, MINPACK-1) for
// m = 49 residuals and n = 5 parameters, one fit per thread, in float64.  The residuals go through exp(), whose
// last bit differs between libm/numpy and CUDA, so fitted parameters agree to ~1e-9 rather than bitwise; the
// keep/reject decisions are identical except for widths that sit within that distance of a limit.
// =============================================================================================================
namespace {

using namespace lm;

// The residuals of one fit evaluated by a whole warp for lm_fit_warp (lm_minpack_warp.cuh): lane l computes residuals l
// and l + 32 (the 49 exp() calls are most of a residual evaluation) and writes them to the warp's shared vector.  Each
// value is the one GaussResiduals computes, bit for bit.
template <bool RATIONAL>
struct WarpResiduals {
    const double* data;       // the 49 window values (shared memory)
    __device__ LMW_NOINLINE void operator()(const double* p, double* f) const {
        const int lane = threadIdx.x & 31;
        const double wx = p[3] + 1e-3, wy = p[4] + 1e-3;
        const double dx2 = 2.0 * (wx * wx), dy2 = 2.0 * (wy * wy);
        auto one = [&](int i) {
            const double X = (double)(i / 7) - 3.0, Y = (double)(i % 7) - 3.0;
            const double a = (p[1] - X) * (p[1] - X) / dx2, b = (p[2] - Y) * (p[2] - Y) / dy2;
            return (RATIONAL ? p[0] / (1.0 + (a + b)) : p[0] * exp(-1.0 * (a + b))) - data[i];
        };
        __syncwarp();                               // earlier readers of f are done
        f[lane] = one(lane);
        if (lane + 32 < LM_M) f[lane + 32] = one(lane + 32);
        __syncwarp();
    }
};

constexpr int GF_THREADS = 128;      // 4 fits per CTA, one warp each
#ifndef GF_MINB
#define GF_MINB 4
#endif

__global__ void __launch_bounds__(GF_THREADS, GF_MINB) gaussfit_kernel(const float* __restrict__ sted, int H, int W, const int* __restrict__ peaks,
                                                            const int* __restrict__ npeaks, int maxp, double pixelsize,
                                                            unsigned char* __restrict__ valid, double* __restrict__ params) {
    __shared__ LmWarpWork work[GF_THREADS / 32];
    const int b = blockIdx.y, k = blockIdx.x * (GF_THREADS / 32) + (threadIdx.x >> 5);      // warp-uniform
    if (k >= npeaks[b] || k >= maxp) return;
    LmWarpWork& w = work[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int row = peaks[((size_t)b * maxp + k) * 2], col = peaks[((size_t)b * maxp + k) * 2 + 1];
    const float* img = sted + (size_t)b * H * W;
    double dmax = -1.0e300;
    for (int i = lane; i < LM_M; i += 32) {
        const int y = row + i / 7 - 3, x = col + i % 7 - 3;
        w.data[i] = (y >= 0 && y < H && x >= 0 && x < W) ? rint((double)img[y * W + x]) : 0.0;     // numpy.pad(..., constant_values=0)
        dmax = fmax(dmax, w.data[i]);
    }
    for (int o = 16; o; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    __syncwarp();
    double p[LM_N] = {dmax, 0.0, 0.0, 1.0, 1.0};
    lm_fit_warp(WarpResiduals<false>{w.data}, w, p, 1e-3, 1.49012e-8, 0.0, 200 * (LM_N + 1), 100.0);
    if (lane != 0) return;
    const double fx = 2.355 * p[3] * pixelsize, fy = 2.355 * p[4] * pixelsize;
    valid[(size_t)b * maxp + k] = !(fx > 200e-9 || fx < 40e-9 || fy > 200e-9 || fy < 40e-9);
    if (params) for (int j = 0; j < LM_N; ++j) params[((size_t)b * maxp + k) * LM_N + j] = p[j];
}

// Test hook: run the solver on n windows with caller-supplied starts; model 0 = the Gaussian, 1 = the rational test model
// (one thread per fit, lm_fit); models 2 / 3 = the same two models through the warp-cooperative lm_fit_warp (one warp per fit).
__global__ void lmfit_test_kernel(const double* __restrict__ data, int n, int model, double ftol, double* __restrict__ params,
                                  int* __restrict__ info) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double d[LM_M], p[LM_N];
    for (int i = 0; i < LM_M; ++i) d[i] = data[(size_t)k * LM_M + i];
    for (int j = 0; j < LM_N; ++j) p[j] = params[(size_t)k * LM_N + j];
    const int rc = model == 0 ? lm_fit(GaussResiduals{d}, p, ftol, 1.49012e-8, 0.0, 200 * (LM_N + 1), 100.0)
                              : lm_fit(RationalResiduals{d}, p, ftol, 1.49012e-8, 0.0, 200 * (LM_N + 1), 100.0);
    for (int j = 0; j < LM_N; ++j) params[(size_t)k * LM_N + j] = p[j];
    info[k] = rc;
}

__global__ void __launch_bounds__(GF_THREADS) lmfit_warp_test_kernel(const double* __restrict__ data, int n, int model, double ftol,
                                                                     double* __restrict__ params, int* __restrict__ info) {
    __shared__ LmWarpWork work[GF_THREADS / 32];
    const int k = blockIdx.x * (GF_THREADS / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= n) return;
    LmWarpWork& w = work[threadIdx.x >> 5];
    for (int i = lane; i < LM_M; i += 32) w.data[i] = data[(size_t)k * LM_M + i];
    __syncwarp();
    double p[LM_N];
    for (int j = 0; j < LM_N; ++j) p[j] = params[(size_t)k * LM_N + j];
    __syncwarp();
    const int rc = model == 2 ? lm_fit_warp(WarpResiduals<false>{w.data}, w, p, ftol, 1.49012e-8, 0.0, 200 * (LM_N + 1), 100.0)
                              : lm_fit_warp(WarpResiduals<true>{w.data}, w, p, ftol, 1.49012e-8, 0.0, 200 * (LM_N + 1), 100.0);
    if (lane != 0) return;
    for (int j = 0; j < LM_N; ++j) params[(size_t)k * LM_N + j] = p[j];
    info[k] = rc;
}

}  // namespace

extern "C" int sted_lmfit_test(const double* data_dev, int n, int model, double ftol, double* params_dev, int32_t* info_dev,
                               void* stream) {
    if (!data_dev || !params_dev || !info_dev || n <= 0 || model < 0 || model > 3) return fail(STED_EINVAL, "bad lmfit arguments");
    int rc = check_device();
    if (rc) return rc;
    if (model >= 2)
        lmfit_warp_test_kernel<<<(n + GF_THREADS / 32 - 1) / (GF_THREADS / 32), GF_THREADS, 0, (cudaStream_t)stream>>>(data_dev, n, model, ftol,
                                                                                                                 params_dev, info_dev);
    else
        lmfit_test_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(data_dev, n, model, ftol, params_dev, info_dev);
    CUDA_TRY(cudaGetLastError());
    return STED_OK;
}

extern "C" int sted_gaussfit_validate(const float* sted_dev, int B, int H, int W, const int32_t* peaks_dev, const int32_t* npeaks_dev,
                                      int max_peaks, double pixelsize, uint8_t* valid_dev, double* params_dev, void* stream) {
    if (!sted_dev || !peaks_dev || !npeaks_dev || !valid_dev) return fail(STED_EINVAL, "null pointer argument");
    if (B <= 0 || H <= 0 || W <= 0 || max_peaks <= 0) return fail(STED_EINVAL, "B, H, W, max_peaks must be positive");
    int rc = check_device();
    if (rc) return rc;
    dim3 grid((max_peaks + GF_THREADS / 32 - 1) / (GF_THREADS / 32), B);
    gaussfit_kernel<<<grid, GF_THREADS, 0, (cudaStream_t)stream>>>(sted_dev, H, W, peaks_dev, npeaks_dev, max_peaks, pixelsize, valid_dev,
                                                          params_dev);
    CUDA_TRY(cudaGetLastError());
    return STED_OK;
}

// =============================================================================================================
// F1 of the nanodomain detections (gym_sted/rewards/objectives.py:425-430): metrics.CentroidDetectionError(truth,
// guess, threshold, algorithm="hungarian").f1_score.  The Hungarian solution on the cost matrix with the pairs at
// distance >= threshold priced out assigns as many allowed pairs as possible (one forbidden pair costs more than all
// allowed ones together), so TP is the size of a MAXIMUM MATCHING of the bipartite graph "truth i - guess j closer than
// the threshold"; FP and FN are what is left on either side.  Coordinates are integer pixels, so the distance test is
// exact in integers (threshold^2 as a double compared with an integer sum of squares).  One thread per image runs
// Kuhn's augmenting-path algorithm with an explicit stack (graphs are tiny and very sparse: <= 9 integer points lie
// within 2 px of a truth); F1 = 2PR/(P+R+1e-6), P = TP/(TP+FP+1e-6), R = TP/(TP+FN+1e-6) in float64, numpy's order.
// =============================================================================================================
namespace {

constexpr int F1_MAXN = 512;         // truths and kept guesses per image

constexpr int F1_THREADS = 128, F1_ADJ = 8;

__global__ void __launch_bounds__(F1_THREADS) f1_match_kernel(const int* __restrict__ peaks, const int* __restrict__ npeaks,
                                                               const unsigned char* __restrict__ valid, int maxp,
                                                               const int* __restrict__ truth, const int* __restrict__ ntruth, int maxt,
                                                               double thr2, int* __restrict__ counts, double* __restrict__ f1,
                                                               int* __restrict__ flags) {
    __shared__ short gy[F1_MAXN], gx[F1_MAXN], ty[F1_MAXN], tx[F1_MAXN];
    __shared__ short match_g[F1_MAXN];               // truth matched to guess g, -1 if free
    __shared__ short stack_t[F1_MAXN], stack_j[F1_MAXN], stack_g[F1_MAXN];
    __shared__ short adj[F1_MAXN][F1_ADJ];           // the guesses within the threshold of truth t, ascending
    __shared__ unsigned char adjn[F1_MAXN];          // how many (255: more than F1_ADJ, scan all guesses instead)
    __shared__ unsigned visited[F1_MAXN / 32];
    __shared__ int s_ng;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int np_ = min(npeaks[b], maxp), nt = min(ntruth[b], maxt);
    // compact the kept guesses (order preserved: warp-wide ballots, warp 0)
    if (tid == 0) s_ng = 0;
    __syncthreads();
    bool overflow = nt > F1_MAXN;
    if (tid < 32) {
        for (int base = 0; base < np_; base += 32) {
            const int k = base + lane;
            const bool keep = k < np_ && (!valid || valid[(size_t)b * maxp + k]);
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            const int at = s_ng + __popc(m & ((1u << lane) - 1));
            if (keep && at < F1_MAXN) {
                gy[at] = (short)peaks[((size_t)b * maxp + k) * 2];
                gx[at] = (short)peaks[((size_t)b * maxp + k) * 2 + 1];
            }
            __syncwarp();
            if (lane == 0) s_ng += __popc(m);
            __syncwarp();
        }
    }
    __syncthreads();
    const int ng_all = s_ng;
    overflow |= ng_all > F1_MAXN;
    const int ng = min(ng_all, F1_MAXN), ntc = min(nt, F1_MAXN);
    for (int i = tid; i < ntc; i += F1_THREADS) {
        ty[i] = (short)truth[((size_t)b * maxt + i) * 2];
        tx[i] = (short)truth[((size_t)b * maxt + i) * 2 + 1];
    }
    for (int i = tid; i < ng; i += F1_THREADS) match_g[i] = -1;
    __syncthreads();
    auto close_pair = [&](int t, int g) {
        const int dy = (int)ty[t] - (int)gy[g], dx = (int)tx[t] - (int)gx[g];
        return (double)(dy * dy + dx * dx) < thr2;
    };
    // admissible pairs, one thread per truth: the search below then walks a few entries instead of every guess
    for (int t = tid; t < ntc; t += F1_THREADS) {
        int n = 0;
        for (int g = 0; g < ng; ++g)
            if (close_pair(t, g)) { if (n < F1_ADJ) adj[t][n] = (short)g; ++n; }
        adjn[t] = n > F1_ADJ ? 255 : (unsigned char)n;
    }
    __syncthreads();
    if (tid != 0) return;
    int tp = 0;
    for (int t0 = 0; t0 < ntc; ++t0) {
        if (adjn[t0] == 0) continue;
        for (int i = 0; i < (ng + 31) / 32; ++i) visited[i] = 0u;
        // iterative DFS: frame d tries the admissible guesses of truth stack_t[d] from position stack_j[d] on, in ascending
        // order; stack_g[d] = guess taken to descend
        int d = 0;
        stack_t[0] = (short)t0; stack_j[0] = 0;
        bool found = false;
        while (d >= 0) {
            const int t = stack_t[d];
            const int na = adjn[t];
            int j = stack_j[d];
            bool descended = false;
            for (;; ++j) {
                int g;
                if (na == 255) { if (j >= ng) break; g = j; if (!close_pair(t, g)) continue; }
                else { if (j >= na) break; g = adj[t][j]; }
                if ((visited[g >> 5] >> (g & 31)) & 1u) continue;
                visited[g >> 5] |= 1u << (g & 31);
                stack_g[d] = (short)g;
                if (match_g[g] < 0) { found = true; break; }
                stack_j[d] = (short)(j + 1);
                ++d;
                stack_t[d] = match_g[g]; stack_j[d] = 0;
                descended = true;
                break;
            }
            if (found) break;
            if (!descended) --d;                       // dead end: back to the caller, which continues after its guess
        }
        if (found) {
            for (int k = d; k >= 0; --k) match_g[stack_g[k]] = stack_t[k];   // flip the augmenting path
            ++tp;
        }
    }
    const int fp = ng - tp, fn = ntc - tp;
    counts[b * 3] = tp; counts[b * 3 + 1] = fp; counts[b * 3 + 2] = fn;
    const double dtp = (double)tp, dfp = (double)fp, dfn = (double)fn;
    const double prec = dtp / (dtp + dfp + 1e-6), rec = dtp / (dtp + dfn + 1e-6);
    f1[b] = 2.0 * prec * rec / (prec + rec + 1e-6);
    if (flags) flags[b] = overflow ? 1 : 0;
}

// ---- apodisation + centring in front of the FFT (objectives.py:150-160): window * (image - mean) + mean rounded to
// float32 like the reference's astype(float32), cropped to odd sides, ifftshift-ed so that a plain forward FFT
// followed by the gather below equals fftshift(fftn(fftshift(image))).  The mean of an integer image is exact (integer
// sum / N), whatever the summation order.
constexpr int AP_THREADS = 256;

__global__ void __launch_bounds__(AP_THREADS) apodize_kernel(const float* __restrict__ img, const double* __restrict__ window,
                                                            int H, int W, int Hc, int Wc, double2* __restrict__ out) {
    __shared__ long long s_part[AP_THREADS / 32];
    __shared__ double s_mean;
    const int b = blockIdx.x;
    const float* im = img + (size_t)b * H * W;
    long long acc = 0;
    for (int i = threadIdx.x; i < H * W; i += AP_THREADS) acc += (long long)rintf(im[i]);
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int i = 0; i < AP_THREADS / 32; ++i) t += s_part[i];
        s_mean = (double)t / (double)(H * W);
    }
    __syncthreads();
    const double mean = s_mean;
    double2* o = out + (size_t)b * Hc * Wc;
    const int sy = Hc / 2, sx = Wc / 2;                  // fftshift of an odd length n moves index i to (i + n/2) % n
    for (int i = threadIdx.x; i < Hc * Wc; i += AP_THREADS) {
        const int y = i / Wc, x = i - y * Wc;
        const double v = (double)(float)(window[y * W + x] * ((double)im[y * W + x] - mean) + mean);
        const int yd = (y + sy) % Hc, xd = (x + sx) % Wc;
        o[yd * Wc + xd] = make_double2(v, 0.0);
    }
}

// ---- after the FFT: centre the spectrum, take the pixels in radius order, zero those at R >= 1, and form the
// phase-only copy Ik/|Ik| (0 where |Ik| = 0 or not finite) — the two inputs of decorr_kernel.
__global__ void __launch_bounds__(256) spectrum_gather_kernel(const double2* __restrict__ spec, const long long* __restrict__ order,
                                                             int N, int n_inside, int Hc, int Wc, double2* __restrict__ ik,
                                                             double2* __restrict__ ikn) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double2 v = make_double2(0.0, 0.0), n = make_double2(0.0, 0.0);
    if (i < n_inside) {
        const int src = (int)order[i];                   // index in the centred (fftshift-ed) spectrum
        const int y = src / Wc, x = src - y * Wc;
        const int ys = (y + Hc - Hc / 2) % Hc, xs = (x + Wc - Wc / 2) % Wc;      // where the plain FFT holds it
        v = spec[(size_t)b * N + (size_t)ys * Wc + xs];
        const double mag = hypot(v.x, v.y);              // numpy.abs of a complex number
        if (mag > 0.0) {
            n = make_double2(v.x / mag, v.y / mag);
            if (!(isfinite(n.x) && isfinite(n.y))) n = make_double2(0.0, 0.0);
        }
    }
    ik[(size_t)b * N + i] = v;
    ikn[(size_t)b * N + i] = n;
}

}  // namespace

extern "C" int sted_f1_match(const int32_t* peaks_dev, const int32_t* npeaks_dev, const uint8_t* valid_dev, int max_peaks,
                             const int32_t* truth_dev, const int32_t* ntruth_dev, int max_truth, int B, double threshold,
                             int32_t* counts_dev, double* f1_dev, int32_t* flags_dev, void* stream) {
    if (!peaks_dev || !npeaks_dev || !truth_dev || !ntruth_dev || !counts_dev || !f1_dev) return fail(STED_EINVAL, "null pointer argument");
    if (B <= 0 || max_peaks <= 0 || max_truth <= 0) return fail(STED_EINVAL, "B, max_peaks, max_truth must be positive");
    int rc = check_device();
    if (rc) return rc;
    f1_match_kernel<<<B, F1_THREADS, 0, (cudaStream_t)stream>>>(peaks_dev, npeaks_dev, valid_dev, max_peaks, truth_dev, ntruth_dev, max_truth,
                                                       threshold * threshold, counts_dev, f1_dev, flags_dev);
    CUDA_TRY(cudaGetLastError());
    return STED_OK;
}

extern "C" int sted_apodize_for_fft(const float* image_dev, const double* window_dev, int B, int H, int W, int Hc, int Wc,
                                    void* out_dev, void* stream) {
    if (!image_dev || !window_dev || !out_dev) return fail(STED_EINVAL, "null pointer argument");
    if (B <= 0 || H <= 0 || W <= 0 || Hc <= 0 || Wc <= 0 || Hc > H || Wc > W) return fail(STED_EINVAL, "bad shape");
    int rc = check_device();
    if (rc) return rc;
    apodize_kernel<<<B, AP_THREADS, 0, (cudaStream_t)stream>>>(image_dev, window_dev, H, W, Hc, Wc, (double2*)out_dev);
    CUDA_TRY(cudaGetLastError());
    return STED_OK;
}

extern "C" int sted_spectrum_gather(const void* spectrum_dev, const int64_t* order_dev, int B, int Hc, int Wc, int n_inside,
                                    void* ik_dev, void* ikn_dev, void* stream) {
    if (!spectrum_dev || !order_dev || !ik_dev || !ikn_dev) return fail(STED_EINVAL, "null pointer argument");
    if (B <= 0 || Hc <= 0 || Wc <= 0 || n_inside < 0 || n_inside > Hc * Wc) return fail(STED_EINVAL, "bad shape");
    int rc = check_device();
    if (rc) return rc;
    const int N = Hc * Wc;
    dim3 grid((N + 255) / 256, B);
    spectrum_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const double2*)spectrum_dev, (const long long*)order_dev, N, n_inside,
                                                                  Hc, Wc, (double2*)ik_dev, (double2*)ikn_dev);
    CUDA_TRY(cudaGetLastError());
    return STED_OK;
}
