/*
 * sted_b200.h — C ABI of the B200-native STED acquisition hot path.
 *
 * Drop-in boundary for the path gym-sted reaches through
 *   pysted.base.Microscope.get_signal_and_bleach   (reference call sites:
 *       gym_sted/envs/mosted_env.py:184,222,227,232)
 *   pysted.base.Microscope.get_effective / fluo.get_k_bleach / detector.get_signal
 *       (gym_sted/utils.py:616-631,646-657)
 *   gym_sted.utils.get_foreground + rewards.objectives.{Signal_Ratio,Bleach,Resolution,
 *       NumberNanodomains}                        (gym_sted/utils.py:16-24,
 *       gym_sted/rewards/objectives.py:57-430)
 *
 * The reference has no FFI layer of its own (pure Python calling pySTED's Cython); these
 * entry points are what a ctypes stub on the reference side binds — see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; every `*_dev` pointer is DEVICE memory owned by the
 *     caller, every `*_host` pointer is HOST memory owned by the caller;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); device entry points
 *     are stream-ordered and never synchronise; `*_host` entry points synchronise before
 *     returning (their outputs are host memory);
 *   - return 0 on success, a negative STED_E* code otherwise; sted_last_error() gives text;
 *   - outputs are fully overwritten; no allocation crosses the ABI (the `*_host` calls keep a
 *     private, growable device workspace per process);
 *   - there is no CPU fallback: without a CUDA device every compute call fails with
 *     STED_ENODEVICE.
 *
 * Device datamap layout ("molecule map"): float32 [B][Hp][Wpitch], Hp = H + K - 1,
 *   Wp = W + K - 1, Wpitch = Wp rounded up to a multiple of 4 (16-byte rows for TMA / float4).
 *   Counts are integer-valued floats (< 2^24) in stochastic mode, expected counts otherwise.
 *   The K/2-pixel frame around the H x W region of interest is the zero padding
 *   pysted.base.Datamap.set_roi(i_ex, "max") adds (gym_sted/utils.py:179-180).
 */
#ifndef STED_B200_H
#define STED_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STED_OK            0
#define STED_EINVAL       -1   /* bad argument (shape, alignment, null pointer, aliasing)   */
#define STED_ENODEVICE    -2   /* no usable CUDA device / driver                            */
#define STED_ECUDA        -3   /* a CUDA runtime call or kernel launch failed                */
#define STED_EUNSUPPORTED -4   /* shape outside the compiled limits (K > STED_MAX_K, ...)    */
#define STED_EOVERFLOW    -5   /* a device-side capacity was exceeded; the outputs are invalid */

#define STED_MAX_K        63   /* largest PSF side (odd)                                     */
#define STED_KPITCH(K)    ((((K) + 1) + 3) & ~3)   /* row pitch of the per-env K x K tables  */

/* flags of sted_acquire* */
#define STED_BLEACH         1u  /* apply in-scan photobleaching (pySTED bleach=True)          */
#define STED_SAMPLE_PHOTONS 2u  /* detector noise: Poisson photon counts (Detector.noise)     */
#define STED_SAMPLE_BLEACH  4u  /* binomial molecule thinning; otherwise expected counts      */
#define STED_HOST_U16      16u  /* sted_acquire_host only: dmap_host and signal_host are uint16 counts (half the bus
                                 * traffic; saturating at 65535).  Needs integer outputs: STED_SAMPLE_PHOTONS, and
                                 * STED_SAMPLE_BLEACH when STED_BLEACH is set.                                     */
#define STED_NO_DETECT      8u  /* sted_acquire leaves the detected power (W) in signal_dev; the
                                   detector is applied later by sted_detect                   */
#define STED_FRAME_MOLECULES 32u /* the K/2 frame of dmap_in_dev may hold molecules (a caller assigned the whole padded
                                  * map, gym_sted/envs/sted_env.py:170): the stochastic scan then draws deaths for the
                                  * whole padded map instead of the H x W region of interest only                   */

#define STED_PRESAMPLED     64u  /* sted_acquire: the workspace already holds this scan's death schedule, written by
                                  * sted_bleach_presample with the same maps, tables, dwell times, flags, seed, offset and
                                  * env_base                                                                          */

/* Fluorophore + STED-pulse parameters of one environment.
 * Mirrors pysted.base.Fluorescence(**defaults.FLUO) (gym_sted/defaults.py:25-42) and the
 * lambda_/tau/rate attributes gym-sted reads from microscope.excitation / microscope.sted
 * (gym_sted/utils.py:620-623). */
typedef struct StedFluoParams {
    double lambda_ex;        /* excitation wavelength (m)                  */
    double lambda_sted;      /* depletion wavelength (m)                   */
    double lambda_fluo;      /* emission wavelength (m)                    */
    double sigma_abs_ex;     /* absorption cross-section at lambda_ex (m2) */
    double sigma_abs_sted;   /* ... at lambda_sted (anti-Stokes excitation) */
    double sigma_ste;        /* stimulated-emission cross-section (m2)     */
    double qy;               /* quantum yield                              */
    double tau;              /* fluorescence lifetime (s)                  */
    double tau_vib;          /* vibrational relaxation (s)                 */
    double tau_tri;          /* triplet lifetime (s)                       */
    double k0, k1, b;        /* bleaching law k0*I + k1*I^b                */
    double triplet_dynamics_frac;
    double sted_tau;         /* STED pulse width (s)                       */
    double sted_rate;        /* laser repetition rate (Hz)                 */
    int32_t anti_stoke;      /* DonutBeam.anti_stoke                       */
    int32_t _pad;
} StedFluoParams;

/* Detector parameters (pysted.base.Detector; gym_sted/defaults.py:21). */
typedef struct StedDetectorParams {
    double pcef;             /* photon collection efficiency  */
    double pdef;             /* photon detection efficiency   */
    double background;       /* background counts per second  */
} StedDetectorParams;

/* Per-env tables produced by sted_make_effective, consumed by sted_acquire*.
 * float32 [B][STED_TABLES][K][KP], KP = STED_KPITCH(K); columns >= K are zero. */
#define STED_TAB_E   0   /* effective PSF E[u][v] (W per molecule) — Microscope.get_effective */
#define STED_TAB_ET  1   /* E[u][v]*exp(-sum_{v'>v} A[u][v']): in-row expected-survival weights */
#define STED_TAB_CH  2   /* suffix hazard CH[u][v] = sum_{v'>=v} A[u][v'], A = k_bleach*pdt    */
#define STED_TAB_RS  3   /* exp(-CH[u][v]): survival of a whole in-row pass starting at tap v     */
#define STED_TAB_E2  4   /* slots 4..7: float32 [K+1][KP][2] = (E[u][v], E[u-1][v]), zero outside: the packed operand
                          * of the gather kernel's FFMA2 (two output rows per instruction)                      */
#define STED_TABLES  8

const char* sted_version(void);
const char* sted_last_error(void);
/* Number of visible CUDA devices (0 when none); never fails. */
int sted_device_count(void);

/* get_effective + get_k_bleach for B environments at once.
 *   psf_cache_dev : float64 [3][K][K]  = Microscope.cache(px): i_ex, i_sted (instantaneous,
 *                   per watt of average power), psf_det
 *   fluo_dev      : StedFluoParams [B]
 *   p_ex/p_sted/pdt_dev : float64 [B]  imaging parameters of each environment
 *   tables_dev    : float32 [B][STED_TABLES][K][KP]  (output)
 *   kbleach_dev   : float64 [B][K][K] bleaching-rate map k_b (1/s), or NULL (output)
 */
int sted_make_effective(const double* psf_cache_dev, const StedFluoParams* fluo_dev,
                        const double* p_ex_dev, const double* p_sted_dev, const double* pdt_dev,
                        int B, int K, float* tables_dev, double* kbleach_dev, void* stream);

/* One raster-scan acquisition for B environments (Microscope.get_signal_and_bleach).
 *   dmap_in_dev   : float32 [B][Hp][Wpitch] molecule maps
 *   dmap_out_dev  : float32 [B][Hp][Wpitch] bleached maps (required, and != dmap_in_dev, when
 *                   STED_BLEACH is set; ignored otherwise)
 *   tables_dev    : from sted_make_effective
 *   pdt_dev       : float64 [B] pixel dwell times
 *   det           : detector parameters (host struct, read during the call)
 *   lambda_fluo   : emission wavelength for the photon conversion (m)
 *   seed, offset  : Philox key / acquisition counter; (seed, offset, env, pixel) fixes every draw,
 *                   env ids are env_base .. env_base+B-1 so results do not depend on sharding
 *   intensity_dev : float32 [B][H][W] detected power per pixel (W), or NULL
 *   signal_dev    : float32 [B][H][W] photon counts (integer-valued when STED_SAMPLE_PHOTONS)
 */
int sted_acquire(const float* dmap_in_dev, float* dmap_out_dev, const float* tables_dev,
                 const double* pdt_dev, int B, int H, int W, int K, uint32_t flags,
                 const StedDetectorParams* det, double lambda_fluo,
                 uint64_t seed, uint64_t offset, int64_t env_base,
                 float* intensity_dev, float* signal_dev,
                 void* workspace_dev, uint64_t workspace_bytes, void* stream);

/* The general scan: imaging parameters that change from dwell to dwell (pySTED accepts (H, W) arrays for pdt, p_ex and
 * p_sted; the timed environments pass `pdt * numpy.ones(shape)` with zeros where the experiment's clock runs out,
 * gym_sted/envs/timed_sted_env.py:141-148,199).  Dwells are visited in the reference's row-major order, each one
 * gathering and bleaching its K x K window before the next starts; one CTA per environment works on a K-row band of
 * the map held in shared memory.  Scalar parameters belong to sted_acquire, which is an order of magnitude faster.
 *   tables_dev   : float32 [B][n_tables][STED_TABLES][K][KP] and
 *   kbleach_dev  : float64 [B][n_tables][K][K]: sted_make_effective's outputs for B * n_tables parameter sets, built at
 *                  the REFERENCE excitation power and dwell time of each environment (only STED_TAB_E is read);
 *                  kbleach_dev may be NULL without STED_BLEACH
 *   pdt_dev      : float64 [B] reference dwell time
 *   w_dwell_dev  : float32 [B][H][W] dwell time of every pixel / reference dwell time, or NULL (= 1)
 *   w_ex_dev     : float32 [B][H][W] excitation power of every pixel / reference power, or NULL (= 1): E and k_b are
 *                  linear in p_ex (without DonutBeam.anti_stoke)
 *   table_id_dev : uint8 [B][H][W] which of the n_tables (<= 256) parameter sets — distinct p_sted values — a dwell
 *                  uses, or NULL (= 0)
 *   others as sted_acquire; the detector uses each pixel's own dwell time.  Stochastic bleaching draws one binomial per
 *   (dwell, window pixel) — pySTED's own chain — keyed by (seed, offset, env, pixel, tap); no workspace is needed.
 * Limit: (W + K - 1) * K map values must fit in shared memory (W + K - 1 <= 600 at K = 59). */
int sted_acquire_dwellmaps(const float* dmap_in_dev, float* dmap_out_dev, const float* tables_dev,
                           const double* kbleach_dev, const double* pdt_dev, const float* w_dwell_dev,
                           const float* w_ex_dev, const uint8_t* table_id_dev, int n_tables,
                           int B, int H, int W, int K, uint32_t flags, const StedDetectorParams* det,
                           double lambda_fluo, uint64_t seed, uint64_t offset, int64_t env_base,
                           float* intensity_dev, float* signal_dev, void* stream);

/* Scratch for STED_BLEACH|STED_SAMPLE_BLEACH: the deaths of the whole scan are drawn before it, staged and
 * bucketed by scan row (12 bytes per event).  max_events must cover the total molecule count of the B maps
 * (every 4096-pixel chunk reserves staging room for all of its molecules).  workspace_dev must be 8-byte
 * aligned.  Other modes accept workspace_dev = NULL.
 * Limits of sted_acquire: B <= 65535 environments per call, H <= 65535, W + K - 1 <= 4095, fewer than 32768
 * molecules per pixel. */
uint64_t sted_bleach_workspace_bytes(int B, int H, uint64_t max_events);
/* First half of a stochastic scan on its own: draws every death of the scan (Philox, keyed as in sted_acquire) and
 * buckets the events by scan row in workspace_dev.  It depends on the maps and on the STED acquisition's tables only,
 * not on the images acquired in between, so a caller can queue it on a second stream — e.g. under the confocal
 * acquisition that precedes the STED one (gym_sted/envs/mosted_env.py:222-227) — and then call sted_acquire with
 * STED_PRESAMPLED on the same arguments.  Arguments as sted_acquire. */
int sted_bleach_presample(const float* dmap_in_dev, const float* tables_dev, const double* pdt_dev,
                          int B, int H, int W, int K, uint32_t flags, uint64_t seed, uint64_t offset,
                          int64_t env_base, void* workspace_dev, uint64_t workspace_bytes, void* stream);
/* Number of deaths the last stochastic scan could not store (0 unless max_events was too small;
 * a non-zero value means that scan's outputs are invalid).  Synchronises `stream`. */
int sted_bleach_dropped_events(const void* workspace_dev, int* dropped_host, void* stream);

/* Detector.get_signal on its own (gym_sted/utils.py:654-657): converts detected power (W), as left
 * in signal_dev by sted_acquire(..., STED_NO_DETECT), to photon counts in place:
 *   mean = floor(I / (h c / lambda_fluo)) * pcef * pdef * pdt + background * pdt,
 *   Poisson(mean) when STED_SAMPLE_PHOTONS is set.  Same (seed, offset, env, pixel) keying. */
int sted_detect(float* signal_dev, const double* pdt_dev, int B, int H, int W, uint32_t flags,
                const StedDetectorParams* det, double lambda_fluo, uint64_t seed, uint64_t offset,
                int64_t env_base, void* stream);

/* Same, from HOST buffers: pads, uploads, runs sted_make_effective + sted_acquire, downloads.
 *   dmap_host     : float32 (uint16 with STED_HOST_U16) [B][H][W] molecule maps of the region of interest
 *                   (un-padded); overwritten with the bleached maps when STED_BLEACH is set
 *   psf_cache_host: float64 [3][K][K]
 *   fluo_host     : StedFluoParams [B];  p_ex/p_sted/pdt_host : float64 [B]
 *   intensity_host : float32 [B][H][W], may be NULL;  signal_host : float32 (uint16 with STED_HOST_U16) [B][H][W]
 */
/* sted_acquire_host keeps its device scratch and its streams per calling host thread, so calls from different threads
 * run concurrently.  A thread that carries work nobody waits for yet (e.g. the confocal image of the next structure)
 * can ask for low-priority streams before its first call: the GPU then runs it in the idle slots of the other
 * threads' kernels. */
int sted_host_thread_priority(int low);

int sted_acquire_host(void* dmap_host, const double* psf_cache_host,
                      const StedFluoParams* fluo_host, const double* p_ex_host,
                      const double* p_sted_host, const double* pdt_host,
                      int B, int H, int W, int K, uint32_t flags,
                      const StedDetectorParams* det, uint64_t seed, uint64_t offset,
                      int64_t env_base, float* intensity_host, void* signal_host);

/* ---- Sessions: a structure stays resident across the acquisitions made of it ------------------------------------
 * gym-sted images every structure several times (confocal, STED + bleaching, confocal: gym_sted/envs/mosted_env.py:
 * 222-234) on ONE pysted Datamap object that carries the bleaching (sequence_sted_env.py:86-92).  A session is that
 * object for B environments: the map is uploaded once, acquisitions advance it on the device, and every call only
 * QUEUES work (own streams; uploads, kernels and downloads of successive calls overlap, and so do the parts of
 * successive acquisitions that do not depend on each other: the tables and the death schedule of an acquisition are
 * prepared under the gather or scan of the one before it, its detector runs under the one after it).  Host buffers handed to
 * sted_session_acquire / _download must stay valid, and are only defined, after the next sted_session_sync.
 * One session belongs to one device (the one current at creation) and must be driven by one host thread at a time.
 *   psf_cache_host : float64 [3][K][K] (copied at creation);  low_priority != 0: streams of the lowest priority, for
 *                    work nobody waits for yet (e.g. the first confocal image of the NEXT structure)
 *   dmap_host      : [B][H][W] un-padded molecule maps, float32, or uint16 with STED_HOST_U16
 *   max_molecules  : upper bound of the total molecule count of the B maps (sizes the event scratch of the stochastic
 *                    scan: 12 bytes each); 0 = 4 per pixel, and sted_session_sync reports STED_EOVERFLOW if a scan
 *                    needed more
 *   flags          : as sted_acquire, plus STED_HOST_U16 for uint16 signal_host
 *   signal_host / intensity_host : [B][H][W] or NULL (result stays on the device / is not wanted)                */
typedef struct StedSession StedSession;
int sted_session_create(int B, int H, int W, int K, const double* psf_cache_host, int low_priority, StedSession** out);
int sted_session_destroy(StedSession* s);
int sted_session_set_fluorophores(StedSession* s, const StedFluoParams* fluo_host /* [B] */);
int sted_session_upload(StedSession* s, const void* dmap_host, uint32_t flags, uint64_t max_molecules);
int sted_session_acquire(StedSession* s, const double* p_ex_host, const double* p_sted_host, const double* pdt_host,
                         uint32_t flags, const StedDetectorParams* det, double lambda_fluo, uint64_t seed, uint64_t offset,
                         int64_t env_base, void* signal_host, float* intensity_host);
/* Device-resident callers (the vectorised environments keep their structures and images on the GPU):
 * sted_session_upload_device pads [B][H][W] maps that already live in device memory (float32, or uint16 with
 * STED_HOST_U16), ready once the work `stream` holds at the time of the call has run; roi_dev must stay unchanged until
 * the session has been synchronised or a stream has waited (sted_session_result) on an acquisition queued after it.
 * sted_session_result hands out the device buffers of the acquisition queued `back` calls ago (0 = the last one,
 * back < 4): float32 [B][H][W] photon counts and detected power (the latter only when that call asked for the
 * intensity); with wait != 0 `stream` is made to wait for that acquisition's detector.  A buffer is overwritten by the
 * fourth acquisition queued after its own.
 * sted_session_kernel_ms: device time (CUDA events on the session's compute stream) of the gather or scan kernel of
 * the acquisition queued `back` calls ago (back < 64); waits for it.
 * sted_session_join: `stream` waits for everything queued on the session so far (all of its streams). */
int sted_session_upload_device(StedSession* s, const void* roi_dev, uint32_t flags, uint64_t max_molecules, void* stream);
int sted_session_result(StedSession* s, int back, float** signal_dev, float** intensity_dev, void* stream, int wait);
int sted_session_kernel_ms(StedSession* s, int back, float* ms);
int sted_session_join(StedSession* s, void* stream);
/* dst takes over src's current maps (device-to-device, queued after everything already queued on both sessions).
 * gym-sted's step ends with the first confocal image of the NEXT structure (mosted_env.py:174-186), which is the
 * structure the next step images three times: a low-priority session uploads and images it while the main session
 * works, then the main session adopts it — one upload per structure.  Same B, H, W, K required. */
int sted_session_adopt(StedSession* dst, StedSession* src);
/* Ordering between sessions: the gather / scan kernels queued on s from now on start only after those queued on `other`
 * so far have finished (other's detector passes and downloads may still be running).  A low-priority session's work is
 * otherwise free to drift to the very end of a step, where a caller that synchronises every step waits for it with an
 * idle GPU; ordering the step's last acquisition after it lets its detector and download run under that acquisition. */
int sted_session_after(StedSession* s, StedSession* other);
/* Current (bleached) maps -> host, same format rules as sted_session_upload. */
int sted_session_download(StedSession* s, void* dmap_host, uint32_t flags);
/* Waits for everything queued on the session; *dropped_events (may be NULL) = deaths that did not fit the scratch. */
int sted_session_sync(StedSession* s, int64_t* dropped_events);
/* Device pointer of the current padded maps [B][Hp][Wpitch] (valid until the next bleaching acquisition / upload). */
int sted_session_device_map(StedSession* s, float** map_dev);

/* ---- Fluorophore fitting (gym_sted/utils.py:442-686, FluorescenceOptimizer.optimize; BleachSampler "uniform"/"choice",
 * utils.py:339-376): for each environment the (sigma_abs, k1, b) that give a confocal signal of sig_target photons and a
 * bleached fraction of bl_target under the stated imaging parameters, found along the path of the reference's optimiser
 * (25 rounds of two L-BFGS-B fits, forward differences of 0.01 in the scaled variables, targets blended in).  A target
 * <= 0 skips that criterion.  Detector noise is replaced by its expectation.
 *   psf_cache_dev : float64 [3][K][K] (Microscope.cache);  blob_dev: float64 [K][K], the test structure of
 *                   expected_confocal_signal (a Gaussian-filtered unit pixel scaled to a peak of 40 molecules), zero
 *                   outside the centred square of half-side blob_radius
 *   fluo_host     : start values (sigma_abs_ex, k1, b = defaults.FLUO) and the fixed parameters;  det_host: detector
 *   out_dev       : float64 [B][5] = sigma_abs (m2), k1, b, signal reached (photons), bleach reached            */
typedef struct StedFluoCriteria {
    double sig_p_ex, sig_p_sted, sig_pdt, sig_target;
    double bl_p_ex, bl_p_sted, bl_pdt, bl_target;
} StedFluoCriteria;
int sted_fluo_optimize(const double* psf_cache_dev, const double* blob_dev, int K, int blob_radius, const StedFluoParams* fluo_host,
                       const StedDetectorParams* det_host, const StedFluoCriteria* criteria_dev, int B, int iterations,
                       double* out_dev, void* stream);

/* Samplers exposed for the statistical tests (same device functions the kernels use).
 *   out_dev[i] ~ Poisson(lambda_dev[i]) / Binomial(n_dev[i], q_dev[i]); counter = i. */
int sted_sample_poisson(const float* lambda_dev, int64_t n, uint64_t seed, uint64_t offset,
                        float* out_dev, void* stream);
int sted_sample_binomial(const float* n_dev, const float* q_dev, int64_t n, uint64_t seed,
                         uint64_t offset, float* out_dev, void* stream);
/* Raw Philox4x32-10 block for known-answer tests: out_host[4]. Host-side, no device needed. */
void sted_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out_host[4]);

/* Foreground masks and the Bleach / SNR objectives of B acquisitions, bit-identical to
 *   get_foreground (skimage Otsu on integer images)        gym_sted/utils.py:16-24
 *   fg_s = otsu(sted) if any(sted) else ones; fg_s *= fg_c gym_sted/envs/mosted_env.py:237-244
 *   Bleach.evaluate, Signal_Ratio(75).evaluate             gym_sted/rewards/objectives.py:91-111
 *   conf1/sted/conf2_dev : float32 [B][H][W] photon counts (integer valued)
 *   fg_c/fg_s_dev        : uint8   [B][H][W] masks (output)
 *   objs_dev             : float64 [B][2] = {Bleach, SNR}   (SNR is 0 without STED foreground)
 *   flags_dev            : int32   [B]  bit 0: SNR < 0 (the reference returns None), bit 1: value range too
 *                          large for the histograms (outputs NaN)
 *   hist_scratch_dev     : int32 [B][2][65536] or NULL; only needed when an image spans > 8192 values */
int sted_foreground_objectives(const float* conf1_dev, const float* sted_dev, const float* conf2_dev,
                               int B, int H, int W, uint8_t* fg_c_dev, uint8_t* fg_s_dev,
                               double* objs_dev, int32_t* flags_dev, int32_t* hist_scratch_dev,
                               void* stream);

/* Bleach / SNR objectives of B acquisitions for masks the CALLER supplies - the form the reference's objective classes
 * have: Bleach.evaluate and Signal_Ratio(75).evaluate take sted_fg / confocal_fg as arguments
 * (gym_sted/rewards/objectives.py:76-111; called through MORewardCalculator.evaluate, gym_sted/rewards/rewards.py:52-67).
 * Arguments as sted_foreground_objectives, except that fg_c_dev / fg_s_dev (uint8, non-zero = foreground) are inputs
 * and fg_s is used as given (the environments multiply it by fg_c before the call, mosted_env.py:244).  Empty
 * selections give NaN like numpy's mean / percentile of nothing; no STED foreground gives SNR = 0. */
int sted_objectives_masked(const float* conf1_dev, const float* sted_dev, const float* conf2_dev,
                           int B, int H, int W, const uint8_t* fg_c_dev, const uint8_t* fg_s_dev,
                           double* objs_dev, int32_t* flags_dev, int32_t* hist_scratch_dev, void* stream);

/* Nanodomain candidates of B STED images: the integer part of NumberNanodomains.evaluate
 * (gym_sted/rewards/objectives.py:416-422): gaussian(sigma 0.5) -> > quantile 0.99 -> remove_small_objects(4,
 * 8-connected) -> label -> peak_local_max(min_distance 2, per label, no border exclusion).
 *   sted_dev    : float32 [B][H][W] photon counts
 *   weights     : host, the 3 distinct taps {w0, w1, w2} of scipy's 5-tap Gaussian kernel for sigma 0.5
 *   k0, gamma   : numpy.quantile's lower order-statistic index and interpolation weight for q = 0.99, N = H*W
 *   scratch_dev : float64 [2][B][H][W]
 *   peaks_dev   : int32 [B][256][2] (row, col) in skimage's order; npeaks_dev: int32 [B]
 *   flags_dev   : int32 [B] non-zero = a capacity of the kernel was exceeded (bit 0: more than 4096 mask pixels,
 *                 bit 1: more than 64 maxima in one component, bit 3: more than 256 peaks); the output is then
 *                 incomplete and the caller must treat it as an error */
int sted_nanodomain_candidates(const float* sted_dev, int B, int H, int W, const double weights[3],
                               int64_t k0, double gamma, double* scratch_dev, int32_t* peaks_dev,
                               int32_t* npeaks_dev, int32_t* flags_dev, void* stream);

/* Resolution objective by image decorrelation (gym_sted/rewards/objectives.py:146-335, 378-385): all 22 decorrelation
 * passes and peak searches of B images, one CTA per image.
 *   ik_dev, ikn_dev : complex128 [B][N] centred Fourier transform of the apodised image and its phase-only version,
 *                     pixels sorted by normalised radius (stable), zero where R >= 1
 *   radius_dev, radius2_dev : float64 [N] the sorted radii and their squares; n_inside = number of radii < 1
 *   r_coarse_dev    : float64 [50] = numpy.linspace(0, 1, 50)
 *   sigma0 = Hc/4, gmax_flat = max(Hc, Wc)/2 (the reference's substitutes when the first peak sits at r = 0)
 *   res_dev         : float64 [B] resolution in nm, capped at res_cap */
int sted_decorrelation_resolution(const void* ik_dev, const void* ikn_dev, int B, int N, int n_inside,
                                  const double* radius_dev, const double* radius2_dev, const double* r_coarse_dev,
                                  double sigma0, double gmax_flat, double pixelsize, double res_cap, double tol,
                                  double* res_dev, void* stream);

/* Gaussian-fit validation of nanodomain candidates (gym_sted/rewards/utils.py:6-49, called from
 * gym_sted/rewards/objectives.py:424): one Levenberg-Marquardt fit (MINPACK LMDIF restated, float64) of a 7x7
 * zero-padded window per candidate; valid = both FWHMs 2.355*w*pixelsize inside [40 nm, 200 nm].
 *   peaks_dev/npeaks_dev : as written by sted_nanodomain_candidates, max_peaks rows per image
 *   valid_dev  : uint8 [B][max_peaks]
 *   params_dev : optional float64 [B][max_peaks][5] fitted (amplitude, cx, cy, wx, wy), may be NULL */
int sted_gaussfit_validate(const float* sted_dev, int B, int H, int W, const int32_t* peaks_dev,
                           const int32_t* npeaks_dev, int max_peaks, double pixelsize, uint8_t* valid_dev,
                           double* params_dev, void* stream);

/* F1 score of the validated detections against the ground-truth nanodomains (gym_sted/rewards/objectives.py:425-430:
 * metrics.CentroidDetectionError(truth, guess, threshold, "hungarian").f1_score).  TP = size of a maximum matching of
 * the pairs closer than `threshold` (what the Hungarian solution on the priced-out cost matrix counts), FP / FN the
 * unmatched guesses / truths; F1 = 2PR/(P+R+1e-6) with P = TP/(TP+FP+1e-6), R = TP/(TP+FN+1e-6), float64.
 *   peaks_dev/npeaks_dev/valid_dev : as written by sted_nanodomain_candidates / sted_gaussfit_validate (valid_dev NULL =
 *                                    keep every candidate), max_peaks rows per image
 *   truth_dev  : int32 [B][max_truth][2] (row, col);  ntruth_dev : int32 [B]
 *   counts_dev : int32 [B][3] = TP, FP, FN;  f1_dev : float64 [B];  flags_dev : int32 [B] (NULL allowed), 1 = more than
 *                512 truths or kept guesses in the image (outputs then cover the first 512 only: treat as an error) */
int sted_f1_match(const int32_t* peaks_dev, const int32_t* npeaks_dev, const uint8_t* valid_dev, int max_peaks,
                  const int32_t* truth_dev, const int32_t* ntruth_dev, int max_truth, int B, double threshold,
                  int32_t* counts_dev, double* f1_dev, int32_t* flags_dev, void* stream);

/* The two element-wise ends of the Resolution objective's Fourier transform (gym_sted/rewards/objectives.py:150-175).
 * sted_apodize_for_fft: out = ifftshift(crop(float32(window * (image - mean) + mean))) as complex128 [B][Hc][Wc]
 *   (image_dev float32 [B][H][W] photon counts, window_dev float64 [H][W] the cosine edge window), ready for a plain
 *   forward FFT.  sted_spectrum_gather: from that FFT (complex128 [B][Hc][Wc]) the centred spectrum in radius order
 *   (order_dev int64 [Hc*Wc]: stable argsort of the normalised radii of the CENTRED spectrum), zero at R >= 1 (the
 *   first n_inside entries are inside), plus its phase-only copy Ik/|Ik| — the inputs of
 *   sted_decorrelation_resolution. */
int sted_apodize_for_fft(const float* image_dev, const double* window_dev, int B, int H, int W, int Hc, int Wc,
                         void* out_dev, void* stream);
int sted_spectrum_gather(const void* spectrum_dev, const int64_t* order_dev, int B, int Hc, int Wc, int n_inside,
                         void* ik_dev, void* ikn_dev, void* stream);

/* ---- env.step plumbing between the acquisitions (gym_sted/envs/contextual_sted_env.py:49-93, mosted_env.py:168-187) ----
 * sted_synapse_generate: the structures a ContextualMOSTED step ends on (`self.synapse_generator(rotate=True)`,
 * mosted_env.py:80-83,174; gym_sted/utils.py:26-78), for B environments on the device.  pysted.exp_data_gen.Synapse is
 * not vendored in the reference; the geometry is the synthetic spine datamap of SURVEY.md 8d: per 64 x 64 cell one
 * elliptical outline band thick_lo..thick_hi px thick holding `molecules` per pixel and n_lo..n_hi single-pixel
 * nanodomains of m_lo..m_hi molecules inside the band, at least min_dist_px apart (dart throwing in the order of a
 * random priority).  Every draw is a Philox4x32-10 word keyed by (seed, env_base + b, cell, offset) and the geometry is
 * float64 +, -, *, /, sqrt only, so oracle/synapse_oracle.py reproduces the outputs bit for bit.
 *   frames_dev : int32 [B][H][W] molecule maps;  coords_dev : int32 [B][max_truth][2] (row, col) of the nanodomains in
 *   cell order, zero filled;  counts_dev : int32 [B];  scratch_dev : sted_synapse_scratch_bytes(B, H, W) bytes. */
typedef struct StedSynapseParams {
    int32_t molecules, n_lo, n_hi, m_lo, m_hi, _pad;
    double min_dist_px, thick_lo, thick_hi;
} StedSynapseParams;
uint64_t sted_synapse_scratch_bytes(int B, int H, int W);
int sted_synapse_generate(const StedSynapseParams* prm, int B, int H, int W, uint64_t seed, uint64_t offset, int64_t env_base,
                          int32_t* frames_dev, int32_t* coords_dev, int32_t* counts_dev, int max_truth, void* scratch_dev,
                          void* stream);
/* Procedural structure maps of the preference environments (PreferenceCountRateMOSTED-*): stand-in for
 * DatamapGenerator (gym_sted/datamaps/datamap.py:68-152), whose U-Net datamaps are not shipped with the reference
 * (defaults.py:6-13).  Per environment up to 24 objects combined by maximum, cut below 0.5, times the molecule count,
 * floored, rotated by rot * 90 degrees (numpy.rot90; square maps) and flipped along rows / columns.
 *   objects_dev : float64 [B][24][4]  filament: cos, sin, offset, width of exp(-d^2 / 2 width^2) with
 *                 d = |x cos + y sin - offset + 3 sin(0.08 (y cos - x sin))|;  cluster: cy, cx, sigma, unused
 *   env_dev     : float64 [B][6] = number of objects, molecule count, rot (0..3), flip rows, flip columns, 1 = filaments
 *   maps_dev    : float32 [B][H][W] molecule counts (integer valued) */
int sted_structure_maps(const double* objects_dev, const double* env_dev, int B, int H, int W, float* maps_dev, void* stream);
/* Datamap.set_roi(i_ex, "max") (gym_sted/utils.py:179-180) for device-resident maps: roi_dev int32 or float32 [B][H][W]
 * -> padded_dev float32 [B][H+K-1][roundup4(W+K-1)] with the zero frame of K/2 pixels. */
int sted_pad_maps(const void* roi_dev, int roi_is_int32, int B, int H, int W, int K, float* padded_dev, void* stream);
/* Image half of the observation (gym_sted/envs/contextual_sted_env.py:84-93): (state, conf1, sted) float32 [B][H][W]
 * photon counts -> uint16 [B][H][W][3], cast like numpy's int64 -> uint16 (wrap modulo 2^16). */
int sted_pack_observation(const float* state_dev, const float* conf1_dev, const float* sted_dev, int B, int H, int W,
                          uint16_t* obs_dev, void* stream);

/* Test hook for the Levenberg-Marquardt solver: n windows of 49 float64 values, params_dev float64 [n][5] holds the
 * starting points on entry and the solutions on return, info_dev int32 [n] MINPACK's info code.
 * model 0: the Gaussian of sted_gaussfit_validate; model 1: a rational peak (no transcendental, bit-reproducible), both
 * with the one-thread solver; models 2 / 3: the same two through the warp-cooperative solver sted_gaussfit_validate runs. */
int sted_lmfit_test(const double* data_dev, int n, int model, double ftol, double* params_dev, int32_t* info_dev,
                    void* stream);

/* Debug: per-phase cycle sums of the scan kernel (non-zero only in SWEEP_PROBE=2 builds). */
int sted_debug_phase_cycles(unsigned long long out_host[8], int reset);

/* FP32-FMA peak microbenchmark (roofline denominator): returns achieved TFLOP/s in *tflops. */
int sted_fma_peak(double* tflops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STED_B200_H */
