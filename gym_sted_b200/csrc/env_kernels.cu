// gym_sted_b200 — the per-step pieces of env.step that sit between the acquisitions.
//
//  * synapse_cell_kernel / synapse_compact_kernel: the structure every ContextualMOSTED step ends on
//    (gym_sted/envs/mosted_env.py:168-187 `_update_datamap`: `self.synapse = self.synapse_generator(rotate=True)`,
//    generator configured at mosted_env.py:80-83 — 5 molecules per structure pixel, 3-15 nanodomains of 100-175 molecules,
//    at least 100 nm apart, valid_thickness (3, 10); gym_sted/utils.py:26-78).  pysted.exp_data_gen.Synapse is not
//    vendored in the reference, so the geometry is the synthetic spine datamap SURVEY.md §8d specifies: per 64 x 64
//    cell one elliptical outline band 3-10 px thick, single-pixel nanodomains inside the band, >= min_dist apart.
//    Every draw is a Philox4x32-10 word keyed by (seed, global env id, cell, structure number) and all geometry is
//    float64 +, -, *, /, sqrt (this file is compiled with -fmad=false), so oracle/synapse_oracle.py reproduces frames
//    and coordinates bit for bit.
//  * pad_maps_kernel: Datamap.set_roi(i_ex, "max") for device-resident int32 / float32 maps (gym_sted/utils.py:179-180).
//  * pack_observation_kernel: the image half of the observation, (state, conf1, sted) stacked on the last axis and cast
//    to uint16 the way numpy casts int64 (wrap modulo 2^16; gym_sted/envs/contextual_sted_env.py:84-93).
#include <cstdint>

#include "common.cuh"
#include "philox.cuh"

namespace {

constexpr int CELL = 64, CELL_PIX = CELL * CELL, SYN_THREADS = 256, SYN_MAXN = 32;
constexpr uint32_t RNG_SYNAPSE = 7u;      // purpose field (top 4 bits of counter word 2); 1-6 are taken by the acquisition

__device__ __forceinline__ double u01(uint32_t w) { return ((double)w + 0.5) * 2.3283064365386963e-10; }   // (w + 0.5) / 2^32
__device__ __forceinline__ uint32_t pick(uint32_t w, int lo, int hi) {        // integer in [lo, hi] (multiply-shift)
    return (uint32_t)lo + (uint32_t)(((uint64_t)w * (uint64_t)(hi - lo + 1)) >> 32);
}

struct CellRng {
    uint32_t k0, k1, env, z, off;
    __device__ Philox4 block(uint32_t idx) const {
        Philox4 c;
        c.x = idx; c.y = env; c.z = z; c.w = off;
        return philox4x32_10(c, k0, k1);
    }
};

// grid (cells_x * cells_y, B), 256 threads.  Philox blocks of a cell: 0 centre / half axes, 1 thickness / count,
// 2.. direction (rejection), 32.. molecules of the accepted nanodomains, 64.. priorities (block 64 + pix / 4, word pix % 4).
__global__ void __launch_bounds__(SYN_THREADS) synapse_cell_kernel(
    const StedSynapseParams prm, const int B, const int H, const int W, const int cells_x, const uint64_t seed,
    const uint32_t offset, const uint32_t env_base, int32_t* __restrict__ frames, int32_t* __restrict__ cell_coords,
    int32_t* __restrict__ cell_counts) {
    __shared__ unsigned long long keys[CELL_PIX];
    __shared__ uint8_t band[CELL_PIX];
    __shared__ double geo[8];
    __shared__ int s_n, s_nband, s_nacc, s_acc[SYN_MAXN], s_mol[SYN_MAXN];
    const int cell = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int cy0 = (cell / cells_x) * CELL, cx0 = (cell % cells_x) * CELL;
    CellRng rng{(uint32_t)seed, (uint32_t)(seed >> 32), env_base + (uint32_t)b, (RNG_SYNAPSE << 28) | (uint32_t)cell, offset};

    if (tid == 0) {
        const Philox4 g0 = rng.block(0), g1 = rng.block(1);
        geo[0] = CELL / 2 + (12.0 * u01(g0.x) - 6.0);        // centre row
        geo[1] = CELL / 2 + (12.0 * u01(g0.y) - 6.0);        // centre column
        geo[2] = 12.0 + 8.0 * u01(g0.z);                     // half axis along the direction
        geo[3] = 9.0 + 7.0 * u01(g0.w);                      // half axis across
        geo[4] = prm.thick_lo + (prm.thick_hi - prm.thick_lo) * u01(g1.x);
        s_n = (int)pick(g1.y, prm.n_lo, prm.n_hi);
        double c = 1.0, s = 0.0;                             // uniform direction without trigonometry: a point of the unit disc
        for (uint32_t i = 2; i < 32; ++i) {
            const Philox4 d = rng.block(i);
            const double x0 = 2.0 * u01(d.x) - 1.0, y0 = 2.0 * u01(d.y) - 1.0, x1 = 2.0 * u01(d.z) - 1.0, y1 = 2.0 * u01(d.w) - 1.0;
            const double r0 = x0 * x0 + y0 * y0, r1 = x1 * x1 + y1 * y1;
            if (r0 <= 1.0 && r0 > 1e-4) { const double r = sqrt(r0); c = x0 / r; s = y0 / r; break; }
            if (r1 <= 1.0 && r1 > 1e-4) { const double r = sqrt(r1); c = x1 / r; s = y1 / r; break; }
        }
        geo[5] = c; geo[6] = s;
        s_nband = 0; s_nacc = 0;
    }
    __syncthreads();
    const double cyc = geo[0], cxc = geo[1], ax = geo[2], bx = geo[3], thick = geo[4], c = geo[5], s = geo[6];
    const double mean_r = 0.5 * (ax + bx), half = thick / 2.0;
    int mine = 0;
    for (int blk = tid; blk < CELL_PIX / 4; blk += SYN_THREADS) {
        const Philox4 pr = rng.block(64u + (uint32_t)blk);
        const uint32_t w[4] = {pr.x, pr.y, pr.z, pr.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int pix = 4 * blk + j;
            const double dy = (double)(pix / CELL) - cyc, dx = (double)(pix % CELL) - cxc;
            const double ru = (dx * c + dy * s) / ax, rv = (dy * c - dx * s) / bx;
            const double rad = sqrt(ru * ru + rv * rv);
            const bool in = fabs(rad - 1.0) * mean_r <= half;
            band[pix] = in;
            keys[pix] = in ? ((1ull << 44) | ((unsigned long long)w[j] << 12) | (unsigned long long)(CELL_PIX - 1 - pix)) : 0ull;
            mine += in;
        }
    }
    if (mine) atomicAdd(&s_nband, mine);
    __syncthreads();
    // bitonic sort, descending: band pixels by (priority, lower pixel index first) in front, everything else (0) behind
    for (int k = 2; k <= CELL_PIX; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < CELL_PIX; i += SYN_THREADS) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long a = keys[i], d = keys[l];
                    const bool desc = (i & k) == 0;
                    if (desc ? a < d : a > d) { keys[i] = d; keys[l] = a; }
                }
            }
            __syncthreads();
        }
    }
    // dart throwing in priority order (warp 0): a candidate is accepted when it is >= min_dist from every accepted one
    if (tid < 32) {
        const int want = min(s_n, SYN_MAXN), nband = s_nband;
        const double min_d2 = prm.min_dist_px * prm.min_dist_px;
        int nacc = 0, my = 0, mx = 0;                    // lane l holds accepted point l
        for (int i = 0; i < nband && nacc < want; ++i) {
            const int pix = CELL_PIX - 1 - (int)(keys[i] & (CELL_PIX - 1));
            const int y = pix / CELL, x = pix % CELL;
            const int dy = y - my, dx = x - mx;
            const bool clash = tid < nacc && (double)(dy * dy + dx * dx) < min_d2;
            if (__ballot_sync(0xffffffffu, clash) == 0u) {
                if (tid == nacc) { my = y; mx = x; s_acc[nacc] = pix; }
                ++nacc;
            }
        }
        if (tid < nacc) {
            const Philox4 m = rng.block(32u + (uint32_t)(tid >> 2));
            const uint32_t w = (tid & 3) == 0 ? m.x : (tid & 3) == 1 ? m.y : (tid & 3) == 2 ? m.z : m.w;
            s_mol[tid] = (int)pick(w, prm.m_lo, prm.m_hi);
        }
        if (tid == 0) s_nacc = nacc;
    }
    __syncthreads();
    const int nacc = s_nacc;
    int32_t* fr = frames + (size_t)b * H * W;
    for (int pix = tid; pix < CELL_PIX; pix += SYN_THREADS) {
        const int y = cy0 + pix / CELL, x = cx0 + pix % CELL;
        if (y < H && x < W) fr[(size_t)y * W + x] = band[pix] ? prm.molecules : 0;
    }
    __syncthreads();
    const size_t slot = ((size_t)b * gridDim.x + cell) * SYN_MAXN;
    if (tid < nacc) {
        const int pix = s_acc[tid];
        const int y = cy0 + pix / CELL, x = cx0 + pix % CELL;
        if (y < H && x < W) fr[(size_t)y * W + x] = s_mol[tid];       // the nanodomain replaces the structure's count
        cell_coords[(slot + tid) * 2] = y;
        cell_coords[(slot + tid) * 2 + 1] = x;
    }
    if (tid == 0) cell_counts[(size_t)b * gridDim.x + cell] = nacc;
}

// one warp per environment: the cells' coordinates in cell order, clipped to the image
__global__ void synapse_compact_kernel(const int B, const int H, const int W, const int ncells, const int max_truth,
                                       const int32_t* __restrict__ cell_coords, const int32_t* __restrict__ cell_counts,
                                       int32_t* __restrict__ coords, int32_t* __restrict__ counts) {
    const int b = blockIdx.x, lane = threadIdx.x;
    int total = 0;
    for (int cell = 0; cell < ncells; ++cell) {
        const int n = cell_counts[(size_t)b * ncells + cell];
        const size_t slot = ((size_t)b * ncells + cell) * SYN_MAXN;
        int y = 0, x = 0;
        bool keep = false;
        if (lane < n) {
            y = cell_coords[(slot + lane) * 2]; x = cell_coords[(slot + lane) * 2 + 1];
            keep = y < H && x < W;
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        const int at = total + __popc(m & ((1u << lane) - 1u));
        if (keep && at < max_truth) { coords[((size_t)b * max_truth + at) * 2] = y; coords[((size_t)b * max_truth + at) * 2 + 1] = x; }
        total += __popc(m);
    }
    for (int i = min(total, max_truth) + lane; i < max_truth; i += 32) { coords[((size_t)b * max_truth + i) * 2] = 0; coords[((size_t)b * max_truth + i) * 2 + 1] = 0; }
    if (lane == 0) counts[b] = min(total, max_truth);
}

// Procedural structure maps of the preference environments (stand-in for the U-Net datamaps the reference crops and
// that are not shipped, gym_sted/datamaps/datamap.py:86-152, defaults.py:6-13): per environment up to 24 objects -
// sinuous filaments exp(-d^2 / 2 w^2) or round clusters exp(-r^2 / 2 s^2) - combined by maximum, cut below 0.5, scaled
// by the molecule count and floored, then rotated by k * 90 degrees and flipped.  One thread per OUTPUT pixel: it maps its
// coordinates back through the flips and the rotation and evaluates the analytic image there (float64).
constexpr int SM_MAXOBJ = 24;
__global__ void __launch_bounds__(256) structure_maps_kernel(const double* __restrict__ obj, const double* __restrict__ env, int H, int W,
                                                             float* __restrict__ out) {
    __shared__ double so[SM_MAXOBJ * 4];
    __shared__ double se[6];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < SM_MAXOBJ * 4; i += blockDim.x) so[i] = obj[(size_t)b * SM_MAXOBJ * 4 + i];
    if (threadIdx.x < 6) se[threadIdx.x] = env[(size_t)b * 6 + threadIdx.x];
    __syncthreads();
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= H * W) return;
    int i = pix / W, j = pix % W;
    const int nobj = (int)se[0], rot = (int)se[2];
    const bool filament = se[5] != 0.0;
    if (se[4] != 0.0) j = W - 1 - j;                       // undo flip along the columns
    if (se[3] != 0.0) i = H - 1 - i;                       // undo flip along the rows
    int y = i, x = j;                                      // undo numpy.rot90(img, rot) (square maps only: rot = 0 otherwise)
    if (rot == 1) { y = j; x = W - 1 - i; }
    else if (rot == 2) { y = H - 1 - i; x = W - 1 - j; }
    else if (rot == 3) { y = H - 1 - j; x = i; }
    const double yy = (double)y, xx = (double)x;
    double img = 0.0;
    for (int k = 0; k < nobj; ++k) {
        const double a0 = so[4 * k], a1 = so[4 * k + 1], a2 = so[4 * k + 2], a3 = so[4 * k + 3];
        double v;
        if (filament) {                                    // a0 = cos, a1 = sin, a2 = offset, a3 = width
            const double d = fabs(xx * a0 + yy * a1 - a2 + 3.0 * sin(0.08 * (yy * a0 - xx * a1)));
            v = exp(-0.5 * ((d / a3) * (d / a3)));
        } else {                                           // a0 = cy, a1 = cx, a2 = sigma
            v = exp(-0.5 * ((yy - a0) * (yy - a0) + (xx - a1) * (xx - a1)) / (a2 * a2));
        }
        img = fmax(img, v);
    }
    if (img < 0.5) img = 0.0;
    out[(size_t)b * H * W + pix] = (float)floor(img * se[1]);
}

template <class T>
__global__ void pad_maps_kernel(const T* __restrict__ roi, float* __restrict__ padded, int B, int H, int W, int p, int Hp, int Wpitch) {
    const size_t n = (size_t)B * Hp * Wpitch;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % Wpitch);
        const int y = (int)((i / Wpitch) % Hp);
        const int b = (int)(i / ((size_t)Wpitch * Hp));
        const int r = y - p, c = x - p;
        padded[i] = (r >= 0 && r < H && c >= 0 && c < W) ? (float)roi[((size_t)b * H + r) * W + c] : 0.f;
    }
}

__global__ void pack_observation_kernel(const float* __restrict__ c0, const float* __restrict__ c1, const float* __restrict__ c2,
                                        size_t n, uint16_t* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        // float -> int64 (truncation, like ndarray.astype) -> uint16 (wrap)
        out[3 * i] = (uint16_t)((long long)c0[i] & 0xFFFF);
        out[3 * i + 1] = (uint16_t)((long long)c1[i] & 0xFFFF);
        out[3 * i + 2] = (uint16_t)((long long)c2[i] & 0xFFFF);
    }
}

}  // namespace

extern "C" {

uint64_t sted_synapse_scratch_bytes(int B, int H, int W) {
    const uint64_t cells = (uint64_t)((H + CELL - 1) / CELL) * ((W + CELL - 1) / CELL);
    return (uint64_t)B * cells * (SYN_MAXN * 2 + 1) * sizeof(int32_t);
}

int sted_synapse_generate(const StedSynapseParams* prm, int B, int H, int W, uint64_t seed, uint64_t offset, int64_t env_base,
                          int32_t* frames_dev, int32_t* coords_dev, int32_t* counts_dev, int max_truth, void* scratch_dev,
                          void* stream) {
    int rc = check_device();
    if (rc) return rc;
    if (!prm || !frames_dev || !coords_dev || !counts_dev || !scratch_dev) return fail(STED_EINVAL, "null pointer argument");
    if (B <= 0 || H <= 0 || W <= 0 || max_truth <= 0) return fail(STED_EINVAL, "B, H, W and max_truth must be positive");
    if (B > 65535) return fail(STED_EUNSUPPORTED, "B=%d exceeds 65535 environments per call", B);
    if (prm->n_lo < 0 || prm->n_hi < prm->n_lo || prm->n_hi > SYN_MAXN)
        return fail(STED_EINVAL, "nanodomains per cell must satisfy 0 <= lo <= hi <= %d", SYN_MAXN);
    if (prm->m_hi < prm->m_lo || prm->thick_hi < prm->thick_lo || prm->min_dist_px < 0)
        return fail(STED_EINVAL, "invalid synapse parameter range");
    const int cx = (W + CELL - 1) / CELL, cy = (H + CELL - 1) / CELL, ncells = cx * cy;
    if (ncells > 4095) return fail(STED_EUNSUPPORTED, "%d cells exceed the 4095 cells of the counter layout", ncells);
    int32_t* cell_coords = (int32_t*)scratch_dev;
    int32_t* cell_counts = cell_coords + (size_t)B * ncells * SYN_MAXN * 2;
    cudaStream_t st = (cudaStream_t)stream;
    synapse_cell_kernel<<<dim3(ncells, B), SYN_THREADS, 0, st>>>(*prm, B, H, W, cx, seed, (uint32_t)offset, (uint32_t)env_base,
                                                                 frames_dev, cell_coords, cell_counts);
    CUDA_TRY(cudaGetLastError());
    synapse_compact_kernel<<<B, 32, 0, st>>>(B, H, W, ncells, max_truth, cell_coords, cell_counts, coords_dev, counts_dev);
    CUDA_TRY(cudaGetLastError());
    return STED_OK;
}

int sted_structure_maps(const double* objects_dev, const double* env_dev, int B, int H, int W, float* maps_dev, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    if (!objects_dev || !env_dev || !maps_dev) return fail(STED_EINVAL, "null pointer argument");
    if (B <= 0 || B > 65535 || H <= 0 || W <= 0) return fail(STED_EINVAL, "B (<= 65535), H, W must be positive");
    structure_maps_kernel<<<dim3((H * W + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(objects_dev, env_dev, H, W, maps_dev);
    CUDA_TRY(cudaGetLastError());
    return STED_OK;
}

int sted_pad_maps(const void* roi_dev, int roi_is_int32, int B, int H, int W, int K, float* padded_dev, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    if (!roi_dev || !padded_dev) return fail(STED_EINVAL, "null pointer argument");
    if (B <= 0 || H <= 0 || W <= 0 || K <= 0 || !(K & 1)) return fail(STED_EINVAL, "B, H, W must be positive and K odd");
    const int Hp = H + K - 1, Wpitch = (W + K - 1 + 3) & ~3;
    const size_t n = (size_t)B * Hp * Wpitch;
    const int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
    if (roi_is_int32) pad_maps_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const int32_t*)roi_dev, padded_dev, B, H, W, K / 2, Hp, Wpitch);
    else pad_maps_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)roi_dev, padded_dev, B, H, W, K / 2, Hp, Wpitch);
    CUDA_TRY(cudaGetLastError());
    return STED_OK;
}

int sted_pack_observation(const float* state_dev, const float* conf1_dev, const float* sted_dev, int B, int H, int W,
                          uint16_t* obs_dev, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    if (!state_dev || !conf1_dev || !sted_dev || !obs_dev) return fail(STED_EINVAL, "null pointer argument");
    if (B <= 0 || H <= 0 || W <= 0) return fail(STED_EINVAL, "B, H, W must be positive");
    const size_t n = (size_t)B * H * W;
    const int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
    pack_observation_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(state_dev, conf1_dev, sted_dev, n, obs_dev);
    CUDA_TRY(cudaGetLastError());
    return STED_OK;
}

}  // extern "C"
