// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_peak scripts/ffma2_peak.cu   (B200: scalar FFMA 71.8, FFMA2 74.1 TFLOP/s)
// FFMA2 (fma.rn.f32x2) throughput probe for sm_100a: packed vs scalar, with and without broadcast operands.
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float s) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f;
    if (MODE == 0) {            // scalar FFMA, 16 accumulators
        float acc[16];
        for (int i = 0; i < 16; ++i) acc[i] = i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], s, a0);
        }
        float t = 0; for (int i = 0; i < 16; ++i) t += acc[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = t;
    } else if (MODE == 1) {     // FFMA2, 8 packed accumulators (=16 fp32), operands fixed pairs
        u64 acc[8];
        for (int i = 0; i < 8; ++i) acc[i] = pack((float)i, (float)i + 0.5f);
        const u64 sp = pack(s, s), ap = pack(a0, a1);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fma2(acc[i], sp, ap);
        }
        float t = 0; for (int i = 0; i < 8; ++i) { float lo, hi; unpack(acc[i], lo, hi); t += lo + hi; }
        out[blockIdx.x * blockDim.x + threadIdx.x] = t;
    } else {                    // FFMA2 with a broadcast operand rebuilt from a scalar each time (gather-like)
        u64 acc[8];
        for (int i = 0; i < 8; ++i) acc[i] = pack((float)i, (float)i + 0.5f);
        float w[4] = {s, s * 1.1f, s * 1.2f, s * 1.3f};
        const u64 e01 = pack(a0, a1), e23 = pack(a1, a0);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const u64 wb = pack(w[k], w[k]);
                acc[2 * k] = fma2(e01, wb, acc[2 * k]);
                acc[2 * k + 1] = fma2(e23, wb, acc[2 * k + 1]);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) w[k] += 1e-7f;     // keep w live and changing (1 FADD per 2 FFMA2)
        }
        float t = 0; for (int i = 0; i < 8; ++i) { float lo, hi; unpack(acc[i], lo, hi); t += lo + hi; }
        out[blockIdx.x * blockDim.x + threadIdx.x] = t;
    }
}

template <int MODE> void run(const char* name, double flops_per_iter_thread) {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    const int iters = 20000;
    probe<MODE><<<148 * 8, 256>>>(d, 100, 0.999f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<148 * 8, 256>>>(d, iters, 0.999f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-40s %.3f ms  %.1f TFLOP/s\n", name, ms, flops_per_iter_thread * iters * 148.0 * 8 * 256 / (ms * 1e-3) * 1e-12);
    cudaFree(d);
}
int main() {
    run<0>("scalar FFMA x16", 32);
    run<1>("FFMA2 x8 (fixed operands)", 32);
    run<2>("FFMA2 x8 (broadcast pack per use)", 32);
    return 0;
}
