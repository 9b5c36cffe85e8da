"""Writes profiles/rNN_fp32_peak.json — the roofline denominator of the FP32-bound kernels with its clock record — from
a bench.py JSON line measured on the GPU box (MEASURED_PEAKS.json, driver-written, holds HBM and bf16 figures only).

    python scripts/record_fp32_peak.py gpurun_out/full_bench.json profiles/r02_fp32_peak.json
"""
import json
import sys

line = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
clk = line["clocks"]
sms, lanes = 148, 128
rec = {
    "fp32_tflops_measured": line["roofline"]["peak"],
    "how": "sted_fma_peak (gym_sted_b200/csrc/sted_kernels.cu: fma_peak_kernel, 1184 CTAs x 256 threads, 16 independent FFMA "
           "chains per thread, best of 5 launches, CUDA events), run inside the same bench.py process right after the timed region",
    "fp32_tflops_theoretical": sms * lanes * 2 * clk["sm_mhz"] * 1e6 / 1e12,
    "theoretical_formula": f"{sms} SMs x {lanes} FP32 lanes x 2 flop x {clk['sm_mhz']} MHz",
    "clocks_during_timed_region": clk,
    "bench_value_acq_per_s": line["value"],
    "gpu": "NVIDIA B200 (sm_100a)",
}
rec["measured_over_theoretical"] = rec["fp32_tflops_measured"] / rec["fp32_tflops_theoretical"]
json.dump(rec, open(sys.argv[2], "w"), indent=1)
print(json.dumps(rec))
