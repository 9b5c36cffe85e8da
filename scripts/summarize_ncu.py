"""Extracts the judged metrics of an `ncu --set full` report into CSV + JSON under profiles/.

    python scripts/summarize_ncu.py gpurun_out/r01_full.ncu-rep profiles/r01_ncu_summary
"""
import csv
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    table = []
    for r in rows[2:]:
        rec = {"kernel": r[idx["Kernel Name"]]}
        for m in METRICS:
            if m in idx:
                rec[m] = r[idx[m]].replace(",", "")
                rec[m + ".unit"] = units[idx[m]]
        stalls = sorted(((float(r[i].replace(",", "")), h) for i, h in enumerate(hdr)
                         if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")
                         and r[i] not in ("", "n/a")), reverse=True)[:6]
        rec["top_stalls_per_issue"] = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(v, 2)
                                       for v, h in stalls}
        table.append(rec)
    with open(out + ".json", "w") as fh:
        json.dump(table, fh, indent=1)
    keys = ["kernel"] + [m for m in METRICS if m in idx]
    with open(out + ".csv", "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(keys + ["top_stalls_per_issue"])
        for rec in table:
            w.writerow([rec.get(k, "") for k in keys] + [json.dumps(rec["top_stalls_per_issue"])])
    print(f"{len(table)} launches -> {out}.csv / .json")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
