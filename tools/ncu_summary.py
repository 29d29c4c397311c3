#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into profiles/<name>.csv and .md.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_name "free-text note"
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def main():
    rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    cols = [m for m in METRICS if m in hdr]
    with open(out + ".csv", "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["kernel"] + cols)
        w.writerow(["unit"] + [units[hdr.index(c)] for c in cols])
        for d in data:
            w.writerow([d[ki]] + [d[hdr.index(c)] for c in cols])
    with open(out + ".md", "w") as fh:
        fh.write(f"# ncu summary: {rep}\n\n{note}\n\n")
        fh.write("| metric | unit | " + " | ".join(d[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:28] for d in data) + " |\n")
        fh.write("|---|---|" + "---|" * len(data) + "\n")
        for c in cols:
            i = hdr.index(c)
            fh.write(f"| {c} | {units[i]} | " + " | ".join(d[i][:14] for d in data) + " |\n")
    print("wrote", out + ".csv", out + ".md")


if __name__ == "__main__":
    main()
