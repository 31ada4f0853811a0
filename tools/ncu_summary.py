#!/usr/bin/env python3
"""Summarise an .ncu-rep (brought back in gpurun_out/) into a small CSV + markdown table under profiles/.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_xxx
"""
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = sorted(h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"))
    cols = [k for k in KEEP if k in idx] + stall
    with open(out + ".csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + cols)
        w.writerow(["unit"] + [units[idx[c]] for c in cols])
        for d in data:
            w.writerow([d[idx["Kernel Name"]].split("(")[0]] + [d[idx[c]] for c in cols])
    with open(out + ".md", "w") as f:
        f.write("| metric | " + " | ".join(d[idx["Kernel Name"]].split("(")[0] for d in data) + " |\n")
        f.write("|---|" + "---|" * len(data) + "\n")
        for c in cols:
            name = c.replace("smsp__average_warps_issue_stalled_", "stall:").replace("_per_issue_active.ratio", "")
            f.write("| %s [%s] | " % (name, units[idx[c]]) + " | ".join(d[idx[c]] for d in data) + " |\n")
    print("wrote", out + ".csv", out + ".md")


if __name__ == "__main__":
    main()
