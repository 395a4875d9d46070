#!/usr/bin/env python
"""Extracts the metrics DESIGN.md quotes from an .ncu-rep (read here, no GPU needed):
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.md"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_adu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tc.sum.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# ncu summary of `%s`\n" % rep.split("/")[-1])
    print("Captured with `ncu --set full --clock-control none --import-source on` under gpurun (B200, sm_100a).\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("## %s\n" % name)
        print("| metric | unit | value |\n|---|---|---|")
        vals = {}
        for i, h in enumerate(hdr):
            if h in KEYS:
                vals[h] = r[i]
                print("| %s | %s | %s |" % (h, units[i], r[i]))
        try:
            inst = float(vals["smsp__inst_executed.sum"])
            cyc = float(vals["smsp__cycles_active.avg"])
            fp64pct = float(vals["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"])
            fp64_inst = fp64pct / 100.0 * cyc / 2.0 * 592
            print("\nDerived: warp instructions executed %.4g, of which FP64 ~%.4g (pipe-active cycles / 2 x 592 "
                  "sub-partitions) = %.1f %% of all issued instructions.\n" % (inst, fp64_inst, 100 * fp64_inst / inst))
        except Exception:
            pass


if __name__ == "__main__":
    main()
