#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into the short text summaries kept in this directory.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep "title line" > profiles/rNN_ncu_<kernel>_summary.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]


def main():
    rep, title = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(title)
    for r in rows[2:]:
        print(f"kernel: {r[hdr.index('Kernel Name')]}")
        for h, u, v in sorted(zip(hdr, units, r)):
            if h in KEYS or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                print(f"{h} [{u}] = {v}")


if __name__ == "__main__":
    main()
