#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into the short text summaries kept in this directory.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep "title line" > profiles/rNN_ncu_<kernel>_summary.txt
       python profiles/ncu_summary.py --digest profiles/r02_ncu_digest.json key=rep[:trials] [key=rep ...]
           machine-readable digest of several captures (what bench.py attaches to its JSON line instead of literals):
           key = the kernel key bench.py asks for (k_trials_dense32, k_trials_table, k_trials_prefix, k_trials_full32)"""
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


def digest(out, specs):
    import json
    import os
    d = {}
    for spec in specs:
        key, rep = spec.split("=", 1)
        trials = None
        if ":" in rep:
            rep, trials = rep.rsplit(":", 1)
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr = rows[0]
        r = rows[2]
        g = lambda name: float(r[hdr.index(name)].replace(",", "")) if name in hdr and r[hdr.index(name)] not in ("", "n/a") else None
        stalls = {h.split("issue_stalled_")[1].split("_per_issue")[0]: float(v) for h, v in zip(hdr, r)
                  if "issue_stalled" in h and h.endswith("per_issue_active.ratio")}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:4]
        d[key] = {
            "kernel": r[hdr.index("Kernel Name")], "capture": os.path.basename(rep), "trials_in_capture": int(trials) if trials else None,
            "duration_ms": g("gpu__time_duration.sum") / 1e6 if units_of(rows, "gpu__time_duration.sum") == "ns" else g("gpu__time_duration.sum"),
            "dram_bytes_per_launch": (g("dram__bytes_read.sum") or 0) * scale_of(rows, "dram__bytes_read.sum") + (g("dram__bytes_write.sum") or 0) * scale_of(rows, "dram__bytes_write.sum"),
            "warp_instructions": g("smsp__inst_executed.sum"),
            "issue_slots_busy_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "issue_slots_busy_pct_of_elapsed": g("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
            "alu_pipe_cycles_active_pct": g("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
            "fma_pipe_cycles_active_pct": g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
            "xu_pipe_inst_pct": g("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
            "lsu_wavefronts_pct_of_peak": g("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            "shared_wavefronts": g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
            "shared_bank_conflicts": g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
            "achieved_occupancy_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "registers_per_thread": g("launch__registers_per_thread"), "threads_per_block": g("launch__block_size"), "grid": g("launch__grid_size"),
            "top_stalls_per_issue": dict(top),
        }
    with open(out, "w") as f:
        json.dump(d, f, indent=1)
    print(json.dumps(d, indent=1))


def units_of(rows, name):
    hdr, units = rows[0], rows[1]
    return units[hdr.index(name)] if name in hdr else None


def scale_of(rows, name):
    u = units_of(rows, name)
    return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def main():
    if sys.argv[1] == "--digest":
        return digest(sys.argv[2], sys.argv[3:])
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
