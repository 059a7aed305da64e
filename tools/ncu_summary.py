#!/usr/bin/env python
"""Summarise one `ncu --set full` report (.ncu-rep, read with `ncu -i ... --page raw --csv`) into the handful of numbers
DESIGN.md / profiles/ quote: duration, grid, registers, instruction mix, issue activity, occupancy, DRAM traffic, top stalls.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [more.ncu-rep ...] > profiles/r2_xxx_ncu_full.md"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def summarise(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(head, r))
        u = dict(zip(head, units))
        out.append(f"### {d['Kernel Name'][:100]}  grid {d['Grid Size']} block {d['Block Size']}  ({path})\n")
        out.append("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            hit = [h for h in head if h == k or h.endswith("." + k)]
            if hit:
                out.append(f"| {k} | {d[hit[0]]} | {u[hit[0]]} |")
        stalls = []
        for h in head:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(d[h].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        out.append("| stalled warps per issue-active cycle, by reason | " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:8]) + " | |")
        out.append("")
    return "\n".join(out)


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(summarise(p))
