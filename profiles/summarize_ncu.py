#!/usr/bin/env python
"""Summarise ncu outputs into small text files for profiles/ (run in the build container; reads gpurun_out/)."""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def launches(path):
    rows = [l for l in open(path) if l.startswith('"')]
    agg = defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(io.StringIO("".join(rows))):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0].replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += float(r["Metric Value"]) / 1e6
    total = sum(v[1] for v in agg.values())
    out = ["kernel,launches,total_ms,avg_ms,share"]
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{k},{n},{ms:.3f},{ms / n:.4f},{ms / total:.4f}")
    return "\n".join(out)


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__inst_executed.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled", "smsp__warps_eligible.avg.per_cycle_active",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]


def raw(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = []
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS or h.startswith("smsp__average_warps_issue_stalled") or h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued"):
            out.append(f"{h},{u},{v}")
        if h in ("Kernel Name", "Grid Size", "Block Size"):
            out.append(f"{h},,{v}")
    return "\n".join(out)


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    print(launches(path) if mode == "launches" else raw(path))
