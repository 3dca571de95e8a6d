#!/usr/bin/env python
"""Per-kernel summary of an .ncu-rep: time, DRAM bytes, occupancy, issue utilisation, smem wavefronts,
top warp-stall reasons.  usage: ncu_stalls.py file.ncu-rep"""
import csv, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
def f(r, name):
    try:
        return float(r[col[name]].replace(",", ""))
    except Exception:
        return float("nan")
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    print("---", r[col["Kernel Name"]][:100])
    unit = {h: rows[1][i] for h, i in col.items()}
    print("  time %.3f %s  dram rd %.3f %s wr %.3f %s  dram%% %.1f  regs %d  warps_active %.1f%%  issue_active %.1f%%" % (
        f(r, "gpu__time_duration.sum"), unit["gpu__time_duration.sum"], f(r, "dram__bytes_read.sum"), unit["dram__bytes_read.sum"],
        f(r, "dram__bytes_write.sum"), unit["dram__bytes_write.sum"],
        f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), f(r, "launch__registers_per_thread"),
        f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")))
    print("  inst %.0f  smem wavefronts %.0f (conflicts %.0f)  l1 lsu-pipe %.1f%%  fma %.1f%%  alu %.1f%%" % (
        f(r, "smsp__inst_executed.sum"), f(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        f(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        f(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        f(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        f(r, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active")))
    vals = sorted(((f(r, h), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in stall), reverse=True)[:7]
    print("  stalls: " + "  ".join("%s=%.2f" % (n, v) for v, n in vals))
