#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/.
  python scripts/ncu_summary.py launches gpurun_out/launches_c2_TAG.csv
  python scripts/ncu_summary.py full gpurun_out/prof_c2_TAG.ncu-rep
"""
import collections
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_lsu.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_wait_per_warp_active.pct',
        'smsp__warp_issue_stalled_not_selected_per_warp_active.pct', 'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct']


def launches(path):
    lines = [l for l in open(path) if l.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(io.StringIO(''.join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"][:100]
        v = float(row["Metric Value"].replace(',', ''))
        v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.3f} ms total (ncu serialised, cold cache)")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t / 1e6:10.3f} ms {100 * t / tot:5.1f}%  n={n:4d}  avg {t / n / 1e3:9.1f} us  {k}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {path}")
    for r in rows[2:]:
        print("---", r[hdr.index('Kernel Name')])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w} = {r[i]} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
