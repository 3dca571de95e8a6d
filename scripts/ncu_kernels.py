#!/usr/bin/env python
"""Aggregate an ncu CSV (gpu__time_duration + dram bytes per launch) by kernel name:
   python scripts/ncu_kernels.py gpurun_out/launches_X.csv"""
import collections, csv, io, re, sys
f = sys.argv[1]
lines = [l for l in open(f) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO(''.join(lines))))
byid = collections.OrderedDict()
for r in rows:
    d = byid.setdefault(r["ID"], {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(',', '')); u = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["us"] = v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
    else:
        d[r["Metric Name"]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
agg = collections.OrderedDict()
for d in byid.values():
    nm = d["name"].replace("exb::", "")
    nm = re.sub(r"NlS<\(int\)(\d), \(int\)-?\d, \(int\)\d, \(int\)\d>", r"S\1", nm)
    nm = re.sub(r"\(int\)", "", nm)[:86]
    if not ("fast" in nm or "pass" in nm or "k1d" in nm or "etdrk" in nm or "copy_b" in nm or "col_" in nm or "forcing" in nm):
        continue
    key = (nm, round(d.get("dram__bytes_read.sum", 0) / 2e7), round(d.get("dram__bytes_write.sum", 0) / 2e7))
    a = agg.setdefault(key, [0, 0.0, 0.0, 0.0]); a[0] += 1; a[1] += d["us"]
    a[2] += d.get("dram__bytes_read.sum", 0); a[3] += d.get("dram__bytes_write.sum", 0)
tot = sum(a[1] for a in agg.values())
print(f"# {f}: total {tot/1e3:.3f} ms over {sum(a[0] for a in agg.values())} launches")
for (nm, _, _), (n, us, rd, wr) in agg.items():
    print(f"n={n:3d} avg {us/n:9.1f} us {100*us/tot:5.1f}%  rd {rd/n/1e6:8.1f} MB wr {wr/n/1e6:8.1f} MB -> {(rd+wr)/us/1e6:5.2f} TB/s  {nm}")
