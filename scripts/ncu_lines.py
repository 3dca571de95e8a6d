#!/usr/bin/env python
"""Hot source lines per kernel from an .ncu-rep captured with --import-source on (needs -lineinfo):
executed warp-instructions aggregated per CUDA source line.  usage: ncu_lines.py file.ncu-rep <kernel substring> [top]"""
import csv, subprocess, sys, collections, os
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
flt = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 40
fpath, fn, hdr = None, None, None
agg = collections.OrderedDict()
for row in csv.reader(txt.splitlines()):
    if not row:
        continue
    if row[0] == "File Path":
        fpath = os.path.basename(row[1]); continue
    if row[0] == "Function Name":
        fn = row[1]; continue
    if row[0] == "Line No":
        hdr = {h: i for i, h in enumerate(row)}
        iexec = [i for i, h in enumerate(row) if h == "Instructions Executed"][0]
        isamp = [i for i, h in enumerate(row) if h == "# Samples"][0]
        iwf = [i for i, h in enumerate(row) if h == "L1 Wavefronts Shared"][0]
        continue
    if hdr is None or fn is None or flt not in fn or row[0] == "":
        continue
    try:
        n = int(row[iexec]); smp = int(row[isamp]); wf = int(row[iwf] or 0)
    except ValueError:
        continue
    d = agg.setdefault(fn, collections.OrderedDict())
    key = (fpath, int(row[0]), row[1].strip()[:110])
    if key in d:
        d[key] = (d[key][0] + n, d[key][1] + smp, d[key][2] + wf)
    else:
        d[key] = (n, smp, wf)
for fn, d in agg.items():
    tot = sum(v[0] for v in d.values()); ts = sum(v[1] for v in d.values()); tw = max(1, sum(v[2] for v in d.values()))
    print(f"=== {fn[:130]}\n    total warp-instr {tot}, samples {ts}")
    key = (lambda kv: -kv[1][2]) if "--wf" in sys.argv else (lambda kv: -kv[1][0])
    for (f, ln, src), (n, smp, wf) in sorted(d.items(), key=key)[:top]:
        print(f"  {100*n/tot:5.1f}% inst {100*smp/max(ts,1):5.1f}% smp {100*wf/tw:5.1f}% smem-wf  {f}:{ln}  {src}")
    break
