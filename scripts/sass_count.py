#!/usr/bin/env python
"""Instruction counts per kernel from `cuobjdump -sass <obj>` (static, whole function).
usage: sass_count.py build/exb_fastnd_n512.o [substring filter]"""
import re, subprocess, sys, collections
obj = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
name = None
cnt = collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cnt[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(2).split(".")[0]
        cnt[name][op] += 1
        cnt[name]["_total"] += 1
for k, c in cnt.items():
    if flt in k:
        top = ", ".join(f"{o}:{n}" for o, n in c.most_common(14) if o != "_total")
        print(f"{c['_total']:6d}  {k[:110]}\n        {top}")
