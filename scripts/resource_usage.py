#!/usr/bin/env python
"""Registers / spill stack / static shared memory of every kernel in build/*.o (cuobjdump, no GPU needed).
usage: python scripts/resource_usage.py > profiles/rNN_resource_usage.txt"""
import glob, os, re, subprocess
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
rows = []
for obj in sorted(glob.glob(os.path.join(root, "build", "*.o"))):
    txt = subprocess.run(["cuobjdump", "--dump-resource-usage", obj], capture_output=True, text=True).stdout
    name = None
    for l in txt.splitlines():
        m = re.match(r"\s*Function (\S+):", l)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", l)
        if m and name:
            rows.append((name,) + tuple(int(x) for x in m.groups()))
            name = None
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump --dump-resource-usage of build/*.o (sm_100a): registers, stack (spill) bytes, static shared memory")
print(f"# {len(rows)} kernels / device functions; {sum(1 for r in rows if r[2])} with a stack frame")
for n, r in sorted(zip(names, rows)):
    print(f"REG {r[1]:4d} STACK {r[2]:4d} SHARED {r[3]:6d}  {n[:170]}")
