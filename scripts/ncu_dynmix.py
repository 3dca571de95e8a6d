#!/usr/bin/env python
"""Dynamic instruction mix per kernel from an .ncu-rep captured with --import-source on:
executed warp-instructions by opcode (SASS page).  usage: ncu_dynmix.py file.ncu-rep [kernel substring]"""
import csv, subprocess, sys, collections, re
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
flt = sys.argv[2] if len(sys.argv) > 2 else ""
kern, hdr, seen = None, None, set()
mix = collections.OrderedDict()
for row in csv.reader(txt.splitlines()):
    if not row:
        continue
    if row[0] == "Kernel Name":
        kern = row[1]
        if kern in seen:
            kern = None  # one launch per kernel is enough
        else:
            seen.add(kern)
            mix[kern] = collections.Counter()
        hdr = None
        continue
    if row[0] == "Address":
        hdr = {h: i for i, h in enumerate(row)}
        continue
    if kern is None or hdr is None:
        continue
    src = row[hdr["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
    if not m:
        continue
    op = m.group(2)
    try:
        n = int(row[hdr["Instructions Executed"]])
    except ValueError:
        continue
    mix[kern][op] += n
    mix[kern]["_total"] += n
for k, c in mix.items():
    if flt not in k:
        continue
    tot = c["_total"]
    print(f"--- {k[:120]}\n  total warp-instructions {tot}")
    groups = {"fp32": ("FADD", "FMUL", "FFMA", "FADD2", "FMUL2", "FFMA2", "FSEL", "FSETP", "FMNMX", "MUFU"),
              "smem": ("LDS", "STS", "LDSM"), "global": ("LDG", "STG", "LDGSTS", "LDL", "STL"),
              "int/addr": ("IMAD", "IADD3", "LEA", "LOP3", "SHF", "ISETP", "VIADD", "SEL", "IABS", "I2FP", "VIMNMX", "MOV", "PRMT", "CS2R", "PLOP3", "R2P", "P2R", "UIADD3", "UMOV", "ULEA", "UIMAD", "S2R", "LDC", "LDCU", "HFMA2", "IADD", "I2F", "F2I", "UISETP", "USEL", "ULOP3", "USHF", "R2UR"),
              "control": ("BRA", "BSSY", "BSYNC", "BAR", "EXIT", "CALL", "RET", "WARPSYNC", "NOP", "YIELD", "DEPBAR", "ERRBAR", "LDGDEPBAR")}
    acc = collections.Counter()
    for op, n in c.items():
        if op == "_total":
            continue
        for g, ops in groups.items():
            if op in ops:
                acc[g] += n
                break
        else:
            acc["other:" + op] += n
    print("  " + "  ".join(f"{g} {100 * n / tot:.1f}%" for g, n in acc.most_common()))
    print("  top: " + ", ".join(f"{o}:{100 * n / tot:.1f}%" for o, n in c.most_common(16) if o != "_total"))
