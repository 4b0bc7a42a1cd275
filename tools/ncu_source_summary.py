#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source sass`: executed instruction mix by opcode and stall reasons.

usage: ncu -i rep.ncu-rep --page source --csv --print-source sass | python tools/ncu_source_summary.py [units]
`units` (optional) divides instruction counts, e.g. theta*steps/32 to get warp-instructions per theta*step.
"""
import csv
import re
import sys
from collections import Counter

units = float(sys.argv[1]) if len(sys.argv) > 1 else None
rows = list(csv.reader(sys.stdin))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr_i]
src, ex, smp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
ops, stalls, samples_by_op = Counter(), Counter(), Counter()
total = 0
for r in rows[hdr_i + 1:]:
    if len(r) <= ex or not r[ex].isdigit():
        continue
    n = int(r[ex])
    t = re.sub(r"^@!?U?P\d+\s+", "", r[src].strip())
    op = t.split()[0].split(".")[0] if t else "?"
    if op in ("IMAD",) and ".MOV" in t:
        op = "IMAD.MOV"
    ops[op] += n
    total += n
    samples_by_op[op] += int(r[smp] or 0)
    for i, c in stall_cols:
        if r[i].isdigit():
            stalls[c] += int(r[i])
print(f"total warp-instructions executed: {total}" + (f"  ({total / units:.1f} per unit)" if units else ""))
fp64 = sum(ops[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
print(f"FP64 pipe (DFMA+DMUL+DADD+DSETP): {fp64} = {100 * fp64 / total:.1f}%" + (f"  ({fp64 / units:.1f} per unit)" if units else ""))
for op, n in ops.most_common(25):
    print(f"  {op:10s} {n:12d} {100 * n / total:5.1f}%" + (f" {n / units:7.1f}/unit" if units else "") + f"  samples {samples_by_op[op]}")
ts = sum(stalls.values())
print("stall samples:", ", ".join(f"{c[6:]} {100 * n / ts:.1f}%" for c, n in stalls.most_common(8)))
