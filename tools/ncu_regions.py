#!/usr/bin/env python
"""Attribute the stall samples of `ncu --page source --csv --print-source sass` output to SASS regions (runs of
consecutive instructions with the same execution count, i.e. loop bodies).   usage: python tools/ncu_regions.py source.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
src, smp, ex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
data = []
for r in rows[hi + 1:]:
    if len(r) <= ex or not r[ex].isdigit():
        continue
    data.append((r[src].strip(), int(r[smp] or 0), int(r[ex])))
tot = sum(d[1] for d in data)
runs = []
for i, d in enumerate(data):
    t = d[0].split()
    op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")).split(".")[0]
    if runs and runs[-1]["ex"] == d[2]:
        r = runs[-1]
        r["n"] += 1; r["smp"] += d[1]; r["end"] = i
    else:
        r = dict(ex=d[2], n=1, smp=d[1], start=i, end=i, ops={})
        runs.append(r)
    r["ops"][op] = r["ops"].get(op, 0) + 1
print(f"{len(data)} SASS instructions, {tot} samples")
for r in runs:
    if r["smp"] > tot * 0.004 or r["n"] > 20:
        top = sorted(r["ops"].items(), key=lambda x: -x[1])[:7]
        print(f"{r['start']:5d}-{r['end']:5d} n={r['n']:4d} exec={r['ex']:10d} samples={r['smp']:6d} ({100 * r['smp'] / tot:4.1f}%) {top}")
