#!/usr/bin/env python
"""Summarise the loops of a cuobjdump -sass listing: for every backward branch, the instruction mix of its body.

usage: cuobjdump -sass file.o | python tools/sass_loops.py [kernel-substring]
"""
import re
import sys
from collections import Counter

pat = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);")
want = sys.argv[1] if len(sys.argv) > 1 else None
kern, ins = None, []


def flush():
    if not ins or (want and want not in (kern or "")):
        return
    print("== kernel", (kern or "")[:110], "total instr", len(ins))
    addr = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, txt) in enumerate(ins):
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", txt)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr:
                body = ins[addr[tgt]:i + 1]
                c = Counter()
                for _, t in body:
                    t = re.sub(r"^@!?U?P\d+\s+", "", t)
                    c[t.split()[0].split(".")[0]] += 1
                fp64 = sum(c[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
                print(f"  loop {tgt:#x}..{a:#x}: {len(body)} instr, fp64={fp64}, MUFU={c['MUFU']}, "
                      f"LDL/STL={c['LDL']}/{c['STL']}, LDG/STG={c['LDG']}/{c['STG']}, CALL={c['CALL']}")
                print("    ", dict(c.most_common(14)))


for line in sys.stdin:
    if "Function :" in line:
        flush()
        kern, ins = line.split("Function :")[1].strip(), []
        continue
    m = pat.match(line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
flush()
