#!/usr/bin/env python
"""tools/sass_loop.py FILE.sass [min_len] -- static look at the loops of one disassembled kernel (cuobjdump -sass output of ONE function):
lists the backward branches (loop back-edges) with the opcode histogram of their bodies, largest first.  Straight-line time-loop bodies
(the fused rollout kernels) show up as one long range whose length is the instruction count of a step when no inner branch skips code."""
import collections
import re
import sys

lines = open(sys.argv[1]).read().splitlines()
min_len = int(sys.argv[2]) if len(sys.argv) > 2 else 200
ins = []
for ln in lines:
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, txt) in enumerate(ins):
    m = re.search(r"\bBRA(?:\.\w+)*\s+(?:[!\w]+,\s*)?`?\(?\.?L?_?x?_?\w*\)?\s*$", txt)
    m2 = re.search(r"BRA.*?0x([0-9a-f]+)", txt)
    if "BRA" in txt and m2:
        t = int(m2.group(1), 16)
        if t < a and t in addr_index and i - addr_index[t] >= min_len:
            loops.append((i - addr_index[t] + 1, addr_index[t], i))
for n, lo, hi in sorted(loops, reverse=True)[:4]:
    c = collections.Counter()
    for _, txt in ins[lo:hi + 1]:
        parts = txt.split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        base = op.split(".")[0]
        c[base if not base.startswith("MUFU") else op] += 1
    print("loop %#x..%#x: %d instructions" % (ins[lo][0], ins[hi][0], n))
    print("   " + ", ".join("%s %d" % kv for kv in c.most_common(40)))
