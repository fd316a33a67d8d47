#!/usr/bin/env python
"""tools/sass_lines.py FILE.cubin KERNEL_SUBSTRING [LO_HEX HI_HEX] -- static instruction count per source line (nvdisasm -g -c, needs -lineinfo),
optionally restricted to an address range (e.g. the time-loop range printed by sass_loop.py).  Lines inlined from other places are attributed to
the innermost line, as in ncu's source page."""
import collections
import re
import subprocess
import sys

cubin, kern = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 62
txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
inside = False
cur = ("?", 0)
count = collections.Counter()
ops = collections.defaultdict(collections.Counter)
for ln in txt:
    if ln.startswith("//---") and ".text." in ln:
        inside = kern in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        a = int(m.group(1), 16)
        if lo <= a <= hi:
            parts = m.group(2).split()
            op = parts[1] if parts[0].startswith("@") else parts[0]
            count[cur] += 1
            ops[cur][op.split(".")[0]] += 1
total = sum(count.values())
print("instructions in range: %d" % total)
srcs = {}
for (f, l), n in count.most_common(int(sys.argv[5]) if len(sys.argv) > 5 else 60):
    if f not in srcs:
        try:
            import glob
            path = [p for p in glob.glob("/root/repo/raptor_b200/csrc/" + f) + glob.glob("/usr/local/cuda/include/crt/" + f)]
            srcs[f] = open(path[0]).read().splitlines() if path else []
        except Exception:
            srcs[f] = []
    src = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
    print("%5d %5.1f%%  %-16s:%-4d %-60s %s" % (n, 100.0 * n / total, f, l, ", ".join("%s %d" % kv for kv in ops[(f, l)].most_common(4)), src))
