#!/usr/bin/env python
"""Aggregate an ncu source-page export (SASS rows with warp-stall samples) by CUDA source line.

    ncu -i rep.ncu-rep --page source --csv > sass.csv
    cuobjdump -xelf all libelg_b200.so ; nvdisasm -g -c <tu>.cubin > lines.txt
    python tools/ncu_source_lines.py sass.csv lines.txt <mangled kernel name> <source file name> [top N]

Rows are joined by instruction order (same function, same SASS).  Inlined helper code is charged to the line of
<source file> it was inlined at.  Prints samples / executed instructions per line and the dominant stall reasons."""
import csv
import re
import sys
from collections import defaultdict

sass_csv, lines_txt, kernel, srcname = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40

# --- nvdisasm: ordered list of source lines, one per instruction of the kernel
cur, inside, per_instr = None, False, []
for ln in open(lines_txt):
    if ln.startswith(".text."):
        inside = ln.strip() == ".text.%s:" % kernel
        continue
    if not inside:
        continue
    if "//## File" in ln:
        m = re.findall(r'File "([^"]+)", line (\d+)', ln)
        hit = [int(n) for f, n in m if f.endswith(srcname)]
        if hit:
            cur = hit[-1]
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        per_instr.append(cur)

rows = list(csv.reader(open(sass_csv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
body = [r for r in rows[h + 1:] if len(r) == len(hdr)]
iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
if len(body) != len(per_instr):
    sys.stderr.write("warning: %d SASS rows vs %d disassembled instructions\n" % (len(body), len(per_instr)))
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
tot = 0
for r, line in zip(body, per_instr):
    s, n = int(r[iS] or 0), int(r[iI] or 0)
    a = agg[line]
    a[0] += s
    a[1] += n
    tot += s
    for c in stall_cols:
        v = int(r[c] or 0)
        if v:
            a[2][hdr[c][6:]] += v
src = {}
try:
    for i, l in enumerate(open(srcname if "/" in srcname else "elg_b200/csrc/" + srcname), 1):
        src[i] = l.rstrip()
except OSError:
    pass
print("total samples %d" % tot)
for line, (s, n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    reasons = ", ".join("%s %d%%" % (k, 100 * v // max(1, s)) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print("%5s %5.1f%% inst %9d  %-42s | %s" % (line, 100.0 * s / tot, n, reasons, src.get(line, "")[:90].strip()))
