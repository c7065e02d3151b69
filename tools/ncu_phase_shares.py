#!/usr/bin/env python
"""Stall samples / executed instructions of an ncu source-page export, summed over line ranges (phases) of a kernel.

    python tools/ncu_phase_shares.py sass.csv lines.txt <mangled kernel> <source file>     (inputs as for ncu_source_lines.py)

The `ranges` table below holds the phase boundaries of rollout_tc.cu at the commit profiles/r02_rollout_tc_ncu.md was
taken from; edit it when the file moves (grep -n PHASE_MARK gives the boundaries)."""
import csv, re, sys
from collections import defaultdict
sass_csv, lines_txt, kernel, srcname = sys.argv[1:5]
ranges = [("helpers / setup (inlined)", 0, 262), ("Q row + list prefetch lambdas", 263, 308), ("step start", 309, 323), ("L1 list validity + exchange", 324, 357),
          ("L2 bit scan + records + features", 358, 432), ("L3 local scores + max/neighbour-bit exchange", 433, 491), ("L4 weights -> TMEM + MMA1", 492, 548),
          ("L5 ol + MMA2 / first QK issue", 549, 623), ("L6 local finalize", 624, 658), ("softmax rounds", 659, 745),
          ("O operand + score issue", 746, 791), ("B3", 792, 879), ("select", 880, 902), ("phase C", 903, 973), ("epilogue", 974, 3000)]
cur, inside, per_instr = None, False, []
for ln in open(lines_txt):
    if ln.startswith(".text."):
        inside = ln.strip() == ".text.%s:" % kernel; continue
    if not inside: continue
    if "//## File" in ln:
        m = re.findall(r'File "([^"]+)", line (\d+)', ln)
        hit = [int(n) for f, n in m if f.endswith(srcname)]
        if hit: cur = hit[-1]
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln): per_instr.append(cur)
rows = list(csv.reader(open(sass_csv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address"); hdr = rows[h]
body = [r for r in rows[h + 1:] if len(r) == len(hdr)]
print(len(body), len(per_instr))
iS, iI, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
stall_cols = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
agg = defaultdict(lambda: [0, 0, 0, defaultdict(int), 0])
tot = toti = 0
for r, line in zip(body, per_instr):
    name = next((n for n, a, b in ranges if line is not None and a <= line <= b), "setup/other")
    a = agg[name]; s = int(r[iS] or 0); n = int(r[iI] or 0)
    a[0] += s; a[1] += n; a[2] += int(r[iT] or 0); a[4] += 1; tot += s; toti += n
    for c in stall_cols:
        v = int(r[c] or 0)
        if v: a[3][hdr[c][6:]] += v
for n, _, _ in ranges:
    s, ni, nt, st, k = agg[n]
    rs = ", ".join("%s %d%%" % (k2, 100 * v // max(1, s)) for k2, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
    print("%-22s samples %5.1f%%  warp-instr %5.1f%% (%6.2fG, %4.1f thr/warp, %5d static)  %s" % (n, 100. * s / tot, 100. * ni / toti, ni / 1e9, nt / max(1, ni), k, rs))
