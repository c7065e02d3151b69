#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name.
    python tools/launch_summary.py gpurun_out/launches.csv [skip_first_n_launches]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
for r in rows[1 + skip:]:
    name = re.sub(r"\(.*", "", r[ki])
    t = float(r[vi].replace(",", ""))
    t = t / 1e3 if r[ui] in ("ns", "nsecond") else (t * 1e3 if r[ui] in ("ms", "msecond") else t)   # -> us
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f%% |" % (k, n, t, 100 * t / tot))
print("| total | %d | %.1f | |" % (sum(a[0] for a in agg.values()), tot))
