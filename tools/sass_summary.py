#!/usr/bin/env python
"""Per-kernel counts of the Blackwell instructions in the built library (cuobjdump -sass), for profiles/.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt

UTCHMMA = tcgen05.mma (kind::f16), LDTM / STTM = tcgen05.ld / st (tensor memory), UTCBAR = tcgen05.commit,
UBLKCP = cp.async.bulk (1-D TMA), UTMALDG = cp.async.bulk.tensor (tiled TMA), SYNCS = mbarrier ops."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "elg_b200", "csrc", "libelg_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "FFMA", "MUFU", "ATOMS", "SHFL", "BAR"]
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1).split(".")[0]
        counts[kern][op] += 1
        counts[kern]["_all"] += 1
print("# SASS instruction counts per kernel of %s (sm_100a)" % os.path.basename(lib))
print("%-70s %7s " % ("kernel", "instr") + " ".join("%7s" % o for o in OPS))
for k, c in counts.items():
    if c["_all"] < 50:
        continue
    print("%-70s %7d " % (k[:70], c["_all"]) + " ".join("%7d" % c[o] for o in OPS))
    total.update(c)
print("%-70s %7d " % ("TOTAL", total["_all"]) + " ".join("%7d" % total[o] for o in OPS))
