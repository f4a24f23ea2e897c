#!/bin/bash
# Blackwell-native evidence from the built library: per kernel, how many tcgen05 / TMEM / TMA / bulk-copy instructions
# its SASS holds (cuobjdump -sass on seesaw_b200/libseesaw_b200.so).  Mnemonics (guides/B200_PROFILING.md):
#   UTCHMMA = tcgen05.mma (fp16 kind)   LDTM / STTM = tcgen05.ld / tcgen05.st (tensor memory)
#   UTMALDG = cp.async.bulk.tensor (TMA tile load)   UBLKCP = cp.async.bulk (1-D bulk copy)
#   UTCBAR = tcgen05.commit   SYNCS = mbarrier ops
# usage: scripts/sass_summary.sh > profiles/r02_sass_summary.txt
set -e
cd "$(dirname "$0")/.."
LIB=seesaw_b200/libseesaw_b200.so
echo "# $(date -u +%Y-%m-%dT%H:%MZ)  cuobjdump -sass $LIB  ($(cuobjdump --version | tail -1))"
echo "# arch: $(cuobjdump -lelf $LIB | head -3 | tr '\n' ' ')"
cuobjdump -sass "$LIB" | python3 -c '
import re, sys, collections
ops = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "LDGSTS"]
counts, cur, order = collections.defaultdict(collections.Counter), None, []
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        order.append(cur)
        continue
    if cur is None:
        continue
    for op in ops:
        if re.search(r"\b" + op + r"\b|\b" + op + r"\.", line):
            counts[cur][op] += 1
import subprocess
def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    except OSError:
        return n
print("%-9s %-7s %-6s %-6s %-8s %-7s %-7s %-6s  kernel" % ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA"))
tot = collections.Counter()
for f in order:
    c = counts[f]
    tot.update(c)
    name = demangle(f)
    name = re.sub(r"\(.*", "", name)[:110]
    print("%-9d %-7d %-6d %-6d %-8d %-7d %-7d %-6d  %s" % (c["UTCHMMA"], c["LDTM"], c["STTM"], c["UTMALDG"], c["UBLKCP"], c["UTCBAR"], c["SYNCS"], c["HMMA"], name))
print("%-9d %-7d %-6d %-6d %-8d %-7d %-7d %-6d  TOTAL (%d kernels)" % (tot["UTCHMMA"], tot["LDTM"], tot["STTM"], tot["UTMALDG"], tot["UBLKCP"], tot["UTCBAR"], tot["SYNCS"], tot["HMMA"], len(order)))
'
