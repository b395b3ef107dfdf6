#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show what a kernel is built from (no GPU needed):
UBLKCP = TMA bulk copy (cp.async.bulk), SYNCS.* = mbarrier operations, ATOMS/ATOMG = shared / global atomics,
SHFL / VOTE = warp shuffles and ballots.   python profiles/sass_evidence.py [dgraph_b200/libdgx.so]"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "dgraph_b200/libdgx.so"
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEEP = ("UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "ATOMS", "ATOMG", "NANOSLEEP", "VOTE", "SHFL", "REDUX", "MATCH", "LDGSTS",
        "HMMA", "UTCHMMA")
fn, cnt, total = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        total[fn] += 1
        if m.group(1).split(".")[0] in KEEP:
            cnt[fn][m.group(1)] += 1
for f in sorted(total, key=lambda x: -total[x]):
    name = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip().split("(")[0]
    print(f"{name}: {total[f]} instructions ({total[f] * 16 / 1024:.1f} KB)")
    for op, n in sorted(cnt[f].items()):
        print(f"    {op:<32} {n}")
