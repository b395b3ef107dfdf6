#!/usr/bin/env python
"""Group an .ncu-rep's per-line instruction counts by enclosing function/region markers.
usage: python profiles/ncu_groups.py <rep> <file> <ntiles>   (regions = top-level functions in <file>)"""
import csv, re, subprocess, sys
rep, fname, ntiles = sys.argv[1], sys.argv[2], float(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, data = "", None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or not r[0].isdigit(): continue
    try: data.append((cur, int(r[0]), int(r[hdr["Instructions Executed"]]), int(r[hdr["# Samples"]])))
    except Exception: pass
# region starts: lines in file that begin a function (heuristic) or carry a '// ----' marker
src = open(fname).read().splitlines()
marks = []
for i, l in enumerate(src, 1):
    if re.match(r"^(template|__device__|__global__|static|struct)", l) or "// ----" in l:
        marks.append((i, l.strip()[:70]))
tot = sum(d[2] for d in data); ts = sum(d[3] for d in data)
acc = {}
base = fname.split("/")[-1]
for f, ln, ie, sm in data:
    if f == base:
        name = "(top)"
        for m, t in marks:
            if m <= ln: name = f"{m}: {t}"
        key = name
    else:
        key = f
    a = acc.setdefault(key, [0, 0]); a[0] += ie; a[1] += sm
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print(f"{100*v[0]/tot:5.1f}% {v[0]/ntiles:8.0f}/tile  samples {100*v[1]/ts:5.1f}% | {k}")
print("total/tile", tot / ntiles)
