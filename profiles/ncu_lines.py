#!/usr/bin/env python
"""Per-CUDA-source-line instruction / stall-sample breakdown of an .ncu-rep captured with
--import-source on (kernels compiled with -lineinfo):  python profiles/ncu_lines.py <rep> [topN]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, data = "", None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or r[0] in ("", "Function Name") or not r[0].isdigit():
        continue
    try:
        ie = int(r[hdr["Instructions Executed"]])
        sm = int(r[hdr["# Samples"]])
    except (ValueError, KeyError):
        continue
    data.append((ie, sm, cur_file, int(r[0]), r[1].strip()[:100]))
tot_i = sum(d[0] for d in data) or 1
tot_s = sum(d[1] for d in data) or 1
print(f"total warp-instructions {tot_i}, samples {tot_s}")
print("---- by instructions executed")
for ie, sm, f, ln, src in sorted(data, reverse=True)[:top]:
    print(f"{ie:>11d} {100 * ie / tot_i:5.1f}%  samples {100 * sm / tot_s:5.1f}% | {f}:{ln} {src}")
print("---- by stall samples")
for ie, sm, f, ln, src in sorted(data, key=lambda d: -d[1])[:top // 2]:
    print(f"{sm:>8d} {100 * sm / tot_s:5.1f}%  inst {100 * ie / tot_i:5.1f}% | {f}:{ln} {src}")
