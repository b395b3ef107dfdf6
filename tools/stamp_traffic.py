#!/usr/bin/env python
"""profiles/filter_kernel_traffic.json from an `ncu --set full` capture of filter_pipe_kernel (read here, no GPU):

    python tools/stamp_traffic.py gpurun_out/<capture>.ncu-rep [workload]

The file carries dram__bytes_read.sum + dram__bytes_write.sum per launch and the hash of the kernel sources it was
measured on; bench.py prints `roofline.traffic` only while that hash matches the sources it runs (a stale capture
prints null instead of an old number)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rep = sys.argv[1]
workload = sys.argv[2] if len(sys.argv) > 2 else "c2"
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
best = None
for r in rows[2:]:
    if "filter_pipe_kernel" not in r[idx["Kernel Name"]]:
        continue
    rd = float(r[idx["dram__bytes_read.sum"]]) * scale[units[idx["dram__bytes_read.sum"]]]
    wr = float(r[idx["dram__bytes_write.sum"]]) * scale[units[idx["dram__bytes_write.sum"]]]
    best = {"dram_bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
            "kernel_us": float(r[idx["gpu__time_duration.sum"]]), "kernel": "dgx::filter_pipe_kernel",
            "workload": workload, "capture": os.path.basename(rep), "kernel_src_sha": bench.kernel_source_sha()}
assert best, "no filter_pipe_kernel launch in the capture"
path = os.path.join(ROOT, "profiles", "filter_kernel_traffic.json")
json.dump(best, open(path, "w"), indent=1)
print(json.dumps(best))
