#!/bin/bash
mkdir -p gpurun_out
tools/ab.sh libdgx.so libdgx.so:DGX_PREPASS_OVERLAP=0 libdgx_slp.so
for T in 4; do
timeout 120 python bench.py --steps 5 --warmup 3 --no-ops --no-dense --e2e-threads $T > gpurun_out/r2_bench6_$T.json 2> gpurun_out/r2_bench6_$T.err || { echo "T=$T FAILED rc=$?"; tail -3 gpurun_out/r2_bench6_$T.err; continue; }
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench6_$T.json").read().strip().splitlines()[-1])
print("T=$T image", *[(k, round(d[k]["ms_per_step"],3), d[k].get("pcie_GBps")) for k in ("e2e","e2e_cached")], d["bit_exact"], d["ms_per_step"])
PY
done
timeout 400 python -m pytest tests -m gpu -x -q -k "not full" 2>&1 | tail -3
