#!/bin/bash
mkdir -p gpurun_out
tools/ab.sh libdgx.so libdgx_nofast.so 2>&1
timeout 400 python -m pytest tests -m gpu -x -q -k "not full" > gpurun_out/r2_t2.log 2>&1; tail -3 gpurun_out/r2_t2.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-ops > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench2.json").read().strip().splitlines()[-1])
for k in ("e2e","e2e_cached","e2e_raw_u64"): print(k, d[k]["ms_per_step"], d[k]["value"], d[k].get("pcie_GBps"), d[k]["bit_exact"])
print(d["ms_per_step"], d["roofline"]["frac"], d["bit_exact"])
PY
tail -5 gpurun_out/r2_bench2.err
