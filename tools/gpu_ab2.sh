#!/bin/bash
mkdir -p gpurun_out
tools/ab.sh "$@" 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "not full" > gpurun_out/r2_t4.log 2>&1; tail -5 gpurun_out/r2_t4.log
for T in 4 8 16; do
timeout 300 python bench.py --steps 5 --warmup 3 --no-ops --no-dense --e2e-threads $T > gpurun_out/r2_bench4_$T.json 2> gpurun_out/r2_bench4.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench4_$T.json").read().strip().splitlines()[-1])
print("T=$T", *[(k, round(d[k]["ms_per_step"],3), d[k].get("pcie_GBps")) for k in ("e2e","e2e_cached","e2e_raw_u64")])
PY
done
