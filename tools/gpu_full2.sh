#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q -k "not full" 2>&1 | tail -3
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench8.json").read().strip().splitlines()[-1])
print("value %.4g ms %.4f frac %.3f exact %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["bit_exact"]))
print("dense", d["dense_variant"]["ms_per_step"], d["dense_variant"]["roofline_frac"])
for k in ("e2e","e2e_cached","e2e_raw_u64"): print(k, round(d[k]["ms_per_step"],3), "%.4g"%d[k]["value"], d[k].get("pcie_GBps"), d[k]["bit_exact"])
for r in d["ops"]: print(r["op"], "|", r["ms"], "ms |", r["frac_of_hbm_peak"], r["check"])
PY
