#!/bin/bash
o=gpurun_out; mkdir -p $o
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > $o/r2c_bench_ref.json 2> $o/r2c_bench_ref.err
timeout 400 python bench.py > $o/r2c_bench.json 2> $o/r2c_bench.err; echo "bench rc=$?"
timeout 200 python bench.py --no-resident --no-e2e --no-ops --steps 50 --warmup 5 > $o/r2c_bench_noresident.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/r2c_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-ops --no-dense > $o/r2c_ncu_bench.log 2>&1
timeout 600 python -m pytest tests -m gpu -q > $o/r2c_tests.log 2>&1; tail -2 $o/r2c_tests.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c_bench.json").read().strip().splitlines()[-1])
r=json.loads(open("gpurun_out/r2c_bench_ref.json").read().strip().splitlines()[-1])
n=json.loads(open("gpurun_out/r2c_bench_noresident.json").read().strip().splitlines()[-1])
print("value %.4g ms %.4f frac %.3f exact %s | no-resident ms %.4f frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["bit_exact"], n["ms_per_step"], n["roofline"]["frac"]))
print("dense", d["dense_variant"]["ms_per_step"], d["dense_variant"]["roofline_frac"])
for k in ("e2e","e2e_cached","e2e_raw_u64"): print(k, round(d[k]["ms_per_step"],3), "%.4g"%d[k]["value"], d[k].get("pcie_GBps"), d[k]["bit_exact"])
print("ref", r["value"], "ratio e2e", d["e2e"]["value"]/r["value"], "ratio value", d["value"]/r["value"])
for o in d["ops"]: print(o["op"], "|", o["ms"], "|", o["frac_of_hbm_peak"], o["check"])
PY
