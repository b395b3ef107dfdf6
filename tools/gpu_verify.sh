#!/bin/bash
# Last check of a round on one GPU: smoke(), the whole GPU suite, one bench line.
o=gpurun_out; mkdir -p $o
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 900 python -m pytest tests -m gpu -q --durations=3 ${PYTEST_K:+-k "$PYTEST_K"} > $o/r2e_tests.log 2>&1; tail -3 $o/r2e_tests.log
timeout 400 python bench.py > $o/r2e_bench.json 2> $o/r2e_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2e_bench.json").read().strip().splitlines()[-1])
print("value %.4g ms %.4f frac %.3f exact %s e2e %.4g launches %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["bit_exact"], d["e2e"]["value"], d["gpu_launches"]))
ops = d["ops"]
if isinstance(ops, dict): print("OPS ERROR", str(ops)[:300])
else:
    for o in ops: print(o["op"], "|", o["ms"], "|", o["frac_of_hbm_peak"], o["check"])
PY
