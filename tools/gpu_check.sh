#!/bin/bash
# GPU parity suite + one short headline bench line (used between kernel iterations).
tag=${1:-chk}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x --timeout 90 > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
timeout 120 python bench.py --e2e-steps 1 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print("value %.4g ms %.4f frac %.3f bit_exact %s"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d.get("bit_exact")))
PY
