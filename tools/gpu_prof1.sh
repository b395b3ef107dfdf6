#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "not full" > gpurun_out/r2_t3.log 2>&1; tail -3 gpurun_out/r2_t3.log
DGX_LIB=$PWD/dgraph_b200/libdgx_prof.so timeout 150 python tools/prof_pipe_waits.py > gpurun_out/r2_pipe_waits_c2.json 2> gpurun_out/r2_pw.err; tail -c 1500 gpurun_out/r2_pipe_waits_c2.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:filter_pipe_kernel -s 3 -c 1 -f -o gpurun_out/r2_pipe_fast \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-ops --no-dense > gpurun_out/r2_ncu_pipe.log 2>&1; tail -2 gpurun_out/r2_ncu_pipe.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-ops --no-dense > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench3.json").read().strip().splitlines()[-1])
for k in ("e2e","e2e_cached","e2e_raw_u64"): print(k, d[k]["ms_per_step"], d[k]["value"], d[k].get("pcie_GBps"), d[k]["bit_exact"])
print(d["ms_per_step"], d["roofline"]["frac"], d["bit_exact"])
PY
tail -5 gpurun_out/r2_bench3.err
