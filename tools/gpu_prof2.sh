#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "packed or cache" > gpurun_out/r2_t5.log 2>&1; tail -3 gpurun_out/r2_t5.log
for T in 2 4 8; do
timeout 300 python bench.py --steps 5 --warmup 3 --no-ops --no-dense --e2e-threads $T > gpurun_out/r2_bench5_$T.json 2> gpurun_out/r2_bench5.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench5_$T.json").read().strip().splitlines()[-1])
print("T=$T zero-copy", *[(k, round(d[k]["ms_per_step"],3), d[k].get("pcie_GBps")) for k in ("e2e","e2e_cached")], d["bit_exact"])
PY
done
DGX_ZERO_COPY=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-ops --no-dense --e2e-threads 4 > gpurun_out/r2_bench5_dma.json 2>> gpurun_out/r2_bench5.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench5_dma.json").read().strip().splitlines()[-1])
print("T=4 dma", *[(k, round(d[k]["ms_per_step"],3), d[k].get("pcie_GBps")) for k in ("e2e","e2e_cached")], d["bit_exact"])
PY
DGX_LIB=$PWD/dgraph_b200/libdgx_prof.so timeout 150 python tools/prof_pipe_waits.py > gpurun_out/r2_pipe_waits_c2.json 2> gpurun_out/r2_pw.err; tail -c 1200 gpurun_out/r2_pipe_waits_c2.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:filter_pipe_kernel -s 3 -c 1 -f -o gpurun_out/r2_pipe_fast2 \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-ops --no-dense > gpurun_out/r2_ncu_pipe.log 2>&1; tail -2 gpurun_out/r2_ncu_pipe.log
