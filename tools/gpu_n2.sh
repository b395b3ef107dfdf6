#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -3
P=$((20000 + RANDOM % 20000))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 5 --no-ops > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench N=$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().splitlines()[-1])
    print("N=$N value %.4g ms %.4f exact %s gather %s e2e %.4g cached %.4g" % (d["value"], d["ms_per_step"], d["bit_exact"], d["gather"], d["e2e"]["value"], d["e2e_cached"]["value"]))
except Exception as e:
    print("parse failed", e)
PY
tail -5 gpurun_out/r2_bench_n$N.err
P=$((20000 + RANDOM % 20000))
NCCL_DEBUG=VERSION timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P tools/bench_multi.py > gpurun_out/r2_multi_n$N.log 2> gpurun_out/r2_multi_n$N.err; echo "multi rc=$?"; cat gpurun_out/r2_multi_n$N.log | cut -c1-400; tail -3 gpurun_out/r2_multi_n$N.err
