#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
for piece in 8388608 1048576; do
P=$((20000 + RANDOM % 20000))
DGX_GATHER_PIECE=$piece timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P tools/bench_multi.py > gpurun_out/r2_multi_n${N}_$piece.log 2> gpurun_out/r2_multi_n$N.err; echo "piece $piece rc=$?"; cat gpurun_out/r2_multi_n${N}_$piece.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'], 'N', d['n_gpus'], 'ms', round(d['ms'],3), 'compute-only', round(d['ms_compute_only'],3), 'check', d['check'])"
done
