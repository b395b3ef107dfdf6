#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
for rs in 0 8 16; do
P=$((20000 + RANDOM % 20000))
DGX_RESERVE_CTAS=$rs timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 30 --warmup 5 --no-ops --no-e2e --no-dense 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('reserve $rs N', d['n_gpus'], 'ms', round(d['ms_per_step'],4), 'value %.4g' % d['value'], d['bit_exact'], d['step_ms_rank0'])"
done
P=$((20000 + RANDOM % 20000))
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P tools/bench_multi.py 2>/dev/null > gpurun_out/r2_multi_n${N}b.log; python -c "
import sys,json
for l in open('gpurun_out/r2_multi_n${N}b.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['config'], 'N', d['n_gpus'], 'ms', round(d['ms'],3), 'compute-only', round(d['ms_compute_only'],3), 'check', d['check'])"
