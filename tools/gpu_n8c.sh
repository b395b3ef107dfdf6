#!/bin/bash
N=${1:-8}
run() {
P=$((20000 + RANDOM % 20000))
env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 30 --warmup 5 --no-ops --no-e2e --no-dense 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$*', 'N', d['n_gpus'], 'ms', round(d['ms_per_step'],4), 'value %.4g' % d['value'], d['bit_exact'], d['step_ms_rank0'])"
}
run A=1
run NCCL_MAX_NCHANNELS=1
run NCCL_MAX_NCHANNELS=2 DGX_RESERVE_CTAS=16
run DGX_RESERVE_CTAS=16
