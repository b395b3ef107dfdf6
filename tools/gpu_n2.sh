#!/bin/bash
o=gpurun_out; mkdir -p $o
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $o/n2_bench.json 2> $o/n2_bench.err; echo "rc=$?"; tail -c 1500 $o/n2_bench.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=2 value %.4g ms %.4f frac %.3f exact %s e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('bit_exact'), d['e2e']['value']))"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -q 2>&1 | tail -2
