#!/bin/bash
o=gpurun_out; mkdir -p $o
timeout 400 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x > $o/n2_multirank.log 2>&1; tail -4 $o/n2_multirank.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/bench_multi.py --pairs 2000 > $o/n2_multi.jsonl 2> $o/n2_multi.err; echo "rc=$?"; cut -c1-420 $o/n2_multi.jsonl; tail -3 $o/n2_multi.err | cut -c1-300
