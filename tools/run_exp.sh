#!/bin/bash
# A/B runs of experimental library builds: tools/run_exp.sh libdgx_X.so ...  (C2 bench line + C1 batch timing each)
for lib in "$@"; do
  out=$(DGX_LIB=$PWD/dgraph_b200/$lib timeout 120 python bench.py --steps 50 --warmup 3 --e2e-steps 1 2>/dev/null | tail -1)
  c2=$(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("%.4f"%d["ms_per_step"], d.get("bit_exact"))')
  c1=$(WORKLOAD=c1 PAIRS=1024 DGX_LIB=$PWD/dgraph_b200/$lib timeout 120 python tools/prof_pipe_waits.py 2>/dev/null | tail -1)
  echo "$lib :: C2 $c2 :: C1 $c1"
done
