#!/bin/bash
# A/B of library builds on the headline step (device-resident, no e2e/ops): tools/ab.sh libdgx.so libdgx_X.so:ENV=VAL ...
i=0
for spec in "$@"; do
  i=$((i+1))
  lib=${spec%%:*}; envs=""; [[ "$spec" == *:* ]] && envs=${spec#*:}
  out=$(env ${envs//,/ } DGX_LIB=$PWD/dgraph_b200/$lib timeout 120 python bench.py --steps 50 --warmup 5 --no-e2e --no-ops 2>gpurun_out/ab_err_$i.log | tail -1)
  echo "$out" | python -c '
import sys,json
d=json.loads(sys.stdin.read())
dv=d.get("dense_variant") or {}
print("'$spec' :: C2 %.4f ms frac %.3f exact %s :: dense %.4f ms frac %.3f exact %s" % (d["ms_per_step"], d["roofline"]["frac"], d["bit_exact"], dv.get("ms_per_step",0), dv.get("roofline_frac",0), dv.get("bit_exact")))' 2>/dev/null || { echo "$spec :: FAILED"; tail -3 gpurun_out/ab_err_$i.log; }
done
