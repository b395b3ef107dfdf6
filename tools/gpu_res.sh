#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "resident or stress" 2>&1 | tail -3
for t in 1 0; do
DGX_PLAN_SHARED=$t timeout 150 python bench.py --steps 50 --warmup 5 --no-e2e --no-ops 2>gpurun_out/res_err.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); dv=d['dense_variant']; print('[shared $t]', 'C2 %.4f ms frac %.3f exact %s | dense %.4f ms frac %.3f %s' % (d['ms_per_step'], d['roofline']['frac'], d['bit_exact'], dv['ms_per_step'], dv['roofline_frac'], dv['bit_exact']))" || tail -3 gpurun_out/res_err.log
done
