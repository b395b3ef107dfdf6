#!/bin/bash
mkdir -p gpurun_out
for t in 16384 0 8192; do
DGX_PRE_THROTTLE=$t timeout 150 python bench.py --steps 50 --warmup 5 --no-e2e --no-ops 2>gpurun_out/res_err.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); dv=d['dense_variant']; print('[throttle $t]', 'C2 %.4f ms frac %.3f exact %s | dense %.4f ms frac %.3f %s' % (d['ms_per_step'], d['roofline']['frac'], d['bit_exact'], dv['ms_per_step'], dv['roofline_frac'], dv['bit_exact']))" || tail -3 gpurun_out/res_err.log
done
