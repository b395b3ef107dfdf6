#!/bin/bash
o=gpurun_out; mkdir -p $o
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "merge" -x > $o/m_tests.log 2>&1; tail -3 $o/m_tests.log
timeout 120 python tools/bench_merge.py --dense 2>&1 | tail -3
echo p128; DGX_LIB=$PWD/dgraph_b200/libdgx_p128.so timeout 120 python tools/bench_merge.py 2>&1 | tail -1
for s in 9 11; do echo stride $s; DGX_MERGE_STRIDE=$s timeout 120 python tools/bench_merge.py 2>&1 | tail -1; done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"mmerge|mplan|msample|merge_kernel|mscan|mcompact|mtail" -c 40 --csv --log-file $o/m_launches.csv python tools/bench_merge.py --reps 1 > $o/m_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(l for l in open('gpurun_out/m_launches.csv') if l.startswith('"')))
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[ki][:40],[]).append(float(r[vi].replace(',','')))
for k,v in agg.items(): print(k,len(v),'avg us',round(sum(v)/len(v)/1e3,1))
PY
