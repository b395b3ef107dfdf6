#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mmerge_kernel -s 1 -c 1 -f -o gpurun_out/r2_mmerge \
    python tools/prof_merge.py 100000000 > gpurun_out/r2_ncu_merge.log 2>&1; tail -3 gpurun_out/r2_ncu_merge.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(msample|mplan|mmerge|mscan|mcompact|merge)_kernel' -c 40 --csv --log-file gpurun_out/r2_merge_launches.csv \
    python tools/prof_merge.py 100000000 > gpurun_out/r2_ncu_merge2.log 2>&1; tail -12 gpurun_out/r2_merge_launches.csv | cut -c1-200
