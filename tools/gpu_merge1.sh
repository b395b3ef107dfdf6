#!/bin/bash
for st in 6 7; do echo "stride $st"; CHECK=1 DGX_MERGE_STRIDE=$st timeout 120 python tools/prof_merge.py 100000000 2>&1 | tail -2; done
