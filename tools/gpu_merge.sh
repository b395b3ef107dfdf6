#!/bin/bash
o=gpurun_out; mkdir -p $o
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "merge" -x > $o/m_tests.log 2>&1; tail -3 $o/m_tests.log
timeout 200 python - <<'PY'
import ctypes as C, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from tools.bench_ops import Lane, zipf_gaps_gpu, thin_gpu, timeit, DEV
gen = torch.Generator(device=DEV); gen.manual_seed(9)
L = Lane()
master = zipf_gaps_gpu(200_000_000, gen)
for k in (64, 256, 1024):
    lists = [thin_gpu(master, 0.5 / k, gen) for _ in range(k)]
    tot = sum(t.numel() for t in lists)
    out = torch.empty(tot + 8, dtype=torch.int64, device=DEV); out_len = torch.zeros(1, dtype=torch.int64, device=DEV)
    res = {}
    for mode in ("grouped", "tree"):
        if mode == "tree":
            os.environ["DGX_MERGE"] = "tree"   # read once at dgx_init: needs a fresh process, so only report grouped here
            break
        ms, _ = timeit(lambda: L.merge(lists, out, out_len), warm=2, reps=5)
        L.sync(); nm = int(out_len.item())
        ok = bool(torch.equal(out[:nm], torch.unique(torch.cat(lists))))
        res[mode] = (round(ms, 3), ok)
    print(json.dumps({"k": k, "total": tot, "ms": res["grouped"][0], "check": res["grouped"][1]}), flush=True)
    del lists, out
PY
DGX_MERGE=tree timeout 200 python - <<'PY'
import json, os, sys
import torch
sys.path.insert(0, os.getcwd())
from tools.bench_ops import Lane, zipf_gaps_gpu, thin_gpu, timeit, DEV
gen = torch.Generator(device=DEV); gen.manual_seed(9)
L = Lane()
master = zipf_gaps_gpu(200_000_000, gen)
for k in (64, 256, 1024):
    lists = [thin_gpu(master, 0.5 / k, gen) for _ in range(k)]
    tot = sum(t.numel() for t in lists)
    out = torch.empty(tot + 8, dtype=torch.int64, device=DEV); out_len = torch.zeros(1, dtype=torch.int64, device=DEV)
    ms, _ = timeit(lambda: L.merge(lists, out, out_len), warm=2, reps=5)
    print(json.dumps({"k": k, "total": tot, "ms_tree": round(ms, 3)}), flush=True)
    del lists, out
PY
