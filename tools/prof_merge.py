#!/usr/bin/env python
"""Small driver for profiling MergeSorted / Decode under ncu: C5 shape at 1/4 scale, a few launches."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_ops as B
L = B.Lane()
gen = torch.Generator(device=B.DEV); gen.manual_seed(5)
total = int(sys.argv[1]) if len(sys.argv) > 1 else 25_000_000
master = B.zipf_gaps_gpu(2 * total, gen)
w = 1.0 / np.arange(1, 65); lens = (w / w.sum() * total).astype(np.int64)
lists = [B.thin_gpu(master, float(l) / master.numel(), gen) for l in lens]
tot = sum(t.numel() for t in lists)
out = torch.empty(tot + 8, dtype=torch.int64, device=B.DEV); out_len = torch.zeros(1, dtype=torch.int64, device=B.DEV)
for _ in range(3):
    L.merge(lists, out, out_len)
L.sync()
ms, _ = B.timeit(lambda: L.merge(lists, out, out_len), warm=1, reps=3)
print("merge ms", ms, "total", tot, "out", int(out_len.item()))
if os.environ.get("CHECK"):
    want = torch.unique(torch.cat(lists))
    n = int(out_len.item())
    print("check", bool(n == want.numel() and torch.equal(out[:n], want)))
