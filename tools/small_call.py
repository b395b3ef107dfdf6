#!/usr/bin/env python
"""Latency of ONE small call (C1: two lists of 1e5, device-resident and through host pointers) for a few values of
DGX_PIPE_MIN_VALUES -- run each value in its own process (the knob is read once)."""
import ctypes as C, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_ops as B
from dgraph_b200 import _lib, algo, pb
L = B.Lane()
rng = np.random.default_rng(1)
for n in (1_000, 10_000, 100_000, 1_000_000):
    a = np.unique(rng.integers(0, 100 * n, int(n * 1.02), dtype=np.uint64))[:n].copy()
    b = np.unique(rng.integers(0, 100 * n, int(n * 1.02), dtype=np.uint64))[:n].copy()
    ta, tb = torch.from_numpy(a.view(np.int64)).to(B.DEV), torch.from_numpy(b.view(np.int64)).to(B.DEV)
    out = torch.empty(n + 8, dtype=torch.int64, device=B.DEV); off = torch.zeros(2, dtype=torch.int64, device=B.DEV)
    ms, _ = B.timeit(lambda: L.filter_batch(0, [ta, tb], [0, 2], out, off), warm=5, reps=50)
    o = pb.List(None)
    la, lb = pb.List(a), pb.List(b)
    for _ in range(5): algo.IntersectWith(la, lb, o)
    t0 = time.perf_counter()
    for _ in range(50): algo.IntersectWith(la, lb, o)
    host_us = (time.perf_counter() - t0) / 50 * 1e6
    print(f"n={n}: device-resident {ms*1e3:.1f} us, host-pointer call {host_us:.1f} us, out {o.Uids.size}")
