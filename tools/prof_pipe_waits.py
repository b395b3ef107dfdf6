#!/usr/bin/env python
"""Wait-time breakdown of filter_pipe_kernel's warp roles on the C2 workload.

Needs an experimental build with -DDGX_PIPE_PROF (exports dgx_debug_pprof):
  make prof        (nvcc ... -DDGX_PIPE_PROF -shared -o dgraph_b200/libdgx_prof.so ...)
  DGX_LIB=$PWD/dgraph_b200/libdgx_prof.so python tools/prof_pipe_waits.py
Instrumented builds are slower; the numbers are shares of each role's own time, not a bench value.
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dgraph_b200 import _lib  # noqa: E402


def main():
    lib = _lib.load()
    _lib.check(lib.dgx_init(0))
    raw = C.CDLL(os.environ.get("DGX_LIB", os.path.join(os.path.dirname(_lib.__file__), "libdgx.so")))
    have_prof = hasattr(raw, "dgx_debug_pprof")  # plain builds: timing only
    if have_prof:
        raw.dgx_debug_pprof.argtypes = [C.POINTER(C.c_uint64)]
    Q = int(os.environ.get("Q", "16"))
    if os.environ.get("WORKLOAD", "c2") == "c1":  # C1 batch: 2048 independent 2-way intersections of 1e5-UID lists
        rng = np.random.default_rng(7)
        queries = []
        for _ in range(int(os.environ.get("PAIRS", "2048"))):
            queries.append([np.unique(rng.integers(0, 10_000_000, 102_000, dtype=np.uint64))[:100_000].copy() for _ in range(2)])
        Q = len(queries)
    elif os.environ.get("WORKLOAD") == "c4":  # C4 shape: 2-way pairs, thinnings p = 0.5 of a per-pair master (half of A survives)
        rng = np.random.default_rng(9)
        queries = []
        for _ in range(int(os.environ.get("PAIRS", "512"))):
            n = int(rng.integers(20_000, 400_000))
            master = np.cumsum(np.minimum(rng.zipf(1.5, 2 * n), 1 << 20).astype(np.uint64), dtype=np.uint64)
            queries.append([master[rng.random(master.size) < 0.5] for _ in range(2)])
        Q = len(queries)
    else:
        queries = bench.make_queries(Q, 0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    lane = lib.dgx_lane_create(0, C.c_void_p(stream.cuda_stream))
    keep, ptrs, lens, k_off = [], [], [], [0]
    for qq in queries:
        for l in qq:
            t = torch.from_numpy(l.view(np.int64)).to(dev)
            keep.append(t); ptrs.append(t.data_ptr()); lens.append(l.size)
        k_off.append(len(ptrs))
    nl = len(ptrs)
    c_ptrs = (C.c_void_p * nl)(*ptrs); c_lens = (C.c_size_t * nl)(*lens); c_koff = (C.c_size_t * (Q + 1))(*k_off)
    out_cap = sum(min(l.size for l in qq) for qq in queries)
    d_out = torch.empty(out_cap + 2, dtype=torch.int64, device=dev)
    d_off = torch.zeros(Q + 1, dtype=torch.int64, device=dev)

    def step():
        _lib.check(lib.dgx_dev_filter_batch(lane, _lib.OP_INTERSECT, c_ptrs, c_lens, c_koff, Q,
                                            C.c_void_p(d_out.data_ptr()), out_cap, C.c_void_p(d_off.data_ptr())))
    buf = (C.c_uint64 * 32)()
    for _ in range(3):
        step()
    _lib.check(lib.dgx_lane_sync(lane))
    if have_prof:
        raw.dgx_debug_pprof(buf)  # clear
    steps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    _lib.check(lib.dgx_lane_sync(lane))
    if not have_prof:
        print(json.dumps({"ms_per_step": e0.elapsed_time(e1) / steps, "workload": os.environ.get("WORKLOAD", "c2")}))
        return
    raw.dgx_debug_pprof(buf)
    v = [int(x) for x in buf]
    def share(a, b):
        return round(a / b, 4) if b else None
    out = {
        "ms_per_step_instrumented": e0.elapsed_time(e1) / steps,
        "consumer": {"wait_stage_full": share(v[1], v[0]), "wait_out_slot": share(v[2], v[0])},
        "tma_warp": {"wait_descriptor": share(v[5], v[4]), "wait_stage_empty": share(v[6], v[4])},
        "metadata_warp": {"wait_ring_slot": share(v[9], v[8])},
        "output_warp": {"wait_slot_full": share(v[13], v[12]), "look_back": share(v[14], v[12])},
        "consumer_phases": {"tile_setup": share(v[16], v[0]), "list_searches": share(v[17], v[0]),
                            "deliver_survivors": share(v[18], v[0])},
        "raw_cycles": v,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
