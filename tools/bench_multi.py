#!/usr/bin/env python
"""Multi-GPU runs of BASELINE configs C4 and C5 (one process per GPU, launched by torchrun):

  C4  batched independent 2-way intersections, sizes ~ power law on [1e4, 1e6]: units are
      independent, so they are partitioned by bytes over the ranks (dgraph_b200/shard.py),
      every rank runs ONE batched libdgx launch on its units, and the results return in pair
      order through an NCCL all-gatherv.  No collective on the data path.
  C5  one huge MergeSorted (k = 64, total --merge-total UIDs) + Difference: the UID space is
      range-partitioned (G-1 splitters from a sample; equal keys land on one rank, so there are no
      cross-rank duplicates), every rank merges its range of every list and subtracts the same
      range of the second operand, results are concatenated in rank order (all-gatherv).  Inputs
      are replicated on every rank here (same seed), so the all-to-all exchange is skipped.
  C5-sharded  the same operation with the lists sharded BY LIST (rank r holds lists r, r + N, ...): splitters from a
      strided sample of every rank's lists, one all-to-all of the slices (shard.exchange_by_range), local merge and
      difference, all-gatherv -- everything inside the timed region.

    python -m torch.distributed.run --nproc-per-node N tools/bench_multi.py [--pairs 10000] [--merge-total 1e8]

Prints one JSON line per config on rank 0.  Times are CUDA events, max over ranks.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dgraph_b200 import _lib, shard  # noqa: E402


def zipf_gaps_gpu(n, gen, dev, a=1.5, clip=1 << 20):
    u = torch.rand(n, device=dev, generator=gen, dtype=torch.float64).clamp_(min=1e-12)
    g = torch.clamp(torch.floor(u.pow(-1.0 / (a - 1.0))), max=float(clip)).to(torch.int64)
    return torch.cumsum(g, 0)


def thin(master, p, gen):
    return master[torch.rand(master.numel(), device=master.device, generator=gen) < p].contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=10000)
    ap.add_argument("--merge-total", type=float, default=1e8)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if os.environ.get("NCCL_DEBUG"):
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's log never goes to stdout (JSON lines only)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    _lib.check(lib.dgx_init(local))
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    lane = lib.dgx_lane_create(local, C.c_void_p(stream.cuda_stream))

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---------------- C4 ----------------------------------------------------------------------
    npairs = args.pairs
    rng = np.random.default_rng(401)
    u = rng.random(npairs)
    sizes = np.minimum((1e4 / (1 - u * (1 - 1e4 / 1e6))).astype(np.int64), 1_000_000)  # power law, alpha = 2
    parts = shard.contiguous_partition([16 * int(s) for s in sizes], world)  # contiguous: no scatter on the way back
    mine = parts[rank]
    gen = torch.Generator(device=dev)
    lists, koff = [], [0]
    for i in mine:
        gen.manual_seed(402 + int(i))  # pair i is the same whichever rank owns it
        master = zipf_gaps_gpu(int(sizes[i]) * 2, gen, dev)
        lists += [thin(master, 0.5, gen), thin(master, 0.5, gen)]
        koff.append(len(lists))
    n = len(lists)
    ptrs = (C.c_void_p * max(n, 1))(*[t.data_ptr() for t in lists])
    lens = (C.c_size_t * max(n, 1))(*[t.numel() for t in lists])
    ckoff = (C.c_size_t * len(koff))(*koff)
    cap = sum(min(lists[2 * j].numel(), lists[2 * j + 1].numel()) for j in range(len(mine)))
    out = torch.empty(cap + 8, dtype=torch.int64, device=dev)
    off = torch.zeros(len(mine) + 1, dtype=torch.int64, device=dev)
    res = {}

    def c4_compute():
        _lib.check(lib.dgx_dev_filter_batch(lane, _lib.OP_INTERSECT, ptrs, lens, ckoff, len(mine),
                                            C.c_void_p(out.data_ptr()), cap, C.c_void_p(off.data_ptr())))

    def c4_step():
        c4_compute()
        if world > 1:
            res["g"] = shard.gatherv_exact(dist, parts, out, off, device=dev)
        else:
            res["g"] = (out, off)

    ms_compute = timed(c4_compute)
    ms = timed(c4_step)
    uids_local = sum(t.numel() for t in lists)
    tot = torch.tensor([uids_local], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    gout, goff = res["g"]
    ok = True
    if rank == 0:  # spot-check a few pairs against torch on this rank (regenerating them)
        for i in (0, npairs // 2, npairs - 1):
            gen.manual_seed(402 + i)
            master = zipf_gaps_gpu(int(sizes[i]) * 2, gen, dev)
            a, b = thin(master, 0.5, gen), thin(master, 0.5, gen)
            want = a[torch.isin(a, b, assume_unique=True)]
            got = gout[int(goff[i]):int(goff[i + 1])]
            ok = ok and bool(torch.equal(got, want))
        print(json.dumps({"config": "C4", "n_gpus": world, "pairs": npairs, "total_uids": int(tot.item()),
                          "ms": ms, "uids_per_s": int(tot.item()) / (ms * 1e-3),
                          "ms_compute_only": ms_compute, "uids_per_s_compute_only": int(tot.item()) / (ms_compute * 1e-3),
                          "out_uids": int(goff[-1].item()), "check": ok,
                          "note": "contiguous shards balanced by bytes, one batched launch per rank, NCCL all-gatherv of results inside the timed region"}),
              flush=True)
    del lists, out

    # ---------------- C5 ----------------------------------------------------------------------
    total = int(args.merge_total)
    gen.manual_seed(501)  # replicated inputs: same seed on every rank
    master = zipf_gaps_gpu(2 * total, gen, dev)
    w = 1.0 / np.arange(1, 65)
    lens5 = (w / w.sum() * total).astype(np.int64)
    lists = [thin(master, float(l) / master.numel(), gen) for l in lens5]
    dlist = thin(master, (total / 10) / master.numel(), gen)
    # splitters: G-1 quantiles of a sample of the longest list (all lists are thinnings of one master)
    q = torch.linspace(0, 1, world + 1, device=dev)[1:-1]
    samp = lists[0][:: max(1, lists[0].numel() // 65536)]
    spl = samp[(q * (samp.numel() - 1)).long()] if world > 1 else torch.empty(0, dtype=torch.int64, device=dev)
    lo = torch.tensor([-1], dtype=torch.int64, device=dev) if rank == 0 else spl[rank - 1:rank]
    sl = []
    for t in lists + [dlist]:
        a = 0 if rank == 0 else int(torch.searchsorted(t, lo, right=False).item())
        b = t.numel() if rank == world - 1 else int(torch.searchsorted(t, spl[rank:rank + 1], right=False).item())
        sl.append(t[a:b])
    mlists, dl = sl[:-1], sl[-1]
    tot_local = sum(t.numel() for t in mlists)
    mptrs = (C.c_void_p * 64)(*[t.data_ptr() for t in mlists])
    mlens = (C.c_size_t * 64)(*[t.numel() for t in mlists])
    mout = torch.empty(tot_local + 8, dtype=torch.int64, device=dev)
    mlen = torch.zeros(1, dtype=torch.int64, device=dev)
    dout = torch.empty(tot_local + 8, dtype=torch.int64, device=dev)
    doff = torch.zeros(2, dtype=torch.int64, device=dev)
    res5 = {}

    def c5_compute():
        c5_step(gather=False)

    def c5_step(gather=True):
        _lib.check(lib.dgx_dev_merge_sorted(lane, mptrs, mlens, 64, C.c_void_p(mout.data_ptr()), tot_local,
                                            C.c_void_p(mlen.data_ptr())))
        # Difference(merged, D) on this rank's range; the merged length stays on the device
        torch.cuda.current_stream().synchronize()  # merged length is needed to size the next call
        nm = int(mlen.item())
        p2 = (C.c_void_p * 2)(mout.data_ptr(), dl.data_ptr())
        l2 = (C.c_size_t * 2)(nm, dl.numel())
        k2 = (C.c_size_t * 2)(0, 2)
        _lib.check(lib.dgx_dev_filter_batch(lane, _lib.OP_DIFFERENCE, p2, l2, k2, 1, C.c_void_p(dout.data_ptr()),
                                            tot_local, C.c_void_p(doff.data_ptr())))
        if world > 1 and gather:
            res5["g"] = shard.gatherv_exact(dist, [[0]] * world, dout, doff, device=dev)
        else:
            res5["g"] = (dout, doff)

    ms_compute = timed(c5_compute)
    ms = timed(c5_step)
    gout, goff = res5["g"]
    if rank == 0:
        allm = torch.unique(torch.cat(lists))
        want = allm[~torch.isin(allm, dlist, assume_unique=True)]
        nres = int(goff[-1].item())
        ok = bool(nres == want.numel() and torch.equal(gout[:nres], want))
        tin = sum(t.numel() for t in lists)
        print(json.dumps({"config": "C5", "n_gpus": world, "k": 64, "total_uids": tin, "ms": ms,
                          "uids_per_s": (tin + dlist.numel()) / (ms * 1e-3), "ms_compute_only": ms_compute,
                          "out_uids": nres, "check": ok,
                          "note": "range-partitioned MergeSorted + Difference, inputs replicated, rank-ordered all-gatherv inside the timed region"}),
              flush=True)
    # ---------------- C5, lists sharded BY LIST ---------------------------------------------------------
    # rank r holds lists r, r + world, ... (and rank 0 the second operand): splitters from a strided sample, ONE
    # all-to-all of the slices (shard.exchange_by_range), then the same local merge + difference + all-gatherv.
    if world > 1:
        mine = lists[rank::world]
        held_d = [dlist] if rank == 0 else []
        res6 = {}

        def merge_runs(runs):
            n = len(runs)
            tot = sum(int(r.numel()) for r in runs)
            out = torch.empty(tot + 8, dtype=torch.int64, device=dev)
            ln = torch.zeros(1, dtype=torch.int64, device=dev)
            if n:
                _lib.check(lib.dgx_dev_merge_sorted(lane, (C.c_void_p * n)(*[r.data_ptr() for r in runs]),
                                                    (C.c_size_t * n)(*[int(r.numel()) for r in runs]), n,
                                                    C.c_void_p(out.data_ptr()), out.numel(), C.c_void_p(ln.data_ptr())))
            torch.cuda.current_stream().synchronize()
            return out[: int(ln.item())]

        def c6_step(gather=True, exchange_only=False):
            part, spl = shard.run_range_merge(dist, mine, (lambda runs: runs) if exchange_only else merge_runs,
                                              gather=False, device=dev)
            if exchange_only:
                return
            dr = shard.exchange_by_range(dist, held_d, spl, device=dev)
            dpart = dr[0] if dr else torch.zeros(0, dtype=torch.int64, device=dev)
            out = torch.empty(int(part.numel()) + 8, dtype=torch.int64, device=dev)
            off = torch.zeros(2, dtype=torch.int64, device=dev)
            if part.numel():
                p2 = (C.c_void_p * 2)(part.data_ptr(), dpart.data_ptr())
                l2 = (C.c_size_t * 2)(int(part.numel()), int(dpart.numel()))
                _lib.check(lib.dgx_dev_filter_batch(lane, _lib.OP_DIFFERENCE, p2, l2, (C.c_size_t * 2)(0, 2), 1,
                                                    C.c_void_p(out.data_ptr()), int(part.numel()), C.c_void_p(off.data_ptr())))
            res6["g"] = shard.gatherv_exact(dist, [[0]] * world, out, off, device=dev) if gather else (out, off)

        ms_x = timed(lambda: c6_step(exchange_only=True))
        ms6 = timed(c6_step)
        gout, goff = res6["g"]
        if rank == 0:
            nres = int(goff[-1].item())
            ok = bool(nres == want.numel() and torch.equal(gout[:nres], want))
            print(json.dumps({"config": "C5-sharded", "n_gpus": world, "k": 64, "total_uids": tin, "ms": ms6,
                              "ms_splitters_and_exchange": ms_x, "uids_per_s": (tin + dlist.numel()) / (ms6 * 1e-3),
                              "out_uids": nres, "check": ok,
                              "note": "lists sharded by list: splitters + all-to-all of the slices + local merge + difference + rank-ordered all-gatherv, all inside the timed region"}),
                  flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
