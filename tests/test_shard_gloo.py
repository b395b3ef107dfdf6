"""world_size-2 gloo test of the multi-GPU host logic (LPT partition + all-gatherv
reassembly).  The compute step is stood in for by the oracle (this is a test of the
sharding plumbing, not of the kernels)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(__file__))


def test_lpt_partition_balanced_and_deterministic():
    from dgraph_b200.shard import lpt_partition

    rng = np.random.default_rng(0)
    costs = rng.integers(1, 10**6, 1000)
    for world in (1, 2, 4, 8):
        parts = lpt_partition(costs, world)
        allu = np.sort(np.concatenate(parts))
        assert np.array_equal(allu, np.arange(1000))
        loads = np.array([costs[p].sum() for p in parts])
        assert loads.max() - loads.min() <= costs.max()
        parts2 = lpt_partition(costs, world)
        assert all(np.array_equal(a, b) for a, b in zip(parts, parts2))
    assert [p.tolist() for p in lpt_partition([5, 5, 5], 2)] == [[0, 2], [1]]
    assert [p.tolist() for p in lpt_partition([], 2)] == [[], []]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import gen
        from dgraph_b200.shard import run_sharded_pairs
        from oracle import pyoracle as orc

        rng = np.random.default_rng(42)  # same inputs on every rank
        npairs = 37
        A, B = [], []
        for i in range(npairs):
            n, m = int(rng.integers(0, 3000)), int(rng.integers(0, 5000))
            master = gen.zipf_gaps(rng, 8000)
            A.append(gen.thin(rng, master, n / 8000))
            B.append(gen.thin(rng, master, m / 8000))

        def compute(units, a_lists, b_lists):
            outs = [orc.intersect_with(a_lists[u], b_lists[u]) for u in units]
            off = np.zeros(len(units) + 1, dtype=np.int64)
            if outs:
                off[1:] = np.cumsum([o.size for o in outs])
            cat = np.concatenate(outs + [np.zeros(0, np.uint64)]).view(np.int64)
            return torch.from_numpy(cat.copy()), torch.from_numpy(off)

        ok = True
        for partition in ("contiguous", "lpt"):
            out, off = run_sharded_pairs(dist, A, B, compute, partition=partition)
            out = out.numpy().view(np.uint64)
            off = off.numpy()
            ok = ok and off.size == npairs + 1
            for i in range(npairs):
                ok = ok and np.array_equal(out[off[i]: off[i + 1]], orc.intersect_with(A[i], B[i]))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_pairs_gloo_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_contiguous_partition():
    from dgraph_b200.shard import contiguous_partition

    rng = np.random.default_rng(1)
    costs = rng.integers(1, 10**6, 997)
    for world in (1, 2, 3, 8):
        parts = contiguous_partition(costs, world)
        assert np.array_equal(np.concatenate(parts), np.arange(997))
        loads = np.array([costs[p].sum() for p in parts])
        assert loads.max() - loads.min() <= 2 * costs.max()
    assert [p.tolist() for p in contiguous_partition([], 3)] == [[], [], []]
    assert [p.tolist() for p in contiguous_partition([5], 3)][-1] == [0] or sum(len(p) for p in contiguous_partition([5], 3)) == 1


def _worker_contig(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dgraph_b200.shard import contiguous_partition, gatherv_contiguous

        rng = np.random.default_rng(7)
        n = 23
        results = [np.sort(rng.integers(0, 1000, int(rng.integers(0, 50)), dtype=np.int64)) for _ in range(n)]
        parts = contiguous_partition([r.size + 1 for r in results], world)
        mine = parts[rank]
        outs = [results[u] for u in mine]
        off = np.zeros(len(mine) + 1, dtype=np.int64)
        if outs:
            off[1:] = np.cumsum([o.size for o in outs])
        cat = np.concatenate(outs + [np.zeros(0, np.int64)])
        out, goff = gatherv_contiguous(dist, torch.from_numpy(cat.copy()), torch.from_numpy(off))
        out, goff = out.numpy(), goff.numpy()
        ok = goff.size == n + 1 and all(np.array_equal(out[goff[i]: goff[i + 1]], results[i]) for i in range(n))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gatherv_contiguous_gloo_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_contig, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def _worker_range(rank, world, port, ret):
    """Range-partitioned MergeSorted over lists sharded by list (SURVEY 8e pattern 2): splitters, all-to-all of the
    slices, local merge (the oracle stands in for the kernel here; tests/test_gpu_multirank.py runs it with
    dgx_dev_merge_sorted underneath), rank-ordered all-gatherv.  Then the second operand cut by the same splitters."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import gen
        from dgraph_b200.shard import exchange_by_range, run_range_merge
        from oracle import pyoracle as orc

        def as_t(a):
            return torch.from_numpy(np.ascontiguousarray(a).view(np.int64).copy())

        def merge(runs):
            if not runs:
                return torch.zeros(0, dtype=torch.int64)
            return as_t(orc.merge_sorted([r.numpy().view(np.uint64) for r in runs]))

        rng = np.random.default_rng(11)  # same inputs on every rank; rank r HOLDS lists r, r + world, ...
        master = gen.zipf_gaps(rng, 60000)
        cases = {
            "skewed": [gen.thin(rng, master, 1.0 / (i + 1)) for i in range(11)],
            "with empties and duplicates": [np.zeros(0, np.uint64), np.sort(rng.integers(0, 500, 4000).astype(np.uint64)),
                                            np.zeros(0, np.uint64), np.arange(100, 900, dtype=np.uint64), np.full(300, 7, np.uint64)],
            "fewer lists than ranks": [master[::3]],
            "nothing at all": [],
            "above 2^63": [np.sort(rng.integers(0, 2**64 - 1, 5000, dtype=np.uint64)) for _ in range(5)]
                          + [np.array([0, 2**63 - 1, 2**63, 2**64 - 1], dtype=np.uint64)],
        }
        ok = True
        for name, lists in cases.items():
            mine = [as_t(l) for l in lists[rank::world]]
            want = orc.merge_sorted(lists) if lists else np.zeros(0, np.uint64)
            out, spl = run_range_merge(dist, mine, merge, oversample=64)
            ok = ok and np.array_equal(out.numpy().view(np.uint64), want)
            part, _ = run_range_merge(dist, mine, merge, gather=False, oversample=64)
            parts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(parts, torch.tensor([part.numel()]))
            ok = ok and sum(int(p.item()) for p in parts) == want.size          # ranges are disjoint: no cross-rank repeats
            # a second operand held by rank 0 only, cut by the same splitters: every rank gets its range of it
            d = gen.thin(rng, master, 0.3)
            got = exchange_by_range(dist, [as_t(d)] if rank == 0 else [], spl)
            sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(sizes, torch.tensor([sum(int(g.numel()) for g in got)]))
            ok = ok and sum(int(s.item()) for s in sizes) == d.size
            if got and part.numel():
                lo, hi = int(part.numpy().view(np.uint64).min()), int(part.numpy().view(np.uint64).max())
                g = torch.cat(got).numpy().view(np.uint64)
                skey = spl.numpy().view(np.uint64)
                if rank > 0:
                    ok = ok and bool((g >= skey[rank - 1]).all())
                if rank < world - 1:
                    ok = ok and bool((g < skey[rank]).all())
                del lo, hi
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world", [2, 3])
def test_range_merge_gloo(world):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_range, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {r: True for r in range(world)}
