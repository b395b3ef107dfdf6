"""Multi-GPU parity on real GPUs: the sharding plumbing with the KERNELS underneath (dgx_dev_filter_batch as the
compute step, NCCL for the all-gatherv; dgx_dev_merge_sorted behind the range-partitioned merge with its NCCL
all-to-all), checked against the oracle.  Skipped when the box has fewer than two GPUs;
the world_size-2 gloo test (tests/test_shard_gloo.py) covers the host logic on CPU."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import ctypes as C, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import gen
from dgraph_b200 import _lib
from dgraph_b200.shard import run_sharded_pairs, run_range_merge, exchange_by_range
from oracle import pyoracle as orc

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()
_lib.check(lib.dgx_init(local))
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
lane = lib.dgx_lane_create(local, C.c_void_p(stream.cuda_stream))
rng = np.random.default_rng(42)          # same inputs on every rank
npairs = 301
A, B = [], []
for i in range(npairs):
    n = int(rng.integers(0, 60000)); m = int(rng.integers(0, 90000))
    master = gen.zipf_gaps(rng, 120000)
    A.append(gen.thin(rng, master, n / 120000)); B.append(gen.thin(rng, master, m / 120000))

def compute(units, a_lists, b_lists):
    keep, ptrs, lens, koff = [], [], [], [0]
    for u in units:
        for l in (a_lists[u], b_lists[u]):
            t = torch.from_numpy(l.view(np.int64)).to(dev)
            keep.append(t); ptrs.append(t.data_ptr()); lens.append(l.size)
        koff.append(len(ptrs))
    n = len(ptrs)
    cap = sum(min(a_lists[u].size, b_lists[u].size) for u in units)
    out = torch.empty(cap + 8, dtype=torch.int64, device=dev)
    off = torch.zeros(len(units) + 1, dtype=torch.int64, device=dev)
    if len(units):
        _lib.check(lib.dgx_dev_filter_batch(lane, _lib.OP_INTERSECT, (C.c_void_p * n)(*ptrs), (C.c_size_t * n)(*lens),
                                            (C.c_size_t * len(koff))(*koff), len(units), C.c_void_p(out.data_ptr()), cap,
                                            C.c_void_p(off.data_ptr())))
        _lib.check(lib.dgx_lane_sync(lane))
    return out, off

ok = True
for partition in ("contiguous", "lpt"):
    out, off = run_sharded_pairs(dist, A, B, compute, device=dev, partition=partition)
    out = out.cpu().numpy().view(np.uint64); off = off.cpu().numpy()
    ok = ok and off.size == npairs + 1
    for i in range(npairs):
        ok = ok and np.array_equal(out[off[i]: off[i + 1]], orc.intersect_with(A[i], B[i]))

# ---- pattern 2: one MergeSorted over lists sharded BY LIST (rank r holds lists r, r + world, ...), then Difference
# against a list held by rank 0 only, cut by the same splitters
def merge(runs):
    if not runs:
        return torch.zeros(0, dtype=torch.int64, device=dev)
    n = len(runs)
    tot = sum(int(r.numel()) for r in runs)
    out = torch.empty(tot + 8, dtype=torch.int64, device=dev)
    out_len = torch.zeros(1, dtype=torch.int64, device=dev)
    _lib.check(lib.dgx_dev_merge_sorted(lane, (C.c_void_p * n)(*[r.data_ptr() for r in runs]),
                                        (C.c_size_t * n)(*[int(r.numel()) for r in runs]), n,
                                        C.c_void_p(out.data_ptr()), out.numel(), C.c_void_p(out_len.data_ptr())))
    _lib.check(lib.dgx_lane_sync(lane))
    return out[: int(out_len.item())]

def diff(a, b):
    ptrs = (C.c_void_p * 2)(a.data_ptr(), b.data_ptr()); lens = (C.c_size_t * 2)(int(a.numel()), int(b.numel()))
    out = torch.empty(int(a.numel()) + 8, dtype=torch.int64, device=dev)
    off = torch.zeros(2, dtype=torch.int64, device=dev)
    _lib.check(lib.dgx_dev_filter_batch(lane, _lib.OP_DIFFERENCE, ptrs, lens, (C.c_size_t * 2)(0, 2), 1,
                                        C.c_void_p(out.data_ptr()), int(a.numel()), C.c_void_p(off.data_ptr())))
    _lib.check(lib.dgx_lane_sync(lane))
    return out[: int(off[1].item())]

rng = np.random.default_rng(7)
master = gen.zipf_gaps(rng, 3000000)
lens = (1500000 / np.arange(1, 25)).astype(int)                      # 24 lists, lengths ~ 1/rank: the multiway path
for name, lists in (("config-5 shape", [gen.thin(rng, master, min(1.0, l / master.size)) for l in lens]),
                    ("few short lists", [gen.thin(rng, master, 0.001) for _ in range(3)]),
                    ("one list", [master[::2]])):
    mine = [torch.from_numpy(l.view(np.int64)).to(dev) for l in lists[rank::world]]
    want = orc.merge_sorted(lists)
    out, spl = run_range_merge(dist, mine, merge, device=dev)
    ok = ok and np.array_equal(out.cpu().numpy().view(np.uint64), want)
    part, spl = run_range_merge(dist, mine, merge, gather=False, device=dev)
    d = gen.thin(rng, master, 0.2)
    dr = exchange_by_range(dist, [torch.from_numpy(d.view(np.int64)).to(dev)] if rank == 0 else [], spl, device=dev)
    dpart = dr[0] if dr else torch.zeros(0, dtype=torch.int64, device=dev)
    res = diff(part, dpart) if part.numel() else part
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([int(res.numel())], device=dev))
    cap = max(max(int(x.item()) for x in sizes), 1)
    buf = torch.zeros(cap, dtype=torch.int64, device=dev); buf[: res.numel()] = res
    allb = [torch.zeros(cap, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(allb, buf)
    got = np.concatenate([allb[r][: int(sizes[r].item())].cpu().numpy().view(np.uint64) for r in range(world)])
    ok = ok and np.array_equal(got, orc.difference(want, d))
t = torch.tensor([int(ok)], device=dev)
dist.all_reduce(t)
if rank == 0:
    print("MULTIRANK_OK" if int(t.item()) == world else "MULTIRANK_FAIL")
dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(600)
def test_sharded_pairs_nccl_kernels():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs (found {n})")
    world = 2 if n < 4 else 4
    env = dict(os.environ)
    env.pop("NCCL_DEBUG", None)
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False, dir=ROOT) as f:
        f.write(WORKER)
        path = f.name
    try:
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                            "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), path],
                           capture_output=True, text=True, timeout=580, cwd=ROOT, env=env)
    finally:
        os.unlink(path)
    assert "MULTIRANK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
