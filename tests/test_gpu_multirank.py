"""Multi-GPU parity on real GPUs: the sharding plumbing with the KERNELS underneath (dgx_dev_filter_batch as the
compute step, NCCL for the all-gatherv), checked against the oracle.  Skipped when the box has fewer than two GPUs;
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
from dgraph_b200.shard import run_sharded_pairs
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
