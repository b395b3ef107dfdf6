"""Generates tests/golden/*.npz.  Run in the authoring container only (reads
/root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

clustered_head.npz: 4096 values (offset 458752: mixed 1/2/3-byte deltas) of the reference's codec benchmark
fixture codec/benchmark/clustered1M.bin.gz (format: benchmark.go:59-88, LE u32
count then LE u32 values), with the pack our oracle encoder produces for them at
BlockSize 256.  The Go reference cannot be executed here, so the recorded bytes
pin the oracle against regressions, not against a Go run.
"""
import gzip
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as orc  # noqa: E402

raw = gzip.open("/root/reference/codec/benchmark/clustered1M.bin.gz", "rb").read()
n = int(np.frombuffer(raw[:4], dtype="<u4")[0])
vals = np.frombuffer(raw[4:4 + 4 * n], dtype="<u4").astype(np.uint64)
assert n == 1048576 and np.all(np.diff(vals.astype(np.int64)) > 0)
head = vals[458752:458752 + 4096].copy()  # slice with 1-, 2- and 3-byte deltas
pack = orc.encode(head, 256)
np.savez_compressed(
    os.path.join(os.path.dirname(__file__), "clustered_head.npz"),
    uids=head, nblocks=np.int64(pack.nblocks), delta_off=pack.delta_off, deltas=pack.deltas,
    base=pack.base, num_uids=pack.num_uids,
)
print("clustered_head.npz:", head.size, "uids", pack.nblocks, "blocks", int(pack.delta_off[-1]), "delta bytes")
