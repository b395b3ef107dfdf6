"""tests/gen.encode_pack_np (the numpy pack generator bench.py uses for its inputs) against the oracle's
restatement of codec.Encode, byte for byte -- two independent writings of the same format."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
import gen  # noqa: E402


def check(orc, uids, bs):
    got = gen.encode_pack_np(uids, bs)
    want = orc.encode(uids, bs)
    if want.is_nil:
        assert got is None
        return
    b, base, num, doff, deltas = got
    assert b == want.block_size
    assert np.array_equal(base, want.base), "base"
    assert np.array_equal(num, want.num_uids), "num_uids"
    assert np.array_equal(doff, want.delta_off), "delta_off"
    assert np.array_equal(deltas, want.deltas), "deltas"


def test_numpy_encoder_matches_oracle(orc):
    rng = np.random.default_rng(1)
    check(orc, np.zeros(0, np.uint64), 256)
    for bs in (0, 1, 2, 5, 10, 100, 256, 257, 1000):
        for n in (1, 2, 4, 5, 9, 255, 256, 257, 1000, 5003):
            check(orc, gen.get_uids(rng, n), bs)
            check(orc, gen.zipf_gaps(rng, n), bs)
    # every delta width, MSB splits, the top of the range
    for shift in (7, 8, 15, 16, 23, 24, 31):
        g = rng.integers(1, 1 << shift, 4000, dtype=np.uint64) + np.uint64(1 << shift)
        check(orc, np.cumsum(g, dtype=np.uint64), 256)
    wide = np.cumsum(rng.integers(1, 1 << 31, 20000, dtype=np.uint64), dtype=np.uint64)
    check(orc, wide, 256)
    check(orc, wide, 10)
    top = np.sort(rng.integers(2**64 - 2**40, 2**64 - 1, 3000, dtype=np.uint64))
    check(orc, np.unique(top), 256)
    check(orc, gen.with_dups(rng, 3000, 2000), 256)       # duplicates: zero deltas


def test_parallel_encoder_matches_oracle(orc):
    rng = np.random.default_rng(2)
    for u, bs, cb in ((gen.zipf_gaps(rng, 300_000), 256, 100),
                      (np.cumsum(rng.integers(1, 1 << 28, 200_000, dtype=np.uint64), dtype=np.uint64), 10, 500),
                      (gen.get_uids(rng, 100_000), 256, 7)):
        got = gen.encode_pack_np_parallel(u, bs, threads=4, chunk_blocks=cb)
        want = orc.encode(u, bs)
        assert np.array_equal(got[1], want.base) and np.array_equal(got[2], want.num_uids)
        assert np.array_equal(got[3], want.delta_off) and np.array_equal(got[4], want.deltas)
