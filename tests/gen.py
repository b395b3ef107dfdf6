"""Seeded synthetic posting-list generators shared by tests and bench.py
(SURVEY.md section 8d).  All lists are sorted uint64; unique unless stated."""
from __future__ import annotations

import numpy as np


def uniform_unique(rng, n, hi):
    """n sorted unique values uniform in [0, hi)."""
    if n == 0:
        return np.zeros(0, np.uint64)
    x = np.unique(rng.integers(0, hi, int(n * 1.2) + 16, dtype=np.uint64))
    while x.size < n:
        x = np.unique(np.concatenate([x, rng.integers(0, hi, n, dtype=np.uint64)]))
    idx = np.sort(rng.choice(x.size, n, replace=False))
    return x[idx]


def zipf_gaps(rng, n, a=1.5, clip=1 << 20, start=0):
    """'Zipf-gap' list: uid[i] = uid[i-1] + min(zipf(a), clip)  (strictly increasing)."""
    if n == 0:
        return np.zeros(0, np.uint64)
    g = np.minimum(rng.zipf(a, n), clip).astype(np.uint64)
    return np.cumsum(g, dtype=np.uint64) + np.uint64(start)


def thin(rng, master, p):
    """independent thinning of a master list with keep-probability p"""
    return master[rng.random(master.size) < p]


def with_dups(rng, n, hi):
    """n sorted values with duplicates, uniform in [0, hi) (reference bench generator shape,
    algo/uidlist_test.go:468-494: sorted rand.Int63n(limit), not de-duplicated)."""
    return np.sort(rng.integers(0, hi, n, dtype=np.uint64))


def fill_nums(rng, n1, n2):
    """fillNums (algo/uidlist_test.go:583-605): n1 common + n2 private random u64 each."""
    common = rng.integers(0, 2**64, n1, dtype=np.uint64)
    block = np.concatenate([common, rng.integers(0, 2**64, n2, dtype=np.uint64)])
    other = np.concatenate([common, rng.integers(0, 2**64, n2, dtype=np.uint64)])
    return np.sort(common), np.sort(block), np.sort(other)


def get_uids(rng, size):
    """getUids (codec/codec_test.go:26-35): start rand.Intn(100), gaps rand.Intn(33)."""
    if size == 0:
        return np.zeros(0, np.uint64)
    gaps = rng.integers(0, 33, size, dtype=np.uint64)
    gaps[0] = rng.integers(0, 100)
    return np.cumsum(gaps, dtype=np.uint64)
