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


def encode_pack_np(uids, block_size: int = 256):
    """codec.Encode (codec/codec.go:57-136, 393-399) vectorised in numpy: input generator for packs.

    Returns (block_size, base u64[nb], num_uids u32[nb], delta_off u64[nb+1], deltas u8[...]) -- the arrays of
    dgx_pack_view -- or None for an empty list (Encoder.Done without Add returns a nil pack).
    Block rule (:116-120): a new block starts when the upper 32 bits change or the block holds BlockSize
    uids (BlockSize 0 behaves like 1).  Payload (:76-96): uint32 deltas in groups of four, the last group
    zero-padded, at least one group per block; a group is a tag byte (value j's length-1 in bits 2j+1:2j)
    followed by the four values, little-endian, 1-4 bytes each (go-groupvarint Encode4).
    Independent of oracle/ (tests/test_gen_encoder.py checks it against the oracle byte for byte)."""
    u = np.ascontiguousarray(np.asarray(uids, dtype=np.uint64))
    n = u.size
    if n == 0:
        return None
    B = max(int(block_size), 1)
    idx = np.arange(n, dtype=np.int64)
    msb = u >> np.uint64(32)
    seg = np.ones(n, dtype=bool)
    seg[1:] = msb[1:] != msb[:-1]
    seg_start = np.maximum.accumulate(np.where(seg, idx, 0))
    bstart = seg | (((idx - seg_start) % B) == 0)
    starts = np.nonzero(bstart)[0]
    nb = starts.size
    num = np.diff(np.append(starts, n)).astype(np.int64)
    base = u[starts]
    ngroups = np.maximum(1, (num - 1 + 3) // 4)
    grp_off = np.concatenate([[0], np.cumsum(ngroups)]).astype(np.int64)
    ng = int(grp_off[-1])
    D = np.zeros((ng, 4), dtype=np.uint32)
    nonstart = ~bstart
    if nonstart.any():
        blk = np.cumsum(bstart) - 1                      # block of every uid
        pos = idx - starts[blk]                          # position inside its block
        d = (u[1:] - u[:-1]).astype(np.uint32)           # same MSB inside a block: fits 32 bits
        sel = np.nonzero(nonstart)[0]
        g = grp_off[blk[sel]] + (pos[sel] - 1) // 4
        D[g, (pos[sel] - 1) % 4] = d[sel - 1]
    lens = 1 + (D >= (1 << 8)).astype(np.int64) + (D >= (1 << 16)) + (D >= (1 << 24))
    tags = ((lens[:, 0] - 1) | ((lens[:, 1] - 1) << 2) | ((lens[:, 2] - 1) << 4) | ((lens[:, 3] - 1) << 6)).astype(np.uint8)
    gbytes = 1 + lens.sum(axis=1)
    goff = np.concatenate([[0], np.cumsum(gbytes)]).astype(np.int64)
    out = np.zeros(int(goff[-1]), dtype=np.uint8)
    out[goff[:-1]] = tags
    voff = goff[:-1] + 1
    for j in range(4):
        for t in range(4):
            m = lens[:, j] > t
            out[voff[m] + t] = ((D[m, j] >> np.uint32(8 * t)) & np.uint32(0xFF)).astype(np.uint8)
        voff = voff + lens[:, j]
    delta_off = goff[grp_off].astype(np.uint64)
    return int(block_size), base, num.astype(np.uint32), delta_off, out


def encode_pack_np_parallel(uids, block_size: int = 256, threads: int = 16, chunk_blocks: int = 16384):
    """encode_pack_np over pieces that are independent by construction -- cuts only where a block starts anyway
    (at a change of the upper 32 bits, or a whole number of blocks after one) -- on a thread pool."""
    from concurrent.futures import ThreadPoolExecutor

    u = np.ascontiguousarray(np.asarray(uids, dtype=np.uint64))
    n = u.size
    B = max(int(block_size), 1)
    if n <= B * chunk_blocks:
        return encode_pack_np(u, block_size)
    msb = u >> np.uint64(32)
    seg_starts = np.concatenate([[0], np.nonzero(msb[1:] != msb[:-1])[0] + 1, [n]])
    cuts = []
    for s, e in zip(seg_starts[:-1], seg_starts[1:]):
        cuts.extend(range(int(s), int(e), B * chunk_blocks))
    cuts.append(n)
    with ThreadPoolExecutor(max_workers=threads) as ex:
        parts = list(ex.map(lambda se: encode_pack_np(u[se[0]: se[1]], block_size), zip(cuts[:-1], cuts[1:])))
    base = np.concatenate([p[1] for p in parts])
    num = np.concatenate([p[2] for p in parts])
    deltas = np.concatenate([p[4] for p in parts])
    shift = np.concatenate([[0], np.cumsum([p[4].size for p in parts])]).astype(np.uint64)
    doff = np.concatenate([p[3][:-1] + shift[i] for i, p in enumerate(parts)] + [shift[-1:]])
    return int(block_size), base, num, doff.astype(np.uint64), deltas
