"""Pins oracle/ against the reference's own tests for codec/ (codec/codec_test.go).

The reference tests only round-trip (no Deltas byte vectors exist in the tree),
so the group-varint byte layout additionally gets hand-derived known answers
from the published go-groupvarint format (see oracle.h PARITY NOTE).
"""
import numpy as np
import pytest

MAXU64 = 0xFFFFFFFFFFFFFFFF


def get_uids(rng, size):
    """getUids, codec/codec_test.go:26-35: start rand.Intn(100), gaps rand.Intn(33)."""
    if size == 0:
        return np.zeros(0, np.uint64)
    gaps = rng.integers(0, 33, size, dtype=np.uint64)
    gaps[0] = rng.integers(0, 100)
    return np.cumsum(gaps, dtype=np.uint64)


# ---- group varint known answers (format restated in oracle.c) ---------------

GV_KATS = [
    ([0, 0, 0, 0], bytes([0x00, 0, 0, 0, 0])),
    ([1, 2, 3, 4], bytes([0x00, 1, 2, 3, 4])),
    ([255, 256, 65535, 65536], bytes([0b10010100, 0xFF, 0x00, 0x01, 0xFF, 0xFF, 0x00, 0x00, 0x01])),
    ([1, 256, 65536, 16777216],
     bytes([0xE4, 0x01, 0x00, 0x01, 0x00, 0x00, 0x01, 0x00, 0x00, 0x00, 0x01])),
    ([0xFFFFFFFF] * 4, bytes([0xFF] + [0xFF] * 16)),
    ([0x12345678, 0x9A, 0xBCDE, 0xF01234],
     bytes([0b10010011, 0x78, 0x56, 0x34, 0x12, 0x9A, 0xDE, 0xBC, 0x34, 0x12, 0xF0])),
]


@pytest.mark.parametrize("vals,enc", GV_KATS)
def test_groupvarint_kat(orc, vals, enc):
    assert orc.gv_encode4(vals) == enc
    assert orc.gv_decode4(enc) == vals
    assert orc.gv_bytes_used(enc[0]) == len(enc)


def test_groupvarint_bytes_used_table(orc):
    for tag in range(256):
        want = 1 + sum(((tag >> (2 * j)) & 3) + 1 for j in range(4))
        assert orc.gv_bytes_used(tag) == want
    rng = np.random.default_rng(1)
    for _ in range(2000):
        widths = rng.integers(1, 5, 4)
        vals = [int(rng.integers(0, 2 ** (8 * w))) for w in widths]
        enc = orc.gv_encode4(vals)
        assert orc.gv_decode4(enc) == vals and orc.gv_bytes_used(enc[0]) == len(enc)


# ---- block layout rules (codec/codec.go:57-127) -----------------------------

def test_block_rules(orc):
    # 1-UID block still carries one all-zero group (5 bytes): loop body runs before the len<=4 test (:76-96)
    p = orc.encode([7], 256)
    assert p.nblocks == 1 and p.num_uids.tolist() == [1] and p.base.tolist() == [7]
    assert p.block_deltas(0) == bytes(5)
    # groups = max(1, ceil((n-1)/4))
    for n in range(1, 40):
        p = orc.encode(np.arange(10, 10 + n, dtype=np.uint64), 256)
        assert p.nblocks == 1
        groups = max(1, -(-(n - 1) // 4))
        assert len(p.block_deltas(0)) == 5 * groups  # all deltas are 1 => 1 byte each
    # BlockSize checked after append with >= : BlockSize 0 and 1 => one uid per block
    for bs in (0, 1):
        p = orc.encode([1, 2, 3], bs)
        assert p.nblocks == 3 and p.num_uids.tolist() == [1, 1, 1]
    # a new block starts when the upper 32 bits change (match32MSB :469)
    p = orc.encode([1, 2, (1 << 32) + 5, (1 << 32) + 6, (2 << 32)], 256)
    assert p.base.tolist() == [1, (1 << 32) + 5, 2 << 32] and p.num_uids.tolist() == [2, 2, 1]
    # exactly block_size uids per full block
    p = orc.encode(np.arange(1, 26, dtype=np.uint64), 10)
    assert p.num_uids.tolist() == [10, 10, 5] and p.base.tolist() == [1, 11, 21]
    # empty input: Encoder.Done() with no Add returns a nil pack
    p = orc.encode([], 128)
    assert p.is_nil and orc.approx_len(p) == 0 and orc.exact_len(p) == 0
    assert orc.decode(p, 0).size == 0


def test_uid_pack_roundtrip(orc):
    # TestUidPack :37-58 (13 random sizes, BlockSize 256); sizes scaled to keep the CPU suite fast
    rng = np.random.default_rng(2)
    sizes = [0, 1, 2, 255, 256, 257, 1023] + [int(rng.integers(0, 300000)) for _ in range(6)]
    for size in sizes:
        expected = get_uids(rng, size)
        pack = orc.encode(expected, 256)
        assert orc.exact_len(pack) == expected.size
        assert np.array_equal(orc.decode(pack, 0), expected)


def make_seek_pack(orc):
    # TestSeek :113-123: 0,10,...,10000 with BlockSize 10
    return orc.encode(np.arange(0, 10001, 10, dtype=np.uint64), 10)


SEEK_TABLE = [
    (0, 0, 0, False), (0, 0, 1, False),
    (100, 100, 0, False), (100, 110, 1, False),
    (1000, 1000, 0, False), (1000, 1010, 1, False),
    (1999, 2000, 0, False), (1999, 2000, 1, False),
    (1101, 1110, 0, False), (1101, 1110, 1, False),
    (10000, 10000, 0, False), (9999, 10000, 1, False),
    (10001, None, 0, True), (10001, None, 1, True),
    (MAXU64, None, 0, True), (MAXU64, None, 1, True),
]


def test_seek_table(orc):
    # TestSeek :125-151
    pack = make_seek_pack(orc)
    dec = orc.Decoder(pack)
    for uid, out, whence, empty in SEEK_TABLE:
        uids = dec.seek(uid, whence)
        if empty:
            assert uids.size == 0, (uid, whence)
        else:
            assert int(uids[0]) == out, (uid, whence)
    # :153-157
    dec.block_idx = 0
    for i in range(100, 10000, 100):
        assert i in dec.linear_seek(i).tolist()


def test_linear_seek(orc):
    # TestLinearSeek :160-187
    N = 10001
    pack = make_seek_pack(orc)
    dec = orc.Decoder(pack)
    for i in range(0, 2 * N, 10):
        uids = dec.linear_seek(i).tolist()
        assert (i in uids) == (i < N)
    for i in range(0, 9990, 10):  # blockIdx now points at the last block
        assert i not in dec.linear_seek(i).tolist()


def test_decoder_seek_and_decode_suffix(orc):
    # TestDecoder :190-217
    N = 10001
    expected = np.arange(3, N, 3, dtype=np.uint64)
    pack = orc.encode(expected, 10)
    dec = orc.Decoder(pack)
    for i in range(3, N, 3 * 7):  # every 7th to bound runtime; each hits all three probes
        for d in (0, 1, 2):
            assert int(dec.seek(i - d, orc.SEEK_START)[0]) == i
        start = i // 3 - 1
        assert np.array_equal(orc.decode(pack, i), expected[start:])
    assert orc.decode(pack, N + 5).size == 0
    assert np.array_equal(orc.decode(pack, 1), expected)


def test_encoding_msb_splits(orc):
    # TestEncoding :306-334
    big = [0xF000000000000000, 0xF00F000000000000, 0x00F00F0000000000, 0x000F0F0000000000, 0x0F0F0F0F00000000]
    rng = np.random.default_rng(3)
    for n in (0, 1, 2, 3, 5, 13, 18, 100, 99, 98):
        ints = np.zeros(n, dtype=np.uint64)
        for i in range(min(50, n)):
            ints[i] = rng.integers(0, 2**32)
        for i in range(50, n):
            ints[i] = int(rng.integers(0, 2**32)) + big[int(rng.integers(0, 5))]
        ints.sort()
        pack = orc.encode(ints, 256)
        assert np.array_equal(orc.decode(pack, 0), ints)


def test_seek_to_block(orc):
    pack = make_seek_pack(orc)
    dec = orc.Decoder(pack)
    dec.seek(0, orc.SEEK_START)
    # block containing 1234 is [1200..1290]; not truncated
    uids = dec.seek_to_block(1234, orc.SEEK_CURRENT)
    assert uids.tolist() == list(range(1200, 1300, 10))
    # exact base hit returns that block for SeekStart
    assert dec.seek_to_block(1300, orc.SEEK_START).tolist() == list(range(1300, 1400, 10))
    # past the block's last uid moves to the next block
    assert int(dec.seek_to_block(1395, orc.SEEK_CURRENT)[0]) == 1400
    assert dec.seek_to_block(20000, orc.SEEK_CURRENT).size == 0


def test_approx_len(orc):
    pack = make_seek_pack(orc)
    assert orc.approx_len(pack) == 101 * 10 and orc.exact_len(pack) == 1001
    dec = orc.Decoder(pack)
    dec.seek(5000, orc.SEEK_START)
    assert dec.approx_len() == 10 * (101 - 50)


def test_golden_clustered_fixture(orc):
    """tests/golden/clustered_head.npz: 4096 values (offset 458752) of the reference's own
    codec/benchmark/clustered1M.bin.gz plus pack geometry recorded by
    tests/golden/make_golden.py (regression pin for the encoder)."""
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "clustered_head.npz")
    g = np.load(path)
    uids = g["uids"]
    pack = orc.encode(uids, 256)
    assert pack.nblocks == int(g["nblocks"])
    assert np.array_equal(pack.delta_off, g["delta_off"])
    assert np.array_equal(pack.deltas, g["deltas"])
    assert np.array_equal(orc.decode(pack, 0), uids)
