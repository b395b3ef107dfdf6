"""Oracle restatement of algo/packed.go (oracle/packed.py) against EVERY known-answer case of the
reference's algo/packed_test.go (cases cited by line), plus algo/heap_test.go's TestPush.

Like the reference, every pack is built with BlockSize 5 (packed_test.go:18-25 newUidPack) so that the
cases cross block boundaries; newUidPack of no uids is the nil pack (Encoder.Done without Add).
SURVEY.md 8(f) row 2: the packed operations are a "next" row -- oracle only for now.
"""
import numpy as np
import pytest

from oracle import packed as pk

MAX_U64 = 2**64 - 1


def new_uid_pack(orc, data):
    return orc.encode(np.asarray(data, dtype=np.uint64), 5)


def dec(orc, pack):
    """codec.Decode(pack, 0): nil pack -> empty."""
    if pack is None:
        return []
    return orc.decode(pack, 0).tolist()


# ---- MergeSortedPacked: packed_test.go:27-118 --------------------------------------------------------
MERGE_PACKED = [
    ([[55]], [55]),                                                                    # :27-32
    ([[1, 3, 6, 8, 10], [2, 4, 5, 7, 15]], [1, 2, 3, 4, 5, 6, 7, 8, 10, 15]),          # :42-55
    ([[1, 3, 6, 8, 10], []], [1, 3, 6, 8, 10]),                                        # :57-63
    ([[], [1, 3, 6, 8, 10]], [1, 3, 6, 8, 10]),                                        # :65-71
    ([[], []], []),                                                                    # :73-79
    ([[11, 13, 16, 18, 20], [12, 14, 15, 15, 16, 16, 17, 25], [1, 2]],
     [1, 2, 11, 12, 13, 14, 15, 16, 17, 18, 20, 25]),                                  # :81-90
    ([[5, 6, 7], [3, 4], [1, 2], []], [1, 2, 3, 4, 5, 6, 7]),                          # :92-100
    ([], []),                                                                          # :102-105
    ([[1, 1, 1]], [1]),                                                                # :107-112
    ([[1, 2, 3, 3, 6], [4, 8, 9]], [1, 2, 3, 4, 6, 8, 9]),                             # :114-120
]


@pytest.mark.parametrize("lists,want", MERGE_PACKED)
def test_merge_sorted_packed(orc, lists, want):
    packs = [new_uid_pack(orc, l) for l in lists]
    for l, p in zip(lists, packs):
        assert dec(orc, p) == l  # the inputs round-trip (packed_test.go:47-48)
    assert dec(orc, pk.merge_sorted_packed(packs)) == want
    # same answer as the unpacked MergeSorted on the decoded lists
    assert orc.merge_sorted([np.asarray(l, dtype=np.uint64) for l in lists]).tolist() == want


# ---- IntersectWithLinPacked: packed_test.go:122-184 -------------------------------------------------
INTERSECT_PACKED = [
    ([1, 2, 3], [], []),                                                               # :122-127
    ([1, 2, 3], [1, 2, 3, 4, 5], [1, 2, 3]),                                           # :129-134
    ([1, 2, 3], [2], [2]),                                                             # :136-141
    ([1, 2, 3], [0, 5], []),                                                           # :143-148
    ([1, 2, 3], [3, 5], [3]),                                                          # :150-155
    ([1, 2, 3, 4, 5, 6, 7, 9], [1, 3, 5, 7, 8, 9], [1, 3, 5, 7, 9]),                   # :157-162
    ([1, 1, 2, 3], [1, 2], [1, 2]),                                                    # :164-169 DupFirst
    ([1, 1, 2, 3, 5], [1, 1, 2, 4], [1, 1, 2]),                                        # :171-176 DupBoth
    ([1, 2, 3, 5], [1, 1, 2, 4], [1, 2]),                                              # :178-183 DupSecond
]


@pytest.mark.parametrize("u,v,want", INTERSECT_PACKED)
def test_intersect_with_lin_packed(orc, u, v, want):
    got = dec(orc, pk.intersect_with_lin_packed(new_uid_pack(orc, u), new_uid_pack(orc, v)))
    assert got == want
    # multiset-min, like IntersectWithLin on the plain lists (uidlist.go:170-193)
    assert orc.intersect_with_lin(np.asarray(u, dtype=np.uint64), np.asarray(v, dtype=np.uint64)).tolist() == want


# ---- IntersectSortedPacked: packed_test.go:186-246 ---------------------------------------------------
INTERSECT_SORTED_PACKED = [
    ([[1, 2, 3], [2, 3, 4, 5]], [2, 3]),                                               # :186-192
    ([[1, 2, 3]], [1, 2, 3]),                                                          # :194-199
    ([], []),                                                                          # :201-204
    ([[100, 101]], [100, 101]),                                                        # :206-211
    ([[1, 2, 3], [2, 3, 4, 5], [4, 5, 6]], []),                                        # :213-220
    ([[10, 12, 13], [2, 3, 4, 13], [4, 5, 6]], []),                                    # :222-229
    ([list(range(1, n + 1)) for n in range(11, 0, -1)], [1]),                          # :231-246
]


@pytest.mark.parametrize("lists,want", INTERSECT_SORTED_PACKED)
def test_intersect_sorted_packed(orc, lists, want):
    assert dec(orc, pk.intersect_sorted_packed([new_uid_pack(orc, l) for l in lists])) == want


def test_intersect_sorted_packed_shadowing_quirk(orc):
    """packed.go:129 re-declares `out` inside the loop: only the two SHORTEST lists are intersected.
    The reference's own cases never notice; this one does, and the restatement keeps the quirk."""
    lists = [[1, 2, 3, 4], [1, 2, 3, 4, 5], [2, 3, 9, 10, 11, 12]]
    got = dec(orc, pk.intersect_sorted_packed([new_uid_pack(orc, l) for l in lists]))
    assert got == [1, 2, 3, 4]                       # ls[0] ∩ ls[1], NOT [2, 3]
    assert orc.intersect_sorted([np.asarray(l, dtype=np.uint64) for l in lists]).tolist() == [2, 3]


# ---- DifferencePacked: packed_test.go:248-308 -------------------------------------------------------
DIFF_PACKED = [
    ([1, 2, 3], [1], [2, 3]),                                                          # :248-255
    ([1, 2, 3], [2], [1, 3]),                                                          # :257-264
    ([1, 2, 3], [3], [1, 2]),                                                          # :266-273
    ([1, 2, 3], [], [1, 2, 3]),                                                        # :275-282
    ([], [1, 2], []),                                                                  # :284-291
    ([1, 2, 3], [2, 3, 4, 5], [1]),                                                    # :293-300 SubSorted1
    ([10, 12, 13], [2, 3, 4, 13], [10, 12]),                                           # :302-309 SubSorted6
]


@pytest.mark.parametrize("u,v,want", DIFF_PACKED)
def test_difference_packed(orc, u, v, want):
    assert dec(orc, pk.difference_packed(new_uid_pack(orc, u), new_uid_pack(orc, v))) == want
    assert orc.difference(np.asarray(u, dtype=np.uint64), np.asarray(v, dtype=np.uint64)).tolist() == want


# ---- IndexOfPacked: packed_test.go:311-339 ------------------------------------------------------------
def test_index_of_packed_1(orc):
    pack = orc.encode(np.arange(1000, dtype=np.uint64), 10)
    for i in range(1000):
        assert pk.index_of_packed(pack, i) == i
    assert pk.index_of_packed(pack, 1000) == -1


def test_index_of_packed_2(orc):
    pack = orc.encode(np.arange(100, dtype=np.uint64), 10)
    for uid in (100, 101, 1000, MAX_U64):
        assert pk.index_of_packed(pack, uid) == -1


def test_index_of_packed_3(orc):
    assert pk.index_of_packed(None, 0) == -1
    assert pk.index_of_packed(None, MAX_U64) == -1


def test_index_of_packed_matches_index_of(orc):
    """IndexOfPacked on Encode(l) == IndexOf on l (uidlist.go:546-552) for hits and misses."""
    rng = np.random.default_rng(5)
    l = np.unique(rng.integers(0, 5000, 700, dtype=np.uint64))
    pack = orc.encode(l, 7)
    for uid in rng.integers(0, 5100, 400, dtype=np.uint64).tolist() + l[:50].tolist():
        assert pk.index_of_packed(pack, int(uid)) == orc.index_of(l, int(uid))


# ---- ApplyFilterPacked: packed_test.go:341-346 ----------------------------------------------------------
def test_apply_filter_uint_packed(orc):
    l = [1, 2, 3, 4, 5, 6, 7, 8, 9]
    res = pk.apply_filter_packed(new_uid_pack(orc, l), lambda a, idx: l[idx] % 2 == 1)
    assert dec(orc, res) == [1, 3, 5, 7, 9]


# ---- reference behaviour beyond its own test cases ---------------------------------------------------------
def test_difference_packed_multi_block_quirks(orc):
    """DifferencePacked (packed.go:140-219) is only correct while v fits one block per u block; the
    restatement keeps what the Go code does (traced by hand), it does not fix it:
      * :212-215 appends the rest of the current u block as soon as the current v block is exhausted, without
        looking at v's next block;
      * once v's decoder runs off its last block, `len(vuids) == 0` (:161) ends the loop and the remaining
        u blocks are dropped.
    The reference's cases (3-4 uids, one block each) never reach either."""
    u = list(range(3, 8))                            # one block [3..7]
    v = list(range(1, 8))                            # blocks [1..5] [6, 7]
    assert dec(orc, pk.difference_packed(new_uid_pack(orc, u), new_uid_pack(orc, v))) == [6, 7]
    assert orc.difference(np.asarray(u, dtype=np.uint64), np.asarray(v, dtype=np.uint64)).tolist() == []
    u = list(range(1, 21))                           # four blocks
    v = [1]
    assert dec(orc, pk.difference_packed(new_uid_pack(orc, u), new_uid_pack(orc, v))) == list(range(2, 11))
    assert orc.difference(np.asarray(u, dtype=np.uint64), np.asarray(v, dtype=np.uint64)).tolist() == list(range(2, 21))


def test_packed_ops_agree_with_plain_ops(orc):
    """Encode -> packed op -> Decode equals the plain-list op on duplicate-free random inputs at several
    block sizes (the reference's packed tests are exactly this shape with BlockSize 5).  DifferencePacked is
    compared only when each side is a single block (see the quirks above)."""
    rng = np.random.default_rng(11)
    for bs in (1, 2, 5, 16, 256):
        for _ in range(6):
            u = np.unique(rng.integers(0, 3000, int(rng.integers(1, 400)), dtype=np.uint64))
            v = np.unique(rng.integers(0, 3000, int(rng.integers(1, 400)), dtype=np.uint64))
            w = np.unique(rng.integers(0, 3000, int(rng.integers(1, 400)), dtype=np.uint64))
            pu, pv, pw = orc.encode(u, bs), orc.encode(v, bs), orc.encode(w, bs)
            assert dec(orc, pk.intersect_with_lin_packed(pu, pv)) == orc.intersect_with_lin(u, v).tolist()
            assert dec(orc, pk.merge_sorted_packed([pu, pv, pw])) == orc.merge_sorted([u, v, w]).tolist()
            if pu.nblocks == 1 and pv.nblocks == 1:
                assert dec(orc, pk.difference_packed(pu, pv)) == orc.difference(u, v).tolist()
    for _ in range(20):  # single-block differences
        u = np.unique(rng.integers(0, 500, int(rng.integers(1, 200)), dtype=np.uint64))
        v = np.unique(rng.integers(0, 500, int(rng.integers(1, 200)), dtype=np.uint64))
        pu, pv = orc.encode(u, 256), orc.encode(v, 256)
        assert pu.nblocks == 1 and pv.nblocks == 1
        assert dec(orc, pk.difference_packed(pu, pv)) == orc.difference(u, v).tolist()


# ---- algo/heap_test.go:14-49 TestPush ------------------------------------------------------------------
def test_heap_push_order(orc):
    """TestPush checks that uint64Heap is a min-heap on elem.val (heap.go:12-37) through Push / Fix / Pop.
    The oracle's heap is internal to MergeSorted, so the same property is checked through it: single-element
    lists come out in pop order, whatever order they were pushed in (5, 3, 4 -> 3, 4, 5 as in :18-26)."""
    got = orc.merge_sorted([np.array([5], dtype=np.uint64), np.array([3], dtype=np.uint64),
                            np.array([4], dtype=np.uint64)], single_heap=True)
    assert got.tolist() == [3, 4, 5]
    got = orc.merge_sorted([np.array([10], dtype=np.uint64), np.array([11], dtype=np.uint64),
                            np.array([1], dtype=np.uint64), np.array([5], dtype=np.uint64)], single_heap=True)
    assert got.tolist() == [1, 5, 10, 11]
