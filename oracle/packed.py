"""CPU restatement of dgraph's algo/packed.go: the set operations on compressed lists (pb.UidPack).

TEST INFRASTRUCTURE ONLY (see oracle/pyoracle.py).  SURVEY.md section 8(f) row 2 marks these functions
"next": the product does not ship packed entry points yet; this module pins their behaviour -- including
the reference's quirks -- against every known-answer case of algo/packed_test.go, so a later GPU version
has an oracle.  Plain Python loops over the oracle's Decoder / Encoder: small inputs only.

Each function follows the Go code decision by decision (block-wise decoders, the same loop exits);
citations are /root/reference/algo/packed.go:line.
"""
from __future__ import annotations

import heapq
from typing import Callable, List, Optional

import numpy as np

from . import pyoracle as orc


def _new_decoder(pack: "orc.Pack") -> "orc.Decoder":
    """codec.NewDecoder (codec/codec.go:146-152): Decoder{Pack} followed by Seek(0, SeekStart)."""
    d = orc.Decoder(pack)
    d.seek(0, orc.SEEK_START)
    return d


class _Encoder:
    """codec.Encoder as the packed functions use it: Add(uid)... Done() (codec/codec.go:57-136).
    Blocks are cut by the rules of Encoder.Add / packBlock, which orc.encode restates; Done() on an
    encoder that never saw a uid returns the nil pack."""

    def __init__(self, block_size: int):
        self.block_size = int(block_size)
        self.uids: List[int] = []

    def add(self, uid: int) -> None:
        self.uids.append(int(uid))

    def done(self) -> "orc.Pack":
        return orc.encode(np.asarray(self.uids, dtype=np.uint64), self.block_size)


def copy_uid_pack(pack: Optional["orc.Pack"]) -> Optional["orc.Pack"]:
    """codec.CopyUidPack (codec/codec.go:474-492): nil stays nil, otherwise a deep copy."""
    if pack is None or pack.is_nil:
        return pack
    return orc.encode(orc.decode(pack, 0), pack.block_size)


def apply_filter_packed(u: "orc.Pack", f: Callable[[int, int], bool]) -> "orc.Pack":
    """ApplyFilterPacked (packed.go:17-33): keep uid number `index` iff f(uid, index)."""
    index = 0
    dec = _new_decoder(u)
    enc = _Encoder(u.block_size)
    while dec.valid():
        for uid in dec.uids():
            if f(int(uid), index):
                enc.add(uid)
            index += 1
        dec.next()
    return enc.done()


def intersect_with_lin_packed(u: Optional["orc.Pack"], v: Optional["orc.Pack"]) -> Optional["orc.Pack"]:
    """IntersectWithLinPacked (packed.go:36-97): two-pointer walk, one decoded block of each side at a time."""
    if u is None or v is None or u.is_nil or v.is_nil:
        # Go: u == nil || v == nil -> nil.  A pack built from no uids IS nil (Encoder.Done), :37-39.
        return None
    u_dec = _new_decoder(u)
    uuids = u_dec.uids()
    v_dec = _new_decoder(v)
    vuids = v_dec.uids()
    u_idx = v_idx = 0
    result = _Encoder(u.block_size)
    while True:
        if len(uuids) == 0 or len(vuids) == 0:      # :50-52
            break
        if u_idx == len(uuids):                     # :55-63
            if u_dec.valid():
                uuids = u_dec.next()
                u_idx = 0
            else:
                break
        if v_idx == len(vuids):                     # :64-71
            if v_dec.valid():
                vuids = v_dec.next()
                v_idx = 0
            else:
                break
        u_len, v_len = len(uuids), len(vuids)
        while u_idx < u_len and v_idx < v_len:      # :76-92
            uid, vid = int(uuids[u_idx]), int(vuids[v_idx])
            if uid > vid:
                v_idx += 1
                while v_idx < v_len and int(vuids[v_idx]) < uid:
                    v_idx += 1
            elif uid == vid:
                result.add(uid)
                v_idx += 1
                u_idx += 1
            else:
                u_idx += 1
                while u_idx < u_len and int(uuids[u_idx]) < vid:
                    u_idx += 1
    return result.done()


def _exact_len(p: Optional["orc.Pack"]) -> int:
    return 0 if p is None else orc.exact_len(p)


def intersect_sorted_packed(lists: List["orc.Pack"]) -> Optional["orc.Pack"]:
    """IntersectSortedPacked (packed.go:107-137).

    Reference quirk kept on purpose: inside the loop `out := IntersectWithLinPacked(out, ls[i].l)`
    declares a NEW variable (:129), so the function returns ls[0] ∩ ls[1] whatever the later lists hold;
    the loop only decides when to stop.  TestIntersectSorted5/6/7Packed pass because the two shortest
    lists already give the final answer."""
    if len(lists) == 0:
        return _Encoder(10).done()                  # :108-111 (an empty encoder: the nil pack)
    ls = sorted(lists, key=_exact_len)              # sort.Slice by length, :119-121 (Python's sort is
    #                                                 stable, Go's is not: ties only permute equal lengths)
    if len(ls) == 1:
        return copy_uid_pack(ls[0])                 # :123-126
    out = intersect_with_lin_packed(ls[0], ls[1])   # :129
    for i in range(2, len(ls)):                     # :131-138
        shadow = intersect_with_lin_packed(out, ls[i])
        if _exact_len(shadow) == 0:
            break
    return out


def difference_packed(u: Optional["orc.Pack"], v: Optional["orc.Pack"]) -> Optional["orc.Pack"]:
    """DifferencePacked (packed.go:140-219)."""
    u_nil = u is None or u.is_nil
    v_nil = v is None or v.is_nil
    if u_nil or v_nil:                              # :141-150
        if v_nil:
            return copy_uid_pack(u)
        return None
    result = _Encoder(u.block_size)
    u_dec = _new_decoder(u)
    uuids = u_dec.uids()
    v_dec = _new_decoder(v)
    vuids = v_dec.uids()
    u_idx = v_idx = 0
    while True:
        if len(uuids) == 0 or len(vuids) == 0:      # :161-163
            break
        if u_idx == len(uuids):                     # :166-174
            if u_dec.valid():
                uuids = u_dec.next()
                u_idx = 0
            else:
                break
        if v_idx == len(vuids):                     # :176-183
            if v_dec.valid():
                vuids = v_dec.next()
                v_idx = 0
            else:
                break
        u_len, v_len = len(uuids), len(vuids)
        while u_idx < u_len and v_idx < v_len:      # :188-210
            uid, vid = int(uuids[u_idx]), int(vuids[v_idx])
            if uid < vid:
                while u_idx < u_len and int(uuids[u_idx]) < vid:
                    result.add(uuids[u_idx])
                    u_idx += 1
            elif uid == vid:
                u_idx += 1
                v_idx += 1
            else:
                v_idx += 1
                while v_idx < v_len and int(vuids[v_idx]) < uid:
                    v_idx += 1
        while u_idx < u_len and v_idx >= v_len:     # :212-215
            result.add(uuids[u_idx])
            u_idx += 1
    return result.done()


def merge_sorted_packed(lists: List[Optional["orc.Pack"]]) -> Optional["orc.Pack"]:
    """MergeSortedPacked (packed.go:222-297): heap of (current value, list), one decoded block per list,
    global de-duplication against the last value written."""
    if len(lists) == 0:
        return None                                 # :223-225
    heap = []                                       # entries: [val, list index, decoder, idx in block, block uids]
    block_size = 0
    for i, l in enumerate(lists):
        if l is None or l.is_nil:
            continue                                # :233-235
        if block_size == 0:
            block_size = l.block_size
        dec = _new_decoder(l)
        block = dec.uids()
        if len(block) == 0:
            continue
        if orc.exact_len(l) > 0:
            heapq.heappush(heap, [int(block[0]), i, dec, 0, block])
    result = _Encoder(block_size)
    empty_result = True
    last = 0
    while heap:                                     # :263-294
        me = heap[0]
        if empty_result or me[0] != last:
            result.add(me[0])
            last = me[0]
            empty_result = False
        dec = me[2]
        last_block = dec.block_idx == dec.pack.nblocks - 1
        if me[3] == len(me[4]) - 1 and last_block:
            heapq.heappop(heap)
            continue
        me[3] += 1
        if me[3] >= len(me[4]):
            me[4] = dec.next()
            me[3] = 0
        me[0] = int(me[4][me[3]])
        heapq.heapreplace(heap, me)                 # heap.Fix(h, 0)
    return result.done()


def index_of_packed(u: Optional["orc.Pack"], uid: int) -> int:
    """IndexOfPacked (packed.go:301-330): position of uid in the whole pack, -1 when absent."""
    if u is None or u.is_nil:
        return -1
    dec = orc.Decoder(u)
    dec.seek(uid, orc.SEEK_START)
    dec.unpack_block()                              # :309-311: Seek may have trimmed the block's head
    uids = dec.uids()
    if len(uids) == 0:
        return -1
    uidx = int(np.searchsorted(uids, np.uint64(uid), side="left"))  # sort.Search(len, uids[i] >= uid)
    # Go indexes uids[uidx] unguarded (:319); Seek only stops on a block whose last uid is >= uid or on
    # the end of the pack (empty uids), so uidx < len(uids) whenever we get here.
    if int(uids[uidx]) != uid:
        return -1
    index = uidx
    num = u.num_uids
    for i in range(dec.block_idx):
        index += int(num[i])
    return index
