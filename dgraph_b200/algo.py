"""Host-side mirror of dgraph's `algo` package over libdgx (sm_100a CUDA).

Same function names, argument meaning and result shapes as
/root/reference/algo/uidlist.go; every body is one C-ABI call into libdgx.so
(include/dgx.h), exactly what the cgo shim in INTEGRATION.md does.  No set
arithmetic happens in Python and there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _lib
from .pb import List, _u64


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _ptr_table(arrs: Sequence[np.ndarray]):
    k = len(arrs)
    ptrs = (C.c_void_p * max(k, 1))(*[a.ctypes.data if a.size else None for a in arrs])
    lens = (C.c_size_t * max(k, 1))(*[a.size for a in arrs])
    return ptrs, lens


def IntersectWith(u: List, v: List, o: List) -> None:
    """algo.IntersectWith(u, v, o) (algo/uidlist.go:142-167): o.Uids = u ∩ v.

    `o` may be `u` (in place).  When o.Uids has room the result is written into
    it (the reference reuses o.Uids[:0]); v is never modified.
    """
    lib = _lib.load()
    uu = np.zeros(0, np.uint64) if u.Uids is None else u.Uids
    vv = np.zeros(0, np.uint64) if v.Uids is None else v.Uids
    cap = min(uu.size, vv.size)
    if o.Uids is not None and o.Uids.size >= cap and o.Uids.flags.writeable:
        out = o.Uids  # dst := o.Uids[:0]
    else:
        out = np.empty(max(cap, 1), dtype=np.uint64)  # make([]uint64, 0, n)
    n = C.c_size_t(0)
    _lib.check(lib.dgx_intersect2(_p(uu), uu.size, _p(vv), vv.size, _p(out), cap, C.byref(n)))
    o.Uids = out[: n.value]


def IntersectSorted(lists: Sequence[List]) -> List:
    """algo.IntersectSorted (algo/uidlist.go:297-329).  No lists -> &pb.List{} (nil Uids)."""
    if len(lists) == 0:
        return List(None)
    lib = _lib.load()
    arrs = [np.zeros(0, np.uint64) if l.Uids is None else l.Uids for l in lists]
    ptrs, lens = _ptr_table(arrs)
    cap = min(a.size for a in arrs)
    out = np.empty(max(cap, 1), dtype=np.uint64)
    n = C.c_size_t(0)
    _lib.check(lib.dgx_intersect_sorted(ptrs, lens, len(arrs), _p(out), cap, C.byref(n)))
    return List(out[: n.value])


def MergeSorted(lists: Sequence[List]) -> List:
    """algo.MergeSorted (algo/uidlist.go:448-542): sorted union, globally de-duplicated."""
    lib = _lib.load()
    arrs = [np.zeros(0, np.uint64) if (l is None or l.Uids is None) else l.Uids for l in lists]
    ptrs, lens = _ptr_table(arrs)
    cap = sum(a.size for a in arrs)
    out = np.empty(max(cap, 1), dtype=np.uint64)
    n = C.c_size_t(0)
    _lib.check(lib.dgx_merge_sorted(ptrs, lens, len(arrs), _p(out), cap, C.byref(n)))
    return List(out[: n.value])


def Difference(u: List, v: List) -> List:
    """algo.Difference (algo/uidlist.go:332-362): u \\ v; nil u or v -> empty non-nil list."""
    if u is None or v is None:
        return List(np.zeros(0, np.uint64))
    lib = _lib.load()
    uu = np.zeros(0, np.uint64) if u.Uids is None else u.Uids
    vv = np.zeros(0, np.uint64) if v.Uids is None else v.Uids
    out = np.empty(max(uu.size, 1), dtype=np.uint64)
    n = C.c_size_t(0)
    _lib.check(lib.dgx_difference(_p(uu), uu.size, _p(vv), vv.size, _p(out), uu.size, C.byref(n)))
    return List(out[: n.value])


def IntersectCompressedWith(pack, afterUID: int, v: List, o: List) -> None:
    """algo.IntersectCompressedWith(pack, afterUID, v, o) (algo/uidlist.go:33-61): o.Uids = v ∩ pack[>= afterUID].

    A nil pack leaves `o` untouched, like the reference (:34-36)."""
    if pack is None:
        return
    from .codec import view_of

    lib = _lib.load()
    vv = np.zeros(0, np.uint64) if v.Uids is None else v.Uids
    out = np.empty(max(vv.size, 1), dtype=np.uint64)
    n = C.c_size_t(0)
    pack = pack.normalized()
    view = view_of(pack)
    _lib.check(lib.dgx_intersect_compressed(C.byref(view), afterUID, _p(vv), vv.size, _p(out), vv.size, C.byref(n)))
    o.Uids = out[: n.value]


def IntersectBatch(a: np.ndarray, a_off: np.ndarray, b: np.ndarray, b_off: np.ndarray):
    """Batched independent IntersectWith in CSR form (dgx_intersect_batch).

    Returns (out, out_off): pair i's result is out[out_off[i]:out_off[i+1]].
    """
    lib = _lib.load()
    a, b = _u64(a), _u64(b)
    a_off, b_off = _u64(a_off), _u64(b_off)
    npairs = a_off.size - 1
    cap = int(np.minimum(np.diff(a_off.astype(np.int64)), np.diff(b_off.astype(np.int64))).sum()) if npairs > 0 else 0
    out = np.empty(max(cap, 1), dtype=np.uint64)
    out_off = np.zeros(npairs + 1, dtype=np.uint64)
    _lib.check(lib.dgx_intersect_batch(_p(a), _p(a_off), _p(b), _p(b_off), npairs, _p(out), _p(out_off), cap))
    return out[: int(out_off[-1])], out_off


def IntersectSortedPacks(packs, keys=None) -> List:
    """algo.IntersectSorted over lists held as UidPacks (dgx_intersect_sorted_packed): the packs cross
    PCIe compressed and are decoded on the device (codec.Decode(p, 0) for each, codec/codec.go:444).

    keys: optional [(key, version)] naming each pack for the HBM-resident cache (0 = anonymous)."""
    from .codec import ExactLen, view_of

    if len(packs) == 0:
        return List(None)
    lib = _lib.load()
    k = len(packs)
    refs = (_lib.PackRef * k)()
    keep = []
    cap = None
    for i, p in enumerate(packs):
        if p is None or p.nblocks == 0:
            refs[i].pack = None
            cap = 0
        else:
            p = p.normalized()
            v = view_of(p)
            keep.append((p, v))
            refs[i].pack = C.pointer(v)
            n = ExactLen(p)
            cap = n if cap is None else min(cap, n)
        refs[i].key, refs[i].version = (keys[i] if keys is not None else (0, 0))
    out = np.empty(max(cap or 0, 1), dtype=np.uint64)
    n = C.c_size_t(0)
    _lib.check(lib.dgx_intersect_sorted_packed(refs, k, _p(out), cap or 0, C.byref(n)))
    return List(out[: n.value])


def IndexOfBatch(u: List, uids) -> np.ndarray:
    """algo.IndexOf(u, uid) for every uid (algo/uidlist.go:546-552): int64 positions, -1 when absent."""
    lib = _lib.load()
    uu = np.zeros(0, np.uint64) if u.Uids is None else u.Uids
    q = _u64(uids)
    idx = np.empty(max(q.size, 1), dtype=np.int64)
    _lib.check(lib.dgx_index_of_batch(_p(uu), uu.size, _p(q), q.size, idx.ctypes.data_as(C.c_void_p)))
    return idx[: q.size]


def IntersectBatchShared(a: np.ndarray, a_off: np.ndarray, b: np.ndarray):
    """Rows a[a_off[i]:a_off[i+1]] each intersected with the one list b (dgx_intersect_batch_shared):
    the `algo.IntersectWith(l, sg.DestUIDs, l)` loop of query/query.go:1425-1438."""
    lib = _lib.load()
    a, b, a_off = _u64(a), _u64(b), _u64(a_off)
    npairs = a_off.size - 1
    cap = int(np.minimum(np.diff(a_off.astype(np.int64)), b.size).sum()) if npairs > 0 else 0
    out = np.empty(max(cap, 1), dtype=np.uint64)
    out_off = np.zeros(npairs + 1, dtype=np.uint64)
    _lib.check(lib.dgx_intersect_batch_shared(_p(a), _p(a_off), npairs, _p(b), b.size, _p(out), _p(out_off), cap))
    return out[: int(out_off[-1])], out_off


def _refs_of(packs, keys=None):
    from .codec import view_of

    k = len(packs)
    refs = (_lib.PackRef * max(k, 1))()
    keep = []
    for i, p in enumerate(packs):
        if p is None or p.nblocks == 0:
            refs[i].pack = None
        else:
            p = p.normalized()
            v = view_of(p)
            keep.append((p, v))
            refs[i].pack = C.pointer(v)
        refs[i].key, refs[i].version = (keys[i] if keys is not None else (0, 0))
    return refs, keep


def _exact(p) -> int:
    from .codec import ExactLen

    return ExactLen(p)


def IntersectWithLinPacked(u, v, blockSize: int = 256):
    """algo.IntersectWithLinPacked(u, v *pb.UidPack) *pb.UidPack (algo/packed.go:35-101): result as a pack."""
    from .codec import _call_pack_out

    lib = _lib.load()
    refs, keep = _refs_of([u, v])
    return _call_pack_out(lambda b, n, d, dl, pnb, pdb, pv: lib.dgx_intersect_packed(
        C.byref(refs[0]), C.byref(refs[1]), blockSize, b, n, d, dl, pnb, pdb, pv), min(_exact(u), _exact(v)), blockSize)


def DifferencePacked(u, v, blockSize: int = 256):
    """algo.DifferencePacked(u, v) (algo/packed.go:141-226): u \\ v as a pack."""
    from .codec import _call_pack_out

    lib = _lib.load()
    refs, keep = _refs_of([u, v])
    return _call_pack_out(lambda b, n, d, dl, pnb, pdb, pv: lib.dgx_difference_packed(
        C.byref(refs[0]), C.byref(refs[1]), blockSize, b, n, d, dl, pnb, pdb, pv), _exact(u), blockSize)


def MergeSortedPacked(lists, blockSize: int = 256):
    """algo.MergeSortedPacked(lists) (algo/packed.go:228-297): sorted de-duplicated union as a pack."""
    from .codec import _call_pack_out

    lib = _lib.load()
    refs, keep = _refs_of(lists)
    return _call_pack_out(lambda b, n, d, dl, pnb, pdb, pv: lib.dgx_merge_sorted_packed(
        refs, len(lists), blockSize, b, n, d, dl, pnb, pdb, pv), sum(_exact(p) for p in lists), blockSize)


def IntersectSortedPacked(lists, blockSize: int = 256):
    """algo.IntersectSortedPacked(lists) (algo/packed.go:103-139) over ALL lists (the reference's loop keeps only
    ls[0] ∩ ls[1]); no lists -> nil pack."""
    from .codec import _call_pack_out

    if len(lists) == 0:
        return None
    lib = _lib.load()
    refs, keep = _refs_of(lists)
    return _call_pack_out(lambda b, n, d, dl, pnb, pdb, pv: lib.dgx_intersect_sorted_packed_out(
        refs, len(lists), blockSize, b, n, d, dl, pnb, pdb, pv), min(_exact(p) for p in lists), blockSize)
