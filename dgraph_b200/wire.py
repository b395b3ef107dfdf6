"""Protobuf wire-format adjacency of the hot path, through libdgx's dgx_wire_* entry points (host only).

Stored posting lists are serialized pb.PostingList values (proto.Unmarshal at posting/list.go:1045,
posting/mvcc.go:634); results travel as pb.List inside pb.Result.uid_matrix (protos/pb.proto:22-24, 76-78).
These helpers go from those bytes to the containers of dgraph_b200.pb and back without a protobuf runtime:
SURVEY.md 8(f) row 3.  No GPU is needed for any of them.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib, pb


def _buf(data) -> tuple:
    b = bytes(data)
    return b, C.c_char_p(b) if b else None, len(b)


def PostingListPack(data) -> Optional[bytes]:
    """The serialized pb.UidPack inside a serialized pb.PostingList (field 1); None when absent (nil pack)."""
    lib = _lib.load()
    b, p, n = _buf(data)
    out, out_len = C.c_void_p(), C.c_size_t(0)
    _lib.check(lib.dgx_wire_posting_list_pack(C.cast(p, C.c_void_p), n, C.byref(out), C.byref(out_len)))
    if not out.value:
        return None
    off = out.value - C.cast(p, C.c_void_p).value
    return b[off:off + out_len.value]


def ParseUidPack(data) -> pb.UidPack:
    """Serialized pb.UidPack -> the struct-of-arrays pb.UidPack that codec.Decode / libdgx take."""
    lib = _lib.load()
    b, p, n = _buf(data)
    vp = C.cast(p, C.c_void_p)
    nb, db = C.c_size_t(0), C.c_size_t(0)
    _lib.check(lib.dgx_wire_pack_measure(vp, n, C.byref(nb), C.byref(db)))
    base = np.zeros(nb.value, np.uint64)
    num = np.zeros(nb.value, np.uint32)
    off = np.zeros(nb.value + 1, np.uint64)
    deltas = np.zeros(db.value, np.uint8)
    view = _lib.PackView()
    _lib.check(lib.dgx_wire_pack_parse(vp, n, base.ctypes.data, num.ctypes.data, off.ctypes.data,
                                       deltas.ctypes.data if db.value else None, nb.value, db.value, C.byref(view)))
    return pb.UidPack(int(view.block_size), base, num, off, deltas)


def ListToWire(l: Optional[pb.List]) -> bytes:
    """pb.List -> its serialized message: a short header in front of the little-endian uids themselves."""
    lib = _lib.load()
    n = 0 if l is None else len(l)
    hdr = (C.c_uint8 * 16)()
    h = lib.dgx_wire_list_header(n, C.cast(hdr, C.c_void_p))
    if n == 0:
        return b""
    return bytes(hdr[:h]) + np.ascontiguousarray(l.Uids, dtype="<u8").tobytes()


def ListFromWire(data) -> pb.List:
    """Serialized pb.List -> pb.List (packed and unpacked encodings)."""
    lib = _lib.load()
    b, p, n = _buf(data)
    vp = C.cast(p, C.c_void_p)
    cnt = C.c_size_t(0)
    _lib.check(lib.dgx_wire_list_decode(vp, n, None, 0, C.byref(cnt)))
    out = np.zeros(cnt.value, np.uint64)
    if cnt.value:
        _lib.check(lib.dgx_wire_list_decode(vp, n, out.ctypes.data, cnt.value, C.byref(cnt)))
    return pb.List(out)


def UidMatrixToWire(out, out_off) -> bytes:
    """CSR result (values, nrows + 1 offsets) -> the uid_matrix rows of a serialized pb.Result."""
    lib = _lib.load()
    vals = np.ascontiguousarray(out, dtype=np.uint64)
    offs = np.ascontiguousarray(out_off, dtype=np.uint64)
    nrows = max(int(offs.size) - 1, 0)
    n = C.c_size_t(0)
    vp = vals.ctypes.data if vals.size else None
    _lib.check(lib.dgx_wire_uid_matrix(vp, offs.ctypes.data, nrows, None, 0, C.byref(n)))
    buf = np.zeros(max(n.value, 1), np.uint8)
    _lib.check(lib.dgx_wire_uid_matrix(vp, offs.ctypes.data, nrows, buf.ctypes.data, n.value, C.byref(n)))
    return buf[:n.value].tobytes()
