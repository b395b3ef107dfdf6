"""ctypes binding of oracle/liboracle.so (CPU restatement of dgraph's algo/ + codec/).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product (dgraph_b200/)
never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

SEEK_START = 0
SEEK_CURRENT = 1
MAX_U64 = 0xFFFFFFFFFFFFFFFF


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)
    ):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


class _Pack(C.Structure):
    _fields_ = [
        ("block_size", C.c_uint32),
        ("nblocks", C.c_size_t),
        ("base", C.POINTER(C.c_uint64)),
        ("num_uids", C.POINTER(C.c_uint32)),
        ("delta_off", C.POINTER(C.c_uint64)),
        ("deltas", C.POINTER(C.c_uint8)),
    ]


_u64p = C.POINTER(C.c_uint64)
_szp = C.POINTER(C.c_size_t)
_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    sz, u64, vp = C.c_size_t, C.c_uint64, C.c_void_p
    L.orc_intersect_with_lin.argtypes = [vp, sz, vp, sz, vp, _szp, _szp, _szp]
    L.orc_intersect_with_lin.restype = None
    L.orc_intersect_with_jump.argtypes = [vp, sz, vp, sz, vp, _szp, _szp, _szp]
    L.orc_intersect_with_jump.restype = None
    L.orc_intersect_with_bin.argtypes = [vp, sz, vp, sz, vp, _szp]
    L.orc_intersect_with_bin.restype = sz
    L.orc_intersect_with.argtypes = [vp, sz, vp, sz, vp]
    L.orc_intersect_with.restype = sz
    L.orc_intersect_with_branch.argtypes = [sz, sz]
    L.orc_intersect_with_branch.restype = C.c_int
    L.orc_intersect_sorted.argtypes = [vp, vp, sz, vp]
    L.orc_intersect_sorted.restype = sz
    L.orc_difference.argtypes = [vp, sz, vp, sz, vp]
    L.orc_difference.restype = sz
    L.orc_merge_sorted.argtypes = [vp, vp, sz, vp]
    L.orc_merge_sorted.restype = sz
    L.orc_internal_merge_sort.argtypes = [vp, vp, sz, vp]
    L.orc_internal_merge_sort.restype = sz
    L.orc_index_of.argtypes = [vp, sz, u64]
    L.orc_index_of.restype = C.c_longlong
    L.orc_gv_encode4.argtypes = [vp, vp]
    L.orc_gv_encode4.restype = sz
    L.orc_gv_decode4.argtypes = [vp, vp]
    L.orc_gv_decode4.restype = None
    L.orc_gv_bytes_used.argtypes = [C.c_uint8]
    L.orc_gv_bytes_used.restype = sz
    L.orc_encode.argtypes = [vp, sz, C.c_int]
    L.orc_encode.restype = C.POINTER(_Pack)
    L.orc_pack_free.argtypes = [C.POINTER(_Pack)]
    L.orc_pack_free.restype = None
    L.orc_approx_len.argtypes = [C.POINTER(_Pack)]
    L.orc_approx_len.restype = sz
    L.orc_exact_len.argtypes = [C.POINTER(_Pack)]
    L.orc_exact_len.restype = sz
    L.orc_decode.argtypes = [C.POINTER(_Pack), u64, vp]
    L.orc_decode.restype = sz
    L.orc_decoder_new.argtypes = [C.POINTER(_Pack)]
    L.orc_decoder_new.restype = vp
    L.orc_decoder_free.argtypes = [vp]
    L.orc_decoder_free.restype = None
    for name in ("unpack_block", "next", "uids"):
        f = getattr(L, "orc_decoder_" + name)
        f.argtypes = [vp, _szp]
        f.restype = _u64p
    L.orc_decoder_seek.argtypes = [vp, u64, C.c_int, _szp]
    L.orc_decoder_seek.restype = _u64p
    L.orc_decoder_seek_to_block.argtypes = [vp, u64, C.c_int, _szp]
    L.orc_decoder_seek_to_block.restype = _u64p
    L.orc_decoder_linear_seek.argtypes = [vp, u64, _szp]
    L.orc_decoder_linear_seek.restype = _u64p
    L.orc_decoder_peek_next_base.argtypes = [vp]
    L.orc_decoder_peek_next_base.restype = u64
    L.orc_decoder_valid.argtypes = [vp]
    L.orc_decoder_valid.restype = C.c_int
    L.orc_decoder_block_idx.argtypes = [vp]
    L.orc_decoder_block_idx.restype = sz
    L.orc_decoder_set_block_idx.argtypes = [vp, sz]
    L.orc_decoder_set_block_idx.restype = None
    L.orc_decoder_approx_len.argtypes = [vp]
    L.orc_decoder_approx_len.restype = sz
    L.orc_intersect_compressed_with.argtypes = [C.POINTER(_Pack), u64, vp, sz, vp]
    L.orc_intersect_compressed_with.restype = sz
    L.orc_intersect_compressed_with_lin_jump.argtypes = [vp, vp, sz, vp]
    L.orc_intersect_compressed_with_lin_jump.restype = sz
    L.orc_intersect_compressed_with_bin.argtypes = [vp, vp, sz, vp]
    L.orc_intersect_compressed_with_bin.restype = sz
    _lib = L
    return L


def _arr(x) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(x, dtype=np.uint64))
    return a


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _lists_args(lists):
    arrs = [_arr(l) for l in lists]
    k = len(arrs)
    ptrs = (C.c_void_p * max(k, 1))(*[a.ctypes.data for a in arrs])
    lens = (C.c_size_t * max(k, 1))(*[a.size for a in arrs])
    return arrs, ptrs, lens, k


# ---- algo ---------------------------------------------------------------


def intersect_with(u, v) -> np.ndarray:
    """algo.IntersectWith(u, v, o) -> o.Uids (algo/uidlist.go:142)."""
    u, v = _arr(u), _arr(v)
    out = np.empty(max(min(u.size, v.size), 1), dtype=np.uint64)
    n = lib().orc_intersect_with(_ptr(u), u.size, _ptr(v), v.size, _ptr(out))
    return out[:n].copy()


def intersect_with_inplace(u: np.ndarray, v) -> int:
    """IntersectWith(u, v, u): result written over u's storage; returns new length."""
    v = _arr(v)
    assert u.dtype == np.uint64 and u.flags.c_contiguous
    return lib().orc_intersect_with(_ptr(u), u.size, _ptr(v), v.size, _ptr(u))


def _two(fn, u, v, ret_pos=False):
    u, v = _arr(u), _arr(v)
    out = np.empty(max(min(u.size, v.size), 1), dtype=np.uint64)
    olen, ri, rk = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    fn(_ptr(u), u.size, _ptr(v), v.size, _ptr(out), C.byref(olen), C.byref(ri), C.byref(rk))
    res = out[: olen.value].copy()
    return (res, ri.value, rk.value) if ret_pos else res


def intersect_with_lin(u, v, ret_pos=False):
    return _two(lib().orc_intersect_with_lin, u, v, ret_pos)


def intersect_with_jump(u, v, ret_pos=False):
    return _two(lib().orc_intersect_with_jump, u, v, ret_pos)


def intersect_with_bin(d, q) -> np.ndarray:
    d, q = _arr(d), _arr(q)
    out = np.empty(max(min(d.size, q.size), 1), dtype=np.uint64)
    olen = C.c_size_t(0)
    lib().orc_intersect_with_bin(_ptr(d), d.size, _ptr(q), q.size, _ptr(out), C.byref(olen))
    return out[: olen.value].copy()


def intersect_with_branch(n: int, m: int) -> int:
    return lib().orc_intersect_with_branch(n, m)


def intersect_sorted(lists) -> np.ndarray:
    arrs, ptrs, lens, k = _lists_args(lists)
    cap = min([a.size for a in arrs], default=0)
    out = np.empty(max(cap, 1), dtype=np.uint64)
    n = lib().orc_intersect_sorted(ptrs, lens, k, _ptr(out))
    return out[:n].copy()


def difference(u, v) -> np.ndarray:
    u, v = _arr(u), _arr(v)
    out = np.empty(max(u.size, 1), dtype=np.uint64)
    n = lib().orc_difference(_ptr(u), u.size, _ptr(v), v.size, _ptr(out))
    return out[:n].copy()


def merge_sorted(lists, single_heap: bool = False) -> np.ndarray:
    arrs, ptrs, lens, k = _lists_args(lists)
    cap = sum(a.size for a in arrs)
    out = np.empty(max(cap, 1), dtype=np.uint64)
    fn = lib().orc_internal_merge_sort if single_heap else lib().orc_merge_sorted
    n = fn(ptrs, lens, k, _ptr(out))
    return out[:n].copy()


def index_of(u, uid: int) -> int:
    u = _arr(u)
    return lib().orc_index_of(_ptr(u), u.size, uid)


# ---- group varint ---------------------------------------------------------


def gv_encode4(vals) -> bytes:
    src = np.asarray(vals, dtype=np.uint32)
    dst = np.zeros(17, dtype=np.uint8)
    n = lib().orc_gv_encode4(dst.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p))
    return bytes(dst[:n])


def gv_decode4(buf: bytes):
    src = np.zeros(max(len(buf), 17) + 16, dtype=np.uint8)
    src[: len(buf)] = np.frombuffer(buf, dtype=np.uint8)
    dst = np.zeros(4, dtype=np.uint32)
    lib().orc_gv_decode4(dst.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p))
    return [int(x) for x in dst]


def gv_bytes_used(tag: int) -> int:
    return lib().orc_gv_bytes_used(tag)


# ---- codec ----------------------------------------------------------------


class Pack:
    """pb.UidPack as produced by codec.Encode (SoA view; None-pack == nil)."""

    def __init__(self, handle):
        self._h = handle  # POINTER(_Pack) or None

    @property
    def is_nil(self) -> bool:
        return not bool(self._h)

    @property
    def block_size(self) -> int:
        return 0 if self.is_nil else self._h.contents.block_size

    @property
    def nblocks(self) -> int:
        return 0 if self.is_nil else self._h.contents.nblocks

    def _np(self, ptr, n, dtype):
        if n == 0:
            return np.zeros(0, dtype=dtype)
        return np.ctypeslib.as_array(ptr, shape=(n,)).copy()

    @property
    def base(self) -> np.ndarray:
        return np.zeros(0, np.uint64) if self.is_nil else self._np(self._h.contents.base, self.nblocks, np.uint64)

    @property
    def num_uids(self) -> np.ndarray:
        return np.zeros(0, np.uint32) if self.is_nil else self._np(self._h.contents.num_uids, self.nblocks, np.uint32)

    @property
    def delta_off(self) -> np.ndarray:
        if self.is_nil:
            return np.zeros(1, np.uint64)
        return self._np(self._h.contents.delta_off, self.nblocks + 1, np.uint64)

    @property
    def deltas(self) -> np.ndarray:
        """All delta bytes concatenated (block b = deltas[delta_off[b]:delta_off[b+1]])."""
        if self.is_nil:
            return np.zeros(0, np.uint8)
        n = int(self.delta_off[-1])
        return self._np(self._h.contents.deltas, n, np.uint8)

    def block_deltas(self, b: int) -> bytes:
        off = self.delta_off
        return bytes(self.deltas[int(off[b]): int(off[b + 1])])

    def handle(self):
        return self._h if self._h else None

    def __del__(self):
        try:
            if self._h:
                lib().orc_pack_free(self._h)
                self._h = None
        except Exception:
            pass


def encode(uids, block_size: int) -> Pack:
    """codec.Encode (codec/codec.go:393)."""
    a = _arr(uids)
    h = lib().orc_encode(_ptr(a), a.size, block_size)
    return Pack(h if h else None)


def approx_len(p: Pack) -> int:
    return lib().orc_approx_len(p.handle())


def exact_len(p: Pack) -> int:
    return lib().orc_exact_len(p.handle())


def decode(p: Pack, seek: int = 0) -> np.ndarray:
    """codec.Decode (codec/codec.go:444)."""
    out = np.empty(max(exact_len(p), 1), dtype=np.uint64)
    n = lib().orc_decode(p.handle(), seek, _ptr(out))
    return out[:n].copy()


class Decoder:
    """codec.Decoder (codec/codec.go:139-384)."""

    def __init__(self, pack: Pack):
        self.pack = pack
        self._d = lib().orc_decoder_new(pack.handle())

    def _ret(self, ptr, ln):
        n = ln.value
        if n == 0:
            return np.zeros(0, dtype=np.uint64)
        return np.ctypeslib.as_array(ptr, shape=(n,)).copy()

    def unpack_block(self):
        ln = C.c_size_t(0)
        return self._ret(lib().orc_decoder_unpack_block(self._d, C.byref(ln)), ln)

    def seek(self, uid, whence):
        ln = C.c_size_t(0)
        return self._ret(lib().orc_decoder_seek(self._d, uid, whence, C.byref(ln)), ln)

    def seek_to_block(self, uid, whence):
        ln = C.c_size_t(0)
        return self._ret(lib().orc_decoder_seek_to_block(self._d, uid, whence, C.byref(ln)), ln)

    def linear_seek(self, uid):
        ln = C.c_size_t(0)
        return self._ret(lib().orc_decoder_linear_seek(self._d, uid, C.byref(ln)), ln)

    def next(self):
        ln = C.c_size_t(0)
        return self._ret(lib().orc_decoder_next(self._d, C.byref(ln)), ln)

    def uids(self):
        ln = C.c_size_t(0)
        return self._ret(lib().orc_decoder_uids(self._d, C.byref(ln)), ln)

    def peek_next_base(self) -> int:
        return lib().orc_decoder_peek_next_base(self._d)

    def valid(self) -> bool:
        return bool(lib().orc_decoder_valid(self._d))

    @property
    def block_idx(self) -> int:
        return lib().orc_decoder_block_idx(self._d)

    @block_idx.setter
    def block_idx(self, v: int):
        lib().orc_decoder_set_block_idx(self._d, v)

    def approx_len(self) -> int:
        return lib().orc_decoder_approx_len(self._d)

    def __del__(self):
        try:
            if self._d:
                lib().orc_decoder_free(self._d)
                self._d = None
        except Exception:
            pass


def intersect_compressed_with(p: Pack, after_uid: int, v) -> np.ndarray:
    v = _arr(v)
    out = np.empty(max(v.size, 1), dtype=np.uint64)
    n = lib().orc_intersect_compressed_with(p.handle(), after_uid, _ptr(v), v.size, _ptr(out))
    return out[:n].copy()


def intersect_compressed_with_lin_jump(dec: Decoder, v) -> np.ndarray:
    v = _arr(v)
    out = np.empty(max(v.size, 1), dtype=np.uint64)
    n = lib().orc_intersect_compressed_with_lin_jump(dec._d, _ptr(v), v.size, _ptr(out))
    return out[:n].copy()


def intersect_compressed_with_bin(dec: Decoder, q) -> np.ndarray:
    q = _arr(q)
    out = np.empty(max(q.size, 1), dtype=np.uint64)
    n = lib().orc_intersect_compressed_with_bin(dec._d, _ptr(q), q.size, _ptr(out))
    return out[:n].copy()
