"""The C-ABI library loads on a CPU-only box and exports every symbol include/dgx.h
declares; without a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dgx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dgx_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from dgraph_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"libdgx.so does not export {n}"
        assert n in _lib.SYMBOLS, f"python binding table misses {n}"
    assert sorted(_lib.SYMBOLS) == names


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dgraph_b200 import _lib, algo, pb
    lib = _lib.load()
    assert lib.dgx_init(-1) == -5  # DGX_ERR_NODEV
    assert b"CUDA" in lib.dgx_last_error()
    o = pb.List(None)
    with pytest.raises(_lib.DgxError):
        algo.IntersectWith(pb.List([1, 2, 3]), pb.List([2, 3]), o)
    with pytest.raises(_lib.DgxError):
        algo.MergeSorted([pb.List([1]), pb.List([2])])
    assert lib.dgx_lane_create(0, None) is None


def test_nil_shapes_need_no_gpu():
    # shapes the Go callers rely on are produced by the host mirror itself
    from dgraph_b200 import algo, codec, pb
    assert algo.IntersectSorted([]).Uids is None
    d = algo.Difference(None, pb.List([1]))
    assert d.Uids is not None and d.Uids.size == 0
    assert codec.Decode(None, 0).size == 0
    assert codec.ApproxLen(None) == 0 and codec.ExactLen(None) == 0
    p = pb.UidPack(10, np.array([1, 50], np.uint64), np.array([3, 2], np.uint32),
                   np.array([0, 5, 10], np.uint64), np.zeros(10, np.uint8))
    assert codec.ApproxLen(p) == 20 and codec.ExactLen(p) == 5
