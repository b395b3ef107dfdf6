"""The two protobuf messages on the hot path, as plain Python containers.

pb.List (protos/pb.proto:22-24, pb.pb.go:541-547): Uids []uint64 (None == nil).
pb.UidPack / pb.UidBlock (protos/pb.proto:379-400) flattened to the struct of
arrays libdgx consumes (dgx_pack_view): block b has Base=base[b],
NumUids=num_uids[b], Deltas=deltas[delta_off[b]:delta_off[b+1]].
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np


def _u64(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64))


class List:
    """pb.List.  `Uids is None` models Go's nil slice."""

    __slots__ = ("Uids",)

    def __init__(self, uids=None):
        self.Uids: Optional[np.ndarray] = None if uids is None else _u64(uids)

    def __len__(self) -> int:
        return 0 if self.Uids is None else int(self.Uids.size)

    def tolist(self):
        return [] if self.Uids is None else self.Uids.tolist()

    def __repr__(self):
        return f"List({self.tolist()!r})"


@dataclass
class UidPack:
    """pb.UidPack as a struct of arrays (see include/dgx.h dgx_pack_view)."""

    block_size: int = 0
    base: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint64))
    num_uids: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    delta_off: np.ndarray = field(default_factory=lambda: np.zeros(1, np.uint64))
    deltas: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint8))

    @property
    def nblocks(self) -> int:
        return int(self.base.size)

    def normalized(self) -> "UidPack":
        return UidPack(
            int(self.block_size),
            np.ascontiguousarray(self.base, dtype=np.uint64),
            np.ascontiguousarray(self.num_uids, dtype=np.uint32),
            np.ascontiguousarray(self.delta_off, dtype=np.uint64),
            np.ascontiguousarray(self.deltas, dtype=np.uint8),
        )
