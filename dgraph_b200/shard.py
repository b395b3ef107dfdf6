"""Multi-GPU sharding of batches of independent set operations (pattern 1).

The reference fans a batch of independent posting-list intersections over
goroutines (x.DivideAndRule, worker/task.go:816-987); here the same batch is
partitioned over the GPUs of one box, one process per GPU.  Units (pairs /
queries) are independent, so there is NO data-path collective; the only exchange
is the final result concatenation (an all-gatherv = all_gather of counts, then of
payloads padded to the largest shard), done with torch.distributed (NCCL on GPUs,
gloo in the CPU tests).

This module is host logic only: which units a rank owns and how results are put
back in unit order.  The compute step is passed in by the caller (libdgx on a
GPU box).
"""
from __future__ import annotations

import os
from typing import Callable, List, Sequence, Tuple

import numpy as np


def lpt_partition(costs: Sequence[int], world: int) -> List[np.ndarray]:
    """Longest-processing-time-first greedy partition of units by cost (bytes).

    Returns, per rank, the sorted unit indices it owns.  Deterministic: ties go to
    the lowest rank, units are visited in (cost desc, index asc) order.
    """
    costs = np.asarray(costs, dtype=np.int64)
    order = np.lexsort((np.arange(costs.size), -costs))
    load = np.zeros(world, dtype=np.int64)
    owner = np.empty(costs.size, dtype=np.int64)
    for u in order:
        r = int(np.argmin(load))  # first minimum -> lowest rank on ties
        owner[u] = r
        load[r] += costs[u]
    return [np.nonzero(owner == r)[0] for r in range(world)]


def contiguous_partition(costs: Sequence[int], world: int) -> List[np.ndarray]:
    """Split units 0..n-1 into `world` CONTIGUOUS blocks of roughly equal cost (cut at the cost
    prefix nearest to r/world of the total).  Imbalance is at most one unit's cost; in exchange a
    rank's results are already contiguous in unit order, so the final concatenation needs no
    scatter: rank payloads are simply laid end to end."""
    costs = np.asarray(costs, dtype=np.int64)
    n = costs.size
    pref = np.concatenate([[0], np.cumsum(costs)])
    total = int(pref[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        c = int(np.searchsorted(pref, target, side="left"))
        cuts.append(min(max(c, cuts[-1]), n))
    cuts.append(n)
    return [np.arange(cuts[r], cuts[r + 1]) for r in range(world)]


def gatherv_contiguous(dist, local_out, local_off, device=None):
    """All-gatherv for contiguous shards (contiguous_partition): returns (out, off) in unit order.
    local_off: this rank's CSR offsets (len = units + 1).  Counts travel first, then payloads padded
    to the largest shard; rank payloads are laid end to end."""
    import torch

    world = dist.get_world_size()
    dev = local_out.device if device is None else device
    meta = torch.tensor([local_off.numel() - 1, int(local_off[-1].item())], dtype=torch.int64, device=dev)
    metas = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(metas, meta)
    nun = [int(m[0].item()) for m in metas]
    tot = [int(m[1].item()) for m in metas]
    max_u, max_t = max(max(nun), 1), max(max(tot), 1)
    lens_local = torch.zeros(max_u, dtype=torch.int64, device=dev)
    lens_local[: local_off.numel() - 1] = local_off[1:] - local_off[:-1]
    lens_all = torch.empty(world * max_u, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(lens_all, lens_local)
    pay_local = torch.empty(max_t, dtype=torch.int64, device=dev)
    pay_local[: int(local_off[-1].item())] = local_out[: int(local_off[-1].item())]
    pay_all = torch.empty(world * max_t, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(pay_all, pay_local)
    lens_g = torch.cat([lens_all[r * max_u: r * max_u + nun[r]] for r in range(world)])
    off_g = torch.zeros(lens_g.numel() + 1, dtype=torch.int64, device=dev)
    off_g[1:] = torch.cumsum(lens_g, 0)
    out_g = torch.cat([pay_all[r * max_t: r * max_t + tot[r]] for r in range(world)])
    return out_g, off_g


def gatherv_exact(dist, parts, local_out, local_off, device=None):
    """All-gatherv for CONTIGUOUS shards (contiguous_partition) with exact sizes: returns (out, off) in unit order,
    identical on every rank.

    1. one all_gather of the per-unit result lengths (shards padded to the largest shard's unit count, which is
       static: the partition is computed from the inputs' sizes on every rank);
    2. ONE host read of those lengths (the only synchronisation): offsets and per-rank totals follow on the host;
    3. every rank's payload travels straight to its final position in `out` -- grouped point-to-point sends and
       receives of the exact sizes (ncclGroupStart/End under NCCL), no padding to the largest shard, no scatter.
    """
    import torch

    world = dist.get_world_size()
    rank = dist.get_rank()
    dev = local_out.device if device is None else device
    n_r = [len(p) for p in parts]
    max_u = max(max(n_r), 1)
    lens_local = torch.zeros(max_u, dtype=torch.int64, device=dev)
    if n_r[rank]:
        lens_local[: n_r[rank]] = local_off[1: n_r[rank] + 1] - local_off[: n_r[rank]]
    lens_all = torch.empty(world * max_u, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(lens_all, lens_local)
    lens_host = lens_all.cpu().numpy().reshape(world, max_u)          # the one host synchronisation
    tot = [int(lens_host[r, : n_r[r]].sum()) for r in range(world)]
    lens_g = np.concatenate([lens_host[r, : n_r[r]] for r in range(world)]) if sum(n_r) else np.zeros(0, np.int64)
    off_host = np.zeros(lens_g.size + 1, dtype=np.int64)
    np.cumsum(lens_g, out=off_host[1:])
    base = np.concatenate([[0], np.cumsum(tot)]).astype(np.int64)
    out_g = torch.empty(max(int(base[-1]), 1), dtype=torch.int64, device=dev)
    if dist.get_backend() == "nccl" and world > 2 and 8 * int(base[-1]) >= (64 << 20):
        # Bandwidth regime on many GPUs: NCCL's point-to-point transfers to 7 peers ran at ~240 GB/s per rank on an
        # 8-GPU box, its all_gather collective (ring / NVLS) runs at several times that.  The collective needs equal
        # shard sizes, so shards travel padded to the largest one (contiguous_partition balances by bytes: the padding
        # is a few percent) and are then moved to their final, compact positions at HBM speed.
        max_t = max(tot)
        send = local_out[:max_t] if local_out.numel() >= max_t else torch.cat(
            [local_out[: tot[rank]], torch.empty(max_t - tot[rank], dtype=torch.int64, device=dev)])
        buf = torch.empty(world * max_t, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(buf, send)
        for r in range(world):
            if tot[r]:
                out_g[int(base[r]): int(base[r]) + tot[r]].copy_(buf[r * max_t: r * max_t + tot[r]])
        return out_g[: int(base[-1])], torch.from_numpy(off_host).to(dev)
    # grouped point-to-point transfers of the exact sizes (ncclGroupStart/End under NCCL), one per peer.  Measured on
    # 2 GPUs, 0.96 GB per peer: 3.9 ms this way; cut into 8 or 64 MiB pieces 8.4 ms; NCCL's uneven all_gather (one
    # broadcast per rank) 6.5 ms.
    piece = int(os.environ.get("DGX_GATHER_PIECE", str(1 << 40)))  # values per transfer (default: one per peer)
    ops = []
    for r in range(world):
        if r == rank:
            if tot[r]:
                out_g[int(base[r]): int(base[r]) + tot[r]].copy_(local_out[: tot[r]])
            continue
        for o in range(0, tot[rank], piece):
            ops.append(dist.P2POp(dist.isend, local_out[o: min(o + piece, tot[rank])], r))
        for o in range(0, tot[r], piece):
            ops.append(dist.P2POp(dist.irecv, out_g[int(base[r]) + o: int(base[r]) + min(o + piece, tot[r])], r))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out_g[: int(base[-1])], torch.from_numpy(off_host).to(dev)


def shard_offsets(counts: np.ndarray) -> np.ndarray:
    out = np.zeros(counts.size + 1, dtype=np.int64)
    np.cumsum(counts, out=out[1:])
    return out


def gatherv_results(dist, local_units: np.ndarray, local_out, local_off, n_units: int, device=None):
    """All-gatherv of per-unit results.

    local_units: global indices of the units this rank computed (ascending).
    local_out:   torch int64 tensor, this rank's results concatenated in local unit order.
    local_off:   torch int64 tensor (len(local_units)+1), CSR offsets into local_out.
    Returns (out, off): torch tensors holding every unit's result in GLOBAL unit order
    (identical on all ranks).
    """
    import torch

    world = dist.get_world_size()
    rank = dist.get_rank()
    dev = local_out.device if device is None else device
    # 1) per-unit lengths of every rank, padded to the largest shard (all_gather needs equal shapes)
    n_local = torch.tensor([len(local_units)], dtype=torch.int64, device=dev)
    n_all = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(n_all, n_local)
    n_all = [int(t.item()) for t in n_all]
    max_units = max(max(n_all), 1)
    lens_local = torch.zeros(max_units, dtype=torch.int64, device=dev)
    ids_local = torch.full((max_units,), -1, dtype=torch.int64, device=dev)
    if len(local_units):
        lens_local[: len(local_units)] = local_off[1:] - local_off[:-1]
        ids_local[: len(local_units)] = torch.as_tensor(np.asarray(local_units), dtype=torch.int64, device=dev)
    lens_all = [torch.zeros_like(lens_local) for _ in range(world)]
    ids_all = [torch.zeros_like(ids_local) for _ in range(world)]
    dist.all_gather(lens_all, lens_local)
    dist.all_gather(ids_all, ids_local)
    # 2) payloads, padded to the largest shard
    tot_local = torch.tensor([int(local_off[-1].item()) if local_off.numel() else 0], dtype=torch.int64, device=dev)
    tot_all = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(tot_all, tot_local)
    tot_all = [int(t.item()) for t in tot_all]
    max_tot = max(max(tot_all), 1)
    pay_local = torch.zeros(max_tot, dtype=torch.int64, device=dev)
    pay_local[: tot_all[rank]] = local_out[: tot_all[rank]]
    pay_all = [torch.zeros_like(pay_local) for _ in range(world)]
    dist.all_gather(pay_all, pay_local)
    # 3) scatter back into global unit order
    lens_g = torch.zeros(n_units, dtype=torch.int64, device=dev)
    for r in range(world):
        if n_all[r]:
            lens_g[ids_all[r][: n_all[r]]] = lens_all[r][: n_all[r]]
    off_g = torch.zeros(n_units + 1, dtype=torch.int64, device=dev)
    off_g[1:] = torch.cumsum(lens_g, 0)
    out_g = torch.zeros(max(int(off_g[-1].item()), 1), dtype=torch.int64, device=dev)
    for r in range(world):
        if n_all[r] == 0 or tot_all[r] == 0:
            continue
        ids = ids_all[r][: n_all[r]]
        ln = lens_all[r][: n_all[r]]
        src_off = torch.cumsum(ln, 0) - ln
        dst_off = off_g[ids]
        # element-wise destination index of every payload value of rank r
        rep = torch.repeat_interleave(torch.arange(n_all[r], device=dev), ln)
        within = torch.arange(tot_all[r], device=dev) - src_off[rep]
        out_g[dst_off[rep] + within] = pay_all[r][: tot_all[r]]
    return out_g[: int(off_g[-1].item())], off_g


def run_sharded_pairs(dist, a_lists, b_lists, compute: Callable, device=None, partition: str = "contiguous"):
    """Intersect pair i = (a_lists[i], b_lists[i]) for all i, sharded over the process group.

    compute(units, a_lists, b_lists) -> (out int64 tensor, off int64 tensor) for the units a
    rank owns (on a GPU box: one dgx_dev_filter_batch launch).  Returns results in pair order.
    partition "contiguous" (default): shards are runs of consecutive pairs balanced by bytes, results come back
    through gatherv_exact (one host sync, exact sizes).  "lpt": longest-processing-time partition (tighter balance
    when a few pairs dominate) with the general, index-scattering gatherv_results.
    """
    world = dist.get_world_size()
    rank = dist.get_rank()
    costs = [8 * (len(a) + len(b)) for a, b in zip(a_lists, b_lists)]
    if partition == "lpt":
        parts = lpt_partition(costs, world)
        mine = parts[rank]
        out, off = compute(mine, a_lists, b_lists)
        return gatherv_results(dist, mine, out, off, len(a_lists), device=device)
    parts = contiguous_partition(costs, world)
    mine = parts[rank]
    out, off = compute(mine, a_lists, b_lists)
    return gatherv_exact(dist, parts, out, off, device=device)
