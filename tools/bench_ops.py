#!/usr/bin/env python
"""Per-operation device-resident benchmarks of libdgx on one B200 (BASELINE configs C1-C5).

Complements bench.py (which times the headline config C2): every row reports CUDA-event
time, input UIDs/s, algorithmic GB/s (SURVEY.md 8d formulas) and the fraction of the
measured HBM peak, plus a correctness check computed independently on the GPU with torch
(sort/unique/isin) or against the original list.  Output: one JSON document on stdout /
--out.

    python tools/bench_ops.py [--out gpurun_out/ops.json] [--small]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dgraph_b200 import _lib  # noqa: E402

DEV = torch.device("cuda", 0)


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def zipf_gaps_gpu(n, gen, a=1.5, clip=1 << 20):
    """Zipf-like gap list on the GPU: gap = min(floor(u^(-1/(a-1))), clip), uid = cumsum(gap)."""
    u = torch.rand(n, device=DEV, generator=gen, dtype=torch.float64).clamp_(min=1e-12)
    g = torch.clamp(torch.floor(u.pow(-1.0 / (a - 1.0))), max=float(clip)).to(torch.int64)
    return torch.cumsum(g, 0)


def thin_gpu(master, p, gen):
    return master[torch.rand(master.numel(), device=DEV, generator=gen) < p].contiguous()


class Lane:
    def __init__(self, lib=None, handle=None):
        self.lib = lib or _lib.load()
        if handle is not None:   # bench.py's lane (already bound to torch's current stream)
            self.h = handle
            return
        _lib.check(self.lib.dgx_init(0))
        self.stream = torch.cuda.Stream(device=DEV)
        torch.cuda.set_stream(self.stream)
        self.h = self.lib.dgx_lane_create(0, C.c_void_p(self.stream.cuda_stream))
        assert self.h

    def sync(self):
        _lib.check(self.lib.dgx_lane_sync(self.h))

    def filter_batch(self, op, lists, k_off, out, out_off):
        self.prepared(op, lists, k_off, out, out_off)()

    def prepared(self, op, lists, k_off, out, out_off):
        """The descriptor tables are marshalled ONCE (20 000 data_ptr() calls take milliseconds of Python): the
        returned closure is what gets timed -- one C-ABI call."""
        n = len(lists)
        ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in lists])
        lens = (C.c_size_t * n)(*[t.numel() for t in lists])
        koff = (C.c_size_t * len(k_off))(*k_off)
        nq, po, cap, poff = len(k_off) - 1, C.c_void_p(out.data_ptr()), out.numel(), C.c_void_p(out_off.data_ptr())

        def call():
            _lib.check(self.lib.dgx_dev_filter_batch(self.h, op, ptrs, lens, koff, nq, po, cap, poff))
        return call

    def merge(self, lists, out, out_len):
        n = len(lists)
        ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in lists])
        lens = (C.c_size_t * n)(*[t.numel() for t in lists])
        _lib.check(self.lib.dgx_dev_merge_sorted(self.h, ptrs, lens, n, C.c_void_p(out.data_ptr()), out.numel(),
                                                 C.c_void_p(out_len.data_ptr())))


def timeit(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    evs[0].record()
    for i in range(reps):
        fn()
        evs[i + 1].record()
    torch.cuda.synchronize()
    ts = [evs[i].elapsed_time(evs[i + 1]) for i in range(reps)]
    return float(np.median(ts)), ts


QUIET = False


def row(name, cfg, ms, uids_in, algo_bytes, ok, extra=None):
    peak, src = peak_gbs()
    gbs = algo_bytes / (ms * 1e-3) / 1e9
    r = {"op": name, "config": cfg, "ms": ms, "uids_in": int(uids_in), "uids_per_s": uids_in / (ms * 1e-3),
         "algorithmic_bytes": int(algo_bytes), "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak,
         "peak_GBps": peak, "peak_source": src, "check": ok}
    if extra:
        r.update(extra)
    if not QUIET:
        print(json.dumps(r), flush=True)
    return r


def quick_rows(lib, lane, dev, peak):
    """bench.py's `ops` key: one compact line per BASELINE config other than the headline, full sizes,
    device-resident, CUDA events, independent torch check; no CPU legs (bench.py has its own)."""
    global DEV, QUIET
    DEV, QUIET = dev, True
    rows = run(Lane(lib, lane), 1, quick=True)
    return [{"op": r["op"], "config": r["config"], "ms": round(r["ms"], 4), "uids_per_s": r["uids_per_s"],
             "algorithmic_GBps": round(r["algorithmic_GBps"], 1), "frac_of_hbm_peak": round(r["algorithmic_GBps"] / peak, 4),
             "check": r["check"]} for r in rows]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--small", action="store_true", help="1/10 sizes (quick functional pass)")
    args = ap.parse_args()
    rows = run(Lane(), 10 if args.small else 1, quick=False)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


def run(L, S, quick):
    import gen as hostgen

    small = S > 1
    gen = torch.Generator(device=DEV)
    gen.manual_seed(1234)
    rows = []

    # ---- C1: 2-way IntersectWith, 2 x 1e5 uniform in [0, 1e7): single call latency + batched x2048
    n1 = 100_000
    pairs = []
    for i in range((512 if quick else 2048) // S):
        a = torch.unique(torch.randint(0, 10_000_000, (int(n1 * 1.02),), device=DEV, generator=gen))[:n1].contiguous()
        b = torch.unique(torch.randint(0, 10_000_000, (int(n1 * 1.02),), device=DEV, generator=gen))[:n1].contiguous()
        pairs.append((a, b))
    out = torch.empty(n1 * len(pairs) + 8, dtype=torch.int64, device=DEV)
    off = torch.zeros(len(pairs) + 1, dtype=torch.int64, device=DEV)
    a, b = pairs[0]
    ms, _ = timeit(L.prepared(0, [a, b], [0, 2], out, off), reps=50)
    L.sync()
    nout = int(off[1].item())
    want = a[torch.isin(a, b, assume_unique=True)]
    ok = bool(torch.equal(out[:nout], want))
    rows.append(row("IntersectWith", "C1: 2 x 1e5 uniform unique in [0,1e7), one call (launch-latency bound, L2 resident)",
                    ms, 2 * n1, 8 * (2 * n1 + nout), ok, {"out": nout}))
    lists, koff = [], [0]
    for a, b in pairs:
        lists += [a, b]
        koff.append(len(lists))
    ms, _ = timeit(L.prepared(0, lists, koff, out, off))
    L.sync()
    offc = off.cpu().numpy()
    ok = True
    for i in (0, len(pairs) // 2, len(pairs) - 1):
        a, b = pairs[i]
        ok = ok and bool(torch.equal(out[offc[i]:offc[i + 1]], a[torch.isin(a, b, assume_unique=True)]))
    tot = sum(a.numel() + b.numel() for a, b in pairs)
    rows.append(row("IntersectWith (batch)", f"C1 x {len(pairs)} pairs in one launch", ms, tot, 8 * (tot + int(offc[-1])), ok))
    del pairs, lists

    # ---- C2 dense variant: k=8, thinnings p=0.9 of a 1.11e6 master, 16 queries
    Q = 16 // (4 if small else 1)
    lists, koff, wants = [], [0], []
    for q in range(Q):
        master = zipf_gaps_gpu(1_110_000, gen)
        ql = [thin_gpu(master, 0.9, gen) for _ in range(8)]
        lists += ql
        koff.append(len(lists))
    cap = sum(min(t.numel() for t in lists[koff[q]:koff[q + 1]]) for q in range(Q))
    out = torch.empty(cap + 8, dtype=torch.int64, device=DEV)
    off = torch.zeros(Q + 1, dtype=torch.int64, device=DEV)
    ms, _ = timeit(L.prepared(0, lists, koff, out, off))
    L.sync()
    offc = off.cpu().numpy()
    w = lists[0]
    for t in lists[1:8]:
        w = w[torch.isin(w, t, assume_unique=True)]
    ok = bool(torch.equal(out[offc[0]:offc[1]], w))
    tot = sum(t.numel() for t in lists)
    rows.append(row("IntersectSorted (dense)", f"C2 variant: k=8 thinnings p=0.9 of a 1.11e6 Zipf master, {Q} queries per launch",
                    ms, tot, 8 * (tot + int(offc[-1])), ok, {"out": int(offc[-1])}))
    del lists

    # ---- C4 (single GPU share): batched 2-way, sizes ~ power law alpha=2 on [1e4, 1e6]
    npairs = 10_000 // S  # the whole 10k-pair batch on one GPU (~7.7 GB of lists)
    rng = np.random.default_rng(401)
    u = rng.random(npairs)
    sizes = np.minimum((1e4 / (1 - u * (1 - 1e4 / 1e6))).astype(np.int64), 1_000_000)
    lists, koff = [], [0]
    for sz in sizes:
        master = zipf_gaps_gpu(int(sz) * 2, gen)
        lists += [thin_gpu(master, 0.5, gen), thin_gpu(master, 0.5, gen)]
        koff.append(len(lists))
    cap = sum(min(lists[2 * i].numel(), lists[2 * i + 1].numel()) for i in range(npairs))
    out = torch.empty(cap + 8, dtype=torch.int64, device=DEV)
    off = torch.zeros(npairs + 1, dtype=torch.int64, device=DEV)
    ms, _ = timeit(L.prepared(0, lists, koff, out, off), reps=5)
    L.sync()
    offc = off.cpu().numpy()
    ok = True
    for i in (0, npairs // 3, npairs - 1):
        a, b = lists[2 * i], lists[2 * i + 1]
        ok = ok and bool(torch.equal(out[offc[i]:offc[i + 1]], a[torch.isin(a, b, assume_unique=True)]))
    tot = sum(t.numel() for t in lists)
    rows.append(row("IntersectWith (batch, skewed sizes)", f"C4: {npairs} pairs in one launch, sizes ~ power law alpha=2 on [1e4,1e6], thinnings p=0.5",
                    ms, tot, 8 * (tot + int(offc[-1])), ok, {"total_uids": tot}))
    del lists, out

    # ---- C5: MergeSorted k=64, total 1e8 (lengths ~ 1/rank), then Difference with a 1e7 list
    total = 100_000_000 // S
    master = zipf_gaps_gpu(2 * total, gen)
    w = 1.0 / np.arange(1, 65)
    lens = (w / w.sum() * total).astype(np.int64)
    lists = [thin_gpu(master, float(l) / master.numel(), gen) for l in lens]
    tot = sum(t.numel() for t in lists)
    out = torch.empty(tot + 8, dtype=torch.int64, device=DEV)
    out_len = torch.zeros(1, dtype=torch.int64, device=DEV)
    ms, _ = timeit(lambda: L.merge(lists, out, out_len), warm=2, reps=5)
    L.sync()
    nm = int(out_len.item())
    want = torch.unique(torch.cat(lists))
    ok = bool(nm == want.numel() and torch.equal(out[:nm], want))
    del want
    rows.append(row("MergeSorted", f"C5: k=64 lists, lengths ~ 1/rank, total {tot} UIDs (thinnings of a {master.numel()} master); single-pass multiway merge (merge_multi.cuh + merge_tile32.cuh)",
                    ms, tot, 8 * (tot + nm), ok, {"out": nm}))
    dlist = thin_gpu(master, 10_000_000 / S / master.numel(), gen)
    merged = out[:nm].clone()
    dout = torch.empty(nm + 8, dtype=torch.int64, device=DEV)
    doff = torch.zeros(2, dtype=torch.int64, device=DEV)
    ms, _ = timeit(L.prepared(1, [merged, dlist], [0, 2], dout, doff), warm=2, reps=5)
    L.sync()
    nd = int(doff[1].item())
    want = merged[~torch.isin(merged, dlist, assume_unique=True)]
    ok = bool(nd == want.numel() and torch.equal(dout[:nd], want))
    rows.append(row("Difference", f"C5: merged ({nm}) minus a {dlist.numel()}-UID list", ms, nm + dlist.numel(),
                    8 * (nm + dlist.numel() + nd), ok, {"out": nd}))
    del lists, out, merged, dout, want, master

    # ---- C3: codec.Decode of a 1e8-UID pack (BlockSize 256), then IntersectSorted with 1e7 and 1e6 lists
    n3 = 100_000_000 // S
    master = zipf_gaps_gpu(n3, gen)
    host = master.cpu().numpy().view(np.uint64)

    class _P:  # the pack's arrays (numpy generator, tests/gen.py; byte-identical to the oracle's Encode)
        pass
    pack = _P()
    _, pack.base, pack.num_uids, pack.delta_off, pack.deltas = hostgen.encode_pack_np_parallel(host, 256, threads=min(32, os.cpu_count() or 1))
    pack.nblocks = pack.base.size
    view = _lib.PackView()
    base, num, doff_, deltas = pack.base, pack.num_uids, pack.delta_off, pack.deltas
    view.block_size, view.nblocks = 256, pack.nblocks
    view.base, view.num_uids, view.delta_off, view.deltas = base.ctypes.data, num.ctypes.data, doff_.ctypes.data, deltas.ctypes.data
    pk = C.c_void_p()
    _lib.check(L.lib.dgx_dev_pack_upload(L.h, C.byref(view), C.byref(pk)))
    pack_bytes = int(deltas.size + 12 * pack.nblocks)
    out = torch.empty(n3 + 8, dtype=torch.int64, device=DEV)
    out_len = torch.zeros(1, dtype=torch.int64, device=DEV)

    def dec():
        _lib.check(L.lib.dgx_dev_decode(L.h, pk, 0, C.c_void_p(out.data_ptr()), n3, C.c_void_p(out_len.data_ptr())))

    ms, _ = timeit(dec, warm=2, reps=5)
    L.sync()
    ok = bool(int(out_len.item()) == n3 and torch.equal(out[:n3], master))
    extra = {"pack_bytes": pack_bytes}
    if not quick:
        from oracle import pyoracle as orc   # CPU comparison leg only

        sample = host[: min(n3, 20_000_000)]
        t0 = time.perf_counter()
        spack = orc.encode(sample, 256)
        t_enc = time.perf_counter() - t0
        t1 = time.perf_counter()
        orc.decode(spack, 0)
        t_dec_cpu = time.perf_counter() - t1
        extra.update({"cpu_oracle_decode_uids_per_s": sample.size / t_dec_cpu, "cpu_oracle_encode_uids_per_s": sample.size / t_enc})
    rows.append(row("codec.Decode", f"C3: {n3}-UID pack, BlockSize 256, {pack.nblocks} blocks, {deltas.size / n3:.3f} delta bytes/UID",
                    ms, n3, pack_bytes + 8 * n3, ok, extra))
    l1 = thin_gpu(master, 0.1, gen)
    l2 = thin_gpu(master, 0.01, gen)
    decoded = out[:n3]
    iout = torch.empty(l2.numel() + 8, dtype=torch.int64, device=DEV)
    ioff = torch.zeros(2, dtype=torch.int64, device=DEV)
    ms_i, _ = timeit(L.prepared(0, [decoded, l1, l2], [0, 3], iout, ioff), warm=2, reps=5)
    L.sync()
    ni = int(ioff[1].item())
    want = l2[torch.isin(l2, l1, assume_unique=True)]
    ok = bool(ni == want.numel() and torch.equal(iout[:ni], want))
    tot = n3 + l1.numel() + l2.numel()
    rows.append(row("IntersectSorted after Decode", f"C3: [decoded {n3}, {l1.numel()}, {l2.numel()}] (driven from the shortest list; long lists are probed, not streamed)",
                    ms_i, tot, 8 * (tot + ni), ok, {"out": ni, "note": "algorithmic bytes count every list once; the kernel reads far fewer"}))
    rows.append(row("Decode + IntersectSorted (pipeline)", "C3 end to end on device: decode to HBM, then intersect",
                    ms + ms_i, tot, pack_bytes + 8 * n3 + 8 * (tot + ni), ok))
    L.lib.dgx_dev_pack_free(pk)
    if quick:
        return rows

    # ---- C3 end to end through the host-pointer C ABI: the pack crosses PCIe COMPRESSED ---------
    # (dgx_decode_intersect_sorted: H2D of the pack + the two lists, decode and intersect on the
    # device, D2H of the result) against the CPU oracle doing codec.Decode + algo.IntersectSorted.
    h1 = l1.cpu().numpy().view(np.uint64)
    h2 = l2.cpu().numpy().view(np.uint64)
    hp = (C.c_void_p * 2)(h1.ctypes.data, h2.ctypes.data)
    hl = (C.c_size_t * 2)(h1.size, h2.size)
    hout = np.empty(h2.size, dtype=np.uint64)
    hn = C.c_size_t(0)

    def e2e():
        _lib.check(L.lib.dgx_decode_intersect_sorted(C.byref(view), 0, hp, hl, 2, hout.ctypes.data_as(C.c_void_p),
                                                      hout.size, C.byref(hn)))

    e2e()
    t0 = time.perf_counter()
    for _ in range(3):
        e2e()
    t_e2e = (time.perf_counter() - t0) / 3
    ok = bool(hn.value == ni and np.array_equal(hout[: hn.value], iout[:ni].cpu().numpy().view(np.uint64)))
    t1 = time.perf_counter()
    dec_cpu = orc.decode(spack, 0)          # the 2e7-UID sample pack, scaled below
    t_dec = (time.perf_counter() - t1) * (n3 / sample.size)
    t1 = time.perf_counter()
    orc.intersect_sorted([host, h1, h2])
    t_int = time.perf_counter() - t1
    r = row("Decode + IntersectSorted (host pointers, e2e)",
            "C3 through dgx_decode_intersect_sorted: pack (compressed) and lists cross PCIe, result comes back",
            t_e2e * 1e3, tot, pack_bytes + 8 * (l1.numel() + l2.numel()) + 8 * ni, ok,
            {"h2d_bytes": pack_bytes + 8 * (h1.size + h2.size), "cpu_oracle_ms": (t_dec + t_int) * 1e3,
             "cpu_oracle_decode_ms": t_dec * 1e3, "cpu_oracle_intersect_ms": t_int * 1e3,
             "speedup_vs_cpu_1thread": (t_dec + t_int) / t_e2e,
             "note": "algorithmic bytes here = bytes that cross PCIe; the CPU figure is the single-threaded oracle (decode scaled from a 2e7-UID sample)"})
    rows.append(r)
    # same call with every input in pinned host memory (dgx_host_alloc): PCIe at DMA speed
    def pin(a):
        ptr = L.lib.dgx_host_alloc(max(a.nbytes, 16))
        C.memmove(ptr, a.ctypes.data, a.nbytes)
        return ptr
    pview = _lib.PackView()
    pview.block_size, pview.nblocks = 256, pack.nblocks
    pins = [pin(base), pin(num), pin(doff_), pin(deltas), pin(h1), pin(h2)]
    pview.base, pview.num_uids, pview.delta_off, pview.deltas = pins[0], pins[1], pins[2], pins[3]
    php = (C.c_void_p * 2)(pins[4], pins[5])

    def e2e_pinned():
        _lib.check(L.lib.dgx_decode_intersect_sorted(C.byref(pview), 0, php, hl, 2, hout.ctypes.data_as(C.c_void_p),
                                                      hout.size, C.byref(hn)))

    e2e_pinned()
    t0 = time.perf_counter()
    for _ in range(3):
        e2e_pinned()
    t_pin = (time.perf_counter() - t0) / 3
    ok = bool(hn.value == ni and np.array_equal(hout[: hn.value], iout[:ni].cpu().numpy().view(np.uint64)))
    rows.append(row("Decode + IntersectSorted (pinned host pointers, e2e)", "as above, inputs in dgx_host_alloc memory",
                    t_pin * 1e3, tot, pack_bytes + 8 * (l1.numel() + l2.numel()) + 8 * ni, ok,
                    {"speedup_vs_cpu_1thread": (t_dec + t_int) / t_pin, "cpu_oracle_ms": (t_dec + t_int) * 1e3}))
    for ptr in pins:
        L.lib.dgx_host_free(ptr)
    return rows


if __name__ == "__main__":
    main()
