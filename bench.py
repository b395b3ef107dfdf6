#!/usr/bin/env python
"""bench.py -- headline benchmark of the posting-list set-op hot path on B200.

Workload (BASELINE.json configs[1]): k-way algo.IntersectSorted, k = 8 Zipf-gap
lists of ~1e6 UIDs each (independent thinnings, p = 0.25, of a shared 4e6 master),
as a batch of Q independent queries per step so that one step's inputs (Q x 64 MB)
exceed the 126 MB L2.  Metric: input UIDs intersected per second, bit-exact vs the
CPU oracle.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU).  Queries are independent, so they
shard across ranks with no data-path collective (weak scaling: Q per rank); the only
exchange is the NCCL all-gatherv that concatenates the per-rank results, and it is
inside the timed region.

`--impl reference` times the reference's CPU algorithm (the C oracle restating
algo/uidlist.go; the Go toolchain is absent) on all host threads, same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

K_LISTS = 8
LIST_LEN = 1_000_000
KEEP_P = 0.25
MASTER = int(LIST_LEN / KEEP_P)
METRIC = "UIDs/sec intersected (bit-exact)"
UNIT = "UIDs/s"


def make_query(seed: int):
    """One config[1] query: 8 thinnings (p=0.25) of a 4e6 Zipf-gap master (a=1.5, clip 2^20)."""
    rng = np.random.default_rng(seed)
    gaps = np.minimum(rng.zipf(1.5, MASTER), 1 << 20).astype(np.uint64)
    master = np.cumsum(gaps, dtype=np.uint64)
    return [master[rng.random(MASTER) < KEEP_P] for _ in range(K_LISTS)]


def make_queries(q: int, rank: int):
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        return list(ex.map(make_query, [201 + 1000 * rank + i for i in range(q)]))


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML during the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.active = False   # only samples taken while the timed region runs are kept
        self.ok = False

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
            }
            self.ok = True
            while not self._halt.is_set():
                if self.active:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                time.sleep(0.002)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_intersect_batch(orc, queries, threads: int) -> float:
    """Seconds for the oracle (algo.IntersectSorted restatement) over all queries on `threads` threads."""
    t0 = time.perf_counter()
    if threads <= 1:
        for q in queries:
            orc.intersect_sorted(q)
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:  # ctypes releases the GIL
            list(ex.map(orc.intersect_sorted, queries))
    return time.perf_counter() - t0


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference CPU algorithm on all host threads (rank 0 only)."""
    if rank != 0:
        return
    from oracle import pyoracle as orc

    orc.lib()
    q = args.queries
    queries = make_queries(q, 0)
    uids = sum(l.size for qq in queries for l in qq)
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_intersect_batch(orc, queries, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_intersect_batch(orc, queries, threads)
    dt = time.perf_counter() - t0
    val = uids * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(q, 1),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{q} queries x {K_LISTS} lists x ~{LIST_LEN} UIDs per step, C restatement of algo.IntersectSorted, queries spread over {threads} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(q: int, world: int):
    return {
        "workload": "configs[1]: k-way IntersectSorted, k=8 Zipf-gap lists of ~1e6 UIDs (thinnings p=0.25 of a 4e6 master), "
                    f"batch of {q} independent queries per GPU per step",
        "k": K_LISTS, "list_len": LIST_LEN, "queries_per_gpu": q, "parallelism": f"queries sharded over {world} GPU(s)",
        "l2": f"inputs {q * K_LISTS * LIST_LEN * 8 / 1e9:.2f} GB per GPU per step > 126 MB L2 (no flush needed)",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--queries", type=int, default=16, help="independent queries per GPU per step")
    ap.add_argument("--e2e-steps", type=int, default=3)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from dgraph_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.pop("NCCL_DEBUG", None)  # any NCCL_DEBUG level prints "NCCL version ..." to stdout before the JSON line
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    _lib.check(lib.dgx_init(local_rank))

    Q = args.queries
    queries = make_queries(Q, rank)
    uids_per_step = sum(l.size for qq in queries for l in qq)

    # ---- device-resident inputs (torch owns the HBM; libdgx gets raw pointers) -------
    # a real (non-default) torch stream: libdgx launches on it and torch.cuda.Event times it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    lane = lib.dgx_lane_create(local_rank, C.c_void_p(stream.cuda_stream))
    assert lane, lib.dgx_last_error()
    d_lists, ptrs, lens, k_off = [], [], [], [0]
    for qq in queries:
        for l in qq:
            t = torch.from_numpy(l.view(np.int64)).to(dev)
            d_lists.append(t)
            ptrs.append(t.data_ptr())
            lens.append(l.size)
        k_off.append(len(ptrs))
    nl = len(ptrs)
    c_ptrs = (C.c_void_p * nl)(*ptrs)
    c_lens = (C.c_size_t * nl)(*lens)
    c_koff = (C.c_size_t * (Q + 1))(*k_off)
    out_cap = sum(min(l.size for l in qq) for qq in queries)
    d_out = torch.empty(out_cap + 2, dtype=torch.int64, device=dev)
    d_off = torch.zeros(Q + 1, dtype=torch.int64, device=dev)
    # N > 1: the all-gatherv is ONE collective per step over preallocated buffers -- every rank sends
    # [count | results padded to a static bound] (no host sync, no per-step allocation).  It runs in
    # stream order after the filter launch (~25 us at N = 2).
    pad = min(int(out_cap), 1 << 16)
    send = torch.empty(1 + pad, dtype=torch.int64, device=dev)
    gathered = torch.empty(world * (1 + pad), dtype=torch.int64, device=dev)

    def step():
        _lib.check(lib.dgx_dev_filter_batch(lane, _lib.OP_INTERSECT, c_ptrs, c_lens, c_koff, Q,
                                            C.c_void_p(d_out.data_ptr()), out_cap, C.c_void_p(d_off.data_ptr())))
        if world > 1:
            torch.cat((d_off[Q:Q + 1], d_out[:pad]), out=send)
            dist.all_gather_into_tensor(gathered, send)

    # ---- parity check (outside the timed region) ------------------------------------
    from oracle import pyoracle as orc

    step()
    torch.cuda.synchronize()
    off = d_off.cpu().numpy().astype(np.uint64)
    res = d_out.cpu().numpy().view(np.uint64)
    want0 = orc.intersect_sorted(queries[0])
    wantl = orc.intersect_sorted(queries[-1])
    bit_exact = bool(np.array_equal(res[int(off[0]):int(off[1])], want0)
                     and np.array_equal(res[int(off[Q - 1]):int(off[Q])], wantl))
    out_uids = int(off[Q])
    if world > 1:  # the gathered block of this rank must be its own [count | results]
        mine = gathered[rank * (1 + pad):(rank + 1) * (1 + pad)].cpu().numpy().view(np.uint64)
        npay = min(out_uids, pad)
        bit_exact = bit_exact and int(mine[0]) == out_uids and bool(np.array_equal(mine[1:1 + npay], res[:npay]))

    # ---- device-resident timing --------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    for _ in range(200):  # NVML init can take a while when 8 ranks start at once
        if sampler.ok or hasattr(sampler, "err"):
            break
        time.sleep(0.01)
    sampler.active = True
    launches0 = lib.dgx_lane_launches(lane)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    # barrier + synchronize immediately before the first timed launch: a rank that enters the loop late
    # (even by one 10 ms sleep above) would make its peers' first gather wait inside their timed region
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    evs[0].record()
    host_t0 = time.perf_counter()
    for i in range(args.steps):
        step()
        evs[i + 1].record()
    host_submit_ms = (time.perf_counter() - host_t0) * 1e3 / args.steps  # host time to enqueue one step
    torch.cuda.synchronize()
    sampler.active = False
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    launches = int(lib.dgx_lane_launches(lane) - launches0)
    total_ms = evs[0].elapsed_time(evs[-1])
    step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        tot = torch.tensor([uids_per_step, launches], dtype=torch.int64, device=dev)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        uids_all, launches_all = int(tot[0].item()), int(tot[1].item())
    else:
        uids_all, launches_all = uids_per_step, launches
    ms_per_step = total_ms / args.steps
    value = uids_all / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (filter_pipe_kernel: one launch per step, ~90 % of it) -----------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    algo_bytes = 8.0 * (uids_per_step + out_uids)  # 8*(sum|L_i| + |Out|), SURVEY 8(d)
    kern_ms = float(np.median(step_ms)) if world == 1 else ms_per_step
    achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "filter_kernel_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            traffic = None
    roofline = {"bound": "hbm", "kernel": "dgx::filter_pipe_kernel (timed with its plan pre-pass: filter_tiles_kernel + filter_plan_kernel)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": kern_ms}

    # ---- end to end through the host-pointer C ABI (pinned host lists, copies timed) ------
    e2e = None
    cpu_baseline = None
    if True:
        host_lists = []
        for qq in queries:
            hq = []
            for l in qq:
                p = lib.dgx_host_alloc(l.nbytes)
                assert p
                C.memmove(p, l.ctypes.data, l.nbytes)
                hq.append((p, l.size))
            host_lists.append(hq)
        res_cap = max(min(l.size for l in qq) for qq in queries)
        h_out = lib.dgx_host_alloc(res_cap * 8)
        n_out = C.c_size_t(0)
        tables = []
        for hq in host_lists:
            tables.append(((C.c_void_p * K_LISTS)(*[p for p, _ in hq]), (C.c_size_t * K_LISTS)(*[n for _, n in hq])))

        # the reference is called from many goroutines at once (x.DivideAndRule, worker/task.go:816);
        # here E2E_THREADS host threads issue the queries, each call borrowing its own lane, so one
        # query's sync / D2H gaps are filled by another query's H2D copies.
        E2E_THREADS = 4
        outs = [(lib.dgx_host_alloc(res_cap * 8), C.c_size_t(0)) for _ in range(E2E_THREADS)]
        pool = ThreadPoolExecutor(max_workers=E2E_THREADS)

        def one_query(args_):
            slot, (tp, tl) = args_
            buf, cnt = outs[slot]
            _lib.check(lib.dgx_intersect_sorted(tp, tl, K_LISTS, buf, res_cap, C.byref(cnt)))
            return cnt.value

        def e2e_step():
            # queries i, i+T, i+2T, ... run on thread i (each thread owns one output buffer)
            def worker(slot):
                tot = 0
                for qi in range(slot, len(tables), E2E_THREADS):
                    tot += one_query((slot, tables[qi]))
                return tot
            return sum(pool.map(worker, range(E2E_THREADS)))

        e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            outn = e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": uids_all * args.e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(uids_per_step * 8), "d2h_bytes_per_step": int(outn * 8 + 8 * Q),
               "api": "dgx_intersect_sorted (host pointers, pinned lists), one call per query, 4 host threads (lane pool)",
               "ms_per_step": 1e3 * dt / args.e2e_steps}
        for hq in host_lists:
            for p, _ in hq:
                lib.dgx_host_free(p)
        lib.dgx_host_free(h_out)
        for buf, _ in outs:
            lib.dgx_host_free(buf)
        pool.shutdown()

    if rank == 0:
        # CPU baseline: the oracle (port of algo.IntersectSorted), 1 thread like the Go code
        sample = queries[: min(Q, 8)]
        cpu_intersect_batch(orc, sample[:1], 1)
        reps, t_cpu = 0, 0.0
        while t_cpu < 2.0 and reps < 8:
            t_cpu += cpu_intersect_batch(orc, sample, 1)
            reps += 1
        s_uids = sum(l.size for qq in sample for l in qq)
        cpu_baseline = {"value": s_uids * reps / t_cpu, "unit": UNIT, "cores": 1, "kind": "port",
                        "sample": f"{len(sample)} of the step's {Q} queries x {reps} repetitions, single thread "
                                  "(algo.IntersectSorted is single-threaded in the reference)",
                        "host": f"{os.cpu_count()} logical CPUs"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": workload_config(Q, world),
            "bit_exact": bit_exact, "out_uids_per_step": out_uids, "host_submit_ms_per_step": round(host_submit_ms, 4),
            "step_ms_rank0": {"median": round(float(np.median(step_ms)), 4), "p10": round(float(np.percentile(step_ms, 10)), 4),
                              "p90": round(float(np.percentile(step_ms, 90)), 4), "max": round(float(np.max(step_ms)), 4)},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": launches_all, "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
