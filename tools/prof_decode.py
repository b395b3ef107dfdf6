#!/usr/bin/env python
"""Small driver for profiling codec.Decode under ncu: 2.5e7-UID pack, a few launches."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_ops as B
from dgraph_b200 import _lib
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen as hostgen
L = B.Lane()
gen = torch.Generator(device=B.DEV); gen.manual_seed(5)
n3 = int(sys.argv[1]) if len(sys.argv) > 1 else 25_000_000
master = B.zipf_gaps_gpu(n3, gen)
host = master.cpu().numpy().view(np.uint64)
_, base, num, doff_, deltas = hostgen.encode_pack_np_parallel(host, 256)
class _P: pass
pack = _P(); pack.nblocks = base.size
view = _lib.PackView()
view.block_size, view.nblocks = 256, pack.nblocks
view.base, view.num_uids, view.delta_off, view.deltas = base.ctypes.data, num.ctypes.data, doff_.ctypes.data, deltas.ctypes.data
pk = C.c_void_p()
_lib.check(L.lib.dgx_dev_pack_upload(L.h, C.byref(view), C.byref(pk)))
out = torch.empty(n3 + 8, dtype=torch.int64, device=B.DEV); out_len = torch.zeros(1, dtype=torch.int64, device=B.DEV)
def dec():
    _lib.check(L.lib.dgx_dev_decode(L.h, pk, 0, C.c_void_p(out.data_ptr()), n3, C.c_void_p(out_len.data_ptr())))
ms, _ = B.timeit(dec, warm=2, reps=5)
L.sync()
print("decode ms", ms, "ok", bool(torch.equal(out[:n3], master)), "GB/s", (deltas.size + 12 * pack.nblocks + 8 * n3) / ms / 1e6)
