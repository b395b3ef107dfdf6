// probe_kernel.cuh -- batched algo.IndexOf (algo/uidlist.go:546-552).
//
// updateDestUids / updateFacetMatrix / updateUidMatrix (query/query.go:1396-1438, 2594-2608) probe
// ONE sorted list (sg.DestUIDs) with every uid of a uid matrix, one sort.Search per uid.  Here the
// probes of a whole matrix are one launch: a thread per uid, binary search through the read-only
// cache (the upper levels of the search tree are shared by every thread and stay in L1/L2, so the
// HBM traffic is about one sector per probe).  Bound: L2/HBM latency, not bandwidth.
#pragma once

#include "common.cuh"

namespace dgx {

// idx[i] = first position of q[i] in the ascending list u[0..n), or -1 when absent.
// q need not be sorted (the matrix rows of the callers are in sort order, not uid order).
__global__ void __launch_bounds__(256) index_of_kernel(const u64* __restrict__ u, u64 n, const u64* __restrict__ q,
                                                       u64 m, long long* __restrict__ idx) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x) {
        const u64 x = ld_stream(q + i);
        const u64 p = lower_bound_g(u, n, x);
        idx[i] = (p < n && ld_probe(u + p) == x) ? (long long)p : -1ll;
    }
}

}  // namespace dgx
