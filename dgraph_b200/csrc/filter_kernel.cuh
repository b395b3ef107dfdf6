// filter_kernel.cuh -- batched "A-driven" sorted-set filter for sm_100a.
//
// One kernel covers algo.IntersectWith (algo/uidlist.go:142), the whole chain of
// algo.IntersectSorted (:297-329, k lists in ONE pass), algo.Difference (:332)
// and the batched row-filter shape of worker/task.go:783 / query/query.go:1425.
//
// Work unit ("tile"): F_TA consecutive values of a query's driving list A (its
// shortest list for intersections, u for Difference).  The tile's candidates
// live in shared memory for the whole tile; every other list L_j of the query
// only contributes the slice [lower_bound(L_j, first), upper_bound(L_j, last))
// which is either streamed through shared memory in coalesced chunks (dense
// slice) or probed by per-candidate binary search (sparse candidates).
// Survivors are compacted in shared memory after every list, so intermediates
// of the k-way chain never touch HBM and a list is read at most once.
// Tile outputs are concatenated with a single-pass decoupled look-back, which
// also yields the CSR offsets of a batch.
//
// Duplicate semantics (pinned by algo/uidlist_test.go:329-348): a candidate that
// is the r-th copy of x in A survives list L_j iff L_j holds more than r copies
// of x, i.e. L_j[lower_bound(x) + r] == x -- multiset-min for intersections and
// multiset difference, exactly what the reference's two-pointer merges produce.
#pragma once

#include "common.cuh"

namespace dgx {

constexpr int F_NT = 256;            // threads per CTA
constexpr int F_VA = 4;              // candidates per thread (blocked)
constexpr int F_TA = F_NT * F_VA;    // candidates per tile
constexpr int F_CAP = 2048;          // values of L_j staged in shared memory at once
constexpr int F_LB = F_NT / 64;      // lists whose slice bounds are searched per batch (2 warps each)

struct FList {
    const u64* ptr;
    u64 len;             // length, or an upper bound when dyn_len != nullptr
    const u64* dyn_len;  // optional: device word holding the actual length
};
struct FTask {
    u64 tile_base;   // index of the task's first tile
    u32 list_first;  // lists[list_first] is A, followed by k-1 filter lists
    u32 k;
};
struct FParams {
    const FTask* tasks;
    const FList* lists;
    u32 ntasks;
    u32 ntiles;
    int op;               // 0 intersect, 1 difference
    u32 stream_ratio;     // stream a slice when slice_len <= stream_ratio * ncand + F_CAP
    u64* out;
    u64 out_cap;
    u64* out_off;         // ntasks + 1
    u64* status;          // ntiles words, zeroed
    u32* ticket;          // zeroed
    int* err;             // set to 1 on out_cap overflow
};

__device__ __forceinline__ u64 flist_len(const FList& L) {
    u64 n = L.len;
    if (L.dyn_len) {
        u64 d = ld_relaxed(L.dyn_len);
        if (d < n) n = d;
    }
    return n;
}

// Rank of candidate idx (value c) among equal values of A: copies before it in
// the tile plus, when the run reaches the tile start, copies before the tile.
__device__ __noinline__ u64 cand_rank(const u64* s_cand, int idx, u64 c, bool has_prev, u64 prev,
                                      const u64* A, u64 a0) {
    int s = idx;
    while (s > 0 && s_cand[s - 1] == c) --s;
    u64 r = (u64)(idx - s);
    if (s == 0 && has_prev && prev == c) r += a0 - lower_bound_g(A, a0, c);
    return r;
}

__global__ void __launch_bounds__(F_NT) filter_kernel(const FParams P) {
    __shared__ __align__(16) u64 s_cand[F_TA];
    __shared__ __align__(16) u64 s_b[F_CAP];
    __shared__ u64 s_r0[F_LB], s_r1[F_LB];
    __shared__ u64 s_prefix;
    __shared__ u32 s_scan[F_NT / 32 + 1];
    __shared__ u32 s_tile;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    if (tid == 0) s_tile = atomicAdd(P.ticket, 1u);
    __syncthreads();
    const u32 tile = s_tile;

    // tile -> task: last task with tile_base <= tile (uniform across the block)
    u32 q;
    {
        u32 lo = 0, hi = P.ntasks;
        while (hi - lo > 1) {
            u32 mid = (lo + hi) >> 1;
            if (P.tasks[mid].tile_base <= (u64)tile) lo = mid; else hi = mid;
        }
        q = lo;
    }
    const FTask T = P.tasks[q];
    const FList LA = P.lists[T.list_first];
    const u64 lenA = flist_len(LA);
    const u64* __restrict__ A = LA.ptr;
    const u64 a0 = (u64)(tile - T.tile_base) * F_TA;
    const int na = a0 < lenA ? (int)((lenA - a0 < (u64)F_TA) ? (lenA - a0) : (u64)F_TA) : 0;

    for (int i = tid; i < na; i += F_NT) s_cand[i] = ld_stream(A + a0 + i);
    const bool has_prev = (a0 > 0) && (na > 0);
    const u64 prev = has_prev ? ld_probe(A + a0 - 1) : 0;
    __syncthreads();

    int ncand = na;
    for (u32 j = 1; j < T.k && ncand > 0; ++j) {
        const int slot = (int)((j - 1) % F_LB);
        if (slot == 0) {
            // slice bounds for lists j .. j+F_LB-1: two warps per list, 32-ary searches
            const u64 lo = s_cand[0], hi = s_cand[ncand - 1];
            const int li = wid >> 1;
            if (j + li < T.k) {
                const FList Lq = P.lists[T.list_first + j + li];
                const bool upper = wid & 1;
                u64 r = warp_bound_g(Lq.ptr, flist_len(Lq), upper ? hi : lo, upper, lane);
                if (lane == 0) (upper ? s_r1 : s_r0)[li] = r;
            }
            __syncthreads();
        }
        const FList Lj = P.lists[T.list_first + j];
        const u64* __restrict__ B = Lj.ptr;
        const u64 lenB = flist_len(Lj);
        const u64 r0 = s_r0[slot], r1 = s_r1[slot];
        const u64 slice = r1 > r0 ? r1 - r0 : 0;

        // my candidates (blocked): indices [F_VA*tid, F_VA*tid + my_n)
        const int base_idx = F_VA * tid;
        const int my_n = ncand - base_idx < 0 ? 0 : (ncand - base_idx > F_VA ? F_VA : ncand - base_idx);
        u64 c[F_VA];
        {
            const ulonglong2* sp = reinterpret_cast<const ulonglong2*>(s_cand + base_idx);
            ulonglong2 v0 = sp[0], v1 = sp[1];
            c[0] = v0.x; c[1] = v0.y; c[2] = v1.x; c[3] = v1.y;
        }
        unsigned dup = 0;  // bit i: candidate i repeats the value before it
        if (my_n > 0) {
            if (base_idx > 0) { if (s_cand[base_idx - 1] == c[0]) dup |= 1u; }
            else if (has_prev && prev == c[0]) dup |= 1u;
        }
#pragma unroll
        for (int i = 1; i < F_VA; ++i)
            if (i < my_n && c[i] == c[i - 1]) dup |= 1u << i;

        unsigned hit = 0;
        if (slice == 0) {
            // nothing of L_j in range: no candidate matches
        } else if (slice <= (u64)P.stream_ratio * (u64)ncand + (u64)F_CAP) {
            // ---- stream the slice through shared memory -------------------------
            unsigned resolved = 0;
            for (u64 cs = r0; cs < r1; cs += F_CAP) {
                const int n = (int)((r1 - cs < (u64)F_CAP) ? (r1 - cs) : (u64)F_CAP);
                const bool last = cs + (u64)n >= r1;
                const u64* src = B + cs;
                const int head = (int)((reinterpret_cast<uintptr_t>(src) >> 3) & 1);
                if (tid == 0 && head) s_b[0] = ld_stream(src);
                const int npairs = (n - head) >> 1;
                for (int i = tid; i < npairs; i += F_NT) {
                    ulonglong2 v = ld_stream2(src + head + 2 * i);
                    s_b[head + 2 * i] = v.x;
                    s_b[head + 2 * i + 1] = v.y;
                }
                if (tid == 32 && ((n - head) & 1)) s_b[n - 1] = ld_stream(src + n - 1);
                __syncthreads();
                const u64 chunk_last = s_b[n - 1];
                int p = 0;
                bool first = true;
#pragma unroll
                for (int i = 0; i < F_VA; ++i) {
                    if (i < my_n && !((resolved >> i) & 1u) && (last || c[i] <= chunk_last)) {
                        const u64 x = c[i];
                        if (first) {
                            p = lower_bound_s(s_b, 0, n, x);
                            first = false;
                        } else {
#pragma unroll
                            for (int s = 0; s < 4; ++s)
                                if (p < n && s_b[p] < x) ++p;
                            if (p < n && s_b[p] < x) p = lower_bound_s(s_b, p + 1, n, x);
                        }
                        bool h;
                        if (!((dup >> i) & 1u)) {
                            h = (p < n) && (s_b[p] == x);
                        } else {
                            const u64 g = cs + (u64)p + cand_rank(s_cand, base_idx + i, x, has_prev, prev, A, a0);
                            h = (g < lenB) && (ld_probe(B + g) == x);
                        }
                        if (h) hit |= 1u << i;
                        resolved |= 1u << i;
                    }
                }
                if (!last) __syncthreads();  // everyone is done with this chunk
            }
        } else {
            // ---- sparse candidates: binary-search the slice in global memory -----
            u64 p = 0;
#pragma unroll
            for (int i = 0; i < F_VA; ++i) {
                if (i < my_n) {
                    const u64 x = c[i];
                    p += lower_bound_g(B + r0 + p, slice - p, x);
                    u64 g = r0 + p;
                    if ((dup >> i) & 1u) g += cand_rank(s_cand, base_idx + i, x, has_prev, prev, A, a0);
                    if ((g < lenB) && (ld_probe(B + g) == x)) hit |= 1u << i;
                }
            }
        }

        const unsigned valid = (1u << my_n) - 1u;
        const unsigned keep = (P.op == 0 ? hit : ~hit) & valid;
        u32 total;
        u32 off = block_exclusive_scan<F_NT>(__popc(keep), s_scan, &total);
        // (the scan's barriers order all reads of s_cand / s_b above before the writes below)
#pragma unroll
        for (int i = 0; i < F_VA; ++i)
            if ((keep >> i) & 1u) s_cand[off++] = c[i];
        __syncthreads();
        ncand = (int)total;
    }

    if (wid == 0) {
        u64 ex = lookback_exclusive(P.status, tile, (u64)ncand, lane);
        if (lane == 0) s_prefix = ex;
    }
    __syncthreads();
    const u64 base = s_prefix;
    if (tid == 0) {
        if ((u64)tile == T.tile_base) P.out_off[q] = base;
        if (tile == P.ntiles - 1) P.out_off[P.ntasks] = base + (u64)ncand;
    }
    if (base + (u64)ncand > P.out_cap) {
        if (tid == 0) atomicExch(P.err, 1);
    } else {
        for (int i = tid; i < ncand; i += F_NT) st_stream(P.out + base + i, s_cand[i]);
    }
}

}  // namespace dgx
