// merge_radix.cuh -- the tile engine of the single-pass multiway MergeSorted, second version.
//
// Same decomposition as merge_multi.cuh (samples -> distinct splitters -> per-(boundary, run) bounds -> one CTA per
// tile), but a tile is not MERGED any more, it is SORTED: all values of a tile lie in [lo, hi] with hi - lo < 2^32 for
// any realistic uid distribution, so the CTA loads the k slices as 32-bit offsets from lo, sorts them with an LSD
// radix sort in shared memory over only the ceil(log2(hi - lo + 1)) bits that can differ (warp-synchronous match
// ranking, per-warp digit counters, one block scan per pass), drops repeats and writes lo + key.  About 4
// warp-instructions per input value against 22 for six levels of pairwise merge-path merges.
//   mplan2_kernel   bounds, run-major (threads of a warp search neighbouring splitters in the SAME run, so their
//                   probes share sectors), each search confined to the gap between two samples of the run (the
//                   samples are already in HBM/L2: 9 steps instead of 24).
//   mmerge2_kernel  tiles numbered by an atomic ticket; a tile that fits one round knows its count after the
//                   de-duplication, obtains its output offset by decoupled look-back and writes its values straight
//                   to their final position (no scratch pass, no scan kernel, no compaction kernel).  Tiles that
//                   need several rounds (more than MR_C values) or whose span does not fit 32 bits (levels engine
//                   of merge_multi.cuh) go through the scratch buffer and are copied once their offset is known.
// algo.MergeSorted: algo/uidlist.go:448-542.
#pragma once

#include "merge_multi.cuh"

namespace dgx {

#ifndef DGX_MR_MATCH
#define DGX_MR_MATCH 1  // 1: hardware match.any; 0: one ballot per digit bit
#endif
constexpr int MR_NT = MM_NT;            // 512 threads
constexpr int MR_NW = MR_NT / 32;       // 16 warps
constexpr int MR_VT = 8;                // keys per thread and round
constexpr int MR_C = MR_NT * MR_VT;     // 4096 keys per round
constexpr int MR_DB = 8;                // widest digit (bits)
constexpr int MR_ND = 1 << MR_DB;
constexpr int MR_ROW = MR_ND + 1;       // counter row of one warp (+1 keeps the digit-major scan off one bank)
constexpr size_t MR_SMEM_RADIX = (size_t)2 * MR_C * sizeof(u32) + (size_t)MR_NW * MR_ROW * sizeof(u32);
constexpr size_t MR_SMEM_LEVELS = (size_t)2 * MM_CP * sizeof(u64);
constexpr size_t MR_SMEM = MR_SMEM_RADIX > MR_SMEM_LEVELS ? MR_SMEM_RADIX : MR_SMEM_LEVELS;
static_assert(MR_NT == 2 * MR_ND, "the counter scan maps thread t to digit t/2, warps (t%2)*8..+7");
static_assert(MR_NW == 16, "counter scan assumes 16 warps");

// ---- bounds, run-major: bounds[j * nbs + b] = position of boundary b in run j ---------------------------------
__global__ void __launch_bounds__(256) mplan2_kernel(const MMParams P) {
    const u32 ns = (u32)(*P.nsplit / P.stride);
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u32 j = (u32)(idx / P.nbs), b = (u32)(idx % P.nbs);
    if (j >= P.k || b > ns + 1) return;
    const u64* p; u64 n;
    mref_resolve(P.runs[j], p, n);
    u64 pos;
    if (b == 0 || n == 0) pos = 0;
    else if (b == ns + 1) pos = n;
    else {
        const u64 S = P.splitters[(u64)b * P.stride - 1];
        // the samples of run j sit at positions q(r) = (r + 1) * n / (s + 1) (msample_kernel): the samples below S
        // and the first one not below S fence the answer in
        const u64* sm = P.samples + P.samp_off[j];
        const u32 s = P.samp_off[j + 1] - P.samp_off[j];
        u32 l = 0, h = s;
        while (l < h) {
            const u32 m = (l + h) >> 1;
            if (ld_probe(sm + m) < S) l = m + 1; else h = m;
        }
        const u64 wlo = l > 0 ? (u64)(((unsigned __int128)l * n) / (s + 1)) + 1 : 0;
        const u64 whi = l < s ? (u64)(((unsigned __int128)(l + 1) * n) / (s + 1)) : n;
        pos = wlo + lower_bound_g(p + wlo, whi - wlo, S);
    }
    P.bounds[(u64)j * P.nbs + b] = pos;
}

// ---- LSD radix sort of n <= MR_C 32-bit keys in shared memory ----------------------------------------------------
// Sorts on bits [0, passes * w).  Keys start in `a`; the sorted keys end up in the returned buffer (a or b).
// Stable: position order = (warp, item, lane) with every warp owning a contiguous range of 32-key items, and the
// counters are scanned digit-major, warp-minor.  All MR_NT threads call; contains barriers.
__device__ __forceinline__ u32* mr_sort(u32* a, u32* b, u32* cnt, u32* s_scan, int n, int passes, int w) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int items = (n + 31) >> 5;
    const int ipw = (items + MR_NW - 1) / MR_NW;  // <= MR_VT
    const int item0 = wid * ipw;
    const u32 mask = (1u << w) - 1u;
    const int nd = 1 << w;
    u32* src = a;
    u32* dst = b;
    for (int ps = 0; ps < passes; ++ps) {
        const int shift = ps * w;
        for (int i = tid; i < MR_NW * MR_ROW; i += MR_NT) cnt[i] = 0;
        __syncthreads();
        u32 keys[MR_VT], loc[MR_VT];
        u32* myc = cnt + wid * MR_ROW;
#pragma unroll
        for (int i = 0; i < MR_VT; ++i) {
            keys[i] = 0; loc[i] = 0;
            const int item = item0 + i;
            if (i < ipw && item < items) {  // warp-uniform
                const int pos = item * 32 + lane;
                const bool act = pos < n;
                const u32 key = act ? src[pos] : 0u;
                const u32 dig = act ? ((key >> shift) & mask) : 0xFFFFFFFFu;
#if DGX_MR_MATCH
                const unsigned peers = __match_any_sync(0xffffffffu, dig);
#else
                unsigned peers = __ballot_sync(0xffffffffu, act);  // lanes with the same digit: one ballot per digit bit
#pragma unroll
                for (int bit = 0; bit < MR_DB; ++bit) {
                    const bool one = (dig >> bit) & 1u;
                    const unsigned m = __ballot_sync(0xffffffffu, one);
                    peers &= one ? m : ~m;
                }
#endif
                const int leader = __ffs(peers) - 1;
                u32 old = 0;
                if (act && lane == leader) { old = myc[dig]; myc[dig] = old + __popc(peers); }
                old = __shfl_sync(0xffffffffu, old, leader);
                keys[i] = key;
                loc[i] = old + __popc(peers & lt);
                __syncwarp();
            }
        }
        __syncthreads();
        {   // exclusive scan of the counters in (digit, warp) order: thread t owns digit t/2, warps (t%2)*8 .. +7
            const int d = tid >> 1, h8 = (tid & 1) * 8;
            u32 c[8], sum = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { c[q] = d < nd ? cnt[(h8 + q) * MR_ROW + d] : 0u; sum += c[q]; }
            u32 tot;
            u32 run = block_exclusive_scan<MR_NT>(sum, s_scan, &tot);
            if (d < nd) {
#pragma unroll
                for (int q = 0; q < 8; ++q) { cnt[(h8 + q) * MR_ROW + d] = run; run += c[q]; }
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < MR_VT; ++i) {
            const int item = item0 + i;
            if (i < ipw && item < items) {
                const int pos = item * 32 + lane;
                if (pos < n) dst[myc[(keys[i] >> shift) & mask] + loc[i]] = keys[i];
            }
        }
        __syncthreads();
        u32* t = src; src = dst; dst = t;
    }
    return src;
}

__global__ void __launch_bounds__(MR_NT, 3) mmerge2_kernel(const MMParams P) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ u64 s_cur[MM_K], s_end[MM_K];
    __shared__ const u64* s_ptr[MM_K];
    __shared__ u32 s_roff[MM_K], s_take[MM_K];
    __shared__ u64 s_r[2][6];
    __shared__ u32 s_cc[MR_VT * MR_NW];
    __shared__ u32 s_scan[MR_NT / 32 + 1];
    __shared__ u64 s_base;
    __shared__ u32 s_tile, s_rtot;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    if (tid == 0) s_tile = atomicAdd(P.ticket, 1u);
    __syncthreads();
    const u32 b = s_tile;
    const u32 ns = (u32)(*P.nsplit / P.stride);
    if (b > ns) return;
    const int k = (int)P.k;

    // ---- the tile's slice of every run; input offset, size, smallest and largest value ---------------------
    {
        u64 c = 0, len = 0, vmin = kU64Max, vmax = 0;
        if (tid < k) {
            const u64* p; u64 n;
            mref_resolve(P.runs[tid], p, n);
            c = P.bounds[(u64)tid * P.nbs + b];
            const u64 e = P.bounds[(u64)tid * P.nbs + b + 1];
            s_ptr[tid] = p; s_cur[tid] = c; s_end[tid] = e;
            len = e - c;
            if (len) { vmin = ld_probe(p + c); vmax = ld_probe(p + e - 1); }
        }
        if (wid < 2) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                c += __shfl_xor_sync(0xffffffffu, c, d);
                len += __shfl_xor_sync(0xffffffffu, len, d);
                const u64 a = __shfl_xor_sync(0xffffffffu, vmin, d);
                const u64 z = __shfl_xor_sync(0xffffffffu, vmax, d);
                vmin = a < vmin ? a : vmin;
                vmax = z > vmax ? z : vmax;
            }
            if (lane == 0) { s_r[wid][0] = c; s_r[wid][1] = len; s_r[wid][2] = vmin; s_r[wid][3] = vmax; }
        }
    }
    __syncthreads();
    const u64 inbase = s_r[0][0] + s_r[1][0];
    const u64 total = s_r[0][1] + s_r[1][1];
    const u64 lo = s_r[0][2] < s_r[1][2] ? s_r[0][2] : s_r[1][2];
    const u64 hi = s_r[0][3] > s_r[1][3] ? s_r[0][3] : s_r[1][3];
    __syncthreads();

    u32 produced = 0;        // values of this tile (in scratch unless `direct`)
    bool direct = false;     // already written to their final position
    if (total == 0) {
        // nothing: the tile still takes part in the look-back below
    } else if (hi - lo > 0xFFFFFFFFull) {
        produced = mm_levels_tile((u64*)s_raw, s_ptr, s_cur, s_end, k, P.scratch + inbase, false, 0);
    } else {
        u32* kA = (u32*)s_raw;
        u32* kB = kA + MR_C;
        u32* cnt = kB + MR_C;
        const u64 span = hi - lo;
        const int bits = span ? 64 - __clzll((long long)span) : 0;
        const int passes = (bits + MR_DB - 1) / MR_DB;
        const int w = passes ? (bits + passes - 1) / passes : 0;
        direct = total <= (u64)MR_C;
        // A value equal to a round's bound can have more copies waiting in runs that had loaded only part of them, so
        // the last value written is remembered across rounds.
        bool have_last = false;
        u32 lastkey = 0;
        for (;;) {
            // ---- how much of every run enters this round ----------------------------------------------------
            u64 rem = 0;
            if (tid < k) rem = s_end[tid] - s_cur[tid];
            u64 tot = rem;
            if (wid < 2) {
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, d);
                if (lane == 0) s_r[wid][0] = tot;
            }
            __syncthreads();
            const u64 total_rem = s_r[0][0] + s_r[1][0];
            if (total_rem == 0) break;
            const bool fits = total_rem <= (u64)MR_C;
            // A tile that does not fit is sorted in rounds: every run contributes a share of the chunk proportional
            // to what it has left (at least one value); only values <= the smallest "last loaded value of a run that
            // has more" are final in a round.
            u64 take = 0, bnd = kU64Max;
            if (tid < k) {
                if (fits) {
                    take = rem;
                } else if (rem > 0) {
                    const u64 share = (u64)(((unsigned __int128)rem * (u64)(MR_C - MM_K)) / total_rem);
                    take = share < 1 ? 1 : share;
                    if (take > rem) take = rem;
                    if (take < rem) bnd = ld_probe(s_ptr[tid] + s_cur[tid] + take - 1);
                }
            }
            if (wid < 2) {
                u32 inc = (u32)take;
                u64 mn = bnd;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const u32 v = __shfl_up_sync(0xffffffffu, inc, d);
                    if (lane >= d) inc += v;
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    const u64 o = __shfl_xor_sync(0xffffffffu, mn, d);
                    mn = o < mn ? o : mn;
                }
                s_roff[tid] = inc - (u32)take;  // exclusive inside the warp; the second warp adds the first's total below
                s_take[tid] = (u32)take;
                if (lane == 31) s_r[wid][1] = inc;
                if (lane == 0) s_r[wid][2] = mn;
            }
            __syncthreads();
            const u32 half = (u32)s_r[0][1];
            const int n = (int)(half + (u32)s_r[1][1]);
            const u64 bound = s_r[0][2] < s_r[1][2] ? s_r[0][2] : s_r[1][2];
            // ---- load the runs' contributions as offsets from lo (warp w: runs w, w + 16, ...) ------------------
            for (int r = wid; r < k; r += MR_NW) {
                const u32 o = s_roff[r] + (r >= 32 ? half : 0u);
                const int cntr = (int)s_take[r];
                const u64* src = s_ptr[r] + s_cur[r];
                for (int i = lane; i < cntr; i += 32) kA[o + i] = (u32)(ld_stream(src + i) - lo);
            }
            __syncthreads();
            const u32* Z = mr_sort(kA, kB, cnt, s_scan, n, passes, w);
            // ---- values <= bound are final this round -----------------------------------------------------------
            int nsafe = n;
            if (!fits) {
                const u32 bd = (u32)(bound - lo);
                int l = 0, h = n;
                while (l < h) {
                    const int m = (l + h) >> 1;
                    if (Z[m] <= bd) l = m + 1; else h = m;
                }
                nsafe = l;
            }
            // ---- drop repeats; position order is (j, thread): chunk (j, warp) holds 32 consecutive positions -------
            u32 kv[MR_VT];
            unsigned keepbits = 0;
#pragma unroll
            for (int j = 0; j < MR_VT; ++j) {
                const int pos = j * MR_NT + tid;
                bool keep = false;
                kv[j] = 0;
                if (pos < nsafe) {
                    const u32 v = Z[pos];
                    kv[j] = v;
                    keep = pos == 0 ? (!have_last || v != lastkey) : (Z[pos - 1] != v);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                if (keep) keepbits |= 1u << j;
                if (lane == 0) s_cc[j * MR_NW + wid] = __popc(bal);
            }
            if (nsafe > 0) { lastkey = Z[nsafe - 1]; have_last = true; }
            __syncthreads();
            if (wid == 0) {  // exclusive scan of the MR_VT * MR_NW = 128 chunk counts: four per lane
                u32 c0 = s_cc[lane * 4], c1 = s_cc[lane * 4 + 1], c2 = s_cc[lane * 4 + 2], c3 = s_cc[lane * 4 + 3];
                const u32 mine = c0 + c1 + c2 + c3;
                u32 inc = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const u32 v = __shfl_up_sync(0xffffffffu, inc, d);
                    if (lane >= d) inc += v;
                }
                const u32 ex = inc - mine;
                s_cc[lane * 4] = ex; s_cc[lane * 4 + 1] = ex + c0; s_cc[lane * 4 + 2] = ex + c0 + c1;
                s_cc[lane * 4 + 3] = ex + c0 + c1 + c2;
                const u32 rtot = __shfl_sync(0xffffffffu, inc, 31);
                if (lane == 0) s_rtot = rtot;
                if (direct) {  // one round: the tile's count is known, fetch its place in the output
                    const u64 base = lookback_exclusive(P.status, b, (u64)rtot, lane);
                    if (lane == 0) {
                        s_base = base;
                        if (b == ns) {
                            *P.out_len = base + rtot;
                            if (base + rtot > P.out_cap) atomicExch(P.err, 1);
                        }
                    }
                }
            }
            __syncthreads();
            {
                const u32 rtot = s_rtot;
                u64* dstp = direct ? P.out + s_base : P.scratch + inbase + produced;
                const u64 room = direct ? (P.out_cap > s_base ? P.out_cap - s_base : 0) : ~0ull;
#pragma unroll
                for (int j = 0; j < MR_VT; ++j) {
                    const bool keep = (keepbits >> j) & 1u;
                    const unsigned bal = __ballot_sync(0xffffffffu, keep);
                    if (keep) {
                        const u32 at = s_cc[j * MR_NW + wid] + __popc(bal & lt);
                        if (at < room) st_stream(dstp + at, lo + kv[j]);
                    }
                }
                produced += rtot;
            }
            // ---- advance the runs past everything that was final ---------------------------------------------------
            if (tid < k) {
                u64 adv = take;
                if (!fits) {
                    const u64* srcp = s_ptr[tid] + s_cur[tid];
                    u64 l = 0, h = take;  // first loaded value of the run that is > bound
                    while (l < h) {
                        const u64 m = l + ((h - l) >> 1);
                        if (ld_probe(srcp + m) <= bound) l = m + 1; else h = m;
                    }
                    adv = l;
                }
                s_cur[tid] += adv;
            }
            __syncthreads();
            if (fits) break;
        }
    }
    if (direct) return;
    // ---- tiles that went through the scratch buffer: place in the output, then one copy --------------------------
    if (wid == 0) {
        const u64 base = lookback_exclusive(P.status, b, (u64)produced, lane);
        if (lane == 0) {
            s_base = base;
            if (b == ns) {
                *P.out_len = base + produced;
                if (base + produced > P.out_cap) atomicExch(P.err, 1);
            }
        }
    }
    __syncthreads();
    {
        const u64 base = s_base;
        const u64* src = P.scratch + inbase;
        for (u32 i = tid; i < produced; i += MR_NT)
            if (base + i < P.out_cap) P.out[base + i] = src[i];
    }
}

}  // namespace dgx
