// compressed_kernel.cuh -- algo.IntersectCompressedWith fused on the device, and the Decoder's seeks.
//
// algo.IntersectCompressedWith(pack, afterUID, v, o) (algo/uidlist.go:33-138) intersects a UidPack with a
// plain list WITHOUT decoding the whole pack: the reference walks block bases (Decoder.SeekToBlock /
// LinearSeek, codec/codec.go:219-271, 349-359) and unpacks only blocks that can hold an element of v.
// Same idea here, data-parallel:
//   icw_ranges_*     block b can only match v[j] with base[b] <= v[j] < base[b+1]: the v-range of every
//                    block (searching the block bases with v, or v with the bases, whichever side is
//                    shorter).  Blocks with an empty range are never touched again.
//   icw_probe_kernel a warp owns 64 consecutive blocks (as in decode_kernel): per batch of 32 it TMA-copies
//                    only the delta bytes between the first and last touched block, walks the tags of
//                    touched blocks, expands each into the warp's shared-memory row and probes it with
//                    the block's slice of v (or probes that slice with the row when the slice is the
//                    longer side).  A match sets keep[j]; nothing decoded ever reaches HBM.
//   compact_kernel   o.Uids = v[keep], order preserved, single pass (decoupled look-back).
// Defined, like the reference's own tests (algo/uidlist_test.go:607-681), for duplicate-free inputs.
#pragma once

#include "decode_kernel.cuh"

namespace dgx {

struct IcwParams {
    DPack pk;
    const u64* v;      // the plain list, ascending
    u64 m;
    u64 after;         // Decoder.Seek(after, SeekStart): only pack uids >= after take part (0: all)
    u32* vlo;          // per block: first index of v inside the block's value range ...
    u32* vhi;          // ... and one past the last (vlo == vhi: block not touched); zero-initialised
    unsigned char* keep;  // per v element: 1 when present in the pack; zero-initialised
};

// Thread per element of v: the block whose range holds v[j]; run boundaries write vlo / vhi.
__global__ void __launch_bounds__(256) icw_ranges_by_v_kernel(const IcwParams P) {
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.m) return;
    const u64 x = P.v[j];
    if (x < P.after) return;
    const u64 nb = P.pk.nblocks;
    // block of x: last b with base[b] <= x
    auto block_of = [&](u64 y) -> long long {
        u64 lo = 0, hi = nb;  // first b with base[b] > y
        while (lo < hi) {
            const u64 mid = lo + ((hi - lo) >> 1);
            if (ld_probe(P.pk.base + mid) <= y) lo = mid + 1; else hi = mid;
        }
        return (long long)lo - 1;
    };
    const long long b = block_of(x);
    if (b < 0) return;
    long long bp = -2, bn = -2;
    if (j > 0) { const u64 xp = P.v[j - 1]; if (xp >= P.after) bp = block_of(xp); }
    if (j + 1 < P.m) bn = block_of(P.v[j + 1]);
    if (bp != b) P.vlo[b] = (u32)j;
    if (bn != b) P.vhi[b] = (u32)(j + 1);
}

// Thread per block: [lower_bound(v, max(base[b], after)), lower_bound(v, base[b+1])).
__global__ void __launch_bounds__(256) icw_ranges_by_block_kernel(const IcwParams P) {
    const u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= P.pk.nblocks) return;
    const u64 first = P.pk.base[b];
    const u64 lo = lower_bound_g(P.v, P.m, first > P.after ? first : P.after);
    const u64 hi = (b + 1 < P.pk.nblocks) ? lower_bound_g(P.v, P.m, P.pk.base[b + 1]) : P.m;
    if (hi > lo) { P.vlo[b] = (u32)lo; P.vhi[b] = (u32)hi; }
}

__global__ void __launch_bounds__(D_NT) icw_probe_kernel(const IcwParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    DWarpSmem& S = reinterpret_cast<DWarpSmem*>(smem_raw)[wid];
    const DPack& pk = P.pk;
    const u64 gw = (u64)blockIdx.x * D_WARPS + wid;
    u64 b = gw * D_BPW;
    const u64 bend = (b + D_BPW < pk.nblocks) ? b + D_BPW : pk.nblocks;
    if (b >= bend) return;
    const u32 bar = smem_u32(&S.mbar);
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    u32 parity = 0;

    while (b < bend) {
        const u64 bi = b + lane;
        const bool in = bi < bend;
        const u32 lo = in ? P.vlo[bi] : 0u, hi = in ? P.vhi[bi] : 0u;
        const unsigned touched = __ballot_sync(0xffffffffu, in && hi > lo);
        if (touched == 0) { b += 32; continue; }  // nothing of v falls into these 32 blocks
        const int f = __ffs(touched) - 1;          // first touched block of the batch
        const u64 doff = in ? pk.delta_off[bi] : 0;
        const u64 dnext = in ? pk.delta_off[bi + 1] : 0;
        const u64 win0 = __shfl_sync(0xffffffffu, doff, f) & ~15ull;
        // blocks f .. l whose bytes fit the window together (l >= f: one block is at most 1088 B at BlockSize 256)
        const bool fits = in && lane >= f && (dnext - win0 <= (u64)D_WB);
        const unsigned fm = __ballot_sync(0xffffffffu, fits) >> f;
        int run = __ffs(~fm) - 1;                  // leading run of fitting blocks, counted from f
        if (fm == (0xffffffffu >> f)) run = 32 - f;
        const u32 num = in ? pk.num[bi] : 0;
        const bool oversize = run == 0 || (num > 256u && lane >= f && lane < f + (run ? run : 1));
        if (__any_sync(0xffffffffu, oversize)) {
            // Generic slow path (non-reference BlockSize): lane-serial walk of the touched blocks from global
            // memory, every decoded uid binary-searched in the block's slice of v.
            const int r = run ? run : 1;
            if (in && lane >= f && lane < f + r && hi > lo) {
                const unsigned char* p = pk.deltas + doff;
                u64 last = pk.base[bi];
                u32 produced = 1;
                auto probe = [&](u64 y) {
                    const u64 pos = lo + lower_bound_g(P.v + lo, (u64)(hi - lo), y);
                    if (pos < hi && ld_probe(P.v + pos) == y) P.keep[pos] = 1;
                };
                probe(last);
                while (produced < num) {
                    const u32 tag = p[0];
                    const unsigned char* d = p + 1;
                    for (int j = 0; j < 4 && produced < num; ++j) {
                        const u32 l = ((tag >> (2 * j)) & 3u) + 1u;
                        u32 v32 = 0;
                        for (u32 k = 0; k < l; ++k) v32 |= (u32)d[k] << (8 * k);
                        d += l;
                        last += v32;
                        ++produced;
                        probe(last);
                    }
                    p += gv_bytes_used(tag);
                }
            }
            b += f + r;
            continue;
        }
        // last touched block inside the run: nothing behind it needs copying
        const unsigned in_run = (run >= 32 ? 0xffffffffu : ((1u << run) - 1u)) << f;
        const unsigned want = touched & in_run;
        const int last_t = 31 - __clz(want);
        const u64 wend = __shfl_sync(0xffffffffu, dnext, last_t);
        const u64 base = in ? pk.base[bi] : 0;
        const u32 ngroups = num > 1 ? (num + 2) / 4 : 0;

        const u32 bytes = (u32)(((wend - win0) + 15ull) & ~15ull);
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar, bytes);
            tma_bulk_g2s(smem_u32(S.payload), pk.deltas + win0, bytes, bar);
        }
        mbar_wait(bar, parity);
        parity ^= 1u;

        if ((want >> lane) & 1u) {  // tag walk of the touched blocks only
            u32 off = (u32)(doff - win0);
            const u32 off_end = (u32)(dnext - win0);
            for (u32 g = 0; g < ngroups; ++g) {
                S.goff[g * D_TSTRIDE + lane] = (unsigned short)off;
                off += gv_bytes_used(S.payload[off]);
                off = off < off_end ? off : off_end;
            }
        }
        __syncwarp();

        const u32* pl32 = reinterpret_cast<const u32*>(S.payload);
        for (unsigned rest = want; rest; rest &= rest - 1) {
            const int i = __ffs(rest) - 1;
            const u32 n_i = __shfl_sync(0xffffffffu, num, i);
            const u64 base_i = __shfl_sync(0xffffffffu, base, i);
            const u32 ng_i = __shfl_sync(0xffffffffu, ngroups, i);
            const u32 lo_i = __shfl_sync(0xffffffffu, lo, i), hi_i = __shfl_sync(0xffffffffu, hi, i);
            if (lane == 0) S.stage[0] = base_i;
            u64 carry = 0;
            for (u32 g0 = 0; g0 < ng_i; g0 += 32) {
                const u32 g = g0 + lane;
                u32 d0 = 0, d1 = 0, d2 = 0, d3 = 0;
                if (g < ng_i) {
                    const u32 off = S.goff[g * D_TSTRIDE + i];
                    const u32 tag = S.payload[off];
                    const u32 a = off + 1;
                    const u32 wb = a >> 2, sh = (a & 3u) * 8u;
                    const u32 w0 = pl32[wb], w1 = pl32[wb + 1], w2 = pl32[wb + 2], w3 = pl32[wb + 3], w4 = pl32[wb + 4];
                    const u32 D0 = __funnelshift_r(w0, w1, sh), D1 = __funnelshift_r(w1, w2, sh);
                    const u32 D2 = __funnelshift_r(w2, w3, sh), D3 = __funnelshift_r(w3, w4, sh);
                    const u32 l0 = (tag & 3u) + 1u, l1 = ((tag >> 2) & 3u) + 1u;
                    const u32 l2 = ((tag >> 4) & 3u) + 1u, l3 = (tag >> 6) + 1u;
                    auto extract = [&](u32 s, u32 l) -> u32 {
                        const u32 wi = s >> 2, sh2 = (s & 3u) * 8u;
                        const u32 lo_ = wi == 0 ? D0 : (wi == 1 ? D1 : (wi == 2 ? D2 : D3));
                        const u32 hi_ = wi == 0 ? D1 : (wi == 1 ? D2 : (wi == 2 ? D3 : 0u));
                        return __funnelshift_r(lo_, hi_, sh2) & (0xffffffffu >> (32u - 8u * l));
                    };
                    d0 = extract(0, l0);
                    d1 = extract(l0, l1);
                    d2 = extract(l0 + l1, l2);
                    d3 = extract(l0 + l1 + l2, l3);
                }
                const u64 p1 = d0, p2 = p1 + d1, p3 = p2 + d2, p4 = p3 + d3;
                u64 inc = p4;
#pragma unroll
                for (int dlt = 1; dlt < 32; dlt <<= 1) {
                    u64 t = __shfl_up_sync(0xffffffffu, inc, dlt);
                    if (lane >= dlt) inc += t;
                }
                const u64 ex = base_i + carry + inc - p4;
                if (g < ng_i) {
                    const u32 idx = 1u + 4u * g;
                    S.stage[idx] = ex + p1;
                    S.stage[idx + 1] = ex + p2;
                    S.stage[idx + 2] = ex + p3;
                    S.stage[idx + 3] = ex + p4;
                }
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
            __syncwarp();
            // ---- probe: the shorter side searches the longer one ----------------------------------
            const u32 cnt = hi_i - lo_i;
            if (cnt <= 4u * n_i) {
                for (u32 j = lo_i + lane; j < hi_i; j += 32) {
                    const u64 x = ld_stream(P.v + j);
                    const int pos = lower_bound_s(S.stage, 0, (int)n_i, x);
                    if (pos < (int)n_i && S.stage[pos] == x) P.keep[j] = 1;
                }
            } else {
                for (u32 e = lane; e < n_i; e += 32) {
                    const u64 y = S.stage[e];
                    const u64 pos = lo_i + lower_bound_g(P.v + lo_i, (u64)cnt, y);
                    if (pos < hi_i && ld_probe(P.v + pos) == y) P.keep[pos] = 1;
                }
            }
            __syncwarp();
        }
        b += f + run;
    }
}

// ---- order-preserving compaction of v by keep[] (single pass, decoupled look-back) ------------
constexpr int CP_NT = 256, CP_VT = 8, CP_TILE = CP_NT * CP_VT;
struct CompactParams {
    const u64* v;
    const unsigned char* keep;
    u64 m;
    u64* out;
    u64 out_cap;
    u64* out_len;
    u64* status;   // ceil(m / CP_TILE) words, zeroed
    u32* ticket;   // zeroed: tiles are numbered in the order their CTAs start (no look-back deadlock)
    int* err;
};
__global__ void __launch_bounds__(CP_NT) compact_kernel(const CompactParams P) {
    __shared__ u32 warp_sums[CP_NT / 32 + 1];
    __shared__ u64 s_base;
    __shared__ u32 s_tile;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(P.ticket, 1u);
    __syncthreads();
    const u32 tile = s_tile;
    const u64 t0 = (u64)tile * CP_TILE;
    // thread owns CP_VT consecutive elements: order inside the tile is thread order
    u32 flags = 0, cnt = 0;
    const u64 e0 = t0 + (u64)tid * CP_VT;
#pragma unroll
    for (int i = 0; i < CP_VT; ++i)
        if (e0 + i < P.m && P.keep[e0 + i]) { flags |= 1u << i; ++cnt; }
    u32 total;
    const u32 ex = block_exclusive_scan<CP_NT>(cnt, warp_sums, &total);
    if (wid == 0) {
        const u64 base = lookback_exclusive(P.status, tile, (u64)total, lane);
        if (lane == 0) {
            s_base = base;
            if (t0 + CP_TILE >= P.m) {
                *P.out_len = base + total;
                if (base + total > P.out_cap) atomicExch(P.err, 1);
            }
        }
    }
    __syncthreads();
    u64 w = s_base + ex;
#pragma unroll
    for (int i = 0; i < CP_VT; ++i)
        if ((flags >> i) & 1u) {
            if (w < P.out_cap) P.out[w] = P.v[e0 + i];
            ++w;
        }
}

// ---- Decoder seeks on the device (codec/codec.go:219-384) --------------------------------------
// One thread restates the control flow of the reference's iterator for ONE call and expands the block it
// lands on; what comes back is what the Go call returns: the uid slice and the decoder's new block index.
// State carried between calls is the block index only (the current block counts as fully unpacked).
enum { SK_SEEK = 0, SK_SEEK_TO_BLOCK = 1, SK_LINEAR_SEEK = 2, SK_NEXT = 3, SK_UNPACK = 4 };
struct SeekParams {
    DPack pk;
    u64 uid;
    int whence;     // 0 SeekStart (>=), 1 SeekCurrent (>)
    int kind;
    u64 block_idx;  // the decoder's blockIdx before the call
    u64* out;       // up to max_num uids
    u64* out_meta;  // [0] = length, [1] = blockIdx after the call
};
__device__ __forceinline__ u32 sk_unpack(const DPack& pk, u64 b, u64* out) {  // Decoder.UnpackBlock (:154-200)
    if (b >= pk.nblocks) return 0;
    const u32 num = pk.num[b];
    const unsigned char* p = pk.deltas + pk.delta_off[b];
    u64 last = pk.base[b];
    u32 produced = 0;
    if (num) out[produced++] = last; else return 0;
    while (produced < num) {
        const u32 tag = p[0];
        const unsigned char* d = p + 1;
        for (int j = 0; j < 4 && produced < num; ++j) {
            const u32 l = ((tag >> (2 * j)) & 3u) + 1u;
            u32 v = 0;
            for (u32 k = 0; k < l; ++k) v |= (u32)d[k] << (8 * k);
            d += l;
            last += v;
            out[produced++] = last;
        }
        p += gv_bytes_used(tag);
    }
    return num;
}
__global__ void pack_seek_kernel(const SeekParams P) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const DPack& pk = P.pk;
    const u64 nb = pk.nblocks;
    const bool gt = P.whence == 1;
    // sort.Search over Blocks[from:] for Base >= uid (SeekStart) / Base > uid (SeekCurrent)
    auto search_base = [&](u64 from) -> u64 {
        u64 lo = from, hi = nb;
        while (lo < hi) {
            const u64 mid = lo + ((hi - lo) >> 1);
            const u64 bs = pk.base[mid];
            if (gt ? (bs > P.uid) : (bs >= P.uid)) hi = mid; else lo = mid + 1;
        }
        return lo;
    };
    u64 blk = 0, len = 0, start = 0;
    auto next = [&]() { blk += 1; len = sk_unpack(pk, blk, P.out); start = 0; };  // Decoder.Next (:370-376)
    switch (P.kind) {
        case SK_UNPACK: blk = P.block_idx; len = sk_unpack(pk, blk, P.out); break;
        case SK_NEXT: blk = P.block_idx; next(); break;
        case SK_LINEAR_SEEK: {  // :349-359: advance while seek >= PeekNextBase (MaxUint64 past the last block)
            blk = P.block_idx;
            while (true) {
                const u64 nxt = (blk + 1 < nb) ? pk.base[blk + 1] : kU64Max;
                if (P.uid < nxt || blk >= nb) break;  // (the reference would spin forever on seek == MaxUint64)
                blk += 1;
            }
            len = sk_unpack(pk, blk, P.out);
            break;
        }
        case SK_SEEK: {  // :279-337
            if (P.uid == 0) { len = sk_unpack(pk, 0, P.out); break; }
            const u64 idx = search_base(0);
            if (idx == 0) { len = sk_unpack(pk, 0, P.out); break; }
            if (idx < nb && pk.base[idx] == P.uid) { blk = idx; len = sk_unpack(pk, blk, P.out); break; }
            blk = idx - 1;
            len = sk_unpack(pk, blk, P.out);
            u64 u = 0;
            while (u < len && (gt ? (P.out[u] <= P.uid) : (P.out[u] < P.uid))) ++u;
            if (u < len) { start = u; len -= u; } else next();
            break;
        }
        default: {  // SK_SEEK_TO_BLOCK, :219-271
            u64 prev = P.block_idx;
            if (P.uid == 0) { len = sk_unpack(pk, 0, P.out); break; }
            if (prev > 0 && prev < nb && P.uid < pk.base[prev]) prev = 0;
            if (prev > nb) prev = nb;
            const u64 idx = search_base(prev);
            if (idx == 0) { len = sk_unpack(pk, 0, P.out); break; }
            if (idx < nb && pk.base[idx] == P.uid) { blk = idx; len = sk_unpack(pk, blk, P.out); break; }
            blk = idx - 1;
            // `if d.blockIdx != prevBlockIdx { d.UnpackBlock() }` (:262-264): when the search was restarted from
            // block 0 for an older uid and lands on block 0, the reference keeps the slice it already held --
            // the block the decoder stood on before the call -- and tests uid against THAT slice.
            len = sk_unpack(pk, blk != prev ? blk : P.block_idx, P.out);
            if (!(len > 0 && P.uid <= P.out[len - 1])) next();
            break;
        }
    }
    if (start) for (u64 i = 0; i < len; ++i) P.out[i] = P.out[start + i];
    P.out_meta[0] = len;
    P.out_meta[1] = blk;
}

}  // namespace dgx
