/*
 * oracle.c -- CPU restatement of dgraph's algo/uidlist.go, algo/heap.go and
 * codec/codec.go (plus the group-varint primitives of go-groupvarint).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  The control flow of every function
 * follows the cited reference lines one decision at a time so that corner
 * semantics (duplicates, empty inputs, dispatch thresholds, block-split rule)
 * are the reference's and not a re-derivation.
 */
#include "oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================== */
/* algo/uidlist.go                                                           */
/* ======================================================================== */

#define ORC_JUMP 32           /* algo/uidlist.go:17 */
#define ORC_LIN_VS_BIN 10     /* algo/uidlist.go:18 */

/* algo/uidlist.go:170-191 */
void orc_intersect_with_lin(const uint64_t* u, size_t n, const uint64_t* v, size_t m,
                            uint64_t* out, size_t* olen, size_t* ri, size_t* rk) {
    size_t i = 0, k = 0, o = *olen;
    while (i < n && k < m) {
        uint64_t uid = u[i];
        uint64_t vid = v[k];
        if (uid > vid) {
            for (k = k + 1; k < m && v[k] < uid; k++) {
            }
        } else if (uid == vid) {
            out[o++] = uid;
            k++;
            i++;
        } else {
            for (i = i + 1; i < n && u[i] < vid; i++) {
            }
        }
    }
    *olen = o;
    if (ri) *ri = i;
    if (rk) *rk = k;
}

/* algo/uidlist.go:195-220 */
void orc_intersect_with_jump(const uint64_t* u, size_t n, const uint64_t* v, size_t m,
                             uint64_t* out, size_t* olen, size_t* ri, size_t* rk) {
    size_t i = 0, k = 0, o = *olen;
    while (i < n && k < m) {
        uint64_t uid = u[i];
        uint64_t vid = v[k];
        if (uid == vid) {
            out[o++] = uid;
            k++;
            i++;
        } else if (k + ORC_JUMP < m && uid > v[k + ORC_JUMP]) {
            k += ORC_JUMP;
        } else if (i + ORC_JUMP < n && vid > u[i + ORC_JUMP]) {
            i += ORC_JUMP;
        } else if (uid > vid) {
            for (k = k + 1; k < m && v[k] < uid; k++) {
            }
        } else {
            for (i = i + 1; i < n && u[i] < vid; i++) {
            }
        }
    }
    *olen = o;
    if (ri) *ri = i;
    if (rk) *rk = k;
}

/* sort.Search(n, f) with f(i) = a[i] >= val */
static size_t search_ge(const uint64_t* a, size_t n, uint64_t val) {
    size_t lo = 0, hi = n;
    while (lo < hi) {
        size_t h = lo + (hi - lo) / 2;
        if (!(a[h] >= val)) lo = h + 1; else hi = h;
    }
    return lo;
}
/* sort.Search(n, f) with f(i) = a[i] > val */
static size_t search_gt(const uint64_t* a, size_t n, uint64_t val) {
    size_t lo = 0, hi = n;
    while (lo < hi) {
        size_t h = lo + (hi - lo) / 2;
        if (!(a[h] > val)) lo = h + 1; else hi = h;
    }
    return lo;
}

/* binIntersect, algo/uidlist.go:254-288.  NOTE: len(d) >= len(q) must hold. */
static void bin_intersect(const uint64_t* d, size_t ld, const uint64_t* q, size_t lq,
                          uint64_t* out, size_t* olen) {
    if (ld == 0 || lq == 0) return;
    size_t midq = lq / 2;
    uint64_t qval = q[midq];
    size_t midd = search_ge(d, ld, qval);

    /* dd = d[0:midd], qq = q[0:midq] */
    if (midd > midq) bin_intersect(d, midd, q, midq, out, olen);
    else bin_intersect(q, midq, d, midd, out, olen);

    if (midd >= ld) return;
    if (d[midd] == qval) {
        out[(*olen)++] = qval;
    } else {
        midd--; /* may wrap to SIZE_MAX exactly like Go's -1; midd+1 restores 0 */
    }
    const uint64_t* dd = d + (midd + 1);
    size_t ldd = ld - (midd + 1);
    const uint64_t* qq = q + (midq + 1);
    size_t lqq = lq - (midq + 1);
    if (ldd > lqq) bin_intersect(dd, ldd, qq, lqq, out, olen);
    else bin_intersect(qq, lqq, dd, ldd, out, olen);
}

/* algo/uidlist.go:226-250 */
size_t orc_intersect_with_bin(const uint64_t* d, size_t ld, const uint64_t* q, size_t lq,
                              uint64_t* out, size_t* olen) {
    if (ld < lq) {
        size_t t = ld; ld = lq; lq = t;
        const uint64_t* tp = d; d = q; q = tp;
    }
    if (ld == 0 || lq == 0 || d[ld - 1] < q[0] || q[lq - 1] < d[0]) return 0;
    uint64_t val = d[0];
    size_t minq = search_ge(q, lq, val);
    val = d[ld - 1];
    size_t maxq = search_gt(q, lq, val);
    bin_intersect(d, ld, q + minq, maxq - minq, out, olen);
    return maxq;
}

/* algo/uidlist.go:156-165 */
int orc_intersect_with_branch(size_t n, size_t m) {
    if (n > m) { size_t t = n; n = m; m = t; }
    if (n == 0) n = 1;
    double ratio = (double)m / (double)n;
    if (ratio < 100) return 0;
    if (ratio < 500) return 1;
    return 2;
}

/* algo/uidlist.go:142-167.  dst := o.Uids[:0]; every branch only ever writes
 * position p after having read positions >= p of u, so out may alias u. */
size_t orc_intersect_with(const uint64_t* u, size_t n, const uint64_t* v, size_t m, uint64_t* out) {
    size_t olen = 0;
    switch (orc_intersect_with_branch(n, m)) {
    case 0: orc_intersect_with_lin(u, n, v, m, out, &olen, NULL, NULL); break;
    case 1: orc_intersect_with_jump(u, n, v, m, out, &olen, NULL, NULL); break;
    default:
        if (out == u) {
            /* the Go code appends into u's storage while recursing over u; the
             * recursion emits in increasing order and never ahead of its read
             * cursor for unique inputs, but to stay independent of that we
             * buffer. */
            size_t cap = n < m ? n : m;
            uint64_t* tmp = (uint64_t*)malloc((cap ? cap : 1) * sizeof(uint64_t));
            orc_intersect_with_bin(u, n, v, m, tmp, &olen);
            memcpy(out, tmp, olen * sizeof(uint64_t));
            free(tmp);
        } else {
            orc_intersect_with_bin(u, n, v, m, out, &olen);
        }
    }
    return olen;
}

typedef struct { const uint64_t* l; size_t length; } list_info; /* :290-293 */

static int cmp_list_info(const void* a, const void* b) {
    size_t x = ((const list_info*)a)->length, y = ((const list_info*)b)->length;
    return (x > y) - (x < y);
}

/* algo/uidlist.go:297-329 */
size_t orc_intersect_sorted(const uint64_t* const* lists, const size_t* lens, size_t k, uint64_t* out) {
    if (k == 0) return 0;
    list_info* ls = (list_info*)malloc(k * sizeof(list_info));
    for (size_t i = 0; i < k; i++) { ls[i].l = lists[i]; ls[i].length = lens[i]; }
    qsort(ls, k, sizeof(list_info), cmp_list_info); /* sort.Slice: unstable, by length */
    size_t olen;
    if (k == 1) {
        memcpy(out, ls[0].l, ls[0].length * sizeof(uint64_t));
        olen = ls[0].length;
        free(ls);
        return olen;
    }
    olen = orc_intersect_with(ls[0].l, ls[0].length, ls[1].l, ls[1].length, out);
    for (size_t i = 2; i < k; i++) {
        olen = orc_intersect_with(out, olen, ls[i].l, ls[i].length, out);
        if (olen == 0) break;
    }
    free(ls);
    return olen;
}

/* algo/uidlist.go:332-362 */
size_t orc_difference(const uint64_t* u, size_t n, const uint64_t* v, size_t m, uint64_t* out) {
    size_t o = 0, i = 0, k = 0;
    while (i < n && k < m) {
        uint64_t uid = u[i];
        uint64_t vid = v[k];
        if (uid < vid) {
            while (i < n && u[i] < vid) {
                out[o++] = u[i];
                i++;
            }
        } else if (uid == vid) {
            i++;
            k++;
        } else {
            for (k = k + 1; k < m && v[k] < uid; k++) {
            }
        }
    }
    while (i < n && k >= m) {
        out[o++] = u[i];
        i++;
    }
    return o;
}

/* ---- algo/heap.go:12-37 + container/heap -------------------------------- */

typedef struct { uint64_t val; size_t list_idx; } heap_elem;
typedef struct { heap_elem* e; size_t len; } u64_heap;

static int heap_less(const u64_heap* h, size_t i, size_t j) { return h->e[i].val < h->e[j].val; }
static void heap_swap(u64_heap* h, size_t i, size_t j) { heap_elem t = h->e[i]; h->e[i] = h->e[j]; h->e[j] = t; }
static void heap_up(u64_heap* h, size_t j) {
    for (;;) {
        if (j == 0) break;
        size_t i = (j - 1) / 2; /* parent */
        if (i == j || !heap_less(h, j, i)) break;
        heap_swap(h, i, j);
        j = i;
    }
}
static int heap_down(u64_heap* h, size_t i0, size_t n) {
    size_t i = i0;
    for (;;) {
        size_t j1 = 2 * i + 1;
        if (j1 >= n) break;
        size_t j = j1;
        size_t j2 = j1 + 1;
        if (j2 < n && heap_less(h, j2, j1)) j = j2;
        if (!heap_less(h, j, i)) break;
        heap_swap(h, i, j);
        i = j;
    }
    return i > i0;
}
static void heap_push(u64_heap* h, heap_elem x) { h->e[h->len++] = x; heap_up(h, h->len - 1); }
static void heap_pop(u64_heap* h) { size_t n = h->len - 1; heap_swap(h, 0, n); heap_down(h, 0, n); h->len--; }
static void heap_fix0(u64_heap* h) { if (!heap_down(h, 0, h->len)) heap_up(h, 0); }

/* internalMergeSortWithBuffer, algo/uidlist.go:392-433 */
static size_t internal_merge_with_buffer(const uint64_t* const* lists, const size_t* lens, size_t k,
                                         uint64_t* output) {
    if (k == 0) return 0;
    u64_heap h;
    h.e = (heap_elem*)malloc(k * sizeof(heap_elem));
    h.len = 0;
    for (size_t i = 0; i < k; i++) {
        if (lists[i] == NULL || lens[i] == 0) continue;
        heap_elem x; x.val = lists[i][0]; x.list_idx = i;
        heap_push(&h, x);
    }
    size_t olen = 0;
    size_t* idx = (size_t*)calloc(k, sizeof(size_t));
    uint64_t last = 0;
    while (h.len > 0) {
        heap_elem me = h.e[0];
        if (olen == 0 || me.val != last) {
            output[olen++] = me.val;
            last = me.val;
        }
        const uint64_t* l = lists[me.list_idx];
        if (idx[me.list_idx] + 1 >= lens[me.list_idx]) { /* idx >= len-1 */
            heap_pop(&h);
        } else {
            idx[me.list_idx]++;
            h.e[0].val = l[idx[me.list_idx]];
            heap_fix0(&h);
        }
    }
    free(idx);
    free(h.e);
    return olen;
}

/* internalMergeSort, algo/uidlist.go:436-446 */
size_t orc_internal_merge_sort(const uint64_t* const* lists, const size_t* lens, size_t k, uint64_t* out) {
    return internal_merge_with_buffer(lists, lens, k, out);
}

#define ORC_NUM_THREADS 10          /* algo/uidlist.go:462 */
#define ORC_NUM_LIST_PER_THREAD 10  /* algo/uidlist.go:463 */

typedef struct {
    const uint64_t* const* lists; const size_t* lens; size_t k;
    uint64_t* buf; size_t result_len; int ran;
} merge_job;

static void* merge_worker(void* arg) {
    merge_job* j = (merge_job*)arg;
    j->result_len = internal_merge_with_buffer(j->lists, j->lens, j->k, j->buf);
    j->ran = 1;
    return NULL;
}

/* MergeSorted -> mergeSortedWithBuffer, algo/uidlist.go:448-542 */
size_t orc_merge_sorted(const uint64_t* const* lists, const size_t* lens, size_t k, uint64_t* out) {
    if (k < ORC_NUM_THREADS * ORC_NUM_LIST_PER_THREAD)
        return orc_internal_merge_sort(lists, lens, k, out);

    size_t chunk_sizes[ORC_NUM_THREADS];
    size_t total_needed = 0;
    size_t chunk = (k + ORC_NUM_THREADS - 1) / ORC_NUM_THREADS;
    for (int i = 0; i < ORC_NUM_THREADS; i++) {
        size_t start = (size_t)i * chunk, end = (size_t)(i + 1) * chunk;
        chunk_sizes[i] = 0;
        if (end > k) end = k;
        if (start > k) continue;
        size_t cap = 0;
        for (size_t j = start; j < end; j++) if (lists[j] != NULL) cap += lens[j];
        chunk_sizes[i] = cap;
        total_needed += cap;
    }
    size_t offsets[ORC_NUM_THREADS];
    offsets[0] = 0;
    for (int i = 1; i < ORC_NUM_THREADS; i++) offsets[i] = offsets[i - 1] + chunk_sizes[i - 1];

    /* bigBuffer := make([]uint64, totalCap) (:457) */
    uint64_t* big = (uint64_t*)malloc((total_needed ? total_needed : 1) * sizeof(uint64_t));
    merge_job jobs[ORC_NUM_THREADS];
    pthread_t th[ORC_NUM_THREADS];
    int started[ORC_NUM_THREADS];
    for (int i = 0; i < ORC_NUM_THREADS; i++) {
        size_t start = (size_t)i * chunk, end = (size_t)(i + 1) * chunk;
        started[i] = 0;
        jobs[i].ran = 0; jobs[i].result_len = 0;
        if (end > k) end = k;
        if (start > k) continue;
        jobs[i].lists = lists + start; jobs[i].lens = lens + start; jobs[i].k = end - start;
        jobs[i].buf = big + offsets[i];
        if (pthread_create(&th[i], NULL, merge_worker, &jobs[i]) == 0) started[i] = 1;
        else merge_worker(&jobs[i]);
    }
    for (int i = 0; i < ORC_NUM_THREADS; i++) if (started[i]) pthread_join(th[i], NULL);

    /* validResults: non-nil and non-empty (:529-535) */
    const uint64_t* vlists[ORC_NUM_THREADS]; size_t vlens[ORC_NUM_THREADS]; size_t nv = 0;
    for (int i = 0; i < ORC_NUM_THREADS; i++) {
        if (jobs[i].ran && jobs[i].result_len > 0) { vlists[nv] = jobs[i].buf; vlens[nv] = jobs[i].result_len; nv++; }
    }
    size_t olen = internal_merge_with_buffer(vlists, vlens, nv, out);
    free(big);
    return olen;
}

/* algo/uidlist.go:546-552 */
long long orc_index_of(const uint64_t* u, size_t n, uint64_t uid) {
    size_t i = search_ge(u, n, uid);
    if (i < n && u[i] == uid) return (long long)i;
    return -1;
}

/* ======================================================================== */
/* group varint (github.com/dgryski/go-groupvarint, not vendored)            */
/* ======================================================================== */
/* One group = 1 tag byte + 4 values.  Value j (0..3) is stored little-endian
 * in len_j = 1..4 bytes, minimal length; (len_j - 1) sits in tag bits
 * [2j+1:2j] (value 0 in the low bits).  BytesUsed[tag] = 1 + sum(len_j).     */

static unsigned gv_len(uint32_t v) {
    if (v < (1u << 8)) return 1;
    if (v < (1u << 16)) return 2;
    if (v < (1u << 24)) return 3;
    return 4;
}

size_t orc_gv_encode4(uint8_t* dst, const uint32_t src[4]) {
    size_t n = 1;
    uint8_t tag = 0;
    for (int j = 0; j < 4; j++) {
        unsigned l = gv_len(src[j]);
        tag |= (uint8_t)((l - 1) << (2 * j));
        for (unsigned b = 0; b < l; b++) dst[n++] = (uint8_t)(src[j] >> (8 * b));
    }
    dst[0] = tag;
    return n;
}

void orc_gv_decode4(uint32_t dst[4], const uint8_t* src) {
    uint8_t tag = src[0];
    size_t n = 1;
    for (int j = 0; j < 4; j++) {
        unsigned l = ((tag >> (2 * j)) & 3u) + 1;
        uint32_t v = 0;
        for (unsigned b = 0; b < l; b++) v |= (uint32_t)src[n++] << (8 * b);
        dst[j] = v;
    }
}

size_t orc_gv_bytes_used(uint8_t tag) {
    return 1 + 4 + (tag & 3u) + ((tag >> 2) & 3u) + ((tag >> 4) & 3u) + ((tag >> 6) & 3u);
}

/* ======================================================================== */
/* codec/codec.go                                                            */
/* ======================================================================== */

typedef struct {
    int block_size;
    orc_pack* pack;      /* e.pack; NULL until first Add */
    uint64_t* uids;      /* e.uids */
    size_t nuids, cap_uids;
    size_t cap_blocks, cap_deltas;
} encoder;

static void pack_reserve_block(encoder* e) {
    orc_pack* p = e->pack;
    if (p->nblocks + 1 > e->cap_blocks) {
        size_t nc = e->cap_blocks ? e->cap_blocks * 2 : 64;
        p->base = (uint64_t*)realloc(p->base, nc * sizeof(uint64_t));
        p->num_uids = (uint32_t*)realloc(p->num_uids, nc * sizeof(uint32_t));
        p->delta_off = (uint64_t*)realloc(p->delta_off, (nc + 1) * sizeof(uint64_t));
        e->cap_blocks = nc;
    }
}
static void pack_reserve_deltas(encoder* e, size_t extra) {
    orc_pack* p = e->pack;
    size_t have = p->delta_off[p->nblocks];
    if (have + extra + 32 > e->cap_deltas) {
        size_t nc = e->cap_deltas ? e->cap_deltas * 2 : 1024;
        while (nc < have + extra + 32) nc *= 2;
        p->deltas = (uint8_t*)realloc(p->deltas, nc);
        e->cap_deltas = nc;
    }
}

/* Encoder.packBlock, codec/codec.go:57-103 */
static void enc_pack_block(encoder* e) {
    if (e->nuids == 0) return;
    orc_pack* p = e->pack;
    pack_reserve_block(e);
    size_t b = p->nblocks;
    p->base[b] = e->uids[0];
    p->num_uids[b] = (uint32_t)e->nuids;

    uint64_t last = e->uids[0];
    const uint64_t* rest = e->uids + 1;
    size_t nrest = e->nuids - 1;
    size_t off = p->delta_off[b];
    for (;;) {
        uint32_t tmp[4];
        for (size_t i = 0; i < 4; i++) {
            if (i >= nrest) {
                tmp[i] = 0; /* padding: Encode4 works on batches of 4 */
            } else {
                tmp[i] = (uint32_t)(rest[i] - last);
                last = rest[i];
            }
        }
        pack_reserve_deltas(e, (off - p->delta_off[b]) + 17);
        off += orc_gv_encode4(p->deltas + off, tmp);
        if (nrest <= 4) { nrest = 0; break; }
        rest += 4; nrest -= 4;
    }
    p->delta_off[b + 1] = off;
    p->nblocks = b + 1;
}

static int match32msb(uint64_t a, uint64_t b) { /* codec/codec.go:469-471 */
    return (a & 0xffffffff00000000ull) == (b & 0xffffffff00000000ull);
}

/* Encoder.Add, codec/codec.go:108-127 */
static void enc_add(encoder* e, uint64_t uid) {
    if (e->pack == NULL) {
        e->pack = (orc_pack*)calloc(1, sizeof(orc_pack));
        e->pack->block_size = (uint32_t)e->block_size;
        e->pack->delta_off = (uint64_t*)calloc(65, sizeof(uint64_t));
        e->pack->base = (uint64_t*)malloc(64 * sizeof(uint64_t));
        e->pack->num_uids = (uint32_t*)malloc(64 * sizeof(uint32_t));
        e->cap_blocks = 64;
    }
    size_t size = e->nuids;
    if (size > 0 && !match32msb(e->uids[size - 1], uid)) {
        enc_pack_block(e);
        e->nuids = 0;
    }
    if (e->nuids + 1 > e->cap_uids) {
        size_t nc = e->cap_uids ? e->cap_uids * 2 : 256;
        e->uids = (uint64_t*)realloc(e->uids, nc * sizeof(uint64_t));
        e->cap_uids = nc;
    }
    e->uids[e->nuids++] = uid;
    if ((long long)e->nuids >= (long long)e->block_size) {
        enc_pack_block(e);
        e->nuids = 0;
    }
}

/* Encode + Encoder.Done, codec/codec.go:129-136, 393-399 */
orc_pack* orc_encode(const uint64_t* uids, size_t n, int block_size) {
    encoder e;
    memset(&e, 0, sizeof(e));
    e.block_size = block_size;
    for (size_t i = 0; i < n; i++) enc_add(&e, uids[i]);
    enc_pack_block(&e); /* Done() */
    free(e.uids);
    if (e.pack) {
        /* zero slack so group reads may touch up to 16 bytes past a group */
        size_t end = e.pack->delta_off[e.pack->nblocks];
        e.pack->deltas = (uint8_t*)realloc(e.pack->deltas, end + 32);
        memset(e.pack->deltas + end, 0, 32);
    }
    return e.pack;
}

void orc_pack_free(orc_pack* p) {
    if (!p) return;
    free(p->base); free(p->num_uids); free(p->delta_off); free(p->deltas); free(p);
}

size_t orc_approx_len(const orc_pack* p) { /* :418-423 */
    if (!p) return 0;
    return p->nblocks * (size_t)p->block_size;
}

size_t orc_exact_len(const orc_pack* p) { /* :427-440 */
    if (!p) return 0;
    size_t num = 0;
    for (size_t b = 0; b < p->nblocks; b++) num += p->num_uids[b];
    return num;
}

struct orc_decoder {
    const orc_pack* pack;
    size_t block_idx;
    uint64_t* buf;    /* storage behind d.uids */
    size_t cap;
    size_t start, len; /* d.uids = buf[start : start+len] */
};

orc_decoder* orc_decoder_new(const orc_pack* p) {
    orc_decoder* d = (orc_decoder*)calloc(1, sizeof(orc_decoder));
    d->pack = p;
    return d;
}
void orc_decoder_free(orc_decoder* d) { if (d) { free(d->buf); free(d); } }

static size_t pack_nblocks(const orc_pack* p) { return p ? p->nblocks : 0; }

/* Decoder.UnpackBlock, codec/codec.go:154-200 */
const uint64_t* orc_decoder_unpack_block(orc_decoder* d, size_t* len) {
    d->start = 0;
    d->len = 0;
    if (d->block_idx >= pack_nblocks(d->pack)) { *len = 0; return d->buf; }
    const orc_pack* p = d->pack;
    size_t b = d->block_idx;
    size_t num = p->num_uids[b];
    if (num + 4 > d->cap) {
        d->cap = num + 4 < 260 ? 260 : num + 4;
        d->buf = (uint64_t*)realloc(d->buf, d->cap * sizeof(uint64_t));
    }
    uint64_t last = p->base[b];
    d->buf[d->len++] = last;
    const uint8_t* enc = p->deltas + p->delta_off[b];
    size_t enc_len = (size_t)(p->delta_off[b + 1] - p->delta_off[b]);
    uint8_t tmp17[17];
    while (d->len < num) {
        const uint8_t* src = enc;
        if (enc_len < 17) {
            /* the reference pads the tail to 17 bytes (:175-187) */
            memset(tmp17, 0, sizeof(tmp17));
            memcpy(tmp17, enc, enc_len);
            src = tmp17;
        }
        uint32_t t[4];
        orc_gv_decode4(t, src);
        size_t used = orc_gv_bytes_used(src[0]);
        enc += used;
        enc_len = enc_len >= used ? enc_len - used : 0;
        for (int i = 0; i < 4; i++) {
            uint64_t sum = last + (uint64_t)t[i];
            d->buf[d->len++] = sum;
            last = sum;
        }
    }
    d->len = num;
    *len = d->len;
    return d->buf;
}

const uint64_t* orc_decoder_next(orc_decoder* d, size_t* len) { /* :377-380 */
    d->block_idx++;
    return orc_decoder_unpack_block(d, len);
}

const uint64_t* orc_decoder_uids(orc_decoder* d, size_t* len) { /* :340-342 */
    *len = d->len;
    return d->buf ? d->buf + d->start : NULL;
}

uint64_t orc_decoder_peek_next_base(const orc_decoder* d) { /* :362-368 */
    size_t bidx = d->block_idx + 1;
    if (bidx < pack_nblocks(d->pack)) return d->pack->base[bidx];
    return UINT64_MAX;
}

int orc_decoder_valid(const orc_decoder* d) { return d->block_idx < pack_nblocks(d->pack); } /* :371-374 */
size_t orc_decoder_block_idx(const orc_decoder* d) { return d->block_idx; }                  /* :382-384 */
void orc_decoder_set_block_idx(orc_decoder* d, size_t idx) { d->block_idx = idx; }

size_t orc_decoder_approx_len(const orc_decoder* d) { /* :203-211 */
    if (!d || !d->pack) return 0;
    return (size_t)d->pack->block_size * (d->pack->nblocks - d->block_idx);
}

/* sort.Search over block bases */
static size_t search_base(const orc_pack* p, size_t from, uint64_t uid, int whence) {
    size_t n = p->nblocks - from, lo = 0, hi = n;
    while (lo < hi) {
        size_t h = lo + (hi - lo) / 2;
        uint64_t base = p->base[h + from];
        int f = (whence == ORC_SEEK_START) ? (base >= uid) : (base > uid);
        if (!f) lo = h + 1; else hi = h;
    }
    return lo;
}

/* Decoder.Seek, codec/codec.go:279-337 */
const uint64_t* orc_decoder_seek(orc_decoder* d, uint64_t uid, int whence, size_t* len) {
    if (d->pack == NULL) { *len = 0; d->len = 0; return NULL; }
    d->block_idx = 0;
    if (uid == 0) return orc_decoder_unpack_block(d, len);
    const orc_pack* p = d->pack;
    size_t idx = search_base(p, 0, uid, whence);
    if (idx == 0) return orc_decoder_unpack_block(d, len);
    if (idx < p->nblocks && p->base[idx] == uid) {
        d->block_idx = idx;
        return orc_decoder_unpack_block(d, len);
    }
    d->block_idx = idx - 1;
    orc_decoder_unpack_block(d, len);
    /* uidx = first position with uids[i] >= uid (SeekStart) / > uid (SeekCurrent) */
    size_t uidx = (whence == ORC_SEEK_START) ? search_ge(d->buf, d->len, uid) : search_gt(d->buf, d->len, uid);
    if (uidx < d->len) {
        d->start = uidx;
        d->len -= uidx;
        *len = d->len;
        return d->buf + d->start;
    }
    return orc_decoder_next(d, len);
}

/* Decoder.SeekToBlock, codec/codec.go:219-271 */
const uint64_t* orc_decoder_seek_to_block(orc_decoder* d, uint64_t uid, int whence, size_t* len) {
    if (d->pack == NULL) { *len = 0; d->len = 0; return NULL; }
    const orc_pack* p = d->pack;
    size_t prev = d->block_idx;
    d->block_idx = 0;
    if (uid == 0) return orc_decoder_unpack_block(d, len);
    if (prev > 0 && prev < p->nblocks && uid < p->base[prev]) prev = 0;
    if (prev > p->nblocks) prev = p->nblocks; /* Go would panic slicing; clamp */
    size_t idx = search_base(p, prev, uid, whence) + prev;
    if (idx == 0) return orc_decoder_unpack_block(d, len);
    if (idx < p->nblocks && p->base[idx] == uid) {
        d->block_idx = idx;
        return orc_decoder_unpack_block(d, len);
    }
    d->block_idx = idx - 1;
    if (d->block_idx != prev || d->len == 0) { /* `|| len==0`: Go would index-panic on an empty d.uids */
        size_t tmp;
        orc_decoder_unpack_block(d, &tmp);
    }
    if (d->len > 0 && uid <= d->buf[d->start + d->len - 1]) {
        *len = d->len;
        return d->buf + d->start;
    }
    return orc_decoder_next(d, len);
}

/* Decoder.LinearSeek, codec/codec.go:349-359 */
const uint64_t* orc_decoder_linear_seek(orc_decoder* d, uint64_t seek, size_t* len) {
    for (;;) {
        uint64_t v = orc_decoder_peek_next_base(d);
        if (seek < v) break;
        d->block_idx++;
    }
    return orc_decoder_unpack_block(d, len);
}

/* Decode, codec/codec.go:444-452 */
size_t orc_decode(const orc_pack* p, uint64_t seek, uint64_t* out) {
    orc_decoder* dec = orc_decoder_new(p);
    size_t olen = 0, len = 0;
    const uint64_t* uids = orc_decoder_seek(dec, seek, ORC_SEEK_START, &len);
    while (len > 0) {
        memcpy(out + olen, uids, len * sizeof(uint64_t));
        olen += len;
        uids = orc_decoder_next(dec, &len);
    }
    orc_decoder_free(dec);
    return olen;
}

/* ---- IntersectCompressedWith*, algo/uidlist.go:33-138 -------------------- */

/* algo/uidlist.go:64-84 */
size_t orc_intersect_compressed_with_lin_jump(orc_decoder* dec, const uint64_t* v, size_t m, uint64_t* out) {
    size_t olen = 0, k = 0, off = 0, ulen = 0;
    const uint64_t* u = orc_decoder_uids(dec, &ulen);
    orc_intersect_with_lin(u, ulen, v + k, m - k, out, &olen, NULL, &off);
    k += off;
    while (k < m) {
        u = orc_decoder_linear_seek(dec, v[k], &ulen);
        if (ulen == 0) break;
        orc_intersect_with_lin(u, ulen, v + k, m - k, out, &olen, NULL, &off);
        if (off == 0) off = 1; /* v[k] isn't in u: move forward */
        k += off;
    }
    return olen;
}

/* algo/uidlist.go:90-138 */
size_t orc_intersect_compressed_with_bin(orc_decoder* dec, const uint64_t* q, size_t lq, uint64_t* out) {
    size_t ld = orc_exact_len(dec->pack);
    size_t olen = 0;
    if (lq == 0) return 0;
    if (ld <= lq) {
        size_t qlen = lq;
        for (;;) {
            size_t blen = 0, off = 0;
            const uint64_t* block = orc_decoder_uids(dec, &blen);
            if (blen == 0) break;
            orc_intersect_with_jump(block, blen, q, qlen, out, &olen, NULL, &off);
            q += off; qlen -= off;
            if (qlen == 0) return olen;
            size_t tmp;
            orc_decoder_next(dec, &tmp);
        }
        return olen;
    }
    size_t ulen = 0;
    const uint64_t* uids = orc_decoder_uids(dec, &ulen);
    size_t qidx = 0;
    for (;;) {
        if (qidx >= lq) return olen;
        uint64_t u = q[qidx];
        if (ulen == 0 || u > uids[ulen - 1]) {
            if (lq * ORC_LIN_VS_BIN < ld) uids = orc_decoder_linear_seek(dec, u, &ulen);
            else uids = orc_decoder_seek_to_block(dec, u, ORC_SEEK_CURRENT, &ulen);
            if (ulen == 0) return olen;
        }
        size_t off = 0;
        orc_intersect_with_jump(uids, ulen, q + qidx, lq - qidx, out, &olen, NULL, &off);
        if (off == 0) off = 1; /* v[k] isn't in u: move forward */
        qidx += off;
    }
}

/* algo/uidlist.go:33-61 */
size_t orc_intersect_compressed_with(const orc_pack* p, uint64_t after_uid,
                                     const uint64_t* v, size_t m_in, uint64_t* out) {
    if (p == NULL) return 0;
    orc_decoder* dec = orc_decoder_new(p);
    size_t tmp;
    orc_decoder_seek(dec, after_uid, ORC_SEEK_START, &tmp);
    size_t n = orc_decoder_approx_len(dec);
    size_t m = m_in;
    if (n > m) { size_t t = n; n = m; m = t; }
    if (n == 0) n = 1;
    double ratio = (double)m / (double)n;
    size_t olen;
    if (ratio < ORC_LIN_VS_BIN) olen = orc_intersect_compressed_with_lin_jump(dec, v, m_in, out);
    else olen = orc_intersect_compressed_with_bin(dec, v, m_in, out);
    orc_decoder_free(dec);
    return olen;
}
