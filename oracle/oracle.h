/*
 * oracle.h -- CPU restatement of dgraph's posting-list set-op hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or
 * as the timed CPU baseline.  The shipped path is dgraph_b200/libdgx.so.
 *
 * Every function cites the reference lines (relative to /root/reference)
 * whose behaviour it restates.  The Go reference cannot be built in this
 * image (no go toolchain, no module cache), so this restatement is pinned
 * against the reference's own known-answer tests (tests/test_oracle_*.py
 * transcribe algo/uidlist_test.go, algo/packed_test.go, codec/codec_test.go).
 *
 * PARITY NOTE (group-varint bytes): the byte layout of UidBlock.Deltas comes
 * from the third-party module github.com/dgryski/go-groupvarint
 * v0.0.0-20230630160417-2bfb7969fb3c, which is NOT vendored under
 * /root/reference, and no reference test pins Deltas bytes (round trips
 * only).  The layout below restates the library's published format; at the
 * byte level it is "parity unpinned".
 */
#ifndef DGX_ORACLE_H
#define DGX_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- algo/uidlist.go --------------------------------------------------- */

/* IntersectWithLin, algo/uidlist.go:170-191.  Appends to out[*olen...].
 * Returns consumed positions through ri/rk (may be NULL). */
void orc_intersect_with_lin(const uint64_t* u, size_t n, const uint64_t* v, size_t m,
                            uint64_t* out, size_t* olen, size_t* ri, size_t* rk);
/* IntersectWithJump, algo/uidlist.go:195-220 (jump = 32, :17). */
void orc_intersect_with_jump(const uint64_t* u, size_t n, const uint64_t* v, size_t m,
                             uint64_t* out, size_t* olen, size_t* ri, size_t* rk);
/* IntersectWithBin + binIntersect, algo/uidlist.go:226-288.  Returns maxq. */
size_t orc_intersect_with_bin(const uint64_t* d, size_t ld, const uint64_t* q, size_t lq,
                              uint64_t* out, size_t* olen);
/* IntersectWith, algo/uidlist.go:142-167.  out may alias u (in place).
 * out must hold min(n,m) values.  Returns the result length. */
size_t orc_intersect_with(const uint64_t* u, size_t n, const uint64_t* v, size_t m, uint64_t* out);
/* Which branch IntersectWith takes: 0 = Lin, 1 = Jump, 2 = Bin (:156-165). */
int orc_intersect_with_branch(size_t n, size_t m);
/* IntersectSorted, algo/uidlist.go:297-329.  out must hold min_i lens[i]. */
size_t orc_intersect_sorted(const uint64_t* const* lists, const size_t* lens, size_t k, uint64_t* out);
/* Difference, algo/uidlist.go:332-362.  out must hold n values. */
size_t orc_difference(const uint64_t* u, size_t n, const uint64_t* v, size_t m, uint64_t* out);
/* MergeSorted, algo/uidlist.go:448-542 (+ heap.go:12-37).  out must hold
 * sum(lens).  Uses 10 threads when k >= 100 exactly like the reference. */
size_t orc_merge_sorted(const uint64_t* const* lists, const size_t* lens, size_t k, uint64_t* out);
/* internalMergeSort (single heap), algo/uidlist.go:392-446. */
size_t orc_internal_merge_sort(const uint64_t* const* lists, const size_t* lens, size_t k, uint64_t* out);
/* IndexOf, algo/uidlist.go:546-552.  Returns -1 when absent. */
long long orc_index_of(const uint64_t* u, size_t n, uint64_t uid);

/* ---- group varint (third party, see PARITY NOTE) ------------------------ */

/* Encode4: writes tag + 4 little-endian values of 1..4 bytes; returns bytes written (5..17). */
size_t orc_gv_encode4(uint8_t* dst, const uint32_t src[4]);
/* Decode4: reads one group at src into dst[4]; never reads beyond the group. */
void orc_gv_decode4(uint32_t dst[4], const uint8_t* src);
/* BytesUsed[tag]. */
size_t orc_gv_bytes_used(uint8_t tag);

/* ---- codec/codec.go ------------------------------------------------------ */

/* pb.UidPack restated as a struct-of-arrays (protos/pb.proto:379-400):
 * block b has Base = base[b], NumUids = num_uids[b] and
 * Deltas = deltas[delta_off[b] .. delta_off[b+1]).  `nil` packs are NULL. */
typedef struct orc_pack {
    uint32_t block_size;
    size_t nblocks;
    uint64_t* base;
    uint32_t* num_uids;
    uint64_t* delta_off; /* nblocks + 1 entries */
    uint8_t* deltas;     /* delta_off[nblocks] bytes + 32 bytes of zero slack */
} orc_pack;

/* Encode / Encoder.Add / packBlock / Done, codec/codec.go:57-136, 393-399.
 * Returns NULL for n == 0 (Encoder.Done with no Add returns a nil pack). */
orc_pack* orc_encode(const uint64_t* uids, size_t n, int block_size);
void orc_pack_free(orc_pack* p);
/* ApproxLen :418-423, ExactLen :427-440. */
size_t orc_approx_len(const orc_pack* p);
size_t orc_exact_len(const orc_pack* p);
/* Decode, codec/codec.go:444-452.  out must hold orc_exact_len(p). */
size_t orc_decode(const orc_pack* p, uint64_t seek, uint64_t* out);

/* Decoder, codec/codec.go:139-384. */
typedef struct orc_decoder orc_decoder;
enum { ORC_SEEK_START = 0, ORC_SEEK_CURRENT = 1 };
orc_decoder* orc_decoder_new(const orc_pack* p); /* Decoder{Pack: p}, no seek */
void orc_decoder_free(orc_decoder* d);
/* Each returns a pointer to the decoder-owned uid slice and its length. */
const uint64_t* orc_decoder_unpack_block(orc_decoder* d, size_t* len);
const uint64_t* orc_decoder_seek(orc_decoder* d, uint64_t uid, int whence, size_t* len);
const uint64_t* orc_decoder_seek_to_block(orc_decoder* d, uint64_t uid, int whence, size_t* len);
const uint64_t* orc_decoder_linear_seek(orc_decoder* d, uint64_t seek, size_t* len);
const uint64_t* orc_decoder_next(orc_decoder* d, size_t* len);
const uint64_t* orc_decoder_uids(orc_decoder* d, size_t* len);
uint64_t orc_decoder_peek_next_base(const orc_decoder* d);
int orc_decoder_valid(const orc_decoder* d);
size_t orc_decoder_block_idx(const orc_decoder* d);
void orc_decoder_set_block_idx(orc_decoder* d, size_t idx);
size_t orc_decoder_approx_len(const orc_decoder* d);

/* IntersectCompressedWith{,LinJump,Bin}, algo/uidlist.go:33-138.
 * out must hold m values.  Returns the result length. */
size_t orc_intersect_compressed_with(const orc_pack* p, uint64_t after_uid,
                                     const uint64_t* v, size_t m, uint64_t* out);
size_t orc_intersect_compressed_with_lin_jump(orc_decoder* dec, const uint64_t* v, size_t m, uint64_t* out);
size_t orc_intersect_compressed_with_bin(orc_decoder* dec, const uint64_t* q, size_t lq, uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif
