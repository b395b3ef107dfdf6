// wire.hpp -- protobuf wire-format adjacency of the hot path (host code, no CUDA).
//
// posting lists reach the path as serialized pb.PostingList values read from
// Badger (proto.Unmarshal at posting/list.go:1045, posting/mvcc.go:634) and leave it as pb.List inside
// pb.Result.uid_matrix (protos/pb.proto:22-24, 76-78).  These functions read / write exactly those bytes so
// that a caller can go from a stored value to the struct-of-arrays dgx_pack_view (and from a result buffer
// to a pb.List message) without materialising Go structs in between.
//
// Standard proto3 wire format, restated from the message definitions (protos/pb.proto:378-408):
//   UidPack     { uint32 block_size = 1; repeated UidBlock blocks = 2; uint64 alloc_ref = 23; }
//   UidBlock    { uint64 base = 1; bytes deltas = 2; uint32 num_uids = 3; }
//   PostingList { UidPack pack = 1; repeated Posting postings = 2; uint64 commit_ts = 3; repeated uint64 splits = 4; }
//   List        { repeated fixed64 uids = 1; }            (packed: tag 0x0A, byte length, little-endian values)
// Unknown fields are skipped by wire type (well-formed groups included, as proto.Unmarshal does); for a
// scalar that appears twice the last value wins; truncated or malformed input is DGX_ERR_ARG.
#pragma once

#include <cstdint>
#include <cstring>

namespace dgx {
namespace wire {

struct Reader {
    const uint8_t* p;
    const uint8_t* end;
    bool varint(uint64_t* v) {
        uint64_t r = 0;
        for (int shift = 0; shift < 64; shift += 7) {
            if (p >= end) return false;
            const uint8_t b = *p++;
            if (shift == 63 && b > 1) return false;  // 10th byte: only bit 63 may be set (protowire: overflow)
            r |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) { *v = r; return true; }
        }
        return false;  // more than 10 bytes
    }
    bool bytes(const uint8_t** s, size_t* n) {
        uint64_t len;
        if (!varint(&len) || len > (uint64_t)(end - p)) return false;
        *s = p;
        *n = (size_t)len;
        p += len;
        return true;
    }
    // Skip one field value of wire type wt (field = its number, needed to match the end of a group).
    bool skip(uint32_t wt, uint32_t field = 0, int depth = 0) {
        uint64_t v;
        const uint8_t* s;
        size_t n;
        switch (wt) {
            case 0: return varint(&v);
            case 1: if (end - p < 8) return false; p += 8; return true;
            case 2: return bytes(&s, &n);
            case 5: if (end - p < 4) return false; p += 4; return true;
            case 3: {  // start of a (deprecated) group: an unknown field like any other, skipped to its end tag
                if (depth > 32) return false;
                for (;;) {
                    uint32_t f2, wt2;
                    bool ok;
                    if (!tag(&f2, &wt2, &ok)) return false;  // input ended inside the group
                    if (wt2 == 4) return f2 == field;
                    if (!skip(wt2, f2, depth + 1)) return false;
                }
            }
            default: return false;  // a stray end-group tag, or wire types 6 / 7
        }
    }
    // next field: false at the end of input or on a malformed tag (ok tells which)
    bool tag(uint32_t* field, uint32_t* wt, bool* ok) {
        *ok = true;
        if (p >= end) return false;
        uint64_t t;
        if (!varint(&t) || (t >> 3) == 0 || (t >> 3) > 0x1fffffffull) { *ok = false; return false; }
        *field = (uint32_t)(t >> 3);
        *wt = (uint32_t)(t & 7);
        return true;
    }
};

struct BlockFields {
    uint64_t base = 0;
    uint32_t num_uids = 0;
    const uint8_t* deltas = nullptr;
    size_t ndeltas = 0;
};

inline bool parse_block(const uint8_t* s, size_t n, BlockFields* b) {
    Reader r{s, s + n};
    uint32_t f, wt;
    bool ok;
    while (r.tag(&f, &wt, &ok)) {
        uint64_t v;
        if (f == 1 && wt == 0) { if (!r.varint(&v)) return false; b->base = v; }
        else if (f == 2 && wt == 2) { if (!r.bytes(&b->deltas, &b->ndeltas)) return false; }
        else if (f == 3 && wt == 0) { if (!r.varint(&v)) return false; b->num_uids = (uint32_t)v; }
        else if (!r.skip(wt, f)) return false;
    }
    return ok;
}

// One pass over a UidPack message.  With arrays == nullptr it only counts (measure); otherwise it fills them
// (caps already checked by the caller through a measure pass).
struct PackArrays {
    uint64_t* base;
    uint32_t* num_uids;
    uint64_t* delta_off;
    uint8_t* deltas;
};
inline bool walk_pack(const uint8_t* buf, size_t len, uint32_t* block_size, size_t* nblocks, size_t* delta_bytes,
                      const PackArrays* a) {
    Reader r{buf, buf + len};
    uint32_t f, wt;
    bool ok;
    size_t nb = 0, db = 0;
    *block_size = 0;
    if (a) a->delta_off[0] = 0;
    while (r.tag(&f, &wt, &ok)) {
        uint64_t v;
        if (f == 1 && wt == 0) {
            if (!r.varint(&v)) return false;
            *block_size = (uint32_t)v;
        } else if (f == 2 && wt == 2) {
            const uint8_t* s;
            size_t n;
            if (!r.bytes(&s, &n)) return false;
            BlockFields b;
            if (!parse_block(s, n, &b)) return false;
            if (a) {
                a->base[nb] = b.base;
                a->num_uids[nb] = b.num_uids;
                if (b.ndeltas) memcpy(a->deltas + db, b.deltas, b.ndeltas);
                a->delta_off[nb + 1] = (uint64_t)(db + b.ndeltas);
            }
            nb += 1;
            db += b.ndeltas;
        } else if (!r.skip(wt, f)) {
            return false;
        }
    }
    *nblocks = nb;
    *delta_bytes = db;
    return ok;
}

}  // namespace wire
}  // namespace dgx
