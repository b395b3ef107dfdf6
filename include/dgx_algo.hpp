// dgx_algo.hpp -- C++ host-side mirror of dgraph's `algo` and `codec` packages over
// the libdgx C ABI (include/dgx.h).
//
// The reference is compiled Go and its toolchain is absent from this image, so the
// host layer above the C ABI is mirrored in C++: same function names, argument
// meaning, aliasing rules and nil/empty result shapes as
//   /root/reference/algo/uidlist.go   (IntersectWith :142, IntersectSorted :297,
//                                      MergeSorted :448, Difference :332)
//   /root/reference/codec/codec.go    (Decode :444, ApproxLen :418, ExactLen :427)
// Every body is one libdgx call -- exactly what the cgo shim (go/algo_dgx.go) does.
// Errors: the Go functions cannot fail; here a non-zero dgx status throws
// dgx::Error so callers (the shim: fall back to the Go code) can react.
#pragma once

#include <algorithm>
#include <cstdint>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#include "dgx.h"

namespace dgx {

struct Error : std::runtime_error {
    int code;
    Error(int c, const char* msg) : std::runtime_error(std::string(msg)), code(c) {}
};
inline void check(int rc) {
    if (rc != DGX_OK) throw Error(rc, dgx_last_error());
}

namespace pb {
// pb.List (protos/pb.proto:22-24).  `nil` models a nil Uids slice.
struct List {
    std::vector<uint64_t> Uids;
    bool nil = false;
    List() = default;
    List(std::initializer_list<uint64_t> v) : Uids(v) {}
    explicit List(std::vector<uint64_t> v) : Uids(std::move(v)) {}
    static List Nil() { List l; l.nil = true; return l; }
};
// pb.UidPack flattened (protos/pb.proto:379-400): block b = {base[b], num_uids[b],
// deltas[delta_off[b] .. delta_off[b+1])}.
struct UidPack {
    uint32_t block_size = 0;
    std::vector<uint64_t> base;
    std::vector<uint32_t> num_uids;
    std::vector<uint64_t> delta_off{0};
    std::vector<uint8_t> deltas;
    dgx_pack_view view() const {
        dgx_pack_view v;
        v.block_size = block_size;
        v.nblocks = base.size();
        v.base = base.data();
        v.num_uids = num_uids.data();
        v.delta_off = delta_off.data();
        v.deltas = deltas.data();
        return v;
    }
};
}  // namespace pb

namespace algo {

// IntersectWith(u, v, o): o.Uids = u ∩ v.  `o` may be `u` (in place); `v` is never modified.
inline void IntersectWith(const pb::List& u, const pb::List& v, pb::List& o) {
    const size_t cap = std::min(u.Uids.size(), v.Uids.size());
    std::vector<uint64_t> out(std::max<size_t>(cap, 1));
    size_t n = 0;
    check(dgx_intersect2(u.Uids.data(), u.Uids.size(), v.Uids.data(), v.Uids.size(), out.data(), cap, &n));
    out.resize(n);
    o.Uids.swap(out);  // after the call: `u` may alias `o`
    o.nil = false;
}

// IntersectSorted(lists): no lists -> &pb.List{} (nil Uids); one list -> a copy.
inline pb::List IntersectSorted(const std::vector<const pb::List*>& lists) {
    if (lists.empty()) return pb::List::Nil();
    std::vector<const uint64_t*> ptrs;
    std::vector<size_t> lens;
    size_t cap = SIZE_MAX;
    for (const pb::List* l : lists) {
        ptrs.push_back(l->Uids.data());
        lens.push_back(l->Uids.size());
        cap = std::min(cap, l->Uids.size());
    }
    pb::List out;
    out.Uids.resize(std::max<size_t>(cap, 1));
    size_t n = 0;
    check(dgx_intersect_sorted(ptrs.data(), lens.data(), lists.size(), out.Uids.data(), cap, &n));
    out.Uids.resize(n);
    return out;
}

// MergeSorted(lists): sorted union, globally de-duplicated; nil / empty lists are skipped.
inline pb::List MergeSorted(const std::vector<const pb::List*>& lists) {
    std::vector<const uint64_t*> ptrs;
    std::vector<size_t> lens;
    size_t total = 0;
    for (const pb::List* l : lists) {
        const bool skip = (l == nullptr) || l->Uids.empty();
        ptrs.push_back(skip ? nullptr : l->Uids.data());
        lens.push_back(skip ? 0 : l->Uids.size());
        total += lens.back();
    }
    pb::List out;
    out.Uids.resize(std::max<size_t>(total, 1));
    size_t n = 0;
    check(dgx_merge_sorted(ptrs.data(), lens.data(), lists.size(), out.Uids.data(), total, &n));
    out.Uids.resize(n);
    return out;
}

// Difference(u, v): u \ v; nil u or v -> empty non-nil list (algo/uidlist.go:333-335).
inline pb::List Difference(const pb::List* u, const pb::List* v) {
    pb::List out;
    if (u == nullptr || v == nullptr) return out;
    out.Uids.resize(std::max<size_t>(u->Uids.size(), 1));
    size_t n = 0;
    check(dgx_difference(u->Uids.data(), u->Uids.size(), v->Uids.data(), v->Uids.size(), out.Uids.data(),
                         u->Uids.size(), &n));
    out.Uids.resize(n);
    return out;
}

}  // namespace algo

namespace codec {

inline size_t ApproxLen(const pb::UidPack* p) { return p ? p->base.size() * (size_t)p->block_size : 0; }
inline size_t ExactLen(const pb::UidPack* p) {
    return p ? std::accumulate(p->num_uids.begin(), p->num_uids.end(), (size_t)0) : 0;
}
// Decode(pack, seek): uids from Decoder.Seek(seek, SeekStart) onward; nil pack -> empty non-nil slice.
inline std::vector<uint64_t> Decode(const pb::UidPack* p, uint64_t seek) {
    std::vector<uint64_t> out;
    if (p == nullptr || p->base.empty()) return out;
    const size_t cap = ExactLen(p);
    out.resize(std::max<size_t>(cap, 1));
    size_t n = 0;
    const dgx_pack_view v = p->view();
    check(dgx_decode(&v, seek, out.data(), cap, &n));
    out.resize(n);
    return out;
}

}  // namespace codec

// Wire-format adjacency (dgx_wire_*, host only): the bytes proto.Unmarshal sees at posting/list.go:1045 and
// the pb.List messages of pb.Result.uid_matrix (protos/pb.proto:22-24, 76-78).
namespace wire {

// Serialized pb.UidPack -> the struct-of-arrays pb::UidPack.
inline pb::UidPack ParseUidPack(const uint8_t* buf, size_t len) {
    size_t nb = 0, db = 0;
    check(dgx_wire_pack_measure(buf, len, &nb, &db));
    pb::UidPack p;
    p.base.resize(nb);
    p.num_uids.resize(nb);
    p.delta_off.assign(nb + 1, 0);
    p.deltas.resize(db);
    dgx_pack_view v;
    check(dgx_wire_pack_parse(buf, len, p.base.data(), p.num_uids.data(), p.delta_off.data(), p.deltas.data(), nb, db, &v));
    p.block_size = v.block_size;
    return p;
}

// Serialized pb.PostingList -> its pb.UidPack; `found` is false when the posting list has no pack (nil).
inline pb::UidPack PostingListPack(const uint8_t* buf, size_t len, bool* found) {
    const uint8_t* sub = nullptr;
    size_t sub_len = 0;
    check(dgx_wire_posting_list_pack(buf, len, &sub, &sub_len));
    if (found) *found = sub != nullptr;
    return sub ? ParseUidPack(sub, sub_len) : pb::UidPack();
}

// pb.List -> its serialized message (header followed by the little-endian uids).
inline std::vector<uint8_t> ListToWire(const pb::List& l) {
    uint8_t hdr[16];
    const size_t h = dgx_wire_list_header(l.Uids.size(), hdr);
    std::vector<uint8_t> out(h + l.Uids.size() * 8);
    std::copy(hdr, hdr + h, out.begin());
    if (!l.Uids.empty()) std::copy((const uint8_t*)l.Uids.data(), (const uint8_t*)l.Uids.data() + l.Uids.size() * 8, out.begin() + h);
    return out;
}

inline pb::List ListFromWire(const uint8_t* buf, size_t len) {
    size_t n = 0;
    check(dgx_wire_list_decode(buf, len, nullptr, 0, &n));
    pb::List l;
    l.Uids.resize(n);
    if (n) check(dgx_wire_list_decode(buf, len, l.Uids.data(), n, &n));
    return l;
}

}  // namespace wire
}  // namespace dgx
