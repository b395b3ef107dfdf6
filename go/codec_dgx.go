//go:build dgx

// Package codec -- cgo shim for codec.Decode (codec/codec.go:444) over libdgx.
// NOT COMPILED IN THIS REPOSITORY (no Go toolchain in the build image); see
// go/algo_dgx.go and INTEGRATION.md.
package codec

/*
#cgo CFLAGS: -I${SRCDIR}/../include
#cgo LDFLAGS: -ldgx
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "dgx.h"
*/
import "C"

import (
	"unsafe"

	"github.com/dgraph-io/dgraph/v25/protos/pb"
)

const dgxMinDecode = 1 << 16

// flatten copies pb.UidPack (Blocks []*UidBlock: pointers to structs, each with its
// own Deltas slice) into the struct-of-arrays dgx_pack_view, in C memory: cgo cannot
// pass Go memory that contains Go pointers.
func flatten(pack *pb.UidPack) (C.dgx_pack_view, func()) {
	nb := len(pack.Blocks)
	total := 0
	for _, b := range pack.Blocks {
		total += len(b.Deltas)
	}
	base := (*[1 << 28]C.uint64_t)(C.malloc(C.size_t(nb*8 + 8)))
	num := (*[1 << 28]C.uint32_t)(C.malloc(C.size_t(nb*4 + 4)))
	off := (*[1 << 28]C.uint64_t)(C.malloc(C.size_t(nb*8 + 8)))
	del := C.malloc(C.size_t(total + 16))
	o := 0
	for i, b := range pack.Blocks {
		base[i] = C.uint64_t(b.Base)
		num[i] = C.uint32_t(b.NumUids)
		off[i] = C.uint64_t(o)
		if len(b.Deltas) > 0 {
			C.memcpy(unsafe.Add(del, o), unsafe.Pointer(&b.Deltas[0]), C.size_t(len(b.Deltas)))
		}
		o += len(b.Deltas)
	}
	off[nb] = C.uint64_t(o)
	var v C.dgx_pack_view
	v.block_size = C.uint32_t(pack.BlockSize)
	v.nblocks = C.size_t(nb)
	v.base = (*C.uint64_t)(unsafe.Pointer(base))
	v.num_uids = (*C.uint32_t)(unsafe.Pointer(num))
	v.delta_off = (*C.uint64_t)(unsafe.Pointer(off))
	v.deltas = (*C.uint8_t)(del)
	return v, func() { C.free(unsafe.Pointer(base)); C.free(unsafe.Pointer(num)); C.free(unsafe.Pointer(off)); C.free(del) }
}

// Decode keeps codec/codec.go:444's contract: uids from Seek(seek, SeekStart) onward,
// a non-nil (possibly empty) slice, sized with ExactLen (ApproxLen is 0 for
// BlockSize-0 packs and only an estimate otherwise).
func Decode(pack *pb.UidPack, seek uint64) []uint64 {
	n := ExactLen(pack)
	if pack == nil || n < dgxMinDecode {
		return decodeGo(pack, seek)
	}
	view, free := flatten(pack)
	defer free()
	out := make([]uint64, n)
	var outLen C.size_t
	if rc := C.dgx_decode(&view, C.uint64_t(seek), (*C.uint64_t)(unsafe.Pointer(&out[0])), C.size_t(n), &outLen); rc != C.DGX_OK {
		return decodeGo(pack, seek)
	}
	return out[:int(outLen)]
}

// DecodeStored decodes the uids of a posting list straight from its stored bytes (the
// value proto.Unmarshal receives at posting/list.go:1045): the pb.UidPack inside the
// serialized pb.PostingList is parsed into the struct-of-arrays view by libdgx itself
// (dgx_wire_*), so no []*pb.UidBlock is ever built.  New entry point (not in the
// reference); nil for a posting list without a pack.
func DecodeStored(value []byte, seek uint64) ([]uint64, bool) {
	if len(value) == 0 {
		return []uint64{}, true
	}
	buf := (*C.uint8_t)(unsafe.Pointer(&value[0]))
	var sub *C.uint8_t
	var subLen C.size_t
	if C.dgx_wire_posting_list_pack(buf, C.size_t(len(value)), &sub, &subLen) != C.DGX_OK {
		return nil, false
	}
	if sub == nil {
		return []uint64{}, true
	}
	var nb, db C.size_t
	if C.dgx_wire_pack_measure(sub, subLen, &nb, &db) != C.DGX_OK {
		return nil, false
	}
	base := C.malloc(nb*8 + 8)
	num := C.malloc(nb*4 + 4)
	off := C.malloc(nb*8 + 16)
	del := C.malloc(db + 16)
	defer func() { C.free(base); C.free(num); C.free(off); C.free(del) }()
	var view C.dgx_pack_view
	if C.dgx_wire_pack_parse(sub, subLen, (*C.uint64_t)(base), (*C.uint32_t)(num), (*C.uint64_t)(off),
		(*C.uint8_t)(del), nb, db, &view) != C.DGX_OK {
		return nil, false
	}
	n := 0
	nums := (*[1 << 28]C.uint32_t)(num)[:int(nb):int(nb)]
	for _, c := range nums {
		n += int(c) // codec.ExactLen
	}
	out := make([]uint64, n)
	var outLen C.size_t
	var outPtr *C.uint64_t
	if n > 0 {
		outPtr = (*C.uint64_t)(unsafe.Pointer(&out[0]))
	}
	if C.dgx_decode(&view, C.uint64_t(seek), outPtr, C.size_t(n), &outLen) != C.DGX_OK {
		return nil, false
	}
	return out[:int(outLen)], true
}

// packRef names a pack for the HBM-resident cache of libdgx: key = hash of the posting list's Badger key,
// version = commit timestamp of the immutable layer the pack was rolled up at (posting.List.minTs).
// key 0 = anonymous (copied for this call only).
type PackRef struct {
	Pack    *pb.UidPack
	Key     uint64
	Version uint64
}

// IntersectCompressedWithRef is algo.IntersectCompressedWith (algo/uidlist.go:33-61), the entry point of
// posting.List.Uids for filtered reads (posting/list.go:1795-1800): o.Uids = v ∩ pack[>= afterUID].  The pack
// crosses PCIe compressed (or not at all when (Key, Version) is already resident); only blocks whose range
// holds an element of v are decoded.  Returns false when the caller must run the Go path.
func IntersectCompressedWithRef(ref PackRef, afterUID uint64, v, o *pb.List) bool {
	if ref.Pack == nil {
		return true // `if pack == nil { return }`: o untouched
	}
	if ApproxLen(ref.Pack)+len(v.Uids) < dgxMinDecode {
		return false
	}
	view, free := flatten(ref.Pack)
	defer free()
	cref := C.dgx_pack_ref{pack: &view, key: C.uint64_t(ref.Key), version: C.uint64_t(ref.Version)}
	out := make([]uint64, len(v.Uids))
	var outLen C.size_t
	var vp, op *C.uint64_t
	if len(v.Uids) > 0 {
		vp = (*C.uint64_t)(unsafe.Pointer(&v.Uids[0]))
		op = (*C.uint64_t)(unsafe.Pointer(&out[0]))
	}
	if C.dgx_intersect_compressed_ref(&cref, C.uint64_t(afterUID), vp, C.size_t(len(v.Uids)), op, C.size_t(len(out)), &outLen) != C.DGX_OK {
		return false
	}
	o.Uids = out[:int(outLen)]
	return true
}

// IntersectSortedPacks is algo.IntersectSorted over lists still held as packs: every pack is decoded on the
// device (codec.Decode(p, 0)) and the k-way intersection runs there; ~1.5 B/UID cross PCIe instead of 8.
func IntersectSortedPacks(refs []PackRef) (*pb.List, bool) {
	if len(refs) == 0 {
		return &pb.List{}, true
	}
	crefs := (*[1 << 20]C.dgx_pack_ref)(C.malloc(C.size_t(len(refs)) * C.size_t(unsafe.Sizeof(C.dgx_pack_ref{}))))
	views := (*[1 << 20]C.dgx_pack_view)(C.malloc(C.size_t(len(refs)) * C.size_t(unsafe.Sizeof(C.dgx_pack_view{}))))
	frees := make([]func(), 0, len(refs))
	defer func() {
		for _, f := range frees {
			f()
		}
		C.free(unsafe.Pointer(crefs))
		C.free(unsafe.Pointer(views))
	}()
	minLen := int(^uint(0) >> 1)
	for i, r := range refs {
		crefs[i] = C.dgx_pack_ref{key: C.uint64_t(r.Key), version: C.uint64_t(r.Version)}
		if r.Pack == nil || len(r.Pack.Blocks) == 0 {
			return &pb.List{Uids: []uint64{}}, true // an empty operand empties the intersection
		}
		v, free := flatten(r.Pack)
		frees = append(frees, free)
		views[i] = v
		crefs[i].pack = &views[i]
		if n := ExactLen(r.Pack); n < minLen {
			minLen = n
		}
	}
	out := make([]uint64, minLen+1)
	var outLen C.size_t
	if C.dgx_intersect_sorted_packed(&crefs[0], C.size_t(len(refs)), (*C.uint64_t)(unsafe.Pointer(&out[0])), C.size_t(minLen), &outLen) != C.DGX_OK {
		return nil, false
	}
	return &pb.List{Uids: out[:int(outLen)]}, true
}

// EncodeDevice is codec.Encode (codec/codec.go:393-399) with the block split and the group-varint bytes produced
// on the device; the result is rebuilt as a pb.UidPack whose blocks share one Deltas backing array.
func EncodeDevice(uids []uint64, blockSize int) (*pb.UidPack, bool) {
	if len(uids) == 0 {
		return nil, true
	}
	var nbCap, dbCap C.size_t
	C.dgx_encode_bound(C.size_t(len(uids)), C.uint32_t(blockSize), &nbCap, &dbCap)
	for attempt := 0; attempt < 2; attempt++ {
		base := make([]uint64, int(nbCap)+1)
		num := make([]uint32, int(nbCap)+1)
		off := make([]uint64, int(nbCap)+2)
		del := make([]byte, int(dbCap)+1)
		nb, db := nbCap, dbCap
		rc := C.dgx_encode((*C.uint64_t)(unsafe.Pointer(&uids[0])), C.size_t(len(uids)), C.uint32_t(blockSize),
			(*C.uint64_t)(unsafe.Pointer(&base[0])), (*C.uint32_t)(unsafe.Pointer(&num[0])),
			(*C.uint64_t)(unsafe.Pointer(&off[0])), (*C.uint8_t)(unsafe.Pointer(&del[0])), &nb, &db, nil)
		if rc == C.DGX_ERR_CAP && (nb > nbCap || db > dbCap) {
			nbCap, dbCap = nb, db // the exact sizes came back: one more try
			continue
		}
		if rc != C.DGX_OK {
			return nil, false
		}
		pack := &pb.UidPack{BlockSize: uint32(blockSize), Blocks: make([]*pb.UidBlock, int(nb))}
		blocks := make([]pb.UidBlock, int(nb))
		for i := range blocks {
			blocks[i] = pb.UidBlock{Base: base[i], NumUids: num[i], Deltas: del[off[i]:off[i+1]:off[i+1]]}
			pack.Blocks[i] = &blocks[i]
		}
		return pack, true
	}
	return nil, false
}
