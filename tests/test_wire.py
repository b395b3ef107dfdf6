"""Wire-format adjacency: libdgx's dgx_wire_* entry points against the real protobuf
runtime.  The messages are declared here exactly as in the reference's protos/pb.proto (List :22-24,
UidBlock / UidPack :378-400, PostingList :402-408; Posting reduced to two of its fields) and serialized by
google.protobuf, which is what the reference's generated pb.pb.go produces / consumes.  Host code only: runs
without a GPU.
"""
import numpy as np
import pytest
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

from dgraph_b200 import _lib, pb, wire

F = descriptor_pb2.FieldDescriptorProto


def _messages():
    fdp = descriptor_pb2.FileDescriptorProto()
    fdp.name, fdp.package, fdp.syntax = "pb_hot_path.proto", "pbt", "proto3"

    def msg(name, fields):
        m = fdp.message_type.add()
        m.name = name
        for fname, num, typ, rep, tname in fields:
            f = m.field.add()
            f.name, f.number, f.type = fname, num, typ
            f.label = F.LABEL_REPEATED if rep else F.LABEL_OPTIONAL
            if tname:
                f.type_name = ".pbt." + tname

    msg("List", [("uids", 1, F.TYPE_FIXED64, True, None)])
    msg("UidBlock", [("base", 1, F.TYPE_UINT64, False, None), ("deltas", 2, F.TYPE_BYTES, False, None),
                     ("num_uids", 3, F.TYPE_UINT32, False, None)])
    msg("UidPack", [("block_size", 1, F.TYPE_UINT32, False, None), ("blocks", 2, F.TYPE_MESSAGE, True, "UidBlock"),
                    ("alloc_ref", 23, F.TYPE_UINT64, False, None)])
    msg("Posting", [("uid", 1, F.TYPE_FIXED64, False, None), ("value", 2, F.TYPE_BYTES, False, None)])
    msg("Result", [("uid_matrix", 1, F.TYPE_MESSAGE, True, "List"), ("counts", 3, F.TYPE_UINT32, True, None),
                    ("intersect_dest", 4, F.TYPE_BOOL, False, None)])
    msg("PostingList", [("pack", 1, F.TYPE_MESSAGE, False, "UidPack"), ("postings", 2, F.TYPE_MESSAGE, True, "Posting"),
                        ("commit_ts", 3, F.TYPE_UINT64, False, None), ("splits", 4, F.TYPE_UINT64, True, None)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fdp)
    return {n: message_factory.GetMessageClass(pool.FindMessageTypeByName("pbt." + n))
            for n in ("List", "UidBlock", "UidPack", "Posting", "PostingList", "Result")}


M = _messages()


def proto_pack(opack, alloc_ref=0):
    """oracle Pack -> protobuf UidPack message."""
    m = M["UidPack"]()
    m.block_size = opack.block_size
    for b in range(opack.nblocks):
        blk = m.blocks.add()
        blk.base = int(opack.base[b])
        blk.deltas = opack.block_deltas(b)
        blk.num_uids = int(opack.num_uids[b])
    m.alloc_ref = alloc_ref
    return m


def assert_same_pack(got: pb.UidPack, opack):
    assert got.block_size == opack.block_size
    assert np.array_equal(got.base, opack.base)
    assert np.array_equal(got.num_uids, opack.num_uids)
    assert np.array_equal(got.delta_off, opack.delta_off)
    assert np.array_equal(got.deltas, opack.deltas)


def test_known_answer_bytes():
    """Hand-derived: UidPack{block_size 5, blocks [{base 1, deltas 00 02, num_uids 2}]}."""
    want = bytes.fromhex("08 05 12 08 08 01 12 02 00 02 18 02")
    m = M["UidPack"]()
    m.block_size = 5
    b = m.blocks.add()
    b.base, b.deltas, b.num_uids = 1, b"\x00\x02", 2
    assert m.SerializeToString() == want
    p = wire.ParseUidPack(want)
    assert (p.block_size, p.nblocks) == (5, 1)
    assert p.base.tolist() == [1] and p.num_uids.tolist() == [2]
    assert p.delta_off.tolist() == [0, 2] and p.deltas.tolist() == [0, 2]


@pytest.mark.parametrize("bs", [1, 5, 256])
def test_uid_pack_roundtrip(orc, bs):
    rng = np.random.default_rng(bs)
    uids = np.unique(np.concatenate([rng.integers(1, 1 << 20, 3000, dtype=np.uint64),
                                     (np.uint64(7) << np.uint64(32)) + rng.integers(0, 5000, 500, dtype=np.uint64)]))
    opack = orc.encode(uids, bs)
    data = proto_pack(opack, alloc_ref=0).SerializeToString()
    got = wire.ParseUidPack(data)
    assert_same_pack(got, opack)
    # the parsed struct of arrays decodes (CPU oracle) to the original uids
    assert np.array_equal(orc.decode(opack, 0), uids)
    # alloc_ref (field 23, two-byte tag) and an unknown trailing field are skipped
    noisy = proto_pack(opack, alloc_ref=12345).SerializeToString() + bytes.fromhex("f8 07 2a") + bytes.fromhex("e5 07 01 02 03 04")
    assert_same_pack(wire.ParseUidPack(noisy), opack)


def test_empty_pack():
    p = wire.ParseUidPack(b"")
    assert p.nblocks == 0 and p.block_size == 0 and p.delta_off.tolist() == [0] and p.deltas.size == 0
    m = M["UidPack"]()
    m.block_size = 256
    p = wire.ParseUidPack(m.SerializeToString())
    assert p.nblocks == 0 and p.block_size == 256


def test_posting_list_pack(orc):
    uids = np.arange(10, 5000, 7, dtype=np.uint64)
    opack = orc.encode(uids, 256)
    pl = M["PostingList"]()
    pl.pack.CopyFrom(proto_pack(opack))
    for u in (11, 12):
        post = pl.postings.add()
        post.uid, post.value = u, b"facet-or-value"
    pl.commit_ts = 99
    pl.splits.extend([1, 1 << 40])
    data = pl.SerializeToString()
    sub = wire.PostingListPack(data)
    assert sub == pl.pack.SerializeToString()
    assert_same_pack(wire.ParseUidPack(sub), opack)
    # a value-only posting list has no pack: nil
    pl2 = M["PostingList"]()
    pl2.postings.add().uid = 5
    assert wire.PostingListPack(pl2.SerializeToString()) is None
    assert wire.PostingListPack(b"") is None


@pytest.mark.parametrize("n", [0, 1, 2, 15, 16, 17, 2047, 2048, 100000])
def test_list_to_wire_matches_protobuf(n):
    rng = np.random.default_rng(n)
    uids = np.sort(rng.integers(0, 2**64 - 1, n, dtype=np.uint64))
    m = M["List"]()
    m.uids.extend(int(x) for x in uids)
    want = m.SerializeToString()
    assert wire.ListToWire(pb.List(uids)) == want
    back = wire.ListFromWire(want)
    assert np.array_equal(back.Uids, uids)
    assert wire.ListToWire(None) == b""


def test_uid_matrix_matches_protobuf():
    """A CSR batch result framed as pb.Result.uid_matrix (protos/pb.proto:76-78), empty rows included."""
    rng = np.random.default_rng(3)
    rows = [np.sort(rng.integers(0, 2**63, int(n), dtype=np.uint64)) for n in (0, 1, 5, 0, 300, 16, 0)]
    out = np.concatenate(rows)
    off = np.concatenate([[0], np.cumsum([r.size for r in rows])]).astype(np.uint64)
    m = M["Result"]()
    for r in rows:
        m.uid_matrix.add().uids.extend(int(x) for x in r)
    assert wire.UidMatrixToWire(out, off) == m.SerializeToString()
    # the caller appends other fields of pb.Result behind the rows
    m.intersect_dest = True
    tail = bytes.fromhex("20 01")
    assert wire.UidMatrixToWire(out, off) + tail == m.SerializeToString()
    back = M["Result"]()
    back.ParseFromString(wire.UidMatrixToWire(out, off))
    assert [list(l.uids) for l in back.uid_matrix] == [r.tolist() for r in rows]
    assert wire.UidMatrixToWire(np.zeros(0, np.uint64), np.zeros(1, np.uint64)) == b""


def test_list_unpacked_and_split_encodings():
    """A parser must accept repeated fixed64 both packed and one-tag-per-value, in any mix."""
    vals = [1, 2**64 - 1, 77, 0x0102030405060708]
    le = [int(v).to_bytes(8, "little") for v in vals]
    unpacked = b"".join(b"\x09" + x for x in le)
    assert wire.ListFromWire(unpacked).tolist() == vals
    mixed = b"\x0a\x10" + le[0] + le[1] + b"\x09" + le[2] + b"\x0a\x08" + le[3]
    assert wire.ListFromWire(mixed).tolist() == vals
    parsed = M["List"]()
    parsed.ParseFromString(mixed)
    assert list(parsed.uids) == vals            # the protobuf runtime agrees


@pytest.mark.parametrize("bad", [
    bytes.fromhex("08"),                         # tag without value
    bytes.fromhex("12 05 08 01"),                # block length runs past the buffer
    bytes.fromhex("12 03 12 05 00"),             # deltas length runs past the block
    bytes.fromhex("0b 08 01"),                   # group that never ends
    bytes.fromhex("0b 08 01 14"),                # group closed by another field's end tag
    bytes.fromhex("0c"),                         # stray end-group tag
    bytes.fromhex("08 ff ff ff ff ff ff ff ff ff ff 01"),  # varint longer than 10 bytes
    bytes.fromhex("00 01"),                      # field number 0
])
def test_malformed_pack_is_rejected(bad):
    with pytest.raises(_lib.DgxError) as e:
        wire.ParseUidPack(bad)
    assert e.value.code == -3  # DGX_ERR_ARG
    # the protobuf runtime rejects the same bytes
    with pytest.raises(Exception):
        M["UidPack"]().ParseFromString(bad)


def test_unknown_groups_are_skipped():
    """A well-formed group is an unknown field like any other (proto.Unmarshal skips it)."""
    m = M["UidPack"]()
    m.block_size = 7
    b = m.blocks.add()
    b.base, b.deltas, b.num_uids = 3, b"\x00\x01\x02", 2
    raw = bytes.fromhex("0b 08 01 13 08 02 14 0c") + m.SerializeToString() + bytes.fromhex("eb 07 ec 07")
    chk = M["UidPack"]()
    chk.ParseFromString(raw)                     # the protobuf runtime accepts it ...
    assert chk.block_size == 7 and len(chk.blocks) == 1
    p = wire.ParseUidPack(raw)                   # ... and so does the parser, with the same content
    assert p.block_size == 7 and p.base.tolist() == [3] and p.deltas.tolist() == [0, 1, 2]


def test_malformed_list_and_caps():
    import ctypes as C
    with pytest.raises(_lib.DgxError):
        wire.ListFromWire(bytes.fromhex("0a 07 01 02 03 04 05 06 07"))      # packed length not a multiple of 8
    with pytest.raises(_lib.DgxError):
        wire.ListFromWire(bytes.fromhex("09 01 02 03"))                     # truncated fixed64
    lib = _lib.load()
    data = wire.ListToWire(pb.List([1, 2, 3]))
    out = np.zeros(2, np.uint64)
    n = C.c_size_t(0)
    assert lib.dgx_wire_list_decode(data, len(data), out.ctypes.data, 2, C.byref(n)) == -4  # DGX_ERR_CAP
    # pack arrays too small
    m = M["UidPack"]()
    b = m.blocks.add()
    b.base, b.deltas, b.num_uids = 9, b"\x00" * 5, 1
    raw = m.SerializeToString()
    view = _lib.PackView()
    off = np.zeros(2, np.uint64)
    assert lib.dgx_wire_pack_parse(raw, len(raw), None, None, off.ctypes.data, None, 0, 0, C.byref(view)) == -4


def test_differential_fuzz_against_protobuf_runtime():
    """Random and mutated byte strings: the parser accepts exactly what the protobuf runtime accepts and reads
    the same fields.  One documented divergence is excluded: a 10-byte varint whose last byte exceeds 1
    overflows 64 bits -- Go's protowire (the reference's runtime, google.golang.org/protobuf) and this parser
    reject it, the Python runtime silently drops the extra bits."""
    import random
    import re

    def ours(data):
        try:
            p = wire.ParseUidPack(data)
        except _lib.DgxError:
            return None
        return (p.block_size, [(int(p.base[i]), int(p.num_uids[i]),
                                p.deltas[int(p.delta_off[i]):int(p.delta_off[i + 1])].tobytes()) for i in range(p.nblocks)])

    def runtime(data):
        m = M["UidPack"]()
        try:
            m.ParseFromString(data)
        except Exception:  # noqa: BLE001  (DecodeError)
            return None
        return (m.block_size, [(b.base, b.num_uids, bytes(b.deltas)) for b in m.blocks])

    rng = random.Random(20260921)
    valid = []
    for _ in range(40):
        m = M["UidPack"]()
        m.block_size = rng.choice([0, 1, 5, 256, 2**32 - 1])
        for _ in range(rng.randrange(0, 5)):
            b = m.blocks.add()
            b.base = rng.choice([0, 1, 2**40, 2**64 - 1])
            b.num_uids = rng.randrange(0, 300)
            b.deltas = bytes(rng.randrange(256) for _ in range(rng.randrange(0, 12)))
        m.alloc_ref = rng.choice([0, 5, 2**63])
        valid.append(m.SerializeToString())
    long_varint = re.compile(rb"[\x80-\xff]{9}")
    checked = accepted = 0
    for it in range(20000):
        if it % 3 == 0:
            data = bytes(rng.randrange(256) for _ in range(rng.randrange(0, 24)))
        else:
            d = bytearray(rng.choice(valid))
            for _ in range(rng.randrange(1, 4)):
                op = rng.randrange(3)
                if op == 0 and d:
                    d[rng.randrange(len(d))] = rng.randrange(256)
                elif op == 1 and d:
                    del d[rng.randrange(len(d))]
                else:
                    d.insert(rng.randrange(len(d) + 1), rng.randrange(256))
            data = bytes(d)
        if long_varint.search(data):
            continue
        want = runtime(data)
        assert ours(data) == want, data.hex()
        checked += 1
        accepted += want is not None
    assert checked > 10000 and accepted > 1000


def test_varint_overflow_follows_go_protowire():
    """protowire.ConsumeVarint: the 10th byte may only carry bit 63."""
    ok = bytes.fromhex("b8 01 ff ff ff ff ff ff ff ff ff 01")      # alloc_ref = 2^64 - 1
    assert wire.ParseUidPack(ok).nblocks == 0
    with pytest.raises(_lib.DgxError):
        wire.ParseUidPack(bytes.fromhex("b8 01 ff ff ff ff ff ff ff ff ff 02"))
    with pytest.raises(_lib.DgxError):
        wire.ParseUidPack(bytes.fromhex("b8 01 80 80 80 80 80 80 80 80 80 80 01"))  # 11 bytes
