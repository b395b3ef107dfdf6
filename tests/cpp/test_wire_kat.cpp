// Known answers for the wire-format adjacency through the C++ mirror (include/dgx_algo.hpp, namespace
// dgx::wire).  Host code only: runs without a device.  Bytes are hand-derived from the proto3 wire format of
// the messages in /root/reference/protos/pb.proto:22-24, 378-408 (tests/test_wire.py checks the same entry
// points against the protobuf runtime).
#include <cstdio>
#include <vector>

#include "dgx_algo.hpp"

static int failures = 0;
#define REQUIRE(c, name) \
    do { if (!(c)) { std::printf("FAIL %s (%s:%d)\n", name, __FILE__, __LINE__); ++failures; } } while (0)

int main() {
    using namespace dgx;
    {   // UidPack{block_size 5, blocks [{base 1, deltas 00 02, num_uids 2}], alloc_ref 7}
        const uint8_t msg[] = {0x08, 0x05, 0x12, 0x08, 0x08, 0x01, 0x12, 0x02, 0x00, 0x02, 0x18, 0x02, 0xb8, 0x01, 0x07};
        pb::UidPack p = wire::ParseUidPack(msg, sizeof(msg));
        REQUIRE(p.block_size == 5 && p.base == std::vector<uint64_t>({1}) && p.num_uids == std::vector<uint32_t>({2}), "pack fields");
        REQUIRE(p.delta_off == std::vector<uint64_t>({0, 2}) && p.deltas == std::vector<uint8_t>({0, 2}), "pack deltas");
        REQUIRE(codec::ExactLen(&p) == 2 && codec::ApproxLen(&p) == 5, "pack lengths");
        // PostingList{pack = that message, commit_ts = 9}
        std::vector<uint8_t> pl = {0x0a, (uint8_t)sizeof(msg)};
        pl.insert(pl.end(), msg, msg + sizeof(msg));
        pl.push_back(0x18); pl.push_back(0x09);
        bool found = false;
        pb::UidPack q = wire::PostingListPack(pl.data(), pl.size(), &found);
        REQUIRE(found && q.base == p.base && q.deltas == p.deltas && q.block_size == 5, "posting list pack");
        const uint8_t no_pack[] = {0x18, 0x09};
        wire::PostingListPack(no_pack, sizeof(no_pack), &found);
        REQUIRE(!found, "posting list without pack");
    }
    {   // List{uids [1, 0x0102030405060708]}
        pb::List l({1, 0x0102030405060708ull});
        const std::vector<uint8_t> want = {0x0a, 0x10, 1, 0, 0, 0, 0, 0, 0, 0, 8, 7, 6, 5, 4, 3, 2, 1};
        REQUIRE(wire::ListToWire(l) == want, "list to wire");
        REQUIRE(wire::ListFromWire(want.data(), want.size()).Uids == l.Uids, "list from wire");
        REQUIRE(wire::ListToWire(pb::List()).empty(), "empty list is the empty message");
        std::vector<uint64_t> big(2048);
        for (size_t i = 0; i < big.size(); ++i) big[i] = i * 3;
        const std::vector<uint8_t> w = wire::ListToWire(pb::List(big));
        REQUIRE(w.size() == 4 + big.size() * 8 && w[0] == 0x0a && w[1] == 0x80 && w[2] == 0x80 && w[3] == 0x01, "header of 16384 bytes: 0a 80 80 01");
        REQUIRE(wire::ListFromWire(w.data(), w.size()).Uids == big, "round trip");
    }
    {   // malformed input throws dgx::Error with DGX_ERR_ARG
        const uint8_t bad[] = {0x12, 0x05, 0x08, 0x01};
        bool threw = false;
        try { wire::ParseUidPack(bad, sizeof(bad)); } catch (const Error& e) { threw = e.code == DGX_ERR_ARG; }
        REQUIRE(threw, "malformed pack");
    }
    if (failures == 0) std::printf("WIRE_KAT_OK\n");
    return failures == 0 ? 0 : 1;
}
