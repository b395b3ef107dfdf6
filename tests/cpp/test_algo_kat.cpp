// Known-answer tests of the C++ host mirror (include/dgx_algo.hpp), transcribing
// /root/reference/algo/uidlist_test.go :25-348 the way the Go tests read.
// Built and run by tests/test_cpp_mirror.py on a GPU box.
#include <cstdio>
#include <cstdlib>

#include "dgx_algo.hpp"

using dgx::pb::List;
using namespace dgx::algo;

static int failures = 0;
#define REQUIRE_EQ(got, want, what)                                                     \
    do {                                                                                \
        if ((got) != (want)) { ++failures; std::printf("FAIL %s (line %d)\n", what, __LINE__); } \
    } while (0)

static List newList(std::initializer_list<uint64_t> v) { return List(v); }
using V = std::vector<uint64_t>;

int main() {
    if (dgx_init(-1) != DGX_OK) { std::printf("no device: %s\n", dgx_last_error()); return 2; }
    {   // TestMergeSorted1..10
        List a = newList({55});
        REQUIRE_EQ(MergeSorted({&a}).Uids, V({55}), "MergeSorted1");
        List b = newList({1, 3, 6, 8, 10}), c = newList({2, 4, 5, 7, 15}), e = newList({});
        REQUIRE_EQ(MergeSorted({&b, &c}).Uids, V({1, 2, 3, 4, 5, 6, 7, 8, 10, 15}), "MergeSorted2");
        REQUIRE_EQ(MergeSorted({&b, &e}).Uids, V({1, 3, 6, 8, 10}), "MergeSorted3");
        REQUIRE_EQ(MergeSorted({&e, &b}).Uids, V({1, 3, 6, 8, 10}), "MergeSorted4");
        REQUIRE_EQ(MergeSorted({&e, &e}).Uids.empty(), true, "MergeSorted5");
        List d1 = newList({11, 13, 16, 18, 20}), d2 = newList({12, 14, 15, 15, 16, 16, 17, 25}), d3 = newList({1, 2});
        REQUIRE_EQ(MergeSorted({&d1, &d2, &d3}).Uids, V({1, 2, 11, 12, 13, 14, 15, 16, 17, 18, 20, 25}), "MergeSorted6");
        List f1 = newList({5, 6, 7}), f2 = newList({3, 4});
        REQUIRE_EQ(MergeSorted({&f1, &f2, &d3, &e}).Uids, V({1, 2, 3, 4, 5, 6, 7}), "MergeSorted7");
        REQUIRE_EQ(MergeSorted({}).Uids.empty(), true, "MergeSorted8");
        List g = newList({1, 1, 1});
        REQUIRE_EQ(MergeSorted({&g}).Uids, V({1}), "MergeSorted9");
        List h1 = newList({1, 2, 3, 3, 6}), h2 = newList({4, 8, 9});
        REQUIRE_EQ(MergeSorted({&h1, &h2}).Uids, V({1, 2, 3, 4, 6, 8, 9}), "MergeSorted10");
    }
    {   // TestIntersectSorted1..6
        List a = newList({1, 2, 3}), b = newList({2, 3, 4, 5}), c = newList({4, 5, 6});
        REQUIRE_EQ(IntersectSorted({&a, &b}).Uids, V({2, 3}), "IntersectSorted1");
        REQUIRE_EQ(IntersectSorted({&a}).Uids, V({1, 2, 3}), "IntersectSorted2");
        List none = IntersectSorted({});
        REQUIRE_EQ(none.nil && none.Uids.empty(), true, "IntersectSorted3");
        List d = newList({100, 101});
        REQUIRE_EQ(IntersectSorted({&d}).Uids, V({100, 101}), "IntersectSorted4");
        REQUIRE_EQ(IntersectSorted({&a, &b, &c}).Uids.empty(), true, "IntersectSorted5");
        List e1 = newList({10, 12, 13}), e2 = newList({2, 3, 4, 13});
        REQUIRE_EQ(IntersectSorted({&e1, &e2, &c}).Uids.empty(), true, "IntersectSorted6");
    }
    {   // TestDiffSorted1..5, TestSubSorted1/6
        List a = newList({1, 2, 3}), e = newList({});
        List v1 = newList({1}), v2 = newList({2}), v3 = newList({3}), v12 = newList({1, 2});
        REQUIRE_EQ(Difference(&a, &v1).Uids, V({2, 3}), "DiffSorted1");
        REQUIRE_EQ(Difference(&a, &v2).Uids, V({1, 3}), "DiffSorted2");
        REQUIRE_EQ(Difference(&a, &v3).Uids, V({1, 2}), "DiffSorted3");
        REQUIRE_EQ(Difference(&a, &e).Uids, V({1, 2, 3}), "DiffSorted4");
        List r = Difference(&e, &v12);
        REQUIRE_EQ(r.Uids.empty() && !r.nil, true, "DiffSorted5");
        List b = newList({2, 3, 4, 5});
        REQUIRE_EQ(Difference(&a, &b).Uids, V({1}), "SubSorted1");
        List c1 = newList({10, 12, 13}), c2 = newList({2, 3, 4, 13});
        REQUIRE_EQ(Difference(&c1, &c2).Uids, V({10, 12}), "SubSorted6");
        REQUIRE_EQ(Difference(nullptr, &a).Uids.empty(), true, "Difference(nil, v)");
    }
    {   // TestUIDListIntersect1..5 (in place: o == u) and the duplicate cases
        List u = newList({1, 2, 3}), v = newList({});
        IntersectWith(u, v, u);
        REQUIRE_EQ(u.Uids.empty(), true, "Intersect1");
        u = newList({1, 2, 3}); v = newList({1, 2, 3, 4, 5});
        IntersectWith(u, v, u);
        REQUIRE_EQ(u.Uids, V({1, 2, 3}), "Intersect2");
        REQUIRE_EQ(v.Uids, V({1, 2, 3, 4, 5}), "Intersect2 v untouched");
        u = newList({1, 2, 3}); v = newList({2});
        IntersectWith(u, v, u);
        REQUIRE_EQ(u.Uids, V({2}), "Intersect3");
        u = newList({1, 2, 3}); v = newList({0, 5});
        IntersectWith(u, v, u);
        REQUIRE_EQ(u.Uids.empty(), true, "Intersect4");
        u = newList({1, 2, 3}); v = newList({3, 5});
        IntersectWith(u, v, u);
        REQUIRE_EQ(u.Uids, V({3}), "Intersect5");
        u = newList({1, 1, 2, 3}); v = newList({1, 2});
        IntersectWith(u, v, u);
        REQUIRE_EQ(u.Uids, V({1, 2}), "IntersectDupFirst");
        u = newList({1, 1, 2, 3, 5}); v = newList({1, 1, 2, 4});
        IntersectWith(u, v, u);
        REQUIRE_EQ(u.Uids, V({1, 1, 2}), "IntersectDupBoth");
        u = newList({1, 2, 3, 5}); v = newList({1, 1, 2, 4});
        IntersectWith(u, v, u);
        REQUIRE_EQ(u.Uids, V({1, 2}), "IntersectDupSecond");
    }
    std::printf("%s: %d failure(s)\n", failures ? "FAILED" : "OK", failures);
    return failures ? 1 : 0;
}
