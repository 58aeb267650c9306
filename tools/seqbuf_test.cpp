// seqbuf_test: SeqBuffer / SlabPool of the host shim (mash_b200/host/seqbuf.hpp): contents survive growth and moves, small records
// stay on the heap, large ones come from 2 MiB-granular slabs that are handed out again after release, from several threads at once.
// Built and run by tests/test_host_shim.py; exit code 0 = all checks passed.
#include "seqbuf.hpp"
#include <cstdio>
#include <random>
#include <string>
#include <thread>
#include <vector>
using mashhost::SeqBuffer;
static int bad = 0;
#define CHECK(c) do { if (!(c)) { if (bad++ < 10) printf("check failed line %d: %s\n", __LINE__, #c); } } while (0)

static void one_thread(unsigned seed)
{
    std::mt19937_64 rng(seed);
    std::vector<std::pair<SeqBuffer, std::string>> kept;
    for (int it = 0; it < 200; it++) {
        SeqBuffer b;
        std::string want;
        const size_t target = (it % 3 == 0) ? rng() % 2000 : (it % 3 == 1 ? rng() % 3000000 : (1u << 20) - 3 + rng() % 7);
        if (it % 5 == 0) b.reserve(target / 2);
        while (want.size() < target) {
            std::string piece(1 + rng() % 70000, (char)('A' + rng() % 20));
            b.append(piece.data(), piece.size());
            want += piece;
        }
        CHECK(b.size() == want.size());
        CHECK(want.empty() || memcmp(b.data(), want.data(), want.size()) == 0);
        CHECK(b.capacity() >= b.size());
        SeqBuffer c(std::move(b));
        CHECK(b.size() == 0 && b.capacity() == 0 && b.data() == nullptr);
        CHECK(c.size() == want.size());
        if (rng() % 4 == 0) { kept.emplace_back(std::move(c), std::move(want)); if (kept.size() > 8) kept.erase(kept.begin()); }
        else { c.clear(); CHECK(c.size() == 0); c.append("ACGT", 4); CHECK(c.size() == 4 && memcmp(c.data(), "ACGT", 4) == 0); }
    }
    for (auto &kv : kept) CHECK(kv.first.size() == kv.second.size() && memcmp(kv.first.data(), kv.second.data(), kv.second.size()) == 0);
}

int main()
{
    {   // a released slab is handed out again
        SeqBuffer a; a.reserve(5u << 20);
        const char *p = a.data(); const size_t cap = a.capacity();
        CHECK(cap >= (5u << 20) && cap % (2u << 20) == 0 && ((uintptr_t)p % (2u << 20)) == 0);
        a = SeqBuffer();
        SeqBuffer b; b.reserve(5u << 20);
        CHECK(b.data() == p && b.capacity() == cap);
        SeqBuffer small; small.reserve(1000);
        CHECK(small.capacity() == 1000);          // heap, exact
    }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < 6; t++) th.emplace_back(one_thread, 100 + t);
    for (auto &t : th) t.join();
    printf("bad %d\n", bad);
    return bad != 0;
}
