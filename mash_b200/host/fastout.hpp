// fastout.hpp -- text output of the pair grids (writeOutput, reference CommandDistance.cpp:247-304, CommandTriangle.cpp:159-198).
//
// The reference prints one pair per `cout << ... << endl`: a formatted write and a flush per line, ~1 M lines/s on one thread, which
// is of the order of its compare rate.  Behind a GPU that compares 10^10 pairs per second the printing IS the run time of `mash dist`,
// so rows are formatted into memory -- several rows at a time on the `-p` threads, each into its own buffer -- and written in row
// order with one fwrite per buffer.  The bytes are the reference's: a double goes through std::to_chars(general, precision 6),
// which is defined as printf's %.6g, which is what `ostream << double` prints with the default precision; tests/test_host_shim.py
// compares the two on a few million values.
#pragma once
#include <charconv>
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

namespace mashhost {

struct OutBuf {
    std::vector<char> b;
    void str(const std::string &s) { b.insert(b.end(), s.begin(), s.end()); }
    void ch(char c) { b.push_back(c); }
    void dbl(double v)          // == `std::cout << v` (general format, 6 significant digits)
    {
        char tmp[48];
        const auto r = std::to_chars(tmp, tmp + sizeof tmp, v, std::chars_format::general, 6);
        b.insert(b.end(), tmp, r.ptr);
    }
    void u64(uint64_t v)
    {
        char tmp[24];
        const auto r = std::to_chars(tmp, tmp + sizeof tmp, v);
        b.insert(b.end(), tmp, r.ptr);
    }
};

// rows [0, n): fn(row, OutBuf &) appends the text of one row (about `rowCost` pairs each).  Rows are taken in slabs of ~2^17 pairs
// per thread (tens of MB of text at most in memory); within a slab contiguous row ranges are formatted on up to `threads` threads,
// then written to stdout in row order.
template <class F>
void writeRows(uint64_t n, int threads, uint64_t rowCost, F fn)
{
    if (n == 0) return;
    std::cout.flush();                                     // whatever went through cout so far precedes these rows
    const uint64_t T = std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)std::max(1, threads), n));
    const uint64_t slab = std::max<uint64_t>(T, T * (((uint64_t)1 << 17) / std::max<uint64_t>(1, rowCost)));
    std::vector<OutBuf> bufs(T);
    for (uint64_t base = 0; base < n; base += slab) {
        const uint64_t m = std::min(slab, n - base);
        auto work = [&](uint64_t t) {
            const uint64_t lo = base + m * t / T, hi = base + m * (t + 1) / T;
            bufs[t].b.clear();
            for (uint64_t r = lo; r < hi; r++) fn(r, bufs[t]);
        };
        std::vector<std::thread> pool;
        for (uint64_t t = 1; t < T; t++) pool.emplace_back(work, t);
        work(0);
        for (auto &th : pool) th.join();
        for (auto &ob : bufs)
            if (!ob.b.empty()) fwrite(ob.b.data(), 1, ob.b.size(), stdout);
    }
}

}  // namespace mashhost
