// seqbuf.hpp -- sequence buffers of the host shim, drawn from a recycling pool of large mappings.
//
// The parse stage of `mash sketch` turns files into sequence records that live until their batch has been sketched.  With
// std::string records every file meant fresh anonymous memory (malloc maps multi-megabyte strings one by one and unmaps them on
// free): first touch of fresh memory costs ~2 us per 4 KB page on a bare host and far more inside a micro-VM (measured on the
// build container: 0.5 GB/s per thread for never-touched guest memory against 13 GB/s for the 8-thread parse on recycled
// buffers, tools/parse_bench.cpp) -- the page faults, not the parser, bounded file -> .msh.  Records are therefore kept in
// SeqBuffer objects whose storage comes from SlabPool: mappings of whole 2 MiB units,
// returned to a size-ordered free list when the record dies and handed out again to the next record of similar size.  The
// steady-state footprint is one batch plus the parsers' look-ahead window.
#pragma once
#include <sys/mman.h>

#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <utility>

namespace mashhost {

class SlabPool {
public:
    static SlabPool &instance() { static SlabPool *pool = new SlabPool; return *pool; }     // never destroyed: buffers may outlive main()

    // A mapping of at least `want` bytes; `cap` receives its size.  A free slab is taken when it is not more than twice (plus one
    // unit) what was asked for, so a chromosome-sized slab is not spent on a plasmid.
    char *get(size_t want, size_t &cap)
    {
        const size_t need = roundUp(want ? want : 1);
        {
            std::lock_guard<std::mutex> lock(mu_);
            auto it = free_.lower_bound(need);
            if (it != free_.end() && it->first <= 2 * need + kUnit) {
                char *p = it->second;
                cap = it->first;
                retained_ -= cap;
                free_.erase(it);
                return p;
            }
        }
        // fresh mapping, aligned to the huge-page size
        const size_t span = need + kUnit;
        char *raw = static_cast<char *>(mmap(NULL, span, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0));
        if (raw == MAP_FAILED) throw std::bad_alloc();
        char *p = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(raw) + kUnit - 1) & ~(uintptr_t)(kUnit - 1));
        if (p > raw) munmap(raw, (size_t)(p - raw));
        const size_t tail = (size_t)(raw + span - (p + need));
        if (tail) munmap(p + need, tail);
        // (no MADV_HUGEPAGE: recycled slabs are already resident, and for the first pass over fresh memory huge-page faults
        // measured slower and far more variable than 4 KB faults in the build container -- compaction -- 1.0-3.6 against 5.7-11 GB/s)
        cap = need;
        return p;
    }

    void put(char *p, size_t cap)
    {
        if (!p) return;
        {
            std::lock_guard<std::mutex> lock(mu_);
            if (retained_ + cap <= kRetainMax) {
                free_.emplace(cap, p);
                retained_ += cap;
                return;
            }
        }
        munmap(p, cap);
    }

private:
    static constexpr size_t kUnit = (size_t)2 << 20;
    static constexpr size_t kRetainMax = (size_t)8 << 30;      // free slabs kept for reuse; beyond this they go back to the system
    static size_t roundUp(size_t n) { return (n + kUnit - 1) & ~(kUnit - 1); }
    std::mutex mu_;
    std::multimap<size_t, char *> free_;
    size_t retained_ = 0;
};

// The part of std::string the reader and the batch code use, on pool storage.  Move-only.
class SeqBuffer {
public:
    SeqBuffer() {}
    SeqBuffer(SeqBuffer &&o) noexcept : p_(o.p_), n_(o.n_), cap_(o.cap_) { o.p_ = 0; o.n_ = o.cap_ = 0; }
    SeqBuffer &operator=(SeqBuffer &&o) noexcept
    {
        if (this != &o) { release(); p_ = o.p_; n_ = o.n_; cap_ = o.cap_; o.p_ = 0; o.n_ = o.cap_ = 0; }
        return *this;
    }
    SeqBuffer(const SeqBuffer &) = delete;
    SeqBuffer &operator=(const SeqBuffer &) = delete;
    ~SeqBuffer() { release(); }

    const char *data() const { return p_; }
    size_t size() const { return n_; }
    size_t capacity() const { return cap_; }
    void clear() { n_ = 0; }
    void reserve(size_t want)
    {
        if (want <= cap_) return;
        size_t cap = want;
        char *q;
        if (want < kPooledFrom) {                // short records (reads, contigs of an -i run): the heap
            q = static_cast<char *>(malloc(want));
            if (!q) throw std::bad_alloc();
        } else {
            q = SlabPool::instance().get(want, cap);
        }
        if (n_) memcpy(q, p_, n_);
        drop(p_, cap_);
        p_ = q;
        cap_ = cap;
    }
    void append(const char *src, size_t len)
    {
        if (n_ + len > cap_) reserve(std::max(n_ + len, 2 * cap_));
        memcpy(p_ + n_, src, len);
        n_ += len;
    }

private:
    static constexpr size_t kPooledFrom = (size_t)1 << 20;     // capacities from 1 MiB up are pool slabs (whole 2 MiB units)
    static void drop(char *p, size_t cap)
    {
        if (!p) return;
        if (cap < kPooledFrom) free(p); else SlabPool::instance().put(p, cap);
    }
    void release() { drop(p_, cap_); p_ = 0; n_ = cap_ = 0; }
    char *p_ = 0;
    size_t n_ = 0, cap_ = 0;
};

}  // namespace mashhost
