// dist_filter.cuh -- partial-key cuckoo filter over the 32-bit dictionary ranks of one reference tile.
//
// Used by dist_probe_kernel (dist.cu): before any (query, reference) pair of a 32-reference tile is merged, the query's
// ranks are looked up in a filter that holds every rank of the tile.  A query none of whose ranks is in the filter shares
// no hash with any of the 32 references, so the reference's merge loop (CommandDistance.cpp:347-365) would end with
// common == 0 for all 32 pairs and only the closed form is written.  The filter has no false negatives; a positive is
// confirmed by an exact binary search before the tile is handed to the merge kernel.
//
// Layout: 2^15 buckets of two 16-bit fingerprints (one 32-bit word per bucket, 128 KB of shared memory), 0 = free slot.
// h = rank * odd constant is a bijection of the 32-bit rank: bucket = low 15 bits (a bijection of the rank's low 15 bits,
// which are uniform for dictionary ranks), fingerprint = top 16 bits (0 -> 1); the alternate bucket is
// bucket ^ f(fingerprint) (partial-key cuckoo hashing), so a lookup is exactly two shared-memory loads and no loop.
// The bit assignment is chosen for the lookup's instruction count: the byte offset of the bucket is (h << 2) & mask and
// the alternate one xor-and of a product, one ALU-pipe instruction each (the probe kernel is bound by that pipe).
// 32 references x 1000 ranks fill 49 % of the slots (two-slot buckets work up to ~84 %).
//
// The functions compile for the host too (tests/test_dist_filter.py drives them through tools/cf_host_test.cpp) --
// there the "atomic" is a plain compare-and-swap.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define MG_HD __host__ __device__ __forceinline__
#else
#define MG_HD inline
#endif

namespace mashgpu {

constexpr uint32_t CF_LOG2_BUCKETS = 15;
constexpr uint32_t CF_BUCKETS = 1u << CF_LOG2_BUCKETS;
constexpr uint32_t CF_MAX_KICKS = 256;

MG_HD uint32_t cf_hash(uint32_t code) { return code * 0x9E3779B1u; }
MG_HD uint32_t cf_fp(uint32_t h) { const uint32_t f = h >> 16; return f ? f : 1u; }
MG_HD uint32_t cf_bucket(uint32_t h) { return h & (CF_BUCKETS - 1u); }
MG_HD uint32_t cf_alt(uint32_t bucket, uint32_t fp) { return bucket ^ ((fp * 0x5BD1E995u) & (CF_BUCKETS - 1u)); }

// does one of the two 16-bit halves of w equal fp (fp != 0)
MG_HD bool cf_word_has(uint32_t w, uint32_t fp)
{
    const uint32_t x = w ^ (fp * 0x00010001u);
    return ((x - 0x00010001u) & ~x & 0x80008000u) != 0;
}

MG_HD uint32_t cf_cas(uint32_t *addr, uint32_t expect, uint32_t value)
{
#if defined(__CUDA_ARCH__)
    return atomicCAS(addr, expect, value);
#else
    const uint32_t old = *addr;
    if (old == expect) *addr = value;
    return old;
#endif
}

MG_HD uint32_t cf_read(const uint32_t *addr)
{
#if defined(__CUDA_ARCH__)
    return *reinterpret_cast<const volatile uint32_t *>(addr);
#else
    return *addr;
#endif
}

// true when fp is (now) in bucket b; false when the bucket is full of other fingerprints
MG_HD bool cf_try_place(uint32_t *tab, uint32_t b, uint32_t fp)
{
    uint32_t w = cf_read(tab + b);
    for (;;) {
        if (cf_word_has(w, fp)) return true;
        uint32_t nw;
        if ((w & 0xFFFFu) == 0) nw = w | fp;
        else if ((w >> 16) == 0) nw = w | (fp << 16);
        else return false;
        const uint32_t old = cf_cas(tab + b, w, nw);
        if (old == w) return true;
        w = old;
    }
}

// Concurrent insert (shared-memory atomics on the device).  Returns false when the random walk gives up; the caller
// then treats the whole tile as "may share" (correct, only slower).
MG_HD bool cf_insert(uint32_t *tab, uint32_t code, uint32_t salt)
{
    const uint32_t h = cf_hash(code);
    uint32_t fp = cf_fp(h);
    const uint32_t b1 = cf_bucket(h), b2 = cf_alt(b1, fp);
    if (cf_try_place(tab, b1, fp)) return true;
    if (cf_try_place(tab, b2, fp)) return true;
    uint32_t rnd = salt * 747796405u + 2891336453u;
    uint32_t b = (rnd >> 31) ? b1 : b2;
    for (uint32_t kick = 0; kick < CF_MAX_KICKS; kick++) {
        rnd = rnd * 1664525u + 1013904223u;
        const uint32_t sh = (rnd >> 31) * 16u;
        uint32_t w = cf_read(tab + b), victim;
        for (;;) {
            if (cf_word_has(w, fp)) return true;           // the same fingerprint arrived by another route
            victim = (w >> sh) & 0xFFFFu;
            const uint32_t nw = (w & ~(0xFFFFu << sh)) | (fp << sh);
            const uint32_t old = cf_cas(tab + b, w, nw);
            if (old == w) break;
            w = old;
        }
        if (victim == 0) return true;                       // the slot had been freed: nothing to move on
        fp = victim;
        b = cf_alt(b, fp);
        if (cf_try_place(tab, b, fp)) return true;
    }
    return false;
}

MG_HD bool cf_lookup(const uint32_t *tab, uint32_t code)
{
    const uint32_t h = cf_hash(code);
    const uint32_t fp = cf_fp(h);
    // byte offsets of the two buckets (same values as 4 * cf_bucket(h) and 4 * cf_alt(cf_bucket(h), fp))
    const uint32_t o1 = (h << 2) & (4u * CF_BUCKETS - 4u);
    const uint32_t o2 = o1 ^ ((fp * (0x5BD1E995u << 2)) & (4u * CF_BUCKETS - 4u));
    const uint32_t f2 = fp * 0x00010001u;
    const uint32_t x1 = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(tab) + o1) ^ f2;
    const uint32_t x2 = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(tab) + o2) ^ f2;
    return ((((x1 - 0x00010001u) & ~x1) | ((x2 - 0x00010001u) & ~x2)) & 0x80008000u) != 0;
}

}  // namespace mashgpu
