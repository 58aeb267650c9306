// pack.cpp -- host feed path: ASCII records -> 2-bit packed stream + list of invalid runs.
//
// What crosses PCIe per base drops from 1 byte to 0.25 byte (+ a few bytes per run of non-ACGT bytes).  This is a
// format conversion only -- the same information addMinHashes derives per base before it hashes anything
// (upper-casing Sketch.cpp:524-530, alphabet test :548-556): code = A0 C1 G2 T3, everything else "invalid".  No
// k-mer, hash or sketch logic runs on the host.
//
// Stream layout (positions = the flat stream of sketch_stream_core: units back to back, one separator position after
// every record): codes[p / 32] holds base p at bits 2*(p % 32); invalid positions (non-alphabet bytes and
// separators) are reported as runs {start, length} and expanded into a bit mask on the device.
#include "pack.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <pthread.h>
#include <thread>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace mashgpu {

namespace {

// Worker threads that outlive a call.  A pack call used to start and join its threads itself: ~14 thread creations per call cost
// as much as packing 300 MB does (measured through the screen feed on the 16-CPU box: 8.2 ms per 302 MB chunk against 2.6 ms of
// packing at the rate the 1 GiB sketch waves reach), so small jobs -- screen chunks -- could not pay for themselves.
class PackPool {
public:
    // never destroyed: workers may be parked at exit.  A forked child has none of the parent's threads: it starts over with a new
    // pool (the old object is leaked; its mutexes may have been held at the time of the fork).
    static PackPool &instance()
    {
        static std::once_flag once;
        std::call_once(once, [] {
            current() = new PackPool;
            pthread_atfork(nullptr, nullptr, [] { current() = new PackPool; });
        });
        return *current();
    }
    // `work` is run on `threads` threads in total, the caller included; it pulls its items from its own atomic counter
    void run(int threads, const std::function<void()> &work)
    {
        std::lock_guard<std::mutex> one_job(run_mu_);
        const int helpers = std::max(0, threads - 1);
        {
            std::unique_lock<std::mutex> lock(mu_);
            while ((int)workers_.size() < helpers) {
                const int id = (int)workers_.size();
                workers_.emplace_back([this, id] { loop(id); });
                workers_.back().detach();
            }
            job_ = &work; want_ = helpers; active_ = helpers; gen_++;
        }
        cv_work_.notify_all();
        work();
        std::unique_lock<std::mutex> lock(mu_);
        cv_done_.wait(lock, [&] { return active_ == 0; });
        job_ = nullptr;
    }

private:
    static PackPool *&current() { static PackPool *p = nullptr; return p; }
    void loop(int id)
    {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void()> *job;
            {
                std::unique_lock<std::mutex> lock(mu_);
                cv_work_.wait(lock, [&] { return gen_ != seen; });
                seen = gen_;
                if (id >= want_) continue;
                job = job_;
            }
            (*job)();
            std::lock_guard<std::mutex> lock(mu_);
            if (--active_ == 0) cv_done_.notify_all();
        }
    }
    std::mutex run_mu_, mu_;
    std::condition_variable cv_work_, cv_done_;
    std::vector<std::thread> workers_;
    const std::function<void()> *job_ = nullptr;
    int want_ = 0, active_ = 0;
    uint64_t gen_ = 0;
};

struct Lut {
    uint8_t v[2][256];   // [preserve_case][byte] -> code (0..3) or 4 = invalid
    Lut()
    {
        for (int pc = 0; pc < 2; pc++)
            for (int b = 0; b < 256; b++) {
                int u = (!pc && b > 96 && b < 123) ? b - 32 : b;
                v[pc][b] = u == 'A' ? 0 : u == 'C' ? 1 : u == 'G' ? 2 : u == 'T' ? 3 : 4;
            }
    }
};
const Lut g_lut;

// scalar: n bytes (n <= 32) -> codes (2 bits each) + invalid bit mask
inline void pack_scalar(const uint8_t *s, int n, int preserve_case, uint64_t &codes, uint32_t &inval)
{
    const uint8_t *lut = g_lut.v[preserve_case ? 1 : 0];
    uint64_t c = 0;
    uint32_t m = 0;
    for (int i = 0; i < n; i++) {
        uint8_t x = lut[s[i]];
        c |= (uint64_t)(x & 3) << (2 * i);
        m |= (uint32_t)(x >> 2) << i;
    }
    codes = c;
    inval = m;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) inline void pack_avx2(const uint8_t *s, int preserve_case, uint64_t &codes, uint32_t &inval)
{
    __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s));
    if (!preserve_case) {
        // only a/c/g/t can become A/C/G/T by clearing bit 5; other bytes stay outside the alphabet either way
        v = _mm256_and_si256(v, _mm256_set1_epi8((char)0xDF));
    }
    // c' = (v >> 1) & 3 : A0 C1 T2 G3 ; expected byte for c' via pshufb ; code = c' ^ (c' >> 1) : A0 C1 G2 T3
    const __m256i c = _mm256_and_si256(_mm256_srli_epi16(v, 1), _mm256_set1_epi8(3));
    const __m256i expect = _mm256_shuffle_epi8(_mm256_setr_epi8('A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                                               'A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0), c);
    const uint32_t valid = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(expect, v));
    const __m256i code = _mm256_xor_si256(c, _mm256_and_si256(_mm256_srli_epi16(c, 1), _mm256_set1_epi8(1)));
    // gather 2 bits per byte: pairs (b0 + 4 b1) -> 16-bit lanes, pairs of those (w0 + 16 w1) -> 32-bit lanes holding 4 bases in a byte
    const __m256i w = _mm256_maddubs_epi16(code, _mm256_set1_epi16(0x0401));
    const __m256i d = _mm256_madd_epi16(w, _mm256_set1_epi32(0x00100001));
    // low byte of each of the 8 dwords -> 8 bytes
    const __m256i sh = _mm256_shuffle_epi8(d, _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                                               0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1));
    const uint32_t lo = (uint32_t)_mm256_extract_epi32(sh, 0), hi = (uint32_t)_mm256_extract_epi32(sh, 4);
    codes = (uint64_t)lo | ((uint64_t)hi << 32);
    inval = ~valid;
}

// 64 bases per step: the bit planes of ASCII bits 1 and 2 (A 00, C 01, T 10, G 11) come out of two byte tests as 64-bit masks,
// code = (b1 ^ b2) + 2 b2 (A0 C1 G2 T3), and PDEP interleaves the two planes into 2-bit codes -- about half the instructions
// per base of the AVX2 routine above
__attribute__((target("avx512f,avx512bw,bmi2"))) inline void pack_avx512(const uint8_t *s, int preserve_case, uint64_t &c0, uint64_t &c1, uint64_t &inval)
{
    __m512i v = _mm512_loadu_si512(reinterpret_cast<const void *>(s));
    if (!preserve_case) v = _mm512_and_si512(v, _mm512_set1_epi8((char)0xDF));
    const __m512i c = _mm512_and_si512(_mm512_srli_epi16(v, 1), _mm512_set1_epi8(3));
    const __m512i expect = _mm512_shuffle_epi8(_mm512_broadcast_i32x4(_mm_setr_epi8('A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)), c);
    const uint64_t valid = _mm512_cmpeq_epi8_mask(expect, v);
    const uint64_t b1 = _mm512_test_epi8_mask(v, _mm512_set1_epi8(2)), b2 = _mm512_test_epi8_mask(v, _mm512_set1_epi8(4));
    const uint64_t lo = b1 ^ b2, hi = b2;
    c0 = _pdep_u64(lo & 0xFFFFFFFFull, 0x5555555555555555ull) | _pdep_u64(hi & 0xFFFFFFFFull, 0xAAAAAAAAAAAAAAAAull);
    c1 = _pdep_u64(lo >> 32, 0x5555555555555555ull) | _pdep_u64(hi >> 32, 0xAAAAAAAAAAAAAAAAull);
    inval = ~valid;
}
#endif

bool have_avx512()
{
#if defined(__x86_64__)
    static const bool ok = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("bmi2");
    return ok;
#else
    return false;
#endif
}

bool have_avx2()
{
#if defined(__x86_64__)
    static const bool ok = __builtin_cpu_supports("avx2");
    return ok;
#else
    return false;
#endif
}

// Packs stream positions [p0, p1) (p0 multiple of 32) of a virtual stream made of segments; positions not covered by a
// segment are separators (invalid).  Appends invalid runs to `runs` in position order.
struct Segment { const uint8_t *src; uint64_t pos, len; };

void pack_range(const Segment *segs, size_t nseg, uint64_t p0, uint64_t p1, int preserve_case, uint64_t *codes, std::vector<PackRun> &runs)
{
    const bool avx2 = have_avx2();
    size_t si = 0;
    // first segment that ends after p0
    {
        size_t lo = 0, hi = nseg;
        while (lo < hi) { size_t mid = (lo + hi) / 2; if (segs[mid].pos + segs[mid].len <= p0) lo = mid + 1; else hi = mid; }
        si = lo;
    }
    uint64_t run_start = 0, run_len = 0;
    auto add_invalid = [&](uint64_t pos, uint64_t len) {
        if (run_len && run_start + run_len == pos) { run_len += len; return; }
        if (run_len) runs.push_back(PackRun{run_start, run_len});
        run_start = pos; run_len = len;
    };
    auto add_mask = [&](uint64_t g, uint32_t m) {          // invalid positions of one 32-base group
        if (m == 0xFFFFFFFFu) { add_invalid(g, 32); return; }
        for (int i = 0; i < 32;) {
            if (!((m >> i) & 1)) { i++; continue; }
            int j = i;
            while (j < 32 && ((m >> j) & 1)) j++;
            add_invalid(g + i, j - i);
            i = j;
        }
    };
#if defined(__x86_64__)
    const bool avx512 = have_avx512();
#endif
    for (uint64_t g = p0; g < p1; g += 32) {
        const uint64_t gend = g + 32 < p1 ? g + 32 : p1;
        uint64_t c = 0;
        while (si < nseg && segs[si].pos + segs[si].len <= g) si++;
#if defined(__x86_64__)
        if (avx512 && si < nseg && segs[si].pos <= g) {
            // as many 64-base steps as lie inside this record and this range
            const uint64_t end = std::min(segs[si].pos + segs[si].len, p1);
            const uint64_t n64 = end > g ? (end - g) / 64 : 0;
            if (n64) {
                const uint8_t *s = segs[si].src + (g - segs[si].pos);
                uint64_t *out = codes + (g - p0) / 32;
                for (uint64_t t = 0; t < n64; t++, s += 64, out += 2) {
                    uint64_t m;
                    pack_avx512(s, preserve_case, out[0], out[1], m);
                    if (m) {
                        const uint64_t gg = g + 64 * t;
                        if ((uint32_t)m) add_mask(gg, (uint32_t)m);
                        if (m >> 32) add_mask(gg + 32, (uint32_t)(m >> 32));
                    }
                }
                g += 64 * n64 - 32;      // the loop header adds the last 32
                continue;
            }
        }
#endif
        if (si < nseg && segs[si].pos <= g && segs[si].pos + segs[si].len >= g + 32 && gend == g + 32) {
            // whole group inside one record: the fast path
            uint32_t m;
            const uint8_t *s = segs[si].src + (g - segs[si].pos);
#if defined(__x86_64__)
            if (avx2) pack_avx2(s, preserve_case, c, m); else
#endif
                pack_scalar(s, 32, preserve_case, c, m);
            if (m) add_mask(g, m);
        } else {
            // group touches a record boundary / separator / the end: per position
            size_t sj = si;
            for (uint64_t p = g; p < gend; p++) {
                while (sj < nseg && segs[sj].pos + segs[sj].len <= p) sj++;
                uint8_t x = 4;
                if (sj < nseg && segs[sj].pos <= p) x = g_lut.v[preserve_case ? 1 : 0][segs[sj].src[p - segs[sj].pos]];
                if (x & 4) add_invalid(p, 1);
                c |= (uint64_t)(x & 3) << (2 * (p - g));
            }
        }
        codes[(g - p0) / 32] = c;
    }
    if (run_len) runs.push_back(PackRun{run_start, run_len});
}

}  // namespace

void pack_stream(const PackSegment *segments, size_t n_segments, uint64_t stream_len, int preserve_case, int threads,
                 uint64_t *codes, std::vector<PackRun> &runs)
{
    runs.clear();
    const uint64_t groups = (stream_len + 31) / 32;
    if (groups == 0) return;
    static_assert(sizeof(PackSegment) == sizeof(Segment), "layout");
    const Segment *segs = reinterpret_cast<const Segment *>(segments);
    if (threads < 1) threads = 1;
    const uint64_t min_groups = 1 << 15;    // at least 1 Mbase per thread
    uint64_t nt = std::min<uint64_t>((uint64_t)threads, (groups + min_groups - 1) / min_groups);
    if (nt <= 1) {
        pack_range(segs, n_segments, 0, stream_len, preserve_case, codes, runs);
        return;
    }
    // dynamic chunks so that threads slowed by page faults / NUMA do not hold up the rest
    const uint64_t chunk_groups = std::max<uint64_t>(min_groups, groups / (nt * 8));
    const uint64_t n_chunks = (groups + chunk_groups - 1) / chunk_groups;
    std::vector<std::vector<PackRun>> chunk_runs(n_chunks);
    std::atomic<uint64_t> next(0);
    auto work = [&]() {
        for (;;) {
            uint64_t ci = next.fetch_add(1);
            if (ci >= n_chunks) return;
            uint64_t g0 = ci * chunk_groups, g1 = std::min(groups, g0 + chunk_groups);
            pack_range(segs, n_segments, g0 * 32, std::min(stream_len, g1 * 32), preserve_case, codes + g0, chunk_runs[ci]);
        }
    };
    PackPool::instance().run((int)nt, work);
    for (auto &cr : chunk_runs)
        for (auto &r : cr) {
            if (!runs.empty() && runs.back().start + runs.back().len == r.start) runs.back().len += r.len;
            else runs.push_back(r);
        }
}

void pack_chunk_mask(const uint8_t *src, uint64_t len, int preserve_case, int threads, uint64_t *codes, uint32_t *inval)
{
    const uint64_t groups = (len + 31) / 32;
    if (groups == 0) return;
    auto range = [&](uint64_t g0, uint64_t g1) {            // groups [g0, g1)
        uint64_t g = g0;
#if defined(__x86_64__)
        if (have_avx512())
            for (; g + 2 <= g1 && (g + 2) * 32 <= len; g += 2) {
                uint64_t m;
                pack_avx512(src + g * 32, preserve_case, codes[g], codes[g + 1], m);
                inval[g] = (uint32_t)m;
                inval[g + 1] = (uint32_t)(m >> 32);
            }
        if (have_avx2())
            for (; g < g1 && (g + 1) * 32 <= len; g++) pack_avx2(src + g * 32, preserve_case, codes[g], inval[g]);
#endif
        for (; g < g1; g++) {
            const int n = (int)std::min<uint64_t>(32, len - g * 32);
            pack_scalar(src + g * 32, n, preserve_case, codes[g], inval[g]);
            if (n < 32) inval[g] |= ~0u << n;
        }
    };
    const uint64_t min_groups = 1 << 15;
    const uint64_t nt = std::min<uint64_t>((uint64_t)std::max(1, threads), (groups + min_groups - 1) / min_groups);
    if (nt <= 1) { range(0, groups); return; }
    const uint64_t chunk_groups = std::max<uint64_t>(min_groups, groups / (nt * 8));
    const uint64_t n_chunks = (groups + chunk_groups - 1) / chunk_groups;
    std::atomic<uint64_t> next(0);
    auto work = [&]() {
        for (;;) {
            const uint64_t ci = next.fetch_add(1);
            if (ci >= n_chunks) return;
            range(ci * chunk_groups, std::min(groups, (ci + 1) * chunk_groups));
        }
    };
    PackPool::instance().run((int)nt, work);
}

int host_pack_threads()
{
    // a container's CPU quota can be far below the visible core count, and threads beyond it are throttled together (measured on
    // this pool's B200 box, 128 vCPUs visible, quota 16: 61 GB/s with 16 threads, 13 GB/s with 128; tools/pack_bench.py)
    int threads = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 96u);
    if (FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r")) {
        long long quota = 0, period = 0;
        if (fscanf(f, "%lld %lld", &quota, &period) == 2 && quota > 0 && period > 0)
            threads = std::min(threads, (int)std::max(1ll, (quota + period - 1) / period));
        fclose(f);
    }
    if (const char *t = getenv("MASHGPU_PACK_THREADS")) threads = std::max(1, atoi(t));
    return threads;
}

}  // namespace mashgpu
