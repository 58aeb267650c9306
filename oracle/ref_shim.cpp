// ref_shim.cpp -- extern "C" access to the REFERENCE'S OWN object code (test infrastructure).
//
// Linked against marbl/Mash's MurmurHash3.cpp, hash.cpp, HashList.cpp, HashSet.cpp,
// HashPriorityQueue.cpp and MinHashHeap.cpp, compiled in place from /root/reference by
// oracle/Makefile into oracle/_ref/libmash_ref.so (git-ignored; never copied into the repo).
// Sketch.cpp / CommandDistance.cpp / CommandScreen.cpp cannot be compiled here (they need the
// generated Cap'n Proto header and GSL/Boost), so the scan loops around the reference's
// getHash() and MinHashHeap are restated below with their source lines cited.
//
// Used for (1) validating oracle/mash_oracle.c against the reference's own hash + heap code and
// (2) the multi-threaded CPU baseline timed by bench.py (`--impl reference`, cpu_baseline).
// The product path never loads this.
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>

#include <zlib.h>

#include "hash.h"
#include "HashList.h"
#include "MinHashHeap.h"
#include "kseq.h"
KSEQ_INIT(gzFile, gzread)            // as the reference instantiates its parser (Sketch.cpp:21)

#define REF_API extern "C" __attribute__((visibility("default")))

struct ref_params {      // same layout as mo_params in mash_oracle.c
    int32_t kmer_size;
    uint32_t seed;
    int32_t use64;
    int32_t noncanonical;
    int32_t preserve_case;
    uint8_t alphabet[256];
};

REF_API uint64_t ref_get_hash(const char *seq, int len, uint32_t seed, int use64)
{
    hash_u h = getHash(seq, len, seed, use64 != 0);   // hash.cpp:10-38 (reference object code)
    return use64 ? h.hash64 : (uint64_t)h.hash32;
}

static char complement_of(unsigned char c)            // Sketch.cpp:1070-1098
{
    static const char tbl[27] = "TVGHNNCDNNMNKNNNNYSAABWNRN";
    return (c >= 'A' && c <= 'Z') ? tbl[c - 'A'] : 'N';
}

// addMinHashes (Sketch.cpp:512-583) restated around the reference's getHash + MinHashHeap::tryInsert.
// Same passes as the reference: upper-case copy, full reverse-complement copy, sliding window.
static void add_min_hashes(MinHashHeap &heap, const char *seq_in, uint64_t length, const ref_params &p)
{
    const int k = p.kmer_size;
    if (length < (uint64_t)k) return;
    std::vector<char> seq(seq_in, seq_in + length), rev;
    if (!p.preserve_case)
        for (uint64_t i = 0; i < length; i++)
            if (seq[i] > 96 && seq[i] < 123) seq[i] -= 32;
    if (!p.noncanonical) {
        rev.resize(length);
        for (uint64_t i = 0; i < length; i++) rev[i] = complement_of((unsigned char)seq[length - 1 - i]);
    }
    uint64_t j = 0;
    for (uint64_t i = 0; i + k <= length; i++) {
        bool bad = false;
        for (; j < i + k; j++) {
            if (!p.alphabet[(unsigned char)seq[j]]) { i = j++; bad = true; break; }
        }
        if (bad) continue;
        const char *fwd = seq.data() + i;
        const char *kmer = fwd;
        if (!p.noncanonical) {
            const char *rc = rev.data() + length - i - k;
            if (memcmp(fwd, rc, k) > 0) kmer = rc;
        }
        heap.tryInsert(getHash(kmer, k, p.seed, p.use64 != 0));
    }
}

static uint64_t emit(const MinHashHeap &heap, bool use64, uint64_t *out_hashes, uint32_t *out_counts)
{
    HashList list(use64);
    std::vector<uint32_t> counts;
    heap.toHashList(list, counts);                     // HashSet.cpp:78-118 (reference object code)
    for (int i = 0; i < list.size(); i++) {
        out_hashes[i] = use64 ? list.at(i).hash64 : (uint64_t)list.at(i).hash32;
        if (out_counts) out_counts[i] = counts[i];
    }
    return (uint64_t)list.size();
}

// One unit = sketchFile's record loop (Sketch.cpp:1202-1282) on pre-parsed records.
REF_API uint64_t ref_sketch_unit(const ref_params *p, uint64_t sketch_size,
                                 uint64_t n_records, const char *const *seqs, const uint64_t *lens,
                                 int reads, uint64_t genome_size,
                                 uint64_t *out_hashes, uint32_t *out_counts, uint64_t *out_length)
{
    MinHashHeap heap(p->use64 != 0, sketch_size, 1, 0);
    uint64_t length = 0;
    for (uint64_t r = 0; r < n_records; r++) {
        if (lens[r] < (uint64_t)p->kmer_size) continue;
        if (!reads) length += lens[r];
        add_min_hashes(heap, seqs[r], lens[r], *p);
    }
    if (reads) length = genome_size ? genome_size : (uint64_t)heap.estimateSetSize();
    if (out_length) *out_length = length;
    return emit(heap, p->use64 != 0, out_hashes, out_counts);
}

// Multi-threaded CPU baseline: one job per unit (single record each), work-stealing counter in
// place of the reference ThreadPool (Sketch.cpp:109,211).  out_hashes is n_units x sketch_size.
REF_API void ref_sketch_many(const ref_params *p, uint64_t sketch_size, uint64_t n_units,
                             const char *const *seqs, const uint64_t *lens, int threads,
                             uint64_t *out_hashes, uint32_t *out_n)
{
    std::atomic<uint64_t> next(0);
    auto work = [&]() {
        for (;;) {
            uint64_t u = next.fetch_add(1);
            if (u >= n_units) return;
            uint64_t len = 0;
            out_n[u] = (uint32_t)ref_sketch_unit(p, sketch_size, 1, &seqs[u], &lens[u], 0, 0,
                                                 out_hashes + u * sketch_size, nullptr, &len);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(work);
    for (auto &t : pool) t.join();
}

// sketchFile's own loop for ONE file per sketch (Sketch.cpp:1186-1282, default non-reads mode): the reference's parser
// (kseq_read) feeds addMinHashes record by record; records shorter than k are skipped and not counted in the length.
// Returns the number of hashes; *out_length = Reference::length.  -1 when the file cannot be opened / is truncated.
static int64_t sketch_one_file(const ref_params *p, uint64_t sketch_size, const char *path, uint64_t *out_hashes, uint64_t *out_length)
{
    gzFile fp = gzopen(path, "r");
    if (!fp) return -1;
    kseq_t *ks = kseq_init(fp);
    MinHashHeap heap(p->use64 != 0, sketch_size, 1, 0);
    uint64_t length = 0;
    int l;
    while ((l = kseq_read(ks)) >= 0) {
        if (l < p->kmer_size) continue;
        length += (uint64_t)l;
        add_min_hashes(heap, ks->seq.s, (uint64_t)l, *p);
    }
    kseq_destroy(ks);
    gzclose(fp);
    if (l != -1) return -1;
    if (out_length) *out_length = length;
    return (int64_t)emit(heap, p->use64 != 0, out_hashes, nullptr);
}

// The CPU arm as `mash sketch -p threads file1 file2 ...` runs it: one job per file, parse included.
REF_API int ref_sketch_files(const ref_params *p, uint64_t sketch_size, uint64_t n_files, const char *const *paths, int threads,
                             uint64_t *out_hashes, uint32_t *out_n, uint64_t *out_length)
{
    std::atomic<uint64_t> next(0);
    std::atomic<int> failed(0);
    auto work = [&]() {
        for (;;) {
            uint64_t u = next.fetch_add(1);
            if (u >= n_files) return;
            int64_t n = sketch_one_file(p, sketch_size, paths[u], out_hashes + u * sketch_size, out_length ? out_length + u : nullptr);
            if (n < 0) { failed = 1; out_n[u] = 0; } else out_n[u] = (uint32_t)n;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(work);
    for (auto &t : pool) t.join();
    return failed.load();
}

// hashSequence (CommandScreen.cpp:484-599) nucleotide path around reference getHash/MinHashHeap;
// table = sorted distinct reference hashes + u32 counters (role of hashCounts, :93-114).
REF_API void ref_hash_sequence(const uint64_t *keys, uint32_t *counts, uint64_t n_keys,
                               const char *seq_in, uint64_t length, const ref_params *p,
                               uint64_t sketch_size, uint64_t *out_hashes, uint32_t *out_n)
{
    MinHashHeap heap(p->use64 != 0, sketch_size);
    const int k = p->kmer_size;
    if (length >= (uint64_t)k) {
        std::vector<char> seq(seq_in, seq_in + length), rev(length);
        if (!p->preserve_case)
            for (uint64_t i = 0; i < length; i++)
                if (seq[i] > 96 && seq[i] < 123) seq[i] -= 32;
        for (uint64_t i = 0; i < length; i++) rev[i] = complement_of((unsigned char)seq[length - 1 - i]);
        int64_t lastGood = -1;
        const int64_t len = (int64_t)length;
        for (int64_t j = 0; j < len - k + 1; j++) {
            while (lastGood < j + k - 1 && lastGood < len - 1) {
                lastGood++;
                if (!p->alphabet[(unsigned char)seq[lastGood]]) j = lastGood + 1;
            }
            if (j > len - k) break;
            const char *fwd = seq.data() + j;
            const char *rc = rev.data() + len - j - k;
            const char *kmer = (p->noncanonical || memcmp(fwd, rc, k) <= 0) ? fwd : rc;
            hash_u h = getHash(kmer, k, p->seed, p->use64 != 0);
            heap.tryInsert(h);
            uint64_t key = p->use64 ? h.hash64 : (uint64_t)h.hash32;
            uint64_t lo = 0, hi = n_keys;
            while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (keys[mid] < key) lo = mid + 1; else hi = mid; }
            if (lo < n_keys && keys[lo] == key) __atomic_fetch_add(&counts[lo], 1u, __ATOMIC_RELAXED);
        }
    }
    if (out_hashes) *out_n = (uint32_t)emit(heap, p->use64 != 0, out_hashes, nullptr);
}

// ---------------------------------------------------------------------------------------------------------
// Screen with the reference's own table type: robin_hood::unordered_map<uint64_t, std::atomic<uint32_t>> hashCounts
// (CommandScreen.cpp:94, 103-110), probed as hashSequence does (:571-575: count(key) == 1 -> hashCounts[key]++).
// Multi-threaded CPU arm of bench.py: one HashInput per chunk (:224-262), one MinHashHeap per worker (:114-119, 239).
// ---------------------------------------------------------------------------------------------------------
#include "robin_hood.h"

struct ref_screen_table {
    robin_hood::unordered_map<uint64_t, std::atomic<uint32_t>> counts;
};

REF_API ref_screen_table *ref_screen_table_new(const uint64_t *keys, uint64_t n_keys)
{
    ref_screen_table *t = new ref_screen_table();
    for (uint64_t i = 0; i < n_keys; i++) t->counts[keys[i]] = 0;
    return t;
}

REF_API void ref_screen_table_free(ref_screen_table *t) { delete t; }

REF_API void ref_screen_table_counts(ref_screen_table *t, const uint64_t *keys, uint64_t n_keys, uint32_t *out)
{
    for (uint64_t i = 0; i < n_keys; i++) {
        auto it = t->counts.find(keys[i]);
        out[i] = it == t->counts.end() ? 0u : it->second.load();
    }
}

static void hash_sequence_map(ref_screen_table *t, MinHashHeap &heap, const char *seq_in, uint64_t length, const ref_params *p)
{
    const int k = p->kmer_size;
    if (length < (uint64_t)k) return;
    std::vector<char> seq(seq_in, seq_in + length), rev(length);
    if (!p->preserve_case)
        for (uint64_t i = 0; i < length; i++)
            if (seq[i] > 96 && seq[i] < 123) seq[i] -= 32;
    for (uint64_t i = 0; i < length; i++) rev[i] = complement_of((unsigned char)seq[length - 1 - i]);
    int64_t lastGood = -1;
    const int64_t len = (int64_t)length;
    for (int64_t j = 0; j < len - k + 1; j++) {
        while (lastGood < j + k - 1 && lastGood < len - 1) {
            lastGood++;
            if (!p->alphabet[(unsigned char)seq[lastGood]]) j = lastGood + 1;
        }
        if (j > len - k) break;
        const char *fwd = seq.data() + j;
        const char *rc = rev.data() + len - j - k;
        const char *kmer = (p->noncanonical || memcmp(fwd, rc, k) <= 0) ? fwd : rc;
        hash_u h = getHash(kmer, k, p->seed, p->use64 != 0);
        heap.tryInsert(h);
        uint64_t key = p->use64 ? h.hash64 : (uint64_t)h.hash32;
        if (t->counts.count(key) == 1) t->counts[key]++;          // CommandScreen.cpp:571-575
    }
}

// chunks: n_chunks '*'-joined read blocks; `threads` workers take chunks in order.  out_hashes/out_n (nullable): bottom-s of
// the whole stream (heaps merged as CommandScreen.cpp:288-302).
REF_API void ref_screen_many(ref_screen_table *t, uint64_t n_chunks, const char *const *chunks, const uint64_t *lens,
                             const ref_params *p, uint64_t sketch_size, int threads, uint64_t *out_hashes, uint32_t *out_n)
{
    std::atomic<uint64_t> next(0);
    std::vector<MinHashHeap *> heaps;
    for (int i = 0; i < threads; i++) heaps.push_back(new MinHashHeap(p->use64 != 0, sketch_size));
    auto work = [&](int w) {
        for (;;) {
            uint64_t c = next.fetch_add(1);
            if (c >= n_chunks) return;
            hash_sequence_map(t, *heaps[w], chunks[c], lens[c], p);
        }
    };
    std::vector<std::thread> pool;
    for (int w = 0; w < threads; w++) pool.emplace_back(work, w);
    for (auto &th : pool) th.join();
    MinHashHeap merged(p->use64 != 0, sketch_size);
    for (MinHashHeap *h : heaps) {
        HashList list(p->use64 != 0);
        h->toHashList(list);
        for (int i = 0; i < list.size(); i++) merged.tryInsert(list.at(i));
        delete h;
    }
    if (out_hashes && out_n) *out_n = (uint32_t)emit(merged, p->use64 != 0, out_hashes, nullptr);
}

// `mash sketch -m min_copies` / `-b bloom_bytes`: the record loop of ref_sketch_unit around the reference's own MinHashHeap
// constructed as sketchFile does (Sketch.cpp:1186: MinHashHeap(use64, minHashesPerWindow, reads ? minCov : 1, memoryBound)).
REF_API uint64_t ref_sketch_unit_m(const ref_params *p, uint64_t sketch_size, uint64_t min_copies, uint64_t bloom_bytes,
                                   uint64_t n_records, const char *const *seqs, const uint64_t *lens,
                                   int reads, uint64_t genome_size,
                                   uint64_t *out_hashes, uint32_t *out_counts, uint64_t *out_length)
{
    MinHashHeap heap(p->use64 != 0, sketch_size, min_copies, bloom_bytes);
    uint64_t length = 0;
    for (uint64_t r = 0; r < n_records; r++) {
        if (lens[r] < (uint64_t)p->kmer_size) continue;
        if (!reads) length += lens[r];
        add_min_hashes(heap, seqs[r], lens[r], *p);
    }
    if (reads) length = genome_size ? genome_size : (uint64_t)heap.estimateSetSize();
    if (out_length) *out_length = length;
    return emit(heap, p->use64 != 0, out_hashes, out_counts);
}

// ... and with `-c target_cov`: the loop stops after the first kept record that brings estimateMultiplicity() to the target
// (Sketch.cpp:1258-1262).  *out_records_used = kept records fed to the heap.
REF_API uint64_t ref_sketch_unit_mc(const ref_params *p, uint64_t sketch_size, uint64_t min_copies, double target_cov,
                                    uint64_t n_records, const char *const *seqs, const uint64_t *lens,
                                    int reads, uint64_t genome_size,
                                    uint64_t *out_hashes, uint32_t *out_counts, uint64_t *out_length, uint64_t *out_records_used)
{
    MinHashHeap heap(p->use64 != 0, sketch_size, min_copies, 0);
    uint64_t length = 0, used = 0;
    for (uint64_t r = 0; r < n_records; r++) {
        if (lens[r] < (uint64_t)p->kmer_size) continue;
        if (!reads) length += lens[r];
        add_min_hashes(heap, seqs[r], lens[r], *p);
        used++;
        if (reads && target_cov > 0 && heap.estimateMultiplicity() >= target_cov) break;
    }
    if (reads) length = genome_size ? genome_size : (uint64_t)heap.estimateSetSize();
    if (out_length) *out_length = length;
    if (out_records_used) *out_records_used = used;
    return emit(heap, p->use64 != 0, out_hashes, out_counts);
}
