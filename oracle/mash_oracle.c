/*
 * mash_oracle.c -- CPU restatement of the marbl/Mash MinHash hot paths.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library, and there
 * only as the checker / CPU baseline.  The product path (mash_b200/csrc, include/mashgpu.h)
 * never links, loads or calls anything in oracle/.
 *
 * Every function restates (does not copy) the algorithm at the cited reference location
 * (paths relative to /root/reference/src/mash/).  Pinning:
 *   - hashes / sketches / shared counts: pinned bit-exactly by test/ref/genomes.json,
 *     test/ref/reads.json, test/ref/genomes.dist, test/ref/screen (tests/test_oracle_golden.py)
 *     and by the reference's own object code (oracle/_ref, tests/test_oracle_vs_ref.py).
 *   - p-values: the reference calls GSL gsl_cdf_binomial_Q / Boost.Math (un-vendored, version
 *     unpinned, CommandDistance.cpp:444-446).  Pinned to the six printed golden values
 *     (6 significant digits) and to 50-digit mpmath fixtures (tests/golden/pvalue_mpmath.json)
 *     at <= 1e-12 relative.  Beyond that: "parity unpinned".
 *   - multiplicity counts: restated incl. the top-of-heap quirk, unpinned by any reference
 *     fixture (stale golden).
 *
 * Plain C11, no dependencies beyond libc/libm.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

#define MO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * MurmurHash3_x64_128 (MurmurHash3.cpp:255-332; fmix64 :81-90).  Public-domain algorithm by
 * Austin Appleby; restated here from its published definition.  Returns h1, optionally h2.
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

static inline uint64_t fmix64(uint64_t k)
{
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

static inline uint64_t load_le(const uint8_t *p, int n)
{
    uint64_t v = 0;
    for (int i = 0; i < n; i++) v |= (uint64_t)p[i] << (8 * i);
    return v;
}

MO_API uint64_t mo_murmur3_x64_128(const void *key, int len, uint32_t seed, uint64_t *h2_out)
{
    const uint8_t *data = (const uint8_t *)key;
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed, h2 = seed;
    int nblocks = len / 16;
    for (int i = 0; i < nblocks; i++) {
        uint64_t k1 = load_le(data + 16 * i, 8), k2 = load_le(data + 16 * i + 8, 8);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t *tail = data + 16 * nblocks;
    int n = len & 15;
    if (n > 8) {
        uint64_t k2 = load_le(tail + 8, n - 8);
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
    }
    if (n > 0) {
        uint64_t k1 = load_le(tail, n > 8 ? 8 : n);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    }
    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    if (h2_out) *h2_out = h2;
    return h1;
}

/* getHash (hash.cpp:10-38): first 8 bytes of the 128-bit digest, or its low 4 bytes. */
MO_API uint64_t mo_get_hash(const char *seq, int len, uint32_t seed, int use64)
{
    uint64_t h1 = mo_murmur3_x64_128(seq, len, seed, NULL);
    return use64 ? h1 : (h1 & 0xffffffffULL);
}

/* ------------------------------------------------------------------------------------------
 * Bottom-s container: MinHashHeap (MinHashHeap.h:10-45, MinHashHeap.cpp:68-146) with
 * multiplicityMinimum == 1 and no Bloom filter; HashSet::toHashList (HashSet.cpp:78-118).
 * A binary max-heap of distinct hashes + an open-addressing map hash -> count.
 * ---------------------------------------------------------------------------------------- */
typedef struct mo_heap {
    uint64_t cap;        /* cardinalityMaximum */
    int use64;
    uint64_t size;       /* distinct hashes held */
    uint64_t *heap;      /* max-heap, heap[0] = top */
    uint64_t heap_n;
    /* open addressing, linear probing, backward-shift deletion */
    uint64_t tcap;       /* power of two */
    uint64_t *tkey;
    uint32_t *tcnt;      /* 0 = empty */
    uint64_t multiplicity_sum;
} mo_heap;

static inline uint64_t slot_of(const mo_heap *h, uint64_t key)
{
    return (key * 0x9E3779B97F4A7C15ULL) >> 11 & (h->tcap - 1);
}

MO_API mo_heap *mo_heap_new(int use64, uint64_t cap)
{
    mo_heap *h = (mo_heap *)calloc(1, sizeof(mo_heap));
    h->cap = cap; h->use64 = use64;
    h->heap = (uint64_t *)malloc(sizeof(uint64_t) * (cap + 2));
    uint64_t t = 16;
    while (t < 4 * (cap + 2)) t <<= 1;
    h->tcap = t;
    h->tkey = (uint64_t *)malloc(sizeof(uint64_t) * t);
    h->tcnt = (uint32_t *)calloc(t, sizeof(uint32_t));
    return h;
}

MO_API void mo_heap_free(mo_heap *h)
{
    if (!h) return;
    free(h->heap); free(h->tkey); free(h->tcnt); free(h);
}

MO_API void mo_heap_clear(mo_heap *h)
{
    h->size = 0; h->heap_n = 0; h->multiplicity_sum = 0;
    memset(h->tcnt, 0, sizeof(uint32_t) * h->tcap);
}

static uint32_t *map_find(mo_heap *h, uint64_t key)
{
    uint64_t i = slot_of(h, key);
    while (h->tcnt[i]) {
        if (h->tkey[i] == key) return &h->tcnt[i];
        i = (i + 1) & (h->tcap - 1);
    }
    return NULL;
}

static void map_insert_new(mo_heap *h, uint64_t key, uint32_t cnt)
{
    uint64_t i = slot_of(h, key);
    while (h->tcnt[i]) i = (i + 1) & (h->tcap - 1);
    h->tkey[i] = key; h->tcnt[i] = cnt;
}

static void map_erase(mo_heap *h, uint64_t key)
{
    uint64_t mask = h->tcap - 1, i = slot_of(h, key);
    while (h->tcnt[i] && h->tkey[i] != key) i = (i + 1) & mask;
    if (!h->tcnt[i]) return;
    uint64_t j = i;
    for (;;) {
        j = (j + 1) & mask;
        if (!h->tcnt[j]) break;
        uint64_t k = slot_of(h, h->tkey[j]);
        /* can entry j move into hole i?  yes unless k lies cyclically in (i, j] */
        int between = (i <= j) ? (i < k && k <= j) : (i < k || k <= j);
        if (!between) { h->tkey[i] = h->tkey[j]; h->tcnt[i] = h->tcnt[j]; i = j; }
    }
    h->tcnt[i] = 0;
}

static void heap_push(mo_heap *h, uint64_t v)
{
    uint64_t i = h->heap_n++;
    while (i > 0) {
        uint64_t p = (i - 1) / 2;
        if (h->heap[p] >= v) break;
        h->heap[i] = h->heap[p]; i = p;
    }
    h->heap[i] = v;
}

static void heap_pop(mo_heap *h)
{
    uint64_t v = h->heap[--h->heap_n], i = 0, n = h->heap_n;
    for (;;) {
        uint64_t c = 2 * i + 1;
        if (c >= n) break;
        if (c + 1 < n && h->heap[c + 1] > h->heap[c]) c++;
        if (h->heap[c] <= v) break;
        h->heap[i] = h->heap[c]; i = c;
    }
    if (n) h->heap[i] = v;
}

/* MinHashHeap::tryInsert (MinHashHeap.cpp:68-146), multiplicityMinimum==1, bloomFilter==0.
 * NB the quirk at :70-74: once full, an occurrence equal to the current top is rejected
 * (hashLessThan is strict) and therefore not counted. */
MO_API void mo_heap_try_insert(mo_heap *h, uint64_t hash)
{
    if (h->cap == 0) return;     /* reference would dereference an empty queue; no caller does this */
    if (h->size < h->cap || hash < h->heap[0]) {
        uint32_t *c = map_find(h, hash);
        if (!c) {
            map_insert_new(h, hash, 1);
            heap_push(h, hash);
            h->size++;
            h->multiplicity_sum += 1;
        } else {
            (*c)++;
            h->multiplicity_sum++;
        }
        if (h->size > h->cap) {
            uint64_t top = h->heap[0];
            uint32_t *tc = map_find(h, top);
            h->multiplicity_sum -= tc ? *tc : 0;
            map_erase(h, top);
            heap_pop(h);
            h->size--;
        }
    }
}

MO_API uint64_t mo_heap_size(const mo_heap *h) { return h->size; }

/* MinHashHeap::estimateSetSize (MinHashHeap.h:45) */
MO_API double mo_heap_estimate_set_size(const mo_heap *h)
{
    if (!h->size) return 0;
    return pow(2.0, h->use64 ? 64.0 : 32.0) * (double)h->size / (double)h->heap[0];
}

/* MinHashHeap::estimateMultiplicity (MinHashHeap.h:44) */
MO_API double mo_heap_estimate_multiplicity(const mo_heap *h)
{
    return h->size ? (double)h->multiplicity_sum / (double)h->size : 0;
}

typedef struct { uint64_t k; uint32_t c; } mo_kc;
static int cmp_kc(const void *a, const void *b)
{
    uint64_t x = ((const mo_kc *)a)->k, y = ((const mo_kc *)b)->k;
    return x < y ? -1 : x > y;
}

/* HashSet::toHashList (HashSet.cpp:78-118): ascending hashes + parallel counts. Returns n. */
MO_API uint64_t mo_heap_to_list(const mo_heap *h, uint64_t *hashes, uint32_t *counts)
{
    mo_kc *v = (mo_kc *)malloc(sizeof(mo_kc) * (h->size + 1));
    uint64_t n = 0;
    for (uint64_t i = 0; i < h->tcap; i++)
        if (h->tcnt[i]) { v[n].k = h->tkey[i]; v[n].c = h->tcnt[i]; n++; }
    qsort(v, n, sizeof(mo_kc), cmp_kc);
    for (uint64_t i = 0; i < n; i++) {
        hashes[i] = v[i].k;
        if (counts) counts[i] = v[i].c;
    }
    free(v);
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Sketch parameters (Sketch.h:34-109) -- only the fields the hot path reads.
 * ---------------------------------------------------------------------------------------- */
typedef struct mo_params {
    int32_t kmer_size;
    uint32_t seed;
    int32_t use64;
    int32_t noncanonical;
    int32_t preserve_case;
    uint8_t alphabet[256];
} mo_params;

/* setAlphabetFromString (Sketch.cpp:1108-1137): fills alphabet[], returns alphabet size and the
 * use64 rule  pow(alphabetSize, k) > pow(2, 32). */
MO_API uint32_t mo_set_alphabet(mo_params *p, const char *characters)
{
    memset(p->alphabet, 0, 256);
    for (const char *c = characters; *c; c++) {
        unsigned char u = (unsigned char)*c;
        if (!p->preserve_case && u > 96 && u < 123) u -= 32;
        p->alphabet[u] = 1;
    }
    uint32_t n = 0;
    for (int i = 0; i < 256; i++) n += p->alphabet[i];
    p->use64 = pow((double)n, (double)p->kmer_size) > pow(2.0, 32.0);
    return n;
}

/* complement[] (Sketch.cpp:1070-1098) restated as a function over 'A'..'Z'. */
static char complement_of(unsigned char c)
{
    static const char tbl[27] = "TVGHNNCDNNMNKNNNNYSAABWNRN";
    if (c >= 'A' && c <= 'Z') return tbl[c - 'A'];
    return 'N'; /* reference indexes out of range here (UB); any such window is skipped anyway */
}

/* addMinHashes (Sketch.cpp:512-583).  Unlike the reference this works on a private copy and
 * leaves `seq` untouched; the sequence of tryInsert calls is identical. */
MO_API void mo_add_min_hashes(mo_heap *heap, const char *seq_in, uint64_t length, const mo_params *p)
{
    int k = p->kmer_size;
    if (length < (uint64_t)k) return; /* callers skip such records (Sketch.cpp:1222-1226) */
    char *seq = (char *)malloc(length ? length : 1);
    char *rev = NULL;
    for (uint64_t i = 0; i < length; i++) {           /* :524-530 */
        char c = seq_in[i];
        if (!p->preserve_case && c > 96 && c < 123) c -= 32;
        seq[i] = c;
    }
    if (!p->noncanonical) {                           /* :532-538, :1100-1106 */
        rev = (char *)malloc(length ? length : 1);
        for (uint64_t i = 0; i < length; i++) rev[i] = complement_of((unsigned char)seq[length - 1 - i]);
    }
    /* :540-576 -- sliding window; `good` = number of consecutive alphabet bytes ending at j */
    uint64_t good = 0;
    for (uint64_t j = 0; j < length; j++) {
        if (!p->alphabet[(unsigned char)seq[j]]) { good = 0; continue; }
        good++;
        if (good < (uint64_t)k) continue;
        uint64_t i = j + 1 - k;
        const char *fwd = seq + i;
        const char *kmer = fwd;
        if (!p->noncanonical) {
            const char *rc = rev + length - i - k;
            if (memcmp(fwd, rc, k) > 0) kmer = rc;    /* :569-571 */
        }
        mo_heap_try_insert(heap, mo_get_hash(kmer, k, p->seed, p->use64));
    }
    free(seq); free(rev);
}

/* Like mo_add_min_hashes but appends every window hash to out[] (for unit tests of the scan).
 * Returns the number of hashes written (<= length-k+1). */
MO_API uint64_t mo_all_hashes(const char *seq_in, uint64_t length, const mo_params *p, uint64_t *out)
{
    int k = p->kmer_size;
    if (length < (uint64_t)k) return 0;
    uint64_t n = 0, good = 0;
    char fwd[40], rc[40];
    for (uint64_t j = 0; j < length; j++) {
        unsigned char c = (unsigned char)seq_in[j];
        if (!p->preserve_case && c > 96 && c < 123) c -= 32;
        if (!p->alphabet[c]) { good = 0; continue; }
        good++;
        if (good < (uint64_t)k) continue;
        uint64_t i = j + 1 - k;
        for (int t = 0; t < k; t++) {
            unsigned char d = (unsigned char)seq_in[i + t];
            if (!p->preserve_case && d > 96 && d < 123) d -= 32;
            fwd[t] = (char)d;
        }
        const char *kmer = fwd;
        if (!p->noncanonical) {
            for (int t = 0; t < k; t++) rc[t] = complement_of((unsigned char)fwd[k - 1 - t]);
            if (memcmp(fwd, rc, k) > 0) kmer = rc;
        }
        out[n++] = mo_get_hash(kmer, k, p->seed, p->use64);
    }
    return n;
}

/* One sketch unit = sketchFile's record loop (Sketch.cpp:1202-1270) / sketchSequence (:1338-1365)
 * over pre-parsed records: records shorter than k are skipped and not counted in length
 * (:1222-1226, :1251-1254).  Output: ascending distinct hashes (+counts), n, length.
 * reads != 0: length = genome_size if nonzero else (uint64_t)estimateSetSize() (:1272-1282). */
MO_API uint64_t mo_sketch_unit(const mo_params *p, uint64_t sketch_size,
                               uint64_t n_records, const char *const *seqs, const uint64_t *lens,
                               int reads, uint64_t genome_size,
                               uint64_t *out_hashes, uint32_t *out_counts, uint64_t *out_length)
{
    mo_heap *h = mo_heap_new(p->use64, sketch_size);
    uint64_t length = 0;
    for (uint64_t r = 0; r < n_records; r++) {
        if (lens[r] < (uint64_t)p->kmer_size) continue;
        if (!reads) length += lens[r];
        mo_add_min_hashes(h, seqs[r], lens[r], p);
    }
    if (reads) length = genome_size ? genome_size : (uint64_t)mo_heap_estimate_set_size(h);
    uint64_t n = mo_heap_to_list(h, out_hashes, out_counts);
    if (out_length) *out_length = length;
    mo_heap_free(h);
    return n;
}

/* ------------------------------------------------------------------------------------------
 * MinHashHeap with multiplicityMinimum > 1 (`mash sketch -m`, MinHashHeap.cpp:96-144): a hash enters the bottom-s only at its
 * m-th accepted occurrence; until then it waits in hashesPending (hash -> count) with a max-queue hashesQueuePending beside it.
 * Restated with two more containers of the same kind as above (an open-addressing count map and a max-heap that may hold
 * "zombies" -- hashes already promoted or purged from the map, :114-118, :131-141).  Bloom filter (-b): not restated.
 * ---------------------------------------------------------------------------------------- */
typedef struct mo_heap_m {
    mo_heap *acc;        /* hashes + hashesQueue */
    mo_heap *pend;       /* hashesPending (map part) + hashesQueuePending (heap part, with zombies) */
    uint64_t m;          /* multiplicityMinimum */
} mo_heap_m;

MO_API mo_heap_m *mo_heap_m_new(int use64, uint64_t cap, uint64_t multiplicity_min)
{
    mo_heap_m *h = (mo_heap_m *)calloc(1, sizeof(mo_heap_m));
    h->acc = mo_heap_new(use64, cap);
    h->pend = mo_heap_new(use64, 64);
    h->m = multiplicity_min < 1 ? 1 : multiplicity_min;
    return h;
}

MO_API void mo_heap_m_free(mo_heap_m *h)
{
    if (!h) return;
    mo_heap_free(h->acc); mo_heap_free(h->pend); free(h);
}

/* the pending containers grow without bound (the reference's do too) */
static void pend_reserve(mo_heap *q)
{
    if (q->heap_n + 2 >= q->cap) {
        q->cap *= 2;
        q->heap = (uint64_t *)realloc(q->heap, sizeof(uint64_t) * (q->cap + 2));
    }
    if (4 * (q->size + 2) >= q->tcap) {          /* rehash at load 1/4 */
        uint64_t ocap = q->tcap, *okey = q->tkey; uint32_t *ocnt = q->tcnt;
        q->tcap = ocap * 4;
        q->tkey = (uint64_t *)malloc(sizeof(uint64_t) * q->tcap);
        q->tcnt = (uint32_t *)calloc(q->tcap, sizeof(uint32_t));
        for (uint64_t i = 0; i < ocap; i++)
            if (ocnt[i]) map_insert_new(q, okey[i], ocnt[i]);
        free(okey); free(ocnt);
    }
}

/* MinHashHeap::tryInsert (MinHashHeap.cpp:68-146) for any multiplicityMinimum, bloomFilter == 0 */
MO_API void mo_heap_m_try_insert(mo_heap_m *hm, uint64_t hash)
{
    mo_heap *h = hm->acc, *q = hm->pend;
    if (h->cap == 0) return;
    if (!(h->size < h->cap || hash < h->heap[0])) return;                  /* :70-74 */
    uint32_t *c = map_find(h, hash);
    if (!c) {                                                              /* :76 hashes.count(hash) == 0 */
        uint32_t *pc = map_find(q, hash);
        const uint64_t pending = pc ? *pc : 0;
        if (hm->m == 1 || pending == hm->m - 1) {                          /* :96 */
            map_insert_new(h, hash, (uint32_t)hm->m);                      /* :98 hashes.insert(hash, multiplicityMinimum) */
            heap_push(h, hash);
            h->size++;
            h->multiplicity_sum += hm->m;
            if (hm->m > 1 && pc) { map_erase(q, hash); q->size--; }        /* :102-108 (stays in the pending queue as a zombie) */
        } else {
            if (!pc) {                                                     /* :112-115 */
                pend_reserve(q);
                heap_push(q, hash);
                map_insert_new(q, hash, 1);                                /* :117 hashesPending.insert(hash, 1) */
                q->size++;
            } else {
                (*pc)++;
            }
        }
    } else {                                                               /* :120-124 */
        (*c)++;
        h->multiplicity_sum++;
    }
    if (h->size > h->cap) {                                                /* :126-144 */
        const uint64_t top = h->heap[0];
        uint32_t *tc = map_find(h, top);
        h->multiplicity_sum -= tc ? *tc : 0;
        map_erase(h, top);
        while (q->heap_n > 0 && top < q->heap[0]) {                        /* :133 hashLessThan(hashesQueue.top(), hashesQueuePending.top()) */
            if (map_find(q, q->heap[0])) { map_erase(q, q->heap[0]); q->size--; }
            heap_pop(q);
        }
        heap_pop(h);
        h->size--;
    }
}

/* sketchFile's record loop with `-m min_copies` (Sketch.cpp:1186 constructs MinHashHeap(use64, s, reads ? minCov : 1, ...)) */
MO_API uint64_t mo_sketch_unit_m(const mo_params *p, uint64_t sketch_size, uint64_t min_copies,
                                 uint64_t n_records, const char *const *seqs, const uint64_t *lens,
                                 int reads, uint64_t genome_size,
                                 uint64_t *out_hashes, uint32_t *out_counts, uint64_t *out_length)
{
    mo_heap_m *hm = mo_heap_m_new(p->use64, sketch_size, min_copies);
    uint64_t length = 0;
    uint64_t *buf = NULL, cap = 0;
    for (uint64_t r = 0; r < n_records; r++) {
        if (lens[r] < (uint64_t)p->kmer_size) continue;
        if (!reads) length += lens[r];
        if (lens[r] > cap) { cap = lens[r]; buf = (uint64_t *)realloc(buf, sizeof(uint64_t) * cap); }
        const uint64_t n = mo_all_hashes(seqs[r], lens[r], p, buf);        /* the hashes addMinHashes feeds to tryInsert, in order */
        for (uint64_t i = 0; i < n; i++) mo_heap_m_try_insert(hm, buf[i]);
    }
    free(buf);
    if (reads) length = genome_size ? genome_size : (uint64_t)mo_heap_estimate_set_size(hm->acc);
    const uint64_t n = mo_heap_to_list(hm->acc, out_hashes, out_counts);
    if (out_length) *out_length = length;
    mo_heap_m_free(hm);
    return n;
}

/* sketchFile's record loop with `-m min_copies -c target_cov` (Sketch.cpp:1186-1282): after every kept record the loop stops when
 * estimateMultiplicity() >= targetCov (:1258-1262, reads mode only).  *out_records_used = records fed to the heap (the "Reads
 * used" line, :1324-1327). */
MO_API uint64_t mo_sketch_unit_mc(const mo_params *p, uint64_t sketch_size, uint64_t min_copies, double target_cov,
                                  uint64_t n_records, const char *const *seqs, const uint64_t *lens,
                                  int reads, uint64_t genome_size,
                                  uint64_t *out_hashes, uint32_t *out_counts, uint64_t *out_length, uint64_t *out_records_used)
{
    mo_heap_m *hm = mo_heap_m_new(p->use64, sketch_size, min_copies);
    uint64_t length = 0, used = 0;
    uint64_t *buf = NULL, cap = 0;
    for (uint64_t r = 0; r < n_records; r++) {
        if (lens[r] < (uint64_t)p->kmer_size) continue;
        if (!reads) length += lens[r];
        if (lens[r] > cap) { cap = lens[r]; buf = (uint64_t *)realloc(buf, sizeof(uint64_t) * cap); }
        const uint64_t n = mo_all_hashes(seqs[r], lens[r], p, buf);
        for (uint64_t i = 0; i < n; i++) mo_heap_m_try_insert(hm, buf[i]);
        used++;
        if (reads && target_cov > 0 && mo_heap_estimate_multiplicity(hm->acc) >= target_cov) break;     /* :1258-1262 */
    }
    free(buf);
    if (reads) length = genome_size ? genome_size : (uint64_t)mo_heap_estimate_set_size(hm->acc);
    const uint64_t n = mo_heap_to_list(hm->acc, out_hashes, out_counts);
    if (out_length) *out_length = length;
    if (out_records_used) *out_records_used = used;
    mo_heap_m_free(hm);
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Binomial upper tail P[Bin(n, r) >= x]  ==  gsl_cdf_binomial_Q(x-1, r, n)
 * (call sites CommandDistance.cpp:444-446, CommandScreen.cpp:611-613).  GSL/Boost are not in
 * the reference tree; this is an independent evaluation of the same mathematical quantity
 * (direct summation of the pmf with a rescaled running product for the first term, in long
 * double), pinned against mpmath in tests.
 * ---------------------------------------------------------------------------------------- */
MO_API double mo_binomial_upper_tail(uint64_t x, double r_in, uint64_t n)
{
    if (x == 0) return 1.0;
    if (x > n) return 0.0;
    long double r = r_in;
    if (r <= 0.0L) return 0.0;
    if (r >= 1.0L) return 1.0;
    long double q = 1.0L - r, odds = r / q;
    if ((long double)x >= ((long double)n + 1.0L) * r) {
        /* first term C(n,x) r^x q^(n-x) as mantissa * 2^e */
        long double m = 1.0L; long e = 0;
        for (uint64_t i = 1; i <= x; i++) {
            m *= ((long double)(n - x + i) / (long double)i) * r;
            if (m < 0x1p-4000L) { m *= 0x1p4000L; e -= 4000; }
            if (m > 0x1p4000L) { m *= 0x1p-4000L; e += 4000; }
        }
        /* q^(n-x) = 2^(y/ln2) */
        long double y = (long double)(n - x) * log1pl(-r) / 0.693147180559945309417232121458176568L;
        long double yi = floorl(y);
        m *= exp2l(y - yi); e += (long)yi;
        /* sum the decreasing terms */
        long double sum = 1.0L, t = 1.0L;
        for (uint64_t i = x; i < n; i++) {
            t *= ((long double)(n - i) / (long double)(i + 1)) * odds;
            sum += t;
            if (t < sum * 1e-22L) break;
        }
        long double res = ldexpl(m * sum, (int)(e < -40000 ? -40000 : e));
        return (double)res;
    } else {
        /* below the mean: 1 - P[X <= x-1], summed downwards from x-1 */
        long double m = 1.0L; long e = 0;
        uint64_t xm = x - 1;
        for (uint64_t i = 1; i <= xm; i++) {
            m *= ((long double)(n - xm + i) / (long double)i) * r;
            if (m < 0x1p-4000L) { m *= 0x1p4000L; e -= 4000; }
            if (m > 0x1p4000L) { m *= 0x1p-4000L; e += 4000; }
        }
        long double y = (long double)(n - xm) * log1pl(-r) / 0.693147180559945309417232121458176568L;
        long double yi = floorl(y);
        m *= exp2l(y - yi); e += (long)yi;
        long double sum = 1.0L, t = 1.0L;
        for (uint64_t i = xm; i > 0; i--) {
            t *= ((long double)i / (long double)(n - i + 1)) / odds;
            sum += t;
            if (t < sum * 1e-22L) break;
        }
        long double lower = ldexpl(m * sum, (int)(e < -40000 ? -40000 : e));
        return (double)(1.0L - lower);
    }
}

/* pValue (CommandDistance.cpp:427-448) */
MO_API double mo_pvalue(uint64_t x, uint64_t length_ref, uint64_t length_query, double kmer_space, uint64_t sketch_size)
{
    if (x == 0) return 1.0;
    double pX = 1. / (1. + kmer_space / length_ref);
    double pY = 1. / (1. + kmer_space / length_query);
    double r = pX * pY / (pX + pY - pX * pY);
    return mo_binomial_upper_tail(x, r, sketch_size);
}

/* pValueWithin (CommandScreen.cpp:601-615) */
MO_API double mo_pvalue_within(uint64_t x, uint64_t set_size, double kmer_space, uint64_t sketch_size)
{
    if (x == 0) return 1.0;
    double r = (double)set_size / kmer_space;
    return mo_binomial_upper_tail(x, r, sketch_size);
}

/* estimateIdentity (CommandScreen.cpp:463-482) */
MO_API double mo_estimate_identity(uint64_t common, uint64_t denom, int kmer_size)
{
    double jaccard = (double)common / (double)denom;
    if (common == denom) return 1.;
    if (common == 0) return 0.;
    return pow(jaccard, 1. / kmer_size);
}

typedef struct mo_pair_output {   /* PairOutput (CommandDistance.h:63-70) */
    uint64_t numer;
    uint64_t denom;
    double distance;
    double pvalue;
    int32_t pass;
    int32_t filled;               /* 0 when the reference leaves numer/denom/distance/pValue unset */
} mo_pair_output;

/* compareSketches (CommandDistance.cpp:336-425).  Hash lists are ascending u64 (32-bit hashes
 * widened), as HashList stores them. */
MO_API void mo_compare_sketches(mo_pair_output *out,
                                const uint64_t *ref, uint64_t n_ref, uint64_t len_ref,
                                const uint64_t *qry, uint64_t n_qry, uint64_t len_qry,
                                uint64_t sketch_size, int kmer_size, double kmer_space,
                                double max_distance, double max_pvalue)
{
    uint64_t i = 0, j = 0, common = 0, denom = 0;
    out->pass = 0; out->filled = 0;
    out->numer = 0; out->denom = 0; out->distance = 0; out->pvalue = 0;
    while (denom < sketch_size && i < n_ref && j < n_qry) {        /* :347-365 */
        if (ref[i] < qry[j]) i++;
        else if (qry[j] < ref[i]) j++;
        else { i++; j++; common++; }
        denom++;
    }
    if (denom < sketch_size) {                                     /* :367-385 */
        if (i < n_ref) denom += n_ref - i;
        if (j < n_qry) denom += n_qry - j;
        if (denom > sketch_size) denom = sketch_size;
    }
    double distance, jaccard = (double)common / (double)denom;     /* :387-407 */
    if (common == denom) distance = 0;
    else if (common == 0) distance = 1.;
    else {
        distance = -log(2 * jaccard / (1. + jaccard)) / kmer_size;
        if (distance > 1) distance = 1;
    }
    if (max_distance >= 0 && distance > max_distance) return;      /* :409-412 */
    out->filled = 1;
    out->numer = common; out->denom = denom; out->distance = distance;
    out->pvalue = mo_pvalue(common, len_ref, len_qry, kmer_space, denom);
    if (max_pvalue >= 0 && out->pvalue > max_pvalue) return;       /* :419-422 */
    out->pass = 1;
}

/* compare (CommandDistance.cpp:306-334) over the whole grid: query-major linear pair index
 * (CommandDistance.cpp:213-232).  Sketch sets are dense arrays with `stride` u64 per sketch. */
MO_API void mo_compare_all(mo_pair_output *out,
                           const uint64_t *ref, const uint32_t *ref_n, const uint64_t *ref_len, uint64_t n_ref, uint64_t ref_stride,
                           const uint64_t *qry, const uint32_t *qry_n, const uint64_t *qry_len, uint64_t n_qry, uint64_t qry_stride,
                           uint64_t sketch_size, int kmer_size, double kmer_space,
                           double max_distance, double max_pvalue,
                           uint64_t q_begin, uint64_t q_end)
{
    for (uint64_t q = q_begin; q < q_end && q < n_qry; q++)
        for (uint64_t r = 0; r < n_ref; r++)
            mo_compare_sketches(&out[q * n_ref + r],
                                ref + r * ref_stride, ref_n[r], ref_len[r],
                                qry + q * qry_stride, qry_n[q], qry_len[q],
                                sketch_size, kmer_size, kmer_space, max_distance, max_pvalue);
}

/* ------------------------------------------------------------------------------------------
 * Screen (CommandScreen.cpp).  Reference hash table: sorted array of distinct reference hashes
 * (role of hashCounts, :93-114) with a parallel u32 counter array.
 * ---------------------------------------------------------------------------------------- */
static int64_t find_sorted(const uint64_t *keys, uint64_t n, uint64_t key)
{
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t mid = (lo + hi) / 2;
        if (keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    return (lo < n && keys[lo] == key) ? (int64_t)lo : -1;
}

/* hashSequence (CommandScreen.cpp:484-599), nucleotide (trans == false) path: every valid window
 * of the '*'-joined chunk is hashed, inserted into the chunk's heap and, if it is a reference hash,
 * its counter is bumped (:571-575). */
MO_API void mo_hash_sequence(mo_heap *heap, const uint64_t *keys, uint32_t *counts, uint64_t n_keys,
                             const char *seq_in, uint64_t length, const mo_params *p)
{
    int k = p->kmer_size;
    if (length < (uint64_t)k) return;
    char fwd[40], rc[40];
    uint64_t good = 0;
    for (uint64_t j = 0; j < length; j++) {
        unsigned char c = (unsigned char)seq_in[j];
        if (!p->preserve_case && c > 96 && c < 123) c -= 32;
        if (!p->alphabet[c]) { good = 0; continue; }
        good++;
        if (good < (uint64_t)k) continue;
        uint64_t i = j + 1 - k;
        for (int t = 0; t < k; t++) {
            unsigned char d = (unsigned char)seq_in[i + t];
            if (!p->preserve_case && d > 96 && d < 123) d -= 32;
            fwd[t] = (char)d;
        }
        const char *kmer = fwd;
        if (!p->noncanonical) {
            for (int t = 0; t < k; t++) rc[t] = complement_of((unsigned char)fwd[k - 1 - t]);
            if (memcmp(fwd, rc, k) > 0) kmer = rc;
        }
        uint64_t hash = mo_get_hash(kmer, k, p->seed, p->use64);
        if (heap) mo_heap_try_insert(heap, hash);
        int64_t at = find_sorted(keys, n_keys, hash);
        if (at >= 0) __atomic_fetch_add(&counts[at], 1u, __ATOMIC_RELAXED);
    }
}

static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}
static int cmp_u32(const void *a, const void *b)
{
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return x < y ? -1 : x > y;
}

/* Build the distinct-hash table of a sketch set (CommandScreen.cpp:93-114).  keys must hold
 * sum(ref_n) entries; returns the number of distinct keys (ascending). */
MO_API uint64_t mo_screen_build_table(const uint64_t *ref, const uint32_t *ref_n, uint64_t n_ref, uint64_t stride, uint64_t *keys)
{
    uint64_t n = 0;
    for (uint64_t r = 0; r < n_ref; r++)
        for (uint32_t i = 0; i < ref_n[r]; i++) keys[n++] = ref[r * stride + i];
    qsort(keys, n, sizeof(uint64_t), cmp_u64);
    uint64_t m = 0;
    for (uint64_t i = 0; i < n; i++)
        if (i == 0 || keys[i] != keys[i - 1]) keys[m++] = keys[i];
    return m;
}

/* Screen reduce (CommandScreen.cpp:322-355, 409-455): per reference sketch shared count, median
 * depth depths[shared/2], identity and p-value.  set_size = (uint64_t)estimateSetSize() of the
 * merged mixture heap (:322).  minCov == 1 (:152). */
MO_API void mo_screen_finish(const uint64_t *ref, const uint32_t *ref_n, uint64_t n_ref, uint64_t stride,
                             const uint64_t *keys, const uint32_t *counts, uint64_t n_keys,
                             uint64_t set_size, int kmer_size, double kmer_space,
                             uint64_t *shared, uint64_t *median, double *identity, double *pvalue)
{
    uint32_t *depths = (uint32_t *)malloc(sizeof(uint32_t) * (stride + 1));
    for (uint64_t r = 0; r < n_ref; r++) {
        uint64_t s = 0;
        for (uint32_t i = 0; i < ref_n[r]; i++) {
            int64_t at = find_sorted(keys, n_keys, ref[r * stride + i]);
            if (at >= 0 && counts[at] >= 1) depths[s++] = counts[at];
        }
        qsort(depths, s, sizeof(uint32_t), cmp_u32);
        shared[r] = s;
        median[r] = s > 0 ? depths[s / 2] : 0;
        identity[r] = mo_estimate_identity(s, ref_n[r], kmer_size);
        pvalue[r] = mo_pvalue_within(s, set_size, kmer_space, ref_n[r]);
    }
    free(depths);
}

/* Screen reduce with `-w` ("winner take all", CommandScreen.cpp:357-407): after the plain shared counts, every reference
 * hash seen in the mixture is re-assigned to ONE of the sketches that contain it -- the one with the highest identity
 * estimate (scores[] = estimateIdentity of the plain shared count, :361-364), ties going to the longer genome (:392-396) --
 * and shared / depths are rebuilt from those assignments (:366-404); median, identity and p-value then follow from the
 * new counts as in mo_screen_finish (:409-436).  The reference walks each hash's sketches in the iteration order of a
 * robin_hood::unordered_set, so a tie in BOTH score and length is broken by an order this restatement does not
 * reproduce: here the lowest sketch index wins (parity unpinned for exact ties; the tests avoid equal lengths). */
MO_API void mo_screen_finish_winner(const uint64_t *ref, const uint32_t *ref_n, const uint64_t *ref_len, uint64_t n_ref, uint64_t stride,
                                    const uint64_t *keys, const uint32_t *counts, uint64_t n_keys,
                                    uint64_t set_size, int kmer_size, double kmer_space,
                                    uint64_t *shared, uint64_t *median, double *identity, double *pvalue)
{
    double *scores = (double *)malloc(sizeof(double) * (n_ref ? n_ref : 1));
    int64_t *winner = (int64_t *)malloc(sizeof(int64_t) * (n_keys ? n_keys : 1));
    for (uint64_t k = 0; k < n_keys; k++) winner[k] = -1;
    for (uint64_t r = 0; r < n_ref; r++) {             /* plain shared counts -> scores */
        uint64_t s = 0;
        for (uint32_t i = 0; i < ref_n[r]; i++) {
            int64_t at = find_sorted(keys, n_keys, ref[r * stride + i]);
            if (at >= 0 && counts[at] >= 1) s++;
        }
        scores[r] = mo_estimate_identity(s, ref_n[r], kmer_size);
    }
    for (uint64_t r = 0; r < n_ref; r++)               /* ascending index: a full tie keeps the earlier sketch */
        for (uint32_t i = 0; i < ref_n[r]; i++) {
            int64_t at = find_sorted(keys, n_keys, ref[r * stride + i]);
            if (at < 0 || counts[at] < 1) continue;
            int64_t w = winner[at];
            if (w < 0 || scores[r] > scores[w] || (scores[r] == scores[w] && ref_len[r] > ref_len[w])) winner[at] = (int64_t)r;
        }
    uint32_t *depths = (uint32_t *)malloc(sizeof(uint32_t) * (stride + 1));
    for (uint64_t r = 0; r < n_ref; r++) {
        uint64_t s = 0;
        for (uint32_t i = 0; i < ref_n[r]; i++) {
            int64_t at = find_sorted(keys, n_keys, ref[r * stride + i]);
            if (at >= 0 && counts[at] >= 1 && winner[at] == (int64_t)r) depths[s++] = counts[at];
        }
        qsort(depths, s, sizeof(uint32_t), cmp_u32);
        shared[r] = s;
        median[r] = s > 0 ? depths[s / 2] : 0;
        identity[r] = mo_estimate_identity(s, ref_n[r], kmer_size);
        pvalue[r] = mo_pvalue_within(s, set_size, kmer_space, ref_n[r]);
    }
    free(depths); free(winner); free(scores);
}
