// scan.cuh -- the k-mer scan + hash + filter kernel (K1 of DESIGN.md), shared by sketch and screen.
//
// Replaces the per-base loop of addMinHashes (reference Sketch.cpp:512-583) and hashSequence
// (CommandScreen.cpp:484-599):  upper-case -> alphabet check -> canonical k-mer -> getHash -> tryInsert.
//
// Layout.  The input is one flat byte stream in HBM (units back to back, records separated by a byte
// outside the alphabet).  A CTA processes TILE window-start positions at a time: the tile's ASCII bytes
// are read once with 128-bit loads, converted to 4-bit codes (A,C,G,T -> 0,1,2,3, anything else -> 8)
// and staged in shared memory as nibble words (8 bases per 32-bit word).  A thread takes a block of 8
// consecutive window starts: it loads the BW words that cover them, builds the reverse-complement block
// in registers (nibble reversal + xor 3), and for each of the 8 windows extracts the forward and the
// reverse-complement k-mer with funnel shifts, picks the canonical one by an integer compare of the
// little-endian nibble words (equivalent to the reference's memcmp because complementing reverses the
// base order), expands it to ASCII with one PRMT per 4 bases and runs MurmurHash3_x64_128.
// Only hashes at or below the tile's coarse threshold leave the fast path.
#pragma once
#include <cstdint>
#include "murmur3.cuh"

namespace mashgpu {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_TILE = 8192;              // window starts per tile
constexpr int SCAN_HALO_WORDS = 8;           // extra nibble words (64 bases) staged past the tile
constexpr int SCAN_WORDS = SCAN_TILE / 8 + SCAN_HALO_WORDS;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
constexpr int SCAN_WARP_TILE = 1024;         // window starts per warp tile (32 lanes x 4 blocks x 8)
constexpr int SCAN_WARP_WORDS = SCAN_WARP_TILE / 8 + SCAN_HALO_WORDS;   // 136 nibble words
constexpr int SCAN_WARP_VECS = SCAN_WARP_WORDS / 2;                     // 68 16-byte vectors
constexpr uint64_t EMPTY_KEY = 0xFFFFFFFFFFFFFFFFULL;

enum ScanMode : int { SCAN_SKETCH = 0, SCAN_SCREEN = 1, SCAN_DUMP = 2, SCAN_COUNT = 3, SCAN_EVENTS = 4 };

struct ScanArgs {
    const uint8_t *stream;        // flat byte stream (ASCII source; NULL when the packed source is used)
    const uint64_t *codes;        // packed source: 2-bit codes, 32 positions per word (A0 C1 G2 T3), see pack.cpp
    const uint32_t *inval;        // packed source: 1 bit per position, set = not in the alphabet / separator / past the end
    uint64_t stream_len;          // positions >= stream_len are treated as separators
    uint64_t tile_begin, tile_end;
    const uint64_t *tile_tmax;    // per-tile coarse threshold (indexed by absolute tile id); NULL -> coarse_t
    uint64_t coarse_t;
    uint32_t seed;
    int use64;
    int preserve_case;
    int mode;
    // candidate tables (sketch units / the screen mixture): open addressing, keys EMPTY_KEY when free
    const uint64_t *unit_start;   // n_units + 1 stream offsets
    uint32_t n_units;
    const uint64_t *unit_t;       // per-unit threshold: keep hash <= unit_t[u]
    const uint64_t *tab_off;      // per-unit slot offset into tab_keys / tab_cnt
    const uint32_t *tab_log2;     // per-unit log2(capacity)
    uint64_t *tab_keys;
    uint32_t *tab_cnt;
    uint64_t *tab_first;          // optional (multiplicity counts): stream position of the first / last occurrence per slot,
    uint64_t *tab_last;           //   needed to reproduce MinHashHeap's top-of-heap counting quirk (sketch.cu, quirk_kernel)
    uint32_t *unit_flags;         // bit0: table overflow
    uint32_t *unit_maxhash;       // occurrences of the hash value 2^64-1 (cannot be a table key)
    int64_t only_unit;            // >= 0: ignore every other unit (exact re-run)
    uint32_t min_copies;          // `-m` (>= 1): tab_first holds min_copies positions per slot, the smallest stream positions of the key
    // screen: reference hash table (distinct keys, EMPTY_KEY when free) with u32 hit counters
    const uint64_t *ref_keys;
    const uint32_t *ref_idx;      // slot -> index of the key in the sorted distinct key list (deterministic across ranks)
    uint32_t *ref_cnt;            // hit counters, indexed by that key index
    uint32_t ref_log2;
    uint64_t ref_hmax;            // largest reference hash (probe prefilter)
    const uint32_t *ref_bitmap;   // optional second prefilter: bit (hash >> ref_bitmap_shift) is set iff some reference hash has these
    uint32_t ref_bitmap_shift;    //   top bits (indexed by value, not hashed: reference hashes crowd the low end of the range)
    uint64_t screen_mix_t;        // screen: largest unit threshold -- only hashes at or below it can enter the mixture's bottom-s
    // dump
    uint64_t *out_hash;
    uint8_t *out_valid;
    // byte-alphabet kernels (any alphabet, non-canonical): byte -> itself (upper-cased unless preserve_case) if it is in
    // the alphabet, else 0
    uint8_t byte_lut[256];
    // SCAN_COUNT: occurrences of one hash at stream positions [count_lo, count_hi]
    uint64_t count_target, count_lo, count_hi;
    uint32_t *count_out;
    // SCAN_EVENTS (`-c`): every k-mer whose hash is at or below the threshold of its position band (the "units" of this pass)
    // is appended as an event {position, hash}; sketch.cu replays the events in stream order through the heap
    uint64_t *ev_pos, *ev_hash;
    unsigned long long *ev_count;
    uint64_t ev_capacity;
};

__device__ __forceinline__ uint32_t slot_hash(uint64_t key, uint32_t log2cap)
{
    return (uint32_t)((key * 0x9E3779B97F4A7C15ULL) >> (64 - log2cap));
}

// Insert-or-count into an open-addressing table. Returns the slot, or -1 when the table is full.
__device__ __forceinline__ int64_t table_add(uint64_t *keys, uint32_t *cnt, uint32_t log2cap, uint64_t key)
{
    const uint32_t mask = (1u << log2cap) - 1;
    uint32_t slot = slot_hash(key, log2cap);
    for (uint32_t probes = 0; probes <= mask; probes++) {
        unsigned long long prev = atomicCAS((unsigned long long *)&keys[slot], (unsigned long long)EMPTY_KEY, (unsigned long long)key);
        if (prev == EMPTY_KEY || prev == key) {
            atomicAdd(&cnt[slot], 1u);
            return slot;
        }
        slot = (slot + 1) & mask;
    }
    return -1;
}

// Slow path: a hash passed the tile's coarse threshold.
static __device__ __noinline__ void scan_emit(const ScanArgs &a, uint32_t hash_lo, uint32_t hash_hi, uint64_t tile_base, uint32_t local_pos)
{
    const uint64_t hash = ((uint64_t)hash_hi << 32) | hash_lo;
    const uint64_t pos = tile_base + local_pos;
    if (a.mode == SCAN_DUMP) {
        a.out_hash[pos] = hash;
        a.out_valid[pos] = 1;
        return;
    }
    if (a.mode == SCAN_COUNT) {
        if (hash == a.count_target && pos >= a.count_lo && pos <= a.count_hi) atomicAdd(a.count_out, 1u);
        return;
    }
    // (screen: the reference table was probed by screen_probe_lanes before the survivors were serialised)
    // unit of this position: last u with unit_start[u] <= pos
    uint32_t lo = 0, hi = a.n_units;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (a.unit_start[mid] <= pos) lo = mid; else hi = mid;
    }
    const uint32_t u = lo;
    if (a.only_unit >= 0 && (int64_t)u != a.only_unit) return;
    if (hash > a.unit_t[u]) return;
    if (a.mode == SCAN_EVENTS) {
        const unsigned long long at = atomicAdd(a.ev_count, 1ull);
        if (at < a.ev_capacity) { a.ev_pos[at] = pos; a.ev_hash[at] = hash; }
        return;
    }
    if (hash == EMPTY_KEY) { atomicAdd(&a.unit_maxhash[u], 1u); return; }
    const int64_t slot = table_add(a.tab_keys + a.tab_off[u], a.tab_cnt + a.tab_off[u], a.tab_log2[u], hash);
    if (slot < 0) { atomicOr(&a.unit_flags[u], 1u); return; }
    if (a.tab_first) {
        // the m smallest positions of the key, ascending: a cascade of atomic minima -- level i ends up with the (i+1)-th smallest
        // position whatever the arrival order, because every value that reaches a level is either its final minimum or is passed
        // on exactly once.  Level m-1 is the position at which MinHashHeap promotes the hash (MinHashHeap.cpp:96-100); m = 1: the
        // first occurrence.
        unsigned long long *f = (unsigned long long *)a.tab_first + (a.tab_off[u] + (uint64_t)slot) * a.min_copies;
        unsigned long long x = pos;
        for (uint32_t i = 0; i < a.min_copies; i++) {
            const unsigned long long old = atomicMin(&f[i], x);
            if (old == EMPTY_KEY) break;            // the level was empty: nothing to pass on
            if (old > x) x = old;
        }
        atomicMax((unsigned long long *)&a.tab_last[a.tab_off[u] + slot], (unsigned long long)pos);
    }
}

// Warp-uniform wrapper around the slow path.  The hot loop never branches on a per-lane condition: a window's pass
// flag is voted, and when any lane has a survivor the whole warp enters here and serialises the survivors with
// ballot/shuffle (lane 0 does the table work).  Keeping the main loop free of divergent branches matters: with a
// per-lane `if (pass) emit()` the warp kept running split into sub-warps between emits (ncu: 19 active lanes on
// average, 1.6x the warp instructions).
// The fast path does not even test window validity (2 instructions per k-mer saved): hashes of windows that contain
// a non-alphabet byte are garbage, and the few that slip under the threshold are rejected here by re-reading the
// window's nibbles from the warp's shared-memory tile.  The fast-path filter only compares the high word (k > 16)
// or the low word (k <= 16) of the hash with the tile threshold; the exact 64-bit test is in scan_emit.
// window validity from the warp's tile: DNA kernels keep nibbles (bit 3 = outside the alphabet), byte kernels bytes (0 = outside)
__device__ __forceinline__ bool window_valid(const uint32_t *sm_tile, uint32_t lp, int k)
{
    if (k > 0) {
        uint32_t bad = 0;
        for (int i = 0; i < k; i++) bad |= sm_tile[(lp + i) >> 3] >> (4 * ((lp + i) & 7));
        return !(bad & 8u);
    }
    const uint8_t *bytes = reinterpret_cast<const uint8_t *>(sm_tile);
    bool ok = true;
    for (int i = 0; i < -k; i++) ok &= bytes[lp + i] != 0;
    return ok;
}

// Screen: hashCounts[key]++ iff key is a reference hash (CommandScreen.cpp:571-575).  Every lane probes for its own window --
// with a reference set that spans the whole hash range (small genomes, plasmids, viruses in the .msh: their bottom-s reaches
// up to 2^64 s / L) EVERY k-mer of the mixture is a candidate, so this cannot go through the serialised survivor path.
// Order: largest reference hash -> value-indexed bitmap (L2 resident, rejects the sparse upper range without touching the
// table) -> open-addressing table in HBM -> on a hit only, the window's validity (the fast path hashes invalid windows too).
static __device__ __noinline__ void screen_probe_lanes(const ScanArgs &a, bool pass, uint32_t hash_lo, uint32_t hash_hi,
                                                       uint32_t local_pos, const uint32_t *sm_tile, int k)
{
    const uint64_t hash = ((uint64_t)hash_hi << 32) | hash_lo;
    bool go = pass && hash <= a.ref_hmax;
    if (go && a.ref_bitmap) {
        const uint64_t b = hash >> a.ref_bitmap_shift;
        go = (__ldg(a.ref_bitmap + (b >> 5)) >> (b & 31)) & 1u;
    }
    if (!go) return;
    const uint32_t mask = (1u << a.ref_log2) - 1;
    uint32_t slot = slot_hash(hash, a.ref_log2);
    for (;;) {
        const uint64_t key = __ldg(a.ref_keys + slot);
        if (key == hash) {
            if (window_valid(sm_tile, local_pos, k)) atomicAdd(&a.ref_cnt[a.ref_idx[slot]], 1u);
            return;
        }
        if (key == EMPTY_KEY) return;
        slot = (slot + 1) & mask;
    }
}

static __device__ __noinline__ void scan_emit_warp(const ScanArgs &a, bool pass, uint32_t hash_lo, uint32_t hash_hi,
                                                   uint64_t tile_base, uint32_t local_pos, const uint32_t *sm_tile, int k,
                                                   uint64_t tmax)
{
    if (a.mode == SCAN_SCREEN) {
        screen_probe_lanes(a, pass, hash_lo, hash_hi, local_pos, sm_tile, k);
        __syncwarp();
        pass = pass && ((((uint64_t)hash_hi) << 32) | hash_lo) <= a.screen_mix_t;      // the rest is about the mixture's bottom-s only
    }
    unsigned m = __ballot_sync(0xFFFFFFFFu, pass);
    const int lane = threadIdx.x & 31;
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t lo = __shfl_sync(0xFFFFFFFFu, hash_lo, src);
        const uint32_t hi = __shfl_sync(0xFFFFFFFFu, hash_hi, src);
        const uint32_t lp = __shfl_sync(0xFFFFFFFFu, local_pos, src);
        if (lane == 0) {
            if (window_valid(sm_tile, lp, k) && ((((uint64_t)hi) << 32) | lo) <= tmax) scan_emit(a, lo, hi, tile_base, lp);
        }
        __syncwarp();
    }
}

// 4 ASCII bytes -> 4 nibbles (in the low 16 bits).  Codes: A=0 C=1 G=2 T=3 (order preserving, complement = xor 3),
// anything else 8.  `fold` = 0xDFDFDFDF to upper-case first (a..z -> A..Z is the only effect that matters), or ~0.
__device__ __forceinline__ uint32_t ascii4_to_nibbles(uint32_t w, uint32_t fold)
{
    const uint32_t u = w & fold;
    const uint32_t t = (u >> 2) & ~(u >> 1) & 0x01010101u;                    // bit2 & ~bit1: 'T'
    const uint32_t expect = (0x41414141u | (u & 0x06060606u)) ^ (t * 0x11u);  // the ACGT byte with these bits 1,2
    const uint32_t bad = u ^ expect;
    const uint32_t nz = (((bad & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | bad) & 0x80808080u;  // bit7 set where byte != 0
    const uint32_t c = (u >> 1) & 0x03030303u;                                // A0 C1 T2 G3
    const uint32_t code = c ^ ((c >> 1) & 0x01010101u);                       // A0 C1 G2 T3
    uint32_t x = code | (nz >> 4);
    x = (x | (x >> 4)) & 0x00FF00FFu;
    x = (x | (x >> 8)) & 0x0000FFFFu;
    return x;
}

// PRMT in its native 4-bit-selector form (bit 3 of a selector nibble = sign replicate; never set for valid codes)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

__device__ __forceinline__ uint32_t nibble_reverse(uint32_t x)
{
    x = __byte_perm(x, 0, 0x0123);
    return ((x & 0x0F0F0F0Fu) << 4) | ((x >> 4) & 0x0F0F0F0Fu);
}

// word i of (block >> 4*OFF nibbles), block = BW little-endian words
template <int OFF, int BW>
__device__ __forceinline__ uint32_t block_word(const uint32_t (&b)[BW], int i)
{
    constexpr int WS = OFF / 8, SH = 4 * (OFF % 8);
    const int lo_i = i + WS, hi_i = i + WS + 1;
    const uint32_t lo = lo_i < BW ? b[lo_i < BW ? lo_i : 0] : 0u;
    if (SH == 0) return lo;
    const uint32_t hi = hi_i < BW ? b[hi_i < BW ? hi_i : 0] : 0u;
    return __funnelshift_r(lo, hi, SH);
}

template <int K>
struct KmerShape {
    static constexpr int NW = (K + 7) / 8;           // nibble words per k-mer
    static constexpr int BW = (K + 7 + 7) / 8;       // nibble words per block of 8 window starts
    static constexpr int NA = 2 * NW;                // ASCII words per k-mer
    static constexpr uint32_t LAST_MASK = (K % 8) ? ((1u << (4 * (K % 8))) - 1u) : 0xFFFFFFFFu;
};

// nibble words (codes 0..3) -> ASCII words, bytes >= K zeroed
template <int K>
__device__ __forceinline__ void expand_ascii(const uint32_t (&c)[KmerShape<K>::NW], uint32_t (&a)[KmerShape<K>::NA], uint32_t POOL)
{
#pragma unroll
    for (int i = 0; i < KmerShape<K>::NW; i++) {
        uint32_t lo = prmt(POOL, 0, c[i]);
#if MG_SEL_FMA
        uint32_t hi = prmt(POOL, 0, shr_fma<16>(c[i]));
#else
        uint32_t hi = prmt(POOL, 0, c[i] >> 16);
#endif
        const int rem_lo = K - 8 * i;      // bases left for word 2i
        const int rem_hi = K - 8 * i - 4;  // bases left for word 2i+1
        if (rem_lo < 4) lo &= (1u << (8 * rem_lo)) - 1u;
        if (rem_hi <= 0) hi = 0;
        else if (rem_hi < 4) hi &= (1u << (8 * rem_hi)) - 1u;
        a[2 * i] = lo;
        a[2 * i + 1] = hi;
    }
}

#ifndef SCAN_RC_SMEM
#define SCAN_RC_SMEM 1         // keep a reverse-complement image of the tile in shared memory (see scan_kernel)
#endif
#ifndef SCAN_VOTE_GROUP
#define SCAN_VOTE_GROUP 4      // windows hashed back to back before one warp vote (1, 2, 4 or 8); see scan_group.
                               // tools/scan_microbench.cu on B200 (k=21 canonical, 4 CTAs/SM): 1 -> 187.7, 2 -> 192.4, 4 -> 193.0, 8 -> 193.6 Gbp/s
#endif

template <int K, bool CANON, int J>
__device__ __forceinline__ bool scan_window(const ScanArgs &a, const uint32_t (&b)[KmerShape<K>::BW],
                                            const uint32_t (&rb)[KmerShape<K>::BW], uint64_t tmax, uint32_t pool, u64x2 &h_out)
{
    using S = KmerShape<K>;
    uint32_t f[S::NW];
#pragma unroll
    for (int i = 0; i < S::NW; i++) {
        f[i] = block_word<J, S::BW>(b, i);
        if (i == S::NW - 1) f[i] &= S::LAST_MASK;
    }
    uint32_t c[S::NW];
    if (CANON) {
        constexpr int RO = 8 * S::BW - K - J;   // nibble offset of the reverse-complement window
        uint32_t r[S::NW];
#pragma unroll
        for (int i = 0; i < S::NW; i++) {
            r[i] = block_word<RO, S::BW>(rb, i);
            if (i == S::NW - 1) r[i] &= S::LAST_MASK;
        }
        // rc < fwd as little-endian multi-word integers  <=>  memcmp(fwd, rc, k) > 0  (Sketch.cpp:569-571):
        // borrow out of rc - fwd.  m = all-ones when the reverse complement is the canonical k-mer.
        uint32_t m;
        if (S::NW == 1) {
            asm("{\n\t.reg .u32 t;\n\tsub.cc.u32 t, %1, %2;\n\tsubc.u32 %0, 0, 0;\n\t}" : "=r"(m) : "r"(r[0]), "r"(f[0]));
        } else if (S::NW == 2) {
            asm("{\n\t.reg .u32 t;\n\tsub.cc.u32 t, %1, %2;\n\tsubc.cc.u32 t, %3, %4;\n\tsubc.u32 %0, 0, 0;\n\t}"
                : "=r"(m) : "r"(r[0]), "r"(f[0]), "r"(r[S::NW > 1 ? 1 : 0]), "r"(f[S::NW > 1 ? 1 : 0]));
        } else if (S::NW == 3) {
            asm("{\n\t.reg .u32 t;\n\tsub.cc.u32 t, %1, %2;\n\tsubc.cc.u32 t, %3, %4;\n\tsubc.cc.u32 t, %5, %6;\n\tsubc.u32 %0, 0, 0;\n\t}"
                : "=r"(m) : "r"(r[0]), "r"(f[0]), "r"(r[S::NW > 1 ? 1 : 0]), "r"(f[S::NW > 1 ? 1 : 0]), "r"(r[S::NW > 2 ? 2 : 0]), "r"(f[S::NW > 2 ? 2 : 0]));
        } else {
            asm("{\n\t.reg .u32 t;\n\tsub.cc.u32 t, %1, %2;\n\tsubc.cc.u32 t, %3, %4;\n\tsubc.cc.u32 t, %5, %6;\n\tsubc.cc.u32 t, %7, %8;\n\tsubc.u32 %0, 0, 0;\n\t}"
                : "=r"(m) : "r"(r[0]), "r"(f[0]), "r"(r[S::NW > 1 ? 1 : 0]), "r"(f[S::NW > 1 ? 1 : 0]), "r"(r[S::NW > 2 ? 2 : 0]), "r"(f[S::NW > 2 ? 2 : 0]),
                  "r"(r[S::NW > 3 ? 3 : 0]), "r"(f[S::NW > 3 ? 3 : 0]));
        }
#pragma unroll
        for (int i = 0; i < S::NW; i++) c[i] = (f[i] & ~m) | (r[i] & m);
    } else {
#pragma unroll
        for (int i = 0; i < S::NW; i++) c[i] = f[i];
    }
    uint32_t asc[S::NA];
    expand_ascii<K>(c, asc, pool);
    u64x2 h = murmur3_h1<K, S::NA>(asc, a.seed);
    if (K <= 16) h.hi = 0;        // 32-bit hashes: |{A,C,G,T}|^k <= 2^32 (reference Sketch.cpp:1136, hash.cpp:31-35)
    // coarse filter on one 32-bit word; validity and the exact threshold are checked in the slow path
    h_out = h;
    return (K <= 16) ? (h.lo <= (uint32_t)tmax) : (h.hi <= (uint32_t)(tmax >> 32));
}

// G consecutive windows (J0 .. J0+G-1) are hashed without a control-flow break in between, which lets ptxas interleave
// their independent multiply chains, then ONE warp vote decides whether anybody has a survivor.
template <int K, bool CANON, int J0, int G>
__device__ __forceinline__ void scan_group(const ScanArgs &a, const uint32_t (&b)[KmerShape<K>::BW], const uint32_t (&rb)[KmerShape<K>::BW],
                                           uint64_t tile_base, uint32_t local0, uint64_t tmax, uint32_t pool, const uint32_t *sm_tile)
{
    u64x2 h[G];
    bool pass[G];
    bool any = false;
    pass[0] = scan_window<K, CANON, J0>(a, b, rb, tmax, pool, h[0]);
    if constexpr (G > 1) pass[1] = scan_window<K, CANON, J0 + (G > 1 ? 1 : 0)>(a, b, rb, tmax, pool, h[G > 1 ? 1 : 0]);
    if constexpr (G > 2) {
        pass[2] = scan_window<K, CANON, J0 + (G > 2 ? 2 : 0)>(a, b, rb, tmax, pool, h[G > 2 ? 2 : 0]);
        pass[3] = scan_window<K, CANON, J0 + (G > 2 ? 3 : 0)>(a, b, rb, tmax, pool, h[G > 2 ? 3 : 0]);
    }
    if constexpr (G > 4) {
        pass[4] = scan_window<K, CANON, J0 + (G > 4 ? 4 : 0)>(a, b, rb, tmax, pool, h[G > 4 ? 4 : 0]);
        pass[5] = scan_window<K, CANON, J0 + (G > 4 ? 5 : 0)>(a, b, rb, tmax, pool, h[G > 4 ? 5 : 0]);
        pass[6] = scan_window<K, CANON, J0 + (G > 4 ? 6 : 0)>(a, b, rb, tmax, pool, h[G > 4 ? 6 : 0]);
        pass[7] = scan_window<K, CANON, J0 + (G > 4 ? 7 : 0)>(a, b, rb, tmax, pool, h[G > 4 ? 7 : 0]);
    }
#pragma unroll
    for (int j = 0; j < G; j++) any |= pass[j];
    if (__any_sync(0xFFFFFFFFu, any)) {
#pragma unroll
        for (int j = 0; j < G; j++)
            if (__any_sync(0xFFFFFFFFu, pass[j])) scan_emit_warp(a, pass[j], h[j].lo, h[j].hi, tile_base, local0 + J0 + j, sm_tile, K, tmax);
    }
}

// 16 ASCII bytes at stream offset `off` -> two nibble words (bytes at or past stream_len become separators)
__device__ __forceinline__ uint2 stage16(const uint4 q, uint64_t off, uint64_t stream_len, uint32_t fold)
{
    uint32_t w0 = 0x88888888u, w1 = 0x88888888u;
    if (off < stream_len) {
        w0 = ascii4_to_nibbles(q.x, fold) | (ascii4_to_nibbles(q.y, fold) << 16);
        w1 = ascii4_to_nibbles(q.z, fold) | (ascii4_to_nibbles(q.w, fold) << 16);
        if (off + 16 > stream_len) {   // ragged end
            const int keep = (int)(stream_len - off);   // 1..15
            if (keep < 8) { w0 |= 0x88888888u << (4 * keep); w1 = 0x88888888u; }
            else if (keep > 8) w1 |= 0x88888888u << (4 * (keep - 8));
            else w1 = 0x88888888u;
        }
    }
    return make_uint2(w0, w1);
}

__device__ __forceinline__ uint4 load16(const uint8_t *stream, uint64_t off, uint64_t stream_len)
{
    if (off < stream_len) return __ldg(reinterpret_cast<const uint4 *>(stream + off));
    return make_uint4(0, 0, 0, 0);
}

// 32 packed positions (2-bit codes + invalid bits) -> 4 nibble words
__device__ __forceinline__ uint32_t spread_codes8(uint32_t x)   // 8 two-bit codes (16 bits) -> 8 nibbles
{
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    return x;
}
__device__ __forceinline__ uint32_t spread_bits8(uint32_t y)    // 8 bits -> bit 3 of 8 nibbles
{
    y = (y | (y << 12)) & 0x000F000Fu;
    y = (y | (y << 6)) & 0x03030303u;
    y = (y | (y << 3)) & 0x11111111u;
    return y << 3;
}
__device__ __forceinline__ uint4 expand32(uint64_t c, uint32_t m)
{
    const uint32_t lo = (uint32_t)c, hi = (uint32_t)(c >> 32);
    uint4 r;
    r.x = spread_codes8(lo & 0xFFFFu) | spread_bits8(m & 0xFFu);
    r.y = spread_codes8(lo >> 16) | spread_bits8((m >> 8) & 0xFFu);
    r.z = spread_codes8(hi & 0xFFFFu) | spread_bits8((m >> 16) & 0xFFu);
    r.w = spread_codes8(hi >> 16) | spread_bits8(m >> 24);
    return r;
}

// Warp-private tiles: every warp owns WARP_TILE window starts at a time, staged in its own slice of shared memory,
// so the only synchronisation is __syncwarp().  The next tile's bytes are prefetched into registers before the
// current tile is hashed (global latency hidden behind ~4000 instructions of work per lane).
#ifndef SCAN_MIN_BLOCKS
#define SCAN_MIN_BLOCKS 4      // resident CTAs per SM requested from ptxas (=> 64 registers).  tools/scan_microbench.cu on B200:
                               // 1 -> 181 Gbp/s (96 regs, 2 CTAs), 3 -> 186, 4 -> 188, 5 -> 183 (spills), 6 -> 177
#endif
template <int K, bool CANON, bool PACKED>
__global__ void __launch_bounds__(SCAN_THREADS, (K > 24 && SCAN_MIN_BLOCKS > 3) ? 3 : SCAN_MIN_BLOCKS) scan_kernel(const __grid_constant__ ScanArgs a)
{
    using S = KmerShape<K>;
    __shared__ __align__(16) uint32_t sm_all[SCAN_WARPS][SCAN_WARP_WORDS];
#if SCAN_RC_SMEM
    // reverse-complement image of the tile (word i = nibble-reversed, complemented word NW-1-i), written once at staging so
    // that a block's reverse-complement words are 4 more LDS (the LSU is idle) instead of ~16 ALU instructions per block
    __shared__ __align__(16) uint32_t smr_all[CANON ? SCAN_WARPS : 1][SCAN_WARP_WORDS];
#endif
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *sm = sm_all[warp];
#if SCAN_RC_SMEM
    uint32_t *smr = smr_all[CANON ? warp : 0];
#endif
    const uint32_t fold = a.preserve_case ? 0xFFFFFFFFu : 0xDFDFDFDFu;
    uint32_t pool;   // bytes 'A','C','G','T' for PRMT; opaque to the compiler so that it stays in one register
    asm volatile("mov.u32 %0, 0x54474341;" : "=r"(pool));

    const uint64_t wt_begin = a.tile_begin * (SCAN_TILE / SCAN_WARP_TILE);
    const uint64_t wt_end = a.tile_end * (SCAN_TILE / SCAN_WARP_TILE);
    const uint64_t wt_stride = (uint64_t)gridDim.x * SCAN_WARPS;
    uint64_t wt = wt_begin + (uint64_t)blockIdx.x * SCAN_WARPS + warp;
    if (wt >= wt_end) return;

    // each lane stages vectors lane, lane+32 and one of the 4 halo vectors (64 + lane%4: eight lanes write the same
    // words with the same values -- no lane-dependent branch, the warp never splits): 16 bytes -> 2 words each
    static_assert(SCAN_WARP_VECS - 64 == 4, "halo = 4 vectors");
    const int v2 = 64 + (lane & 3);
    uint4 q0, q1, q2;                      // ASCII source: three 16-byte vectors per lane
    uint64_t pc0 = 0, pc1 = 0;             // packed source: group (lane) and one of the 2 halo groups (32 + lane%2)
    uint32_t pm0 = 0, pm1 = 0;
    const int hg = 32 + (lane & 1);
    if (PACKED) {
        const uint64_t g0 = wt * (uint64_t)(SCAN_WARP_TILE / 32);
        pc0 = __ldg(a.codes + g0 + lane); pm0 = __ldg(a.inval + g0 + lane);
        pc1 = __ldg(a.codes + g0 + hg);   pm1 = __ldg(a.inval + g0 + hg);
    } else {
        const uint64_t base = wt * (uint64_t)SCAN_WARP_TILE;
        q0 = load16(a.stream, base + 16ull * lane, a.stream_len);
        q1 = load16(a.stream, base + 16ull * (lane + 32), a.stream_len);
        q2 = load16(a.stream, base + 16ull * v2, a.stream_len);
    }
    for (; wt < wt_end; wt += wt_stride) {
        const uint64_t base = wt * (uint64_t)SCAN_WARP_TILE;
        if (PACKED) {
            const uint4 e0 = expand32(pc0, pm0), e1 = expand32(pc1, pm1);
            *reinterpret_cast<uint4 *>(sm + 4 * lane) = e0;
            *reinterpret_cast<uint4 *>(sm + 4 * hg) = e1;
#if SCAN_RC_SMEM
            if (CANON) {
                *reinterpret_cast<uint4 *>(smr + (SCAN_WARP_WORDS - 4 - 4 * lane)) =
                    make_uint4(nibble_reverse(e0.w) ^ 0x33333333u, nibble_reverse(e0.z) ^ 0x33333333u, nibble_reverse(e0.y) ^ 0x33333333u, nibble_reverse(e0.x) ^ 0x33333333u);
                *reinterpret_cast<uint4 *>(smr + (SCAN_WARP_WORDS - 4 - 4 * hg)) =
                    make_uint4(nibble_reverse(e1.w) ^ 0x33333333u, nibble_reverse(e1.z) ^ 0x33333333u, nibble_reverse(e1.y) ^ 0x33333333u, nibble_reverse(e1.x) ^ 0x33333333u);
            }
#endif
        } else {
            const uint2 s0 = stage16(q0, base + 16ull * lane, a.stream_len, fold);
            const uint2 s1 = stage16(q1, base + 16ull * (lane + 32), a.stream_len, fold);
            const uint2 s2 = stage16(q2, base + 16ull * v2, a.stream_len, fold);
            *reinterpret_cast<uint2 *>(sm + 2 * lane) = s0;
            *reinterpret_cast<uint2 *>(sm + 2 * (lane + 32)) = s1;
            *reinterpret_cast<uint2 *>(sm + 2 * v2) = s2;
#if SCAN_RC_SMEM
            if (CANON) {
                *reinterpret_cast<uint2 *>(smr + (SCAN_WARP_WORDS - 2 - 2 * lane)) = make_uint2(nibble_reverse(s0.y) ^ 0x33333333u, nibble_reverse(s0.x) ^ 0x33333333u);
                *reinterpret_cast<uint2 *>(smr + (SCAN_WARP_WORDS - 2 - 2 * (lane + 32))) = make_uint2(nibble_reverse(s1.y) ^ 0x33333333u, nibble_reverse(s1.x) ^ 0x33333333u);
                *reinterpret_cast<uint2 *>(smr + (SCAN_WARP_WORDS - 2 - 2 * v2)) = make_uint2(nibble_reverse(s2.y) ^ 0x33333333u, nibble_reverse(s2.x) ^ 0x33333333u);
            }
#endif
        }
        __syncwarp();
        const uint64_t next = wt + wt_stride;
        if (next < wt_end) {   // prefetch
            if (PACKED) {
                const uint64_t g0 = next * (uint64_t)(SCAN_WARP_TILE / 32);
                pc0 = __ldg(a.codes + g0 + lane); pm0 = __ldg(a.inval + g0 + lane);
                pc1 = __ldg(a.codes + g0 + hg);   pm1 = __ldg(a.inval + g0 + hg);
            } else {
                const uint64_t nb = next * (uint64_t)SCAN_WARP_TILE;
                q0 = load16(a.stream, nb + 16ull * lane, a.stream_len);
                q1 = load16(a.stream, nb + 16ull * (lane + 32), a.stream_len);
                q2 = load16(a.stream, nb + 16ull * v2, a.stream_len);
            }
        }
        __syncwarp();
        const uint64_t tmax = a.tile_tmax ? a.tile_tmax[base / SCAN_TILE] : a.coarse_t;
#pragma unroll 1
        for (int g = lane; g < SCAN_WARP_TILE / 8; g += 32) {
            uint32_t b[S::BW], rb[S::BW];
#pragma unroll
            for (int i = 0; i < S::BW; i++) b[i] = sm[g + i];
            if (CANON) {
#if SCAN_RC_SMEM
#pragma unroll
                for (int i = 0; i < S::BW; i++) rb[i] = smr[SCAN_WARP_WORDS - S::BW - g + i];
#else
#pragma unroll
                for (int i = 0; i < S::BW; i++) rb[i] = nibble_reverse(b[S::BW - 1 - i]) ^ 0x33333333u;
#endif
            }
            const uint32_t local0 = 8u * g;
            constexpr int G = SCAN_VOTE_GROUP;
            scan_group<K, CANON, 0, G>(a, b, rb, base, local0, tmax, pool, sm);
            if constexpr (G < 8) scan_group<K, CANON, (G < 8 ? G : 0), G>(a, b, rb, base, local0, tmax, pool, sm);
            if constexpr (G < 4) {
                scan_group<K, CANON, (G < 4 ? 2 * G : 0), G>(a, b, rb, base, local0, tmax, pool, sm);
                scan_group<K, CANON, (G < 4 ? 3 * G : 0), G>(a, b, rb, base, local0, tmax, pool, sm);
            }
            if constexpr (G < 2) {
                scan_group<K, CANON, 4, 1>(a, b, rb, base, local0, tmax, pool, sm);
                scan_group<K, CANON, 5, 1>(a, b, rb, base, local0, tmax, pool, sm);
                scan_group<K, CANON, 6, 1>(a, b, rb, base, local0, tmax, pool, sm);
                scan_group<K, CANON, 7, 1>(a, b, rb, base, local0, tmax, pool, sm);
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Byte-alphabet kernel: any alphabet (protein, custom -z), k-mers hashed as they stand (the reference makes every
// non-nucleotide alphabet non-canonical, sketchParameterSetup.cpp:79-95).  Same warp-private tiling; the tile is kept
// as bytes (alphabet bytes, 0 = outside the alphabet), a lane takes 4 consecutive window starts per block and builds each
// window with PRMT byte funnels.  Validity, exact threshold and table work stay in the shared slow path.
// ---------------------------------------------------------------------------------------------------------------
constexpr int GEN_WARP_WORDS = SCAN_WARP_TILE / 4 + 16;     // 1024 + 64 bytes

template <int K>
__global__ void __launch_bounds__(SCAN_THREADS) scan_bytes_kernel(const __grid_constant__ ScanArgs a)
{
    constexpr int NWB = (K + 3) / 4;              // words per k-mer
    constexpr int BWB = (K + 3 + 3) / 4;          // words per block of 4 window starts
    constexpr int NA = 2 * ((K + 7) / 8);         // words murmur3_h1 expects
    __shared__ __align__(16) uint32_t sm_all[SCAN_WARPS][GEN_WARP_WORDS];
    __shared__ uint8_t lut[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *sm = sm_all[warp];
    for (int i = threadIdx.x; i < 256; i += SCAN_THREADS) lut[i] = a.byte_lut[i];
    __syncthreads();

    const uint64_t wt_begin = a.tile_begin * (SCAN_TILE / SCAN_WARP_TILE), wt_end = a.tile_end * (SCAN_TILE / SCAN_WARP_TILE);
    const uint64_t wt_stride = (uint64_t)gridDim.x * SCAN_WARPS;
    for (uint64_t wt = wt_begin + (uint64_t)blockIdx.x * SCAN_WARPS + warp; wt < wt_end; wt += wt_stride) {
        const uint64_t base = wt * (uint64_t)SCAN_WARP_TILE;
        // stage 1088 bytes = 68 vectors of 16: lanes take vectors lane, lane+32, 64 + lane%4
        const int vec[3] = {lane, lane + 32, 64 + (lane & 3)};
#pragma unroll
        for (int t = 0; t < 3; t++) {
            const uint64_t off = base + 16ull * vec[t];
            uint4 q = load16(a.stream, off, a.stream_len);
            uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t o = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const uint64_t pos = off + 4 * j + b;
                    const uint32_t c = pos < a.stream_len ? lut[(w[j] >> (8 * b)) & 0xFFu] : 0u;
                    o |= c << (8 * b);
                }
                w[j] = o;
            }
            *reinterpret_cast<uint4 *>(sm + 4 * vec[t]) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        __syncwarp();
        const uint64_t tmax = a.tile_tmax ? a.tile_tmax[base / SCAN_TILE] : a.coarse_t;
#pragma unroll 1
        for (int g = lane; g < SCAN_WARP_TILE / 4; g += 32) {
            uint32_t b[BWB];
#pragma unroll
            for (int i = 0; i < BWB; i++) b[i] = sm[g + i];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t asc[NA];
#pragma unroll
                for (int i = 0; i < NA; i++) {
                    uint32_t v = 0;
                    if (i < NWB) {
                        const uint32_t lo = b[i], hi = (i + 1 < BWB) ? b[i + 1] : 0u;
                        v = j == 0 ? lo : __funnelshift_r(lo, hi, 8 * j);
                        const int rem = K - 4 * i;
                        if (rem < 4) v &= (1u << (8 * rem)) - 1u;
                    }
                    asc[i] = v;
                }
                u64x2 h = murmur3_h1<K, NA>(asc, a.seed);
                if (!a.use64) h.hi = 0;
                const bool pass = a.use64 ? (h.hi <= (uint32_t)(tmax >> 32)) : (h.lo <= (uint32_t)tmax);
                if (__any_sync(0xFFFFFFFFu, pass)) scan_emit_warp(a, pass, h.lo, h.hi, base, 4u * g + j, sm, -K, tmax);
            }
        }
        __syncwarp();
    }
}

// Host-side launcher table (defined in scan_inst_*.cu)
typedef void (*scan_launch_fn)(const ScanArgs &a, int grid, cudaStream_t stream);
typedef int (*scan_occupancy_fn)();
scan_launch_fn get_scan_launcher(int k, bool canonical, bool packed);
scan_launch_fn get_bytes_launcher(int k);

}  // namespace mashgpu
