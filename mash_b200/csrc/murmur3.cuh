// murmur3.cuh -- MurmurHash3_x64_128 (first 64 bits) for k-mers of compile-time length K <= 32,
// fed from ASCII words held in registers.  Follows the published algorithm the reference uses
// (reference MurmurHash3.cpp:255-332, fmix64 :81-90; getHash hash.cpp:10-38 keeps h1 only).
#pragma once
#include <cstdint>

namespace mashgpu {

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

__device__ __forceinline__ uint64_t fmix64(uint64_t k)
{
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

__device__ __forceinline__ uint64_t pack64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// a[i] = ASCII bytes 4i..4i+3 of the k-mer (little endian, byte 0 = first base), bytes >= K zeroed.
// NA = 2*ceil(K/8) words.
template <int K, int NA>
__device__ __forceinline__ uint64_t murmur3_h1(const uint32_t (&a)[NA], uint32_t seed)
{
    constexpr uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    constexpr int NB = K / 16, TAIL = K & 15;
    uint64_t h1 = seed, h2 = seed;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        uint64_t k1 = pack64(a[4 * b], a[4 * b + 1]);
        uint64_t k2 = pack64(a[4 * b + 2], a[4 * b + 3]);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    if constexpr (TAIL > 8) {
        uint64_t k2 = pack64(a[4 * NB + 2], a[4 * NB + 3]);
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
    }
    if constexpr (TAIL > 0) {
        uint64_t k1 = pack64(a[4 * NB], a[4 * NB + 1]);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    }
    h1 ^= (uint64_t)K; h2 ^= (uint64_t)K;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2;
    return h1;
}

}  // namespace mashgpu
