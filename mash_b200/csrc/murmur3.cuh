// murmur3.cuh -- MurmurHash3_x64_128 (first 64 bits) for k-mers of compile-time length K <= 32,
// fed from ASCII words held in registers.  Follows the published algorithm the reference uses
// (reference MurmurHash3.cpp:255-332, fmix64 :81-90; getHash hash.cpp:10-38 keeps h1 only).
#pragma once
#include <cstdint>

namespace mashgpu {

// 64-bit values are carried as two 32-bit halves so that every operation maps onto one SASS instruction:
//   x * const  -> IMAD.WIDE.U32 + 2 IMAD          rotl -> 2 SHF (funnel)        add -> IADD3 + IADD3.X
struct u64x2 { uint32_t lo, hi; };

__device__ __forceinline__ u64x2 mul_const(u64x2 a, uint64_t c)
{
    const uint32_t clo = (uint32_t)c, chi = (uint32_t)(c >> 32);
    u64x2 r;
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %4;\n\tmov.b64 {%0, %1}, t;\n\t"
        "mad.lo.u32 %1, %2, %5, %1;\n\tmad.lo.u32 %1, %3, %4, %1;\n\t}"
        : "=r"(r.lo), "=&r"(r.hi) : "r"(a.lo), "r"(a.hi), "r"(clo), "r"(chi));
    return r;
}

template <int R>
__device__ __forceinline__ u64x2 rotl(u64x2 a)
{
    u64x2 r;
    if constexpr (R == 32) { r.lo = a.hi; r.hi = a.lo; }
    else if constexpr (R < 32) { r.hi = __funnelshift_l(a.lo, a.hi, R); r.lo = __funnelshift_l(a.hi, a.lo, R); }
    else { r.hi = __funnelshift_l(a.hi, a.lo, R - 32); r.lo = __funnelshift_l(a.lo, a.hi, R - 32); }
    return r;
}

__device__ __forceinline__ u64x2 add64(u64x2 a, u64x2 b)
{
    u64x2 r;
    asm("add.cc.u32 %0, %2, %4;\n\taddc.u32 %1, %3, %5;" : "=r"(r.lo), "=r"(r.hi) : "r"(a.lo), "r"(a.hi), "r"(b.lo), "r"(b.hi));
    return r;
}

__device__ __forceinline__ u64x2 xor64(u64x2 a, u64x2 b) { return u64x2{a.lo ^ b.lo, a.hi ^ b.hi}; }

// a * 5 + c
__device__ __forceinline__ u64x2 mul5_add(u64x2 a, uint32_t c)
{
    u64x2 r;
    asm("{\n\t.reg .u64 t, cc;\n\tcvt.u64.u32 cc, %3;\n\tmad.wide.u32 t, %2, 5, cc;\n\tmov.b64 {%0, %1}, t;\n\t}"
        : "=r"(r.lo), "=r"(r.hi) : "r"(a.lo), "r"(c));
    r.hi = a.hi * 5u + r.hi;
    return r;
}

__device__ __forceinline__ u64x2 fmix(u64x2 k)
{
    k.lo ^= k.hi >> 1;                       // k ^= k >> 33
    k = mul_const(k, 0xff51afd7ed558ccdULL);
    k.lo ^= k.hi >> 1;
    k = mul_const(k, 0xc4ceb9fe1a85ec53ULL);
    k.lo ^= k.hi >> 1;
    return k;
}

// a[i] = ASCII bytes 4i..4i+3 of the k-mer (little endian, byte 0 = first base), bytes >= K zeroed.
// NA = 2*ceil(K/8) words.  Returns h1 (the first 8 bytes of the 128-bit digest).
template <int K, int NA>
__device__ __forceinline__ u64x2 murmur3_h1(const uint32_t (&a)[NA], uint32_t seed)
{
    constexpr uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    constexpr int NB = K / 16, TAIL = K & 15;
    u64x2 h1{seed, 0u}, h2{seed, 0u};
#pragma unroll
    for (int b = 0; b < NB; b++) {
        u64x2 k1{a[4 * b], a[4 * b + 1]}, k2{a[4 * b + 2], a[4 * b + 3]};
        k1 = mul_const(k1, c1); k1 = rotl<31>(k1); k1 = mul_const(k1, c2); h1 = xor64(h1, k1);
        h1 = rotl<27>(h1); h1 = add64(h1, h2); h1 = mul5_add(h1, 0x52dce729u);
        k2 = mul_const(k2, c2); k2 = rotl<33>(k2); k2 = mul_const(k2, c1); h2 = xor64(h2, k2);
        h2 = rotl<31>(h2); h2 = add64(h2, h1); h2 = mul5_add(h2, 0x38495ab5u);
    }
    if constexpr (TAIL > 8) {
        u64x2 k2{a[4 * NB + 2], a[4 * NB + 3]};
        k2 = mul_const(k2, c2); k2 = rotl<33>(k2); k2 = mul_const(k2, c1); h2 = xor64(h2, k2);
    }
    if constexpr (TAIL > 0) {
        u64x2 k1{a[4 * NB], a[4 * NB + 1]};
        k1 = mul_const(k1, c1); k1 = rotl<31>(k1); k1 = mul_const(k1, c2); h1 = xor64(h1, k1);
    }
    h1.lo ^= (uint32_t)K; h2.lo ^= (uint32_t)K;
    h1 = add64(h1, h2); h2 = add64(h2, h1);
    h1 = fmix(h1); h2 = fmix(h2);
    h1 = add64(h1, h2);
    return h1;
}

}  // namespace mashgpu
