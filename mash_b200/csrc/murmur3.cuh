// murmur3.cuh -- MurmurHash3_x64_128 (first 64 bits) for k-mers of compile-time length K <= 32,
// fed from ASCII words held in registers.  Follows the published algorithm the reference uses
// (reference MurmurHash3.cpp:255-332, fmix64 :81-90; getHash hash.cpp:10-38 keeps h1 only).
#pragma once
#include <cstdint>

namespace mashgpu {

// 64-bit values are carried as two 32-bit halves so that every operation maps onto one SASS instruction:
//   x * const  -> IMAD.WIDE.U32 + 2 IMAD          rotl -> 2 SHF (funnel)        add -> IADD3 + IADD3.X
struct u64x2 { uint32_t lo, hi; };

// ---- pipe balancing -----------------------------------------------------------------------------------------
// The hot loop is bound by the ALU pipe (SHF/LOP3/IADD3/PRMT; ncu r01: alu 76 %, fma 26-31 % of peak).  Shifts,
// rotates and 64-bit adds can also be done by the integer multiply-add unit (IMAD / IMAD.WIDE / IMAD.HI on the FMA
// pipe) if the multiplier is a power of two -- but ptxas strength-reduces a literal power of two back into
// SHF/LEA/IADD3.  Reading the multipliers from constant memory keeps them opaque (the IMAD takes the constant-bank
// operand directly, no register needed).  Each MG_*_FMA switch moves one class of operations to the FMA pipe.
// Measured on B200 (tools/scan_microbench.cu, k=21 canonical, 2 Gbp): all off 187.5 Gbp/s; rot 170.7; shr 173.2;
// add 168.2; sel 183.4; all on 128.3 -- IMAD.WIDE/IMAD.HI cost more issue slots than the SHF/IADD3 they replace,
// so every switch defaults to 0.  Kept as documented dead ends.
#ifndef MG_ROT_FMA
#define MG_ROT_FMA 0
#endif
#ifndef MG_SHR_FMA
#define MG_SHR_FMA 0
#endif
#ifndef MG_ADD_FMA
#define MG_ADD_FMA 0
#endif
#ifndef MG_SEL_FMA
#define MG_SEL_FMA 0
#endif
// g_pow2[i] = 2^i (i = 0..31)
static __constant__ uint32_t g_pow2[32] = {
    1u << 0, 1u << 1, 1u << 2, 1u << 3, 1u << 4, 1u << 5, 1u << 6, 1u << 7, 1u << 8, 1u << 9, 1u << 10, 1u << 11, 1u << 12, 1u << 13, 1u << 14, 1u << 15,
    1u << 16, 1u << 17, 1u << 18, 1u << 19, 1u << 20, 1u << 21, 1u << 22, 1u << 23, 1u << 24, 1u << 25, 1u << 26, 1u << 27, 1u << 28, 1u << 29, 1u << 30, 1u << 31};

// x >> S on the FMA pipe: high word of x * 2^(32-S)
template <int S>
__device__ __forceinline__ uint32_t shr_fma(uint32_t x)
{
    static_assert(S >= 1 && S <= 31, "shift");
    uint32_t r;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(g_pow2[32 - S]));
    return r;
}

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
    else if constexpr (MG_ROT_FMA) {
        // rotl by R' = R mod 32 of (h:l) [halves pre-swapped when R > 32]:  l*2^R' (wide) + {h >> (32-R'), h << R'}
        const uint32_t l = R < 32 ? a.lo : a.hi, h = R < 32 ? a.hi : a.lo;
        const uint32_t c = g_pow2[R & 31];
        asm("{\n\t.reg .u64 t;\n\t.reg .u32 x, y;\n\tmul.hi.u32 x, %3, %4;\n\tmul.lo.u32 y, %3, %4;\n\tmov.b64 t, {x, y};\n\t"
            "mad.wide.u32 t, %2, %4, t;\n\tmov.b64 {%0, %1}, t;\n\t}"
            : "=r"(r.lo), "=r"(r.hi) : "r"(l), "r"(h), "r"(c));
    }
    else if constexpr (R < 32) { r.hi = __funnelshift_l(a.lo, a.hi, R); r.lo = __funnelshift_l(a.hi, a.lo, R); }
    else { r.hi = __funnelshift_l(a.hi, a.lo, R - 32); r.lo = __funnelshift_l(a.lo, a.hi, R - 32); }
    return r;
}

__device__ __forceinline__ u64x2 add64(u64x2 a, u64x2 b)
{
    u64x2 r;
#if MG_ADD_FMA
    // a.lo * 1 + b (wide) carries into the high word; then + a.hi
    asm("{\n\t.reg .u64 t;\n\tmov.b64 t, {%4, %5};\n\tmad.wide.u32 t, %2, %6, t;\n\tmov.b64 {%0, %1}, t;\n\tmad.lo.u32 %1, %3, %6, %1;\n\t}"
        : "=r"(r.lo), "=&r"(r.hi) : "r"(a.lo), "r"(a.hi), "r"(b.lo), "r"(b.hi), "r"(g_pow2[0]));
#else
    asm("add.cc.u32 %0, %2, %4;\n\taddc.u32 %1, %3, %5;" : "=r"(r.lo), "=r"(r.hi) : "r"(a.lo), "r"(a.hi), "r"(b.lo), "r"(b.hi));
#endif
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

__device__ __forceinline__ uint32_t shr1(uint32_t x)
{
#if MG_SHR_FMA
    return shr_fma<1>(x);
#else
    return x >> 1;
#endif
}

__device__ __forceinline__ u64x2 fmix(u64x2 k)
{
    k.lo ^= shr1(k.hi);                      // k ^= k >> 33
    k = mul_const(k, 0xff51afd7ed558ccdULL);
    k.lo ^= shr1(k.hi);
    k = mul_const(k, 0xc4ceb9fe1a85ec53ULL);
    k.lo ^= shr1(k.hi);
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
