// binom.cuh -- P[Binomial(n, r) >= x] in double precision on the device.
//
// This is the quantity the reference obtains from gsl_cdf_binomial_Q(x-1, r, n) / Boost.Math
// (reference CommandDistance.cpp:444-446, CommandScreen.cpp:611-613).  Direct summation of the pmf from x
// towards the tail with geometrically decreasing terms; the first term C(n,x) r^x (1-r)^(n-x) is a running
// product with power-of-two rescaling (no lgamma), which keeps the relative error ~1e-14 for n up to 1e4
// (SURVEY.md appendix C).  Below the mean the complementary lower tail is summed instead.
#pragma once
#include <cstdint>

namespace mashgpu {

// C(n, x) r^x (1-r)^(n-x) as mant * 2^expo (mant may be far from 1; only the product is meaningful)
__device__ __forceinline__ void binom_pmf_scaled(uint64_t x, double r, uint64_t n, double &mant, int &expo)
{
    double m = 1.0;
    int e = 0;
    const double base = (double)(n - x);
    for (uint64_t i = 1; i <= x; i++) {
        m *= ((base + (double)i) / (double)i) * r;
        if (m < 0x1p-500) { m *= 0x1p500; e -= 500; }
        else if (m > 0x1p500) { m *= 0x1p-500; e += 500; }
    }
    // (1-r)^(n-x) = 2^(y), y = (n-x) log2(1-r)
    const double y = base * log1p(-r) * 1.4426950408889634074;   // 1/ln 2
    const double yi = floor(y);
    m *= exp2(y - yi);
    // yi can be hugely negative (result underflows to 0): clamp so the int conversion is safe
    e += (int)fmax(yi, -100000.0);
    mant = m;
    expo = e;
}

// Table of C(n0, x), x = 0..n0, as mant * 2^expo, built on the host in long double (dist.cu: n0 = the job's sketch size,
// which is the denominator of almost every pair).  With it the first term costs a 2 log2(x)-step power instead of x
// multiply-divide steps -- the p-value of a pair with many shared hashes used to cost more than its merge.
struct BinomTable { const double *mant; const int *expo; uint64_t n0; };

// r^x as mant * 2^expo by binary exponentiation on frexp-normalised factors (relative error <= ~x ulp, as the running product)
__device__ __forceinline__ void pow_scaled(double r, uint64_t x, double &mant, int &expo)
{
    int er;
    double base = frexp(r, &er);        // r = base * 2^er, base in [0.5, 1)
    int ebase = 0, eacc = 0, k;
    double acc = 1.0;
    const long long e_lin = (long long)er * (long long)x;
    while (x) {
        if (x & 1) { acc = frexp(acc * base, &k); eacc += ebase + k; }
        x >>= 1;
        if (x) { base = frexp(base * base, &k); ebase = 2 * ebase + k; }
    }
    const long long e = e_lin + eacc;
    mant = acc;
    expo = e < -1000000 ? -1000000 : (int)e;
}

// same quantity as binom_pmf_scaled, C(n,x) taken from the table (requires n == tab.n0)
__device__ __forceinline__ void binom_pmf_scaled_tab(uint64_t x, double r, uint64_t n, const BinomTable &tab, double &mant, int &expo)
{
    double pm; int pe;
    pow_scaled(r, x, pm, pe);
    double m = tab.mant[x] * pm;
    int e = tab.expo[x] + pe;
    const double base = (double)(n - x);
    const double y = base * log1p(-r) * 1.4426950408889634074;   // 1/ln 2
    const double yi = floor(y);
    m *= exp2(y - yi);
    e += (int)fmax(yi, -100000.0);
    mant = m;
    expo = e;
}

__device__ __forceinline__ double binomial_upper_tail(uint64_t x, double r, uint64_t n, const BinomTable *tab = nullptr)
{
    if (x == 0) return 1.0;
    if (x > n) return 0.0;
    if (!(r > 0.0)) return 0.0;
    if (r >= 1.0) return 1.0;
    const double odds = r / (1.0 - r);
    if ((double)x >= ((double)n + 1.0) * r) {
        double m; int e;
        if (tab && tab->mant && n == tab->n0) binom_pmf_scaled_tab(x, r, n, *tab, m, e);
        else binom_pmf_scaled(x, r, n, m, e);
        double sum = 1.0, t = 1.0;
        for (uint64_t i = x; i < n; i++) {
            t *= ((double)(n - i) / (double)(i + 1)) * odds;
            sum += t;
            if (t < sum * 1e-18) break;
        }
        return ldexp(m * sum, e < -100000 ? -100000 : e);
    } else {
        const uint64_t xm = x - 1;
        double m; int e;
        binom_pmf_scaled(xm, r, n, m, e);
        double sum = 1.0, t = 1.0;
        for (uint64_t i = xm; i > 0; i--) {
            t *= ((double)i / (double)(n - i + 1)) / odds;
            sum += t;
            if (t < sum * 1e-18) break;
        }
        return 1.0 - ldexp(m * sum, e < -100000 ? -100000 : e);
    }
}

// pValue, reference CommandDistance.cpp:427-448
__device__ __forceinline__ double mash_pvalue(uint64_t x, uint64_t len_ref, uint64_t len_qry, double kmer_space, uint64_t sketch_size,
                                              const BinomTable *tab = nullptr)
{
    if (x == 0) return 1.0;
    const double pX = 1. / (1. + kmer_space / (double)len_ref);
    const double pY = 1. / (1. + kmer_space / (double)len_qry);
    const double r = pX * pY / (pX + pY - pX * pY);
    return binomial_upper_tail(x, r, sketch_size, tab);
}

}  // namespace mashgpu
