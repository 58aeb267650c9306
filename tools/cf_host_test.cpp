// Host-side check of the cuckoo filter used by dist_probe_kernel (mash_b200/csrc/dist_filter.cuh): every inserted rank must be
// found (no false negatives), the random-walk insert must succeed at the tile's load, and the false-positive rate is printed.
// Usage: cf_host_test <mode> <seed>   mode 0 = 32 unrelated rows x 1000 ranks, 1 = 32 related rows, 2 = dense sequential ranks
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>
#include <vector>
#include "../mash_b200/csrc/dist_filter.cuh"
using namespace mashgpu;
int main(int argc, char **argv)
{
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const uint32_t seed = argc > 2 ? (uint32_t)atoi(argv[2]) : 1;
    std::mt19937_64 rng(seed);
    std::vector<uint32_t> tab(CF_BUCKETS, 0);
    std::set<uint32_t> members;
    std::vector<uint32_t> stream;
    const uint32_t D = mode == 2 ? 40000u : 100000000u;
    if (mode == 0) {
        for (int r = 0; r < 32; r++) for (int i = 0; i < 1000; i++) stream.push_back((uint32_t)(rng() % D));
    } else if (mode == 1) {
        std::vector<uint32_t> base;
        for (int i = 0; i < 2000; i++) base.push_back((uint32_t)(rng() % D));
        for (int r = 0; r < 32; r++) for (int i = 0; i < 1000; i++) stream.push_back((rng() % 10) ? base[rng() % 2000] : (uint32_t)(rng() % D));
    } else {
        for (uint32_t i = 0; i < 32000; i++) stream.push_back(i + (uint32_t)(rng() % 3) * 32000u % D);
    }
    int failed = 0;
    uint32_t salt = 12345;
    for (uint32_t c : stream) { members.insert(c); if (!cf_insert(tab.data(), c, salt++)) failed++; }
    int missing = 0;
    for (uint32_t c : members) if (!cf_lookup(tab.data(), c)) missing++;
    // the probe kernel's one-product form must address the same buckets and carry the same fingerprint
    int form_mismatch = 0;
    for (int i = 0; i < 200000; i++) {
        const uint32_t c = (uint32_t)rng();
        uint32_t o1, o2, f2;
        cf_offsets(c, o1, o2, f2);
        const uint32_t h = cf_hash(c), fp = cf_fp(h);
        form_mismatch += o1 != 4u * cf_bucket(h) || o2 != 4u * cf_alt(cf_bucket(h), fp) || f2 != fp * 0x00010001u ||
                         (fp & 0x4000u) != 0 || (fp & 1u) == 0 || fp > 0xFFFFu;
    }
    // reference-id side table: after marking, every (rank, reference) must be named by its slot or flagged as shared
    std::vector<uint8_t> ids(2 * CF_BUCKETS, (uint8_t)CF_ID_NONE);
    const size_t per_ref = (stream.size() + 31) / 32;
    for (size_t i = 0; i < stream.size(); i++) cf_mark_ids(tab.data(), ids.data(), stream[i], (uint32_t)(i / per_ref));      // one reference after the other
    int owner_missing = 0; uint64_t owner_multi = 0;
    for (size_t i = 0; i < stream.size(); i++) {
        bool multi = false;
        const uint32_t bits = cf_owner_bits(tab.data(), ids.data(), stream[i], &multi);
        if (!multi && !((bits >> (i / per_ref)) & 1u)) owner_missing++;
        owner_multi += multi;
    }
    uint64_t used = 0;
    for (uint32_t w : tab) used += ((w & 0xFFFF) != 0) + ((w >> 16) != 0);
    uint64_t fp = 0, probes = 0;
    for (int i = 0; i < 2000000; i++) {
        const uint32_t c = (uint32_t)(rng() % D);
        if (members.count(c)) continue;
        probes++;
        fp += cf_lookup(tab.data(), c);
    }
    printf("{\"distinct\": %zu, \"slots_used\": %llu, \"insert_failures\": %d, \"missing\": %d, \"form_mismatch\": %d, \"false_positive_rate\": %.3g, "
           "\"owner_missing\": %d, \"owner_multi_fraction\": %.4g}\n",
           members.size(), (unsigned long long)used, failed, missing, form_mismatch, probes ? (double)fp / probes : 0.0,
           owner_missing, (double)owner_multi / stream.size());
    return (missing == 0 && failed == 0 && form_mismatch == 0 && owner_missing == 0) ? 0 : 1;
}
