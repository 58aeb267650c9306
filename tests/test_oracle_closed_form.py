"""CPU check of the two identities the dist tile prefilter rests on (mash_b200/csrc/dist.cu), against
the oracle's restatement of compareSketches (reference CommandDistance.cpp:336-425):
  1. a pair whose lists share no hash among their first s' entries has numer 0, denom min(s', |A| + |B|) (lists cut at
     s'), distance 1 (0 when denom is 0) and p-value 1;
  2. hashes at index >= s' of a list never take part in the merge (so the filter may ignore them)."""
import numpy as np


def disjoint_sets(rng, n, s, ragged=True):
    """n sketches over disjoint value ranges (so no two share a hash), optionally ragged / empty."""
    H = np.full((n, s), np.uint64(0xFFFFFFFFFFFFFFFF), dtype=np.uint64)
    N = np.zeros(n, np.uint32)
    for g in range(n):
        m = int(rng.integers(0, s + 1)) if (ragged and g % 3 == 0) else s
        v = np.sort(rng.choice(1 << 20, m, replace=False).astype(np.uint64)) * np.uint64(n) + np.uint64(g)   # residue g mod n
        H[g, :m] = v
        N[g] = m
    L = rng.integers(1_000_000, 6_000_000, n).astype(np.uint64)
    return H, N, L


def test_closed_form_of_pairs_without_shared_hashes(oracle):
    rng = np.random.Generator(np.random.PCG64(5))
    for s, s_cmp in ((64, 64), (64, 20), (7, 7), (1, 1)):
        H, N, L = disjoint_sets(rng, 24, s)
        for md, mp in ((1.0, 1.0), (0.5, 1.0), (1.0, 0.5), (-1.0, -1.0)):
            got = oracle.compare_all(H, N, L, H, N, L, s_cmp, 21, 4.0 ** 21, max_distance=md, max_pvalue=mp)
            for q in range(24):
                for r in range(24):
                    if q == r:
                        continue
                    g = got[q, r]
                    denom = min(s_cmp, min(int(N[q]), s_cmp) + min(int(N[r]), s_cmp))
                    dist = 0.0 if denom == 0 else 1.0
                    if md >= 0 and dist > md:                      # CommandDistance.cpp:409-412: returns before filling
                        assert not g["pass"]
                        continue
                    assert g["numer"] == 0 and g["denom"] == denom and g["distance"] == dist and g["pvalue"] == 1.0
                    assert bool(g["pass"]) == (not (mp >= 0 and 1.0 > mp))


def test_entries_past_the_sketch_size_never_matter(oracle):
    rng = np.random.Generator(np.random.PCG64(9))
    s_list, s_cmp = 40, 25
    base = np.sort(rng.choice(1 << 16, 3 * s_list, replace=False).astype(np.uint64))
    H = np.stack([np.sort(rng.choice(base, s_list, replace=False)) for _ in range(12)])
    N = np.full(12, s_list, np.uint32)
    L = rng.integers(1_000_000, 6_000_000, 12).astype(np.uint64)
    full = oracle.compare_all(H, N, L, H, N, L, s_cmp, 21, 4.0 ** 21)
    cut = H.copy()
    cut[:, s_cmp:] = np.uint64(0xFFFFFFFFFFFFFFFF)
    Ncut = np.full(12, s_cmp, np.uint32)
    short = oracle.compare_all(cut, Ncut, L, cut, Ncut, L, s_cmp, 21, 4.0 ** 21)
    for key in ("numer", "denom", "distance", "pvalue", "pass"):
        assert np.array_equal(full[key], short[key]), key
