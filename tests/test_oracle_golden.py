"""Pin the oracle (oracle/mash_oracle.c) against every golden vector the reference's own tests hold
for the hot paths (reference Makefile.in:94-115, test/ref/*; doc/sphinx/tutorials.rst:24,56-57),
murmur known answers from the reference's object code, and mpmath for the binomial tail."""
import numpy as np
import pytest

from fixtures import fmt_g

K, S, SEED = 21, 1000, 42


@pytest.fixture(scope="module")
def genome_sketches(oracle, golden):
    p = oracle.params(k=K, seed=SEED)
    out = []
    for fname, recs in golden.genomes:
        h, _, length = oracle.sketch_unit([r[2] for r in recs], p, s=S)
        out.append((h, length))
    return out


@pytest.fixture(scope="module")
def reads_sketch(oracle, golden):
    p = oracle.params(k=K, seed=SEED)
    h, _, length = oracle.sketch_unit(golden.reads_round_robin(), p, s=S, reads=True)
    return h, length


def test_murmur_known_answers(oracle, golden):
    for c in golden.murmur_kat:
        assert oracle.get_hash(c["kmer"].encode(), c["seed"], c["use64"]) == int(c["hash"]), c
    # SURVEY.md 8(c) spot values
    assert oracle.get_hash(b"A" * 21, 42, True) == 18154334747705351023
    assert oracle.get_hash(b"ACGT" * 4, 42, False) == 2886031495


def test_use64_rule(oracle):
    # Sketch.cpp:1136: use64 <=> alphabetSize^k > 2^32
    assert oracle.params(k=16).use64 == 0
    assert oracle.params(k=17).use64 == 1
    assert oracle.params(k=7, alphabet="ACDEFGHIKLMNPQRSTVWY").use64 == 0
    assert oracle.params(k=8, alphabet="ACDEFGHIKLMNPQRSTVWY").use64 == 1


def test_genome_sketches_match_golden_json(genome_sketches, golden):
    # `mash sketch genome1.fna genome2.fna genome3.fna` -> test/ref/genomes.json
    for i, (h, length) in enumerate(genome_sketches):
        gh, glen, _, _ = golden.golden_sketch(i)
        assert length == glen
        assert h.size == 1000 and np.array_equal(h, gh)


def test_reads_sketch_matches_golden_json(reads_sketch, golden):
    # `mash sketch -r -I reads reads1.fastq reads2.fastq` -> test/ref/reads.json (length = estimateSetSize)
    h, length = reads_sketch
    gh, glen = golden.golden_reads_sketch()
    assert np.array_equal(h, gh)
    assert length == glen == 502359


def test_dist_matches_golden(oracle, genome_sketches, reads_sketch, golden):
    # `mash dist genomes.msh reads.msh` -> test/ref/genomes.dist
    p = oracle.params(k=K, seed=SEED)
    ks = oracle.kmer_space(p)
    qh, qlen = reads_sketch
    for (rh, rlen), line in zip(genome_sketches, golden.dist_lines):
        o = oracle.compare_sketches(rh, rlen, qh, qlen, S, K, ks)
        assert o.pass_ == 1
        assert [fmt_g(o.distance), fmt_g(o.pvalue), f"{o.numer}/{o.denom}"] == line[2:5]


def test_tutorial_known_answers(oracle, genome_sketches):
    # doc/sphinx/tutorials.rst:24,56-57
    p = oracle.params(k=K, seed=SEED)
    ks = oracle.kmer_space(p)
    (h1, l1), (h2, l2), (h3, l3) = genome_sketches
    o = oracle.compare_sketches(h1, l1, h2, l2, S, K, ks)
    assert (fmt_g(o.distance), fmt_g(o.pvalue), o.numer, o.denom) == ("0.0222766", "0", 456, 1000)
    o = oracle.compare_sketches(h1, l1, h3, l3, S, K, ks)
    assert (fmt_g(o.distance), fmt_g(o.pvalue), o.numer, o.denom) == ("0", "0", 1000, 1000)


def test_screen_matches_golden(oracle, genome_sketches, golden):
    # `mash screen genomes.msh reads1.fastq reads2.fastq` -> test/ref/screen
    p = oracle.params(k=K, seed=SEED)
    ref = np.stack([h for h, _ in genome_sketches])
    ref_n = np.full(3, 1000, np.uint32)
    # CommandScreen.cpp:224-262: '*' + read, reads of length >= k only
    chunk = b"".join(b"*" + r for r in golden.reads_round_robin() if len(r) >= K)
    res = oracle.screen(ref, ref_n, [chunk], p, s=S)
    for i, line in enumerate(golden.screen_lines):
        got = [fmt_g(res["identity"][i]), f"{res['shared'][i]}/1000", str(res["median"][i]), fmt_g(res["pvalue"][i])]
        assert got == line[:4]


def test_screen_chunking_invariant(oracle, genome_sketches, golden):
    # splitting the stream into several '*'-joined chunks (1 MiB in the reference) changes nothing
    p = oracle.params(k=K, seed=SEED)
    ref = np.stack([h for h, _ in genome_sketches])
    ref_n = np.full(3, 1000, np.uint32)
    reads = [r for r in golden.reads_round_robin() if len(r) >= K]
    one = oracle.screen(ref, ref_n, [b"".join(b"*" + r for r in reads)], p)
    many = oracle.screen(ref, ref_n, [b"".join(b"*" + r for r in reads[i:i + 300]) for i in range(0, len(reads), 300)], p)
    assert np.array_equal(one["counts"], many["counts"]) and one["set_size"] == many["set_size"]
    assert np.array_equal(one["mixture"], many["mixture"])


def test_binomial_tail_vs_mpmath(oracle, golden):
    worst = 0.0
    for c in golden.pvalue_cases:
        truth = float(c["p"])
        got = oracle.binomial_upper_tail(c["x"], float(c["r"]), c["n"])
        if truth == 0.0 or truth < 1e-300:
            assert got <= 1e-300
            continue
        rel = abs(got - truth) / truth
        worst = max(worst, rel)
        assert rel <= 1e-12, (c, got)
    assert worst < 1e-13


def test_compare_edge_cases(oracle):
    ks = 4.0 ** 21
    a = np.array([1, 5, 9, 12], np.uint64)
    b = np.array([2, 5, 7, 12, 20, 30], np.uint64)
    # union: 1 2 5 7 9 12 | 20 30; sketch_size 6 -> common 2 (5,12) of 6
    o = oracle.compare_sketches(a, 1000, b, 1000, 6, 21, ks)
    assert (o.numer, o.denom) == (2, 6)
    # list runs out before s: denom = min(s, denom + remaining)  (CommandDistance.cpp:367-385)
    o = oracle.compare_sketches(a, 1000, b, 1000, 100, 21, ks)
    assert (o.numer, o.denom) == (2, 8)
    o = oracle.compare_sketches(a, 1000, b, 1000, 7, 21, ks)
    assert (o.numer, o.denom) == (2, 7)
    # empty vs empty: denom 0, common==denom -> distance 0, p 1
    e = np.array([], np.uint64)
    o = oracle.compare_sketches(e, 1000, e, 1000, 10, 21, ks)
    assert (o.numer, o.denom, o.distance, o.pvalue, o.pass_) == (0, 0, 0.0, 1.0, 1)
    # distance filter leaves fields unset (CommandDistance.cpp:409-412)
    o = oracle.compare_sketches(a, 1000, np.array([3, 4], np.uint64), 1000, 6, 21, ks, max_distance=0.5)
    assert o.pass_ == 0 and o.filled == 0
