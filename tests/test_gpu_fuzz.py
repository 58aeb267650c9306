"""Seeded randomised parity sweep: sketch -> dist -> screen through the C ABI vs the oracle, over random k, s, seeds,
strand modes, case modes and ragged inputs (the reference's own tests only exercise k=21, s=1000)."""
import numpy as np
import pytest

from fixtures import synth_genome, mutate

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", range(24))
def test_random_pipeline(gpu, oracle, case):
    rng = np.random.Generator(np.random.PCG64(9000 + case))
    k = int(rng.integers(1, 33))
    s = int(rng.choice([1, 7, 64, 333, 1000, 1035]))
    seed = int(rng.integers(0, 2**32))
    noncanonical = bool(rng.integers(0, 2))
    preserve_case = bool(rng.integers(0, 2))
    p = gpu.params(k=k, s=s, seed=seed, noncanonical=noncanonical, preserve_case=preserve_case)
    po = oracle.params(k=k, seed=seed, noncanonical=noncanonical, preserve_case=preserve_case)
    base = synth_genome(case, int(rng.integers(2_000, 120_000)), n_runs=int(rng.integers(0, 6)), lower_frac=float(rng.choice([0, 0.05, 0.5])))
    units = []
    for u in range(int(rng.integers(2, 7))):
        g = mutate(base, float(rng.choice([0, 0.002, 0.02, 0.2])), 100 * case + u)
        n_rec = int(rng.integers(1, 4))
        cuts = np.sort(rng.integers(0, g.size, n_rec - 1))
        recs = [bytes(x) for x in np.split(g, cuts)]
        if rng.random() < 0.3:
            recs.append(b"")
        if rng.random() < 0.3:
            recs.append(bytes(rng.integers(0, 256, int(rng.integers(1, 200)), dtype=np.uint8)))
        units.append(recs)
    recs, uor = [], []
    for u, rs in enumerate(units):
        recs += rs; uor += [u] * len(rs)
    h, n, length, c = gpu.sketch(recs, p, unit_of_record=uor, n_units=len(units), counts=True)
    for u, rs in enumerate(units):
        oh, oc, olen = oracle.sketch_unit(rs, po, s=s, counts=True)
        assert length[u] == olen and n[u] == oh.size
        assert np.array_equal(h[u, :n[u]], oh) and np.array_equal(c[u, :n[u]], oc), (case, k, s, u)
    # dist all-vs-all on what was just sketched
    ks = 4.0 ** k
    res = gpu.dist(h, n, np.maximum(length, 1), sketch_size=s, k=k, kmer_space=ks)
    want = oracle.compare_all(h, n, np.maximum(length, 1), h, n, np.maximum(length, 1), s, k, ks)
    assert np.array_equal(res["numer"], want["numer"]) and np.array_equal(res["denom"], want["denom"])
    assert np.all(np.abs(res["distance"] - want["distance"]) <= 1e-12)
    big = want["pvalue"] > 1e-305
    assert np.all(np.abs(res["pvalue"][big] - want["pvalue"][big]) <= 1e-12 * want["pvalue"][big])
    # screen: reads drawn from the first unit
    if s <= 1000:
        src = np.frombuffer(b"".join(units[0]), np.uint8)
        if src.size > 200:
            reads = [bytes(src[a:a + 100]) for a in rng.integers(0, src.size - 100, 300)]
            chunk = b"".join(b"*" + r for r in reads if len(r) >= k)
            job = gpu.screen_open(h, n, p)
            job.feed(chunk[: len(chunk) // 2]); job.feed(chunk[len(chunk) // 2:])
            got = job.finish(); job.close()
            # splitting a chunk in the middle of a read breaks that read's k-mers at the cut: feed the same split to the oracle
            ws = oracle.screen(h, n, [chunk[: len(chunk) // 2], chunk[len(chunk) // 2:]], po, s=s)
            assert np.array_equal(got["shared"], ws["shared"]) and np.array_equal(got["median"], ws["median"])
            assert got["set_size"] == ws["set_size"] and np.array_equal(got["mixture"], ws["mixture"])
