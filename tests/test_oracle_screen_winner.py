"""CPU checks of the oracle's restatement of `mash screen -w` (CommandScreen.cpp:357-407; oracle/mash_oracle.c
mo_screen_finish_winner): every reference hash seen in the mixture is counted for exactly one sketch -- the best
(identity estimate, then genome length) among those containing it."""
import numpy as np

from fixtures import synth_genome, mutate


def build(oracle, k=21, s=400):
    po = oracle.params(k=k)
    a = synth_genome(1, 120_000)
    genomes = [a, mutate(a, 0.01, 2), mutate(a, 0.05, 3), synth_genome(4, 120_000), a[:100_000].copy()]
    lengths = np.array([g.size + 10 * i for i, g in enumerate(genomes)], np.uint64)       # all different: no full ties
    ref = np.full((len(genomes), s), np.uint64(2 ** 64 - 1)); ref_n = np.zeros(len(genomes), np.uint32)
    for i, g in enumerate(genomes):
        h, _, _ = oracle.sketch_unit([bytes(g)], po, s=s)
        ref[i, :h.size] = h; ref_n[i] = h.size
    rng = np.random.Generator(np.random.PCG64(11))
    reads = []
    for _ in range(6000):
        g = genomes[int(rng.integers(0, 2))]               # the mixture holds genomes 0 and 1 only
        p0 = int(rng.integers(0, g.size - 150))
        reads.append(bytes(g[p0:p0 + 150]))
    chunks = [b"".join(b"*" + r for r in reads[i:i + 2000]) for i in range(0, len(reads), 2000)]
    return po, ref, ref_n, lengths, chunks


def test_winner_counts_every_seen_hash_once(oracle):
    po, ref, ref_n, lengths, chunks = build(oracle)
    plain = oracle.screen(ref, ref_n, chunks, po, s=400)
    win = oracle.screen(ref, ref_n, chunks, po, s=400, winner=True, ref_len=lengths)
    seen = int((plain["counts"] >= 1).sum())
    assert seen > 0 and int(win["shared"].sum()) == seen                 # each seen hash goes to exactly one sketch
    assert int(plain["shared"].sum()) > seen                              # the related genomes shared hashes before
    assert np.all(win["shared"] <= plain["shared"])
    best = int(np.argmax(plain["identity"]))
    assert win["shared"][best] == plain["shared"][best]                   # the best sketch keeps everything it had
    assert win["set_size"] == plain["set_size"] and np.array_equal(win["mixture"], plain["mixture"])
    # identity / p-value follow from the new counts (CommandScreen.cpp:420-428)
    for i in range(ref.shape[0]):
        sh, n = int(win["shared"][i]), int(ref_n[i])
        want = 1.0 if sh == n else (0.0 if sh == 0 else (sh / n) ** (1.0 / 21))
        assert abs(win["identity"][i] - want) <= 1e-15
        assert (win["pvalue"][i] == 1.0) == (sh == 0)


def test_winner_is_a_no_op_without_overlap(oracle):
    po = oracle.params(k=21)
    genomes = [synth_genome(20 + i, 60_000) for i in range(3)]
    ref = np.full((3, 200), np.uint64(2 ** 64 - 1)); ref_n = np.zeros(3, np.uint32)
    for i, g in enumerate(genomes):
        h, _, _ = oracle.sketch_unit([bytes(g)], po, s=200)
        ref[i, :h.size] = h; ref_n[i] = h.size
    chunk = b"".join(b"*" + bytes(g[a:a + 150]) for g in genomes[:2] for a in range(0, 59_000, 75))
    plain = oracle.screen(ref, ref_n, [chunk], po, s=200)
    win = oracle.screen(ref, ref_n, [chunk], po, s=200, winner=True, ref_len=np.array([3, 2, 1], np.uint64))
    for key in ("shared", "median", "identity", "pvalue"):
        assert np.array_equal(plain[key], win[key])
