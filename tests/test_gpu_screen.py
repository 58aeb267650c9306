"""GPU parity tests for hot path 3 (screen), through the C ABI, against the oracle and test/ref/screen."""
import numpy as np
import pytest

from fixtures import synth_genome, mutate, fmt_g

pytestmark = pytest.mark.gpu


def run_screen(gpu, ref, ref_n, p, chunks):
    job = gpu.screen_open(ref, ref_n, p)
    try:
        for c in chunks:
            job.feed(c)
        return job.finish()
    finally:
        job.close()


def check(res, want):
    assert res["set_size"] == want["set_size"]
    assert np.array_equal(res["mixture"], want["mixture"])
    assert np.array_equal(res["shared"], want["shared"])
    assert np.array_equal(res["median"], want["median"])
    assert np.all(np.abs(res["identity"] - want["identity"]) <= 1e-12)
    po, pg = want["pvalue"], res["pvalue"]
    big = po > 1e-305
    assert np.all(np.abs(pg[big] - po[big]) <= 1e-12 * po[big])
    assert np.all(pg[~big] <= 1.0000001e-305)


def test_golden_screen(gpu, golden):
    # BASELINE config 1: `mash screen genomes.msh reads1.fastq reads2.fastq` -> test/ref/screen
    p = gpu.params(k=21, s=1000)
    ref = np.stack([golden.golden_sketch(i)[0] for i in range(3)])
    reads = [r for r in golden.reads_round_robin() if len(r) >= 21]
    chunk = b"".join(b"*" + r for r in reads)           # CommandScreen.cpp:224-262
    res = run_screen(gpu, ref, np.full(3, 1000, np.uint32), p, [chunk])
    for i, line in enumerate(golden.screen_lines):
        got = [fmt_g(res["identity"][i]), f"{res['shared'][i]}/1000", str(res["median"][i]), fmt_g(res["pvalue"][i])]
        assert got == line[:4]


@pytest.mark.parametrize("k,s", [(21, 1000), (16, 200), (11, 64)])
def test_synthetic_screen_matches_oracle(gpu, oracle, k, s):
    p = gpu.params(k=k, s=s)
    po = oracle.params(k=k)
    genomes = [synth_genome(40 + i, 150_000) for i in range(4)]
    genomes.append(mutate(genomes[0], 0.02, 7))
    ref = np.full((len(genomes) + 1, s), np.uint64(2**64 - 1)); ref_n = np.zeros(len(genomes) + 1, np.uint32)
    for i, g in enumerate(genomes):
        h, _, _ = oracle.sketch_unit([bytes(g)], po, s=s)
        ref[i, :h.size] = h; ref_n[i] = h.size
    # last reference: a short sketch (fewer than s hashes)
    h, _, _ = oracle.sketch_unit([bytes(synth_genome(999, 300))], po, s=s)
    ref[-1, :h.size] = h; ref_n[-1] = h.size
    rng = np.random.Generator(np.random.PCG64(4242))
    reads = []
    for _ in range(6000):
        g = genomes[int(rng.integers(0, 2))]
        a = int(rng.integers(0, g.size - 150))
        r = g[a:a + 150].copy()
        if rng.random() < 0.1:
            r[int(rng.integers(0, 150))] = ord("N")
        if rng.random() < 0.05:
            r = r[:int(rng.integers(1, 40))]                # shorter than k sometimes
        reads.append(bytes(r))
    kept = [r for r in reads if len(r) >= k]
    chunks = [b"".join(b"*" + r for r in kept[i:i + 1500]) for i in range(0, len(kept), 1500)]
    want = oracle.screen(ref, ref_n, chunks, po, s=s)
    check(run_screen(gpu, ref, ref_n, p, chunks), want)
    # chunking invariance (the reference flushes every 1 MiB; any split gives the same answer)
    check(run_screen(gpu, ref, ref_n, p, [b"".join(chunks)]), want)


def test_empty_stream(gpu, oracle):
    p = gpu.params(k=21, s=100)
    ref = np.arange(1, 101, dtype=np.uint64)[None, :] * np.uint64(1 << 40)
    res = run_screen(gpu, ref, np.array([100], np.uint32), p, [b"*NNNNNNNNNNNNNNNNNNNNNNNNNNNNNN"])
    assert res["set_size"] == 0 and res["shared"][0] == 0 and res["pvalue"][0] == 1.0 and res["identity"][0] == 0.0


def test_winner_take_all_matches_oracle(gpu, oracle):
    # `mash screen -w` (CommandScreen.cpp:357-407): hashes shared by several sketches are counted for the best one only
    from test_oracle_screen_winner import build
    po, ref, ref_n, lengths, chunks = build(oracle)
    p = gpu.params(k=21, s=400)
    want_plain = oracle.screen(ref, ref_n, chunks, po, s=400)
    want = oracle.screen(ref, ref_n, chunks, po, s=400, winner=True, ref_len=lengths)
    job = gpu.screen_open(ref, ref_n, p, ref_len=lengths)
    try:
        for c in chunks:
            job.feed(c)
        check(job.finish(), want_plain)            # finish is repeatable: plain first, then with the reallocation
        job.set_winner(True)
        res = job.finish()
    finally:
        job.close()
    check(res, want)
    assert not np.array_equal(want["shared"], want_plain["shared"])


def _refs_and_reads(oracle, po, s, n_reads, seed, with_tiny=True):
    genomes = [synth_genome(140 + i, 120_000) for i in range(3)]
    tiny = [synth_genome(700 + i, 150 + 37 * i) for i in range(6)] if with_tiny else []      # fewer k-mers than s: the sketch is every k-mer, hashes up to ~2^64
    allg = genomes + tiny
    ref = np.full((len(allg), s), np.uint64(2**64 - 1)); ref_n = np.zeros(len(allg), np.uint32)
    for i, g in enumerate(allg):
        h, _, _ = oracle.sketch_unit([bytes(g)], po, s=s)
        ref[i, :h.size] = h; ref_n[i] = h.size
    rng = np.random.Generator(np.random.PCG64(seed))
    reads = []
    pool = genomes[:2] + tiny[:2]
    for _ in range(n_reads):
        g = pool[int(rng.integers(0, len(pool)))]
        a = int(rng.integers(0, max(1, g.size - 150)))
        r = g[a:a + 150].copy()
        if rng.random() < 0.05:
            r[int(rng.integers(0, r.size))] = ord("N")
        reads.append(bytes(r))
    return ref, ref_n, reads


@pytest.mark.parametrize("bitmap", ["1", "0"])
def test_table_spanning_the_whole_hash_range(gpu, oracle, monkeypatch, bitmap):
    # A reference .msh with small genomes: their bottom-s reaches up to ~2^64, so EVERY k-mer of the mixture is a table
    # candidate (no "largest reference hash" shortcut).  Per-lane probe, with and without the value-indexed bitmap.
    monkeypatch.setenv("MASHGPU_SCREEN_BITMAP", bitmap)
    s = 300
    p = gpu.params(k=21, s=s)
    po = oracle.params(k=21)
    ref, ref_n, reads = _refs_and_reads(oracle, po, s, 5000, seed=11)
    assert ref[:, :][np.arange(ref.shape[0]), ref_n - 1].max() > np.uint64(2**63)       # the table really spans the range
    chunks = [b"".join(b"*" + r for r in reads[i:i + 1000]) for i in range(0, len(reads), 1000)]
    want = oracle.screen(ref, ref_n, chunks, po, s=s)
    check(run_screen(gpu, ref, ref_n, p, chunks), want)
    assert want["shared"][3] > 0 and want["shared"][5] == 0


@pytest.mark.parametrize("host_pack", ["0", "1"])
def test_many_small_host_chunks_are_joined(gpu, oracle, monkeypatch, host_pack):
    # the reference feeds 1 MiB HashInputs; smaller host chunks are joined inside the library before a kernel pass
    # (host_pack: the joined chunk goes up as ASCII, or 2-bit packed by the host threads with an invalid-position mask)
    monkeypatch.setenv("MASHGPU_SCREEN_HOST_PACK", host_pack)
    s = 200
    p = gpu.params(k=21, s=s)
    po = oracle.params(k=21)
    ref, ref_n, reads = _refs_and_reads(oracle, po, s, 3000, seed=12, with_tiny=False)
    chunks = [b"".join(b"*" + r for r in reads[i:i + 7]) for i in range(0, len(reads), 7)]       # ~1 KB each
    want = oracle.screen(ref, ref_n, chunks, po, s=s)
    check(run_screen(gpu, ref, ref_n, p, chunks), want)


@pytest.mark.parametrize("host_pack,threads", [("0", "1"), ("1", "1"), ("1", "5")])
def test_large_host_chunks_are_pipelined(gpu, oracle, monkeypatch, host_pack, threads):
    # chunks above the joining threshold go through the two-buffer pipeline (copy of chunk i+1 overlaps the kernels of chunk i);
    # with the host packer the chunk crosses PCIe as 2-bit codes + an invalid mask (lower case, N, '*' separators, a ragged tail)
    monkeypatch.setenv("MASHGPU_SCREEN_HOST_PACK", host_pack)
    monkeypatch.setenv("MASHGPU_PACK_THREADS", threads)
    s = 300
    p = gpu.params(k=21, s=s)
    po = oracle.params(k=21)
    ref, ref_n, reads = _refs_and_reads(oracle, po, s, 90_000, seed=13)
    per = 30_000                                                   # 30 000 x 151 B = 4.5 MB per chunk
    reads = [r.lower() if i % 17 == 0 else r for i, r in enumerate(reads)]        # case folding happens in the packer on that path
    chunks = [b"".join(b"*" + r for r in reads[i:i + per]) for i in range(0, len(reads), per)]
    chunks[1] = chunks[1] + b"*ACGTACGTACGTACGTACGTACGTAC"                           # a length that is not a multiple of 32 or 64
    assert min(len(c) for c in chunks) >= 4 << 20
    want = oracle.screen(ref, ref_n, chunks, po, s=s)
    job = gpu.screen_open(ref, ref_n, p)
    try:
        for c in chunks:
            buf = bytearray(c)
            job.feed(bytes(buf))
            buf[:] = b"N" * len(buf)          # the caller may reuse its buffer as soon as feed() returns
        res = job.finish()
    finally:
        job.close()
    check(res, want)
