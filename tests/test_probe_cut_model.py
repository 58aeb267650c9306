"""CPU model of dist_probe_kernel's cut-off (mash_b200/csrc/dist.cu): a query stops being looked up in a tile's filter once
`index + (lower bound of the ranks every reference of the tile has below it)` reaches the sketch size.  The rule must never
drop a (query, tile) combination whose merge would count a shared hash, and for the combinations it leaves unflagged the closed
form (common = 0, denom = min(s', |A| + |B|)) must equal the reference's merge (CommandDistance.cpp:347-385)."""
import numpy as np
import pytest

from test_pair_merge_model import reference_merge

PAD = 0xFFFFFFFF
GROUP = 128


def probe_model(refs, B, S):
    """Returns (flagged, ranks_looked_up): the group loop of the kernel with exact lookups (no filter false positives)."""
    mx = []
    for l in range(32):
        mx.append(max((int(A[32 * l]) if 32 * l < min(len(A), S) else PAD) for A in refs))
    assert all(mx[l] <= mx[l + 1] for l in range(31))
    keys = set()
    for A in refs:
        keys.update(int(x) for x in A[:S])
    nB = min(len(B), S)
    cut_at = S + 31
    base, looked = 0, 0
    while base < nB:
        grp = B[base:min(base + GROUP, nB)]
        looked += len(grp)
        if any(int(x) in keys for x in grp):
            return True, looked
        if base + GROUP > nB:
            break                                   # ragged last group: nothing behind it
        base += GROUP
        last = int(grp[-1])
        L = sum(1 for l in range(32) if mx[l] < last)
        if base + 32 * L >= cut_at:
            break
    return False, looked


def make_case(rng, S, n_refs, universe, related):
    def sketch(n):
        return np.sort(rng.choice(universe, size=n, replace=False)).astype(np.uint64)
    sizes = [S] * n_refs
    for i in range(n_refs):
        r = rng.random()
        if r < 0.15:
            sizes[i] = int(rng.integers(0, S))          # short list (genome smaller than the sketch)
        elif r < 0.25:
            sizes[i] = S + int(rng.integers(1, 50))     # rows may be longer than s' (dist -s smaller than the stored size)
    refs = [sketch(n) for n in sizes]
    nB = S if rng.random() < 0.7 else int(rng.integers(0, S + 20))
    B = sketch(nB)
    if related and len(B):
        # plant shared hashes at chosen depths of the query, into one reference
        a = int(rng.integers(0, n_refs))
        n_sh = int(rng.integers(1, 4))
        where = rng.integers(0, len(B), size=n_sh)
        A = set(int(x) for x in refs[a])
        A.update(int(B[w]) for w in where)
        refs[a] = np.array(sorted(A), dtype=np.uint64)
    return refs, B


@pytest.mark.parametrize("S", [1, 31, 100, 400, 1000, 1035])
def test_cut_never_drops_a_counted_shared_hash(S):
    rng = np.random.default_rng(1234 + S)
    cut_happened = 0
    for trial in range(120):
        universe = int(rng.choice([4 * S + 64, 40 * S + 64, 1000 * S + 64]))
        refs, B = make_case(rng, S, int(rng.integers(1, 33)), universe, related=trial % 2 == 0)
        flagged, looked = probe_model(refs, B, S)
        cut_happened += looked < min(len(B), S)
        if flagged:
            continue                                 # the combination is merged exactly, whatever the merge finds
        for A in refs:
            common, denom = reference_merge([int(x) for x in A], [int(x) for x in B], S)
            assert common == 0
            assert denom == min(S, len(A) + len(B))
    if S >= 400:
        assert cut_happened > 10                     # the rule does something


def test_similar_size_unrelated_sketches_stop_near_the_middle():
    rng = np.random.default_rng(7)
    S = 1000
    def sketch():
        return np.unique(rng.integers(0, 2**62, size=S + 8))[:S]      # (ranks of unrelated sketches: no collisions to speak of)
    refs = [sketch() for _ in range(32)]
    looked = []
    for _ in range(20):
        B = sketch()
        flagged, n = probe_model(refs, B, S)
        assert not flagged
        looked.append(n)
    assert max(looked) <= 768 and np.mean(looked) <= 700
