"""CPU model of dist_pair_kernel's algorithm (mash_b200/csrc/dist.cu): the zooming union count with odd run lengths and reads past a
lane's ranges ("an exhausted side always holds an element that cannot win"), lane for lane as the kernel does it, against the
reference's sequential merge (CommandDistance.cpp:347-385).  Guards the reasoning the kernel's missing bounds checks rest on:
ragged and empty rows, ties at run boundaries, rows longer than the sketch size, every sketch size class (1 ... 10 000)."""
import numpy as np
import pytest

PAD = 0xFFFFFFFF


def reference_merge(A, B, S):
    """compareSketches' loop: common, denom (CommandDistance.cpp:347-385)"""
    i = j = common = denom = 0
    while denom < S and i < len(A) and j < len(B):
        if A[i] < B[j]:
            i += 1
        elif B[j] < A[i]:
            j += 1
        else:
            i += 1; j += 1; common += 1
        denom += 1
    if denom < S:
        if i < len(A):
            denom += len(A) - i
        if j < len(B):
            denom += len(B) - j
        denom = min(denom, S)
    return common, denom


def kernel_model(rowA, rowB, S):
    """One warp of dist_pair_kernel.  rowA / rowB: ascending distinct ranks (any length); returns (common, taken)."""
    nA, nB = min(len(rowA), S), min(len(rowB), S)
    sA = list(rowA[:nA]) + [PAD]            # staged row + sentinel (the kernel overwrites position nA)
    sB = list(rowB[:nB]) + [PAD]
    run0 = ((S + 31) // 32) | 1
    i0, i1, j0, j1 = 0, nA, 0, nB
    need, common, taken = S, 0, 0
    first = True
    for _round in range(64):
        lenB = j1 - j0
        run = run0 if first else (lenB + 31) // 32
        if run > 1:
            run |= 1
        first = False
        lanes = []
        for lane in range(32):
            jb = min(j0 + lane * run, j1)
            je = min(jb + run, j1)
            if lane == 0:
                ib = i0
            elif jb >= j1:
                ib = i1
            else:
                v = sB[jb]
                lo, hi = i0, i1
                while lo < hi:
                    mid = (lo + hi) >> 1
                    if sA[mid] < v:
                        lo = mid + 1
                    else:
                        hi = mid
                ib = lo
            lanes.append([jb, je, ib])
        us, ts, ies = [], [], []
        for lane in range(32):
            jb, je, ib = lanes[lane]
            ie = lanes[lane + 1][2] if lane < 31 else i1
            ies.append(ie)
            u = t = 0
            if ib < ie or jb < je:
                pa, pb = ib, jb
                av, bv = sA[pa], sB[pb]                       # natural reads: may be past the lane's ranges
                while True:
                    adv_a, adv_b = av <= bv, bv <= av
                    if adv_a:
                        pa += 1
                    if adv_b:
                        pb += 1
                    if adv_a and adv_b:
                        t += 1
                    if adv_a:
                        av = sA[pa]
                    if adv_b:
                        bv = sB[pb]
                    u += 1
                    if not (pa < ie or pb < je):
                        break
                assert pa == ie and pb == je, "a lane ran past its ranges: the no-bounds-check argument is broken"
            us.append(u); ts.append(t)
        U = np.cumsum(us); T = np.cumsum(ts)
        if U[-1] <= need:
            common += int(T[-1]); taken += int(U[-1])
            return common, taken
        L = int(np.argmax(U >= need))
        Uprev, Tprev = int(U[L] - us[L]), int(T[L] - ts[L])
        common += Tprev; taken += Uprev; need -= Uprev
        i0, i1, j0, j1 = lanes[L][2], ies[L], lanes[L][0], lanes[L][1]
        if j1 - j0 <= 1:
            if j1 > j0:
                b = sB[j0]
                k = sum(1 for i in range(i0, i1) if sA[i] < b)
                tie = i0 + k < i1 and sA[i0 + k] == b
                if tie and need >= k + 1:
                    common += 1
            taken += need
            return common, taken
    raise AssertionError("zoom did not terminate")


def rows(rng, n, universe):
    return np.sort(rng.choice(universe, n, replace=False)).astype(np.int64)


@pytest.mark.parametrize("S", [1, 2, 31, 32, 33, 100, 1000, 1035, 1036, 2500, 10000])
def test_zooming_union_count_equals_sequential_merge(S):
    rng = np.random.Generator(np.random.PCG64(S))
    cases = 0
    for trial in range(60):
        kind = trial % 6
        nA = int(rng.integers(0, S + 40)) if kind in (0, 1) else S
        nB = int(rng.integers(0, S + 40)) if kind in (0, 2) else S
        if kind == 5:
            nA, nB = (0, S) if trial % 2 else (S, 0)
        universe = max(4 * (nA + nB + 2), 64) if kind != 3 else max(nA, nB) + 5          # kind 3: dense universe -> many ties
        A = rows(rng, nA, universe)
        B = rows(rng, nB, universe)
        if kind == 4 and nA and nB:                                                      # related rows: B = A with some entries replaced
            keep = rng.random(nA) < 0.7
            fresh = rows(rng, nA, 8 * universe) + universe
            B = np.unique(np.where(keep, A, fresh))[:nB]
        want = reference_merge(list(A), list(B), S)
        got = kernel_model(list(A), list(B), S)
        assert got == want, (S, trial, kind, nA, nB, got, want)
        cases += 1
    assert cases == 60
