"""CPU model of mashgpu_sketch_reads' `-c` algorithm (mash_b200/csrc/sketch.cu): exact heap tops at a few prefixes bound the
k-mers that can pass MinHashHeap's gate afterwards; only those are collected as events and replayed in stream order.  The
replay here is the oracle's own heap (mo_heap_m_*), so what this checks is the band argument: that no k-mer the full stream
would have fed through the gate is missing from the event list -- the result must equal the oracle run on the whole stream."""
import ctypes as C

import numpy as np
import pytest

from test_oracle_vs_ref import _read_set


def banded_replay(oracle, reads, p, s, m, c, first_band=3000):
    k = p.kmer_size
    kept = [r for r in reads if len(r) >= k]
    hashes = [oracle.all_hashes(r, p) for r in kept]                 # per read, in stream order
    n_kmers = np.array([h.size for h in hashes])
    ends = np.cumsum(n_kmers)                                        # k-mer index after each read
    total = int(ends[-1])
    # band starts at read boundaries near total / 2^j; threshold = exact top of the bottom-s(-m) sketch of the prefix
    cuts, target = [], total // 2
    while target >= first_band:
        r = int(np.searchsorted(ends, target, side="right"))        # reads fully inside the prefix
        if r > 0 and (not cuts or ends[r - 1] < cuts[-1][0]):
            cuts.append((int(ends[r - 1]), r))
        target //= 2
    cuts.reverse()
    bands = [(0, 0, None)]                                           # (first k-mer index, first read, threshold or None = keep all)
    for pos, r in cuts:
        allh = np.concatenate(hashes[:r])
        u, cnt = np.unique(allh, return_counts=True)
        q = u[cnt >= m]
        bands.append((pos, r, int(q[s - 1]) if q.size >= s else None))
    # events: k-mers at or below their band's threshold
    L = oracle.lib
    L.mo_heap_m_new.restype = C.c_void_p; L.mo_heap_m_new.argtypes = [C.c_int, C.c_uint64, C.c_uint64]
    L.mo_heap_m_try_insert.argtypes = [C.c_void_p, C.c_uint64]
    L.mo_heap_m_free.argtypes = [C.c_void_p]
    hm = L.mo_heap_m_new(int(p.use64), s, m)
    acc = C.cast(hm, C.POINTER(C.c_void_p))[0]                      # mo_heap_m.acc
    n_events = 0
    used = len(kept)
    band_of_read = np.searchsorted([b[1] for b in bands], np.arange(len(kept)), side="right") - 1
    for r, h in enumerate(hashes):
        thr = bands[band_of_read[r]][2]
        ev = h if thr is None else h[h <= np.uint64(thr)]
        n_events += ev.size
        for x in ev:
            L.mo_heap_m_try_insert(hm, int(x))
        if c > 0 and L.mo_heap_size(acc) and L.mo_heap_estimate_multiplicity(acc) >= c:
            used = r + 1
            break
    out = np.empty(s, np.uint64); cnt = np.empty(s, np.uint32)
    n = L.mo_heap_to_list(acc, out.ctypes.data_as(C.POINTER(C.c_uint64)), cnt.ctypes.data_as(C.POINTER(C.c_uint32)))
    L.mo_heap_m_free(hm)
    return out[:n].copy(), cnt[:n].copy(), used, n_events, total


@pytest.mark.parametrize("m,c,s,cov", [(1, 3.0, 200, 12), (2, 4.0, 100, 15), (1, 1e9, 200, 6), (3, 2.5, 100, 10), (2, 1e9, 150, 8)])
def test_banded_event_replay_equals_full_stream(oracle, m, c, s, cov):
    p = oracle.params(k=21)
    L = oracle.lib
    L.mo_heap_estimate_multiplicity.restype = C.c_double; L.mo_heap_estimate_multiplicity.argtypes = [C.c_void_p]
    L.mo_heap_size.restype = C.c_uint64; L.mo_heap_size.argtypes = [C.c_void_p]
    reads = _read_set(900 + m + s, 30_000, 300 * cov, err=0.005)
    want_h, want_c, _, want_used = oracle.sketch_unit_mc(reads, p, s=s, min_copies=m, target_cov=c, counts=True)
    h, cnt, used, n_events, total = banded_replay(oracle, reads, p, s, m, c)
    assert used == want_used and np.array_equal(h, want_h) and np.array_equal(cnt, want_c)
    assert n_events < total / 3                  # the bands really filter
