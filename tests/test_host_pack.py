"""Host feed path (mash_b200/csrc/pack.cpp): 2-bit packing + invalid-run list, against an independent numpy
restatement.  CPU only."""
import ctypes as C

import numpy as np
import pytest

from fixtures import synth_genome


def numpy_pack(records, preserve_case):
    stream = bytearray()
    for r in records:
        stream += bytes(r) + b"\x00"
    a = np.frombuffer(bytes(stream), np.uint8).copy()
    if not preserve_case:
        low = (a > 96) & (a < 123)
        a[low] -= 32
    code = np.full(a.size, 4, np.uint8)
    for c, v in ((ord("A"), 0), (ord("C"), 1), (ord("G"), 2), (ord("T"), 3)):
        code[a == c] = v
    n = a.size
    groups = (n + 31) // 32
    padded = np.zeros(groups * 32, np.uint64)
    padded[:n] = code & 3
    shifts = (2 * (np.arange(groups * 32) % 32)).astype(np.uint64)
    words = np.bitwise_or.reduce((padded << shifts).reshape(groups, 32), axis=1)
    inv = code == 4
    runs = []
    i = 0
    idx = np.flatnonzero(inv)
    if idx.size:
        breaks = np.flatnonzero(np.diff(idx) > 1)
        starts = np.concatenate([[idx[0]], idx[breaks + 1]])
        ends = np.concatenate([idx[breaks], [idx[-1]]])
        runs = [(int(s), int(e - s + 1)) for s, e in zip(starts, ends)]
    return words, runs, n


def host_pack(records, preserve_case=False, threads=1):
    import mash_b200
    lib = mash_b200.load_library()
    p = mash_b200.SketchParams()
    p.kmer_size = 21
    p.preserve_case = int(preserve_case)
    lib.mashgpu_set_alphabet(C.byref(p), b"ACGT")
    bufs = [np.frombuffer(bytes(r), np.uint8) for r in records]
    ptrs = (C.c_void_p * max(1, len(bufs)))(*[b.ctypes.data if b.size else None for b in bufs])
    lens = np.array([b.size for b in bufs], np.uint64)
    total = int(lens.sum()) + len(bufs)
    codes = np.zeros(max(1, (total + 31) // 32), np.uint64)
    cap = total + 1
    runs = np.zeros(2 * cap, np.uint64)
    n_runs = C.c_uint64(0)
    u64p = C.POINTER(C.c_uint64)
    rc = lib.mashgpu_host_pack(C.byref(p), len(bufs), C.cast(ptrs, C.c_void_p), lens.ctypes.data_as(u64p), threads,
                               codes.ctypes.data_as(u64p), runs.ctypes.data_as(u64p), cap, C.byref(n_runs))
    assert rc == 0
    r = runs[:2 * n_runs.value].reshape(-1, 2)
    return codes[:(total + 31) // 32], [(int(a), int(b)) for a, b in r], total


@pytest.mark.parametrize("preserve_case", [False, True])
@pytest.mark.parametrize("threads", [1, 7])
def test_pack_matches_numpy(preserve_case, threads):
    rng = np.random.Generator(np.random.PCG64(1))
    recs = [bytes(synth_genome(1, 3_000_017, n_runs=30, lower_frac=0.03)), b"", b"ACGT", b"N" * 100, bytes(synth_genome(2, 31)),
            bytes(synth_genome(3, 32)), bytes(synth_genome(4, 33)), bytes(rng.integers(0, 256, 5000, dtype=np.uint8)),
            bytes(synth_genome(5, 1_500_000)), b"acgtnACGTN*-", bytes(synth_genome(6, 64))]
    codes, runs, total = host_pack(recs, preserve_case, threads)
    want_codes, want_runs, n = numpy_pack(recs, preserve_case)
    assert total == n
    assert runs == want_runs
    # codes at invalid positions are "don't care": compare under the validity mask
    valid = np.ones(((n + 31) // 32) * 32, bool)
    valid[n:] = False
    for s, l in want_runs:
        valid[s:s + l] = False
    shifts = (2 * (np.arange(valid.size) % 32)).astype(np.uint64)
    got = (np.repeat(codes, 32) >> shifts) & np.uint64(3)
    want = (np.repeat(want_codes, 32) >> shifts) & np.uint64(3)
    assert np.array_equal(got[valid], want[valid])


def test_pack_empty():
    codes, runs, total = host_pack([])
    assert total == 0 and runs == []


def test_chunk_mask_packer_matches_byte_by_byte_restatement(tmp_path):
    # pack_chunk_mask (host packer of the screen feed: codes + invalid-position mask, no runs; AVX-512 / AVX2 / scalar paths, 1 and 5
    # threads through the persistent worker pool, lengths 0..49 and random ones up to 40 MB) has no C-ABI entry point of its own: a small
    # C++ harness links pack.cpp and compares every position with the definition
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "pack_mask_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(root, "mash_b200", "csrc"), os.path.join(root, "tools", "pack_mask_test.cpp"),
                           os.path.join(root, "mash_b200", "csrc", "pack.cpp"), "-o", exe, "-pthread"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "bad 0" in out.stdout, out.stdout[-2000:]
