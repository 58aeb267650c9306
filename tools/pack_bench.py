"""Host packer throughput (mashgpu_host_pack, no GPU work) for a range of thread counts -- sizing of the opt-in packed feed path."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mash_b200 import _capi

lib = _capi.load_library()
p = _capi.SketchParams()
p.kmer_size = 21; p.sketch_size = 1000; p.seed = 42
lib.mashgpu_set_alphabet(C.byref(p), b"ACGT")
n, L = 256, 5_000_000
rng = np.random.default_rng(1)
base = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, L)]
recs = [base.copy() for _ in range(n)]
ptrs = (C.c_void_p * n)(*[r.ctypes.data for r in recs])
lens = np.full(n, L, np.uint64)
stream_len = int((lens + 1).sum())
codes = np.zeros((stream_len + 31) // 32, np.uint64)
runs = np.zeros(2 * 100000, np.uint64); nr = C.c_uint64(0)
print("nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), flush=True)
for th in [int(x) for x in (sys.argv[1:] or ["1", "8", "16", "24", "32", "48", "64", "96", "128"])]:
    best = 1e9
    for rep in range(3):
        t = time.perf_counter()
        lib.mashgpu_host_pack(C.byref(p), n, ptrs, lens.ctypes.data_as(C.POINTER(C.c_uint64)), th, codes.ctypes.data_as(C.POINTER(C.c_uint64)),
                              runs.ctypes.data_as(C.POINTER(C.c_uint64)), 100000, C.byref(nr))
        best = min(best, time.perf_counter() - t)
    print(th, "threads:", round(stream_len / best / 1e9, 2), "GB/s", flush=True)
