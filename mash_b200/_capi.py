"""ctypes mirror of include/mashgpu.h.

Names follow the reference's seams: Engine.sketch ~ sketchFile/sketchSequence jobs (Sketch.cpp:1147-1365),
Engine.dist ~ compare/compareSketches (CommandDistance.cpp:306-448), Engine.screen_* ~ hashSequence + reduce
(CommandScreen.cpp:484-599, 288-455).  Every call goes through libmashgpu.so; nothing here computes hashes,
merges or p-values on the host.
"""
import ctypes as C
import os

import numpy as np

ALPHABET_NUCLEOTIDE = "ACGT"                 # reference Sketch.h:25
ALPHABET_PROTEIN = "ACDEFGHIKLMNPQRSTVWY"    # reference Sketch.h:26

_HERE = os.path.dirname(os.path.abspath(__file__))

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)
f64p = C.POINTER(C.c_double)


class MashGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mashgpu error {code}: {msg}")
        self.code = code


class SketchParams(C.Structure):
    """mashgpu_sketch_params (Sketch::Parameters subset, reference Sketch.h:34-109)."""
    _fields_ = [("kmer_size", C.c_int32), ("sketch_size", C.c_uint32), ("seed", C.c_uint32),
                ("use64", C.c_int32), ("noncanonical", C.c_int32), ("preserve_case", C.c_int32),
                ("alphabet", C.c_uint8 * 256), ("min_copies", C.c_uint32), ("target_cov", C.c_double)]

    @property
    def kmer_space(self):
        return float(sum(1 for a in self.alphabet if a)) ** self.kmer_size   # Sketch.cpp:509


class DistParams(C.Structure):
    """mashgpu_dist_params (arguments of compareSketches, reference CommandDistance.h:92)."""
    _fields_ = [("sketch_size", C.c_uint64), ("kmer_size", C.c_int32), ("kmer_space", C.c_double),
                ("max_distance", C.c_double), ("max_pvalue", C.c_double)]


class SketchSet(C.Structure):
    _fields_ = [("n", C.c_uint64), ("stride", C.c_uint64), ("hashes", C.c_void_p),
                ("n_hashes", C.c_void_p), ("length", C.c_void_p), ("on_device", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("scan_kernel_ms", C.c_double), ("scan_kernel_launches", C.c_uint64),
                ("dist_kernel_ms", C.c_double), ("dist_kernel_launches", C.c_uint64), ("exact_reruns", C.c_uint64)]


def lib_path():
    return os.path.join(_HERE, "libmashgpu.so")


_lib = None


def load_library():
    """Load libmashgpu.so (built in-tree by __graft_entry__.build()). No fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build the CUDA extension first "
                          "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
    L = C.CDLL(path)
    L.mashgpu_set_alphabet.restype = C.c_uint32
    L.mashgpu_set_alphabet.argtypes = [C.POINTER(SketchParams), C.c_char_p]
    L.mashgpu_device_count.restype = C.c_int
    L.mashgpu_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.mashgpu_destroy.argtypes = [C.c_void_p]
    L.mashgpu_destroy.restype = None
    L.mashgpu_last_error.restype = C.c_char_p
    L.mashgpu_last_error.argtypes = [C.c_void_p]
    L.mashgpu_sketch_batch.argtypes = [C.c_void_p, C.POINTER(SketchParams), C.c_uint64, C.c_void_p, u64p, u32p, C.c_uint64,
                                       u64p, u32p, u32p, u64p]
    L.mashgpu_sketch_stream_dev.argtypes = [C.c_void_p, C.POINTER(SketchParams), C.c_void_p, u64p, C.c_uint64,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mashgpu_sketch_reads.argtypes = [C.c_void_p, C.POINTER(SketchParams), C.c_uint64, C.c_void_p, u64p, u64p, u32p, u32p, u64p]
    L.mashgpu_sketch_batch_packed.argtypes = [C.c_void_p, C.POINTER(SketchParams), u64p, C.c_uint64, u64p, C.c_uint64, u64p, C.c_uint64, u64p, u32p, u32p]
    L.mashgpu_host_pack.argtypes = [C.POINTER(SketchParams), C.c_uint64, C.c_void_p, u64p, C.c_int, u64p, u64p, C.c_uint64, u64p]
    L.mashgpu_hash_windows.argtypes = [C.c_void_p, C.POINTER(SketchParams), C.c_void_p, C.c_uint64, u64p, u8p]
    L.mashgpu_dist_open.argtypes = [C.c_void_p, C.POINTER(SketchSet), C.POINTER(SketchSet), C.POINTER(DistParams), C.POINTER(C.c_void_p)]
    L.mashgpu_dist_run.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, u32p, u32p, f64p, f64p, u8p]
    L.mashgpu_dist_run_dev.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mashgpu_dist_run_list.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, u64p, u32p, u32p, f64p, f64p, u64p]
    L.mashgpu_dist_close.argtypes = [C.c_void_p]
    L.mashgpu_dist_set_prefilter.argtypes = [C.c_void_p, C.c_int]
    L.mashgpu_dist_set_triangle.argtypes = [C.c_void_p, C.c_int]
    L.mashgpu_dist_prefilter_stats.argtypes = [C.c_void_p, u64p, u64p, C.POINTER(C.c_int)]
    L.mashgpu_dist_pair_stats.argtypes = [C.c_void_p, u64p]
    L.mashgpu_dist.argtypes = [C.c_void_p, C.POINTER(SketchSet), C.POINTER(SketchSet), C.POINTER(DistParams), u32p, u32p, f64p, f64p, u8p]
    L.mashgpu_screen_open.argtypes = [C.c_void_p, C.POINTER(SketchParams), C.POINTER(SketchSet), C.POINTER(C.c_void_p)]
    L.mashgpu_screen_feed.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.mashgpu_screen_feed_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.mashgpu_screen_finish.argtypes = [C.c_void_p, u64p, u64p, f64p, f64p, u64p, u64p, u32p]
    L.mashgpu_screen_counters.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), u64p]
    L.mashgpu_screen_merge_mixture.argtypes = [C.c_void_p, u64p, C.c_uint32]
    L.mashgpu_screen_close.argtypes = [C.c_void_p]
    L.mashgpu_screen_mixture_dev.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    L.mashgpu_screen_merge_mixtures_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64]
    L.mashgpu_dict_local_sort.argtypes = [C.c_void_p, C.POINTER(SketchSet), C.c_uint64, C.c_void_p, C.c_void_p, u64p, C.c_void_p]
    L.mashgpu_dict_split.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, u64p, C.c_uint32, u64p, C.c_void_p]
    L.mashgpu_dict_rank.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, u64p, C.c_void_p]
    L.mashgpu_dict_scatter.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, u64p, u64p, C.c_uint32, C.POINTER(SketchSet), C.c_uint64,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
    L.mashgpu_dist_open_encoded.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64,
                                            C.POINTER(DistParams), C.POINTER(C.c_void_p)]
    L.mashgpu_screen_set_winner.argtypes = [C.c_void_p, C.c_int]
    L.mashgpu_set_timing.argtypes = [C.c_void_p, C.c_int]
    L.mashgpu_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats), C.c_int]
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _as_u8(s):
    if isinstance(s, (bytes, bytearray, memoryview)):
        return np.frombuffer(s, dtype=np.uint8)
    return np.ascontiguousarray(s, dtype=np.uint8)


class _Set:
    """Keeps the numpy arrays behind a mashgpu_sketch_set alive."""

    def __init__(self, hashes, n_hashes, length=None, on_device=False, n=None, stride=None):
        if on_device:
            self.c = SketchSet(n, stride, hashes, n_hashes, length, 1)
            return
        self.h = np.ascontiguousarray(hashes, np.uint64)
        if self.h.ndim != 2:
            raise ValueError("hashes must be (n, stride)")
        self.n = np.ascontiguousarray(n_hashes, np.uint32)
        self.l = np.ascontiguousarray(length if length is not None else np.ones(self.h.shape[0]), np.uint64)
        self.c = SketchSet(self.h.shape[0], self.h.shape[1], self.h.ctypes.data, self.n.ctypes.data, self.l.ctypes.data, 0)


class Engine:
    """One mashgpu context (one CUDA device)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.mashgpu_create(device, C.byref(h))
        if rc != 0:
            raise MashGpuError(rc, (self.lib.mashgpu_last_error(None) or b"").decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.mashgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise MashGpuError(rc, (self.lib.mashgpu_last_error(self.h) or b"").decode())

    # ---- parameters (sketchParameterSetup / setAlphabetFromString) --------------------------------------
    def params(self, k=21, s=1000, seed=42, alphabet=ALPHABET_NUCLEOTIDE, noncanonical=False, preserve_case=False, min_copies=1, target_cov=0.0):
        p = SketchParams()
        p.min_copies = min_copies
        p.target_cov = target_cov
        p.kmer_size = k
        p.sketch_size = s
        p.seed = seed
        p.noncanonical = int(noncanonical)
        p.preserve_case = int(preserve_case)
        self.lib.mashgpu_set_alphabet(C.byref(p), alphabet.encode())
        return p

    # ---- hot path 1 ----------------------------------------------------------------------------------------
    def sketch(self, records, p, unit_of_record=None, n_units=None, counts=False):
        """records: list of byte strings / uint8 arrays.  Returns (hashes (n_units, s) u64, n u32, length u64[, counts])."""
        bufs = [_as_u8(r) for r in records]
        ptrs = (C.c_void_p * max(1, len(bufs)))(*[b.ctypes.data if b.size else None for b in bufs])
        lens = np.array([b.size for b in bufs], dtype=np.uint64)
        if unit_of_record is None:
            n_units = len(bufs)
            uor = None
        else:
            uor = np.ascontiguousarray(unit_of_record, np.uint32)
            if n_units is None:
                n_units = int(uor.max()) + 1 if uor.size else 0
        s = p.sketch_size
        out = np.zeros((n_units, s), np.uint64)
        out_n = np.zeros(n_units, np.uint32)
        out_len = np.zeros(n_units, np.uint64)
        out_c = np.zeros((n_units, s), np.uint32) if counts else None
        self._check(self.lib.mashgpu_sketch_batch(self.h, C.byref(p), len(bufs), C.cast(ptrs, C.c_void_p), _p(lens, u64p),
                                                  _p(uor, u32p), n_units, _p(out, u64p), _p(out_c, u32p), _p(out_n, u32p), _p(out_len, u64p)))
        return (out, out_n, out_len, out_c) if counts else (out, out_n, out_len)

    def sketch_reads(self, records, p, counts=False):
        """`mash sketch -r [-m] [-c]` of one read set.  Returns (hashes[:n], counts[:n] | None, records_used)."""
        bufs = [_as_u8(r) for r in records]
        ptrs = (C.c_void_p * max(1, len(bufs)))(*[b.ctypes.data if b.size else None for b in bufs])
        lens = np.array([b.size for b in bufs], dtype=np.uint64)
        s = p.sketch_size
        out = np.zeros(s, np.uint64); out_c = np.zeros(s, np.uint32) if counts else None
        n = C.c_uint32(0); used = C.c_uint64(0)
        self._check(self.lib.mashgpu_sketch_reads(self.h, C.byref(p), len(bufs), C.cast(ptrs, C.c_void_p), _p(lens, u64p), _p(out, u64p), _p(out_c, u32p),
                                                  C.byref(n), C.byref(used)))
        return out[:n.value].copy(), (out_c[:n.value].copy() if counts else None), used.value

    def host_pack(self, records, p, threads=1):
        """records -> (codes u64[ceil(L/32)], runs u64[n, 2], record_start u64[n_records + 1]) in the packed stream format
        (every record followed by one separator position).  No GPU involved."""
        bufs = [_as_u8(r) for r in records]
        ptrs = (C.c_void_p * max(1, len(bufs)))(*[b.ctypes.data if b.size else None for b in bufs])
        lens = np.array([b.size for b in bufs], dtype=np.uint64)
        total = int(lens.sum()) + len(bufs)
        codes = np.zeros(max(1, (total + 31) // 32), np.uint64)
        cap = 1 << 16
        while True:
            runs = np.zeros(2 * cap, np.uint64)
            n_runs = C.c_uint64(0)
            self._check(self.lib.mashgpu_host_pack(C.byref(p), len(bufs), C.cast(ptrs, C.c_void_p), _p(lens, u64p), threads, _p(codes, u64p),
                                                   _p(runs, u64p), cap, C.byref(n_runs)))
            if n_runs.value <= cap:
                break
            cap = int(n_runs.value)
        starts = np.zeros(len(bufs) + 1, np.uint64)
        starts[1:] = np.cumsum(lens + np.uint64(1))
        return codes, runs[:2 * n_runs.value].reshape(-1, 2).copy(), starts

    def sketch_packed(self, codes, stream_len, runs, unit_start, p, counts=False):
        """mashgpu_sketch_batch_packed: units of a caller-packed 2-bit stream.  Returns (hashes, n[, counts])."""
        codes = np.ascontiguousarray(codes, np.uint64); runs = np.ascontiguousarray(runs, np.uint64).reshape(-1)
        us = np.ascontiguousarray(unit_start, np.uint64)
        n_units = us.size - 1
        s = p.sketch_size
        out = np.zeros((n_units, s), np.uint64); out_n = np.zeros(n_units, np.uint32)
        out_c = np.zeros((n_units, s), np.uint32) if counts else None
        self._check(self.lib.mashgpu_sketch_batch_packed(self.h, C.byref(p), _p(codes, u64p), stream_len, _p(runs, u64p) if runs.size else None, runs.size // 2,
                                                         _p(us, u64p), n_units, _p(out, u64p), _p(out_c, u32p), _p(out_n, u32p)))
        return (out, out_n, out_c) if counts else (out, out_n)

    def sketch_stream_dev(self, p, d_stream_ptr, unit_start, d_out_hashes, d_out_n, d_out_counts=None, stream=None):
        us = np.ascontiguousarray(unit_start, np.uint64)
        self._check(self.lib.mashgpu_sketch_stream_dev(self.h, C.byref(p), d_stream_ptr, _p(us, u64p), us.size - 1,
                                                       d_out_hashes, d_out_counts, d_out_n, stream))

    def hash_windows(self, seq, p):
        b = _as_u8(seq)
        n = max(0, b.size - p.kmer_size + 1)
        h = np.zeros(max(1, n), np.uint64)
        v = np.zeros(max(1, n), np.uint8)
        self._check(self.lib.mashgpu_hash_windows(self.h, C.byref(p), b.ctypes.data, b.size, _p(h, u64p), _p(v, u8p)))
        return h[:n], v[:n].astype(bool)

    # ---- hot path 2 ----------------------------------------------------------------------------------------
    def dist(self, ref, ref_n, ref_len, qry=None, qry_n=None, qry_len=None, *, sketch_size, k, kmer_space,
             max_distance=1.0, max_pvalue=1.0):
        """All pairs, query-major.  Returns a dict of (n_qry, n_ref) arrays: numer, denom, distance, pvalue, pass."""
        job = self.dist_open(ref, ref_n, ref_len, qry, qry_n, qry_len, sketch_size=sketch_size, k=k, kmer_space=kmer_space,
                             max_distance=max_distance, max_pvalue=max_pvalue)
        try:
            return job.run(0, job.n_qry)
        finally:
            job.close()

    def dist_open(self, ref, ref_n, ref_len, qry=None, qry_n=None, qry_len=None, *, sketch_size, k, kmer_space,
                  max_distance=1.0, max_pvalue=1.0):
        return DistJob(self, ref, ref_n, ref_len, qry, qry_n, qry_len, sketch_size, k, kmer_space, max_distance, max_pvalue)

    def dist_open_encoded(self, d_rows, d_n_eff, d_length, n_rows, ref_begin, ref_count, *, sketch_size, k, kmer_space,
                          max_distance=1.0, max_pvalue=1.0, keepalive=None):
        """Job over dictionary-encoded rows in device memory (pointers); references = rows [ref_begin, ref_begin + ref_count)."""
        return DistJob.encoded(self, d_rows, d_n_eff, d_length, n_rows, ref_begin, ref_count, sketch_size, k, kmer_space,
                               max_distance, max_pvalue, keepalive)

    # ---- sharded dictionary build (device pointers; see mash_b200/shard.py) ----------------------------------
    def dict_local_sort(self, d_hashes, d_n_hashes, n, stride, sketch_size, d_keys, d_slots, stream=None):
        st = SketchSet(n, stride, d_hashes, d_n_hashes, None, 1)
        nv = C.c_uint64(0)
        self._check(self.lib.mashgpu_dict_local_sort(self.h, C.byref(st), sketch_size, d_keys, d_slots, C.byref(nv), stream))
        return nv.value

    def dict_split(self, d_keys, n, splitters, stream=None):
        spl = np.ascontiguousarray(splitters, np.uint64)
        counts = np.zeros(spl.size + 1, np.uint64)
        self._check(self.lib.mashgpu_dict_split(self.h, d_keys, n, _p(spl, u64p), spl.size + 1, _p(counts, u64p), stream))
        return [int(c) for c in counts]

    def dict_rank(self, d_keys, n, d_codes, stream=None):
        nd = C.c_uint64(0)
        self._check(self.lib.mashgpu_dict_rank(self.h, d_keys, n, d_codes, C.byref(nd), stream))
        return nd.value

    def dict_scatter(self, d_codes, d_slots, seg_counts, seg_base, d_n_hashes, n, stride, sketch_size, d_rows, d_n_eff, stream=None):
        sc = np.ascontiguousarray(seg_counts, np.uint64); sb = np.ascontiguousarray(seg_base, np.uint64)
        st = SketchSet(n, stride, 1, d_n_hashes, None, 1)          # the hashes themselves are not read again
        self._check(self.lib.mashgpu_dict_scatter(self.h, d_codes, d_slots, _p(sc, u64p), _p(sb, u64p), sc.size, C.byref(st), sketch_size,
                                                  d_rows, d_n_eff, stream))

    # ---- hot path 3 ----------------------------------------------------------------------------------------
    def screen_open(self, ref, ref_n, p, ref_len=None):
        return ScreenJob(self, ref, ref_n, p, ref_len)

    # ---- instrumentation -----------------------------------------------------------------------------------
    def set_timing(self, on=True):
        self._check(self.lib.mashgpu_set_timing(self.h, int(on)))

    def stats(self, reset=False):
        st = Stats()
        self._check(self.lib.mashgpu_get_stats(self.h, C.byref(st), int(reset)))
        return {f: getattr(st, f) for f, _ in Stats._fields_}


class DistJob:
    def __init__(self, eng, ref, ref_n, ref_len, qry, qry_n, qry_len, sketch_size, k, kmer_space, max_distance, max_pvalue):
        self.eng = eng
        if isinstance(ref, _Set):
            self._ref = ref
            self._qry = qry
        else:
            self._ref = _Set(ref, ref_n, ref_len)
            self._qry = None if qry is None else _Set(qry, qry_n, qry_len)
        self.n_ref = int(self._ref.c.n)
        self.n_qry = self.n_ref if self._qry is None else int(self._qry.c.n)
        self.params = DistParams(sketch_size, k, kmer_space, max_distance, max_pvalue)
        h = C.c_void_p()
        eng._check(eng.lib.mashgpu_dist_open(eng.h, C.byref(self._ref.c), C.byref(self._qry.c) if self._qry is not None else None,
                                             C.byref(self.params), C.byref(h)))
        self.h = h

    @classmethod
    def encoded(cls, eng, d_rows, d_n_eff, d_length, n_rows, ref_begin, ref_count, sketch_size, k, kmer_space, max_distance, max_pvalue,
                keepalive=None):
        self = cls.__new__(cls)
        self.eng = eng
        self._ref = keepalive        # the caller's device arrays must outlive the job
        self._qry = None
        self.n_ref, self.n_qry = int(ref_count), int(n_rows)
        self.params = DistParams(sketch_size, k, kmer_space, max_distance, max_pvalue)
        h = C.c_void_p()
        eng._check(eng.lib.mashgpu_dist_open_encoded(eng.h, d_rows, d_n_eff, d_length, n_rows, ref_begin, ref_count, C.byref(self.params), C.byref(h)))
        self.h = h
        return self

    def run(self, q_begin, q_count):
        n = q_count * self.n_ref
        numer = np.zeros(n, np.uint32); denom = np.zeros(n, np.uint32)
        dist = np.zeros(n, np.float64); pv = np.zeros(n, np.float64); ok = np.zeros(n, np.uint8)
        self.eng._check(self.eng.lib.mashgpu_dist_run(self.h, q_begin, q_count, _p(numer, u32p), _p(denom, u32p), _p(dist, f64p), _p(pv, f64p), _p(ok, u8p)))
        shp = (q_count, self.n_ref)
        return {"numer": numer.reshape(shp), "denom": denom.reshape(shp), "distance": dist.reshape(shp),
                "pvalue": pv.reshape(shp), "pass": ok.reshape(shp).astype(bool)}

    def run_list(self, q_begin, q_count, capacity):
        """Passing pairs only, sorted by pair index. Returns (n_pass, dict of arrays) -- arrays empty when n_pass > capacity."""
        idx = np.zeros(capacity, np.uint64); numer = np.zeros(capacity, np.uint32); denom = np.zeros(capacity, np.uint32)
        dist = np.zeros(capacity, np.float64); pv = np.zeros(capacity, np.float64); n = C.c_uint64(0)
        self.eng._check(self.eng.lib.mashgpu_dist_run_list(self.h, q_begin, q_count, capacity, _p(idx, u64p), _p(numer, u32p), _p(denom, u32p),
                                                           _p(dist, f64p), _p(pv, f64p), C.byref(n)))
        m = n.value if n.value <= capacity else 0
        return n.value, {"index": idx[:m], "numer": numer[:m], "denom": denom[:m], "distance": dist[:m], "pvalue": pv[:m]}

    def run_dev(self, q_begin, q_count, d_numer=None, d_denom=None, d_distance=None, d_pvalue=None, d_pass=None, stream=None):
        self.eng._check(self.eng.lib.mashgpu_dist_run_dev(self.h, q_begin, q_count, d_numer, d_denom, d_distance, d_pvalue, d_pass, stream))

    def set_prefilter(self, mode):
        """-1 = auto (default), 0 = merge every pair, 1 = always probe the reference tiles first."""
        self.eng._check(self.eng.lib.mashgpu_dist_set_prefilter(self.h, int(mode)))

    def set_triangle(self, on=True):
        """Self comparison, lower triangle only (pairs with r >= q are not computed or written)."""
        self.eng._check(self.eng.lib.mashgpu_dist_set_triangle(self.h, int(on)))

    def prefilter_stats(self):
        probed = C.c_uint64(0); flagged = C.c_uint64(0); active = C.c_int(0)
        self.eng._check(self.eng.lib.mashgpu_dist_prefilter_stats(self.h, C.byref(probed), C.byref(flagged), C.byref(active)))
        pairs = C.c_uint64(0)
        self.eng._check(self.eng.lib.mashgpu_dist_pair_stats(self.h, C.byref(pairs)))
        return {"combos_probed": probed.value, "combos_flagged": flagged.value, "active": bool(active.value), "pairs_from_lists": pairs.value}

    def close(self):
        if self.h:
            self.eng.lib.mashgpu_dist_close(self.h)
            self.h = None


class ScreenJob:
    def __init__(self, eng, ref, ref_n, p, ref_len=None):
        self.eng = eng
        self._ref = ref if isinstance(ref, _Set) else _Set(ref, ref_n, ref_len)
        self.n_ref = int(self._ref.c.n)
        self.p = p
        h = C.c_void_p()
        eng._check(eng.lib.mashgpu_screen_open(eng.h, C.byref(p), C.byref(self._ref.c), C.byref(h)))
        self.h = h

    def feed(self, chunk):
        b = _as_u8(chunk)
        self.eng._check(self.eng.lib.mashgpu_screen_feed(self.h, b.ctypes.data, b.size))

    def feed_dev(self, d_ptr, length):
        self.eng._check(self.eng.lib.mashgpu_screen_feed_dev(self.h, d_ptr, length))

    def set_winner(self, on=True):
        """`mash screen -w`: reallocate every seen reference hash to the best sketch containing it before the reduce."""
        self.eng._check(self.eng.lib.mashgpu_screen_set_winner(self.h, int(on)))

    def counters(self):
        """(device pointer, slot count) of the uint32 hit counters, for an all-reduce across ranks."""
        ptr = C.c_void_p(); n = C.c_uint64(0)
        self.eng._check(self.eng.lib.mashgpu_screen_counters(self.h, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def merge_mixture(self, hashes):
        h = np.ascontiguousarray(hashes, np.uint64)
        self.eng._check(self.eng.lib.mashgpu_screen_merge_mixture(self.h, _p(h, u64p), h.size))

    def mixture_dev(self):
        """(device pointer of the running bottom-s list, device pointer of its length) for an all-gather across ranks."""
        a = C.c_void_p(); b = C.c_void_p()
        self.eng._check(self.eng.lib.mashgpu_screen_mixture_dev(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def merge_mixtures_dev(self, d_hashes, d_n, n_lists, stride):
        self.eng._check(self.eng.lib.mashgpu_screen_merge_mixtures_dev(self.h, d_hashes, d_n, n_lists, stride))

    def mixture(self):
        """The running bottom-s list of the mixture (ascending u64)."""
        return self.finish_mixture_only()

    def finish_mixture_only(self):
        s = self.p.sketch_size
        set_size = C.c_uint64(0); mix = np.zeros(s, np.uint64); mix_n = C.c_uint32(0)
        self.eng._check(self.eng.lib.mashgpu_screen_finish(self.h, None, None, None, None, C.byref(set_size), _p(mix, u64p), C.byref(mix_n)))
        return mix[:mix_n.value].copy()

    def finish(self):
        n, s = self.n_ref, self.p.sketch_size
        shared = np.zeros(n, np.uint64); median = np.zeros(n, np.uint64)
        ident = np.zeros(n, np.float64); pv = np.zeros(n, np.float64)
        set_size = C.c_uint64(0); mix = np.zeros(s, np.uint64); mix_n = C.c_uint32(0)
        self.eng._check(self.eng.lib.mashgpu_screen_finish(self.h, _p(shared, u64p), _p(median, u64p), _p(ident, f64p), _p(pv, f64p),
                                                           C.byref(set_size), _p(mix, u64p), C.byref(mix_n)))
        return dict(shared=shared, median=median, identity=ident, pvalue=pv, set_size=set_size.value, mixture=mix[:mix_n.value].copy())

    def close(self):
        if self.h:
            self.eng.lib.mashgpu_screen_close(self.h)
            self.h = None
