"""ctypes bindings for the checker libraries in oracle/ (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (mash_b200) never does.

  Oracle     -> oracle/libmash_oracle.so   (plain-C restatement, oracle/mash_oracle.c)
  RefLib     -> oracle/_ref/libmash_ref.so (the reference's own hash/heap object code + ref_shim.cpp)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
f64p = C.POINTER(C.c_double)


class Params(C.Structure):
    """mo_params / ref_params (Sketch::Parameters subset, Sketch.h:34-109)."""
    _fields_ = [("kmer_size", C.c_int32), ("seed", C.c_uint32), ("use64", C.c_int32),
                ("noncanonical", C.c_int32), ("preserve_case", C.c_int32),
                ("alphabet", C.c_uint8 * 256)]


class PairOutput(C.Structure):
    _fields_ = [("numer", C.c_uint64), ("denom", C.c_uint64), ("distance", C.c_double),
                ("pvalue", C.c_double), ("pass_", C.c_int32), ("filled", C.c_int32)]


ALPHABET_NUCLEOTIDE = "ACGT"                       # Sketch.h:25
ALPHABET_PROTEIN = "ACDEFGHIKLMNPQRSTVWY"          # Sketch.h:26


def build(force=False):
    """Compile oracle/ (and oracle/_ref when /root/reference is present)."""
    lib = os.path.join(HERE, "libmash_oracle.so")
    if force or not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(os.path.join(HERE, "mash_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "libmash_oracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/src/mash"):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


def _ptr(a, t):
    return a.ctypes.data_as(t)


def _seq_arrays(seqs):
    """list of bytes/np.uint8 arrays -> (keepalive, char** array, u64 lens)."""
    bufs = [np.frombuffer(s, dtype=np.uint8) if isinstance(s, (bytes, bytearray)) else np.ascontiguousarray(s, dtype=np.uint8) for s in seqs]
    ptrs = (C.c_void_p * len(bufs))(*[b.ctypes.data if b.size else 0 for b in bufs])
    lens = np.array([b.size for b in bufs], dtype=np.uint64)
    return bufs, ptrs, lens


class Oracle:
    def __init__(self):
        build()
        self.lib = L = C.CDLL(os.path.join(HERE, "libmash_oracle.so"))
        L.mo_murmur3_x64_128.restype = C.c_uint64
        L.mo_murmur3_x64_128.argtypes = [C.c_char_p, C.c_int, C.c_uint32, u64p]
        L.mo_get_hash.restype = C.c_uint64
        L.mo_get_hash.argtypes = [C.c_char_p, C.c_int, C.c_uint32, C.c_int]
        L.mo_set_alphabet.restype = C.c_uint32
        L.mo_set_alphabet.argtypes = [C.POINTER(Params), C.c_char_p]
        L.mo_all_hashes.restype = C.c_uint64
        L.mo_all_hashes.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(Params), u64p]
        L.mo_sketch_unit.restype = C.c_uint64
        L.mo_sketch_unit.argtypes = [C.POINTER(Params), C.c_uint64, C.c_uint64, C.c_void_p, u64p,
                                     C.c_int, C.c_uint64, u64p, u32p, u64p]
        L.mo_binomial_upper_tail.restype = C.c_double
        L.mo_binomial_upper_tail.argtypes = [C.c_uint64, C.c_double, C.c_uint64]
        L.mo_pvalue.restype = C.c_double
        L.mo_pvalue.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_uint64]
        L.mo_pvalue_within.restype = C.c_double
        L.mo_pvalue_within.argtypes = [C.c_uint64, C.c_uint64, C.c_double, C.c_uint64]
        L.mo_estimate_identity.restype = C.c_double
        L.mo_estimate_identity.argtypes = [C.c_uint64, C.c_uint64, C.c_int]
        L.mo_compare_sketches.restype = None
        L.mo_compare_sketches.argtypes = [C.POINTER(PairOutput), u64p, C.c_uint64, C.c_uint64,
                                          u64p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int,
                                          C.c_double, C.c_double, C.c_double]
        L.mo_compare_all.restype = None
        L.mo_compare_all.argtypes = [C.POINTER(PairOutput),
                                     u64p, u32p, u64p, C.c_uint64, C.c_uint64,
                                     u64p, u32p, u64p, C.c_uint64, C.c_uint64,
                                     C.c_uint64, C.c_int, C.c_double, C.c_double, C.c_double,
                                     C.c_uint64, C.c_uint64]
        L.mo_heap_new.restype = C.c_void_p
        L.mo_heap_new.argtypes = [C.c_int, C.c_uint64]
        L.mo_heap_free.argtypes = [C.c_void_p]
        L.mo_heap_try_insert.argtypes = [C.c_void_p, C.c_uint64]
        L.mo_heap_size.restype = C.c_uint64
        L.mo_heap_size.argtypes = [C.c_void_p]
        L.mo_heap_estimate_set_size.restype = C.c_double
        L.mo_heap_estimate_set_size.argtypes = [C.c_void_p]
        L.mo_heap_to_list.restype = C.c_uint64
        L.mo_heap_to_list.argtypes = [C.c_void_p, u64p, u32p]
        L.mo_hash_sequence.restype = None
        L.mo_hash_sequence.argtypes = [C.c_void_p, u64p, u32p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(Params)]
        L.mo_screen_build_table.restype = C.c_uint64
        L.mo_screen_build_table.argtypes = [u64p, u32p, C.c_uint64, C.c_uint64, u64p]
        L.mo_screen_finish_winner.restype = None
        L.mo_screen_finish_winner.argtypes = [u64p, u32p, u64p, C.c_uint64, C.c_uint64, u64p, u32p, C.c_uint64,
                                              C.c_uint64, C.c_int, C.c_double, u64p, u64p, f64p, f64p]
        L.mo_screen_finish.restype = None
        L.mo_screen_finish.argtypes = [u64p, u32p, C.c_uint64, C.c_uint64, u64p, u32p, C.c_uint64,
                                       C.c_uint64, C.c_int, C.c_double, u64p, u64p, f64p, f64p]

    # ---- parameters -------------------------------------------------------------------------
    def params(self, k=21, seed=42, alphabet=ALPHABET_NUCLEOTIDE, noncanonical=False, preserve_case=False):
        p = Params()
        p.kmer_size = k
        p.seed = seed
        p.noncanonical = int(noncanonical)
        p.preserve_case = int(preserve_case)
        self.alphabet_size = self.lib.mo_set_alphabet(C.byref(p), alphabet.encode())
        return p

    @staticmethod
    def kmer_space(p):
        n = sum(p.alphabet)
        return float(n) ** p.kmer_size       # Sketch.cpp:509  pow(alphabetSize, kmerSize)

    # ---- hashing / sketching ---------------------------------------------------------------
    def get_hash(self, kmer: bytes, seed=42, use64=True):
        return self.lib.mo_get_hash(kmer, len(kmer), seed, int(use64))

    def all_hashes(self, seq, p):
        b = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, np.uint8)
        out = np.empty(max(1, b.size), dtype=np.uint64)
        n = self.lib.mo_all_hashes(b.ctypes.data, b.size, C.byref(p), _ptr(out, u64p))
        return out[:n].copy()

    def sketch_unit(self, records, p, s=1000, reads=False, genome_size=0, counts=False):
        """records: list of byte strings -> (hashes u64 ascending, counts|None, length)."""
        bufs, ptrs, lens = _seq_arrays(records)
        out = np.empty(s, dtype=np.uint64)
        cnt = np.empty(s, dtype=np.uint32)
        length = C.c_uint64(0)
        n = self.lib.mo_sketch_unit(C.byref(p), s, len(bufs), C.cast(ptrs, C.c_void_p), _ptr(lens, u64p),
                                    int(reads), genome_size, _ptr(out, u64p), _ptr(cnt, u32p), C.byref(length))
        return out[:n].copy(), (cnt[:n].copy() if counts else None), length.value

    def sketch_unit_m(self, records, p, s=1000, min_copies=1, reads=True, genome_size=0, counts=False):
        """`mash sketch -r -m min_copies` (MinHashHeap.cpp:96-144): hashes enter the bottom-s at their m-th accepted occurrence."""
        L = self.lib
        L.mo_sketch_unit_m.restype = C.c_uint64
        L.mo_sketch_unit_m.argtypes = [C.POINTER(Params), C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, u64p,
                                       C.c_int, C.c_uint64, u64p, u32p, u64p]
        bufs, ptrs, lens = _seq_arrays(records)
        out = np.empty(s, dtype=np.uint64)
        cnt = np.empty(s, dtype=np.uint32)
        length = C.c_uint64(0)
        n = L.mo_sketch_unit_m(C.byref(p), s, min_copies, len(bufs), C.cast(ptrs, C.c_void_p), _ptr(lens, u64p),
                               int(reads), genome_size, _ptr(out, u64p), _ptr(cnt, u32p), C.byref(length))
        return out[:n].copy(), (cnt[:n].copy() if counts else None), length.value

    def sketch_unit_mc(self, records, p, s=1000, min_copies=1, target_cov=0.0, genome_size=0, counts=False):
        """`mash sketch -r -m min_copies -c target_cov`: returns (hashes, counts|None, length, records_used)."""
        L = self.lib
        L.mo_sketch_unit_mc.restype = C.c_uint64
        L.mo_sketch_unit_mc.argtypes = [C.POINTER(Params), C.c_uint64, C.c_uint64, C.c_double, C.c_uint64, C.c_void_p, u64p,
                                        C.c_int, C.c_uint64, u64p, u32p, u64p, u64p]
        bufs, ptrs, lens = _seq_arrays(records)
        out = np.empty(s, dtype=np.uint64); cnt = np.empty(s, dtype=np.uint32)
        length = C.c_uint64(0); used = C.c_uint64(0)
        n = L.mo_sketch_unit_mc(C.byref(p), s, min_copies, target_cov, len(bufs), C.cast(ptrs, C.c_void_p), _ptr(lens, u64p),
                                1, genome_size, _ptr(out, u64p), _ptr(cnt, u32p), C.byref(length), C.byref(used))
        return out[:n].copy(), (cnt[:n].copy() if counts else None), length.value, used.value

    # ---- dist --------------------------------------------------------------------------------
    def binomial_upper_tail(self, x, r, n):
        return self.lib.mo_binomial_upper_tail(x, r, n)

    def pvalue(self, x, len_ref, len_qry, kmer_space, sketch_size):
        return self.lib.mo_pvalue(x, len_ref, len_qry, kmer_space, sketch_size)

    def compare_sketches(self, ref, len_ref, qry, len_qry, sketch_size, k, kmer_space, max_distance=1.0, max_pvalue=1.0):
        ref = np.ascontiguousarray(ref, np.uint64)
        qry = np.ascontiguousarray(qry, np.uint64)
        o = PairOutput()
        self.lib.mo_compare_sketches(C.byref(o), _ptr(ref, u64p), ref.size, len_ref, _ptr(qry, u64p), qry.size, len_qry,
                                     sketch_size, k, kmer_space, max_distance, max_pvalue)
        return o

    def compare_all(self, ref, ref_n, ref_len, qry, qry_n, qry_len, sketch_size, k, kmer_space,
                    max_distance=1.0, max_pvalue=1.0, q_begin=0, q_end=None):
        """Dense (n x stride) u64 arrays. Returns structured arrays in query-major order."""
        ref = np.ascontiguousarray(ref, np.uint64); qry = np.ascontiguousarray(qry, np.uint64)
        ref_n = np.ascontiguousarray(ref_n, np.uint32); qry_n = np.ascontiguousarray(qry_n, np.uint32)
        ref_len = np.ascontiguousarray(ref_len, np.uint64); qry_len = np.ascontiguousarray(qry_len, np.uint64)
        nr, nq = ref.shape[0], qry.shape[0]
        if q_end is None:
            q_end = nq
        out = (PairOutput * (nr * nq))()
        self.lib.mo_compare_all(out, _ptr(ref, u64p), _ptr(ref_n, u32p), _ptr(ref_len, u64p), nr, ref.shape[1],
                                _ptr(qry, u64p), _ptr(qry_n, u32p), _ptr(qry_len, u64p), nq, qry.shape[1],
                                sketch_size, k, kmer_space, max_distance, max_pvalue, q_begin, q_end)
        a = np.frombuffer(out, dtype=np.dtype([("numer", "<u8"), ("denom", "<u8"), ("distance", "<f8"),
                                               ("pvalue", "<f8"), ("pass", "<i4"), ("filled", "<i4")]))
        return a.reshape(nq, nr).copy()

    # ---- screen ------------------------------------------------------------------------------
    def screen(self, ref, ref_n, chunks, p, s=1000, winner=False, ref_len=None):
        """ref: (n x stride) u64; chunks: list of '*'-joined byte strings (CommandScreen.cpp:224-262).
        winner: `-w` reallocation (CommandScreen.cpp:357-407), needs ref_len.
        Returns dict(shared, median, identity, pvalue, set_size, counts, keys, mixture)."""
        ref = np.ascontiguousarray(ref, np.uint64); ref_n = np.ascontiguousarray(ref_n, np.uint32)
        nr, stride = ref.shape
        keys = np.empty(max(1, int(ref_n.sum())), dtype=np.uint64)
        nk = self.lib.mo_screen_build_table(_ptr(ref, u64p), _ptr(ref_n, u32p), nr, stride, _ptr(keys, u64p))
        keys = keys[:nk].copy()
        counts = np.zeros(max(1, nk), dtype=np.uint32)
        heap = self.lib.mo_heap_new(int(p.use64), s)
        for ch in chunks:
            b = np.frombuffer(ch, dtype=np.uint8) if isinstance(ch, (bytes, bytearray)) else np.ascontiguousarray(ch, np.uint8)
            self.lib.mo_hash_sequence(heap, _ptr(keys, u64p), _ptr(counts, u32p), nk, b.ctypes.data, b.size, C.byref(p))
        set_size = int(np.uint64(self.lib.mo_heap_estimate_set_size(heap))) if self.lib.mo_heap_size(heap) else 0
        mix = np.empty(s, dtype=np.uint64)
        nm = self.lib.mo_heap_to_list(heap, _ptr(mix, u64p), None)
        self.lib.mo_heap_free(heap)
        shared = np.zeros(nr, np.uint64); median = np.zeros(nr, np.uint64)
        ident = np.zeros(nr, np.float64); pv = np.zeros(nr, np.float64)
        if winner:
            ref_len = np.ascontiguousarray(ref_len, np.uint64)
            self.lib.mo_screen_finish_winner(_ptr(ref, u64p), _ptr(ref_n, u32p), _ptr(ref_len, u64p), nr, stride, _ptr(keys, u64p), _ptr(counts, u32p), nk,
                                             set_size, p.kmer_size, self.kmer_space(p), _ptr(shared, u64p), _ptr(median, u64p),
                                             _ptr(ident, f64p), _ptr(pv, f64p))
        else:
            self.lib.mo_screen_finish(_ptr(ref, u64p), _ptr(ref_n, u32p), nr, stride, _ptr(keys, u64p), _ptr(counts, u32p), nk,
                                      set_size, p.kmer_size, self.kmer_space(p), _ptr(shared, u64p), _ptr(median, u64p),
                                      _ptr(ident, f64p), _ptr(pv, f64p))
        return dict(shared=shared, median=median, identity=ident, pvalue=pv, set_size=set_size,
                    counts=counts[:nk], keys=keys, mixture=mix[:nm].copy())


class RefLib:
    """The reference's own hash + heap object code (oracle/_ref/libmash_ref.so)."""

    @staticmethod
    def path():
        return os.path.join(HERE, "_ref", "libmash_ref.so")

    @staticmethod
    def available():
        return os.path.exists(RefLib.path())

    def __init__(self):
        self.lib = L = C.CDLL(self.path())
        L.ref_get_hash.restype = C.c_uint64
        L.ref_get_hash.argtypes = [C.c_char_p, C.c_int, C.c_uint32, C.c_int]
        L.ref_sketch_unit.restype = C.c_uint64
        L.ref_sketch_unit.argtypes = [C.POINTER(Params), C.c_uint64, C.c_uint64, C.c_void_p, u64p,
                                      C.c_int, C.c_uint64, u64p, u32p, u64p]
        L.ref_sketch_many.restype = None
        L.ref_sketch_many.argtypes = [C.POINTER(Params), C.c_uint64, C.c_uint64, C.c_void_p, u64p, C.c_int, u64p, u32p]
        L.ref_hash_sequence.restype = None
        L.ref_hash_sequence.argtypes = [u64p, u32p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(Params),
                                        C.c_uint64, u64p, u32p]

    def get_hash(self, kmer: bytes, seed=42, use64=True):
        return self.lib.ref_get_hash(kmer, len(kmer), seed, int(use64))

    def sketch_unit(self, records, p, s=1000, reads=False, genome_size=0, counts=False):
        bufs, ptrs, lens = _seq_arrays(records)
        out = np.empty(s, dtype=np.uint64); cnt = np.empty(s, dtype=np.uint32)
        length = C.c_uint64(0)
        n = self.lib.ref_sketch_unit(C.byref(p), s, len(bufs), C.cast(ptrs, C.c_void_p), _ptr(lens, u64p),
                                     int(reads), genome_size, _ptr(out, u64p), _ptr(cnt, u32p), C.byref(length))
        return out[:n].copy(), (cnt[:n].copy() if counts else None), length.value

    def sketch_unit_m(self, records, p, s=1000, min_copies=1, bloom_bytes=0, reads=True, genome_size=0, counts=False):
        L = self.lib
        L.ref_sketch_unit_m.restype = C.c_uint64
        L.ref_sketch_unit_m.argtypes = [C.POINTER(Params), C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, u64p,
                                        C.c_int, C.c_uint64, u64p, u32p, u64p]
        bufs, ptrs, lens = _seq_arrays(records)
        out = np.empty(s, dtype=np.uint64); cnt = np.empty(s, dtype=np.uint32)
        length = C.c_uint64(0)
        n = L.ref_sketch_unit_m(C.byref(p), s, min_copies, bloom_bytes, len(bufs), C.cast(ptrs, C.c_void_p), _ptr(lens, u64p),
                                int(reads), genome_size, _ptr(out, u64p), _ptr(cnt, u32p), C.byref(length))
        return out[:n].copy(), (cnt[:n].copy() if counts else None), length.value

    def sketch_unit_mc(self, records, p, s=1000, min_copies=1, target_cov=0.0, genome_size=0, counts=False):
        L = self.lib
        L.ref_sketch_unit_mc.restype = C.c_uint64
        L.ref_sketch_unit_mc.argtypes = [C.POINTER(Params), C.c_uint64, C.c_uint64, C.c_double, C.c_uint64, C.c_void_p, u64p,
                                         C.c_int, C.c_uint64, u64p, u32p, u64p, u64p]
        bufs, ptrs, lens = _seq_arrays(records)
        out = np.empty(s, dtype=np.uint64); cnt = np.empty(s, dtype=np.uint32)
        length = C.c_uint64(0); used = C.c_uint64(0)
        n = L.ref_sketch_unit_mc(C.byref(p), s, min_copies, target_cov, len(bufs), C.cast(ptrs, C.c_void_p), _ptr(lens, u64p),
                                 1, genome_size, _ptr(out, u64p), _ptr(cnt, u32p), C.byref(length), C.byref(used))
        return out[:n].copy(), (cnt[:n].copy() if counts else None), length.value, used.value

    def sketch_many(self, seqs, p, s=1000, threads=1):
        bufs, ptrs, lens = _seq_arrays(seqs)
        out = np.zeros((len(bufs), s), dtype=np.uint64); out_n = np.zeros(len(bufs), dtype=np.uint32)
        self.lib.ref_sketch_many(C.byref(p), s, len(bufs), C.cast(ptrs, C.c_void_p), _ptr(lens, u64p), threads,
                                 _ptr(out, u64p), _ptr(out_n, u32p))
        return out, out_n

    def sketch_files(self, paths, p, s=1000, threads=1):
        """One sketch per FASTA/FASTQ(.gz) file through the reference's own parser (kseq.h) + hash + heap: `mash sketch -p`
        as the CPU runs it.  Returns (hashes (n x s), n_hashes, lengths)."""
        arr = (C.c_char_p * len(paths))(*[os.fsencode(q) for q in paths])
        out = np.zeros((len(paths), s), dtype=np.uint64); out_n = np.zeros(len(paths), dtype=np.uint32)
        lens = np.zeros(len(paths), dtype=np.uint64)
        self.lib.ref_sketch_files.restype = C.c_int
        rc = self.lib.ref_sketch_files(C.byref(p), C.c_uint64(s), C.c_uint64(len(paths)), arr, C.c_int(threads), _ptr(out, u64p), _ptr(out_n, u32p), _ptr(lens, u64p))
        if rc:
            raise RuntimeError("ref_sketch_files: a file could not be read")
        return out, out_n, lens

    def hash_sequence(self, keys, counts, chunk, p, s=1000):
        b = np.frombuffer(chunk, dtype=np.uint8) if isinstance(chunk, (bytes, bytearray)) else np.ascontiguousarray(chunk, np.uint8)
        out = np.empty(s, np.uint64); n = C.c_uint32(0)
        self.lib.ref_hash_sequence(_ptr(keys, u64p), _ptr(counts, u32p), keys.size, b.ctypes.data, b.size, C.byref(p), s,
                                   _ptr(out, u64p), C.byref(n))
        return out[:n.value].copy()

    # ---- screen with the reference's own table type (robin_hood map of atomics), multi-threaded ----------------------
    def screen_table(self, keys):
        L = self.lib
        L.ref_screen_table_new.restype = C.c_void_p
        L.ref_screen_table_new.argtypes = [u64p, C.c_uint64]
        L.ref_screen_table_free.argtypes = [C.c_void_p]
        L.ref_screen_table_counts.argtypes = [C.c_void_p, u64p, C.c_uint64, u32p]
        L.ref_screen_many.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, u64p, C.POINTER(Params), C.c_uint64, C.c_int, u64p, u32p]
        keys = np.ascontiguousarray(keys, np.uint64)
        return L.ref_screen_table_new(_ptr(keys, u64p), keys.size)

    def screen_table_free(self, t):
        self.lib.ref_screen_table_free(C.c_void_p(t))

    def screen_table_counts(self, t, keys):
        keys = np.ascontiguousarray(keys, np.uint64)
        out = np.zeros(max(1, keys.size), np.uint32)
        self.lib.ref_screen_table_counts(C.c_void_p(t), _ptr(keys, u64p), keys.size, _ptr(out, u32p))
        return out[:keys.size]

    def screen_many(self, t, chunks, p, s=1000, threads=1):
        """hashSequence over '*'-joined chunks on `threads` workers (one chunk = one HashInput); returns the mixture bottom-s."""
        bufs, ptrs, lens = _seq_arrays(chunks)
        out = np.empty(s, np.uint64); n = C.c_uint32(0)
        self.lib.ref_screen_many(C.c_void_p(t), len(bufs), C.cast(ptrs, C.c_void_p), _ptr(lens, u64p), C.byref(p), s, threads,
                                 _ptr(out, u64p), C.byref(n))
        return out[:n.value].copy()
