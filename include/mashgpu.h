/*
 * mashgpu.h -- C ABI of the B200-native MinHash engine (libmashgpu.so, sm_100a).
 *
 * This is the drop-in boundary for the three data-parallel hot paths of marbl/Mash.  The reference has
 * no FFI; its seams are the ThreadPool worker functions and the pure-compute functions underneath them
 * (paths below are relative to the reference's src/mash/):
 *
 *   mashgpu_sketch_batch      replaces  sketchFile / sketchSequence      Sketch.cpp:1147-1365 (Sketch.h:229-234)
 *                                       addMinHashes                     Sketch.cpp:512-583   (Sketch.h:226)
 *                                       getHash                          hash.cpp:10-38       (hash.h:21)
 *                                       MinHashHeap::tryInsert/toHashList MinHashHeap.cpp:68-146, HashSet.cpp:78-118
 *   mashgpu_dist_*            replaces  compare / compareSketches / pValue CommandDistance.cpp:306-448 (CommandDistance.h:91-93)
 *   mashgpu_screen_*          replaces  hashSequence + the shared/median/identity/p-value reduce
 *                                       CommandScreen.cpp:93-114, 484-599, 288-355, 409-455, 463-482, 601-615
 *
 * Conventions: plain pointers and sizes only; the caller owns every buffer; every function returns an int
 * status (0 = MASHGPU_OK), never throws and never calls exit() (the reference prints to cerr and exit(1)s,
 * Sketch.cpp:1294-1312 -- the host shim maps statuses back to those messages); a context is bound to one
 * CUDA device and must not be used from two threads at once.  Hashes are always carried as uint64_t, 32-bit
 * hashes (use64 == 0) zero-extended, exactly as HashList stores hash_u (HashList.h:13-37, hash.h:15-19).
 * Unlike addMinHashes (Sketch.cpp:524-530) the sequence buffers are NOT modified.
 *
 * There is no CPU fallback: without a CUDA device mashgpu_create fails.
 */
#ifndef MASHGPU_H
#define MASHGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define MASHGPU_OK 0
#define MASHGPU_ERR_INVALID 1      /* bad argument */
#define MASHGPU_ERR_CUDA 2         /* CUDA runtime error (message in mashgpu_last_error) */
#define MASHGPU_ERR_UNSUPPORTED 3  /* parameter combination outside the GPU path (see mashgpu_sketch_params) */
#define MASHGPU_ERR_NOMEM 4

typedef struct mashgpu_ctx mashgpu_ctx;

/* Sketch::Parameters (Sketch.h:34-109), the fields the hot path reads. */
typedef struct mashgpu_sketch_params {
    int32_t kmer_size;        /* kmerSize, 1..32 (Command.cpp:168) */
    uint32_t sketch_size;     /* minHashesPerWindow (s) */
    uint32_t seed;            /* hash seed (default 42) */
    int32_t use64;            /* alphabetSize^k > 2^32 (Sketch.cpp:1136); filled by mashgpu_set_alphabet */
    int32_t noncanonical;
    int32_t preserve_case;
    uint8_t alphabet[256];    /* Parameters::alphabet; filled by mashgpu_set_alphabet */
    uint32_t min_copies;      /* `-m`: MinHashHeap multiplicityMinimum (Sketch.cpp:1186, reads mode; MinHashHeap.cpp:96-118): a hash enters
                                 the sketch at its m-th occurrence.  0 or 1 = off.  The sketch is then the s smallest hashes seen at
                                 least m times, and the multiplicities follow the heap exactly (tests/test_gpu_sketch.py). */
    double target_cov;        /* `-c`: stop after the first read that brings the heap's average multiplicity to this value
                                 (Sketch.cpp:1258-1262).  Only mashgpu_sketch_reads looks at it; <= 0 = off. */
} mashgpu_sketch_params;
/* GPU path coverage: alphabet == {A,C,G,T} (canonical or not, any k 1..32, either case mode) runs the 4-bit
 * packed DNA kernels; any other alphabet requires noncanonical != 0 (the protein setting of the reference,
 * sketchParameterSetup.cpp) and runs the byte-alphabet kernels.  Other combinations return
 * MASHGPU_ERR_UNSUPPORTED (never a CPU fallback). */

/* setAlphabetFromString (Sketch.cpp:1108-1137): fills alphabet[] and use64 from kmer_size/preserve_case.
 * Returns the alphabet size. */
uint32_t mashgpu_set_alphabet(mashgpu_sketch_params *p, const char *characters);

int mashgpu_device_count(void);
int mashgpu_create(int device, mashgpu_ctx **ctx);
void mashgpu_destroy(mashgpu_ctx *ctx);
/* Message of the last failure on this context (ctx == NULL: of the last failed mashgpu_create). */
const char *mashgpu_last_error(const mashgpu_ctx *ctx);

/* ---------------------------------------------------------------------------------------------------------
 * Hot path 1: sketching.
 *
 * n_records pre-parsed records (what kseq_read delivers: sequence bytes, kseq.h:170-208) are grouped into
 * n_units sketch units: unit_of_record[r] (non-decreasing; NULL = record r is unit r).  A unit is one
 * sketchFile job (all records of the listed files, k-mers never span records) or one sketchSequence job.
 * Records shorter than kmer_size are skipped and not counted in the length (Sketch.cpp:1222-1226,1251-1254).
 *
 * Outputs, per unit u, in input order (ThreadPool ordering contract, ThreadPool.hxx:87-118):
 *   out_hashes[u*sketch_size .. +out_n[u])  the min(s, #distinct) smallest distinct hashes, ascending
 *   out_counts (nullable)                    multiplicity of each hash (HashSet counts)
 *   out_n[u]                                 number of hashes
 *   out_length[u]                            sum of the kept records' lengths (Reference::length, non -r mode)
 * ------------------------------------------------------------------------------------------------------- */
int mashgpu_sketch_batch(mashgpu_ctx *ctx, const mashgpu_sketch_params *params,
                         uint64_t n_records, const char *const *seq, const uint64_t *len,
                         const uint32_t *unit_of_record, uint64_t n_units,
                         uint64_t *out_hashes, uint32_t *out_counts, uint32_t *out_n, uint64_t *out_length);

/* One sketch of a read set, `mash sketch -r [-m min_copies] [-c target_cov]` (sketchFile with parameters.reads, Sketch.cpp:1186-1282):
 * all records in the order given (the caller interleaves several files round robin as Sketch.cpp:1202-1270 does).  With
 * target_cov > 0 the result is the heap as it stood after the first kept record that brought estimateMultiplicity() to the target
 * (Sketch.cpp:1258-1262) and *out_records_used (the "Reads used" line, :1324-1327) is the number of kept records up to and
 * including it; otherwise all kept records are used.  The stop depends on the order of the reads; it is found exactly: the
 * exact top of the heap at geometrically spaced prefixes bounds which k-mers can pass the heap's gate afterwards, those
 * k-mers are collected as events and replayed in stream order through the reference's heap logic on the device
 * (DESIGN.md 2.6).  out_counts nullable.  DNA alphabet only. */
int mashgpu_sketch_reads(mashgpu_ctx *ctx, const mashgpu_sketch_params *params,
                         uint64_t n_records, const char *const *seq, const uint64_t *len,
                         uint64_t *out_hashes, uint32_t *out_counts, uint32_t *out_n, uint64_t *out_records_used);

/* Same computation on a sequence stream already resident in device memory.  d_stream holds the units back to
 * back; unit u occupies bytes [unit_start[u], unit_start[u+1]) (host array of n_units+1 offsets); records inside
 * a unit are separated by at least one byte outside the alphabet (e.g. 0).  d_stream must be readable up to
 * unit_start[n_units] rounded up to 16 bytes.  d_out_* are device pointers (d_out_counts nullable).
 * `stream` is a cudaStream_t (NULL = the context's stream).  All work is enqueued on that stream; the call synchronises
 * it once at the end (it has to read the per-unit status flags to decide whether a unit needs the exact re-run). */
int mashgpu_sketch_stream_dev(mashgpu_ctx *ctx, const mashgpu_sketch_params *params,
                              const void *d_stream, const uint64_t *unit_start, uint64_t n_units,
                              uint64_t *d_out_hashes, uint32_t *d_out_counts, uint32_t *d_out_n, void *stream);

/* Host feed path, exposed for tests (no GPU involved): packs records the way mashgpu_sketch_batch does before the
 * H2D copy -- every record followed by one separator position -- into 2-bit codes (A0 C1 G2 T3, 32 positions per
 * uint64 word) plus the list of runs of positions that are not in the alphabet (upper-casing unless preserve_case;
 * separators included).  codes must hold ceil(stream_len/32) words with stream_len = sum(len[r] + 1); runs receives
 * up to runs_capacity {start, length} pairs; *n_runs is the number of runs found. */
int mashgpu_host_pack(const mashgpu_sketch_params *params, uint64_t n_records, const char *const *seq, const uint64_t *len,
                      int threads, uint64_t *codes, uint64_t *runs, uint64_t runs_capacity, uint64_t *n_runs);

/* Sketching from a stream the caller already keeps 2-bit packed (the format mashgpu_host_pack writes): a collection cached in
 * this form crosses PCIe at a quarter of a byte per base and no host thread touches the bases again, so the call runs at the
 * rate of the scan kernel instead of the rate of PCIe-ASCII.  codes: ceil(stream_len / 32) words, base p at bits 2 (p % 32) of
 * codes[p / 32] (A0 C1 G2 T3); runs: n_runs {start, length} pairs, ascending and disjoint, covering every position that is not
 * a base of the alphabet (the separator after each record included); unit u = positions [unit_start[u], unit_start[u+1])
 * (n_units + 1 ascending offsets, unit_start[n_units] <= stream_len).  Same outputs as mashgpu_sketch_batch (without
 * out_length: record lengths are the caller's).  Pinned host buffers make the copies asynchronous.  DNA alphabet only. */
int mashgpu_sketch_batch_packed(mashgpu_ctx *ctx, const mashgpu_sketch_params *params,
                                const uint64_t *codes, uint64_t stream_len, const uint64_t *runs, uint64_t n_runs,
                                const uint64_t *unit_start, uint64_t n_units,
                                uint64_t *out_hashes, uint32_t *out_counts, uint32_t *out_n);

/* getHash over every window (hash.cpp:10-38 applied as in Sketch.cpp:540-576): out_hash[i]/out_valid[i] for each
 * window start i in [0, len-k]; invalid windows (a byte outside the alphabet) have out_valid[i] == 0.
 * Host buffers.  Diagnostic/test entry point of the scan+hash kernel. */
int mashgpu_hash_windows(mashgpu_ctx *ctx, const mashgpu_sketch_params *params,
                         const char *seq, uint64_t len, uint64_t *out_hash, uint8_t *out_valid);

/* ---------------------------------------------------------------------------------------------------------
 * Hot path 2: all-pairs sketch comparison.
 * ------------------------------------------------------------------------------------------------------- */
/* A Sketch's reference list (Sketch.h:131-139): n rows of `stride` uint64 slots; row i holds n_hashes[i]
 * ascending distinct hashes (Reference::hashesSorted) and has sequence length length[i] (Reference::length).
 * on_device != 0: hashes/n_hashes/length are device pointers. */
typedef struct mashgpu_sketch_set {
    uint64_t n;
    uint64_t stride;
    const uint64_t *hashes;
    const uint32_t *n_hashes;
    const uint64_t *length;
    int32_t on_device;
} mashgpu_sketch_set;

/* Arguments of compareSketches (CommandDistance.h:92). */
typedef struct mashgpu_dist_params {
    uint64_t sketch_size;   /* min(s_query, s_ref)  (CommandDistance.cpp:313-315) */
    int32_t kmer_size;
    double kmer_space;      /* alphabetSize^k as double (Sketch.cpp:509) */
    double max_distance;    /* -d, < 0 disables (CommandDistance.cpp:409) */
    double max_pvalue;      /* -v, < 0 disables (CommandDistance.cpp:419) */
} mashgpu_dist_params;

typedef struct mashgpu_dist_job mashgpu_dist_job;

/* Uploads both sets (qry == NULL or qry == ref: all-vs-all of one set), builds the order-preserving 32-bit
 * dictionary of all hashes and keeps everything resident. */
int mashgpu_dist_open(mashgpu_ctx *ctx, const mashgpu_sketch_set *ref, const mashgpu_sketch_set *qry,
                      const mashgpu_dist_params *params, mashgpu_dist_job **job);

/* Compares queries [q_begin, q_begin+q_count) with every reference: pair (q, r) is written at index
 * (q - q_begin) * n_ref + r, the reference's query-major order (CommandDistance.cpp:213-232, 306-334).
 * Per pair (PairOutput, CommandDistance.h:63-70): numer (shared hashes), denom, distance, pvalue, pass.
 * Sketch sizes up to 1035 run the tiled lockstep merge (32 references per shared-memory tile); larger ones (`-s 10000`) a
 * warp-per-pair merge; beyond ~28 000 hashes per sketch one thread per pair from global memory.
 * Where the reference leaves numer/denom/distance/pValue unset (distance > max_distance, :409-412) this
 * engine still writes the computed numer/denom/distance and pvalue = 0, with pass = 0.
 * Any output pointer may be NULL.  Host buffers. */
int mashgpu_dist_run(mashgpu_dist_job *job, uint64_t q_begin, uint64_t q_count,
                     uint32_t *numer, uint32_t *denom, double *distance, double *pvalue, uint8_t *pass);
/* Same with device output buffers, enqueued on `stream` (cudaStream_t, NULL = context stream).  A job holds one set of
 * scratch buffers (work lists, deferred p-value queue, pass list): at most ONE run of a job may be in flight -- enqueue the
 * runs of a job on a single stream, or synchronise between runs on different streams. */
int mashgpu_dist_run_dev(mashgpu_dist_job *job, uint64_t q_begin, uint64_t q_count,
                         uint32_t *d_numer, uint32_t *d_denom, double *d_distance, double *d_pvalue, uint8_t *d_pass,
                         void *stream);
/* Filtered runs (-d / -v): only the pairs with pass == 1, which is all that writeOutput prints
 * (CommandDistance.cpp:270-287), as a compact list sorted by pair index (q - q_begin) * n_ref + r, i.e. in the
 * reference's output order.  *n_pass = number of passing pairs; if it exceeds `capacity` nothing is returned and the
 * caller retries with a larger capacity (or uses mashgpu_dist_run).  Host buffers of `capacity` entries. */
int mashgpu_dist_run_list(mashgpu_dist_job *job, uint64_t q_begin, uint64_t q_count, uint64_t capacity,
                          uint64_t *pair_index, uint32_t *numer, uint32_t *denom, double *distance, double *pvalue, uint64_t *n_pass);
int mashgpu_dist_close(mashgpu_dist_job *job);

/* Lower-triangular enumeration for a self comparison (qry == NULL): `mash triangle` compares row i with rows 0..i-1 only
 * (CommandTriangle.cpp:200-214).  on != 0: pairs with r >= q are neither computed nor written and never enter a pass list;
 * whole reference tiles above the diagonal are skipped.  Their output slots are zero with mashgpu_dist_run (host buffers)
 * and keep whatever they held with mashgpu_dist_run_dev (device buffers). */
int mashgpu_dist_set_triangle(mashgpu_dist_job *job, int on);

/* Tile prefilter of the merge (no counterpart in the reference, which merges every pair, CommandDistance.cpp:347-365;
 * the results are identical).  Before a tile of 32 references is merged with a query, the query's hashes are looked up in
 * a filter built over the tile's hashes; a query that shares no hash with any of the 32 references gets the closed form
 * of an empty intersection (common 0, denom min(s, |A|+|B|), distance 1, p-value 1) for all 32 pairs and is not merged.
 * mode: -1 auto (default: on, and switched off for the rest of the job once more than half of the probed
 * (query, tile) combinations turned out to share hashes), 0 off, 1 on.  Environment MASHGPU_DIST_PREFILTER=0|1
 * overrides the default at mashgpu_dist_open. */
int mashgpu_dist_set_prefilter(mashgpu_dist_job *job, int mode);
/* Counters of the prefilter since the job was opened: (query, tile) combinations probed, those that went on to the merge,
 * and whether the next run would probe.  Synchronises the device.  Any pointer may be NULL. */
int mashgpu_dist_prefilter_stats(mashgpu_dist_job *job, uint64_t *combos_probed, uint64_t *combos_flagged, int *active);
/* The prefilter also learns WHICH references of a tile a query may share a hash with (one owner byte per filter slot).  A
 * combination with at most 4 candidate references (MASHGPU_DIST_PAIR_MAX, 0 = off) is not merged as a whole: its candidate
 * pairs go to a list and are merged one warp per pair, the other references of the tile get the closed form.  This keeps the
 * run fast when related sketches are scattered over the collection instead of stored next to each other.
 * *pairs_merged_from_lists: pairs that took this path since the job was opened.  Synchronises the device. */
int mashgpu_dist_pair_stats(mashgpu_dist_job *job, uint64_t *pairs_merged_from_lists);

/* ---- Sharded dictionary build (multi-GPU all-vs-all).  With the reference axis sharded over G ranks the
 * dictionary of mashgpu_dist_open would be rebuilt in full on every rank.  These entry points split it by HASH RANGE
 * instead (a sample sort): a rank sorts only its own rows' hashes, the sorted keys are cut at G-1 splitters, every rank
 * receives one hash range from all ranks (the caller's all-to-all, NCCL), ranks the distinct values of its range, and the
 * 32-bit codes travel back; the owners scatter them into rows of sketch_size+1 codes which are then all-gathered (4 B per
 * hash instead of 8) and opened with mashgpu_dist_open_encoded.  Codes of range d are offset by the number of distinct
 * values in the ranges below it, so the encoding is order- and equality-preserving over the whole collection, which is
 * all the merge of compareSketches (CommandDistance.cpp:347-365) depends on.  All pointers named d_* are device memory;
 * `stream` is a cudaStream_t (NULL = the context's stream); the calls that return a host value synchronise it. */
/* Hashes i < min(n_hashes, sketch_size+1, stride) of every row, ascending, with their slot row * (sketch_size+1) + i.
 * d_keys / d_slots must hold set->n * min(stride, sketch_size+1) entries; *n_valid = number of entries written. */
int mashgpu_dict_local_sort(mashgpu_ctx *ctx, const mashgpu_sketch_set *set, uint64_t sketch_size,
                            uint64_t *d_keys, uint32_t *d_slots, uint64_t *n_valid, void *stream);
/* counts[p] = number of keys of the ascending list d_keys[0..n) in part p of n_parts: part p holds
 * splitters[p-1] <= key < splitters[p] (splitters: host array of n_parts-1 ascending values). */
int mashgpu_dict_split(mashgpu_ctx *ctx, const uint64_t *d_keys, uint64_t n, const uint64_t *splitters, uint32_t n_parts,
                       uint64_t *counts, void *stream);
/* d_codes[t] = number of distinct keys smaller than d_keys[t] (any order of d_keys; equal keys get equal codes);
 * *n_distinct = number of distinct keys. */
int mashgpu_dict_rank(mashgpu_ctx *ctx, const uint64_t *d_keys, uint64_t n, uint32_t *d_codes, uint64_t *n_distinct, void *stream);
/* Rows of sketch_size+1 codes from the codes of this rank's sorted keys: entry t (in the order of mashgpu_dict_local_sort)
 * belongs to segment g = the part it was sent to (seg_counts[g] consecutive entries, host arrays of n_segs) and becomes
 * d_rows[d_slots[t]] = d_codes[t] + seg_base[g]; every other position of the n_rows rows is the padding code 0xFFFFFFFF.
 * d_n_eff[row] = min(n_hashes, sketch_size+1, stride).  Fails when a code would reach the padding value. */
int mashgpu_dict_scatter(mashgpu_ctx *ctx, const uint32_t *d_codes, const uint32_t *d_slots, const uint64_t *seg_counts,
                         const uint64_t *seg_base, uint32_t n_segs, const mashgpu_sketch_set *set, uint64_t sketch_size,
                         uint32_t *d_rows, uint32_t *d_n_eff, void *stream);
/* A job over rows that are already dictionary-encoded (device memory, owned by the caller, must outlive the job):
 * n_rows rows of sketch_size+1 codes, d_n_eff / d_length per row.  The queries are all n_rows rows, the references rows
 * [ref_begin, ref_begin + ref_count): pair (q, r) of a run is query row q against row ref_begin + r, written at
 * (q - q_begin) * ref_count + r.  mashgpu_dist_set_triangle compares q with rows below it only (ref_begin + r < q). */
int mashgpu_dist_open_encoded(mashgpu_ctx *ctx, const uint32_t *d_rows, const uint32_t *d_n_eff, const uint64_t *d_length,
                              uint64_t n_rows, uint64_t ref_begin, uint64_t ref_count, const mashgpu_dist_params *params,
                              mashgpu_dist_job **job);

/* One-shot convenience: open + run over all queries + close (the whole `compare` grid). */
int mashgpu_dist(mashgpu_ctx *ctx, const mashgpu_sketch_set *ref, const mashgpu_sketch_set *qry,
                 const mashgpu_dist_params *params,
                 uint32_t *numer, uint32_t *denom, double *distance, double *pvalue, uint8_t *pass);

/* ---------------------------------------------------------------------------------------------------------
 * Hot path 3: screen (containment of reference sketches in a read stream).
 * ------------------------------------------------------------------------------------------------------- */
typedef struct mashgpu_screen_job mashgpu_screen_job;

/* Builds the distinct-hash table of the reference sketches with zeroed counters (CommandScreen.cpp:93-114). */
int mashgpu_screen_open(mashgpu_ctx *ctx, const mashgpu_sketch_params *params, const mashgpu_sketch_set *refs,
                        mashgpu_screen_job **job);
/* One HashInput (CommandScreen.h:104-133): a '*'-joined chunk of reads (CommandScreen.cpp:224-262; any byte
 * outside the alphabet separates reads).  Host buffer; chunks may be of any size: chunks below 4 MiB are joined before a kernel
 * pass, larger ones are pipelined (the copy of chunk i+1 overlaps the kernels of chunk i).  The call returns as soon as the
 * caller's buffer may be reused; results are complete at mashgpu_screen_finish.  (MASHGPU_SCREEN_HOST_PACK=1: the chunk is 2-bit
 * packed on the host threads and uploaded with an invalid-position mask instead of as ASCII.) */
int mashgpu_screen_feed(mashgpu_screen_job *job, const char *chunk, uint64_t len);
/* Same for a chunk already in device memory (readable up to len rounded up to 16). */
int mashgpu_screen_feed_dev(mashgpu_screen_job *job, const void *d_chunk, uint64_t len);
/* Reduce (CommandScreen.cpp:288-355, 409-455): per reference sketch i the number of its hashes seen at least
 * once (shared), the median multiplicity depths[shared/2], identity (estimateIdentity, :463-482) and p-value
 * (pValueWithin, :601-615); *set_size = (uint64_t)estimateSetSize() of the mixture's bottom-s heap (:322).
 * mixture_hashes (nullable, sketch_size slots) / mixture_n receive that bottom-s list.  Host buffers. */
int mashgpu_screen_finish(mashgpu_screen_job *job, uint64_t *shared, uint64_t *median, double *identity,
                          double *pvalue, uint64_t *set_size, uint64_t *mixture_hashes, uint32_t *mixture_n);
/* `mash screen -w` ("winner take all", CommandScreen.cpp:357-407): on != 0 makes mashgpu_screen_finish re-assign every
 * reference hash seen in the mixture to the one sketch, among those containing it, with the highest identity estimate
 * (ties: the longer genome, refs->length at mashgpu_screen_open; a tie in both goes to the lowest index -- the reference
 * leaves it to the iteration order of a hash set) before shared / median / identity / p-value are computed. */
int mashgpu_screen_set_winner(mashgpu_screen_job *job, int on);
/* Multi-GPU screening (reads sharded over ranks, table replicated): the hit counters live in device memory as one
 * uint32 per distinct reference hash, in ascending hash order, so counter i means the same hash on every rank that
 * opened the job with the same sketches.  *d_counters / *n_counters expose that array for an all-reduce(sum) (NCCL via torch.distributed); the
 * mixture bottom-s lists of the other ranks are folded in with mashgpu_screen_merge_mixture (ascending distinct hashes,
 * host buffer) -- bottom-s of a union is the bottom-s of the union of bottom-s lists (the reference merges its
 * per-thread heaps the same way, CommandScreen.cpp:288-302).  Then mashgpu_screen_finish as usual. */
int mashgpu_screen_counters(mashgpu_screen_job *job, uint32_t **d_counters, uint64_t *n_counters);
int mashgpu_screen_merge_mixture(mashgpu_screen_job *job, const uint64_t *hashes, uint32_t n);
/* The same without a host round trip per rank: *d_mix / *d_mix_n expose the running bottom-s list (sketch_size uint64 slots,
 * ascending, *d_mix_n of them valid) for an all-gather; mashgpu_screen_merge_mixtures_dev folds n_lists lists lying in
 * device memory (list i = d_hashes[i * stride ...], d_n[i] ascending distinct hashes; the job's own list may be among
 * them) into the job's list with one kernel and one read-back of the new top. */
int mashgpu_screen_mixture_dev(mashgpu_screen_job *job, uint64_t **d_mix, uint32_t **d_mix_n);
int mashgpu_screen_merge_mixtures_dev(mashgpu_screen_job *job, const uint64_t *d_hashes, const uint32_t *d_n, uint32_t n_lists, uint64_t stride);
int mashgpu_screen_close(mashgpu_screen_job *job);

/* ---------------------------------------------------------------------------------------------------------
 * Instrumentation (bench.py): kernel launch counter and device time of the dominant kernels.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct mashgpu_stats {
    uint64_t kernel_launches;      /* kernels launched by this context since the last reset */
    double scan_kernel_ms;         /* accumulated device time of the sketch/screen scan kernel (CUDA events) */
    uint64_t scan_kernel_launches;
    double dist_kernel_ms;         /* accumulated device time of the merge kernel */
    uint64_t dist_kernel_launches;
    uint64_t exact_reruns;         /* units that needed the exact re-run path */
} mashgpu_stats;
/* timing != 0 brackets the dominant kernels with CUDA events (adds a sync at mashgpu_get_stats). */
int mashgpu_set_timing(mashgpu_ctx *ctx, int timing);
int mashgpu_get_stats(mashgpu_ctx *ctx, mashgpu_stats *out, int reset);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* MASHGPU_H */
