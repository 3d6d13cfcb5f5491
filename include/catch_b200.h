/*
 * catch_b200.h -- C ABI of libcatchb200.so, the sm_100a implementation of the CATCH
 * probe-coverage + set-cover + near-duplicate hot path.
 *
 * The reference (broadinstitute/catch) is pure Python and has no FFI; this header is the
 * boundary a maintainer would bind with ctypes from the reference's filter classes
 * (see INTEGRATION.md).  Each entry point names the reference code it replaces
 * (paths relative to the reference tree).
 *
 * Conventions: every function returns 0 on success and a negative cb_status on error;
 * cb_last_error(ctx) gives the message.  Handles are opaque and released with the matching
 * *_free call.  Calls are synchronous (they return after the context's stream has been
 * synchronised) and a cb_ctx must only be used from one host thread at a time.
 * Host pointers are plain caller-owned memory; nothing here takes or returns torch types.
 */
#ifndef CATCH_B200_H
#define CATCH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cb_ctx cb_ctx;
typedef struct cb_targets cb_targets;
typedef struct cb_probes cb_probes;
typedef struct cb_cover cb_cover;

enum cb_status {
    CB_OK = 0,
    CB_ERR_CUDA = -1,        /* a CUDA runtime call failed */
    CB_ERR_ARG = -2,         /* invalid argument */
    CB_ERR_UNSUPPORTED = -3, /* outside the supported envelope */
    CB_ERR_NOMEM = -4,
    CB_ERR_STATE = -5,       /* e.g. set cover could not reach the requested coverage */
    CB_ERR_COMM = -6         /* NCCL failure */
};

/* Limits of the device kernels. */
#define CB_MAX_PROBE_LEN 256     /* bases per probe */
#define CB_MAX_SYMBOL_BITS 8     /* bit planes per base */
#define CB_MAX_MISMATCHES 31
#define CB_MAX_RANKS 8           /* GPUs of one box that can share a sharded set cover */

/* Hybridisation model parameters: probe.py:1274-1346
 * probe_covers_sequence_by_longest_common_substring(mismatches, lcf_thres, island). */
typedef struct {
    int32_t mismatches;
    int32_t lcf_thres;
    int32_t island_of_exact_match;
    int32_t cover_extension;     /* filter/set_cover_filter.py:429-432 */
    int32_t k;                   /* seed (k-mer) length of the probe map, probe.py:507-577 */
} cb_hyb_params;

/* Per-call timings (CUDA events on the context's stream, milliseconds) and counters. */
typedef struct {
    double ms_h2d;               /* host->device copies inside the call */
    double ms_pack;              /* ASCII -> bit-plane packing */
    double ms_seed_index;        /* K2 */
    double ms_scan_count;        /* K3, counting pass */
    double ms_scan_emit;         /* K3, emitting pass */
    double ms_merge;             /* K4 */
    double ms_universe;          /* K5 + gain initialisation */
    double ms_greedy;            /* K6-K8 persistent greedy kernel */
    double ms_d2h;
    double ms_total;             /* wall time of the call on the device timeline */
    int64_t n_seed_entries;
    int64_t n_seed_lookups;      /* target positions looked up */
    int64_t n_candidate_hits;    /* bucket entries examined */
    int64_t n_raw_ranges;        /* cover ranges emitted before merging */
    int64_t n_intervals;         /* merged (probe, genome) intervals */
    int64_t n_picks;
    int64_t n_kernel_launches;
    int64_t bytes_algorithmic;   /* DESIGN.md formula, for the roofline */
    int64_t reserved[8];
} cb_stats;

/* ---- context --------------------------------------------------------------------- */
int cb_init(int device_id, cb_ctx **out);
void cb_destroy(cb_ctx *ctx);
const char *cb_last_error(cb_ctx *ctx);
/* Library build identification ("catch_b200 <version> sm_100a"). */
const char *cb_version(void);
/* Benchmark hygiene: overwrite a buffer larger than the L2 cache (256 MiB) so the next call
 * starts with a cold L2. */
int cb_flush_l2(cb_ctx *ctx);
/* Grow the device's stream-ordered memory pool so that it holds at least `bytes` of free memory: later
 * calls then take their work buffers from the pool without going to the driver (which may wait for
 * running kernels).  Optional; useful before the first call and wherever several contexts share a device. */
int cb_pool_reserve(cb_ctx *ctx, int64_t bytes);
/* Benchmark calibration: the rate (32-bit integer operations per second, whole chip) at which this GPU
 * executes independent chains of LOP3 / SHF / IADD3, the instruction mix the scan kernel is bound by.
 * The denominator of the integer-op roofline bench.py reports beside the HBM one. */
int cb_intop_rate(cb_ctx *ctx, double *ops_per_s);

/* Page-locked host staging memory owned by the context: slot 0..5 (0/1: probes and targets of the call in progress; 2/3 and 4/5: two more
 * pairs, so that the next groupings can be gathered while the current one is on the device), at least `bytes` bytes, valid
 * until the next cb_host_buffer call for the same slot with a larger size (or cb_destroy).  Filling
 * the sequences straight into it lets the uploads below run as true asynchronous DMA instead of
 * going through the driver's bounce buffer. */
int cb_host_buffer(cb_ctx *ctx, int32_t slot, int64_t bytes, void **out);

/* ---- packing (K1) ------------------------------------------------------------------
 * Replaces the per-hit str -> np.array('U1') conversions of probe.py:1074 and
 * Probe.from_str (probe.py:344): sequences go to the device once, as `bits` bit planes
 * per base.  `lut` maps every input byte to a code < 2^bits; two bases match iff their
 * codes are equal (plain character equality, utils/longest_common_substring.py:110, so the
 * host must give distinct codes to distinct characters, 'N' included).
 *
 * Targets: `n_seqs` sequences concatenated in `ascii`, sequence i at
 * [seq_off[i], seq_off[i+1]).  seq_genome[i] (non-decreasing, 0..n_genomes-1) is the genome
 * ("universe", filter/set_cover_filter.py:414-416) the sequence belongs to. */
int cb_upload_targets(cb_ctx *ctx, const uint8_t *ascii, const int64_t *seq_off, int64_t n_seqs,
                      const int32_t *seq_genome, int32_t n_genomes, const uint8_t lut[256],
                      int32_t bits, cb_targets **out, cb_stats *stats);
void cb_targets_free(cb_targets *t);

/* Probes: probe i is ascii[probe_off[i] .. probe_off[i+1]), at most CB_MAX_PROBE_LEN. */
int cb_upload_probes(cb_ctx *ctx, const uint8_t *ascii, const int64_t *probe_off, int64_t n_probes,
                     const uint8_t lut[256], int32_t bits, cb_probes **out, cb_stats *stats);
void cb_probes_free(cb_probes *p);

/* Both uploads of one grouping in a single call, with the code table derived on the device: the
 * distinct bytes of the two buffers (A, C, G, T always included) get dense codes in byte order,
 * *bits_out = ceil(log2(#symbols)) planes.  This is what SetCoverFilter._filter does per grouping
 * before _make_sets (filter/set_cover_filter.py:359-470 receives str sequences; here they are
 * packed once).
 * Probes come either with explicit offsets (probe_off != NULL, probes_bytes/sep ignored) or as ONE
 * buffer of probes_bytes bytes in which consecutive probes are separated by a single `sep` byte
 * (what '\n'.join(...) produces; no trailing separator); the offsets are then found here, which
 * spares the host a per-probe length pass.  CB_ERR_ARG if the number of separators is not
 * n_probes - 1.  probe_len_out (nullable) receives the n_probes probe lengths. */
int cb_upload_group(cb_ctx *ctx, const uint8_t *probes_ascii, int64_t probes_bytes, const int64_t *probe_off,
                    int64_t n_probes, int32_t sep, const uint8_t *targets_ascii, const int64_t *seq_off,
                    int64_t n_seqs, const int32_t *seq_genome, int32_t n_genomes, int32_t *probe_len_out,
                    int32_t *bits_out, cb_probes **probes_out, cb_targets **targets_out, cb_stats *stats);

/* The same on a native worker thread: _begin returns immediately, _end waits for the draw and
 * returns its status.  key/pos/out must stay valid and untouched in between.  Lets the caller
 * overlap the (host-only) seed draw with cb_upload_group. */
typedef struct cb_rng_job cb_rng_job;
cb_rng_job *cb_mt19937_randint_begin(uint32_t *key, int32_t *pos, uint32_t bound, int64_t n, void *out,
                                     int32_t elem_size /* 4: int32 output, 1: uint8 output */);
int cb_mt19937_randint_end(cb_rng_job *job);

/* Host-side helper (no device work): lengths of the n strings held in `buf`, consecutive strings
 * separated by one `sep` byte (see cb_upload_group).  CB_ERR_ARG unless exactly n-1 separators occur. */
int cb_split_lengths(const uint8_t *buf, int64_t bytes, int64_t n, int32_t sep, int32_t *len_out);

/* 1 in *has_dup if two probes may have the same sequence (decided from a 64-bit hash of the
 * packed probe: equal sequences always report 1, distinct ones almost never).  Lets the host skip
 * its duplicate bookkeeping (filter/set_cover_filter.py:408-412) in the common duplicate-free case. */
int cb_probes_have_duplicates(cb_ctx *ctx, const cb_probes *probes, int32_t *has_dup);

/* Host-side helper (no device work): continue numpy's legacy MT19937 stream exactly as
 * RandomState.randint(0, bound, size=n) / np.random.choice(bound, n) would (masked rejection
 * sampling on 32-bit outputs, numpy/random/src/distributions/distributions.c), so the seed draws of
 * probe.py:393-396 can be replayed without per-probe Python overhead.  key[624]/pos are the state
 * from np.random.get_state() and are updated in place. */
int cb_mt19937_randint(uint32_t *key, int32_t *pos, uint32_t bound, int64_t n, int32_t *out);
/* The portable (non-SIMD) loop of the same function: identical output; exported so that both code
 * paths can be compared on one machine. */
int cb_mt19937_randint_scalar(uint32_t *key, int32_t *pos, uint32_t bound, int64_t n, int32_t *out);
/* Same draws stored as bytes (bound <= 256): the form cb_coverage_uniform takes. */
int cb_mt19937_randint_u8(uint32_t *key, int32_t *pos, uint32_t bound, int64_t n, uint8_t *out);

/* ---- stage A: coverage (K2-K4) -----------------------------------------------------
 * Replaces SetCoverFilter._make_sets (filter/set_cover_filter.py:359-470), i.e.
 * probe.SharedKmerProbeMap.construct (probe.py:684-763) + open_probe_finding_pool +
 * find_probe_covers_in_sequence over every target sequence (probe.py:1008-1271) with the
 * predicate of probe.py:1328-1344, followed by the +-cover_extension / clip / genome-offset
 * step (:429-439) and interval.IntervalSet merging (:462-466).
 *
 * seed_off/seed_pos: CSR over probes of the selected seed start positions.  Set semantics, as in
 * the reference (probe.py:398): a position may repeat and the order is free; every position must
 * satisfy pos + k <= len(probe).  The host generates them (pigeonhole rule or a replay of
 * numpy's legacy RNG, probe.py:356-504) because they depend on host RNG state. */
int cb_coverage(cb_ctx *ctx, const cb_probes *probes, const cb_targets *targets,
                const cb_hyb_params *params, const int64_t *seed_off, const int32_t *seed_pos,
                cb_cover **out, cb_stats *stats);
/* Same as cb_coverage when every probe has the same number of seed positions (the reference's
 * random mode draws 20 per probe, probe.py:393-396; pigeonhole mode gives L/k per probe):
 * seed_pos is a dense [n_probes][seeds_per_probe] byte matrix, no offsets, no host-side narrowing. */
int cb_coverage_uniform(cb_ctx *ctx, const cb_probes *probes, const cb_targets *targets,
                        const cb_hyb_params *params, const uint8_t *seed_pos, int32_t seeds_per_probe,
                        cb_cover **out, cb_stats *stats);
/* The general form, used by the probe-sharded multi-GPU path: exactly one of (seed_off, seed_pos) and
 * seed_pos_u8 is given, and only the probes [probe_lo, probe_hi) are scanned (probe_hi < 0: all).  The
 * uniform byte matrix then holds probe_hi - probe_lo rows, the first one for probe_lo; the CSR form keeps
 * its n_probes + 1 offsets.  The cover keeps GLOBAL probe ids: rows outside the range are empty. */
int cb_coverage_range(cb_ctx *ctx, const cb_probes *probes, const cb_targets *targets,
                      const cb_hyb_params *params, const int64_t *seed_off, const int32_t *seed_pos,
                      const uint8_t *seed_pos_u8, int32_t seeds_per_probe, int64_t probe_lo, int64_t probe_hi,
                      cb_cover **out, cb_stats *stats);
/* The ranges the scan emits, one record per (probe, diagonal, mismatch-free run that holds a selected seed)
 * that passes the predicate, NOT merged and NOT extended by cover_extension (params->cover_extension must be
 * 0): five uint32 per record = probe index, sequence index, start and end inside that sequence
 * (probe.py:1095-1106 cover_start / cover_end), and the sequence position of the seed hit that produced it
 * (probe.py:1062 `i`); records come in no particular order.  This is what find_probe_covers_in_sequence(merge_overlapping=False)
 * consumers (coverage_analysis.py:228-231, `sorted(set(...))` per probe and sequence) and the adapter
 * filter's interval scheduling (adapter_filter.py:191-238, which also needs the order in which probes are
 * first found) are built on.  *records is allocated by the library (cb_free_host), *n_records records. */
int cb_coverage_records(cb_ctx *ctx, const cb_probes *probes, const cb_targets *targets,
                        const cb_hyb_params *params, const int64_t *seed_off, const int32_t *seed_pos,
                        int64_t *n_records, uint32_t **records, cb_stats *stats);
void cb_free_host(void *p);
void cb_cover_free(cb_cover *c);

/* Number of merged (probe, genome, start, end) intervals held by a cover. */
int64_t cb_cover_num_intervals(const cb_cover *c);
/* Copy them out, sorted by (probe, genome, start); start/end are genome coordinates
 * (positions of all sequences of the genome laid end to end, :438-439).  Arrays are
 * caller-allocated with cb_cover_num_intervals() entries. */
int cb_cover_export(cb_ctx *ctx, const cb_cover *c, int64_t *probe_id, int32_t *genome,
                    int64_t *start, int64_t *end);

/* Build a cover from host-side intervals instead of cb_coverage (the `sets` argument of
 * set_cover.approx_multiuniverse, utils/set_cover.py:147, in flat form): interval i says probe
 * probe_id[i] covers [start[i], end[i]) of genome genome[i].  Intervals may come in any order and
 * may overlap; they are merged per (probe, genome) like interval.IntervalSet does.
 * genome_len[g] bounds the coordinates of genome g. */
int cb_cover_import(cb_ctx *ctx, int64_t n_probes, int32_t n_genomes, const int64_t *genome_len,
                    int64_t n_intervals, const int64_t *probe_id, const int32_t *genome,
                    const int64_t *start, const int64_t *end, cb_cover **out);

/* ---- stage B: greedy multi-universe set cover (K5-K8) -------------------------------
 * Replaces set_cover.approx_multiuniverse(sets, costs=1, universe_p, ranks,
 * use_intervalsets=True) (utils/set_cover.py:147-615) as called from
 * filter/set_cover_filter.py:136-142.
 * ranks: one int32 per probe or NULL (all equal).  universe_p: one double per genome or NULL
 * (1.0).  sel_ids: caller-allocated, capacity n_probes; filled in PICK ORDER. */
int cb_setcover(cb_ctx *ctx, const cb_cover *cover, const int32_t *ranks, const double *universe_p,
                int64_t *sel_ids, int64_t *n_sel, cb_stats *stats);

/* The same with per-set costs (approx_multiuniverse's `costs`, utils/set_cover.py:147,426: each pick
 * minimises float(cost) / gain, smallest id on ties).  costs: one nonnegative finite double per probe, or
 * NULL (all 1).  SetCoverFilter itself only ever passes unit costs (filter/set_cover_filter.py:759). */
int cb_setcover_costs(cb_ctx *ctx, const cb_cover *cover, const double *costs, const int32_t *ranks,
                      const double *universe_p, int64_t *sel_ids, int64_t *n_sel, cb_stats *stats);

/* ---- multi-GPU: probe sharding inside one grouping ------------------------------------------
 * One process per GPU.  Rank 0 obtains an NCCL unique id, the host distributes the 128 bytes to all
 * ranks (any side channel), every rank joins with cb_comm_init.  Each rank then runs cb_coverage
 * on ALL targets of the grouping and its own contiguous block of probes
 * [probe_lo, probe_lo + local_n), and cb_cover_allgather assembles the full cover on every rank
 * (NCCL broadcasts over NVLink; the only exchange step of the path).  cb_setcover on that cover
 * gives the same picks on every rank. */
int cb_comm_unique_id(cb_ctx *ctx, uint8_t out[128]);
int cb_comm_init(cb_ctx *ctx, const uint8_t id[128], int32_t rank, int32_t n_ranks);
int cb_comm_destroy(cb_ctx *ctx);
int cb_cover_allgather(cb_ctx *ctx, const cb_cover *local, int64_t probe_lo, int64_t n_probes_total,
                       cb_cover **out);

/* ---- multi-GPU: sharded set cover (the reference loop is utils/set_cover.py:448-613) --
 * The candidate probes of ONE grouping are sharded over the GPUs of a box; every rank holds the cover
 * of its own probes only (cb_coverage_range), its gains and its interval index, plus a replica of
 * the universe bit set.  The greedy loop runs as one persistent kernel per GPU; once per round the
 * kernels exchange their active candidates through peer-mapped memory (stores and loads over
 * NVLink inside the kernel, no host involvement), see csrc/rounds.cu.
 *
 * Set-up, once per process group (or whenever a larger area is needed):
 *   cb_exchange_alloc    (re)allocates this rank's exchange area (cudaMalloc, zeroed);
 *   cb_exchange_handle   returns its CUDA IPC handle (64 bytes) and its device address;
 *   cb_exchange_attach   maps the areas of all ranks: `handles` = n_ranks * 64 bytes gathered from all
 *                        ranks (one process per GPU), or `addresses` = n_ranks device pointers when the
 *                        ranks are contexts of ONE process (tests: several ranks on one device; then
 *                        grid_limit > 0 caps each rank's persistent grid so that all fit together).
 *                        The caller must put a barrier between the attach calls and the first
 *                        cb_setcover_sharded.
 *   cb_exchange_required bytes a cover needs in the area (the maximum over ranks must fit everywhere).
 * cb_setcover_sharded is collective: every rank calls it with the cover of its own probes
 * [probe_lo, probe_hi) (global ids, universe_p == 1, unit costs) and receives the same picks in pick
 * order.  A rank that does not arrive within 30 s makes the others fail with CB_ERR_COMM. */
int cb_exchange_alloc(cb_ctx *ctx, int64_t bytes);
int64_t cb_exchange_bytes(cb_ctx *ctx);
int cb_exchange_handle(cb_ctx *ctx, uint8_t out[64], uint64_t *address);
int cb_exchange_attach(cb_ctx *ctx, int32_t rank, int32_t n_ranks, const uint8_t *handles,
                       const uint64_t *addresses, int32_t grid_limit);
int cb_exchange_required(cb_ctx *ctx, const cb_cover *cover, int64_t *bytes);
int cb_setcover_sharded(cb_ctx *ctx, const cb_cover *cover, int64_t probe_lo, int64_t probe_hi,
                        const int32_t *ranks, int64_t *sel_ids, int64_t *n_sel, cb_stats *stats);
/* The same in two steps: _begin does everything that is local to the rank (work buffers, universe bits,
 * gains, interval index; it never waits for another rank and returns with the stream idle), _end is the
 * collective part (the persistent kernel, the picks) and releases the job.  Lets a caller overlap the
 * set-up with other work, and lets ranks that share one device (tests) finish all host-side set-up
 * before the first persistent kernel starts. */
typedef struct cb_job cb_job;
int cb_setcover_sharded_begin(cb_ctx *ctx, const cb_cover *cover, int64_t probe_lo, int64_t probe_hi,
                              const int32_t *ranks, cb_job **job);
int cb_setcover_sharded_end(cb_ctx *ctx, cb_job *job, int64_t *sel_ids, int64_t *n_sel, cb_stats *stats);

/* ---- near-duplicate filter (K9-K12) --------------------------------------------------
 * Replaces NearDuplicateFilter._filter (filter/near_duplicate_filter.py:47-103) with
 * lsh.NearNeighborLookup (utils/lsh.py:239-320).  `probes` are the DISTINCT probes in
 * priority order (multiplicity descending, stable; :61-66).  keep[i] = 1 iff probe i is kept.
 *
 * MinHash family (utils/lsh.py:74-148, N=1, use_fast_str_hash=True under PYTHONHASHSEED=0):
 * a/b hold n_tables*k_concat drawn parameters, function f of table t at index t*k_concat+f.
 * The inner string hash is CPython's SipHash-1-3 with a zero key over the k-mer's bytes, so
 * this entry point needs the probes' ASCII (passed again here). */
int cb_minhash_neardup(cb_ctx *ctx, const uint8_t *ascii, const int64_t *probe_off, int64_t n_probes,
                       const uint32_t *a, const uint32_t *b, int32_t n_tables, int32_t k_concat,
                       int32_t kmer_size, double dist_thres, uint8_t *keep, cb_stats *stats);

/* Hamming family (utils/lsh.py:16-45): positions holds n_tables*k_concat sampled indices;
 * all probes must have the same length. */
int cb_hamming_neardup(cb_ctx *ctx, const uint8_t *ascii, const int64_t *probe_off, int64_t n_probes,
                       const int32_t *positions, int32_t n_tables, int32_t k_concat,
                       int32_t dist_thres, uint8_t *keep, cb_stats *stats);

/* DuplicateFilter._filter (filter/duplicate_filter.py:20-26, list(OrderedDict.fromkeys(input))) and
 * the multiplicity count of NearDuplicateFilter (:61-63) on the device: identical sequences of the
 * probe list are grouped (exact comparison of the packed bit planes).  first_idx / count
 * (caller-allocated, capacity n_probes; count may be NULL) receive, for every distinct sequence in
 * order of first occurrence, the list index of that occurrence and the number of occurrences. */
int cb_group_duplicates(cb_ctx *ctx, const uint8_t *ascii, const int64_t *probe_off, int64_t n_probes,
                        int64_t *first_idx, int32_t *count, int64_t *n_distinct, cb_stats *stats);

/* The whole of NearDuplicateFilter._filter (filter/near_duplicate_filter.py:47-103) for one probe
 * list WITH its duplicates, in list order: identical sequences are grouped on the device
 * (occurrences[p] += 1, :61-63), the distinct ones are ranked by multiplicity descending and first
 * occurrence ascending (the stable sorted(..., reverse=True) of :64-66), and the LSH filter above
 * runs on them.  family: 0 = MinHash (a, b, kmer_size used), 1 = Hamming (positions used).
 * kept_first_idx (caller-allocated, capacity n_probes) receives, in priority order, the list index
 * of the FIRST occurrence of every kept sequence -- the Probe object the reference's dict keeps. */
int cb_neardup_filter(cb_ctx *ctx, const uint8_t *ascii, const int64_t *probe_off, int64_t n_probes,
                      int32_t family, const uint32_t *a, const uint32_t *b, const int32_t *positions,
                      int32_t n_tables, int32_t k_concat, int32_t kmer_size, double dist_thres,
                      int64_t *kept_first_idx, int64_t *n_kept, int64_t *n_distinct, cb_stats *stats);

/* ---- genome clustering: MinHash sketches of whole sequences ---------------
 * Replaces cluster.make_signatures_with_minhash (utils/cluster.py:29-46) with the hash function of
 * lsh.MinHashFamily(kmer_size, N).make_h() (utils/lsh.py:74-148, md5 inner hash :106-111): for every k-mer x
 * of a sequence, v = (a * int(md5(x).hexdigest(), 16) + b) mod (2^31 - 1); the sketch is the N smallest v in
 * sorted order, as a multiset, and a sequence with fewer than N k-mers counts each of its k-mers
 * ceil(N / num_kmers) times (:131-139).  a, b are the two draws of make_h (random.randint(1, p),
 * random.randint(0, p)); the host makes them because they come from Python's `random`.
 * Sequence i is ascii[seq_off[i] .. seq_off[i+1]); CB_ERR_ARG if one is shorter than kmer_size (the
 * reference asserts, :117).  kmer_size <= 55, N <= 1024.
 * stats: ms_scan_emit = hashing kernel, ms_merge = selection kernel, n_seed_lookups = bases hashed. */
typedef struct cb_sketches cb_sketches;
int cb_sketch_sequences(cb_ctx *ctx, const uint8_t *ascii, const int64_t *seq_off, int64_t n_seqs, int32_t kmer_size,
                        int32_t N, uint64_t a, uint64_t b, cb_sketches **out, cb_stats *stats);
/* Sketches computed elsewhere (n_seqs rows of N sorted values) as a device object. */
int cb_sketches_import(cb_ctx *ctx, const uint32_t *sig, int64_t n_seqs, int32_t N, cb_sketches **out);
/* Copy the sketches out: sig holds n_seqs * N values, row i = the signature tuple h(seq_i). */
int cb_sketches_export(cb_ctx *ctx, const cb_sketches *sk, uint32_t *sig);
void cb_sketches_free(cb_sketches *sk);
/* MinHashFamily.estimate_jaccard_dist (utils/lsh.py:166-214) of sketch rows[r] against EVERY sketch:
 * out[r * n_seqs + c], the double the reference's Python arithmetic gives (1.0 - intersect / union).  One call
 * serves one step of find_connected_components' search (utils/cluster.py:270-290). */
int cb_sketch_dist_rows(cb_ctx *ctx, const cb_sketches *sk, const int64_t *rows, int64_t n_rows, double *out);
/* The same rows, reduced to what a search step uses: for each of rows[r] the sketches c with
 * distance <= threshold, as (column, distance) pairs in ascending column order.  Row r's pairs are
 * idx[row_off[r] .. row_off[r+1]) / dist[...]; row_off (n_rows + 1 entries) is caller-allocated, *idx and *dist are
 * allocated by the library (cb_free_host; NULL when nothing is within the threshold).  The distances are the same
 * doubles cb_sketch_dist_rows returns; only the copy back shrinks (a row of 40 000 doubles to a few thousand pairs). */
int cb_sketch_near_rows(cb_ctx *ctx, const cb_sketches *sk, const int64_t *rows, int64_t n_rows, double threshold,
                        int64_t *row_off, uint32_t **idx, double **dist);
/* cluster.create_condensed_dist_matrix (utils/cluster.py:103-195): all n(n-1)/2 distances in scipy's condensed
 * order, rounded to float32 as the reference's shared c_float array does (:141-142). */
int cb_sketch_dist_condensed(cb_ctx *ctx, const cb_sketches *sk, float *out);

#ifdef __cplusplus
}
#endif
#endif /* CATCH_B200_H */
