// Internal declarations shared by the translation units of libcatchb200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/catch_b200.h"

// ---------------------------------------------------------------------------------------
// Device data layout
//
// Sequences are held as bit planes ("bit-sliced"): plane w, bit i = bit w of the symbol code
// of base i.  Two bases differ iff any plane differs, so the mismatch mask of a probe against
// a target window is OR_w (probe_plane_w XOR window_plane_w): one XOR+OR per plane per 64
// bases, for any alphabet size (2 planes for ACGT, 3 with N, up to 8 for arbitrary bytes).
//
// Targets: all sequences of a group back to back; target coordinate g is stored at bit
// (g + CB_FRONT_PAD) of every plane so windows that hang off the left end read zeros.
// Probes: probe p owns planes*NW consecutive words, [plane][word], NW = ceil(maxlen/64).
// Universe coordinates: genome u owns bits [ubase[u], ubase[u]+genome_len[u]) of the universe
// bit set; ubase is 64-aligned and leaves >= 1 unused bit between genomes so intervals of
// different genomes never touch.
// ---------------------------------------------------------------------------------------
#define CB_FRONT_PAD 256          // bits of zeros in front of target position 0
#define CB_BACK_PAD_WORDS 8       // zero words after the last base
#define CB_TILE 1024              // target positions per scan tile
#define CB_TILE_WORDS ((CB_TILE + 2 * CB_FRONT_PAD) / 64 + 2)   // staged words per plane (even)

struct cb_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;         // kernels launched since the counter was last reset
    void *flush_buf = nullptr;    // cb_flush_l2 scratch
    unsigned flush_val = 0;
    void *pinned[6] = {};         // cb_host_buffer staging (page-locked), grown on demand
    size_t pinned_bytes[6] = {};
    void *comm = nullptr;         // ncclComm_t once cb_comm_init has run
    int rank = 0, n_ranks = 1;
    // exchange area of the sharded set cover (cb_exchange_*): this context's own area and the areas of
    // the other ranks as they are mapped in this process
    unsigned char *xarea = nullptr;
    size_t xarea_bytes = 0;
    unsigned char *xpeer[CB_MAX_RANKS] = {};
    bool xpeer_ipc[CB_MAX_RANKS] = {};
    int xrank = 0, xn_ranks = 1;
    int xgrid_limit = 0;          // > 0: cap on the persistent kernel's grid (several ranks sharing one device)
    bool xarea_poisoned = false;
    // stage A: ranges emitted per (probe x target base) by the last scan and the parameters it ran with; sizes the
    // range list of the next scan with the same parameters without the counting pre-pass (see cb_coverage_impl)
    double range_density = 0.0;
    int32_t range_density_key[5] = {-1, -1, -1, -1, -1};
};

struct cb_targets {
    cb_ctx *ctx = nullptr;
    int bits = 0;
    int64_t n_seqs = 0;
    int32_t n_genomes = 0;
    int64_t total_bases = 0;       // T
    int64_t plane_words = 0;       // words per plane (even)
    uint64_t *d_planes = nullptr;  // [bits][plane_words]
    int64_t *d_seq_start = nullptr;    // [n_seqs+1] target coordinate of each sequence
    int32_t *d_seq_genome = nullptr;   // [n_seqs]
    uint32_t *d_seq_ubase = nullptr;   // [n_seqs] universe bit of the sequence's first base
    uint32_t *d_ubase = nullptr;       // [n_genomes+1] universe bit of each genome's first base
    std::vector<int64_t> h_seq_start;
    std::vector<uint32_t> h_ubase;     // [n_genomes+1]
    std::vector<int64_t> h_genome_len; // [n_genomes]
    int64_t universe_bits = 0;         // total bits (multiple of 64)
    uint8_t lut[256];
};

struct cb_probes {
    cb_ctx *ctx = nullptr;
    int bits = 0;
    int64_t n_probes = 0;
    int nw = 0;                        // words per plane per probe
    int max_len = 0;
    uint64_t *d_words = nullptr;       // [n_probes][bits][nw]
    int32_t *d_len = nullptr;          // [n_probes]
    uint8_t lut[256];
};

struct cb_cover {
    cb_ctx *ctx = nullptr;
    int64_t n_probes = 0;
    int32_t n_genomes = 0;
    int64_t n_intervals = 0;
    int64_t universe_bits = 0;
    uint32_t max_interval_len = 0;
    int64_t *d_iv_off = nullptr;       // [n_probes+1]
    uint2 *d_iv = nullptr;             // [n_intervals] (start, end) universe coordinates, sorted per probe
    uint32_t *d_ubase = nullptr;       // [n_genomes+1] (own copy)
    std::vector<uint32_t> h_ubase;
    std::vector<int64_t> h_genome_len;
};

// ---------------------------------------------------------------------------------------
// Error handling
// ---------------------------------------------------------------------------------------
#define CB_CUDA(ctx, call)                                                                   \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            char buf__[512];                                                                 \
            snprintf(buf__, sizeof buf__, "%s:%d: %s failed: %s", __FILE__, __LINE__, #call, \
                     cudaGetErrorString(e__));                                               \
            (ctx)->err = buf__;                                                              \
            return CB_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)

#define CB_TRY(expr)                \
    do {                            \
        int r__ = (expr);           \
        if (r__ != CB_OK) return r__; \
    } while (0)

static inline int cb_fail(cb_ctx *ctx, int code, const char *msg)
{
    if (ctx) ctx->err = msg;
    return code;
}

// Device memory comes from CUDA's stream-ordered allocator (cudaMallocAsync on the context's
// stream; the pool keeps freed blocks, see cb_init), so the many short-lived work buffers of a
// call cost no cudaMalloc/cudaFree round trips and no implicit device synchronisation.
extern thread_local cudaStream_t cb_tls_stream;     // stream of the context serving this thread's call

static inline cudaError_t cb_dev_alloc(cudaStream_t st, void **p, size_t bytes)
{
    return cudaMallocAsync(p, bytes ? bytes : 1, st);
}
static inline void cb_dev_free(cudaStream_t st, void *p)
{
    if (p) cudaFreeAsync(p, st);
}

// RAII work buffer, freed in stream order when it goes out of scope.
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { cb_dev_free(cb_tls_stream, p); }
    cudaError_t alloc(size_t count)
    {
        cb_dev_free(cb_tls_stream, p);
        p = nullptr;
        n = count;
        return cb_dev_alloc(cb_tls_stream, (void **)&p, count * sizeof(T));
    }
    T *release() { T *q = p; p = nullptr; return q; }
};

struct EventTimer {
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t s;
    explicit EventTimer(cudaStream_t st) : s(st) { cudaEventCreate(&a); cudaEventCreate(&b); }
    ~EventTimer() { cudaEventDestroy(a); cudaEventDestroy(b); }
    void start() { cudaEventRecord(a, s); }
    // records the end event; the value is read later with ms() after a sync
    void stop() { cudaEventRecord(b, s); }
    double ms() { float f = 0; cudaEventSynchronize(b); cudaEventElapsedTime(&f, a, b); return f; }
};

// ---------------------------------------------------------------------------------------
// Launchers implemented in the .cu files
// ---------------------------------------------------------------------------------------
// scan.cu
int cb_exclusive_scan_u32_to_i64(cb_ctx *ctx, const uint32_t *d_in, int64_t *d_out, int64_t n,
                                 int64_t *h_total);
// d_out has n+1 entries: d_out[i] = sum of d_in[0..i), d_out[n] = total.

// pack.cu
int cb_launch_pack_targets(cb_ctx *ctx, const uint8_t *d_ascii, int64_t total, const uint8_t *d_lut,
                           int bits, uint64_t *d_planes, int64_t plane_words);
int cb_launch_pack_probes(cb_ctx *ctx, const uint8_t *d_ascii, const int64_t *d_off, int gap, int64_t n_probes,
                          const uint8_t *d_lut, int bits, int nw, uint64_t *d_words, int32_t *d_len);
int cb_launch_byte_presence(cb_ctx *ctx, const uint8_t *d_buf, int64_t n, uint32_t *d_present);

int cb_probes_have_duplicates_impl(cb_ctx *ctx, const cb_probes *probes, int32_t *has_dup);
int cb_intop_rate_impl(cb_ctx *ctx, double *ops_per_s);

// coverage.cu
int cb_coverage_impl(cb_ctx *ctx, const cb_probes *probes, const cb_targets *targets,
                     const cb_hyb_params *params, const int64_t *seed_off, const int32_t *seed_pos,
                     const uint8_t *seed_pos_u8, int32_t seeds_per_probe, int64_t probe_lo, int64_t probe_hi,
                     cb_cover **out, cb_stats *stats, std::vector<uint32_t> *raw_records = nullptr);

int cb_cover_import_impl(cb_ctx *ctx, int64_t n_probes, int32_t n_genomes, const int64_t *genome_len,
                         int64_t n_intervals, const int64_t *probe_id, const int32_t *genome,
                         const int64_t *start, const int64_t *end, cb_cover **out);

// setcover.cu
int cb_setcover_impl(cb_ctx *ctx, const cb_cover *cover, const double *costs, const int32_t *ranks,
                     const double *universe_p, int64_t *sel_ids, int64_t *n_sel, cb_stats *stats);

// rounds.cu
int64_t cb_rounds_exchange_bytes(const cb_cover *cover);
int cb_setcover_rounds_impl(cb_ctx *ctx, const cb_cover *cover, int64_t lo, int64_t hi, const int32_t *ranks,
                            bool sharded, int64_t *sel_ids, int64_t *n_sel, cb_stats *stats);
struct cb_rounds_job;
int cb_rounds_begin_impl(cb_ctx *ctx, const cb_cover *cover, int64_t lo, int64_t hi, const int32_t *ranks,
                         cb_rounds_job **out);
int cb_rounds_end_impl(cb_rounds_job *job, int64_t *sel_ids, int64_t *n_sel, cb_stats *stats);

// neardup.cu
int cb_minhash_neardup_impl(cb_ctx *ctx, const uint8_t *ascii, const int64_t *probe_off, int64_t n_probes,
                            const uint32_t *a, const uint32_t *b, int32_t n_tables, int32_t k_concat,
                            int32_t kmer_size, double dist_thres, uint8_t *keep, cb_stats *stats);
int cb_neardup_filter_impl(cb_ctx *ctx, const uint8_t *ascii, const int64_t *probe_off, int64_t n, int32_t family,
                           const uint32_t *pa, const uint32_t *pb, const int32_t *positions, int32_t n_tables,
                           int32_t k_concat, int32_t kmer, double dist_thres, int64_t *kept_first_idx,
                           int64_t *n_kept, int64_t *n_distinct_out, cb_stats *stats);
int cb_group_duplicates_impl(cb_ctx *ctx, const uint8_t *ascii, const int64_t *probe_off, int64_t n,
                             int64_t *first_idx, int32_t *count, int64_t *n_distinct, cb_stats *stats);
int cb_hamming_neardup_impl(cb_ctx *ctx, const uint8_t *ascii, const int64_t *probe_off, int64_t n_probes,
                            const int32_t *positions, int32_t n_tables, int32_t k_concat,
                            int32_t dist_thres, uint8_t *keep, cb_stats *stats);

// cluster.cu
int cb_sketch_sequences_impl(cb_ctx *ctx, const uint8_t *ascii, const int64_t *seq_off, int64_t n, int32_t k,
                             int32_t N, uint64_t a, uint64_t b, cb_sketches **out, cb_stats *stats);
int cb_sketch_import_impl(cb_ctx *ctx, const uint32_t *sig, int64_t n, int32_t N, cb_sketches **out);
int cb_sketches_export_impl(cb_ctx *ctx, const cb_sketches *sk, uint32_t *sig);
int cb_sketch_dist_rows_impl(cb_ctx *ctx, const cb_sketches *sk, const int64_t *rows, int64_t n_rows, double *out);
int cb_sketch_dist_condensed_impl(cb_ctx *ctx, const cb_sketches *sk, float *out);
int cb_sketch_near_rows_impl(cb_ctx *ctx, const cb_sketches *sk, const int64_t *rows, int64_t n_rows, double threshold,
                             int64_t *row_off, uint32_t **idx, double **dist);
