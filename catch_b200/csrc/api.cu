// C ABI of libcatchb200.so (see include/catch_b200.h): context, packing (uploads), exports and
// the thin wrappers around the stage implementations.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "internal.cuh"

static int upload_common_lut(cb_ctx *ctx, const uint8_t lut[256], int bits, DevBuf<uint8_t> &d_lut)
{
    if (bits < 1 || bits > CB_MAX_SYMBOL_BITS) return cb_fail(ctx, CB_ERR_ARG, "bits must be in [1, 8]");
    for (int i = 0; i < 256; i++)
        if (bits < 8 && (lut[i] >> bits)) return cb_fail(ctx, CB_ERR_ARG, "lut code does not fit in `bits`");
    CB_CUDA(ctx, d_lut.alloc(256));
    CB_CUDA(ctx, cudaMemcpyAsync(d_lut.p, lut, 256, cudaMemcpyHostToDevice, ctx->stream));
    return CB_OK;
}

// ---- staging and packing shared by cb_upload_targets / cb_upload_probes / cb_upload_group
extern "C" void cb_targets_free(cb_targets *t);
extern "C" void cb_probes_free(cb_probes *p);

// host tables (sequence starts, universe layout), allocations and the host->device copy of the
// bytes; *out is allocated here and owned by the caller from then on
static int targets_stage(cb_ctx *ctx, const uint8_t *ascii, const int64_t *seq_off, int64_t n_seqs,
                         const int32_t *seq_genome, int32_t n_genomes, cb_targets **out, DevBuf<uint8_t> &d_ascii)
{
    if (n_seqs < 0 || n_genomes < 0) return cb_fail(ctx, CB_ERR_ARG, "bad argument");
    if (n_seqs > 0 && (!seq_off || !seq_genome)) return cb_fail(ctx, CB_ERR_ARG, "null sequence table");
    cudaStream_t st = ctx->stream;
    const int64_t T = n_seqs ? seq_off[n_seqs] - seq_off[0] : 0;
    if (T > 0 && !ascii) return cb_fail(ctx, CB_ERR_ARG, "null ascii");
    cb_targets *t = new cb_targets();
    *out = t;
    t->ctx = ctx;
    t->n_seqs = n_seqs;
    t->n_genomes = n_genomes;
    t->total_bases = T;
    t->h_seq_start.resize((size_t)n_seqs + 1);
    t->h_genome_len.assign((size_t)n_genomes, 0);
    std::vector<int64_t> seq_in_genome((size_t)n_seqs, 0);
    for (int64_t i = 0; i < n_seqs; i++) {
        const int64_t len = seq_off[i + 1] - seq_off[i];
        if (len < 0) return cb_fail(ctx, CB_ERR_ARG, "seq_off not monotone");
        const int32_t g = seq_genome[i];
        if (g < 0 || g >= n_genomes || (i > 0 && g < seq_genome[i - 1]))
            return cb_fail(ctx, CB_ERR_ARG, "seq_genome must be non-decreasing and < n_genomes");
        t->h_seq_start[(size_t)i] = seq_off[i] - seq_off[0];
        seq_in_genome[(size_t)i] = t->h_genome_len[(size_t)g];      // length_so_far, set_cover_filter.py:418,453
        t->h_genome_len[(size_t)g] += len;
    }
    t->h_seq_start[(size_t)n_seqs] = T;
    t->h_ubase.resize((size_t)n_genomes + 1);
    uint64_t ub = 0;
    for (int32_t g = 0; g < n_genomes; g++) {
        t->h_ubase[(size_t)g] = (uint32_t)ub;
        ub = (ub + (uint64_t)t->h_genome_len[(size_t)g] + 1 + 63) & ~63ull;    // >= 1 spare bit, 64-aligned
        if (ub >= 0xffffff00ull) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "target group exceeds 2^32 universe bits");
    }
    t->h_ubase[(size_t)n_genomes] = (uint32_t)ub;
    t->universe_bits = (int64_t)ub;
    std::vector<uint32_t> h_seq_ubase((size_t)n_seqs);
    for (int64_t i = 0; i < n_seqs; i++)
        h_seq_ubase[(size_t)i] = t->h_ubase[(size_t)seq_genome[i]] + (uint32_t)seq_in_genome[(size_t)i];

    int64_t pw = (T + CB_FRONT_PAD + 63) / 64 + CB_TILE_WORDS + CB_BACK_PAD_WORDS;
    pw = (pw + 1) & ~1ll;
    t->plane_words = pw;
    CB_CUDA(ctx, cb_dev_alloc(st, (void **)&t->d_seq_start, sizeof(int64_t) * (size_t)(n_seqs + 1)));
    CB_CUDA(ctx, cb_dev_alloc(st, (void **)&t->d_seq_genome, sizeof(int32_t) * (size_t)(n_seqs + 1)));
    CB_CUDA(ctx, cb_dev_alloc(st, (void **)&t->d_seq_ubase, sizeof(uint32_t) * (size_t)(n_seqs + 1)));
    CB_CUDA(ctx, cb_dev_alloc(st, (void **)&t->d_ubase, sizeof(uint32_t) * (size_t)(n_genomes + 1)));
    CB_CUDA(ctx, d_ascii.alloc((size_t)T));
    if (T) CB_CUDA(ctx, cudaMemcpyAsync(d_ascii.p, ascii + seq_off[0], (size_t)T, cudaMemcpyHostToDevice, st));
    CB_CUDA(ctx, cudaMemcpyAsync(t->d_seq_start, t->h_seq_start.data(), sizeof(int64_t) * (size_t)(n_seqs + 1), cudaMemcpyHostToDevice, st));
    if (n_seqs) {
        CB_CUDA(ctx, cudaMemcpyAsync(t->d_seq_genome, seq_genome, sizeof(int32_t) * (size_t)n_seqs, cudaMemcpyHostToDevice, st));
        CB_CUDA(ctx, cudaMemcpyAsync(t->d_seq_ubase, h_seq_ubase.data(), sizeof(uint32_t) * (size_t)n_seqs, cudaMemcpyHostToDevice, st));
    }
    CB_CUDA(ctx, cudaMemcpyAsync(t->d_ubase, t->h_ubase.data(), sizeof(uint32_t) * (size_t)(n_genomes + 1), cudaMemcpyHostToDevice, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));      // h_seq_ubase goes out of scope
    return CB_OK;
}

static int targets_pack(cb_ctx *ctx, cb_targets *t, const uint8_t *d_ascii, const uint8_t lut[256], int bits)
{
    DevBuf<uint8_t> d_lut;
    CB_TRY(upload_common_lut(ctx, lut, bits, d_lut));
    t->bits = bits;
    memcpy(t->lut, lut, 256);
    CB_CUDA(ctx, cb_dev_alloc(ctx->stream, (void **)&t->d_planes, sizeof(uint64_t) * (size_t)t->plane_words * (size_t)bits));
    return cb_launch_pack_targets(ctx, d_ascii, t->total_bases, d_lut.p, bits, t->d_planes, t->plane_words);
}

// Probes either with explicit offsets (probe_off != NULL) or as one buffer of `bytes` bytes in
// which consecutive probes are separated by one `sep` byte (offsets found here with memchr).
static int probes_stage(cb_ctx *ctx, const uint8_t *ascii, int64_t bytes, const int64_t *probe_off, int64_t n_probes,
                        int sep, int32_t *len_out, cb_probes **out, DevBuf<uint8_t> &d_ascii, DevBuf<int64_t> &d_off)
{
    if (n_probes < 0) return cb_fail(ctx, CB_ERR_ARG, "bad argument");
    if (n_probes > 0 && !ascii && (probe_off ? probe_off[n_probes] > probe_off[0] : bytes > 0))
        return cb_fail(ctx, CB_ERR_ARG, "null probe bytes");
    if (n_probes >= 0xfffffff0ll) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "too many probes");
    cudaStream_t st = ctx->stream;
    std::vector<int64_t> rel((size_t)n_probes + 1, 0);
    int64_t total = 0;
    const int gap = probe_off ? 0 : 1;
    if (probe_off) {
        for (int64_t i = 0; i <= n_probes; i++) rel[(size_t)i] = probe_off[i] - probe_off[0];
        total = rel[(size_t)n_probes];
        ascii = ascii ? ascii + probe_off[0] : ascii;
    } else if (n_probes > 0) {
        if (sep < 0 || sep > 255 || bytes < n_probes - 1) return cb_fail(ctx, CB_ERR_ARG, "bad separator / size");
        const uint8_t *q = ascii, *end = ascii + bytes;
        for (int64_t i = 0; i < n_probes; i++) {
            rel[(size_t)i] = q - ascii;
            const uint8_t *hit = (q < end) ? (const uint8_t *)memchr(q, sep, (size_t)(end - q)) : nullptr;
            if (i + 1 < n_probes) {
                if (!hit) return cb_fail(ctx, CB_ERR_ARG, "fewer separators than probes");
                q = hit + 1;
            } else if (hit) {
                return cb_fail(ctx, CB_ERR_ARG, "separator byte occurs inside a probe");
            }
        }
        rel[(size_t)n_probes] = bytes + 1;          // as if a separator followed the last probe
        total = bytes;
    }
    int max_len = 0;
    for (int64_t i = 0; i < n_probes; i++) {
        const int64_t len = rel[(size_t)i + 1] - rel[(size_t)i] - gap;
        if (len < 0) return cb_fail(ctx, CB_ERR_ARG, "probe_off not monotone");
        if (len > CB_MAX_PROBE_LEN) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "probe longer than CB_MAX_PROBE_LEN (256)");
        if (len > max_len) max_len = (int)len;
        if (len_out) len_out[i] = (int32_t)len;
    }
    cb_probes *p = new cb_probes();
    *out = p;
    p->ctx = ctx;
    p->n_probes = n_probes;
    p->max_len = max_len;
    p->nw = max_len ? (max_len + 63) / 64 : 1;
    CB_CUDA(ctx, cb_dev_alloc(st, (void **)&p->d_len, sizeof(int32_t) * (size_t)(n_probes ? n_probes : 1)));
    CB_CUDA(ctx, d_ascii.alloc((size_t)total));
    CB_CUDA(ctx, d_off.alloc((size_t)n_probes + 1));
    if (n_probes) {
        if (total) CB_CUDA(ctx, cudaMemcpyAsync(d_ascii.p, ascii, (size_t)total, cudaMemcpyHostToDevice, st));
        CB_CUDA(ctx, cudaMemcpyAsync(d_off.p, rel.data(), sizeof(int64_t) * (size_t)(n_probes + 1), cudaMemcpyHostToDevice, st));
        CB_CUDA(ctx, cudaStreamSynchronize(st));    // `rel` goes out of scope
    }
    return CB_OK;
}

static int probes_pack(cb_ctx *ctx, cb_probes *p, const uint8_t *d_ascii, const int64_t *d_off, int gap,
                       const uint8_t lut[256], int bits)
{
    DevBuf<uint8_t> d_lut;
    CB_TRY(upload_common_lut(ctx, lut, bits, d_lut));
    p->bits = bits;
    memcpy(p->lut, lut, 256);
    CB_CUDA(ctx, cb_dev_alloc(ctx->stream, (void **)&p->d_words,
                              sizeof(uint64_t) * (size_t)(p->n_probes ? p->n_probes : 1) * (size_t)bits * (size_t)p->nw));
    return cb_launch_pack_probes(ctx, d_ascii, d_off, gap, p->n_probes, d_lut.p, bits, p->nw, p->d_words, p->d_len);
}

thread_local cudaStream_t cb_tls_stream = nullptr;

extern "C" {

const char *cb_version(void) { return "catch_b200 0.1 sm_100a"; }

int cb_init(int device_id, cb_ctx **out)
{
    if (!out) return CB_ERR_ARG;
    *out = nullptr;
    cb_ctx *ctx = new cb_ctx();
    ctx->device = device_id;
    cudaError_t e = cudaSetDevice(device_id);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device_id);
    if (e != cudaSuccess) {
        // keep the context alive so the caller can read the message, but mark it unusable
        ctx->err = std::string("cb_init: ") + cudaGetErrorString(e);
        *out = ctx;
        return CB_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    {   // keep freed blocks in the stream-ordered pool instead of returning them to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device_id) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    if (prop.major < 10) {
        ctx->err = "cb_init: this library is built for sm_100a (Blackwell) only";
        *out = ctx;
        return CB_ERR_UNSUPPORTED;
    }
    {   // Work buffers of a call add up to a few hundred MB .. a few GB; growing the pool piecemeal in the middle of
        // a call was measured to cost up to 400 ms (the driver maps new memory with the device idle), so the pool is
        // grown once here: CB_POOL_RESERVE_MB (default 8192, 0 = leave it to the first calls), never more than a
        // quarter of the device's free memory.
        long long mb = 8192;
        if (const char *e = getenv("CB_POOL_RESERVE_MB")) mb = atoll(e);
        size_t free_b = 0, total_b = 0;
        cudaMemPool_t pool;
        unsigned long long have = 0;
        if (mb > 0 && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess &&
            cudaDeviceGetDefaultMemPool(&pool, device_id) == cudaSuccess &&
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &have) == cudaSuccess) {
            unsigned long long want = (unsigned long long)mb << 20;
            if (want > free_b / 4) want = free_b / 4;
            if (have < want) {
                void *p = nullptr;
                if (cudaMallocAsync(&p, (size_t)(want - have), ctx->stream) == cudaSuccess) {
                    cudaFreeAsync(p, ctx->stream);
                    cudaStreamSynchronize(ctx->stream);
                } else {
                    cudaGetLastError();          // not fatal: the calls grow the pool as they go
                }
            }
        }
    }
    *out = ctx;
    return CB_OK;
}

void cb_destroy(cb_ctx *ctx)
{
    if (!ctx) return;
    cb_comm_destroy(ctx);
    cb_exchange_alloc(ctx, 0);
    if (ctx->flush_buf) cudaFree(ctx->flush_buf);
    for (int i = 0; i < 6; i++)
        if (ctx->pinned[i]) cudaFreeHost(ctx->pinned[i]);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *cb_last_error(cb_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int cb_host_buffer(cb_ctx *ctx, int32_t slot, int64_t bytes, void **out)
{
    if (!ctx || !out || slot < 0 || slot > 5 || bytes < 0) return cb_fail(ctx, CB_ERR_ARG, "bad argument");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    if ((size_t)bytes > ctx->pinned_bytes[slot]) {
        // the previous buffer may still feed an asynchronous copy
        CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->pinned[slot]) CB_CUDA(ctx, cudaFreeHost(ctx->pinned[slot]));
        ctx->pinned[slot] = nullptr;
        ctx->pinned_bytes[slot] = 0;
        size_t want = (size_t)bytes + ((size_t)bytes >> 2) + 4096;      // headroom: sizes drift between groupings
        CB_CUDA(ctx, cudaHostAlloc(&ctx->pinned[slot], want, cudaHostAllocDefault));
        ctx->pinned_bytes[slot] = want;
    }
    *out = ctx->pinned[slot];
    return CB_OK;
}

int cb_flush_l2(cb_ctx *ctx)
{
    if (!ctx) return CB_ERR_ARG;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    const size_t bytes = 256ull << 20;
    if (!ctx->flush_buf) CB_CUDA(ctx, cudaMalloc(&ctx->flush_buf, bytes));
    CB_CUDA(ctx, cudaMemsetAsync(ctx->flush_buf, (int)(++ctx->flush_val & 0xff), bytes, ctx->stream));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CB_OK;
}

int cb_pool_reserve(cb_ctx *ctx, int64_t bytes)
{
    if (!ctx || bytes < 0) return cb_fail(ctx, CB_ERR_ARG, "bad argument");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    void *p = nullptr;
    CB_CUDA(ctx, cudaMallocAsync(&p, (size_t)(bytes ? bytes : 1), ctx->stream));
    CB_CUDA(ctx, cudaFreeAsync(p, ctx->stream));        // stays in the pool (release threshold: never)
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CB_OK;
}

int cb_intop_rate(cb_ctx *ctx, double *ops_per_s)
{
    if (!ctx || !ops_per_s) return cb_fail(ctx, CB_ERR_ARG, "null argument");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    return cb_intop_rate_impl(ctx, ops_per_s);
}

int cb_upload_targets(cb_ctx *ctx, const uint8_t *ascii, const int64_t *seq_off, int64_t n_seqs,
                      const int32_t *seq_genome, int32_t n_genomes, const uint8_t lut[256], int32_t bits,
                      cb_targets **out, cb_stats *stats)
{
    if (!ctx || !out || !lut) return cb_fail(ctx, CB_ERR_ARG, "bad argument");
    *out = nullptr;
    if (stats) memset(stats, 0, sizeof *stats);
    ctx->launches = 0;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    cudaStream_t st = ctx->stream;
    EventTimer t_all(st), t_h2d(st), t_pack(st);
    t_all.start();
    cb_targets *t = nullptr;
    struct Guard { cb_targets **t; ~Guard() { if (*t) cb_targets_free(*t); } } guard{&t};
    DevBuf<uint8_t> d_ascii;
    t_h2d.start();
    CB_TRY(targets_stage(ctx, ascii, seq_off, n_seqs, seq_genome, n_genomes, &t, d_ascii));
    t_h2d.stop();
    t_pack.start();
    CB_TRY(targets_pack(ctx, t, d_ascii.p, lut, bits));
    t_pack.stop();
    t_all.stop();
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    if (stats) {
        stats->ms_h2d = t_h2d.ms();
        stats->ms_pack = t_pack.ms();
        stats->ms_total = t_all.ms();
        stats->n_kernel_launches = ctx->launches;
    }
    *out = t;
    t = nullptr;
    return CB_OK;
}

void cb_targets_free(cb_targets *t)
{
    if (!t) return;
    cudaStream_t st = t->ctx->stream;
    cb_dev_free(st, t->d_planes);
    cb_dev_free(st, t->d_seq_start);
    cb_dev_free(st, t->d_seq_genome);
    cb_dev_free(st, t->d_seq_ubase);
    cb_dev_free(st, t->d_ubase);
    delete t;
}

int cb_upload_probes(cb_ctx *ctx, const uint8_t *ascii, const int64_t *probe_off, int64_t n_probes,
                     const uint8_t lut[256], int32_t bits, cb_probes **out, cb_stats *stats)
{
    if (!ctx || !out || !lut) return cb_fail(ctx, CB_ERR_ARG, "bad argument");
    *out = nullptr;
    if (stats) memset(stats, 0, sizeof *stats);
    ctx->launches = 0;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    cudaStream_t st = ctx->stream;
    EventTimer t_all(st), t_h2d(st), t_pack(st);
    t_all.start();
    cb_probes *p = nullptr;
    struct Guard { cb_probes **p; ~Guard() { if (*p) cb_probes_free(*p); } } guard{&p};
    DevBuf<uint8_t> d_ascii;
    DevBuf<int64_t> d_off;
    t_h2d.start();
    CB_TRY(probes_stage(ctx, ascii, 0, probe_off, n_probes, -1, nullptr, &p, d_ascii, d_off));
    t_h2d.stop();
    t_pack.start();
    CB_TRY(probes_pack(ctx, p, d_ascii.p, d_off.p, 0, lut, bits));
    t_pack.stop();
    t_all.stop();
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    if (stats) {
        stats->ms_h2d = t_h2d.ms();
        stats->ms_pack = t_pack.ms();
        stats->ms_total = t_all.ms();
        stats->n_kernel_launches = ctx->launches;
    }
    *out = p;
    p = nullptr;
    return CB_OK;
}

int cb_upload_group(cb_ctx *ctx, const uint8_t *probes_ascii, int64_t probes_bytes, const int64_t *probe_off,
                    int64_t n_probes, int32_t sep, const uint8_t *targets_ascii, const int64_t *seq_off,
                    int64_t n_seqs, const int32_t *seq_genome, int32_t n_genomes, int32_t *probe_len_out,
                    int32_t *bits_out, cb_probes **probes_out, cb_targets **targets_out, cb_stats *stats)
{
    if (!ctx || !probes_out || !targets_out) return cb_fail(ctx, CB_ERR_ARG, "bad argument");
    *probes_out = nullptr;
    *targets_out = nullptr;
    if (stats) memset(stats, 0, sizeof *stats);
    ctx->launches = 0;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    cudaStream_t st = ctx->stream;
    EventTimer t_all(st), t_h2d(st), t_pack(st);
    t_all.start();
    cb_probes *p = nullptr;
    cb_targets *t = nullptr;
    struct Guard {
        cb_probes **p; cb_targets **t;
        ~Guard() { if (*p) cb_probes_free(*p); if (*t) cb_targets_free(*t); }
    } guard{&p, &t};
    DevBuf<uint8_t> d_pascii, d_tascii;
    DevBuf<int64_t> d_off;
    DevBuf<uint32_t> d_present;
    t_h2d.start();
    CB_TRY(targets_stage(ctx, targets_ascii, seq_off, n_seqs, seq_genome, n_genomes, &t, d_tascii));
    CB_TRY(probes_stage(ctx, probes_ascii, probes_bytes, probe_off, n_probes, probe_off ? -1 : sep, probe_len_out,
                        &p, d_pascii, d_off));
    t_h2d.stop();
    t_pack.start();
    // code table: the distinct bytes of both buffers (ACGT always) get dense codes in byte order,
    // found on the device since the bytes are there anyway
    CB_CUDA(ctx, d_present.alloc(512));
    CB_CUDA(ctx, cudaMemsetAsync(d_present.p, 0, sizeof(uint32_t) * 512, st));
    const int64_t pbytes = probe_off ? (n_probes ? probe_off[n_probes] - probe_off[0] : 0) : probes_bytes;
    CB_TRY(cb_launch_byte_presence(ctx, d_pascii.p, pbytes, d_present.p));
    CB_TRY(cb_launch_byte_presence(ctx, d_tascii.p, t->total_bases, d_present.p + 256));
    uint32_t h_present[512];
    CB_CUDA(ctx, cudaMemcpyAsync(h_present, d_present.p, sizeof h_present, cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    if (!probe_off && sep >= 0 && sep < 256) h_present[sep] = 0;          // separators are not symbols
    uint8_t lut[256];
    memset(lut, 0, sizeof lut);
    int n_sym = 0;
    for (int c = 0; c < 256; c++) {
        const bool here = h_present[c] || h_present[256 + c] || c == 'A' || c == 'C' || c == 'G' || c == 'T';
        if (here) lut[c] = (uint8_t)n_sym++;
    }
    int bits = 1;
    while ((1 << bits) < n_sym) bits++;
    CB_TRY(targets_pack(ctx, t, d_tascii.p, lut, bits));
    CB_TRY(probes_pack(ctx, p, d_pascii.p, d_off.p, probe_off ? 0 : 1, lut, bits));
    t_pack.stop();
    t_all.stop();
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    if (bits_out) *bits_out = bits;
    if (stats) {
        stats->ms_h2d = t_h2d.ms();
        stats->ms_pack = t_pack.ms();
        stats->ms_total = t_all.ms();
        stats->n_kernel_launches = ctx->launches;
    }
    *probes_out = p;
    *targets_out = t;
    p = nullptr;
    t = nullptr;
    return CB_OK;
}

void cb_probes_free(cb_probes *p)
{
    if (!p) return;
    cb_dev_free(p->ctx->stream, p->d_words);
    cb_dev_free(p->ctx->stream, p->d_len);
    delete p;
}

// (the MT19937 replay, cb_mt19937_*, lives in rng.cpp: host-only code with an AVX-512 path)

int cb_split_lengths(const uint8_t *buf, int64_t bytes, int64_t n, int32_t sep, int32_t *len_out)
{
    if (n < 0 || bytes < 0 || sep < 0 || sep > 255 || (n > 0 && !len_out) || (bytes > 0 && !buf)) return CB_ERR_ARG;
    const uint8_t *q = buf, *end = buf + bytes;
    for (int64_t i = 0; i < n; i++) {
        const uint8_t *hit = (q < end) ? (const uint8_t *)memchr(q, sep, (size_t)(end - q)) : nullptr;
        if (i + 1 < n) {
            if (!hit) return CB_ERR_ARG;
            len_out[i] = (int32_t)(hit - q);
            q = hit + 1;
        } else {
            if (hit) return CB_ERR_ARG;
            len_out[i] = (int32_t)(end - q);
        }
    }
    return (n == 0 && bytes > 0) ? CB_ERR_ARG : CB_OK;
}

int cb_probes_have_duplicates(cb_ctx *ctx, const cb_probes *probes, int32_t *has_dup)
{
    if (!ctx || !probes || !has_dup) return cb_fail(ctx, CB_ERR_ARG, "null argument");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    return cb_probes_have_duplicates_impl(ctx, probes, has_dup);
}

int cb_coverage(cb_ctx *ctx, const cb_probes *probes, const cb_targets *targets, const cb_hyb_params *params,
                const int64_t *seed_off, const int32_t *seed_pos, cb_cover **out, cb_stats *stats)
{
    if (!ctx) return CB_ERR_ARG;
    if (stats) memset(stats, 0, sizeof *stats);
    ctx->launches = 0;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    return cb_coverage_impl(ctx, probes, targets, params, seed_off, seed_pos, nullptr, 0, 0, -1, out, stats);
}

int cb_coverage_uniform(cb_ctx *ctx, const cb_probes *probes, const cb_targets *targets, const cb_hyb_params *params,
                        const uint8_t *seed_pos, int32_t seeds_per_probe, cb_cover **out, cb_stats *stats)
{
    if (!ctx) return CB_ERR_ARG;
    if (!seed_pos) return cb_fail(ctx, CB_ERR_ARG, "null seed positions");
    if (stats) memset(stats, 0, sizeof *stats);
    ctx->launches = 0;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    return cb_coverage_impl(ctx, probes, targets, params, nullptr, nullptr, seed_pos, seeds_per_probe, 0, -1, out, stats);
}

int cb_coverage_range(cb_ctx *ctx, const cb_probes *probes, const cb_targets *targets, const cb_hyb_params *params,
                      const int64_t *seed_off, const int32_t *seed_pos, const uint8_t *seed_pos_u8,
                      int32_t seeds_per_probe, int64_t probe_lo, int64_t probe_hi, cb_cover **out, cb_stats *stats)
{
    if (!ctx) return CB_ERR_ARG;
    if (stats) memset(stats, 0, sizeof *stats);
    ctx->launches = 0;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    return cb_coverage_impl(ctx, probes, targets, params, seed_off, seed_pos, seed_pos_u8, seeds_per_probe, probe_lo,
                            probe_hi, out, stats);
}

int cb_coverage_records(cb_ctx *ctx, const cb_probes *probes, const cb_targets *targets, const cb_hyb_params *params,
                        const int64_t *seed_off, const int32_t *seed_pos, int64_t *n_records, uint32_t **records,
                        cb_stats *stats)
{
    if (!ctx) return CB_ERR_ARG;
    if (!n_records || !records || !targets) return cb_fail(ctx, CB_ERR_ARG, "null argument");
    *n_records = 0;
    *records = nullptr;
    if (stats) memset(stats, 0, sizeof *stats);
    ctx->launches = 0;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    if (targets->total_bases >= 0xffffffffll) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "target group exceeds 2^32 bases");
    std::vector<uint32_t> raw;
    cb_cover *cov = nullptr;
    const int rc = cb_coverage_impl(ctx, probes, targets, params, seed_off, seed_pos, nullptr, 0, 0, -1, &cov, stats, &raw);
    if (cov) cb_cover_free(cov);
    if (rc != CB_OK) return rc;
    const size_t n = raw.size() / 4;
    if (n) {
        uint32_t *buf = (uint32_t *)malloc(n * 5 * sizeof(uint32_t));
        if (!buf) return cb_fail(ctx, CB_ERR_NOMEM, "out of host memory");
        // records -> (probe, sequence, start, end, hit position): universe and target coordinates become
        // positions inside the sequence the range lies in
        const std::vector<int64_t> &ss = targets->h_seq_start;
        std::vector<uint32_t> seq_ubase((size_t)targets->n_seqs);
        CB_CUDA(ctx, cudaMemcpy(seq_ubase.data(), targets->d_seq_ubase, sizeof(uint32_t) * (size_t)targets->n_seqs, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; i++) {
            const uint32_t hit = raw[4 * i + 3];
            const size_t q = (size_t)(std::upper_bound(ss.begin(), ss.end(), (int64_t)hit) - ss.begin()) - 1;
            buf[5 * i] = raw[4 * i];
            buf[5 * i + 1] = (uint32_t)q;
            buf[5 * i + 2] = raw[4 * i + 1] - seq_ubase[q];
            buf[5 * i + 3] = raw[4 * i + 2] - seq_ubase[q];
            buf[5 * i + 4] = (uint32_t)((int64_t)hit - ss[q]);
        }
        *records = buf;
    }
    *n_records = (int64_t)n;
    return CB_OK;
}

void cb_free_host(void *p) { free(p); }

void cb_cover_free(cb_cover *c)
{
    if (!c) return;
    cb_dev_free(c->ctx->stream, c->d_iv_off);
    cb_dev_free(c->ctx->stream, c->d_iv);
    cb_dev_free(c->ctx->stream, c->d_ubase);
    delete c;
}

int64_t cb_cover_num_intervals(const cb_cover *c) { return c ? c->n_intervals : 0; }

int cb_cover_export(cb_ctx *ctx, const cb_cover *c, int64_t *probe_id, int32_t *genome, int64_t *start, int64_t *end)
{
    if (!ctx || !c) return cb_fail(ctx, CB_ERR_ARG, "null argument");
    const int64_t E = c->n_intervals, P = c->n_probes;
    if (E == 0) return CB_OK;
    if (!probe_id || !genome || !start || !end) return cb_fail(ctx, CB_ERR_ARG, "null output array");
    std::vector<int64_t> off((size_t)P + 1);
    std::vector<uint2> iv((size_t)E);
    CB_CUDA(ctx, cudaMemcpyAsync(off.data(), c->d_iv_off, sizeof(int64_t) * (size_t)(P + 1), cudaMemcpyDeviceToHost, ctx->stream));
    CB_CUDA(ctx, cudaMemcpyAsync(iv.data(), c->d_iv, sizeof(uint2) * (size_t)E, cudaMemcpyDeviceToHost, ctx->stream));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const std::vector<uint32_t> &ub = c->h_ubase;
    for (int64_t p = 0; p < P; p++)
        for (int64_t i = off[(size_t)p]; i < off[(size_t)p + 1]; i++) {
            const uint32_t s = iv[(size_t)i].x, e = iv[(size_t)i].y;
            const int32_t g = (int32_t)(std::upper_bound(ub.begin(), ub.end(), s) - ub.begin()) - 1;
            probe_id[i] = p;
            genome[i] = g;
            start[i] = (int64_t)s - (int64_t)ub[(size_t)g];
            end[i] = (int64_t)e - (int64_t)ub[(size_t)g];
        }
    return CB_OK;
}

int cb_cover_import(cb_ctx *ctx, int64_t n_probes, int32_t n_genomes, const int64_t *genome_len,
                    int64_t n_intervals, const int64_t *probe_id, const int32_t *genome,
                    const int64_t *start, const int64_t *end, cb_cover **out)
{
    if (!ctx) return CB_ERR_ARG;
    ctx->launches = 0;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    return cb_cover_import_impl(ctx, n_probes, n_genomes, genome_len, n_intervals, probe_id, genome, start, end, out);
}

int cb_setcover(cb_ctx *ctx, const cb_cover *cover, const int32_t *ranks, const double *universe_p,
                int64_t *sel_ids, int64_t *n_sel, cb_stats *stats)
{
    if (!ctx) return CB_ERR_ARG;
    if (stats) memset(stats, 0, sizeof *stats);
    ctx->launches = 0;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    return cb_setcover_impl(ctx, cover, nullptr, ranks, universe_p, sel_ids, n_sel, stats);
}

int cb_setcover_costs(cb_ctx *ctx, const cb_cover *cover, const double *costs, const int32_t *ranks,
                      const double *universe_p, int64_t *sel_ids, int64_t *n_sel, cb_stats *stats)
{
    if (!ctx) return CB_ERR_ARG;
    if (stats) memset(stats, 0, sizeof *stats);
    ctx->launches = 0;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    return cb_setcover_impl(ctx, cover, costs, ranks, universe_p, sel_ids, n_sel, stats);
}

int cb_minhash_neardup(cb_ctx *ctx, const uint8_t *ascii, const int64_t *probe_off, int64_t n_probes,
                       const uint32_t *a, const uint32_t *b, int32_t n_tables, int32_t k_concat,
                       int32_t kmer_size, double dist_thres, uint8_t *keep, cb_stats *stats)
{
    if (!ctx) return CB_ERR_ARG;
    if (stats) memset(stats, 0, sizeof *stats);
    ctx->launches = 0;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    return cb_minhash_neardup_impl(ctx, ascii, probe_off, n_probes, a, b, n_tables, k_concat, kmer_size,
                                   dist_thres, keep, stats);
}

int cb_neardup_filter(cb_ctx *ctx, const uint8_t *ascii, const int64_t *probe_off, int64_t n_probes, int32_t family,
                      const uint32_t *a, const uint32_t *b, const int32_t *positions, int32_t n_tables,
                      int32_t k_concat, int32_t kmer_size, double dist_thres, int64_t *kept_first_idx,
                      int64_t *n_kept, int64_t *n_distinct, cb_stats *stats)
{
    if (!ctx) return CB_ERR_ARG;
    if (stats) memset(stats, 0, sizeof *stats);
    ctx->launches = 0;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    return cb_neardup_filter_impl(ctx, ascii, probe_off, n_probes, family, a, b, positions, n_tables, k_concat,
                                  kmer_size, dist_thres, kept_first_idx, n_kept, n_distinct, stats);
}

int cb_group_duplicates(cb_ctx *ctx, const uint8_t *ascii, const int64_t *probe_off, int64_t n_probes,
                        int64_t *first_idx, int32_t *count, int64_t *n_distinct, cb_stats *stats)
{
    if (!ctx) return CB_ERR_ARG;
    if (stats) memset(stats, 0, sizeof *stats);
    ctx->launches = 0;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    return cb_group_duplicates_impl(ctx, ascii, probe_off, n_probes, first_idx, count, n_distinct, stats);
}

int cb_hamming_neardup(cb_ctx *ctx, const uint8_t *ascii, const int64_t *probe_off, int64_t n_probes,
                       const int32_t *positions, int32_t n_tables, int32_t k_concat, int32_t dist_thres,
                       uint8_t *keep, cb_stats *stats)
{
    if (!ctx) return CB_ERR_ARG;
    if (stats) memset(stats, 0, sizeof *stats);
    ctx->launches = 0;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    return cb_hamming_neardup_impl(ctx, ascii, probe_off, n_probes, positions, n_tables, k_concat,
                                   dist_thres, keep, stats);
}

int cb_sketch_sequences(cb_ctx *ctx, const uint8_t *ascii, const int64_t *seq_off, int64_t n_seqs, int32_t kmer_size,
                        int32_t N, uint64_t a, uint64_t b, cb_sketches **out, cb_stats *stats)
{
    if (!ctx) return CB_ERR_ARG;
    if (stats) memset(stats, 0, sizeof *stats);
    ctx->launches = 0;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    return cb_sketch_sequences_impl(ctx, ascii, seq_off, n_seqs, kmer_size, N, a, b, out, stats);
}

int cb_sketches_import(cb_ctx *ctx, const uint32_t *sig, int64_t n_seqs, int32_t N, cb_sketches **out)
{
    if (!ctx) return CB_ERR_ARG;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    return cb_sketch_import_impl(ctx, sig, n_seqs, N, out);
}

int cb_sketches_export(cb_ctx *ctx, const cb_sketches *sk, uint32_t *sig)
{
    if (!ctx) return CB_ERR_ARG;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    return cb_sketches_export_impl(ctx, sk, sig);
}

int cb_sketch_dist_rows(cb_ctx *ctx, const cb_sketches *sk, const int64_t *rows, int64_t n_rows, double *out)
{
    if (!ctx) return CB_ERR_ARG;
    ctx->launches = 0;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    return cb_sketch_dist_rows_impl(ctx, sk, rows, n_rows, out);
}

int cb_sketch_near_rows(cb_ctx *ctx, const cb_sketches *sk, const int64_t *rows, int64_t n_rows, double threshold,
                        int64_t *row_off, uint32_t **idx, double **dist)
{
    if (!ctx) return CB_ERR_ARG;
    ctx->launches = 0;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    return cb_sketch_near_rows_impl(ctx, sk, rows, n_rows, threshold, row_off, idx, dist);
}

int cb_sketch_dist_condensed(cb_ctx *ctx, const cb_sketches *sk, float *out)
{
    if (!ctx) return CB_ERR_ARG;
    ctx->launches = 0;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    return cb_sketch_dist_condensed_impl(ctx, sk, out);
}

}  // extern "C"
