// Multi-GPU exchange for probe sharding inside one grouping: every rank scans all
// target genomes against its own contiguous block of probes (stage A is embarrassingly parallel
// over probes), then the per-probe coverage intervals are all-gathered over NCCL (NVLink /
// NVSwitch) so that every rank holds the full cover and runs the identical greedy selection.
// This is the one real exchange step of the path; groupings themselves shard with no
// communication at all (catch_b200/parallel.py).
//
// NCCL is resolved at run time (dlopen of libnccl.so.2) so that single-GPU use neither needs nor
// loads it, and so that a process that already carries an NCCL (e.g. torch's) shares that copy.
// The handful of NCCL types the calls below need are declared here (their layout and values are
// part of NCCL 2's stable ABI), so building the library does not need the NCCL headers either.
#include <dlfcn.h>

#include <cstring>

#include "internal.cuh"

typedef struct { char internal[128]; } ncclUniqueId;      // NCCL_UNIQUE_ID_BYTES
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;                                   // ncclSuccess == 0
typedef int ncclDataType_t;
enum { ncclSuccess = 0, ncclInt64 = 4, ncclUint64 = 5 };

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi *nccl_api(cb_ctx *ctx)
{
    static NcclApi api;
    if (api.handle) return &api;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        ctx->err = std::string("cannot load libnccl.so.2: ") + dlerror();
        return nullptr;
    }
#define LOAD(field, sym)                                           \
    api.field = (decltype(api.field))dlsym(h, sym);                \
    if (!api.field) { ctx->err = "libnccl lacks " sym; return nullptr; }
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(AllGather, "ncclAllGather")
    LOAD(Broadcast, "ncclBroadcast")
    LOAD(GroupStart, "ncclGroupStart")
    LOAD(GroupEnd, "ncclGroupEnd")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    api.handle = h;
    return &api;
}

#define CB_NCCL(ctx, api, call)                                                                      \
    do {                                                                                             \
        ncclResult_t r__ = (call);                                                                   \
        if (r__ != ncclSuccess) {                                                                    \
            (ctx)->err = std::string(#call " failed: ") + (api)->GetErrorString(r__);                \
            return CB_ERR_COMM;                                                                      \
        }                                                                                            \
    } while (0)

__global__ void rebase_offsets_kernel(int64_t *off, int64_t n, int64_t base)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        off[i] += base;
}

}  // namespace

extern "C" {

int cb_comm_unique_id(cb_ctx *ctx, uint8_t out[128])
{
    if (!ctx || !out) return CB_ERR_ARG;
    NcclApi *api = nccl_api(ctx);
    if (!api) return CB_ERR_COMM;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    CB_NCCL(ctx, api, api->GetUniqueId(&id));
    memcpy(out, &id, 128);
    return CB_OK;
}

int cb_comm_init(cb_ctx *ctx, const uint8_t id_bytes[128], int32_t rank, int32_t n_ranks)
{
    if (!ctx || !id_bytes || rank < 0 || rank >= n_ranks) return cb_fail(ctx, CB_ERR_ARG, "bad communicator arguments");
    NcclApi *api = nccl_api(ctx);
    if (!api) return CB_ERR_COMM;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id_bytes, 128);
    ncclComm_t comm;
    CB_NCCL(ctx, api, api->CommInitRank(&comm, n_ranks, id, rank));
    ctx->comm = (void *)comm;
    ctx->rank = rank;
    ctx->n_ranks = n_ranks;
    return CB_OK;
}

int cb_comm_destroy(cb_ctx *ctx)
{
    if (!ctx || !ctx->comm) return CB_OK;
    NcclApi *api = nccl_api(ctx);
    if (api) api->CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
    return CB_OK;
}

int cb_cover_allgather(cb_ctx *ctx, const cb_cover *local, int64_t probe_lo, int64_t n_probes_total,
                       cb_cover **out)
{
    if (!ctx || !local || !out) return cb_fail(ctx, CB_ERR_ARG, "null argument");
    if (!ctx->comm) return cb_fail(ctx, CB_ERR_STATE, "cb_comm_init has not been called");
    NcclApi *api = nccl_api(ctx);
    if (!api) return CB_ERR_COMM;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    cudaStream_t st = ctx->stream;
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    const int R = ctx->n_ranks;

    // where every rank's shard starts, and how many intervals it holds
    DevBuf<int64_t> d_meta;
    CB_CUDA(ctx, d_meta.alloc((size_t)3 * (R + 1)));
    int64_t mine[3] = {probe_lo, local->n_intervals, (int64_t)local->max_interval_len};
    CB_CUDA(ctx, cudaMemcpyAsync(d_meta.p + 3 * R, mine, sizeof mine, cudaMemcpyHostToDevice, st));
    CB_NCCL(ctx, api, api->AllGather(d_meta.p + 3 * R, d_meta.p, 3, ncclInt64, comm, st));
    std::vector<int64_t> meta((size_t)3 * R);
    CB_CUDA(ctx, cudaMemcpyAsync(meta.data(), d_meta.p, sizeof(int64_t) * 3 * R, cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    std::vector<int64_t> p_lo((size_t)R + 1, 0), i_lo((size_t)R + 1, 0);
    uint32_t max_len = 0;
    bool tiles = meta[0] == 0;
    for (int r = 0; r < R; r++) {
        p_lo[(size_t)r] = meta[(size_t)3 * r];
        i_lo[(size_t)r + 1] = i_lo[(size_t)r] + meta[(size_t)3 * r + 1];
        if ((uint32_t)meta[(size_t)3 * r + 2] > max_len) max_len = (uint32_t)meta[(size_t)3 * r + 2];
        if (r && p_lo[(size_t)r] < p_lo[(size_t)r - 1]) tiles = false;
    }
    p_lo[(size_t)R] = n_probes_total;
    if (!tiles || p_lo[(size_t)R - 1] > n_probes_total)
        return cb_fail(ctx, CB_ERR_ARG, "probe shards do not tile [0, n_probes_total) in rank order");
    // the local cover either holds just the shard's rows, or (cb_coverage_range) all n_probes_total rows
    // with the ones outside the shard empty
    const int64_t my_rows = p_lo[(size_t)ctx->rank + 1] - p_lo[(size_t)ctx->rank];
    const bool global_ids = local->n_probes == n_probes_total && local->n_probes != my_rows;
    if (!global_ids && local->n_probes != my_rows)
        return cb_fail(ctx, CB_ERR_ARG, "local cover does not match the rank's shard");
    const int64_t *local_off = local->d_iv_off + (global_ids ? probe_lo : 0);
    const int64_t P = n_probes_total, E = i_lo[(size_t)R];

    cb_cover *cov = new cb_cover();
    struct Guard { cb_cover *c; ~Guard() { if (c) cb_cover_free(c); } } guard{cov};
    cov->ctx = ctx;
    cov->n_probes = P;
    cov->n_genomes = local->n_genomes;
    cov->n_intervals = E;
    cov->universe_bits = local->universe_bits;
    cov->max_interval_len = max_len;
    cov->h_ubase = local->h_ubase;
    cov->h_genome_len = local->h_genome_len;
    CB_CUDA(ctx, cb_dev_alloc(st, (void **)&cov->d_ubase, sizeof(uint32_t) * (size_t)(cov->n_genomes + 1)));
    CB_CUDA(ctx, cudaMemcpyAsync(cov->d_ubase, local->d_ubase, sizeof(uint32_t) * (size_t)(cov->n_genomes + 1),
                                 cudaMemcpyDeviceToDevice, st));
    CB_CUDA(ctx, cb_dev_alloc(st, (void **)&cov->d_iv_off, sizeof(int64_t) * (size_t)(P + 1)));
    CB_CUDA(ctx, cb_dev_alloc(st, (void **)&cov->d_iv, sizeof(uint2) * (size_t)(E ? E : 1)));

    // every rank broadcasts its slice into place (grouped: one fused NCCL operation)
    CB_NCCL(ctx, api, api->GroupStart());
    for (int r = 0; r < R; r++) {
        const int64_t np = p_lo[(size_t)r + 1] - p_lo[(size_t)r], ni = meta[(size_t)3 * r + 1];
        if (np > 0)
            CB_NCCL(ctx, api, api->Broadcast(local_off, cov->d_iv_off + p_lo[(size_t)r], (size_t)np, ncclInt64, r, comm, st));
        if (ni > 0)
            CB_NCCL(ctx, api, api->Broadcast(local->d_iv, cov->d_iv + i_lo[(size_t)r], (size_t)ni, ncclUint64, r, comm, st));
    }
    CB_NCCL(ctx, api, api->GroupEnd());
    for (int r = 0; r < R; r++) {
        const int64_t np = p_lo[(size_t)r + 1] - p_lo[(size_t)r];
        if (np > 0 && i_lo[(size_t)r] != 0) {
            rebase_offsets_kernel<<<ctx->sm_count, 256, 0, st>>>(cov->d_iv_off + p_lo[(size_t)r], np, i_lo[(size_t)r]);
            ctx->launches++;
        }
    }
    CB_CUDA(ctx, cudaMemcpyAsync(cov->d_iv_off + P, &E, sizeof(int64_t), cudaMemcpyHostToDevice, st));
    CB_CUDA(ctx, cudaGetLastError());
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    guard.c = nullptr;
    *out = cov;
    return CB_OK;
}

// ---------------------------------------------------------------------------------------
// Exchange areas of the sharded set cover (rounds.cu): one cudaMalloc'ed block per rank that the
// other ranks map into their address space (CUDA IPC across processes, plain pointers inside one
// process), so that the persistent kernels of all ranks can store into and load from each other's
// memory over NVLink.
// ---------------------------------------------------------------------------------------
static void exchange_detach(cb_ctx *ctx)
{
    for (int r = 0; r < CB_MAX_RANKS; r++) {
        if (ctx->xpeer[r] && ctx->xpeer_ipc[r]) cudaIpcCloseMemHandle(ctx->xpeer[r]);
        ctx->xpeer[r] = nullptr;
        ctx->xpeer_ipc[r] = false;
    }
    ctx->xn_ranks = 1;
    ctx->xrank = 0;
}

int cb_exchange_alloc(cb_ctx *ctx, int64_t bytes)
{
    if (!ctx || bytes < 0) return cb_fail(ctx, CB_ERR_ARG, "bad argument");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    exchange_detach(ctx);
    if (ctx->xarea) CB_CUDA(ctx, cudaFree(ctx->xarea));
    ctx->xarea = nullptr;
    ctx->xarea_bytes = 0;
    ctx->xarea_poisoned = false;
    if (bytes == 0) return CB_OK;
    const size_t want = ((size_t)bytes + (1u << 21) - 1) & ~(size_t)((1u << 21) - 1);
    CB_CUDA(ctx, cudaMalloc((void **)&ctx->xarea, want));
    CB_CUDA(ctx, cudaMemset(ctx->xarea, 0, want));
    ctx->xarea_bytes = want;
    return CB_OK;
}

int64_t cb_exchange_bytes(cb_ctx *ctx) { return ctx ? (int64_t)ctx->xarea_bytes : 0; }

int cb_exchange_handle(cb_ctx *ctx, uint8_t out[64], uint64_t *address)
{
    if (!ctx || !out) return cb_fail(ctx, CB_ERR_ARG, "null argument");
    if (!ctx->xarea) return cb_fail(ctx, CB_ERR_STATE, "cb_exchange_alloc has not been called");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CB_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->xarea));
    memcpy(out, &h, 64);
    if (address) *address = (uint64_t)(uintptr_t)ctx->xarea;
    return CB_OK;
}

int cb_exchange_attach(cb_ctx *ctx, int32_t rank, int32_t n_ranks, const uint8_t *handles, const uint64_t *addresses,
                       int32_t grid_limit)
{
    if (!ctx || rank < 0 || rank >= n_ranks || n_ranks > CB_MAX_RANKS || (!handles && !addresses))
        return cb_fail(ctx, CB_ERR_ARG, "bad exchange arguments");
    if (!ctx->xarea) return cb_fail(ctx, CB_ERR_STATE, "cb_exchange_alloc has not been called");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    exchange_detach(ctx);
    for (int r = 0; r < n_ranks; r++) {
        if (r == rank) { ctx->xpeer[r] = ctx->xarea; continue; }
        if (addresses) {                       // same process: the peer's device pointer is valid here
            ctx->xpeer[r] = (unsigned char *)(uintptr_t)addresses[r];
        } else {
            cudaIpcMemHandle_t h;
            memcpy(&h, handles + 64 * (size_t)r, 64);
            void *p = nullptr;
            CB_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            ctx->xpeer[r] = (unsigned char *)p;
            ctx->xpeer_ipc[r] = true;
        }
    }
    // a fresh epoch on every rank: the areas were zeroed by cb_exchange_alloc or are re-zeroed here
    CB_CUDA(ctx, cudaMemset(ctx->xarea, 0, 4096));
    ctx->xrank = rank;
    ctx->xn_ranks = n_ranks;
    ctx->xgrid_limit = grid_limit;
    ctx->xarea_poisoned = false;
    return CB_OK;
}

int cb_exchange_required(cb_ctx *ctx, const cb_cover *cover, int64_t *bytes)
{
    if (!ctx || !cover || !bytes) return cb_fail(ctx, CB_ERR_ARG, "null argument");
    *bytes = cb_rounds_exchange_bytes(cover);
    return CB_OK;
}

int cb_setcover_sharded(cb_ctx *ctx, const cb_cover *cover, int64_t probe_lo, int64_t probe_hi, const int32_t *ranks,
                        int64_t *sel_ids, int64_t *n_sel, cb_stats *stats)
{
    if (!ctx) return CB_ERR_ARG;
    if (!cover || !sel_ids || !n_sel) return cb_fail(ctx, CB_ERR_ARG, "null argument");
    if (stats) memset(stats, 0, sizeof *stats);
    ctx->launches = 0;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    *n_sel = 0;
    return cb_setcover_rounds_impl(ctx, cover, probe_lo, probe_hi, ranks, true, sel_ids, n_sel, stats);
}

int cb_setcover_sharded_begin(cb_ctx *ctx, const cb_cover *cover, int64_t probe_lo, int64_t probe_hi,
                              const int32_t *ranks, cb_job **job)
{
    if (!ctx) return CB_ERR_ARG;
    if (!cover || !job) return cb_fail(ctx, CB_ERR_ARG, "null argument");
    *job = nullptr;
    ctx->launches = 0;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    cb_rounds_job *j = nullptr;
    const int rc = cb_rounds_begin_impl(ctx, cover, probe_lo, probe_hi, ranks, &j);
    *job = reinterpret_cast<cb_job *>(j);
    return rc;
}

int cb_setcover_sharded_end(cb_ctx *ctx, cb_job *job, int64_t *sel_ids, int64_t *n_sel, cb_stats *stats)
{
    if (!ctx) return CB_ERR_ARG;
    if (!job || !sel_ids || !n_sel) return cb_fail(ctx, CB_ERR_ARG, "null argument");
    if (stats) memset(stats, 0, sizeof *stats);
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_tls_stream = ctx->stream;
    *n_sel = 0;
    return cb_rounds_end_impl(reinterpret_cast<cb_rounds_job *>(job), sel_ids, n_sel, stats);
}

}  // extern "C"
