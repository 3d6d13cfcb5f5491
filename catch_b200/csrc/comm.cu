// Multi-GPU exchange for probe sharding inside one grouping: every rank scans all
// target genomes against its own contiguous block of probes (stage A is embarrassingly parallel
// over probes), then the per-probe coverage intervals are all-gathered over NCCL (NVLink /
// NVSwitch) so that every rank holds the full cover and runs the identical greedy selection.
// This is the one real exchange step of the path; groupings themselves shard with no
// communication at all (catch_b200/parallel.py).
//
// NCCL is resolved at run time (dlopen of libnccl.so.2) so that single-GPU use neither needs nor
// loads it, and so that a process that already carries an NCCL (e.g. torch's) shares that copy.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "internal.cuh"

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

    // sizes of every rank's shard
    DevBuf<int64_t> d_meta;
    CB_CUDA(ctx, d_meta.alloc((size_t)3 * (R + 1)));
    int64_t mine[3] = {local->n_probes, local->n_intervals, (int64_t)local->max_interval_len};
    CB_CUDA(ctx, cudaMemcpyAsync(d_meta.p + 3 * R, mine, sizeof mine, cudaMemcpyHostToDevice, st));
    CB_NCCL(ctx, api, api->AllGather(d_meta.p + 3 * R, d_meta.p, 3, ncclInt64, comm, st));
    std::vector<int64_t> meta((size_t)3 * R);
    CB_CUDA(ctx, cudaMemcpyAsync(meta.data(), d_meta.p, sizeof(int64_t) * 3 * R, cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    std::vector<int64_t> p_lo((size_t)R + 1, 0), i_lo((size_t)R + 1, 0);
    uint32_t max_len = 0;
    for (int r = 0; r < R; r++) {
        p_lo[(size_t)r + 1] = p_lo[(size_t)r] + meta[(size_t)3 * r];
        i_lo[(size_t)r + 1] = i_lo[(size_t)r] + meta[(size_t)3 * r + 1];
        if ((uint32_t)meta[(size_t)3 * r + 2] > max_len) max_len = (uint32_t)meta[(size_t)3 * r + 2];
    }
    if (p_lo[(size_t)R] != n_probes_total || p_lo[(size_t)ctx->rank] != probe_lo)
        return cb_fail(ctx, CB_ERR_ARG, "probe shards do not tile [0, n_probes_total) in rank order");
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
        const int64_t np = meta[(size_t)3 * r], ni = meta[(size_t)3 * r + 1];
        if (np > 0)
            CB_NCCL(ctx, api, api->Broadcast(local->d_iv_off, cov->d_iv_off + p_lo[(size_t)r], (size_t)np, ncclInt64, r, comm, st));
        if (ni > 0)
            CB_NCCL(ctx, api, api->Broadcast(local->d_iv, cov->d_iv + i_lo[(size_t)r], (size_t)ni, ncclUint64, r, comm, st));
    }
    CB_NCCL(ctx, api, api->GroupEnd());
    for (int r = 0; r < R; r++) {
        const int64_t np = meta[(size_t)3 * r];
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

}  // extern "C"
