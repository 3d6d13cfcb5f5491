// Device-wide exclusive prefix sum (u32 counts -> i64 offsets) used to build every CSR on the
// path (seed-index buckets, per-probe range lists, interval lists, position blocks).
// Three launches: per-chunk sums, scan of the chunk sums by one block, per-chunk rescan.
#include "internal.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ unsigned long long warp_incl_scan(unsigned long long v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Inclusive scan across a block of SCAN_THREADS threads; returns this thread's inclusive value
// and the block total through `total`.
__device__ __forceinline__ unsigned long long block_incl_scan(unsigned long long v,
                                                              unsigned long long *s_warp,
                                                              unsigned long long &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = warp_incl_scan(v);
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = lane < (SCAN_THREADS / 32) ? s_warp[lane] : 0ull;
        w = warp_incl_scan(w);
        if (lane < (SCAN_THREADS / 32)) s_warp[lane] = w;
    }
    __syncthreads();
    unsigned long long base = warp ? s_warp[warp - 1] : 0ull;
    total = s_warp[SCAN_THREADS / 32 - 1];
    __syncthreads();
    return inc + base;
}

__global__ void __launch_bounds__(SCAN_THREADS)
chunk_sums_kernel(const uint32_t *__restrict__ in, int64_t n, unsigned long long *__restrict__ sums)
{
    __shared__ unsigned long long s_warp[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK;
    unsigned long long v = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        int64_t idx = base + (int64_t)i * SCAN_THREADS + threadIdx.x;
        if (idx < n) v += in[idx];
    }
    unsigned long long total;
    block_incl_scan(v, s_warp, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_sums_kernel(unsigned long long *sums, int64_t n_chunks, int64_t *out_total)
{
    __shared__ unsigned long long s_warp[SCAN_THREADS / 32];
    unsigned long long carry = 0;
    for (int64_t base = 0; base < n_chunks; base += SCAN_THREADS) {
        int64_t idx = base + threadIdx.x;
        unsigned long long v = idx < n_chunks ? sums[idx] : 0ull;
        unsigned long long total;
        unsigned long long inc = block_incl_scan(v, s_warp, total);
        if (idx < n_chunks) sums[idx] = carry + inc - v;     // exclusive
        carry += total;
    }
    if (threadIdx.x == 0) *out_total = (int64_t)carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
chunk_rescan_kernel(const uint32_t *__restrict__ in, int64_t n,
                    const unsigned long long *__restrict__ sums, int64_t *__restrict__ out,
                    const int64_t *__restrict__ total)
{
    __shared__ unsigned long long s_warp[SCAN_THREADS / 32];
    // thread t owns SCAN_ITEMS consecutive items so the scan is a single block scan of the
    // per-thread sums followed by a serial fix-up
    const int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        int64_t idx = base + i;
        v[i] = idx < n ? in[idx] : 0u;
        s += v[i];
    }
    unsigned long long tot;
    unsigned long long inc = block_incl_scan(s, s_warp, tot);
    unsigned long long run = sums[blockIdx.x] + inc - s;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        int64_t idx = base + i;
        if (idx < n) out[idx] = (int64_t)run;
        run += v[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = *total;
}

__global__ void write_zero_total(int64_t *out) { out[0] = 0; }

}  // namespace

int cb_exclusive_scan_u32_to_i64(cb_ctx *ctx, const uint32_t *d_in, int64_t *d_out, int64_t n,
                                 int64_t *h_total)
{
    if (n <= 0) {
        write_zero_total<<<1, 1, 0, ctx->stream>>>(d_out);
        ctx->launches++;
        CB_CUDA(ctx, cudaGetLastError());
        if (h_total) *h_total = 0;
        return CB_OK;
    }
    const int64_t n_chunks = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    DevBuf<unsigned long long> sums;
    DevBuf<int64_t> total;
    CB_CUDA(ctx, sums.alloc((size_t)n_chunks));
    CB_CUDA(ctx, total.alloc(1));
    chunk_sums_kernel<<<(unsigned)n_chunks, SCAN_THREADS, 0, ctx->stream>>>(d_in, n, sums.p);
    scan_sums_kernel<<<1, SCAN_THREADS, 0, ctx->stream>>>(sums.p, n_chunks, total.p);
    chunk_rescan_kernel<<<(unsigned)n_chunks, SCAN_THREADS, 0, ctx->stream>>>(d_in, n, sums.p, d_out, total.p);
    ctx->launches += 3;
    CB_CUDA(ctx, cudaGetLastError());
    if (h_total) {
        CB_CUDA(ctx, cudaMemcpyAsync(h_total, total.p, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
        // sums/total are freed when this function returns; cudaFree synchronises implicitly
    }
    return CB_OK;
}
