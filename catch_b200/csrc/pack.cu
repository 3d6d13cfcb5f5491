// K1: ASCII -> bit planes.  Replaces the per-hit str -> np.array('U1') conversion of the
// reference (probe.py:1074) and Probe.from_str (probe.py:344): every sequence is packed once.
// A warp packs 64 bases at a time: each lane loads two bytes (coalesced), maps them through the
// 256-entry code table in shared memory, and one __ballot_sync per plane yields 32 bits of that
// plane directly.
#include "internal.cuh"

namespace {

constexpr int PACK_THREADS = 256;

__global__ void __launch_bounds__(PACK_THREADS)
pack_targets_kernel(const uint8_t *__restrict__ ascii, int64_t total, const uint8_t *__restrict__ lut,
                    int bits, uint64_t *__restrict__ planes, int64_t plane_words)
{
    __shared__ uint8_t s_lut[256];
    s_lut[threadIdx.x & 255] = lut[threadIdx.x & 255];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * PACK_THREADS + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * PACK_THREADS) >> 5;
    const int64_t n_words = (total + 63) >> 6;
    for (int64_t w = warp; w < n_words; w += n_warps) {
        const int64_t g0 = w << 6;
        const int64_t i0 = g0 + lane, i1 = g0 + 32 + lane;
        const unsigned c0 = i0 < total ? s_lut[ascii[i0]] : 0u;
        const unsigned c1 = i1 < total ? s_lut[ascii[i1]] : 0u;
        for (int b = 0; b < bits; b++) {
            const unsigned lo = __ballot_sync(0xffffffffu, (c0 >> b) & 1u);
            const unsigned hi = __ballot_sync(0xffffffffu, (c1 >> b) & 1u);
            if (lane == b)
                planes[(int64_t)b * plane_words + (CB_FRONT_PAD >> 6) + w] =
                    ((uint64_t)hi << 32) | (uint64_t)lo;
        }
    }
}

__global__ void __launch_bounds__(PACK_THREADS)
pack_probes_kernel(const uint8_t *__restrict__ ascii, const int64_t *__restrict__ off, int gap, int64_t n_probes,
                   const uint8_t *__restrict__ lut, int bits, int nw, uint64_t *__restrict__ words,
                   int32_t *__restrict__ lens)
{
    __shared__ uint8_t s_lut[256];
    s_lut[threadIdx.x & 255] = lut[threadIdx.x & 255];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * PACK_THREADS + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * PACK_THREADS) >> 5;
    for (int64_t p = warp; p < n_probes; p += n_warps) {
        const int64_t beg = off[p];
        const int len = (int)(off[p + 1] - beg) - gap;      // gap: separator bytes between probes
        if (lane == 0) lens[p] = len;
        uint64_t *dst = words + p * (int64_t)bits * nw;
        for (int w = 0; w < nw; w++) {
            const int j0 = w * 64 + lane, j1 = j0 + 32;
            const unsigned c0 = j0 < len ? s_lut[ascii[beg + j0]] : 0u;
            const unsigned c1 = j1 < len ? s_lut[ascii[beg + j1]] : 0u;
            for (int b = 0; b < bits; b++) {
                const unsigned lo = __ballot_sync(0xffffffffu, (c0 >> b) & 1u);
                const unsigned hi = __ballot_sync(0xffffffffu, (c1 >> b) & 1u);
                if (lane == b) dst[b * nw + w] = ((uint64_t)hi << 32) | (uint64_t)lo;
            }
        }
    }
}

}  // namespace

int cb_launch_pack_targets(cb_ctx *ctx, const uint8_t *d_ascii, int64_t total, const uint8_t *d_lut,
                           int bits, uint64_t *d_planes, int64_t plane_words)
{
    CB_CUDA(ctx, cudaMemsetAsync(d_planes, 0, sizeof(uint64_t) * (size_t)bits * (size_t)plane_words, ctx->stream));
    if (total == 0) return CB_OK;
    int64_t n_words = (total + 63) >> 6;
    int64_t blocks = (n_words * 32 + PACK_THREADS - 1) / PACK_THREADS;
    int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    pack_targets_kernel<<<(unsigned)blocks, PACK_THREADS, 0, ctx->stream>>>(d_ascii, total, d_lut, bits,
                                                                            d_planes, plane_words);
    ctx->launches++;
    CB_CUDA(ctx, cudaGetLastError());
    return CB_OK;
}

int cb_launch_pack_probes(cb_ctx *ctx, const uint8_t *d_ascii, const int64_t *d_off, int gap, int64_t n_probes,
                          const uint8_t *d_lut, int bits, int nw, uint64_t *d_words, int32_t *d_len)
{
    if (n_probes == 0) return CB_OK;
    int64_t blocks = (n_probes * 32 + PACK_THREADS - 1) / PACK_THREADS;
    int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    pack_probes_kernel<<<(unsigned)blocks, PACK_THREADS, 0, ctx->stream>>>(d_ascii, d_off, gap, n_probes, d_lut,
                                                                           bits, nw, d_words, d_len);
    ctx->launches++;
    CB_CUDA(ctx, cudaGetLastError());
    return CB_OK;
}


// ---- which byte values occur (for the code table): per-CTA 256-entry presence in shared memory
namespace {
__global__ void __launch_bounds__(PACK_THREADS)
byte_presence_kernel(const uint8_t *__restrict__ buf, int64_t n, uint32_t *__restrict__ present)
{
    __shared__ uint32_t s_present[256];
    s_present[threadIdx.x] = 0u;
    __syncthreads();
    const int64_t tid = (int64_t)blockIdx.x * PACK_THREADS + threadIdx.x, nthr = (int64_t)gridDim.x * PACK_THREADS;
    // 16 bytes per load once the pointer is aligned; the ragged head and tail byte-wise
    const int64_t head = min(n, (int64_t)((16 - ((uintptr_t)buf & 15)) & 15));
    const int64_t n_vec = (n - head) >> 4;
    const uint4 *v = reinterpret_cast<const uint4 *>(buf + head);
    uint32_t last = 0xffffffffu;
    for (int64_t i = tid; i < n_vec; i += nthr) {
        const uint4 q = v[i];
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (w[j] == last) continue;                 // runs of the same 4 bytes are common
            last = w[j];
#pragma unroll
            for (int b = 0; b < 4; b++) s_present[(w[j] >> (8 * b)) & 0xffu] = 1u;
        }
    }
    for (int64_t i = tid; i < head; i += nthr) s_present[buf[i]] = 1u;
    for (int64_t i = head + (n_vec << 4) + tid; i < n; i += nthr) s_present[buf[i]] = 1u;
    __syncthreads();
    if (s_present[threadIdx.x]) present[threadIdx.x] = 1u;
}
}  // namespace

int cb_launch_byte_presence(cb_ctx *ctx, const uint8_t *d_buf, int64_t n, uint32_t *d_present)
{
    if (n <= 0) return CB_OK;
    int64_t blocks = (n / 16 + PACK_THREADS - 1) / PACK_THREADS;
    if (blocks < 1) blocks = 1;
    if (blocks > (int64_t)ctx->sm_count * 8) blocks = (int64_t)ctx->sm_count * 8;
    byte_presence_kernel<<<(unsigned)blocks, PACK_THREADS, 0, ctx->stream>>>(d_buf, n, d_present);
    ctx->launches++;
    CB_CUDA(ctx, cudaGetLastError());
    return CB_OK;
}

// ---- duplicate detection: 64-bit hash of every packed probe into an open-addressing table
namespace {
__global__ void dup_probe_kernel(const uint64_t *__restrict__ words, const int32_t *__restrict__ lens,
                                 int64_t n_probes, int wpp, unsigned long long *__restrict__ table,
                                 uint32_t mask, int *__restrict__ flag)
{
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_probes;
         p += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long h = 0x9E3779B97F4A7C15ull ^ (unsigned long long)lens[p];
        // every word is folded in through a full avalanche step (splitmix64 finaliser): with one multiply and one
        // shift only, a flipped top bit stays a top bit through the multiplication and two substitutions at bases
        // 63 and 96 of a 100-nt probe cancelled each other (3 of 16 V-All-shape groupings reported duplicates)
        for (int w = 0; w < wpp; w++) {
            h ^= words[p * wpp + w];
            h ^= h >> 30;
            h *= 0xBF58476D1CE4E5B9ull;
            h ^= h >> 27;
            h *= 0x94D049BB133111EBull;
            h ^= h >> 31;
        }
        if (h == 0ull) h = 1ull;
        uint32_t slot = (uint32_t)(h >> 13) & mask;
        for (;;) {
            const unsigned long long prev = atomicCAS(&table[slot], 0ull, h);
            if (prev == 0ull) break;
            if (prev == h) { *flag = 1; break; }
            slot = (slot + 1) & mask;
        }
    }
}
}  // namespace

int cb_probes_have_duplicates_impl(cb_ctx *ctx, const cb_probes *probes, int32_t *has_dup)
{
    *has_dup = 0;
    const int64_t P = probes->n_probes;
    if (P < 2) return CB_OK;
    int64_t cap = 1024;
    while (cap < 2 * P) cap <<= 1;
    DevBuf<unsigned long long> table;
    DevBuf<int> flag;
    CB_CUDA(ctx, table.alloc((size_t)cap));
    CB_CUDA(ctx, flag.alloc(1));
    CB_CUDA(ctx, cudaMemsetAsync(table.p, 0, sizeof(unsigned long long) * (size_t)cap, ctx->stream));
    CB_CUDA(ctx, cudaMemsetAsync(flag.p, 0, sizeof(int), ctx->stream));
    int64_t blocks = (P + PACK_THREADS - 1) / PACK_THREADS;
    if (blocks > (int64_t)ctx->sm_count * 16) blocks = (int64_t)ctx->sm_count * 16;
    dup_probe_kernel<<<(unsigned)blocks, PACK_THREADS, 0, ctx->stream>>>(probes->d_words, probes->d_len, P,
                                                                        probes->bits * probes->nw, table.p,
                                                                        (uint32_t)(cap - 1), flag.p);
    ctx->launches++;
    CB_CUDA(ctx, cudaGetLastError());
    int h = 0;
    CB_CUDA(ctx, cudaMemcpyAsync(&h, flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *has_dup = h;
    return CB_OK;
}


// ---------------------------------------------------------------------------------------
// Integer-throughput calibration for the roofline of the scan kernel (which is bound by the integer
// ALU pipe, not by HBM): every thread runs eight independent chains of LOP3 / SHF / IADD3 -- the
// instructions the mismatch-mask and extension code is made of -- and the achieved rate is reported
// as integer ALU instructions per second (thread level) over the whole chip.  One step of a chain is
// two instructions in the SASS: LOP3 (xor) and LEA.HI (shift + add fused).
// ---------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) intop_kernel(uint32_t *out, int iters, uint32_t seed)
{
    uint32_t x[8];
#pragma unroll
    for (int c = 0; c < 8; c++) x[c] = seed + threadIdx.x * 8u + c + blockIdx.x * 2048u;
    const uint32_t y = seed ^ 0x9e3779b9u;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < 8; c++) x[c] = (x[c] ^ y) + (x[c] >> 7);      // LOP3, LEA.HI
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) acc ^= x[c];
    if (acc == 0x12345u) out[0] = acc;                                     // keeps the chains alive
}
}  // namespace

int cb_intop_rate_impl(cb_ctx *ctx, double *ops_per_s)
{
    cudaStream_t st = ctx->stream;
    DevBuf<uint32_t> d_out;
    CB_CUDA(ctx, d_out.alloc(1));
    const int iters = 4096, grid = ctx->sm_count * 8;
    intop_kernel<<<grid, 256, 0, st>>>(d_out.p, 64, 1u);                   // warm-up
    EventTimer t(st);
    double best = 0.0;
    for (int rep = 0; rep < 3; rep++) {
        t.start();
        intop_kernel<<<grid, 256, 0, st>>>(d_out.p, iters, 2u + rep);
        t.stop();
        CB_CUDA(ctx, cudaGetLastError());
        const double ms = t.ms();
        const double ops = 2.0 * 8.0 * (double)iters * 256.0 * (double)grid;     // LOP3 + LEA.HI per step
        if (ms > 0 && ops / (ms * 1e-3) > best) best = ops / (ms * 1e-3);
    }
    ctx->launches += 4;
    *ops_per_s = best;
    return CB_OK;
}
