// Stage B: greedy multi-universe set cover (K5 universe build, K6 gains, K7 argmax, K8 apply).
//
// Replaces utils/set_cover.py:147-615 approx_multiuniverse(use_intervalsets=True) as called by
// filter/set_cover_filter.py:136-142 (cost 1 for every set).
//
//   universes[u] = union of every set's intervals in u (:302-320)      -> universe bit set U
//   num_left_to_cover[u] = |U_u| - int(|U_u| - p_u |U_u|) (:362-373)   -> host, from device popcounts
//   each pick: among the sets of the current rank, minimise cost / sum_u min(left_u, |s_u & U_u|)
//   = maximise the integer gain, smallest set id on ties (:393-433, :483-520); gain 0 everywhere
//   -> next rank (:522-526); U_u -= s_u, left_u = max(0, |U_u| - uncoverable_u) (:528-550);
//   stop when every left_u is 0 (:448).
//
// This file: the dispatcher and the general kernel (some p_u < 1, or non-unit costs): ONE persistent
// cooperative kernel for the whole loop, one pick per iteration, grid-wide barriers between the
// phases of a pick.  With p_u < 1 the min() clamp is active and gains are recomputed from U for every
// probe at every pick ("full" mode); with costs but p_u == 1 gains are kept exact incrementally.
// The default case (unit costs, every p_u == 1) is handled by rounds.cu: parallel rounds, optionally
// sharded over several GPUs.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "internal.cuh"

namespace {

constexpr int GREEDY_THREADS = 256;

struct GreedyParams {
    int64_t n_probes;
    int32_t n_genomes;
    int64_t n_intervals;
    const int64_t *iv_off;
    const uint2 *iv;
    const uint32_t *iv_genome;     // full mode only
    const uint32_t *ubase;
    unsigned long long *U;         // universe bit set
    int64_t u_words;
    uint32_t *gain;                // [n_probes]
    const uint32_t *rank_idx;      // [n_probes]
    int32_t n_ranks;
    // incremental mode
    const int64_t *blk_off;        // [n_blocks+1]
    const uint4 *blk_items;        // (start, end, probe, -)
    int64_t n_blocks;
    uint32_t max_len;
    unsigned long long *remaining; // total uncovered bits still to cover
    // full mode
    long long *u_size;             // [n_genomes] current |U_u|
    const long long *uncoverable;  // [n_genomes]
    unsigned int *n_left;          // [2] universes with left > 0
    int full_mode;
    // control / output
    unsigned long long *key;       // [2] argmax slots
    long long *sel;                // [n_probes] picks in order
    long long *n_sel;
    int *status;
    unsigned long long *barrier;   // grid barrier arrival counter
    unsigned long long *phase_ns;  // [4] time CTA 0 spent in: argmax, barrier, delta, barrier
    const double *costs;           // [n_probes] or nullptr (all 1): two-barrier kernel only
    uint32_t *idmin;               // [2] smallest id among the sets at the minimum ratio (cost mode)
};

// ---- K5: universe = union of all intervals
__global__ void universe_build_kernel(const uint2 *__restrict__ iv, int64_t n, unsigned long long *U)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint2 r = iv[i];
        if (r.x >= r.y) continue;
        const uint32_t w0 = r.x >> 6, w1 = (r.y - 1) >> 6;
        for (uint32_t w = w0; w <= w1; w++) {
            unsigned long long m = ~0ull;
            if (w == w0) m &= ~0ull << (r.x & 63);
            if (w == w1) m &= ~0ull >> (63 - ((r.y - 1) & 63));
            if ((U[w] & m) != m) atomicOr(&U[w], m);
        }
    }
}

// |U_u| for every genome: one warp per genome over its (64-aligned) word range
__global__ void universe_size_kernel(const unsigned long long *__restrict__ U, const uint32_t *__restrict__ ubase,
                                     int32_t n_genomes, long long *__restrict__ u_size)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = warp; u < n_genomes; u += n_warps) {
        const uint32_t w0 = ubase[u] >> 6, w1 = ubase[u + 1] >> 6;
        long long c = 0;
        for (uint32_t w = w0 + lane; w < w1; w += 32) c += __popcll(U[w]);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) u_size[u] = c;
    }
}

// genome of every interval (full mode) by binary search over ubase
__global__ void interval_genome_kernel(const uint2 *__restrict__ iv, int64_t n, const uint32_t *__restrict__ ubase,
                                       int32_t n_genomes, uint32_t *__restrict__ iv_genome)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t s = iv[i].x;
        int lo = 0, hi = n_genomes;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (ubase[mid] <= s) lo = mid; else hi = mid;
        }
        iv_genome[i] = (uint32_t)lo;
    }
}

// index of intervals by 64-position block of their start
template <bool SCATTER>
__global__ void block_index_kernel(const int64_t *__restrict__ iv_off, const uint2 *__restrict__ iv, int64_t n_probes,
                                   uint32_t *__restrict__ count, const int64_t *__restrict__ blk_off,
                                   uint32_t *__restrict__ cursor, uint4 *__restrict__ items)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < n_probes; p += n_warps) {
        for (int64_t i = iv_off[p] + lane; i < iv_off[p + 1]; i += 32) {
            const uint2 r = iv[i];
            const uint32_t b = r.x >> 6;
            if (!SCATTER) atomicAdd(&count[b], 1u);
            else {
                const uint32_t slot = atomicAdd(&cursor[b], 1u);
                items[blk_off[b] + slot] = make_uint4(r.x, r.y, (uint32_t)p, 0u);
            }
        }
    }
}

__device__ __forceinline__ uint32_t popcount_range_cg(const unsigned long long *U, uint32_t s, uint32_t e);

// initial gains: sum over (probe, genome) of min(left_u, |s_u & U_u|)
__device__ __forceinline__ void recompute_gains(const GreedyParams &G, int64_t gtid, int64_t gsize)
{
    for (int64_t p = gtid; p < G.n_probes; p += gsize) {
        uint32_t total = 0;
        int64_t i = G.iv_off[p];
        const int64_t e = G.iv_off[p + 1];
        while (i < e) {
            const uint32_t u = G.iv_genome[i];
            long long c = 0;
            while (i < e && G.iv_genome[i] == u) {
                const uint2 r = G.iv[i];
                c += popcount_range_cg(G.U, r.x, r.y);
                i++;
            }
            long long left = __ldcg(&G.u_size[u]) - G.uncoverable[u];
            if (left < 0) left = 0;
            total += (uint32_t)(c < left ? c : left);
        }
        G.gain[p] = total;
    }
}

__global__ void gains_init_kernel(const int64_t *__restrict__ iv_off, const uint2 *__restrict__ iv, int64_t n_probes,
                                  uint32_t *__restrict__ gain)
{
    // incremental mode: every bit of every interval is in U at the start
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < n_probes; p += n_warps) {
        uint32_t c = 0;
        for (int64_t i = iv_off[p] + lane; i < iv_off[p + 1]; i += 32) c += iv[i].y - iv[i].x;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) gain[p] = c;
    }
}

// ---- grid-wide barrier for the persistent kernel: monotone arrival counter, one arrival per
// CTA, volatile polling (no fence inside the loop), one fence after.  Mutable global data is read with
// __ldcg / atomics (L2), so no L1 invalidation is needed after the barrier.
__device__ __forceinline__ void grid_barrier(unsigned long long *counter, unsigned long long &target)
{
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1ull);
        while (*(volatile unsigned long long *)counter < target) { }
        __threadfence();
    }
    __syncthreads();
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ uint32_t popcount_range_cg(const unsigned long long *U, uint32_t s, uint32_t e)
{
    if (s >= e) return 0;
    const uint32_t w0 = s >> 6, w1 = (e - 1) >> 6;
    uint32_t c = 0;
    for (uint32_t w = w0; w <= w1; w++) {
        unsigned long long m = ~0ull;
        if (w == w0) m &= ~0ull << (s & 63);
        if (w == w1) m &= ~0ull >> (63 - ((e - 1) & 63));
        c += __popcll(__ldcg(U + w) & m);
    }
    return c;
}

// ---- the persistent greedy kernel
__global__ void __launch_bounds__(GREEDY_THREADS)
greedy_kernel(const GreedyParams G)
{
    __shared__ unsigned long long s_key[GREEDY_THREADS / 32];
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gsize = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    int cur_rank = 0;
    long long n_picks = 0;
    long long prev = -1;
    unsigned long long bar_target = 0;
    unsigned long long t_phase[4] = {0, 0, 0, 0}, t_last = 0;
    const bool timing = (gtid == 0);
    if (timing) t_last = globaltimer_ns();
    auto lap = [&](int i) {
        if (timing) {
            const unsigned long long t = globaltimer_ns();
            t_phase[i] += t - t_last;
            t_last = t;
        }
    };

    for (unsigned it = 0;; it++) {
        // ---- apply the previous pick: U -= s (K8), count what it newly covered
        if (prev >= 0) {
            const int64_t i0 = G.iv_off[prev], i1 = G.iv_off[prev + 1];
            for (int64_t i = i0 + gtid; i < i1; i += gsize) {
                const uint2 r = G.iv[i];
                if (r.x >= r.y) continue;
                const uint32_t w0 = r.x >> 6, w1 = (r.y - 1) >> 6;
                uint32_t c = 0;
                for (uint32_t w = w0; w <= w1; w++) {
                    unsigned long long m = ~0ull;
                    if (w == w0) m &= ~0ull << (r.x & 63);
                    if (w == w1) m &= ~0ull >> (63 - ((r.y - 1) & 63));
                    const unsigned long long old = atomicAnd(&G.U[w], ~m);
                    c += __popcll(old & m);
                }
                if (c) {
                    if (G.full_mode) atomicAdd((unsigned long long *)&G.u_size[G.iv_genome[i]], (unsigned long long)(-(long long)c));
                    else atomicAdd(G.remaining, (unsigned long long)(-(long long)c));
                }
            }
        }
        if (G.full_mode) {
            grid_barrier(G.barrier, bar_target);
            if (gtid == 0) G.n_left[(it + 1) & 1] = 0;
            recompute_gains(G, gtid, gsize);
            unsigned cnt = 0;
            for (int64_t u = gtid; u < G.n_genomes; u += gsize)
                if (__ldcg(&G.u_size[u]) - G.uncoverable[u] > 0) cnt++;
            if (cnt) atomicAdd(&G.n_left[it & 1], cnt);
            grid_barrier(G.barrier, bar_target);
        }
        // ---- K7 argmax over the current rank: max gain, smallest id.  With costs (set_cover.py:426:
        // ratio = float(cost) / gain, strict '<' over ascending ids) the key of this first pass is the
        // ratio alone -- the bitwise complement of its order-preserving integer image, so that the
        // LARGEST key is the smallest ratio; the smallest id at that ratio is found in a second pass.
        auto ratio_key = [&](int64_t p, uint32_t g) -> unsigned long long {
            const long long b = __double_as_longlong(G.costs[p] / (double)g);
            const unsigned long long u = (unsigned long long)b ^ ((unsigned long long)(b >> 63) | 0x8000000000000000ull);
            return ~u;
        };
        unsigned long long best = 0;
        for (int64_t p = gtid; p < G.n_probes; p += gsize) {
            const uint32_t g = __ldcg(&G.gain[p]);
            if (g && G.rank_idx[p] == (uint32_t)cur_rank) {
                const unsigned long long key = G.costs ? ratio_key(p, g)
                    : (((unsigned long long)g << 32) | (unsigned long long)(0xffffffffu - (uint32_t)p));
                best = key > best ? key : best;
            }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
            best = t > best ? t : best;
        }
        if (lane == 0) s_key[warp] = best;
        __syncthreads();
        if (warp == 0) {
            best = lane < GREEDY_THREADS / 32 ? s_key[lane] : 0ull;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
                best = t > best ? t : best;
            }
            if (lane == 0 && best) atomicMax(&G.key[it & 1], best);
        }
        lap(0);
        grid_barrier(G.barrier, bar_target);
        lap(1);

        // ---- decide
        const bool done = G.full_mode ? (__ldcg(&G.n_left[it & 1]) == 0) : (__ldcg(G.remaining) == 0ull);
        if (done) break;
        const unsigned long long key = __ldcg(&G.key[it & 1]);
        if (gtid == 0) {
            G.key[(it + 1) & 1] = 0ull;
            if (G.costs) G.idmin[(it + 1) & 1] = 0xffffffffu;     // also when the rank advances below
        }
        if (key == 0ull) {                      // nothing in this rank covers anything needed (:522-526)
            cur_rank++;
            prev = -1;
            if (cur_rank >= G.n_ranks) {
                if (gtid == 0) *G.status = CB_ERR_STATE;
                break;
            }
            grid_barrier(G.barrier, bar_target);
            continue;
        }
        long long w = (long long)(0xffffffffu - (uint32_t)(key & 0xffffffffull));
        if (G.costs) {
            // second pass: smallest id among the sets whose ratio equals the minimum
            uint32_t mine = 0xffffffffu;
            for (int64_t p = gtid; p < G.n_probes; p += gsize) {
                const uint32_t g = __ldcg(&G.gain[p]);
                if (g && G.rank_idx[p] == (uint32_t)cur_rank && ratio_key(p, g) == key) { mine = (uint32_t)p; break; }
            }
            if (mine != 0xffffffffu) atomicMin(&G.idmin[it & 1], mine);
            grid_barrier(G.barrier, bar_target);
            w = (long long)__ldcg(&G.idmin[it & 1]);
        }
        if (gtid == 0) G.sel[n_picks] = w;
        n_picks++;
        prev = w;

        // ---- K6 (incremental): take the winner's still-uncovered bits out of every overlapping
        // interval.  Intervals are indexed by the 64-position block of their start, blocks are
        // consecutive in blk_items, so the candidates of one winner interval are ONE contiguous
        // range of items; a CTA takes a winner interval, its threads stride over the range.
        if (!G.full_mode) {
            const int64_t i0 = G.iv_off[w], i1 = G.iv_off[w + 1];
            for (int64_t i = i0 + blockIdx.x; i < i1; i += gridDim.x) {
                uint2 r = G.iv[i];
                // narrow the winner interval to the span of its still-uncovered bits; most late
                // picks re-cover ground in most genomes and need no gain update there at all
                {
                    uint32_t first = 0xffffffffu, last = 0;
                    if (r.x < r.y) {
                        const uint32_t w0 = r.x >> 6, w1 = (r.y - 1) >> 6;
                        for (uint32_t w = w0; w <= w1; w++) {
                            unsigned long long m = ~0ull;
                            if (w == w0) m &= ~0ull << (r.x & 63);
                            if (w == w1) m &= ~0ull >> (63 - ((r.y - 1) & 63));
                            const unsigned long long v = __ldcg(G.U + w) & m;
                            if (v) {
                                if (first == 0xffffffffu) first = (w << 6) + (uint32_t)__ffsll((long long)v) - 1u;
                                last = (w << 6) + 63u - (uint32_t)__clzll((long long)v);
                            }
                        }
                    }
                    if (first == 0xffffffffu) continue;
                    r.x = first;
                    r.y = last + 1u;
                }
                const int64_t lo_pos = (int64_t)r.x - (int64_t)G.max_len + 1;
                const int64_t b_lo = (lo_pos > 0 ? lo_pos : 0) >> 6;
                int64_t b_hi = ((int64_t)r.y - 1) >> 6;
                if (b_hi >= G.n_blocks) b_hi = G.n_blocks - 1;
                const int64_t x0 = G.blk_off[b_lo], x1 = G.blk_off[b_hi + 1];
                // all item loads of a batch are issued before any is consumed (memory-level
                // parallelism: the phase is a chain of dependent DRAM/L2 round trips otherwise)
                constexpr int BATCH = 4;
                for (int64_t xb = x0 + threadIdx.x; xb < x1; xb += (int64_t)GREEDY_THREADS * BATCH) {
                    uint4 item[BATCH];
#pragma unroll
                    for (int u = 0; u < BATCH; u++) {
                        const int64_t x = xb + (int64_t)u * GREEDY_THREADS;
                        item[u] = x < x1 ? __ldg(G.blk_items + x) : make_uint4(0u, 0u, 0u, 0u);
                    }
                    uint32_t os[BATCH], oe[BATCH];
                    unsigned long long w0v[BATCH], w1v[BATCH], w2v[BATCH];
#pragma unroll
                    for (int u = 0; u < BATCH; u++) {
                        os[u] = max(item[u].x, r.x);
                        oe[u] = min(item[u].y, r.y);
                        w0v[u] = w1v[u] = w2v[u] = 0ull;
                        if (os[u] < oe[u]) {
                            // overlaps are at most max_len long; the common case spans <= 3 words
                            const uint32_t wa = os[u] >> 6, wb = (oe[u] - 1) >> 6;
                            w0v[u] = __ldcg(G.U + wa);
                            if (wb > wa) w1v[u] = __ldcg(G.U + wa + 1);
                            if (wb > wa + 1) w2v[u] = __ldcg(G.U + wa + 2);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < BATCH; u++) {
                        if (os[u] < oe[u]) {
                            const uint32_t wa = os[u] >> 6, wb = (oe[u] - 1) >> 6;
                            uint32_t dlt;
                            if (wb <= wa + 2) {
                                const unsigned long long mlo = ~0ull << (os[u] & 63);
                                const unsigned long long mhi = ~0ull >> (63 - ((oe[u] - 1) & 63));
                                if (wb == wa) dlt = __popcll(w0v[u] & mlo & mhi);
                                else if (wb == wa + 1) dlt = __popcll(w0v[u] & mlo) + __popcll(w1v[u] & mhi);
                                else dlt = __popcll(w0v[u] & mlo) + __popcll(w1v[u]) + __popcll(w2v[u] & mhi);
                            } else {
                                dlt = popcount_range_cg(G.U, os[u], oe[u]);
                            }
                            if (dlt) atomicSub(&G.gain[item[u].z], dlt);
                        }
                    }
                }
            }
        }
        lap(2);
        grid_barrier(G.barrier, bar_target);
        lap(3);
    }
    if (gtid == 0) {
        *G.n_sel = n_picks;
        for (int i = 0; i < 4; i++) G.phase_ns[i] = t_phase[i];
    }
}


}  // namespace

int cb_setcover_impl(cb_ctx *ctx, const cb_cover *cover, const double *costs, const int32_t *ranks,
                     const double *universe_p, int64_t *sel_ids, int64_t *n_sel, cb_stats *stats)
{
    if (!cover || !sel_ids || !n_sel) return cb_fail(ctx, CB_ERR_ARG, "null argument");
    cudaStream_t st = ctx->stream;
    const int64_t P = cover->n_probes;
    const int32_t NG = cover->n_genomes;
    const int64_t E = cover->n_intervals;
    *n_sel = 0;
    if (P == 0 || E == 0) {
        if (stats) stats->n_picks = 0;
        return CB_OK;
    }
    // non-unit costs (never produced by SetCoverFilter, set_cover_filter.py:759, but part of
    // approx_multiuniverse's contract) and p_u < 1 take the general kernel below
    bool use_costs = false;
    if (costs)
        for (int64_t p = 0; p < P; p++) {
            if (!(costs[p] >= 0.0) || costs[p] > 1e300) return cb_fail(ctx, CB_ERR_ARG, "costs must be nonnegative and finite");
            if (costs[p] != 1.0) use_costs = true;
        }
    bool all_full = true;
    for (int32_t u = 0; universe_p && u < NG; u++) {
        if (!(universe_p[u] >= 0.0 && universe_p[u] <= 1.0)) return cb_fail(ctx, CB_ERR_ARG, "universe_p must be in [0,1]");
        if (universe_p[u] != 1.0) all_full = false;
    }
    const char *mode_env = getenv("CB_GREEDY");
    const bool force_general = (mode_env && !strcmp(mode_env, "legacy")) ||
                               (getenv("CB_SETCOVER_FULL") && getenv("CB_SETCOVER_FULL")[0] == '1');
    if (all_full && !use_costs && !force_general)
        return cb_setcover_rounds_impl(ctx, cover, 0, P, ranks, false, sel_ids, n_sel, stats);

    EventTimer t_all(st), t_uni(st), t_greedy(st);
    t_all.start();
    t_uni.start();
    const int wide = ctx->sm_count * 8;
    const int64_t u_words = cover->universe_bits >> 6;

    DevBuf<unsigned long long> d_U, d_key, d_remaining, d_barrier;
    DevBuf<long long> d_usize, d_uncov, d_sel, d_nsel;
    DevBuf<uint32_t> d_gain, d_rank, d_ivg, d_bcount, d_bcursor;
    DevBuf<unsigned int> d_nleft;
    DevBuf<int> d_status;
    DevBuf<int64_t> d_boff;
    DevBuf<uint4> d_items;
    CB_CUDA(ctx, d_U.alloc((size_t)u_words + 1));
    CB_CUDA(ctx, cudaMemsetAsync(d_U.p, 0, sizeof(unsigned long long) * ((size_t)u_words + 1), st));
    CB_CUDA(ctx, d_usize.alloc((size_t)NG));
    universe_build_kernel<<<wide, 256, 0, st>>>(cover->d_iv, E, d_U.p);
    universe_size_kernel<<<wide, 256, 0, st>>>(d_U.p, cover->d_ubase, NG, d_usize.p);
    ctx->launches += 2;
    CB_CUDA(ctx, cudaGetLastError());
    std::vector<long long> h_usize((size_t)NG), h_uncov((size_t)NG);
    CB_CUDA(ctx, cudaMemcpyAsync(h_usize.data(), d_usize.p, sizeof(long long) * (size_t)NG, cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));

    // utils/set_cover.py:362-373: int(len(U) - p*len(U)) evaluated in IEEE doubles on the host,
    // product and difference rounded separately (no fused multiply-add)
    bool full_mode = false;
    unsigned long long remaining = 0;
    for (int32_t u = 0; u < NG; u++) {
        const double p = universe_p ? universe_p[u] : 1.0;
        volatile double len = (double)h_usize[(size_t)u];
        volatile double prod = p * len;
        volatile double diff = len - prod;
        h_uncov[(size_t)u] = (long long)diff;
        if (h_uncov[(size_t)u] != 0) full_mode = true;
        remaining += (unsigned long long)h_usize[(size_t)u];
    }
    if (const char *force = getenv("CB_SETCOVER_FULL")) if (force[0] == '1') full_mode = true;
    // every universe has to be covered completely after all (p_u so close to 1 that nothing may stay
    // uncovered): same problem as p_u == 1
    if (!full_mode && !use_costs && !force_general)
        return cb_setcover_rounds_impl(ctx, cover, 0, P, ranks, false, sel_ids, n_sel, stats);

    // ranks -> dense indices in ascending order of rank value (:349)
    std::vector<uint32_t> h_rank((size_t)P, 0u);
    int32_t n_ranks = 1;
    if (ranks) {
        std::vector<int32_t> vals(ranks, ranks + P);
        std::sort(vals.begin(), vals.end());
        vals.erase(std::unique(vals.begin(), vals.end()), vals.end());
        n_ranks = (int32_t)vals.size();
        for (int64_t p = 0; p < P; p++)
            h_rank[(size_t)p] = (uint32_t)(std::lower_bound(vals.begin(), vals.end(), ranks[p]) - vals.begin());
    }
    CB_CUDA(ctx, d_rank.alloc((size_t)P));
    CB_CUDA(ctx, cudaMemcpyAsync(d_rank.p, h_rank.data(), sizeof(uint32_t) * (size_t)P, cudaMemcpyHostToDevice, st));
    CB_CUDA(ctx, d_gain.alloc((size_t)P));
    CB_CUDA(ctx, d_key.alloc(2));
    CB_CUDA(ctx, cudaMemsetAsync(d_key.p, 0, sizeof(unsigned long long) * 2, st));
    CB_CUDA(ctx, d_remaining.alloc(1));
    CB_CUDA(ctx, cudaMemcpyAsync(d_remaining.p, &remaining, sizeof remaining, cudaMemcpyHostToDevice, st));
    CB_CUDA(ctx, d_uncov.alloc((size_t)NG));
    CB_CUDA(ctx, cudaMemcpyAsync(d_uncov.p, h_uncov.data(), sizeof(long long) * (size_t)NG, cudaMemcpyHostToDevice, st));
    CB_CUDA(ctx, d_sel.alloc((size_t)P));
    CB_CUDA(ctx, d_nsel.alloc(1));
    CB_CUDA(ctx, cudaMemsetAsync(d_nsel.p, 0, sizeof(long long), st));
    CB_CUDA(ctx, d_nleft.alloc(2));
    CB_CUDA(ctx, cudaMemsetAsync(d_nleft.p, 0, sizeof(unsigned int) * 2, st));
    CB_CUDA(ctx, d_status.alloc(1));
    CB_CUDA(ctx, cudaMemsetAsync(d_status.p, 0, sizeof(int), st));
    CB_CUDA(ctx, d_barrier.alloc(8));
    CB_CUDA(ctx, cudaMemsetAsync(d_barrier.p, 0, sizeof(unsigned long long) * 8, st));

    GreedyParams G;
    memset(&G, 0, sizeof G);
    G.n_probes = P;
    G.n_genomes = NG;
    G.n_intervals = E;
    G.iv_off = cover->d_iv_off;
    G.iv = cover->d_iv;
    G.ubase = cover->d_ubase;
    G.U = d_U.p;
    G.u_words = u_words;
    G.gain = d_gain.p;
    G.rank_idx = d_rank.p;
    G.n_ranks = n_ranks;
    G.max_len = cover->max_interval_len;
    G.remaining = d_remaining.p;
    G.u_size = d_usize.p;
    G.uncoverable = d_uncov.p;
    G.n_left = d_nleft.p;
    G.full_mode = full_mode ? 1 : 0;
    G.key = d_key.p;
    G.sel = d_sel.p;
    G.n_sel = d_nsel.p;
    G.status = d_status.p;
    G.barrier = d_barrier.p;
    G.phase_ns = d_barrier.p + 1;
    DevBuf<double> d_costs;
    DevBuf<uint32_t> d_idmin;
    if (use_costs) {
        // ratio keys: the complement of the order-preserving integer image of cost/gain in IEEE double,
        // then the smallest id at the minimum ratio in a second pass
        const uint32_t init[2] = {0xffffffffu, 0xffffffffu};
        CB_CUDA(ctx, d_costs.alloc((size_t)P));
        CB_CUDA(ctx, d_idmin.alloc(2));
        CB_CUDA(ctx, cudaMemcpyAsync(d_costs.p, costs, sizeof(double) * (size_t)P, cudaMemcpyHostToDevice, st));
        CB_CUDA(ctx, cudaMemcpyAsync(d_idmin.p, init, sizeof init, cudaMemcpyHostToDevice, st));
        G.costs = d_costs.p;
        G.idmin = d_idmin.p;
    }
    if (full_mode) {
        CB_CUDA(ctx, d_ivg.alloc((size_t)E));
        interval_genome_kernel<<<wide, 256, 0, st>>>(cover->d_iv, E, cover->d_ubase, NG, d_ivg.p);
        ctx->launches++;
        G.iv_genome = d_ivg.p;
    } else {
        const int64_t n_blocks = u_words + 1;
        CB_CUDA(ctx, d_bcount.alloc((size_t)n_blocks));
        CB_CUDA(ctx, d_bcursor.alloc((size_t)n_blocks));
        CB_CUDA(ctx, d_boff.alloc((size_t)n_blocks + 1));
        CB_CUDA(ctx, d_items.alloc((size_t)E));
        CB_CUDA(ctx, cudaMemsetAsync(d_bcount.p, 0, sizeof(uint32_t) * (size_t)n_blocks, st));
        CB_CUDA(ctx, cudaMemsetAsync(d_bcursor.p, 0, sizeof(uint32_t) * (size_t)n_blocks, st));
        block_index_kernel<false><<<wide, 256, 0, st>>>(cover->d_iv_off, cover->d_iv, P, d_bcount.p, nullptr, nullptr, nullptr);
        ctx->launches++;
        CB_TRY(cb_exclusive_scan_u32_to_i64(ctx, d_bcount.p, d_boff.p, n_blocks, nullptr));
        block_index_kernel<true><<<wide, 256, 0, st>>>(cover->d_iv_off, cover->d_iv, P, nullptr, d_boff.p, d_bcursor.p, d_items.p);
        gains_init_kernel<<<wide, 256, 0, st>>>(cover->d_iv_off, cover->d_iv, P, d_gain.p);
        ctx->launches += 2;
        if (E >= 0xfffffff0ll) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "more than 2^32 intervals in one grouping");
        G.blk_off = d_boff.p;
        G.blk_items = d_items.p;
        G.n_blocks = n_blocks;
    }
    CB_CUDA(ctx, cudaGetLastError());
    t_uni.stop();

    // ---- persistent cooperative launch: as many co-resident blocks as the device allows (at most 2 per SM)
    int per_sm = 0;
    CB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, greedy_kernel, GREEDY_THREADS, 0));
    if (per_sm < 1) return cb_fail(ctx, CB_ERR_CUDA, "greedy kernel does not fit on an SM");
    int want = 2;
    if (const char *e = getenv("CB_GREEDY_BLOCKS_PER_SM")) want = atoi(e) > 0 ? atoi(e) : want;
    if (per_sm > want) per_sm = want;
    const int grid = per_sm * ctx->sm_count;
    void *args[] = {(void *)&G};
    t_greedy.start();
    CB_CUDA(ctx, cudaLaunchCooperativeKernel((void *)greedy_kernel, dim3(grid), dim3(GREEDY_THREADS), args, 0, st));
    ctx->launches++;
    t_greedy.stop();
    t_all.stop();

    long long h_nsel = 0;
    int h_status = 0;
    unsigned long long h_phase[4] = {0, 0, 0, 0};
    CB_CUDA(ctx, cudaMemcpyAsync(h_phase, d_barrier.p + 1, sizeof h_phase, cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaMemcpyAsync(&h_nsel, d_nsel.p, sizeof h_nsel, cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaMemcpyAsync(&h_status, d_status.p, sizeof h_status, cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    if (h_status != 0) return cb_fail(ctx, CB_ERR_STATE, "set cover ran out of ranks before reaching the requested coverage");
    if (h_nsel > 0) {
        static_assert(sizeof(long long) == sizeof(int64_t), "int64");
        CB_CUDA(ctx, cudaMemcpyAsync(sel_ids, d_sel.p, sizeof(int64_t) * (size_t)h_nsel, cudaMemcpyDeviceToHost, st));
        CB_CUDA(ctx, cudaStreamSynchronize(st));
    }
    *n_sel = h_nsel;
    if (stats) {
        stats->ms_universe = t_uni.ms();
        stats->ms_greedy = t_greedy.ms();
        stats->ms_total = t_all.ms();
        stats->n_picks = h_nsel;
        stats->n_intervals = E;
        stats->n_kernel_launches = ctx->launches;
        for (int i = 0; i < 4; i++) stats->reserved[i] = (int64_t)h_phase[i];   // ns: argmax, barrier, apply, barrier
    }
    return CB_OK;
}
