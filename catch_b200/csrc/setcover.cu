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
// The whole loop is ONE persistent cooperative kernel (grid-wide barriers between the phases of a
// pick), so a pick costs two grid barriers instead of kernel launches plus a host round trip.
// Two modes:
//   incremental (every p_u == 1, so the min() clamp is inactive): gain[p] is kept exact by
//     subtracting, for every interval that overlaps the winner, the number of still-uncovered
//     bits in the overlap; overlapping intervals are found through an index of intervals bucketed
//     by start position (64-position blocks).
//   full (some p_u < 1): gains are recomputed from U for every probe at every pick.
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
    const uint2 *ivx;              // per interval: [x0, x1) range of blk_items that can overlap it
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
    // candidate list of the incremental kernel
    uint32_t *list;                // probes whose gain was >= tau when the list was built
    uint32_t *list_n;
    uint32_t list_cap;
    unsigned long long *pub;       // [0] sequence number, [1] message (kind << 32 | probe)
    unsigned long long *ctr;       // [0] list rebuilds, [1] picks served from a list / rounds, [2] active candidates summed over rounds
    // parallel-rounds kernel
    unsigned long long *mark;      // [u_words+1] highest key among the active candidates touching the word
    const double *costs;           // [n_probes] or nullptr (all 1): two-barrier kernel only
    uint32_t *idmin;               // [2] smallest id among the sets at the minimum ratio (cost mode)
    uint32_t *flag;                // [list_cap] conflict flag of active candidate a in the current round
    uint32_t *hist;                // [2][64] level histograms of the list rebuilds, alternating
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

// per interval: the contiguous range of indexed items whose start lies in
// (start - max_len, end), i.e. every interval that can overlap it
__global__ void item_range_kernel(const uint2 *__restrict__ iv, int64_t n, const int64_t *__restrict__ blk_off,
                                  int64_t n_blocks, uint32_t max_len, uint2 *__restrict__ ivx)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint2 r = iv[i];
        const int64_t lo_pos = (int64_t)r.x - (int64_t)max_len + 1;
        const int64_t b_lo = (lo_pos > 0 ? lo_pos : 0) >> 6;
        int64_t b_hi = ((int64_t)r.y - 1) >> 6;
        if (b_hi >= n_blocks) b_hi = n_blocks - 1;
        ivx[i] = make_uint2((uint32_t)blk_off[b_lo], (uint32_t)blk_off[b_hi + 1]);
    }
}

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


constexpr int APPLY_WORDS = 64;          // winner intervals up to 64 words are staged in shared memory

// apply() for a winner interval longer than APPLY_WORDS*64 positions: same work against L2
__device__ __forceinline__ void apply_long_interval(const GreedyParams &G, uint2 r, uint2 xr)
{
    for (int64_t x = (int64_t)xr.x + threadIdx.x; x < (int64_t)xr.y; x += GREEDY_THREADS) {
        const uint4 item = __ldg(G.blk_items + x);
        const uint32_t os = max(item.x, r.x), oe = min(item.y, r.y);
        if (os < oe) {
            const uint32_t dlt = popcount_range_cg(G.U, os, oe);
            if (dlt) atomicSub(&G.gain[item.z], dlt);
        }
    }
    __syncthreads();
    const uint32_t w0 = r.x >> 6, w1 = (r.y - 1) >> 6;
    uint32_t c = 0;
    for (uint32_t wd = w0 + threadIdx.x; wd <= w1; wd += GREEDY_THREADS) {
        unsigned long long m = ~0ull;
        if (wd == w0) m &= ~0ull << (r.x & 63);
        if (wd == w1) m &= ~0ull >> (63 - ((r.y - 1) & 63));
        const unsigned long long old = atomicAnd(&G.U[wd], ~m);
        c += __popcll(old & m);
    }
    if (c) atomicAdd(G.remaining, (unsigned long long)(-(long long)c));
    __syncthreads();
}

// apply() for ONE winner interval, by one CTA: stage the interval's still-uncovered bits in shared
// memory, subtract the uncovered bits of every overlap from the gain of the overlapping interval's
// probe, then clear exactly the staged bits.  CTA-uniform (contains __syncthreads).
__device__ __forceinline__ void apply_interval(const GreedyParams &G, int64_t i, unsigned long long *s_u)
{
    const uint2 r = G.iv[i];
    const uint2 xr = G.ivx[i];
    if (r.x >= r.y) return;
    const uint32_t w0 = r.x >> 6, w1 = (r.y - 1) >> 6;
    const uint32_t nwords = w1 - w0 + 1;
    if (nwords > (uint32_t)APPLY_WORDS) {         // very long interval: L2 path (CTA-uniform)
        apply_long_interval(G, r, xr);
        return;
    }
    // the winner interval's still-uncovered bits, staged in shared memory: they are what
    // every overlap is counted against AND exactly what has to be cleared afterwards
    unsigned long long mine = 0ull;
    if (threadIdx.x < nwords) {
        unsigned long long m = ~0ull;
        const uint32_t wd = w0 + threadIdx.x;
        if (wd == w0) m &= ~0ull << (r.x & 63);
        if (wd == w1) m &= ~0ull >> (63 - ((r.y - 1) & 63));
        mine = __ldcg(G.U + wd) & m;
        s_u[threadIdx.x] = mine;
    }
    // candidate items are fetched while the bits are in flight
    constexpr int BATCH = 4;
    const int64_t x0 = xr.x, x1 = xr.y;
    uint4 item[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; u++) {
        const int64_t x = x0 + threadIdx.x + (int64_t)u * GREEDY_THREADS;
        item[u] = x < x1 ? __ldg(G.blk_items + x) : make_uint4(0u, 0u, 0u, 0u);
    }
    const int any = __syncthreads_or(mine != 0ull);
    if (any) {
        int64_t xb = x0 + threadIdx.x;
        for (;;) {
#pragma unroll
            for (int u = 0; u < BATCH; u++) {
                const uint32_t os = max(item[u].x, r.x), oe = min(item[u].y, r.y);
                if (os < oe) {
                    const uint32_t wa = (os >> 6) - w0, wb = ((oe - 1) >> 6) - w0;
                    uint32_t dlt = 0;
                    for (uint32_t q = wa; q <= wb; q++) {
                        unsigned long long m = ~0ull;
                        if (q == wa) m &= ~0ull << (os & 63);
                        if (q == wb) m &= ~0ull >> (63 - ((oe - 1) & 63));
                        dlt += __popcll(s_u[q] & m);
                    }
                    if (dlt) atomicSub(&G.gain[item[u].z], dlt);
                }
            }
            xb += (int64_t)GREEDY_THREADS * BATCH;
            if (xb - threadIdx.x >= x1) break;          // CTA-uniform
#pragma unroll
            for (int u = 0; u < BATCH; u++) {
                const int64_t x = xb + (int64_t)u * GREEDY_THREADS;
                item[u] = x < x1 ? __ldg(G.blk_items + x) : make_uint4(0u, 0u, 0u, 0u);
            }
        }
        // clear exactly the bits that were set (nobody else touches this range)
        if (mine) {
            atomicAnd(&G.U[w0 + threadIdx.x], ~mine);
            atomicAdd(G.remaining, (unsigned long long)(-(long long)__popcll(mine)));
        }
    }
    __syncthreads();                             // s_u is reused by the next interval
}

// ---------------------------------------------------------------------------------------
// Incremental greedy, one grid-wide rendezvous per pick.
//
//   apply(w):  a CTA takes a winner interval: (1) narrow it to the span of its still-uncovered
//              bits, (2) for every indexed interval overlapping that span subtract the uncovered
//              bits in the overlap from the owner's gain, (3) __syncthreads, (4) clear the bits.
//              The bits a CTA reads in (2) lie inside ITS winner interval and winner intervals
//              are disjoint, so (2) and (4) of different CTAs never interfere: no grid barrier
//              between "update gains" and "clear U".
//   pick:      gains only ever decrease.  When a candidate list is built, it holds every probe of
//              the current rank with gain >= tau; all other probes stay below tau for ever.  As
//              long as the best CURRENT gain inside the list is >= tau it is the global maximum
//              (ties broken by the id embedded in the key), so CTA 0 alone finds the next winner
//              from the list after all CTAs have arrived, and publishes it; the others spin on the
//              published sequence number.  When the list's best falls below tau the list is
//              rebuilt with a full argmax + collect pass.
// ---------------------------------------------------------------------------------------
constexpr unsigned long long MSG_WINNER = 1, MSG_REBUILD = 2, MSG_DONE = 3;

__device__ __forceinline__ void arrive_only(unsigned long long *counter, unsigned long long &target)
{
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1ull);
    }
}

__global__ void __launch_bounds__(GREEDY_THREADS)
greedy_inc_kernel(const GreedyParams G)
{
    __shared__ unsigned long long s_key[GREEDY_THREADS / 32];
    __shared__ unsigned long long s_msg;
    __shared__ unsigned long long s_u[APPLY_WORDS];
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gsize = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    int cur_rank = 0;
    long long n_picks = 0;
    unsigned long long bar_target = 0, seq = 0, n_rebuilds = 0, n_listed = 0;
    uint32_t tau = 1;
    int band_shift = 4;                 // list threshold = gmax - gmax >> band_shift
    unsigned rb = 0;                    // rebuild counter: selects the argmax slot
    bool need_rebuild = true;
    long long w = -1;
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
    auto block_max = [&](unsigned long long best) -> unsigned long long {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
            best = t > best ? t : best;
        }
        if (lane == 0) s_key[warp] = best;
        __syncthreads();
        best = lane < GREEDY_THREADS / 32 ? s_key[lane] : 0ull;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
            best = t > best ? t : best;
        }
        __syncthreads();
        return best;                    // valid in every thread of warp 0 (and all warps: same reduction)
    };

    for (unsigned it = 0;; it++) {
        if (need_rebuild) {
            // ---- full argmax over the current rank
            unsigned long long best = 0;
            for (int64_t p = gtid; p < G.n_probes; p += gsize) {
                const uint32_t g = __ldcg(&G.gain[p]);
                if (g && G.rank_idx[p] == (uint32_t)cur_rank) {
                    const unsigned long long key = ((unsigned long long)g << 32) | (unsigned long long)(0xffffffffu - (uint32_t)p);
                    best = key > best ? key : best;
                }
            }
            best = block_max(best);
            const unsigned slot_k = rb & 1u;
            rb++;
            if (threadIdx.x == 0 && best) atomicMax(&G.key[slot_k], best);
            if (gtid == 0) *G.list_n = 0;
            grid_barrier(G.barrier, bar_target);
            if (__ldcg(G.remaining) == 0ull) break;
            const unsigned long long key = __ldcg(&G.key[slot_k]);
            if (gtid == 0) G.key[slot_k ^ 1u] = 0ull;
            if (key == 0ull) {                  // rank exhausted (:522-526)
                cur_rank++;
                if (cur_rank >= G.n_ranks) {
                    if (gtid == 0) *G.status = CB_ERR_STATE;
                    break;
                }
                grid_barrier(G.barrier, bar_target);
                continue;
            }
            w = (long long)(0xffffffffu - (uint32_t)(key & 0xffffffffull));
            const uint32_t gmax = (uint32_t)(key >> 32);
            // ---- collect the candidate list: everything within a band below the maximum
            for (;;) {
                const uint32_t band = gmax >> band_shift;
                tau = gmax - band;
                if (tau < 1u) tau = 1u;
                for (int64_t p = gtid; p < G.n_probes; p += gsize) {
                    const uint32_t g = __ldcg(&G.gain[p]);
                    if (g >= tau && G.rank_idx[p] == (uint32_t)cur_rank) {
                        const uint32_t slot = atomicAdd(G.list_n, 1u);
                        if (slot < G.list_cap) G.list[slot] = (uint32_t)p;
                    }
                }
                grid_barrier(G.barrier, bar_target);
                const uint32_t n = __ldcg(G.list_n);
                if (n <= G.list_cap) {
                    // aim for a few hundred candidates: widen the band when the list is short
                    if (n < G.list_cap / 16 && band_shift > 1) band_shift--;
                    break;
                }
                // too many candidates: everyone has read n; empty the list and narrow the band
                grid_barrier(G.barrier, bar_target);
                if (gtid == 0) *G.list_n = 0;
                grid_barrier(G.barrier, bar_target);
                if (band == 0u) {
                    // more than list_cap probes tie at the maximum: run without a list
                    // (the empty list forces a full argmax for every pick)
                    tau = 0xffffffffu;
                    break;
                }
                band_shift++;
            }
            need_rebuild = false;
            n_rebuilds++;
            lap(0);
        }

        // ---- apply the pick w
        if (gtid == 0) G.sel[n_picks] = w;
        n_picks++;
        {
            const int64_t i0 = G.iv_off[w], i1 = G.iv_off[w + 1];
            for (int64_t i = i0 + blockIdx.x; i < i1; i += gridDim.x) apply_interval(G, i, s_u);
        }
        lap(2);

        // ---- rendezvous: everyone arrives, CTA 0 picks from the list and publishes
        arrive_only(G.barrier, bar_target);
        seq++;
        if (blockIdx.x == 0) {
            if (threadIdx.x == 0) {
                while (*(volatile unsigned long long *)G.barrier < bar_target) { }
                __threadfence();
            }
            __syncthreads();
            const uint32_t n = __ldcg(G.list_n);
            unsigned long long best = 0;
            for (uint32_t i = threadIdx.x; i < n; i += GREEDY_THREADS) {
                const uint32_t p = G.list[i];
                const uint32_t g = __ldcg(&G.gain[p]);
                if (g) {
                    const unsigned long long key = ((unsigned long long)g << 32) | (unsigned long long)(0xffffffffu - p);
                    best = key > best ? key : best;
                }
            }
            best = block_max(best);
            if (threadIdx.x == 0) {
                unsigned long long msg;
                if (__ldcg(G.remaining) == 0ull) msg = MSG_DONE << 32;
                else if ((uint32_t)(best >> 32) >= tau) msg = (MSG_WINNER << 32) | (unsigned long long)(0xffffffffu - (uint32_t)(best & 0xffffffffull));
                else msg = MSG_REBUILD << 32;
                G.pub[1] = msg;
                __threadfence();
                *(volatile unsigned long long *)G.pub = seq;
                s_msg = msg;
            }
            __syncthreads();
        } else {
            if (threadIdx.x == 0) {
                while (*(volatile unsigned long long *)G.pub < seq) { }
                __threadfence();
                s_msg = *(volatile unsigned long long *)(G.pub + 1);
            }
            __syncthreads();
        }
        const unsigned long long msg = s_msg;
        __syncthreads();
        lap(3);
        const unsigned long long kind = msg >> 32;
        if (kind == MSG_DONE) break;
        if (kind == MSG_REBUILD) { need_rebuild = true; continue; }
        w = (long long)(uint32_t)(msg & 0xffffffffull);
        n_listed++;
    }
    if (gtid == 0) {
        *G.n_sel = n_picks;
        for (int i = 0; i < 4; i++) G.phase_ns[i] = t_phase[i];
        G.ctr[0] = n_rebuilds;
        G.ctr[1] = n_listed;
    }
}

// apply() for ONE winner interval by ONE WARP (parallel-rounds kernel: thousands of winner intervals
// per round, each a short dependent chain of loads, so many of them must be in flight per SM).
// Same work as apply_interval; `s_uw` is the warp's private APPLY_WORDS-word staging area.
__device__ __forceinline__ void apply_interval_warp(const GreedyParams &G, int64_t i, unsigned long long *s_uw, int lane)
{
    const uint2 r = G.iv[i];
    const uint2 xr = G.ivx[i];
    if (r.x >= r.y) return;
    const uint32_t w0 = r.x >> 6, w1 = (r.y - 1) >> 6;
    const uint32_t nwords = w1 - w0 + 1;
    const int64_t x0 = xr.x, x1 = xr.y;
    if (nwords > (uint32_t)APPLY_WORDS) {          // very long interval: count against L2, no staging
        for (int64_t x = x0 + lane; x < x1; x += 32) {
            const uint4 item = __ldg(G.blk_items + x);
            const uint32_t os = max(item.x, r.x), oe = min(item.y, r.y);
            if (os < oe) {
                const uint32_t dlt = popcount_range_cg(G.U, os, oe);
                if (dlt) atomicSub(&G.gain[item.z], dlt);
            }
        }
        __syncwarp();
        uint32_t c = 0;
        for (uint32_t wd = w0 + lane; wd <= w1; wd += 32) {
            unsigned long long m = ~0ull;
            if (wd == w0) m &= ~0ull << (r.x & 63);
            if (wd == w1) m &= ~0ull >> (63 - ((r.y - 1) & 63));
            const unsigned long long old = atomicAnd(&G.U[wd], ~m);
            c += __popcll(old & m);
        }
        if (c) atomicAdd(G.remaining, (unsigned long long)(-(long long)c));
        __syncwarp();
        return;
    }
    unsigned long long mine[APPLY_WORDS / 32];
#pragma unroll
    for (int h = 0; h < APPLY_WORDS / 32; h++) {
        const uint32_t q = (uint32_t)lane + 32u * h;
        mine[h] = 0ull;
        if (q < nwords) {
            unsigned long long m = ~0ull;
            const uint32_t wd = w0 + q;
            if (wd == w0) m &= ~0ull << (r.x & 63);
            if (wd == w1) m &= ~0ull >> (63 - ((r.y - 1) & 63));
            mine[h] = __ldcg(G.U + wd) & m;
            s_uw[q] = mine[h];
        }
    }
    constexpr int BATCH = 8;
    uint4 item[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; u++) {
        const int64_t x = x0 + lane + (int64_t)u * 32;
        item[u] = x < x1 ? __ldg(G.blk_items + x) : make_uint4(0u, 0u, 0u, 0u);
    }
    bool have = false;
#pragma unroll
    for (int h = 0; h < APPLY_WORDS / 32; h++) have |= mine[h] != 0ull;
    __syncwarp();                                   // the staged words are read by the other lanes below
    if (!__any_sync(0xffffffffu, have)) return;    // everything here is covered already
    int64_t xb = x0 + lane;
    for (;;) {
#pragma unroll
        for (int u = 0; u < BATCH; u++) {
            const uint32_t os = max(item[u].x, r.x), oe = min(item[u].y, r.y);
            if (os < oe) {
                const uint32_t wa = (os >> 6) - w0, wb = ((oe - 1) >> 6) - w0;
                uint32_t dlt = 0;
                for (uint32_t q = wa; q <= wb; q++) {
                    unsigned long long m = ~0ull;
                    if (q == wa) m &= ~0ull << (os & 63);
                    if (q == wb) m &= ~0ull >> (63 - ((oe - 1) & 63));
                    dlt += __popcll(s_uw[q] & m);
                }
                if (dlt) atomicSub(&G.gain[item[u].z], dlt);
            }
        }
        xb += 32 * BATCH;
        if (xb - lane >= x1) break;                 // warp-uniform
#pragma unroll
        for (int u = 0; u < BATCH; u++) {
            const int64_t x = xb + (int64_t)u * 32;
            item[u] = x < x1 ? __ldg(G.blk_items + x) : make_uint4(0u, 0u, 0u, 0u);
        }
    }
    // clear exactly the bits that were set (nobody else touches them)
    uint32_t c = 0;
#pragma unroll
    for (int h = 0; h < APPLY_WORDS / 32; h++)
        if (mine[h]) {
            atomicAnd(&G.U[w0 + lane + 32u * h], ~mine[h]);
            c += __popcll(mine[h]);
        }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0 && c) atomicAdd(G.remaining, (unsigned long long)(-(long long)c));
    __syncwarp();                                   // s_uw is reused by the warp's next interval
}

// ---------------------------------------------------------------------------------------
// Incremental greedy in PARALLEL ROUNDS (every p_u == 1).
//
// Sequential greedy picks the probe with the largest key = (gain, smallest id) again and again.
// Call two probes in conflict when they share a still-uncovered universe bit.  A probe whose key
// is larger than the key of every probe it conflicts with ("local maximum") keeps its gain until
// it is picked -- a conflicting neighbour would have to become the global maximum first, and it
// cannot while the probe is there -- and it IS picked in the end, because nobody else can cover
// its bits before it.  So all local maxima can be applied at once: the selected SET and every
// probe's gain at pick time are those of the sequential loop.  Keys at pick time are strictly
// decreasing along the sequential pick sequence, therefore sorting the picks by that key (done
// by the host part of cb_setcover) restores the sequential pick ORDER (utils/set_cover.py:
// 393-433, 483-526 -- the order matters because the reference builds its result set by .add()
// in pick order).
//
// Local maxima are searched among the candidate list only: it holds every probe (of the current
// rank) with gain >= tau, i.e. a prefix of the global key order, so every probe with a larger
// key than a list member is itself in the list.  Per round, three phases separated by grid
// barriers:
//   mark:   every active candidate writes atomicMax(mark[w], key) for each universe word w in
//           which one of its intervals still has uncovered bits (one thread per interval);
//   check:  a candidate is accepted iff mark[w] == its key in all of those words (conflicts are
//           detected at word granularity: false conflicts only postpone a pick);
//   apply:  marks are reset; the intervals of ALL accepted probes are spread over the CTAs and
//           applied as in the one-pick kernel (accepted probes share no word with uncovered bits,
//           so their staged bit ranges are disjoint).
// The active candidates and the winners are compacted redundantly by every CTA into its own
// shared memory (same inputs, same deterministic scan), which saves a barrier and all counters.
// ---------------------------------------------------------------------------------------
constexpr int PAR_GAIN_LEVELS = 14;
constexpr int PAR_ID_LEVELS = 33;
constexpr int PAR_LIST_CAP = 4096;          // upper limit of the candidate list (a runtime cap <= this is used)

template <typename F>
__device__ __forceinline__ void for_each_word(uint2 r, F f)
{
    if (r.x >= r.y) return;
    const uint32_t w0 = r.x >> 6, w1 = (r.y - 1) >> 6;
    for (uint32_t w = w0; w <= w1; w++) {
        unsigned long long m = ~0ull;
        if (w == w0) m &= ~0ull << (r.x & 63);
        if (w == w1) m &= ~0ull >> (63 - ((r.y - 1) & 63));
        f(w, m);
    }
}

__global__ void __launch_bounds__(GREEDY_THREADS, 3)
greedy_par_kernel(const GreedyParams G)
{
    __shared__ unsigned long long s_key[GREEDY_THREADS / 32];
    __shared__ unsigned long long s_u[(GREEDY_THREADS / 32) * APPLY_WORDS];   // one staging area per warp
    __shared__ uint32_t s_part[GREEDY_THREADS / 32];
    __shared__ uint32_t s_hist[64];                  // level histogram of a list rebuild
    extern __shared__ uint32_t s_dyn[];
    // per CTA, list_cap entries each: active candidates (probe, gain, first interval, exclusive prefix
    // of the interval counts) and the winners among them (first interval, prefix)
    uint32_t *s_p = s_dyn, *s_g = s_p + G.list_cap, *s_i0 = s_g + G.list_cap, *s_base = s_i0 + G.list_cap,
             *s_wi0 = s_base + G.list_cap + 1, *s_wbase = s_wi0 + G.list_cap;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gsize = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NWARP = GREEDY_THREADS / 32;
    constexpr int PER = PAR_LIST_CAP / GREEDY_THREADS;
    // exclusive prefix of v over the CTA (thread order); adds the CTA total to `total`
    auto block_scan = [&](uint32_t v, uint32_t &total) -> uint32_t {
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        __syncthreads();                 // s_part may still be read from the previous scan
        if (lane == 31) s_part[warp] = inc;
        __syncthreads();
        uint32_t before = 0, all = 0;
#pragma unroll
        for (int q = 0; q < NWARP; q++) {
            const uint32_t t = s_part[q];
            if (q < warp) before += t;
            all += t;
        }
        total += all;
        return before + inc - v;
    };

    int cur_rank = 0;
    long long n_picks = 0;
    unsigned long long bar_target = 0, n_rebuilds = 0, n_rounds = 0, n_active_sum = 0;
    uint32_t tau = 1;
    unsigned rb = 0;
    bool need_rebuild = true;
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
    auto block_max = [&](unsigned long long best) -> unsigned long long {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
            best = t > best ? t : best;
        }
        if (lane == 0) s_key[warp] = best;
        __syncthreads();
        best = lane < NWARP ? s_key[lane] : 0ull;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
            best = t > best ? t : best;
        }
        __syncthreads();
        return best;
    };

    for (;;) {
        if (need_rebuild) {
            // ---- full argmax over the current rank, then the candidate list (band below the maximum)
            unsigned long long best = 0;
            for (int64_t p = gtid; p < G.n_probes; p += gsize) {
                const uint32_t g = __ldcg(&G.gain[p]);
                if (g && G.rank_idx[p] == (uint32_t)cur_rank) {
                    const unsigned long long key = ((unsigned long long)g << 32) | (unsigned long long)(0xffffffffu - (uint32_t)p);
                    best = key > best ? key : best;
                }
            }
            best = block_max(best);
            const unsigned slot_k = rb & 1u;
            rb++;
            if (threadIdx.x == 0 && best) atomicMax(&G.key[slot_k], best);
            if (gtid == 0) *G.list_n = 0;
            grid_barrier(G.barrier, bar_target);
            if (__ldcg(G.remaining) == 0ull) break;
            const unsigned long long key = __ldcg(&G.key[slot_k]);
            if (gtid == 0) G.key[slot_k ^ 1u] = 0ull;
            if (key == 0ull) {                  // rank exhausted (:522-526)
                cur_rank++;
                if (cur_rank >= G.n_ranks) {
                    if (gtid == 0) *G.status = CB_ERR_STATE;
                    break;
                }
                grid_barrier(G.barrier, bar_target);
                continue;
            }
            const uint32_t wmax = 0xffffffffu - (uint32_t)(key & 0xffffffffull);
            const uint32_t gmax = (uint32_t)(key >> 32);
            // ---- choose the list threshold from a histogram, so that the list holds as many of the
            // top candidates as fit: gain levels tau_0 = 1, tau_s = gmax - (gmax >> s) (s = 1..12),
            // tau_13 = gmax; and, should more probes than fit TIE at gmax, id levels
            // "id < ceil(P / 2^j)" among those ties.  Either way the list is a prefix of the key order.
            auto tau_of = [&](int lv) -> uint32_t {
                if (lv == 0) return 1u;
                if (lv >= PAR_GAIN_LEVELS - 1) return gmax;
                const uint32_t t = gmax - (gmax >> lv);
                return t < 1u ? 1u : t;
            };
            // id levels halve the span from the smallest tied id (wmax) to the end: sequential greedy
            // consumes ties from the low ids upwards, so the remaining ones sit in [wmax, P)
            auto idthr_of = [&](int j) -> unsigned long long {
                const unsigned long long span = (unsigned long long)G.n_probes - (unsigned long long)wmax;
                return (unsigned long long)wmax + ((span + (1ull << j) - 1ull) >> j);
            };
            if (threadIdx.x < 64) s_hist[threadIdx.x] = 0u;
            __syncthreads();
            for (int64_t p = gtid; p < G.n_probes; p += gsize) {
                const uint32_t g = __ldcg(&G.gain[p]);
                if (g && G.rank_idx[p] == (uint32_t)cur_rank) {
                    int lv = PAR_GAIN_LEVELS - 1;
                    while (lv > 0 && g < tau_of(lv)) lv--;          // largest level the probe qualifies for
                    atomicAdd(&s_hist[lv], 1u);
                    if (g == gmax) {
                        int j = 0;
                        while (j + 1 < PAR_ID_LEVELS && (unsigned long long)p < idthr_of(j + 1)) j++;
                        atomicAdd(&s_hist[PAR_GAIN_LEVELS + j], 1u);
                    }
                }
            }
            __syncthreads();
            uint32_t *hist_now = G.hist + 64 * slot_k, *hist_next = G.hist + 64 * (slot_k ^ 1u);
            if (threadIdx.x < 64 && s_hist[threadIdx.x]) atomicAdd(&hist_now[threadIdx.x], s_hist[threadIdx.x]);
            grid_barrier(G.barrier, bar_target);
            if (gtid < 64) hist_next[gtid] = 0u;                    // last read one rebuild ago
            uint32_t id_thr = 0xffffffffu;
            {
                // counts are suffix sums: a probe at level lv also qualifies for every wider level
                uint32_t c = 0;
                int pick = -1;
                for (int lv = PAR_GAIN_LEVELS - 1; lv >= 0; lv--) {
                    c += __ldcg(&hist_now[lv]);
                    if (c <= G.list_cap) pick = lv; else break;
                }
                if (pick >= 0) {
                    tau = tau_of(pick);
                } else {                                            // more ties at gmax than the list holds
                    tau = gmax;
                    c = 0;
                    int pj = -1;
                    for (int j = PAR_ID_LEVELS - 1; j >= 0; j--) {
                        c += __ldcg(&hist_now[PAR_GAIN_LEVELS + j]);
                        if (c <= G.list_cap) pj = j; else break;
                    }
                    // the smallest tied id is wmax; an id level that holds no tie at all (or none that
                    // fits) leaves the argmax alone in the list
                    unsigned long long thr = pj >= 0 ? idthr_of(pj) : 0ull;
                    if (thr <= (unsigned long long)wmax) thr = (unsigned long long)wmax + 1ull;
                    id_thr = thr > 0xffffffffull ? 0xffffffffu : (uint32_t)thr;
                }
            }
            for (int64_t p = gtid; p < G.n_probes; p += gsize) {
                const uint32_t g = __ldcg(&G.gain[p]);
                if (g >= tau && (uint32_t)p < id_thr && G.rank_idx[p] == (uint32_t)cur_rank) {
                    const uint32_t slot = atomicAdd(G.list_n, 1u);
                    if (slot < G.list_cap) G.list[slot] = (uint32_t)p;      // always true, by the counts
                }
            }
            grid_barrier(G.barrier, bar_target);
            need_rebuild = false;
            n_rebuilds++;
            lap(0);
        }

        // ---- active candidates: list entries whose gain is still >= tau.  Every CTA derives the
        // same compact arrays in its own shared memory (gains are stable until the next apply), so
        // no global list, counter or extra barrier is needed.
        const uint32_t n_list = min(__ldcg(G.list_n), G.list_cap);
        uint32_t n_act = 0;
        {
            uint32_t pv[PER], gv[PER];
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const uint32_t c = (uint32_t)u * GREEDY_THREADS + threadIdx.x;
                pv[u] = c < n_list ? __ldcg(&G.list[c]) : 0xffffffffu;
            }
#pragma unroll
            for (int u = 0; u < PER; u++) gv[u] = pv[u] != 0xffffffffu ? __ldcg(&G.gain[pv[u]]) : 0u;
#pragma unroll
            for (int u = 0; u < PER; u++) {
                if ((uint32_t)u * GREEDY_THREADS >= n_list) break;          // CTA-uniform
                const bool act = gv[u] >= tau && pv[u] != 0xffffffffu;
                const uint32_t before = n_act;
                const uint32_t pos = before + block_scan(act ? 1u : 0u, n_act);
                if (act) { s_p[pos] = pv[u]; s_g[pos] = gv[u]; }
            }
            __syncthreads();
        }
        if (n_act == 0u) {                       // the list is used up
            need_rebuild = true;
            continue;
        }
        uint32_t total_pairs = 0;
        {
            uint32_t i0v[PER], cntv[PER];
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const uint32_t a = (uint32_t)u * GREEDY_THREADS + threadIdx.x;
                i0v[u] = cntv[u] = 0;
                if (a < n_act) {
                    const uint32_t p = s_p[a];
                    const int64_t x = G.iv_off[p], y = G.iv_off[p + 1];
                    i0v[u] = (uint32_t)x;
                    cntv[u] = (uint32_t)(y - x);
                }
            }
#pragma unroll
            for (int u = 0; u < PER; u++) {
                if ((uint32_t)u * GREEDY_THREADS >= n_act) break;           // CTA-uniform
                const uint32_t a = (uint32_t)u * GREEDY_THREADS + threadIdx.x;
                const uint32_t before = total_pairs;
                const uint32_t pos = before + block_scan(cntv[u], total_pairs);
                if (a < n_act) { s_i0[a] = i0v[u]; s_base[a] = pos; }
            }
            if (threadIdx.x == 0) s_base[n_act] = total_pairs;
            // the conflict flags of the previous round have been read by everybody (barrier since)
            if (blockIdx.x == 0)
                for (uint32_t a = threadIdx.x; a < G.list_cap; a += GREEDY_THREADS) G.flag[a] = 0u;
            __syncthreads();
        }
        auto pair_of = [&](uint32_t f, uint32_t &a) -> int64_t {
            uint32_t lo = 0, hi = n_act;
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (s_base[mid] <= f) lo = mid; else hi = mid;
            }
            a = lo;
            return (int64_t)s_i0[lo] + (f - s_base[lo]);
        };
        auto key_of = [&](uint32_t a) -> unsigned long long {
            return ((unsigned long long)s_g[a] << 32) | (unsigned long long)(0xffffffffu - s_p[a]);
        };

        // ---- mark: one thread per (candidate, interval)
        for (int64_t f = gtid; f < (int64_t)total_pairs; f += gsize) {
            uint32_t a;
            const int64_t i = pair_of((uint32_t)f, a);
            const unsigned long long key = key_of(a);
            for_each_word(G.iv[i], [&](uint32_t w, unsigned long long m) {
                if (__ldcg(G.U + w) & m) atomicMax(&G.mark[w], key);
            });
        }
        grid_barrier(G.barrier, bar_target);

        // ---- check
        for (int64_t f = gtid; f < (int64_t)total_pairs; f += gsize) {
            uint32_t a;
            const int64_t i = pair_of((uint32_t)f, a);
            const unsigned long long key = key_of(a);
            bool conflict = false;
            for_each_word(G.iv[i], [&](uint32_t w, unsigned long long m) {
                if ((__ldcg(G.U + w) & m) && __ldcg(G.mark + w) != key) conflict = true;
            });
            if (conflict) G.flag[a] = 1u;
        }
        grid_barrier(G.barrier, bar_target);

        // ---- winners = active candidates without a conflict (same compaction in every CTA)
        uint32_t n_win = 0;
        {
            uint32_t fl[PER];
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const uint32_t a = (uint32_t)u * GREEDY_THREADS + threadIdx.x;
                fl[u] = a < n_act ? __ldcg(&G.flag[a]) : 1u;
            }
            uint32_t wsum = 0;
#pragma unroll
            for (int u = 0; u < PER; u++) {
                if ((uint32_t)u * GREEDY_THREADS >= n_act) break;           // CTA-uniform
                const uint32_t a = (uint32_t)u * GREEDY_THREADS + threadIdx.x;
                const bool win = fl[u] == 0u;
                const uint32_t cnt = win ? s_base[a + 1] - s_base[a] : 0u;
                const uint32_t before_n = n_win, before_s = wsum;
                const uint32_t j = before_n + block_scan(win ? 1u : 0u, n_win);
                const uint32_t base = before_s + block_scan(cnt, wsum);
                if (win) {
                    s_wi0[j] = s_i0[a];
                    s_wbase[j] = base;
                    if (blockIdx.x == 0) G.sel[n_picks + j] = (long long)key_of(a);
                }
            }
            if (threadIdx.x == 0) s_wbase[n_win] = wsum;
            __syncthreads();
        }
        lap(1);
        n_rounds++;
        n_active_sum += n_act;
        n_picks += n_win;

        // ---- reset the marks of this round
        for (int64_t f = gtid; f < (int64_t)total_pairs; f += gsize) {
            uint32_t a;
            const int64_t i = pair_of((uint32_t)f, a);
            for_each_word(G.iv[i], [&](uint32_t w, unsigned long long) { G.mark[w] = 0ull; });
        }
        // ---- apply every accepted probe: (winner, interval) pairs are dealt round-robin to the CTAs
        {
            const uint32_t total = s_wbase[n_win];
            for (uint32_t f = (uint32_t)warp * gridDim.x + blockIdx.x; f < total; f += gridDim.x * NWARP) {
                uint32_t lo = 0, hi = n_win;
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (s_wbase[mid] <= f) lo = mid; else hi = mid;
                }
                apply_interval_warp(G, (int64_t)s_wi0[lo] + (f - s_wbase[lo]), s_u + warp * APPLY_WORDS, lane);
            }
        }
        lap(2);
        grid_barrier(G.barrier, bar_target);
        lap(3);
        if (__ldcg(G.remaining) == 0ull) break;
    }
    if (gtid == 0) {
        *G.n_sel = n_picks;
        for (int i = 0; i < 4; i++) G.phase_ns[i] = t_phase[i];
        G.ctr[0] = n_rebuilds;
        G.ctr[1] = n_rounds;
        G.ctr[2] = n_active_sum;
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
    EventTimer t_all(st), t_uni(st), t_greedy(st);
    t_all.start();
    t_uni.start();
    const int wide = ctx->sm_count * 8;
    const int64_t u_words = cover->universe_bits >> 6;

    DevBuf<unsigned long long> d_U, d_key, d_remaining, d_barrier, d_pub, d_mark;
    DevBuf<uint32_t> d_list, d_winners;
    DevBuf<long long> d_usize, d_uncov, d_sel, d_nsel;
    DevBuf<uint32_t> d_gain, d_rank, d_ivg, d_bcount, d_bcursor;
    DevBuf<unsigned int> d_nleft;
    DevBuf<int> d_status;
    DevBuf<int64_t> d_boff;
    DevBuf<uint4> d_items;
    DevBuf<uint2> d_ivx;
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
        if (!(p >= 0.0 && p <= 1.0)) return cb_fail(ctx, CB_ERR_ARG, "universe_p must be in [0,1]");
        volatile double len = (double)h_usize[(size_t)u];
        volatile double prod = p * len;
        volatile double diff = len - prod;
        h_uncov[(size_t)u] = (long long)diff;
        if (h_uncov[(size_t)u] != 0) full_mode = true;
        remaining += (unsigned long long)h_usize[(size_t)u];
    }
    if (const char *force = getenv("CB_SETCOVER_FULL")) if (force[0] == '1') full_mode = true;

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
    // which kernel: parallel rounds (default), one pick per rendezvous ("inc"), or the two-barrier
    // kernel that recomputes clamped gains ("legacy"; the only one that handles p_u < 1)
    // non-unit costs (never produced by SetCoverFilter, set_cover_filter.py:759, but part of
    // approx_multiuniverse's contract): the two-barrier kernel with ratio keys
    bool use_costs = false;
    if (costs)
        for (int64_t p = 0; p < P; p++) {
            if (!(costs[p] >= 0.0) || costs[p] > 1e300) return cb_fail(ctx, CB_ERR_ARG, "costs must be nonnegative and finite");
            if (costs[p] != 1.0) use_costs = true;
        }
    const char *mode_env = getenv("CB_GREEDY");
    const bool legacy = full_mode || use_costs || (mode_env && !strcmp(mode_env, "legacy")) ||
                        (getenv("CB_GREEDY_LEGACY") && getenv("CB_GREEDY_LEGACY")[0] == '1');
    const bool par = !legacy && !(mode_env && !strcmp(mode_env, "inc"));
    uint32_t list_cap = par ? 2048u : 4096u;          // par: <= PAR_LIST_CAP; 2048 leaves room for 3 CTAs per SM
    if (const char *e = getenv("CB_GREEDY_LIST_CAP")) {
        const int v = atoi(e);
        if (v >= 1 && (uint32_t)v <= (par ? (uint32_t)PAR_LIST_CAP : 4096u)) list_cap = (uint32_t)v;
    }
    CB_CUDA(ctx, d_list.alloc(list_cap + 1));
    CB_CUDA(ctx, d_pub.alloc(8));
    CB_CUDA(ctx, cudaMemsetAsync(d_pub.p, 0, sizeof(unsigned long long) * 8, st));
    CB_CUDA(ctx, cudaMemsetAsync(d_list.p, 0, sizeof(uint32_t) * (list_cap + 1), st));
    G.list = d_list.p;
    G.list_n = d_list.p + list_cap;
    G.list_cap = list_cap;
    G.pub = d_pub.p;
    G.ctr = d_pub.p + 4;
    DevBuf<double> d_costs;
    DevBuf<uint32_t> d_idmin;
    if (use_costs) {
        const uint32_t init[2] = {0xffffffffu, 0xffffffffu};
        CB_CUDA(ctx, d_costs.alloc((size_t)P));
        CB_CUDA(ctx, d_idmin.alloc(2));
        CB_CUDA(ctx, cudaMemcpyAsync(d_costs.p, costs, sizeof(double) * (size_t)P, cudaMemcpyHostToDevice, st));
        CB_CUDA(ctx, cudaMemcpyAsync(d_idmin.p, init, sizeof init, cudaMemcpyHostToDevice, st));
        G.costs = d_costs.p;
        G.idmin = d_idmin.p;
    }
    if (par) {
        CB_CUDA(ctx, d_mark.alloc((size_t)u_words + 1));
        CB_CUDA(ctx, cudaMemsetAsync(d_mark.p, 0, sizeof(unsigned long long) * ((size_t)u_words + 1), st));
        CB_CUDA(ctx, d_winners.alloc(list_cap + 128));
        CB_CUDA(ctx, cudaMemsetAsync(d_winners.p, 0, sizeof(uint32_t) * (list_cap + 128), st));
        G.mark = d_mark.p;
        G.flag = d_winners.p;
        G.hist = d_winners.p + list_cap;
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
        CB_CUDA(ctx, d_ivx.alloc((size_t)E));
        item_range_kernel<<<wide, 256, 0, st>>>(cover->d_iv, E, d_boff.p, n_blocks, cover->max_interval_len, d_ivx.p);
        ctx->launches++;
        G.blk_off = d_boff.p;
        G.blk_items = d_items.p;
        G.n_blocks = n_blocks;
        G.ivx = d_ivx.p;
    }
    CB_CUDA(ctx, cudaGetLastError());
    t_uni.stop();

    // ---- persistent cooperative launch: as many co-resident blocks as the device allows
    void *kernel = legacy ? (void *)greedy_kernel : par ? (void *)greedy_par_kernel : (void *)greedy_inc_kernel;
    int per_sm = 0;
    size_t dyn_smem = 0;
    if (legacy) CB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, greedy_kernel, GREEDY_THREADS, 0));
    else if (par) {
        dyn_smem = sizeof(uint32_t) * (6 * (size_t)list_cap + 2);
        CB_CUDA(ctx, cudaFuncSetAttribute(greedy_par_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
        CB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, greedy_par_kernel, GREEDY_THREADS, dyn_smem));
    }
    else CB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, greedy_inc_kernel, GREEDY_THREADS, 0));
    if (per_sm < 1) return cb_fail(ctx, CB_ERR_CUDA, "greedy kernel does not fit on an SM");
    int want = par ? 3 : 2;
    if (const char *e = getenv("CB_GREEDY_BLOCKS_PER_SM")) want = atoi(e) > 0 ? atoi(e) : want;
    if (per_sm > want) per_sm = want;
    const int grid = per_sm * ctx->sm_count;
    void *args[] = {(void *)&G};
    t_greedy.start();
    CB_CUDA(ctx, cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(GREEDY_THREADS), args, dyn_smem, st));
    ctx->launches++;
    t_greedy.stop();
    t_all.stop();

    long long h_nsel = 0;
    int h_status = 0;
    unsigned long long h_phase[4] = {0, 0, 0, 0}, h_ctr[3] = {0, 0, 0};
    CB_CUDA(ctx, cudaMemcpyAsync(h_phase, d_barrier.p + 1, sizeof h_phase, cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaMemcpyAsync(h_ctr, d_pub.p + 4, sizeof h_ctr, cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaMemcpyAsync(&h_nsel, d_nsel.p, sizeof h_nsel, cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaMemcpyAsync(&h_status, d_status.p, sizeof h_status, cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    if (h_status != 0) return cb_fail(ctx, CB_ERR_STATE, "set cover ran out of ranks before reaching the requested coverage");
    if (h_nsel > 0) {
        static_assert(sizeof(long long) == sizeof(int64_t), "int64");
        CB_CUDA(ctx, cudaMemcpyAsync(sel_ids, d_sel.p, sizeof(int64_t) * (size_t)h_nsel, cudaMemcpyDeviceToHost, st));
        CB_CUDA(ctx, cudaStreamSynchronize(st));
        if (par) {
            // the kernel reports each pick's key at pick time, (gain << 32) | (2^32-1 - id), in no
            // particular order inside a round; the sequential loop picks rank by rank and, inside a
            // rank, in strictly decreasing key order (see greedy_par_kernel)
            uint64_t *keys = reinterpret_cast<uint64_t *>(sel_ids);
            auto id_of = [](uint64_t k) { return (int64_t)(0xffffffffu - (uint32_t)(k & 0xffffffffull)); };
            if (n_ranks > 1)
                std::sort(keys, keys + h_nsel, [&](uint64_t a, uint64_t b) {
                    const uint32_t ra = h_rank[(size_t)id_of(a)], rb = h_rank[(size_t)id_of(b)];
                    return ra != rb ? ra < rb : a > b;
                });
            else
                std::sort(keys, keys + h_nsel, [](uint64_t a, uint64_t b) { return a > b; });
            for (long long i = 0; i < h_nsel; i++) sel_ids[i] = id_of(keys[i]);
        }
    }
    *n_sel = h_nsel;
    if (stats) {
        stats->ms_universe = t_uni.ms();
        stats->ms_greedy = t_greedy.ms();
        stats->ms_total = t_all.ms();
        stats->n_picks = h_nsel;
        stats->n_intervals = E;
        stats->n_kernel_launches = ctx->launches;
        for (int i = 0; i < 4; i++) stats->reserved[i] = (int64_t)h_phase[i];   // ns: argmax/rebuild, barrier, apply, rendezvous
        stats->reserved[4] = (int64_t)h_ctr[0];      // candidate-list rebuilds
        stats->reserved[5] = (int64_t)h_ctr[1];      // picks served from a list (inc) / rounds (par)
        stats->reserved[6] = (int64_t)h_ctr[2];      // active candidates summed over the rounds (par)
    }
    return CB_OK;
}
