// Stage B, default path: exact greedy multi-universe set cover in PARALLEL ROUNDS, on one GPU or
// with the candidate probes sharded over the GPUs of one box.
//
// Replaces utils/set_cover.py:147-615 approx_multiuniverse(use_intervalsets=True) for the case every
// caller of SetCoverFilter produces by default: unit costs and p_u == 1 (coverage 1.0), with or
// without ranks.  (p_u < 1 and non-unit costs take the two-barrier kernel of setcover.cu.)
//
// Why rounds are exact.  Sequential greedy picks the probe with the largest key = (gain, smallest
// id) again and again (:393-433, :483-526).  Call two probes in conflict when they share a
// still-uncovered universe bit.  A probe whose key is larger than the key of every probe it
// conflicts with keeps its gain until it is picked -- a conflicting neighbour would have to become
// the global maximum first, and it cannot while the probe is there -- and it IS picked in the end,
// because nobody else can cover its bits before it.  So all such local maxima can be applied at
// once: the selected SET and every probe's gain at pick time are those of the sequential loop.
// Keys at pick time are strictly decreasing along the sequential pick sequence, therefore sorting
// the picks by that key (host part below) restores the sequential pick ORDER, which the reference's
// output order depends on (set.add() in pick order, filter/set_cover_filter.py:893-900).
//
// Local maxima are searched among a candidate list: every probe (of the current rank) with gain
// >= tau, i.e. a prefix of the global key order, so every probe with a larger key than a list
// member is itself a list member.  One round:
//   exchange: every GPU compacts its still-active list entries (gain >= tau) and PUSHES
//            (probe, gain, first interval, #intervals) into the exchange area of every GPU
//            (stores over NVLink into peer-mapped memory), then a cross-GPU barrier;
//   mark:    every GPU runs the conflict detection for ALL active candidates: one thread per
//            (candidate, interval), intervals of a remote candidate are read straight from the
//            owner's memory, atomicMax(mark[w], key) for each universe word with uncovered bits;
//   check:   a candidate is accepted iff mark[w] == its key in all of those words (word
//            granularity: false conflicts only postpone a pick);
//   apply:   every GPU applies ALL accepted probes to ITS OWN copy of the universe bit set and to
//            the gains of ITS OWN probes (interval index of the local probes only).
// The universe bit set, the marks and the accepted list are replicated (every GPU computes the
// same values from the same inputs); gains, the interval index and the candidate search -- the
// parts whose cost grows with the number of probes -- are sharded.  The only cross-GPU traffic per
// round is the candidate push (a few KB) and the interval reads of the active candidates.  On one
// GPU the same kernel runs with n_ranks = 1 and the exchange degenerates to a grid barrier.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <ctime>

#include "internal.cuh"

namespace {

constexpr int RT = 256;                     // threads per CTA
constexpr int NWARP = RT / 32;
constexpr int GAIN_LEVELS = 14;
constexpr int ID_LEVELS = 33;
constexpr int LIST_CAP_MAX = 4096;
constexpr int APPLY_WORDS = 64;             // winner intervals up to 64 words are staged in shared memory
constexpr unsigned long long WAIT_NS_DEFAULT = 30ull * 1000000000ull;   // a wait longer than this aborts the call

// Exchange area of one rank (cb_exchange_*): header, candidate slots, universe bit set, intervals.
struct XHeader {
    unsigned long long flag[CB_MAX_RANKS];        // flag[r]: last barrier epoch rank r announced to this rank
    unsigned long long epoch;                     // barrier epoch this rank has reached (persists across calls)
    unsigned long long pad[7];
    unsigned long long key[2][CB_MAX_RANKS];      // list rebuild: best key among rank r's probes
    uint32_t hist[2][CB_MAX_RANKS][64];           // list rebuild: level histogram of rank r's probes
    uint32_t list_n[2][CB_MAX_RANKS];             // list rebuild: length of rank r's candidate list
    // followed (at RParams::slot_off) by the lists [2][ranks][LIST_CAP_MAX] (uint4: probe, first interval,
    // intervals, -) and the per-round gains [2][ranks][LIST_CAP_MAX] (uint32) of every rank
};

struct RParams {
    int64_t n_probes;               // probes of the whole grouping (ids are global)
    int64_t lo, hi;                 // this rank's probes
    const int64_t *iv_off;          // [n_probes+1] local CSR (rows outside [lo, hi) are empty)
    const uint32_t *rank_idx;       // [n_probes] dense rank of every probe (only [lo, hi) is read)
    int32_t n_ranks_cover;          // number of distinct ranks (set_cover.py:349)
    uint32_t *gain;                 // [n_probes]
    unsigned long long *U;          // universe bit set of this rank (inside its exchange area)
    unsigned long long *mark;       // [u_words+1]
    int64_t u_words;
    // interval index of the LOCAL probes: items bucketed by the 64-position block of their start
    const int64_t *blk_off;         // [n_blocks+1]
    const uint2 *items;             // x = start, y = len << pbits | probe
    int64_t n_blocks;
    uint32_t max_item_len;
    int pbits;
    // control
    unsigned long long *remaining;  // uncovered bits still to cover (replicated)
    unsigned long long *barrier;    // local arrival counter
    unsigned long long *release;    // local release word of the cross-GPU barrier
    unsigned long long *key_local;  // [2]
    uint32_t *hist_local;           // [2][64]
    uint4 *list;                    // [list_cap] this rank's candidate list: probe, first interval, intervals, -
    uint32_t *list_n, list_cap;
    uint32_t *conf;                 // [LIST_CAP_MAX / 4] conflict flags (bytes) of the round's active candidates
    unsigned int *work;             // [2] next (winner, interval) pair to apply, alternating between rounds
    long long *sel, *n_sel;
    int *status;
    unsigned long long *phase_ns;   // [4]
    unsigned long long *ctr;        // [3]
    // the ranks
    int rank, n_ranks;
    unsigned char *xa[CB_MAX_RANKS];       // exchange area of every rank, as mapped HERE
    const uint2 *iv[CB_MAX_RANKS];         // cover intervals of every rank (inside its exchange area)
    int64_t slot_off;                      // byte offset of the list / gain slots in an exchange area
    unsigned long long wait_ns;            // a barrier wait longer than this aborts the call
    unsigned long long *diag;              // first wait that timed out: epoch << 8 | kind
};

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Spin until pred() holds; gives up (and flags the call as failed) after WAIT_NS, so that a rank that
// never arrives turns into an error on the others instead of a hang.
template <typename Pred>
__device__ __forceinline__ bool spin_until(const RParams &G, Pred pred, unsigned long long what = 0)
{
    if (pred()) return true;
    const unsigned long long t0 = globaltimer_ns();
    for (unsigned it = 1;; it++) {
        if (pred()) return true;
        if ((it & 0xfffu) == 0u) {
            if (*(volatile int *)G.status != 0) return false;
            if (globaltimer_ns() - t0 > G.wait_ns) {
                if (atomicCAS(G.status, 0, CB_ERR_COMM) == 0) *G.diag = what;
                return false;
            }
        }
    }
}

// Grid barrier of this GPU: monotone arrival counter, one arrival per CTA.
__device__ __forceinline__ void grid_barrier(const RParams &G, unsigned long long &target)
{
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(G.barrier, 1ull);
        const unsigned long long want = target;
        spin_until(G, [&] { return *(volatile unsigned long long *)G.barrier >= want; }, (want << 8) | 4ull);
        __threadfence();
    }
    __syncthreads();
}

// Cross-GPU barrier with a payload (n_ranks > 1; with one rank it is the plain grid barrier and the
// payload is NOT run: callers read their local copies instead).  All CTAs of this GPU arrive; CTA 0 then runs payload()
// (which stores this rank's contribution into the exchange area of every rank), announces the new
// epoch to every peer and waits for theirs, and releases the other CTAs.  After the call the
// contributions of ALL ranks for this epoch are visible in the local exchange area (read them with
// __ldcg).  Slots are double-buffered by epoch parity: a rank can be at most one barrier ahead of
// the slowest reader.
template <typename Payload>
__device__ __forceinline__ void xbarrier(const RParams &G, unsigned long long &target, unsigned long long &epoch,
                                         Payload payload)
{
    if (G.n_ranks == 1) {                         // nothing to exchange: the symmetric barrier is cheaper
        epoch++;
        grid_barrier(G, target);
        return;
    }
    __syncthreads();
    target += gridDim.x;
    epoch++;
    const unsigned long long e = epoch;
    if (blockIdx.x != 0) {
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(G.barrier, 1ull);
            spin_until(G, [&] { return ld_acquire_gpu(G.release) >= e; }, (e << 8) | 2ull);
        }
        __syncthreads();
        return;
    }
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(G.barrier, 1ull);
        const unsigned long long want = target;
        spin_until(G, [&] { return *(volatile unsigned long long *)G.barrier >= want; }, (e << 8) | 1ull);
        __threadfence();
    }
    __syncthreads();
    payload();
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < G.n_ranks && (int)threadIdx.x != G.rank) {
        XHeader *peer = reinterpret_cast<XHeader *>(G.xa[threadIdx.x]);
        XHeader *mine = reinterpret_cast<XHeader *>(G.xa[G.rank]);
        st_release_sys(&peer->flag[G.rank], e);
        spin_until(G, [&] { return ld_acquire_sys(&mine->flag[threadIdx.x]) >= e; }, (e << 8) | 3ull | ((unsigned long long)threadIdx.x << 4));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        st_release_gpu(G.release, e);
    }
    __syncthreads();
}

__device__ __forceinline__ uint4 *list_slot(const RParams &G, int on_rank, unsigned slot, int from_rank)
{
    return reinterpret_cast<uint4 *>(G.xa[on_rank] + G.slot_off) + ((size_t)slot * CB_MAX_RANKS + from_rank) * LIST_CAP_MAX;
}
__device__ __forceinline__ uint32_t *gain_slot(const RParams &G, int on_rank, unsigned slot, int from_rank)
{
    return reinterpret_cast<uint32_t *>(G.xa[on_rank] + G.slot_off + sizeof(uint4) * 2 * CB_MAX_RANKS * LIST_CAP_MAX) +
           ((size_t)slot * CB_MAX_RANKS + from_rank) * LIST_CAP_MAX;
}

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

__device__ __forceinline__ uint32_t popcount_range_cg(const unsigned long long *U, uint32_t s, uint32_t e)
{
    uint32_t c = 0;
    for_each_word(make_uint2(s, e), [&](uint32_t w, unsigned long long m) { c += __popcll(__ldcg(U + w) & m); });
    return c;
}

// apply() of ONE accepted interval r by ONE WARP: stage the interval's still-uncovered bits in the
// warp's slice of shared memory, subtract the uncovered bits of every overlap from the gain of the
// overlapping (local) interval's probe, then clear exactly the staged bits in this GPU's universe.
// The indexed items that can overlap r are ONE contiguous range: every item whose start lies in
// (r.x - max_item_len, r.y).  Reads stay inside the accepted interval and accepted intervals share no
// word with uncovered bits, so no barrier separates "update gains" from "clear U".
__device__ __forceinline__ void apply_interval_warp(const RParams &G, uint2 r, unsigned long long *s_uw, int lane)
{
    if (r.x >= r.y) return;
    const uint32_t w0 = r.x >> 6, w1 = (r.y - 1) >> 6;
    const uint32_t nwords = w1 - w0 + 1;
    const int64_t lo_pos = (int64_t)r.x - (int64_t)G.max_item_len + 1;
    const int64_t b_lo = (lo_pos > 0 ? lo_pos : 0) >> 6;
    int64_t b_hi = ((int64_t)r.y - 1) >> 6;
    if (b_hi >= G.n_blocks) b_hi = G.n_blocks - 1;
    const uint32_t pmask = (1u << G.pbits) - 1u;
    if (nwords > (uint32_t)APPLY_WORDS) {          // very long interval: count against L2, no staging
        const int64_t x0 = __ldg(G.blk_off + b_lo), x1 = __ldg(G.blk_off + b_hi + 1);
        for (int64_t x = x0 + lane; x < x1; x += 32) {
            const uint2 item = __ldg(G.items + x);
            const uint32_t os = max(item.x, r.x), oe = min(item.x + (item.y >> G.pbits), r.y);
            if (os < oe) {
                const uint32_t dlt = popcount_range_cg(G.U, os, oe);
                if (dlt) atomicSub(&G.gain[item.y & pmask], dlt);
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
    bool have = false;
#pragma unroll
    for (int h = 0; h < APPLY_WORDS / 32; h++) have |= mine[h] != 0ull;
    __syncwarp();                                   // the staged words are read by the other lanes below
    if (!__any_sync(0xffffffffu, have)) return;    // everything here is covered already
    const int64_t x0 = __ldg(G.blk_off + b_lo), x1 = __ldg(G.blk_off + b_hi + 1);
    constexpr int BATCH = 8;
    for (int64_t xb = x0 + lane; xb - lane < x1; xb += 32 * BATCH) {     // warp-uniform trip count
        uint2 item[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; u++) {
            const int64_t x = xb + (int64_t)u * 32;
            item[u] = x < x1 ? __ldg(G.items + x) : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < BATCH; u++) {
            const uint32_t os = max(item[u].x, r.x), oe = min(item[u].x + (item[u].y >> G.pbits), r.y);
            if (os < oe) {
                const uint32_t wa = (os >> 6) - w0, wb = ((oe - 1) >> 6) - w0;
                uint32_t dlt = 0;
                for (uint32_t q = wa; q <= wb; q++) {
                    unsigned long long m = ~0ull;
                    if (q == wa) m &= ~0ull << (os & 63);
                    if (q == wb) m &= ~0ull >> (63 - ((oe - 1) & 63));
                    dlt += __popcll(s_uw[q] & m);
                }
                if (dlt) atomicSub(&G.gain[item[u].y & pmask], dlt);
            }
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

// ---- set-up, pass 1 over the local intervals: universe bits, block counts of the index, gains
// (every bit of every interval is uncovered at the start, so gain = total length).  An interval
// longer than max_piece is indexed as several consecutive pieces (gains are additive over pieces).
template <bool SCATTER>
__global__ void index_kernel(const int64_t *__restrict__ iv_off, const uint2 *__restrict__ iv, int64_t lo, int64_t hi,
                             uint32_t max_piece, int pbits, unsigned long long *U, uint32_t *__restrict__ gain,
                             uint32_t *__restrict__ count, const int64_t *__restrict__ blk_off,
                             uint32_t *__restrict__ cursor, uint2 *__restrict__ items)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = lo + warp; p < hi; p += n_warps) {
        uint32_t c = 0;
        for (int64_t i = iv_off[p] + lane; i < iv_off[p + 1]; i += 32) {
            const uint2 r = iv[i];
            if (r.x >= r.y) continue;
            if (!SCATTER) {
                c += r.y - r.x;
                for_each_word(r, [&](uint32_t w, unsigned long long m) {
                    if ((U[w] & m) != m) atomicOr(&U[w], m);
                });
            }
            for (uint32_t s = r.x; s < r.y; s += max_piece) {
                const uint32_t len = min(max_piece, r.y - s);
                const uint32_t b = s >> 6;
                if (!SCATTER) atomicAdd(&count[b], 1u);
                else {
                    const uint32_t slot = atomicAdd(&cursor[b], 1u);
                    items[blk_off[b] + slot] = make_uint2(s, (len << pbits) | (uint32_t)p);
                }
            }
        }
        if (!SCATTER) {
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0) gain[p] = c;
        }
    }
}

__global__ void __launch_bounds__(RT, 2)
greedy_rounds_kernel(const RParams G)
{
    __shared__ unsigned long long s_key[NWARP];
    __shared__ unsigned long long s_u[NWARP * APPLY_WORDS];   // one staging area per warp
    __shared__ unsigned long long s_part[NWARP];
    __shared__ uint32_t s_hist[64];                  // level histogram of a list rebuild
    __shared__ uint32_t s_rn[CB_MAX_RANKS + 1];      // prefix of the per-rank list lengths
    __shared__ uint32_t s_conf[LIST_CAP_MAX / 4];    // conflict flags (one byte each) of the active candidates
    extern __shared__ uint32_t s_dyn[];
    // per CTA, list_cap entries each.  The candidate LIST (all ranks' entries, rank order): probe,
    // first interval, number of intervals, owner rank -- loaded once per list rebuild; gain -- loaded
    // every round.  The ACTIVE candidates of the round (list slots whose gain is still >= tau) and the
    // WINNERS among them: list slot + exclusive prefix of the interval counts.
    const uint32_t cap = G.list_cap;
    uint32_t *s_p = s_dyn, *s_i0 = s_p + cap, *s_cnt = s_i0 + cap, *s_g = s_cnt + cap, *s_base = s_g + cap,
             *s_wbase = s_base + cap + 1;
    uint16_t *s_act = reinterpret_cast<uint16_t *>(s_wbase + cap + 1), *s_wact = s_act + cap;
    unsigned char *s_own = reinterpret_cast<unsigned char *>(s_wact + cap);
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gsize = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    XHeader *xh = reinterpret_cast<XHeader *>(G.xa[G.rank]);
    const int R = G.n_ranks, me = G.rank;
    const uint32_t per = (cap + RT - 1) / RT;        // consecutive list slots handled by one thread

    // exclusive prefix of a (count, sum) pair over the CTA in thread order; returns the totals too
    auto block_scan2 = [&](uint32_t cnt, uint32_t sum, uint32_t &cnt_before, uint32_t &sum_before, uint32_t &cnt_all,
                           uint32_t &sum_all) {
        const unsigned long long v = ((unsigned long long)cnt << 40) | (unsigned long long)sum;   // sum < 2^40, cnt < 2^24
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        __syncthreads();                 // s_part may still be read from the previous scan
        if (lane == 31) s_part[warp] = inc;
        __syncthreads();
        unsigned long long before = 0, all = 0;
#pragma unroll
        for (int q = 0; q < NWARP; q++) {
            const unsigned long long t = s_part[q];
            if (q < warp) before += t;
            all += t;
        }
        const unsigned long long ex = before + inc - v;
        cnt_before = (uint32_t)(ex >> 40);
        sum_before = (uint32_t)(ex & 0xffffffffffull);
        cnt_all = (uint32_t)(all >> 40);
        sum_all = (uint32_t)(all & 0xffffffffffull);
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

    int cur_rank = 0;
    long long n_picks = 0;
    unsigned long long bar_target = 0, n_rebuilds = 0, n_rounds = 0, n_active_sum = 0;
    unsigned long long epoch = __ldcg(&xh->epoch);
    uint32_t tau = 1, id_thr = 0xffffffffu, n_list = 0;
    unsigned long long stamp = 0;                   // round stamp of the marks, < 2^(32 - pbits)
    unsigned rb = 0;
    bool need_rebuild = true;
    unsigned long long t_phase[4] = {0, 0, 0, 0}, t_last = 0, t_fine[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, t_f = 0;
    const bool timing = (gtid == 0);
    if (timing) t_last = globaltimer_ns();
    auto lap = [&](int i) {
        if (timing) {
            const unsigned long long t = globaltimer_ns();
            t_phase[i] += t - t_last;
            t_last = t;
        }
    };
    // finer split of a round for CB_TRACE: [0] gain exchange barrier, [1] compaction, [2] mark, [3] barrier, [4] check,
    // [5] barrier, [6] winners, [7] mark reset, [8] apply
    auto fine = [&](int i) {
        if (timing) {
            const unsigned long long t = globaltimer_ns();
            t_fine[i] += t - t_f;
            t_f = t;
        }
    };
    // CTA-uniform (a barrier inside): every thread of the CTA takes the same branch even if the status changes
    // while it is being read
    auto failed = [&]() -> bool { return __syncthreads_or(*(volatile int *)G.status != 0) != 0; };

    // ---- prologue: the universe is the union of every rank's intervals (utils/set_cover.py:302-320).
    // The set-up kernel ORed the local intervals into this rank's bit set; OR in the peers' words
    // (monotone, so reading a word a peer is still completing is harmless) and count the bits.
    xbarrier(G, bar_target, epoch, [] {});
    {
        unsigned long long c = 0;
        for (int64_t w = gtid; w < G.u_words; w += gsize) {
            unsigned long long v = __ldcg(G.U + w), mine = v;
            for (int r = 0; r < R; r++)
                if (r != me) v |= *(volatile const unsigned long long *)(G.xa[r] + ((const unsigned char *)G.U - G.xa[me]) + 8 * w);
            if (v != mine) G.U[w] = v;
            c += __popcll(v);
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0 && c) atomicAdd(G.remaining, c);
    }
    xbarrier(G, bar_target, epoch, [] {});           // nobody clears a bit while a peer may still read it

    for (;;) {
        if (failed()) break;
        if (need_rebuild) {
            // ---- largest key among the probes of the current rank, over all GPUs
            unsigned long long best = 0;
            for (int64_t p = G.lo + gtid; p < G.hi; p += gsize) {
                const uint32_t g = __ldcg(&G.gain[p]);
                if (g && G.rank_idx[p] == (uint32_t)cur_rank) {
                    const unsigned long long key = ((unsigned long long)g << 32) | (unsigned long long)(0xffffffffu - (uint32_t)p);
                    best = key > best ? key : best;
                }
            }
            best = block_max(best);
            const unsigned ks = rb & 1u;
            rb++;
            if (threadIdx.x == 0 && best) atomicMax(&G.key_local[ks], best);
            if (gtid == 0) *G.list_n = 0;
            xbarrier(G, bar_target, epoch, [&] {
                if ((int)threadIdx.x < R) {
                    XHeader *peer = reinterpret_cast<XHeader *>(G.xa[threadIdx.x]);
                    peer->key[epoch & 1ull][me] = __ldcg(&G.key_local[ks]);
                }
            });
            if (failed() || __ldcg(G.remaining) == 0ull) break;
            unsigned long long key = R == 1 ? __ldcg(&G.key_local[ks]) : 0ull;
            for (int r = 0; r < R && R > 1; r++) {
                const unsigned long long k2 = __ldcg(&xh->key[epoch & 1ull][r]);
                key = k2 > key ? k2 : key;
            }
            if (gtid == 0) G.key_local[ks ^ 1u] = 0ull;
            if (key == 0ull) {                  // rank exhausted (:522-526)
                cur_rank++;
                if (cur_rank >= G.n_ranks_cover) {
                    if (gtid == 0) atomicCAS(G.status, 0, CB_ERR_STATE);
                    break;
                }
                grid_barrier(G, bar_target);
                continue;
            }
            const uint32_t wmax = 0xffffffffu - (uint32_t)(key & 0xffffffffull);
            const uint32_t gmax = (uint32_t)(key >> 32);
            // ---- choose the list threshold from a histogram, so that the list holds as many of the
            // top candidates as fit: gain levels tau_0 = 1, tau_s = gmax - (gmax >> s) (s = 1..12),
            // tau_13 = gmax; and, should more probes than fit TIE at gmax, id levels
            // "id < wmax + ceil((P - wmax) / 2^j)" among those ties (sequential greedy consumes ties from
            // the low ids upwards, so the remaining ones sit in [wmax, P)).  Either way the list is a
            // prefix of the key order.
            auto tau_of = [&](int lv) -> uint32_t {
                if (lv == 0) return 1u;
                if (lv >= GAIN_LEVELS - 1) return gmax;
                const uint32_t t = gmax - (gmax >> lv);
                return t < 1u ? 1u : t;
            };
            auto idthr_of = [&](int j) -> unsigned long long {
                const unsigned long long span = (unsigned long long)G.n_probes - (unsigned long long)wmax;
                return (unsigned long long)wmax + ((span + (1ull << j) - 1ull) >> j);
            };
            if (threadIdx.x < 64) s_hist[threadIdx.x] = 0u;
            __syncthreads();
            for (int64_t p = G.lo + gtid; p < G.hi; p += gsize) {
                const uint32_t g = __ldcg(&G.gain[p]);
                if (g && G.rank_idx[p] == (uint32_t)cur_rank) {
                    int lv = GAIN_LEVELS - 1;
                    while (lv > 0 && g < tau_of(lv)) lv--;          // largest level the probe qualifies for
                    atomicAdd(&s_hist[lv], 1u);
                    if (g == gmax) {
                        int j = 0;
                        while (j + 1 < ID_LEVELS && (unsigned long long)p < idthr_of(j + 1)) j++;
                        atomicAdd(&s_hist[GAIN_LEVELS + j], 1u);
                    }
                }
            }
            __syncthreads();
            uint32_t *hist_now = G.hist_local + 64 * ks, *hist_next = G.hist_local + 64 * (ks ^ 1u);
            if (threadIdx.x < 64 && s_hist[threadIdx.x]) atomicAdd(&hist_now[threadIdx.x], s_hist[threadIdx.x]);
            xbarrier(G, bar_target, epoch, [&] {
                if (threadIdx.x < 64) {
                    const uint32_t v = __ldcg(&hist_now[threadIdx.x]);
                    for (int r = 0; r < R; r++)
                        reinterpret_cast<XHeader *>(G.xa[r])->hist[epoch & 1ull][me][threadIdx.x] = v;
                }
            });
            if (failed()) break;
            if (gtid < 64) hist_next[gtid] = 0u;                    // last read one rebuild ago
            // the 64 level counts of all ranks, fetched by 64 threads at once
            if (threadIdx.x < 64) {
                uint32_t c = 0;
                if (R == 1) c = __ldcg(&hist_now[threadIdx.x]);
                else
                    for (int r = 0; r < R; r++) c += __ldcg(&xh->hist[epoch & 1ull][r][threadIdx.x]);
                s_hist[threadIdx.x] = c;
            }
            __syncthreads();
            id_thr = 0xffffffffu;
            {
                // counts are suffix sums: a probe at level lv also qualifies for every wider level
                uint32_t c = 0;
                int pick = -1;
                for (int lv = GAIN_LEVELS - 1; lv >= 0; lv--) {
                    c += s_hist[lv];
                    if (c <= cap) pick = lv; else break;
                }
                if (pick >= 0) {
                    tau = tau_of(pick);
                } else {                                            // more ties at gmax than the list holds
                    tau = gmax;
                    c = 0;
                    int pj = -1;
                    for (int j = ID_LEVELS - 1; j >= 0; j--) {
                        c += s_hist[GAIN_LEVELS + j];
                        if (c <= cap) pj = j; else break;
                    }
                    // the smallest tied id is wmax; an id level that holds no tie at all (or none that
                    // fits) leaves the argmax alone in the list
                    unsigned long long thr = pj >= 0 ? idthr_of(pj) : 0ull;
                    if (thr <= (unsigned long long)wmax) thr = (unsigned long long)wmax + 1ull;
                    id_thr = thr > 0xffffffffull ? 0xffffffffu : (uint32_t)thr;
                }
            }
            __syncthreads();                                        // s_hist is reused by the next rebuild
            // ---- the list of this rank: probe, first interval, number of intervals
            for (int64_t p = G.lo + gtid; p < G.hi; p += gsize) {
                const uint32_t g = __ldcg(&G.gain[p]);
                if (g >= tau && (uint32_t)p < id_thr && G.rank_idx[p] == (uint32_t)cur_rank) {
                    const uint32_t slot = atomicAdd(G.list_n, 1u);
                    const int64_t x = G.iv_off[p], y = G.iv_off[p + 1];
                    if (slot < cap) G.list[slot] = make_uint4((uint32_t)p, (uint32_t)x, (uint32_t)(y - x), 0u);   // always, by the counts
                }
            }
            // ---- every rank's list goes to every rank; each CTA keeps all of them in shared memory
            xbarrier(G, bar_target, epoch, [&] {
                const unsigned slot = (unsigned)(epoch & 1ull);
                const uint32_t n_mine = min(__ldcg(G.list_n), cap);
                for (uint32_t i = threadIdx.x; i < n_mine; i += RT) {
                    const uint4 rec = __ldcg(&G.list[i]);
                    for (int r = 0; r < R; r++) list_slot(G, r, slot, me)[i] = rec;
                }
                if (threadIdx.x == 0)
                    for (int r = 0; r < R; r++) reinterpret_cast<XHeader *>(G.xa[r])->list_n[slot][me] = n_mine;
            });
            if (failed()) break;
            {
                const unsigned slot = (unsigned)(epoch & 1ull);
                if (threadIdx.x == 0) {
                    uint32_t acc = 0;
                    for (int r = 0; r < R; r++) {
                        s_rn[r] = acc;
                        acc += min(R == 1 ? __ldcg(G.list_n) : __ldcg(&xh->list_n[slot][r]), cap);
                    }
                    s_rn[R] = acc;
                }
                __syncthreads();
                n_list = min(s_rn[R], cap);
                for (uint32_t i = threadIdx.x; i < n_list; i += RT) {
                    int r = 0;
                    while (r + 1 < R && s_rn[r + 1] <= i) r++;
                    const uint4 rec = R == 1 ? __ldcg(&G.list[i]) : __ldcg(list_slot(G, me, slot, r) + (i - s_rn[r]));
                    s_p[i] = rec.x;
                    s_i0[i] = rec.y;
                    s_cnt[i] = rec.z;
                    s_own[i] = (unsigned char)r;
                }
                __syncthreads();
            }
            need_rebuild = false;
            n_rebuilds++;
            lap(0);
        }

        if (timing) t_f = globaltimer_ns();
        // ---- next round stamp; when it is about to wrap, every mark is wiped (visible to all after the
        // barrier that follows)
        stamp++;
        if (stamp >= (1ull << (32 - G.pbits)) - 1ull) {
            for (int64_t w = gtid; w <= G.u_words; w += gsize) G.mark[w] = 0ull;
            stamp = 1;
        }
        // ---- the current gains of the list entries.  One GPU: a grid barrier (gains are final after it),
        // then every CTA reads them where they are.  Several GPUs: CTA 0 pushes the gains of this rank's
        // entries to every rank inside the cross-GPU barrier.
        xbarrier(G, bar_target, epoch, [&] {
            const unsigned slot = (unsigned)(epoch & 1ull);
            const uint32_t lo_i = s_rn[me], n_mine = s_rn[me + 1] - s_rn[me];
            for (uint32_t i = threadIdx.x; i < n_mine; i += RT) {
                const uint32_t g = __ldcg(&G.gain[s_p[lo_i + i]]);
                for (int r = 0; r < R; r++) gain_slot(G, r, slot, me)[i] = g;
            }
        });
        if (failed() || __ldcg(G.remaining) == 0ull) break;
        fine(0);
        // ---- active candidates = list slots whose gain is still >= tau, compacted in list order by every
        // CTA for itself (same inputs, same result): thread t looks at `per` consecutive slots
        uint32_t n_act = 0, total_pairs = 0;
        {
            const unsigned slot = (unsigned)(epoch & 1ull);
            const uint32_t i_lo = threadIdx.x * per, i_hi = min(i_lo + per, n_list);
            uint32_t c = 0, sum = 0;
            constexpr int PER_MAX = LIST_CAP_MAX / RT;
            uint32_t gv[PER_MAX];
#pragma unroll
            for (int j = 0; j < PER_MAX; j++) {                  // independent loads: all in flight together
                const uint32_t i = i_lo + (uint32_t)j;
                gv[j] = 0u;
                if ((uint32_t)j < per && i < i_hi) {
                    if (R == 1) gv[j] = __ldcg(&G.gain[s_p[i]]);
                    else {
                        const int r = s_own[i];
                        gv[j] = __ldcg(gain_slot(G, me, slot, r) + (i - s_rn[r]));
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < PER_MAX; j++) {
                const uint32_t i = i_lo + (uint32_t)j;
                if ((uint32_t)j < per && i < i_hi) {
                    s_g[i] = gv[j];
                    if (gv[j] >= tau) { c++; sum += s_cnt[i]; }
                }
            }
            uint32_t cb, sb;
            block_scan2(c, sum, cb, sb, n_act, total_pairs);
            for (uint32_t i = i_lo; i < i_hi; i++)
                if (s_g[i] >= tau) {
                    s_act[cb] = (uint16_t)i;
                    s_base[cb] = sb;
                    cb++;
                    sb += s_cnt[i];
                }
            if (threadIdx.x == 0) s_base[n_act] = total_pairs;
            // the conflict bits of the previous round have been read by everybody (barrier since)
            if (blockIdx.x == 0)
                for (uint32_t q = threadIdx.x; q < LIST_CAP_MAX / 4; q += RT) G.conf[q] = 0u;
            __syncthreads();
        }
        if (n_act == 0u) {                       // the list is used up
            need_rebuild = true;
            continue;
        }
        fine(1);
        // (candidate a, interval f - s_base[a]) for the f-th pair of the round
        auto pair_of = [&](uint32_t f, uint32_t &a, uint32_t &slot_i) -> uint2 {
            uint32_t lo = 0, hi = n_act;
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (s_base[mid] <= f) lo = mid; else hi = mid;
            }
            a = lo;
            slot_i = s_act[lo];
            return G.iv[s_own[slot_i]][(int64_t)s_i0[slot_i] + (f - s_base[lo])];      // peer memory when the owner is remote
        };
        auto key_of = [&](uint32_t slot_i) -> unsigned long long {
            return ((unsigned long long)s_g[slot_i] << 32) | (unsigned long long)(0xffffffffu - s_p[slot_i]);
        };
        // Mark value of a candidate: round stamp | gain | (2^pbits - 1 - probe).  Within a round the order is
        // the key order; a newer round beats every older mark, so marks are never reset between rounds (they
        // are wiped when the stamp is about to wrap, see below).
        const int pbits = G.pbits;
        auto mkey_of = [&](uint32_t slot_i) -> unsigned long long {
            return (stamp << (32 + pbits)) | ((unsigned long long)s_g[slot_i] << pbits) |
                   (unsigned long long)(((1u << pbits) - 1u) - s_p[slot_i]);
        };

        // ---- mark: one thread per (candidate, interval).  The first MC pairs of a thread stay in registers
        // (interval, candidate, which of its words still hold uncovered bits) for the check phase; all loads
        // of a step are issued before any is used.
        constexpr int MC = 3, MW = 8;
        uint2 c_r[MC];
        uint32_t c_a[MC], c_si[MC], c_live[MC];
#pragma unroll
        for (int k = 0; k < MC; k++) {
            const int64_t f = gtid + (int64_t)k * gsize;
            c_r[k] = make_uint2(0u, 0u);
            c_a[k] = c_si[k] = c_live[k] = 0u;
            if (f < (int64_t)total_pairs) c_r[k] = pair_of((uint32_t)f, c_a[k], c_si[k]);
        }
#pragma unroll
        for (int k = 0; k < MC; k++) {
            const uint2 r = c_r[k];
            if (r.x >= r.y) continue;
            const uint32_t w0 = r.x >> 6, w1 = (r.y - 1) >> 6;
            if (w1 - w0 < (uint32_t)MW) {
                unsigned long long u[MW];
#pragma unroll
                for (int j = 0; j < MW; j++) u[j] = (w0 + j <= w1) ? __ldcg(G.U + w0 + j) : 0ull;
                uint32_t live = 0;
#pragma unroll
                for (int j = 0; j < MW; j++) {
                    unsigned long long m = ~0ull;
                    if (j == 0) m &= ~0ull << (r.x & 63);
                    if (w0 + j == w1) m &= ~0ull >> (63 - ((r.y - 1) & 63));
                    if (w0 + j <= w1 && (u[j] & m)) live |= 1u << j;
                }
                c_live[k] = live;
                const unsigned long long mk = mkey_of(c_si[k]);
#pragma unroll
                for (int j = 0; j < MW; j++)
                    if ((live >> j) & 1u) atomicMax(&G.mark[w0 + j], mk);
            } else {                                    // long interval: word by word, re-read in the check phase
                c_live[k] = 0x80000000u;
                const unsigned long long mk = mkey_of(c_si[k]);
                for_each_word(r, [&](uint32_t w, unsigned long long m) {
                    if (__ldcg(G.U + w) & m) atomicMax(&G.mark[w], mk);
                });
            }
        }
        for (int64_t f = gtid + (int64_t)MC * gsize; f < (int64_t)total_pairs; f += gsize) {     // rarely: more pairs
            uint32_t a, si;
            const uint2 r = pair_of((uint32_t)f, a, si);
            const unsigned long long mk = mkey_of(si);
            for_each_word(r, [&](uint32_t w, unsigned long long m) {
                if (__ldcg(G.U + w) & m) atomicMax(&G.mark[w], mk);
            });
        }
        fine(2);
        grid_barrier(G, bar_target);
        fine(3);

        // ---- check: accepted iff the candidate holds the mark of every word it marked
#pragma unroll
        for (int k = 0; k < MC; k++) {
            const uint2 r = c_r[k];
            if (r.x >= r.y) continue;
            const unsigned long long mk = mkey_of(c_si[k]);
            const uint32_t w0 = r.x >> 6;
            bool conflict = false;
            if (!(c_live[k] & 0x80000000u)) {
                unsigned long long mv[MW];
#pragma unroll
                for (int j = 0; j < MW; j++) mv[j] = ((c_live[k] >> j) & 1u) ? __ldcg(G.mark + w0 + j) : mk;
#pragma unroll
                for (int j = 0; j < MW; j++) conflict |= mv[j] != mk;
            } else {
                for_each_word(r, [&](uint32_t w, unsigned long long m) {
                    if ((__ldcg(G.U + w) & m) && __ldcg(G.mark + w) != mk) conflict = true;
                });
            }
            if (conflict) reinterpret_cast<volatile unsigned char *>(G.conf)[c_a[k]] = 1;     // plain byte store: no contention
        }
        for (int64_t f = gtid + (int64_t)MC * gsize; f < (int64_t)total_pairs; f += gsize) {
            uint32_t a, si;
            const uint2 r = pair_of((uint32_t)f, a, si);
            const unsigned long long mk = mkey_of(si);
            bool conflict = false;
            for_each_word(r, [&](uint32_t w, unsigned long long m) {
                if ((__ldcg(G.U + w) & m) && __ldcg(G.mark + w) != mk) conflict = true;
            });
            if (conflict) reinterpret_cast<volatile unsigned char *>(G.conf)[a] = 1;
        }
        fine(4);
        grid_barrier(G, bar_target);
        if (failed()) break;
        fine(5);

        // ---- winners = active candidates without a conflict (same compaction in every CTA)
        uint32_t n_win = 0, win_pairs = 0;
        {
            for (uint32_t q = threadIdx.x; q < (n_act + 3) / 4; q += RT) s_conf[q] = __ldcg(&G.conf[q]);
            __syncthreads();
            const unsigned char *s_cf = reinterpret_cast<const unsigned char *>(s_conf);
            const uint32_t pa = (n_act + RT - 1) / RT;
            const uint32_t a_lo = min(threadIdx.x * pa, n_act), a_hi = min(a_lo + pa, n_act);
            uint32_t c = 0, sum = 0;
            for (uint32_t a = a_lo; a < a_hi; a++)
                if (!s_cf[a]) { c++; sum += s_cnt[s_act[a]]; }
            uint32_t cb, sb;
            block_scan2(c, sum, cb, sb, n_win, win_pairs);
            for (uint32_t a = a_lo; a < a_hi; a++)
                if (!s_cf[a]) {
                    const uint32_t si = s_act[a];
                    s_wact[cb] = (uint16_t)si;
                    s_wbase[cb] = sb;
                    if (blockIdx.x == 0) G.sel[n_picks + cb] = (long long)key_of(si);
                    cb++;
                    sb += s_cnt[si];
                }
            if (threadIdx.x == 0) s_wbase[n_win] = win_pairs;
            __syncthreads();
        }
        lap(1);
        fine(6);
        n_rounds++;
        n_active_sum += n_act;
        n_picks += n_win;

        fine(7);
        // ---- apply every accepted probe.  The cost of a (winner, interval) pair varies a lot (number of indexed
        // items it overlaps, contention on the gains), so the pairs are handed out dynamically: a warp takes the
        // next one from a global counter when it is done with its own.  Two counters alternate between rounds; the
        // idle one is reset here, a full round (three barriers) before it is used again.
        {
            unsigned int *work = G.work + (n_rounds & 1ull);
            if (gtid == 0) G.work[(n_rounds & 1ull) ^ 1ull] = 0u;
            for (;;) {
                uint32_t f = 0;
                if (lane == 0) f = atomicAdd(work, 1u);
                f = __shfl_sync(0xffffffffu, f, 0);
                if (f >= win_pairs) break;
                uint32_t lo = 0, hi = n_win;
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (s_wbase[mid] <= f) lo = mid; else hi = mid;
                }
                const uint32_t si = s_wact[lo];
                const uint2 r = G.iv[s_own[si]][(int64_t)s_i0[si] + (f - s_wbase[lo])];
                apply_interval_warp(G, r, s_u + warp * APPLY_WORDS, lane);
            }
        }
        lap(2);
        fine(8);
    }
    if (gtid == 0) {
        for (int i = 0; i < 12; i++) G.ctr[8 + i] = t_fine[i];
        xh->epoch = epoch;
        *G.n_sel = n_picks;
        for (int i = 0; i < 4; i++) G.phase_ns[i] = t_phase[i];
        G.ctr[0] = n_rebuilds;
        G.ctr[1] = n_rounds;
        G.ctr[2] = n_active_sum;
    }
}

// CB_TRACE=1: host-side timestamps of the steps of a call on stderr (debugging of multi-rank runs)
void trace(const cb_ctx *ctx, const char *what)
{
    static const bool on = getenv("CB_TRACE") != nullptr;
    if (!on) return;
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    fprintf(stderr, "[cb %p xrank %d] %.6f %s\n", (const void *)ctx, ctx->xrank, ts.tv_sec % 1000 + ts.tv_nsec * 1e-9, what);
}

size_t slot_bytes() { return (sizeof(uint4) + sizeof(uint32_t)) * 2 * CB_MAX_RANKS * (size_t)LIST_CAP_MAX; }
size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

// Bytes of exchange area the set cover over `cover` needs on this rank.
int64_t cb_rounds_exchange_bytes(const cb_cover *cover)
{
    const size_t u_words = (size_t)(cover->universe_bits >> 6) + 1;
    return (int64_t)(align_up(sizeof(XHeader)) + align_up(slot_bytes()) + align_up(8 * u_words) +
                     align_up(sizeof(uint2) * (size_t)(cover->n_intervals ? cover->n_intervals : 1)));
}

// One set cover in two steps.  prepare(): everything that is local to this rank (work buffers, universe
// bits, gains, interval index) -- it never waits for another rank.  run(): the persistent kernel and the
// read-back of the picks -- with several ranks this is the collective part.
struct cb_rounds_job {
    cb_ctx *ctx = nullptr;
    bool sharded = false;
    int64_t P = 0, E = 0, n_items = 0;
    int32_t n_ranks = 1;
    uint32_t list_cap = 3072u;
    std::vector<uint32_t> h_rank;
    DevBuf<unsigned char> d_area;
    DevBuf<uint32_t> d_rank, d_gain, d_bcount, d_bcursor, d_small;
    DevBuf<unsigned long long> d_mark, d_ctl;
    DevBuf<long long> d_sel;
    DevBuf<int64_t> d_boff;
    DevBuf<uint2> d_items;
    RParams G;
    EventTimer t_all, t_uni, t_greedy;
    explicit cb_rounds_job(cb_ctx *c) : ctx(c), t_all(c->stream), t_uni(c->stream), t_greedy(c->stream) {}
    int prepare(const cb_cover *cover, int64_t lo, int64_t hi, const int32_t *ranks, bool sharded_);
    int run(int64_t *sel_ids, int64_t *n_sel, cb_stats *stats);
};

int cb_rounds_job::prepare(const cb_cover *cover, int64_t lo, int64_t hi, const int32_t *ranks, bool sharded_)
{
    cudaStream_t st = ctx->stream;
    sharded = sharded_;
    P = cover->n_probes;
    E = cover->n_intervals;
    const int R = sharded ? ctx->xn_ranks : 1, me = sharded ? ctx->xrank : 0;
    if (lo < 0 || hi < lo || hi > P) return cb_fail(ctx, CB_ERR_ARG, "bad probe range");
    if (sharded && ctx->xarea_poisoned)
        return cb_fail(ctx, CB_ERR_STATE, "an earlier sharded call failed; attach the exchange areas again");
    if (P >= (1ll << 25)) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "more than 2^25 probes in one grouping");
    if (E >= 0xfffffff0ll) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "more than 2^32 intervals in one grouping");
    trace(ctx, "rounds: prepare");
    t_all.start();
    t_uni.start();
    const int wide = ctx->sm_count * 8;
    const int64_t u_words = cover->universe_bits >> 6;
    const int64_t n_blocks = u_words + 1;

    list_cap = 3072u;                             // <= LIST_CAP_MAX; with 2 CTAs per SM the measured optimum (profiles/README_r02.md)
    if (const char *e = getenv("CB_GREEDY_LIST_CAP")) {
        const int v = atoi(e);
        if (v >= 1 && v <= LIST_CAP_MAX) list_cap = (uint32_t)v;
    }

    // ---- exchange area: the context's shared one when sharded, a work buffer otherwise
    const size_t off_slots = align_up(sizeof(XHeader));
    const size_t off_U = off_slots + align_up(slot_bytes());
    const size_t off_iv = off_U + align_up(8 * ((size_t)u_words + 1));
    const size_t need = off_iv + align_up(sizeof(uint2) * (size_t)(E ? E : 1));
    unsigned char *area = nullptr;
    if (sharded) {
        if (R < 2 || !ctx->xarea) return cb_fail(ctx, CB_ERR_STATE, "cb_exchange_attach has not been called");
        if (need > ctx->xarea_bytes) return cb_fail(ctx, CB_ERR_STATE, "exchange area too small (cb_exchange_alloc)");
        area = ctx->xarea;
    } else {
        CB_CUDA(ctx, d_area.alloc(off_iv));            // intervals stay where they are
        area = d_area.p;
        CB_CUDA(ctx, cudaMemsetAsync(area, 0, sizeof(XHeader), st));
    }
    unsigned long long *d_U = reinterpret_cast<unsigned long long *>(area + off_U);
    CB_CUDA(ctx, cudaMemsetAsync(d_U, 0, 8 * ((size_t)u_words + 1), st));
    const uint2 *d_iv_mine = cover->d_iv;
    if (sharded) {
        if (E) CB_CUDA(ctx, cudaMemcpyAsync(area + off_iv, cover->d_iv, sizeof(uint2) * (size_t)E, cudaMemcpyDeviceToDevice, st));
        d_iv_mine = reinterpret_cast<const uint2 *>(area + off_iv);
    }

    // ---- ranks -> dense indices in ascending order of rank value (:349)
    n_ranks = 1;
    CB_CUDA(ctx, d_rank.alloc((size_t)(P ? P : 1)));
    if (ranks) {
        h_rank.assign((size_t)P, 0u);
        std::vector<int32_t> vals(ranks, ranks + P);
        std::sort(vals.begin(), vals.end());
        vals.erase(std::unique(vals.begin(), vals.end()), vals.end());
        n_ranks = (int32_t)vals.size();
        for (int64_t p = 0; p < P; p++)
            h_rank[(size_t)p] = (uint32_t)(std::lower_bound(vals.begin(), vals.end(), ranks[p]) - vals.begin());
        CB_CUDA(ctx, cudaMemcpyAsync(d_rank.p, h_rank.data(), sizeof(uint32_t) * (size_t)P, cudaMemcpyHostToDevice, st));
    } else {
        CB_CUDA(ctx, cudaMemsetAsync(d_rank.p, 0, sizeof(uint32_t) * (size_t)(P ? P : 1), st));
    }

    // ---- work buffers
    int pbits = 1;
    while ((1ll << pbits) < P) pbits++;
    const uint32_t max_piece = (1u << (32 - pbits)) - 1u;
    CB_CUDA(ctx, d_gain.alloc((size_t)(P ? P : 1)));
    CB_CUDA(ctx, cudaMemsetAsync(d_gain.p, 0, sizeof(uint32_t) * (size_t)(P ? P : 1), st));
    CB_CUDA(ctx, d_bcount.alloc((size_t)n_blocks));
    CB_CUDA(ctx, d_bcursor.alloc((size_t)n_blocks));
    CB_CUDA(ctx, d_boff.alloc((size_t)n_blocks + 1));
    CB_CUDA(ctx, cudaMemsetAsync(d_bcount.p, 0, sizeof(uint32_t) * (size_t)n_blocks, st));
    CB_CUDA(ctx, cudaMemsetAsync(d_bcursor.p, 0, sizeof(uint32_t) * (size_t)n_blocks, st));
    CB_CUDA(ctx, d_mark.alloc((size_t)u_words + 1));
    CB_CUDA(ctx, cudaMemsetAsync(d_mark.p, 0, 8 * ((size_t)u_words + 1), st));
    // control block: [0] remaining, [1] barrier, [2] release, [3..4] key_local, [5..8] phase ns, [9..11] counters,
    // [12] n_sel, [13] status, [14] diagnostics
    CB_CUDA(ctx, d_ctl.alloc(40));
    CB_CUDA(ctx, cudaMemsetAsync(d_ctl.p, 0, 8 * 40, st));
    // small u32 block: list[4 * list_cap] (uint4 entries), conf[LIST_CAP_MAX / 4], hist_local[128], list_n[4], work[4]
    const size_t n_small = 4 * (size_t)list_cap + LIST_CAP_MAX / 4 + 128 + 4 + 4;
    CB_CUDA(ctx, d_small.alloc(n_small));
    CB_CUDA(ctx, cudaMemsetAsync(d_small.p, 0, sizeof(uint32_t) * n_small, st));
    CB_CUDA(ctx, d_sel.alloc((size_t)(P ? P : 1)));

    // ---- set-up: universe bits + block counts + gains, offsets, index items
    index_kernel<false><<<wide, 256, 0, st>>>(cover->d_iv_off, cover->d_iv, lo, hi, max_piece, pbits, d_U, d_gain.p,
                                              d_bcount.p, nullptr, nullptr, nullptr);
    ctx->launches++;
    n_items = 0;
    CB_TRY(cb_exclusive_scan_u32_to_i64(ctx, d_bcount.p, d_boff.p, n_blocks, &n_items));
    if (n_items >= 0xfffffff0ll) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "more than 2^32 index items");
    CB_CUDA(ctx, d_items.alloc((size_t)(n_items ? n_items : 1)));
    index_kernel<true><<<wide, 256, 0, st>>>(cover->d_iv_off, cover->d_iv, lo, hi, max_piece, pbits, nullptr, nullptr,
                                             nullptr, d_boff.p, d_bcursor.p, d_items.p);
    ctx->launches++;
    CB_CUDA(ctx, cudaGetLastError());
    t_uni.stop();

    memset(&G, 0, sizeof G);
    G.n_probes = P;
    G.lo = lo;
    G.hi = hi;
    G.iv_off = cover->d_iv_off;
    G.rank_idx = d_rank.p;
    G.n_ranks_cover = n_ranks;
    G.gain = d_gain.p;
    G.U = d_U;
    G.mark = d_mark.p;
    G.u_words = u_words;
    G.blk_off = d_boff.p;
    G.items = d_items.p;
    G.n_blocks = n_blocks;
    G.max_item_len = std::min(cover->max_interval_len, max_piece);
    G.pbits = pbits;
    G.remaining = d_ctl.p;
    G.barrier = d_ctl.p + 1;
    G.release = d_ctl.p + 2;
    G.key_local = d_ctl.p + 3;
    G.phase_ns = d_ctl.p + 5;
    G.ctr = d_ctl.p + 9;
    G.n_sel = reinterpret_cast<long long *>(d_ctl.p + 12);
    G.status = reinterpret_cast<int *>(d_ctl.p + 13);
    G.diag = d_ctl.p + 14;
    G.list = reinterpret_cast<uint4 *>(d_small.p);           // cudaMallocAsync blocks are 256-byte aligned
    G.list_cap = list_cap;
    G.conf = d_small.p + 4 * (size_t)list_cap;
    G.hist_local = G.conf + LIST_CAP_MAX / 4;
    G.list_n = G.hist_local + 128;
    G.work = G.list_n + 4;
    G.sel = d_sel.p;
    G.rank = me;
    G.n_ranks = R;
    G.slot_off = (int64_t)off_slots;
    G.wait_ns = WAIT_NS_DEFAULT;
    if (const char *e = getenv("CB_WAIT_MS")) if (atoll(e) > 0) G.wait_ns = (unsigned long long)atoll(e) * 1000000ull;
    for (int r = 0; r < R; r++) {
        G.xa[r] = sharded ? ctx->xpeer[r] : area;
        G.iv[r] = sharded ? reinterpret_cast<const uint2 *>(ctx->xpeer[r] + off_iv) : d_iv_mine;
    }
    G.xa[me] = area;
    G.iv[me] = d_iv_mine;
    {
        // once per device: changing a function attribute waits for running instances of the function, which
        // would stall a rank behind another rank's persistent kernel when several contexts share a device
        static bool attr_set[64] = {};
        const size_t dyn_max = sizeof(uint32_t) * (6 * (size_t)LIST_CAP_MAX + 2) + 5 * (size_t)LIST_CAP_MAX;
        if (ctx->device < 0 || ctx->device >= 64 || !attr_set[ctx->device]) {
            CB_CUDA(ctx, cudaFuncSetAttribute(greedy_rounds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_max));
            if (ctx->device >= 0 && ctx->device < 64) attr_set[ctx->device] = true;
        }
    }
    trace(ctx, "rounds: prepared");
    return CB_OK;
}

int cb_rounds_job::run(int64_t *sel_ids, int64_t *n_sel, cb_stats *stats)
{
    cudaStream_t st = ctx->stream;
    const int R = G.n_ranks, me = G.rank;
    // ---- persistent launch: cooperative (co-residency guaranteed), as many blocks as the device holds (<= 2 per SM)
    const size_t dyn_smem = sizeof(uint32_t) * (6 * (size_t)list_cap + 2) + 5 * (size_t)list_cap;
    int per_sm = 0;
    CB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, greedy_rounds_kernel, RT, dyn_smem));
    if (per_sm < 1) return cb_fail(ctx, CB_ERR_CUDA, "greedy kernel does not fit on an SM");
    int want = 2;
    if (const char *e = getenv("CB_GREEDY_BLOCKS_PER_SM")) want = atoi(e) > 0 ? atoi(e) : want;
    if (per_sm > want) per_sm = want;
    int grid = per_sm * ctx->sm_count;
    if (ctx->xgrid_limit > 0 && grid > ctx->xgrid_limit) grid = ctx->xgrid_limit;   // several ranks on one device (tests)
    void *args[] = {(void *)&G};
    trace(ctx, "rounds: launching the greedy kernel");
    t_greedy.start();
    if (ctx->xgrid_limit > 0 && sharded) {
        // several ranks share this device (tests): cooperative launches of different streams do not overlap,
        // so the kernels go out as plain launches; the caller's grid limit keeps all of them co-resident
        greedy_rounds_kernel<<<grid, RT, dyn_smem, st>>>(G);
        CB_CUDA(ctx, cudaGetLastError());
    } else {
        CB_CUDA(ctx, cudaLaunchCooperativeKernel((void *)greedy_rounds_kernel, dim3(grid), dim3(RT), args, dyn_smem, st));
    }
    ctx->launches++;
    t_greedy.stop();
    t_all.stop();

    unsigned long long h_ctl[40];
    CB_CUDA(ctx, cudaMemcpyAsync(h_ctl, d_ctl.p, sizeof h_ctl, cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    trace(ctx, "rounds: greedy kernel finished");
    if (getenv("CB_TRACE"))
        fprintf(stderr, "[cb] round phases (us, CTA 0): xchg %.0f compact %.0f mark %.0f bar %.0f check %.0f bar %.0f winners %.0f reset %.0f "
                        "apply %.0f | rebuilds %llu rounds %llu\n", h_ctl[17] / 1e3, h_ctl[18] / 1e3, h_ctl[19] / 1e3, h_ctl[20] / 1e3,
                h_ctl[21] / 1e3, h_ctl[22] / 1e3, h_ctl[23] / 1e3, h_ctl[24] / 1e3, h_ctl[25] / 1e3, h_ctl[9], h_ctl[10]);
    const long long h_nsel = (long long)h_ctl[12];
    const int h_status = (int)(uint32_t)h_ctl[13];
    if (h_status == CB_ERR_COMM) {
        ctx->xarea_poisoned = sharded;
        char msg[256];
        snprintf(msg, sizeof msg, "set cover: a barrier wait exceeded the time limit (rank %d of %d, barrier %llu, kind %llu, "
                 "peer %llu)", me, R, (unsigned long long)(h_ctl[14] >> 8), (unsigned long long)(h_ctl[14] & 15),
                 (unsigned long long)((h_ctl[14] >> 4) & 15));
        ctx->err = msg;
        return CB_ERR_COMM;
    }
    if (h_status != 0) return cb_fail(ctx, CB_ERR_STATE, "set cover ran out of ranks before reaching the requested coverage");
    *n_sel = 0;
    if (h_nsel > 0) {
        static_assert(sizeof(long long) == sizeof(int64_t), "int64");
        CB_CUDA(ctx, cudaMemcpyAsync(sel_ids, d_sel.p, sizeof(int64_t) * (size_t)h_nsel, cudaMemcpyDeviceToHost, st));
        CB_CUDA(ctx, cudaStreamSynchronize(st));
        // the kernel reports each pick's key at pick time, (gain << 32) | (2^32-1 - id), in no
        // particular order inside a round; the sequential loop picks rank by rank and, inside a
        // rank, in strictly decreasing key order
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
    *n_sel = h_nsel;
    if (stats) {
        stats->ms_universe = t_uni.ms();
        stats->ms_greedy = t_greedy.ms();
        stats->ms_total = t_all.ms();
        stats->n_picks = h_nsel;
        stats->n_intervals = E;
        stats->n_kernel_launches = ctx->launches;
        for (int i = 0; i < 4; i++) stats->reserved[i] = (int64_t)h_ctl[5 + i];   // ns: list rebuilds, exchange+mark+check, apply, -
        stats->reserved[4] = (int64_t)h_ctl[9];      // candidate-list rebuilds
        stats->reserved[5] = (int64_t)h_ctl[10];     // rounds
        stats->reserved[6] = (int64_t)h_ctl[11];     // active candidates summed over the rounds
        stats->reserved[7] = n_items;
    }
    return CB_OK;
}

// `cover` holds the intervals of probes [lo, hi) of a grouping of cover->n_probes probes (rows outside
// are empty).  With sharded = true every rank of the exchange group must make this call with its
// own shard; all of them receive the same picks.
int cb_setcover_rounds_impl(cb_ctx *ctx, const cb_cover *cover, int64_t lo, int64_t hi, const int32_t *ranks,
                            bool sharded, int64_t *sel_ids, int64_t *n_sel, cb_stats *stats)
{
    cb_rounds_job job(ctx);
    CB_TRY(job.prepare(cover, lo, hi, ranks, sharded));
    return job.run(sel_ids, n_sel, stats);
}

int cb_rounds_begin_impl(cb_ctx *ctx, const cb_cover *cover, int64_t lo, int64_t hi, const int32_t *ranks,
                         cb_rounds_job **out)
{
    cb_rounds_job *job = new cb_rounds_job(ctx);
    const int rc = job->prepare(cover, lo, hi, ranks, true);
    if (rc == CB_OK) {
        cudaStreamSynchronize(ctx->stream);           // nothing of the set-up is still queued when run() starts
        *out = job;
    } else {
        delete job;
    }
    return rc;
}

int cb_rounds_end_impl(cb_rounds_job *job, int64_t *sel_ids, int64_t *n_sel, cb_stats *stats)
{
    const int rc = job->run(sel_ids, n_sel, stats);
    delete job;
    return rc;
}
