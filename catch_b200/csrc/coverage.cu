// Stage A: probe coverage of the target genomes (K2 seed index, K3 scan, K4 merge).
//
// Replaces (reference paths relative to catch/):
//   probe.py:356-504,684-763   seed-map construction  -> seed_index_* kernels (bucketed hash CSR)
//   probe.py:1008-1119         _find_probe_covers_in_subsequence -> scan_kernel
//   utils/longest_common_substring.py:59-159 k_lcf_around_anchor -> anchored_extend (bit scans)
//   probe.py:1328-1344         lcf() predicate -> inside process_hit
//   utils/interval.py:288-316  merge_overlapping, filter/set_cover_filter.py:429-439,462-466
//                              -> emitted ranges are extended/clipped/offset at emit time and
//                                 merged per probe by merge_kernel
//
// Semantics kept bit for bit: a (probe, diagonal) pair is examined iff one of the probe's
// SELECTED seeds matches the target exactly at an in-bounds position; every matching seed gives
// its own anchored range; ranges are unioned.  One thread owns a (probe, diagonal): the thread
// that arrived through the smallest matching seed; it walks the remaining seeds of the probe.
#include <cstring>

#include "internal.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int POS_PER_THREAD = CB_TILE / SCAN_THREADS;
constexpr int TW = CB_TILE_WORDS;

struct ScanParams {
    // targets
    const uint64_t *planes;
    int64_t plane_words;
    int bits;
    int64_t total_bases;
    int64_t n_seqs;
    const int64_t *seq_start;
    const uint32_t *seq_ubase;
    // probes
    const uint64_t *pwords;
    const int32_t *plen;
    // seeds (CSR, ascending within a probe)
    const uint32_t *seed_off;
    const uint8_t *seed_pos;
    // seed index
    const int64_t *bucket_off;
    const uint64_t *entries;
    uint32_t bucket_mask;
    // hybridisation model
    int m, lcf, island, ext, k;
    // output
    uint32_t *rec_count;          // count pass: ranges per probe
    const int64_t *rec_off;       // emit pass
    uint32_t *rec_cursor;
    uint64_t *rec;
    // scheduling / stats
    int64_t n_tiles;
    unsigned long long *tile_counter;
    unsigned long long *stat_hits;
    unsigned long long *stat_lookups;
};

// ---------------------------------------------------------------------------------------
// small multi-word bit-mask helpers (NW 64-bit words, bit j = probe position j)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t range_word(int lo, int hi, int wi)
{
    // bits of word wi inside [lo, hi)
    int l = lo - wi * 64, h = hi - wi * 64;
    if (l < 0) l = 0;
    if (h > 64) h = 64;
    if (l >= h) return 0ull;
    uint64_t m = (h == 64) ? ~0ull : ((1ull << h) - 1ull);
    return m & (~0ull << l);
}

template <int NW>
__device__ __forceinline__ bool any_in_range(const uint64_t (&M)[NW], int lo, int hi)
{
    uint64_t acc = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) acc |= M[w] & range_word(lo, hi, w);
    return acc != 0;
}

template <int NW>
__device__ __forceinline__ int lowest_set(const uint64_t (&M)[NW])
{
#pragma unroll
    for (int w = 0; w < NW; w++)
        if (M[w]) return w * 64 + __ffsll((long long)M[w]) - 1;
    return -1;
}

template <int NW>
__device__ __forceinline__ int highest_set(const uint64_t (&M)[NW])
{
#pragma unroll
    for (int w = NW - 1; w >= 0; w--)
        if (M[w]) return w * 64 + 63 - __clzll((long long)M[w]);
    return -1;
}

template <int NW>
__device__ __forceinline__ void clear_bit(uint64_t (&M)[NW], int b)
{
#pragma unroll
    for (int w = 0; w < NW; w++)
        if ((b >> 6) == w) M[w] &= ~(1ull << (b & 63));
}

__device__ __forceinline__ uint64_t mix64(uint64_t h, uint64_t v)
{
    h ^= v;
    h *= 0x9E3779B97F4A7C15ull;
    h ^= h >> 32;
    return h;
}
__device__ __forceinline__ uint64_t fin64(uint64_t h)
{
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 29;
    return h;
}

// 64 bits starting at bit `off` of a word array (caller guarantees idx+1 is readable)
__device__ __forceinline__ uint64_t read64(const uint64_t *w, int off)
{
    const int idx = off >> 6, sh = off & 63;
    const uint64_t lo = w[idx];
    const uint64_t hi = w[idx + 1];
    return sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
}

// same, bounded word array of n words (words past the end read as zero)
__device__ __forceinline__ uint64_t read64_bounded(const uint64_t *w, int n, int off)
{
    const int idx = off >> 6, sh = off & 63;
    const uint64_t lo = idx < n ? __ldg(w + idx) : 0ull;
    const uint64_t hi = (idx + 1) < n ? __ldg(w + idx + 1) : 0ull;
    return sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
}

// Hash of a k-mer given a reader of 64-bit fields; identical for probes and targets.
template <typename Reader>
__device__ __forceinline__ uint64_t kmer_hash(int bits, int k, Reader rd)
{
    uint64_t h = 0x243F6A8885A308D3ull;
    for (int b = 0; b < bits; b++) {
        for (int c = 0; c * 64 < k; c++) {
            uint64_t v = rd(b, c * 64);
            const int nb = k - c * 64;
            if (nb < 64) v &= (1ull << nb) - 1ull;
            h = mix64(h, v);
        }
    }
    return fin64(h);
}

// ---------------------------------------------------------------------------------------
// K2: seed index.  Entry = probe << 32 | pos << 24 | tag24.
// ---------------------------------------------------------------------------------------
__global__ void seed_expand_kernel(const uint32_t *__restrict__ seed_off, int64_t n_probes,
                                   uint32_t *__restrict__ entry_probe)
{
    // one warp per probe writes the probe id of each of its seed entries
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < n_probes; p += n_warps)
        for (uint32_t e = seed_off[p] + lane; e < seed_off[p + 1]; e += 32) entry_probe[e] = (uint32_t)p;
}

template <bool SCATTER>
__global__ void seed_index_kernel(const uint32_t *__restrict__ entry_probe,
                                  const uint8_t *__restrict__ seed_pos, int64_t n_entries,
                                  const uint64_t *__restrict__ pwords, int bits, int nw, int k,
                                  uint32_t bucket_mask, uint32_t *__restrict__ bucket_count,
                                  const int64_t *__restrict__ bucket_off, uint32_t *__restrict__ cursor,
                                  uint64_t *__restrict__ entries)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_entries;
         e += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t p = entry_probe[e];
        const int pos = seed_pos[e];
        const uint64_t *pw = pwords + (int64_t)p * bits * nw;
        const uint64_t h = kmer_hash(bits, k, [&](int b, int o) { return read64_bounded(pw + b * nw, nw, pos + o); });
        const uint32_t bucket = (uint32_t)h & bucket_mask;
        if (!SCATTER) {
            atomicAdd(&bucket_count[bucket], 1u);
        } else {
            const uint32_t slot = atomicAdd(&cursor[bucket], 1u);
            entries[bucket_off[bucket] + slot] =
                ((uint64_t)p << 32) | ((uint64_t)pos << 24) | (uint64_t)(h >> 40);
        }
    }
}

// ---------------------------------------------------------------------------------------
// K3: scan
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine; SASS: UBLKCP), completion on an mbarrier
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// utils/longest_common_substring.py:59-159 on bit masks.  M: mismatch mask of the alignment,
// valid for probe positions [a, b); anchor [s, s+k) is mismatch free.  Returns the length and
// writes the start (probe coordinate); bef0/aft0 give the exact-match extents for the island test.
template <int NW>
__device__ __forceinline__ int anchored_extend(const uint64_t (&M)[NW], int a, int b, int s, int k, int m,
                                               int &start, int &exact_len)
{
    uint64_t ML[NW], MR[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) {
        ML[w] = M[w] & range_word(a, s, w);
        MR[w] = M[w] & range_word(s + k, b, w);
    }
    // after[j] = distance from the anchor's end to the (j+1)-th mismatch on the right (:125),
    // packed one byte each; entries past the last mismatch mean "to the end of the alignment" (:147-150)
    uint64_t aft[4] = {0, 0, 0, 0};
    const int after_full = b - (s + k);
    int n_right = 0;
    for (int j = 0; j <= m; j++) {
        const int pos = lowest_set(MR);
        if (pos < 0) break;
        const uint64_t v = (uint64_t)(pos - (s + k));
#pragma unroll
        for (int q = 0; q < 4; q++)
            if ((j >> 3) == q) aft[q] |= v << ((j & 7) * 8);
        clear_bit(MR, pos);
        n_right++;
    }
    auto after = [&](int j) -> int {
        if (j >= n_right) return after_full;
        uint64_t word = aft[0];
#pragma unroll
        for (int q = 1; q < 4; q++)
            if ((j >> 3) == q) word = aft[q];
        return (int)((word >> ((j & 7) * 8)) & 0xffull);
    };
    int best_len = -1, best_start = -1;
    int bef0 = 0;
    for (int i = 0; i <= m; i++) {
        // before[i]: distance to the (i+1)-th mismatch on the left, else everything to `a` (:140-146)
        const int hp = highest_set(ML);
        const int bef = hp >= 0 ? (s - 1 - hp) : (s - a);
        if (i == 0) bef0 = bef;
        const int tot = bef + k + after(m - i);
        if (tot > best_len) { best_len = tot; best_start = s - bef; }     // strict '>' (:154)
        if (hp < 0) break;       // further i: same `before`, `after` can only shrink
        clear_bit(ML, hp);
    }
    start = best_start;
    exact_len = bef0 + k + after(0);
    return best_len;
}

template <int NW, bool EMIT>
__device__ __forceinline__ void process_hit(const ScanParams &P, const uint64_t *s_tile, int64_t t0, int64_t g,
                                            int64_t qs, int64_t qe, uint32_t q_ubase, uint32_t p, int pos)
{
    const int L = P.plen[p];
    const int k = P.k;
    const int64_t d = g - pos;                       // target coordinate of probe position 0
    const int off = (int)(d - t0) + CB_FRONT_PAD;    // bit offset in the staged tile
    uint64_t M[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) M[w] = 0ull;
    const uint64_t *pw = P.pwords + (int64_t)p * P.bits * NW;
    for (int b = 0; b < P.bits; b++) {
#pragma unroll
        for (int w = 0; w < NW; w++) M[w] |= __ldg(pw + b * NW + w) ^ read64(s_tile + b * TW, off + 64 * w);
    }
    // probe.py:1075-1094: the alignment is clipped to the sequence on both sides
    const int a = (int)max((int64_t)0, qs - d);
    const int bnd = (int)min((int64_t)L, qe - d);
    if (any_in_range<NW>(M, pos, pos + k)) return;   // bucket/tag collision: k-mer differs

    // ownership: the smallest in-bounds, exactly matching seed of this probe on this diagonal
    uint32_t e = P.seed_off[p];
    const uint32_t se = P.seed_off[p + 1];
    for (; e < se; e++) {
        const int s = P.seed_pos[e];
        if (s >= pos) break;
        if (s >= a && s + k <= bnd && !any_in_range<NW>(M, s, s + k)) return;
    }
    const int64_t qlen = qe - qs;
    int thres = P.lcf;                                // probe.py:1332
    if (L < thres) thres = L;
    if (qlen < (int64_t)thres) thres = (int)qlen;

    uint32_t cur_s = 0, cur_e = 0;
    bool have = false;
    uint32_t n_out = 0;
    auto flush = [&]() {
        if (EMIT) {
            const uint32_t slot = atomicAdd(&P.rec_cursor[p], 1u);
            P.rec[P.rec_off[p] + slot] = ((uint64_t)cur_s << 32) | (uint64_t)cur_e;
        } else {
            n_out++;
        }
    };
    for (; e < se; e++) {
        const int s = P.seed_pos[e];
        if (!(s >= a && s + k <= bnd)) continue;
        if (any_in_range<NW>(M, s, s + k)) continue;
        int start, exact_len;
        const int len = anchored_extend<NW>(M, a, bnd, s, k, P.m, start, exact_len);
        if (len < thres) continue;
        if (P.island > 0) {                           // probe.py:1335-1342
            const int ex = (P.m == 0) ? len : exact_len;
            if (ex < P.island) continue;
        }
        // sequence-local range, then +-cover_extension, clip, universe offset
        // (filter/set_cover_filter.py:429-439)
        int64_t rs = d + start - qs, re = rs + len;
        rs -= P.ext;
        re += P.ext;
        if (rs < 0) rs = 0;
        if (re > qlen) re = qlen;
        const uint32_t us = q_ubase + (uint32_t)rs, ue = q_ubase + (uint32_t)re;
        if (have && us <= cur_e && ue >= cur_s) {     // overlaps or touches the pending range
            cur_s = min(cur_s, us);
            cur_e = max(cur_e, ue);
        } else {
            if (have) flush();
            cur_s = us;
            cur_e = ue;
            have = true;
        }
    }
    if (have) flush();
    if (!EMIT && n_out) atomicAdd(&P.rec_count[p], n_out);
}

template <int NW, bool EMIT>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_kernel(const ScanParams P)
{
    __shared__ __align__(128) uint64_t s_tile[CB_MAX_SYMBOL_BITS * TW];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_begin_lo[CB_TILE];       // bucket begin (entries < 2^32)
    __shared__ uint32_t s_cum[CB_TILE + 1];
    __shared__ uint32_t s_tag[CB_TILE];
    __shared__ uint32_t s_seq[CB_TILE];
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    __shared__ long long s_tile_id;
    __shared__ long long s_qrange[2];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phase = 0;
    unsigned long long local_hits = 0, local_lookups = 0;

    for (;;) {
        if (tid == 0) s_tile_id = (long long)atomicAdd(P.tile_counter, 1ull);
        __syncthreads();
        const int64_t tile = s_tile_id;
        if (tile >= P.n_tiles) break;
        const int64_t t0 = tile * CB_TILE;

        // ---- stage the target tile (with a FRONT_PAD halo on both sides) through the TMA engine
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&s_bar, (uint32_t)(P.bits * TW * 8));
            for (int b = 0; b < P.bits; b++)
                tma_bulk_g2s(s_tile + b * TW, P.planes + (int64_t)b * P.plane_words + (t0 >> 6), TW * 8, &s_bar);
        }
        // sequence range of the tile (two binary searches)
        if (tid < 2) {
            int64_t g = tid == 0 ? t0 : min(t0 + CB_TILE - 1, P.total_bases - 1);
            int64_t lo = 0, hi = P.n_seqs;          // largest q with seq_start[q] <= g
            while (hi - lo > 1) {
                const int64_t mid = (lo + hi) >> 1;
                if (P.seq_start[mid] <= g) lo = mid; else hi = mid;
            }
            s_qrange[tid] = lo;
        }
        mbar_wait(&s_bar, phase);
        phase ^= 1;
        __syncthreads();

        // ---- phase 1: one seed-index lookup per target position
        uint32_t cnt[POS_PER_THREAD];
        uint32_t tsum = 0;
#pragma unroll
        for (int r = 0; r < POS_PER_THREAD; r++) {
            const int j = tid * POS_PER_THREAD + r;
            const int64_t g = t0 + j;
            cnt[r] = 0;
            if (g < P.total_bases) {
                int64_t lo = s_qrange[0], hi = s_qrange[1] + 1;
                while (hi - lo > 1) {
                    const int64_t mid = (lo + hi) >> 1;
                    if (P.seq_start[mid] <= g) lo = mid; else hi = mid;
                }
                const int64_t qe = P.seq_start[lo + 1];
                s_seq[j] = (uint32_t)lo;
                if (g + P.k <= qe) {                 // probe.py:1062 i in [0, len-k]
                    const uint64_t h = kmer_hash(P.bits, P.k, [&](int b, int o) {
                        return read64(s_tile + b * TW, j + CB_FRONT_PAD + o);
                    });
                    const uint32_t bucket = (uint32_t)h & P.bucket_mask;
                    const int64_t b0 = P.bucket_off[bucket], b1 = P.bucket_off[bucket + 1];
                    s_begin_lo[j] = (uint32_t)b0;
                    s_tag[j] = (uint32_t)(h >> 40);
                    cnt[r] = (uint32_t)(b1 - b0);
                    local_lookups++;
                }
            }
            tsum += cnt[r];
        }
        // block exclusive scan of the per-position hit counts
        uint32_t inc = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = lane < SCAN_THREADS / 32 ? s_warp[lane] : 0u;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            if (lane < SCAN_THREADS / 32) s_warp[lane] = w;
        }
        __syncthreads();
        uint32_t run = inc - tsum + (warp ? s_warp[warp - 1] : 0u);
#pragma unroll
        for (int r = 0; r < POS_PER_THREAD; r++) {
            s_cum[tid * POS_PER_THREAD + r] = run;
            run += cnt[r];
        }
        if (tid == SCAN_THREADS - 1) s_cum[CB_TILE] = run;
        __syncthreads();
        const uint32_t total = s_cum[CB_TILE];
        local_hits += (tid == 0) ? total : 0;

        // ---- phase 2: candidate hits, spread evenly over the block
        for (uint32_t h = tid; h < total; h += SCAN_THREADS) {
            int lo = 0, hi = CB_TILE;                // largest j with s_cum[j] <= h
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_cum[mid] <= h) lo = mid; else hi = mid;
            }
            const int j = lo;
            const uint64_t ent = __ldg(P.entries + (int64_t)s_begin_lo[j] + (h - s_cum[j]));
            if ((uint32_t)(ent & 0xffffffull) != s_tag[j]) continue;
            const uint32_t q = s_seq[j];
            process_hit<NW, EMIT>(P, s_tile, t0, t0 + j, P.seq_start[q], P.seq_start[q + 1], P.seq_ubase[q],
                                  (uint32_t)(ent >> 32), (int)((ent >> 24) & 0xff));
        }
        __syncthreads();                              // tile + scan arrays are reused by the next tile
    }
    if (!EMIT) {
        if (local_hits) atomicAdd(P.stat_hits, local_hits);
        if (local_lookups) atomicAdd(P.stat_lookups, local_lookups);
    }
}

// ---------------------------------------------------------------------------------------
// K4: per-probe sort + merge of the emitted ranges.
// One block per probe.  Normalised bitonic network (every compare-exchange orders min -> lower
// index, so virtual +inf padding to a power of two needs no storage) in shared memory when the
// probe's ranges fit, otherwise in place in global memory.  Then warp 0 merges overlapping or
// touching ranges (utils/interval.py:304-314: start <= curr_end) in 32-wide chunks.
// ---------------------------------------------------------------------------------------
constexpr int MERGE_THREADS = 128;
constexpr int MERGE_SMEM_CAP = 4096;

__device__ __forceinline__ void bitonic_sort(uint64_t *a, uint32_t n)
{
    uint32_t n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (uint32_t size = 2; size <= n2; size <<= 1) {
        // first step of each merge: compare i with its mirror inside the block of `size`
        for (uint32_t t = threadIdx.x; t < n2 / 2; t += blockDim.x) {
            const uint32_t blk = t / (size / 2), o = t % (size / 2);
            const uint32_t i = blk * size + o, j = blk * size + size - 1 - o;
            if (j < n) {
                const uint64_t x = a[i], y = a[j];
                if (x > y) { a[i] = y; a[j] = x; }
            }
        }
        __syncthreads();
        for (uint32_t stride = size / 4; stride >= 1; stride >>= 1) {
            for (uint32_t t = threadIdx.x; t < n2 / 2; t += blockDim.x) {
                const uint32_t i = (t / stride) * stride * 2 + (t % stride), j = i + stride;
                if (j < n) {
                    const uint64_t x = a[i], y = a[j];
                    if (x > y) { a[i] = y; a[j] = x; }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(MERGE_THREADS)
merge_kernel(const int64_t *__restrict__ rec_off, uint64_t *__restrict__ rec, int64_t n_probes,
             uint32_t *__restrict__ n_merged, uint32_t *__restrict__ max_len)
{
    __shared__ uint64_t s_rec[MERGE_SMEM_CAP];
    uint32_t local_max = 0;
    for (int64_t p = blockIdx.x; p < n_probes; p += gridDim.x) {
        const int64_t o0 = rec_off[p];
        const uint32_t n = (uint32_t)(rec_off[p + 1] - o0);
        if (n == 0) {
            if (threadIdx.x == 0) n_merged[p] = 0;
            continue;
        }
        uint64_t *g = rec + o0;
        uint64_t *a = g;
        const bool in_smem = n <= MERGE_SMEM_CAP;
        if (in_smem) {
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s_rec[i] = g[i];
            a = s_rec;
        }
        __syncthreads();
        if (n > 1) bitonic_sort(a, n);
        // merge by warp 0; output is written in place at the front of the probe's slice
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            uint32_t carry_max = 0, carry_gs = 0, n_out = 0;
            for (uint32_t base = 0; base < n; base += 32) {
                const uint32_t idx = base + lane;
                const bool valid = idx < n;
                const uint64_t r = valid ? a[idx] : ~0ull;
                const uint32_t s = (uint32_t)(r >> 32), e = valid ? (uint32_t)r : 0u;
                const uint64_t rn = (idx + 1 < n) ? a[idx + 1] : ~0ull;     // next start (lookahead)
                const uint32_t s_next = (uint32_t)(rn >> 32);
                // exclusive running max of the ends
                uint32_t inc = e;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc = max(inc, t);
                }
                inc = max(inc, carry_max);
                uint32_t exc = __shfl_up_sync(0xffffffffu, inc, 1);
                if (lane == 0) exc = carry_max;
                const bool head = valid && (idx == 0 || s > exc);
                // start of the group each element belongs to
                uint32_t gs = head ? s : 0u;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, gs, o);
                    if (lane >= o) gs = max(gs, t);
                }
                gs = max(gs, carry_gs);
                const bool tail = valid && (idx + 1 == n || s_next > inc);
                const unsigned tails = __ballot_sync(0xffffffffu, tail);
                __syncwarp();
                if (tail) {
                    const uint32_t k = n_out + __popc(tails & ((1u << lane) - 1u));
                    // k <= idx, and every element at index <= base+31 is already in registers
                    g[k] = ((uint64_t)gs << 32) | (uint64_t)inc;
                    local_max = max(local_max, inc - gs);
                }
                n_out += __popc(tails);
                carry_max = __shfl_sync(0xffffffffu, inc, 31);
                carry_gs = __shfl_sync(0xffffffffu, gs, 31);
                __syncwarp();
            }
            if (lane == 0) n_merged[p] = n_out;
        }
        __syncthreads();
    }
    if (threadIdx.x < 32) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
        if (threadIdx.x == 0 && local_max) atomicMax(max_len, local_max);
    }
}

__global__ void compact_kernel(const int64_t *__restrict__ rec_off, const uint64_t *__restrict__ rec,
                               const int64_t *__restrict__ iv_off, int64_t n_probes, uint2 *__restrict__ iv)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < n_probes; p += n_warps) {
        const int64_t src = rec_off[p], dst = iv_off[p];
        const int64_t n = iv_off[p + 1] - dst;
        for (int64_t i = lane; i < n; i += 32) {
            const uint64_t r = rec[src + i];
            iv[dst + i] = make_uint2((uint32_t)(r >> 32), (uint32_t)r);
        }
    }
}

template <int NW>
int launch_scan(cb_ctx *ctx, const ScanParams &P, bool emit, int grid)
{
    if (emit) scan_kernel<NW, true><<<grid, SCAN_THREADS, 0, ctx->stream>>>(P);
    else scan_kernel<NW, false><<<grid, SCAN_THREADS, 0, ctx->stream>>>(P);
    ctx->launches++;
    CB_CUDA(ctx, cudaGetLastError());
    return CB_OK;
}

int launch_scan_nw(cb_ctx *ctx, int nw, const ScanParams &P, bool emit, int grid)
{
    switch (nw) {
    case 1: return launch_scan<1>(ctx, P, emit, grid);
    case 2: return launch_scan<2>(ctx, P, emit, grid);
    case 3: return launch_scan<3>(ctx, P, emit, grid);
    case 4: return launch_scan<4>(ctx, P, emit, grid);
    }
    return cb_fail(ctx, CB_ERR_UNSUPPORTED, "probe longer than CB_MAX_PROBE_LEN");
}

}  // namespace

int cb_coverage_impl(cb_ctx *ctx, const cb_probes *probes, const cb_targets *targets,
                     const cb_hyb_params *hp, const int64_t *seed_off, const int32_t *seed_pos,
                     cb_cover **out, cb_stats *stats)
{
    if (!probes || !targets || !hp || !out) return cb_fail(ctx, CB_ERR_ARG, "null argument");
    if (probes->bits != targets->bits || memcmp(probes->lut, targets->lut, 256) != 0)
        return cb_fail(ctx, CB_ERR_ARG, "probes and targets were packed with different alphabets");
    if (hp->mismatches < 0 || hp->mismatches > CB_MAX_MISMATCHES)
        return cb_fail(ctx, CB_ERR_UNSUPPORTED, "mismatches outside [0, CB_MAX_MISMATCHES]");
    if (hp->k < 1 || hp->k > CB_MAX_PROBE_LEN) return cb_fail(ctx, CB_ERR_ARG, "seed length k out of range");
    if (hp->cover_extension < 0 || hp->lcf_thres < 0 || hp->island_of_exact_match < 0)
        return cb_fail(ctx, CB_ERR_ARG, "negative hybridisation parameter");
    const int64_t P = probes->n_probes;
    cudaStream_t st = ctx->stream;
    EventTimer t_all(st), t_idx(st), t_cnt(st), t_emit(st), t_merge(st);
    t_all.start();

    cb_cover *cov = new cb_cover();
    cov->ctx = ctx;
    cov->n_probes = P;
    cov->n_genomes = targets->n_genomes;
    cov->universe_bits = targets->universe_bits;
    cov->h_ubase = targets->h_ubase;
    cov->h_genome_len = targets->h_genome_len;
    struct Guard { cb_cover *c; ~Guard() { if (c) cb_cover_free(c); } } guard{cov};
    CB_CUDA(ctx, cudaMalloc((void **)&cov->d_ubase, sizeof(uint32_t) * (size_t)(targets->n_genomes + 1)));
    CB_CUDA(ctx, cudaMemcpyAsync(cov->d_ubase, targets->d_ubase, sizeof(uint32_t) * (size_t)(targets->n_genomes + 1),
                                 cudaMemcpyDeviceToDevice, st));
    CB_CUDA(ctx, cudaMalloc((void **)&cov->d_iv_off, sizeof(int64_t) * (size_t)(P + 1)));

    const int64_t n_entries = P ? seed_off[P] : 0;
    bool empty = (P == 0 || targets->total_bases == 0 || n_entries == 0);
    if (!empty) {
        // validate + narrow the seed CSR on the host
        if (n_entries >= (int64_t)0xffffffffll) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "too many seed entries");
    }
    if (empty) {
        CB_CUDA(ctx, cudaMemsetAsync(cov->d_iv_off, 0, sizeof(int64_t) * (size_t)(P + 1), st));
        CB_CUDA(ctx, cudaStreamSynchronize(st));
        if (stats) memset(stats, 0, sizeof *stats);
        guard.c = nullptr;
        *out = cov;
        return CB_OK;
    }

    std::vector<uint32_t> h_soff((size_t)P + 1);
    std::vector<uint8_t> h_spos((size_t)n_entries);
    for (int64_t p = 0; p <= P; p++) h_soff[(size_t)p] = (uint32_t)seed_off[p];
    for (int64_t p = 0; p < P; p++) {
        if (seed_off[p + 1] < seed_off[p]) return cb_fail(ctx, CB_ERR_ARG, "seed_off not monotone");
        int prev = -1;
        for (int64_t e = seed_off[p]; e < seed_off[p + 1]; e++) {
            const int s = seed_pos[e];
            if (s <= prev) return cb_fail(ctx, CB_ERR_ARG, "seed positions must be distinct and ascending per probe");
            if (s < 0 || s > 255) return cb_fail(ctx, CB_ERR_ARG, "seed position out of range");
            h_spos[(size_t)e] = (uint8_t)s;
            prev = s;
        }
    }
    DevBuf<uint32_t> d_soff, d_eprobe, d_bcount, d_bcursor;
    DevBuf<uint8_t> d_spos;
    DevBuf<int64_t> d_boff;
    DevBuf<uint64_t> d_entries;
    DevBuf<unsigned long long> d_ctr;        // [0] tile counter, [1] hits, [2] lookups
    CB_CUDA(ctx, d_soff.alloc((size_t)P + 1));
    CB_CUDA(ctx, d_spos.alloc((size_t)n_entries));
    CB_CUDA(ctx, d_eprobe.alloc((size_t)n_entries));
    CB_CUDA(ctx, d_entries.alloc((size_t)n_entries));
    CB_CUDA(ctx, d_ctr.alloc(4));
    CB_CUDA(ctx, cudaMemcpyAsync(d_soff.p, h_soff.data(), sizeof(uint32_t) * (size_t)(P + 1), cudaMemcpyHostToDevice, st));
    CB_CUDA(ctx, cudaMemcpyAsync(d_spos.p, h_spos.data(), (size_t)n_entries, cudaMemcpyHostToDevice, st));

    // ---- K2 seed index
    t_idx.start();
    int64_t nb = 1024;
    while (nb < 2 * n_entries) nb <<= 1;
    CB_CUDA(ctx, d_bcount.alloc((size_t)nb));
    CB_CUDA(ctx, d_bcursor.alloc((size_t)nb));
    CB_CUDA(ctx, d_boff.alloc((size_t)nb + 1));
    CB_CUDA(ctx, cudaMemsetAsync(d_bcount.p, 0, sizeof(uint32_t) * (size_t)nb, st));
    CB_CUDA(ctx, cudaMemsetAsync(d_bcursor.p, 0, sizeof(uint32_t) * (size_t)nb, st));
    const int wide = ctx->sm_count * 8;
    seed_expand_kernel<<<wide, 256, 0, st>>>(d_soff.p, P, d_eprobe.p);
    seed_index_kernel<false><<<wide, 256, 0, st>>>(d_eprobe.p, d_spos.p, n_entries, probes->d_words, probes->bits,
                                                   probes->nw, hp->k, (uint32_t)(nb - 1), d_bcount.p, nullptr,
                                                   nullptr, nullptr);
    ctx->launches += 2;
    CB_CUDA(ctx, cudaGetLastError());
    CB_TRY(cb_exclusive_scan_u32_to_i64(ctx, d_bcount.p, d_boff.p, nb, nullptr));
    seed_index_kernel<true><<<wide, 256, 0, st>>>(d_eprobe.p, d_spos.p, n_entries, probes->d_words, probes->bits,
                                                  probes->nw, hp->k, (uint32_t)(nb - 1), nullptr, d_boff.p,
                                                  d_bcursor.p, d_entries.p);
    ctx->launches++;
    CB_CUDA(ctx, cudaGetLastError());
    t_idx.stop();

    // ---- K3 count pass
    DevBuf<uint32_t> d_rcount, d_rcursor;
    DevBuf<int64_t> d_roff;
    CB_CUDA(ctx, d_rcount.alloc((size_t)P));
    CB_CUDA(ctx, d_rcursor.alloc((size_t)P));
    CB_CUDA(ctx, d_roff.alloc((size_t)P + 1));
    CB_CUDA(ctx, cudaMemsetAsync(d_rcount.p, 0, sizeof(uint32_t) * (size_t)P, st));
    CB_CUDA(ctx, cudaMemsetAsync(d_rcursor.p, 0, sizeof(uint32_t) * (size_t)P, st));
    CB_CUDA(ctx, cudaMemsetAsync(d_ctr.p, 0, sizeof(unsigned long long) * 4, st));

    ScanParams sp;
    sp.planes = targets->d_planes;
    sp.plane_words = targets->plane_words;
    sp.bits = targets->bits;
    sp.total_bases = targets->total_bases;
    sp.n_seqs = targets->n_seqs;
    sp.seq_start = targets->d_seq_start;
    sp.seq_ubase = targets->d_seq_ubase;
    sp.pwords = probes->d_words;
    sp.plen = probes->d_len;
    sp.seed_off = d_soff.p;
    sp.seed_pos = d_spos.p;
    sp.bucket_off = d_boff.p;
    sp.entries = d_entries.p;
    sp.bucket_mask = (uint32_t)(nb - 1);
    sp.m = hp->mismatches;
    sp.lcf = hp->lcf_thres;
    sp.island = hp->island_of_exact_match;
    sp.ext = hp->cover_extension;
    sp.k = hp->k;
    sp.rec_count = d_rcount.p;
    sp.rec_off = nullptr;
    sp.rec_cursor = d_rcursor.p;
    sp.rec = nullptr;
    sp.n_tiles = (targets->total_bases + CB_TILE - 1) / CB_TILE;
    sp.tile_counter = d_ctr.p;
    sp.stat_hits = d_ctr.p + 1;
    sp.stat_lookups = d_ctr.p + 2;
    int64_t grid64 = sp.n_tiles < (int64_t)ctx->sm_count * 4 ? sp.n_tiles : (int64_t)ctx->sm_count * 4;
    const int grid = (int)grid64;

    t_cnt.start();
    CB_TRY(launch_scan_nw(ctx, probes->nw, sp, false, grid));
    t_cnt.stop();
    int64_t n_raw = 0;
    CB_TRY(cb_exclusive_scan_u32_to_i64(ctx, d_rcount.p, d_roff.p, P, &n_raw));
    if (n_raw >= (int64_t)0x7fffffff00ll) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "too many cover ranges");

    // ---- K3 emit pass
    DevBuf<uint64_t> d_rec;
    CB_CUDA(ctx, d_rec.alloc((size_t)n_raw));
    CB_CUDA(ctx, cudaMemsetAsync(d_ctr.p, 0, sizeof(unsigned long long), st));    // tile counter only
    sp.rec_off = d_roff.p;
    sp.rec = d_rec.p;
    t_emit.start();
    CB_TRY(launch_scan_nw(ctx, probes->nw, sp, true, grid));
    t_emit.stop();

    // ---- K4 merge
    DevBuf<uint32_t> d_nmerged, d_maxlen;
    CB_CUDA(ctx, d_nmerged.alloc((size_t)P));
    CB_CUDA(ctx, d_maxlen.alloc(1));
    CB_CUDA(ctx, cudaMemsetAsync(d_maxlen.p, 0, sizeof(uint32_t), st));
    t_merge.start();
    {
        int64_t g = P < (int64_t)ctx->sm_count * 16 ? P : (int64_t)ctx->sm_count * 16;
        merge_kernel<<<(unsigned)g, MERGE_THREADS, 0, st>>>(d_roff.p, d_rec.p, P, d_nmerged.p, d_maxlen.p);
        ctx->launches++;
        CB_CUDA(ctx, cudaGetLastError());
    }
    int64_t n_iv = 0;
    CB_TRY(cb_exclusive_scan_u32_to_i64(ctx, d_nmerged.p, cov->d_iv_off, P, &n_iv));
    CB_CUDA(ctx, cudaMalloc((void **)&cov->d_iv, sizeof(uint2) * (size_t)(n_iv ? n_iv : 1)));
    compact_kernel<<<wide, 256, 0, st>>>(d_roff.p, d_rec.p, cov->d_iv_off, P, cov->d_iv);
    ctx->launches++;
    CB_CUDA(ctx, cudaGetLastError());
    t_merge.stop();
    t_all.stop();
    cov->n_intervals = n_iv;

    unsigned long long h_ctr[4];
    CB_CUDA(ctx, cudaMemcpyAsync(h_ctr, d_ctr.p, sizeof h_ctr, cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaMemcpyAsync(&cov->max_interval_len, d_maxlen.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    if (stats) {
        stats->ms_seed_index = t_idx.ms();
        stats->ms_scan_count = t_cnt.ms();
        stats->ms_scan_emit = t_emit.ms();
        stats->ms_merge = t_merge.ms();
        stats->ms_total = t_all.ms();
        stats->n_seed_entries = n_entries;
        stats->n_seed_lookups = (int64_t)h_ctr[2];
        stats->n_candidate_hits = (int64_t)h_ctr[1];
        stats->n_raw_ranges = n_raw;
        stats->n_intervals = n_iv;
        stats->n_kernel_launches = ctx->launches;
    }
    guard.c = nullptr;
    *out = cov;
    return CB_OK;
}
