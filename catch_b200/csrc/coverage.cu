// Stage A: probe coverage of the target genomes (K2 seed index, K3 scan, K4 merge).
//
// Replaces (reference paths relative to catch/):
//   probe.py:356-504,684-763   seed-map construction  -> seed_index_* kernels (bucketed hash CSR)
//   probe.py:1008-1119         _find_probe_covers_in_subsequence -> scan_kernel
//   utils/longest_common_substring.py:59-159 k_lcf_around_anchor -> anchored_extend (bit scans)
//   probe.py:1328-1344         lcf() predicate -> inside run_owner_task
//   utils/interval.py:288-316  merge_overlapping, filter/set_cover_filter.py:429-439,462-466
//                              -> emitted ranges are extended/clipped/offset at emit time and
//                                 merged per probe by merge_kernel
//
// Semantics kept bit for bit: a (probe, diagonal) pair is examined iff one of the probe's
// SELECTED seeds matches the target exactly at an in-bounds position; every matching seed gives
// its own anchored range; ranges are unioned.
//
// How the scan is organised (one CTA per tile of CB_TILE target positions, persistent grid):
//   1. the tile's bit planes (+ halo) are staged into shared memory by the TMA engine
//      (cp.async.bulk + mbarrier);
//   2. one seed-index lookup per target position gives a bucket of candidate (probe, seed) hits;
//      a block-wide prefix sum spreads the hits evenly over the threads;
//   3. per hit, a CHEAP test decides whether the hit has to be evaluated at all.  Seeds inside one
//      mismatch-free run of a diagonal see the same mismatches on both sides and therefore give the
//      same range, so one seed per run is enough.  Every index entry carries the distance g to the
//      probe's nearest lower selected seed and the g probe bases in between; if those bases also
//      match the target (and the lower seed starts inside the sequence), the lower seed matches on
//      this diagonal, lies in the same run and yields the same range: the hit is dropped after one
//      16-byte entry load and a masked compare against the staged tile -- the probe record is never
//      touched.  What survives is exactly one hit per (probe, diagonal, run that holds a seed);
//   4. survivors are compacted into a shared-memory queue and processed densely: mismatch mask
//      M = OR over planes of (probe XOR window), one anchored extension around the hit's own seed
//      (ffs/clz walks over at most 2(m+1) mismatches), threshold/island tests, +-e, clip, universe
//      offset, and a warp-aggregated append to a global range list that is later bucketed by probe.
#include <cstdlib>
#include <cstring>

#include "internal.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
#ifndef SCAN_MIN_BLOCKS
#define SCAN_MIN_BLOCKS 4
#endif
constexpr int POS_PER_THREAD = CB_TILE / SCAN_THREADS;
constexpr int TW = CB_TILE_WORDS;
constexpr int QUEUE_FLUSH = 512;          // the survivor queue is drained once it holds more than this

struct ScanParams {
    // targets
    const uint64_t *planes;
    int64_t plane_words;
    int bits;
    int64_t total_bases;
    int64_t n_seqs;
    const int64_t *seq_start;
    const uint32_t *seq_ubase;
    // probes: record = bits*NW plane words followed by NW words of seed mask
    const uint64_t *precs;
    int prec_words;
    const int32_t *plen;
    // seed index
    const int64_t *bucket_off;
    const ulonglong2 *entries;     // x: probe << 32 | pos << 24 | tag24; y: g | prefix fields (see seed_index_kernel)
    int pw;                        // bases per prefix field = 56 / bits
    uint32_t bucket_mask;
    // hybridisation model
    int m, lcf, island, ext, k;
    // output: global range list (probe, start, end, -) + per-probe counts
    uint4 *rec;
    unsigned long long rec_cap;
    unsigned long long *rec_cursor;
    uint32_t *rec_count;
    int keep_hit;                   // records carry the hit position (cb_coverage_records) instead of the per-probe sequence number
    // scheduling / stats
    int64_t n_tiles;
    unsigned long long *tile_counter;
    unsigned long long *stat_hits;
    unsigned long long *stat_lookups;
    unsigned long long *stat_owners;
    int count_only;               // 1: only count candidate hits (capacity pre-pass)
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

template <int NW>
__device__ __forceinline__ bool test_bit(const uint64_t (&M)[NW], int b)
{
    uint64_t v = 0;
#pragma unroll
    for (int w = 0; w < NW; w++)
        if ((b >> 6) == w) v = M[w];
    return (v >> (b & 63)) & 1ull;
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
// Seed positions arrive as a CSR that may hold repeats in any order (the reference draws with
// replacement and keeps a set, probe.py:393-398).  One warp per probe ORs them into the probe's
// seed mask (the last NW words of its record) and counts the distinct positions.
// seed_off == nullptr: every probe has exactly `uniform` positions, probe p at [(p-lo)*uniform, (p-lo+1)*uniform).
// Probes outside [lo, hi) get an empty mask: they belong to another rank's shard.
__global__ void seed_mask_kernel(const int64_t *__restrict__ seed_off, int uniform,
                                 const uint8_t *__restrict__ seed_pos,
                                 const int32_t *__restrict__ plen, int k, int64_t n_probes, int64_t lo, int64_t hi,
                                 uint64_t *__restrict__ precs, int prec_words, int nw,
                                 uint32_t *__restrict__ n_distinct, int *__restrict__ bad)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < n_probes; p += n_warps) {
        uint64_t mask[4] = {0, 0, 0, 0};
        const int limit = plen[p] - k;             // last admissible seed start
        int64_t e0 = 0, e1 = 0;
        if (p >= lo && p < hi) {
            e0 = seed_off ? seed_off[p] : (p - lo) * uniform;
            e1 = seed_off ? seed_off[p + 1] : e0 + uniform;
        }
        for (int64_t e = e0 + lane; e < e1; e += 32) {
            const int s = seed_pos[e];
            if (s > limit) { *bad = 1; continue; }
#pragma unroll
            for (int w = 0; w < 4; w++)
                if ((s >> 6) == w) mask[w] |= 1ull << (s & 63);
        }
        uint32_t c = 0;
#pragma unroll
        for (int w = 0; w < 4; w++) {
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) mask[w] |= __shfl_xor_sync(0xffffffffu, mask[w], o);
            c += __popcll(mask[w]);
        }
#pragma unroll
        for (int w = 0; w < 4; w++)
            if (lane == w && w < nw) precs[p * (int64_t)prec_words + (prec_words - nw) + w] = mask[w];
        if (lane == 0) n_distinct[p] = c;
    }
}

// one thread per probe lists its distinct seed positions (ascending) as index entries
__global__ void seed_expand_kernel(const uint64_t *__restrict__ precs, int prec_words, int nw, int64_t n_probes,
                                   const int64_t *__restrict__ entry_off, uint32_t *__restrict__ entry_probe,
                                   uint8_t *__restrict__ entry_pos)
{
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_probes;
         p += (int64_t)gridDim.x * blockDim.x) {
        int64_t o = entry_off[p];
        for (int w = 0; w < nw; w++) {
            uint64_t m = precs[p * (int64_t)prec_words + (prec_words - nw) + w];
            while (m) {
                const int b = __ffsll((long long)m) - 1;
                m &= m - 1;
                entry_probe[o] = (uint32_t)p;
                entry_pos[o] = (uint8_t)(w * 64 + b);
                o++;
            }
        }
    }
}

// Entry layout (16 bytes):
//   x = probe << 32 | pos << 24 | tag24 (bits 40.. of the k-mer hash)
//   y = g | field_0 << 8 | field_1 << (8 + pw) | ...   with pw = 56 / bits bases per field:
//       g = distance from this seed down to the probe's nearest lower selected seed (0 when there is
//       none within min(pw, k) bases), field_b = bits [pos - g, pos) of probe plane b.  The scan uses it
//       to drop hits whose lower neighbour seed matches on the same diagonal (see scan_kernel).
template <bool SCATTER>
__global__ void seed_index_kernel(const uint32_t *__restrict__ entry_probe,
                                  const uint8_t *__restrict__ seed_pos, int64_t n_entries,
                                  const uint64_t *__restrict__ precs, int prec_words, int bits, int nw, int k,
                                  uint32_t bucket_mask, uint32_t *__restrict__ bucket_count,
                                  const int64_t *__restrict__ bucket_off, uint32_t *__restrict__ cursor,
                                  ulonglong2 *__restrict__ entries)
{
    const int pw = 56 / bits;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_entries;
         e += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t p = entry_probe[e];
        const int pos = seed_pos[e];
        const uint64_t *pwd = precs + (int64_t)p * prec_words;
        const uint64_t h = kmer_hash(bits, k, [&](int b, int o) { return read64_bounded(pwd + b * nw, nw, pos + o); });
        const uint32_t bucket = (uint32_t)h & bucket_mask;
        if (!SCATTER) {
            atomicAdd(&bucket_count[bucket], 1u);
        } else {
            // nearest selected seed below pos (seed mask = last nw words of the record)
            const uint64_t *sm = pwd + bits * nw;
            int prev = -1;
            {
                int w = pos >> 6;
                uint64_t m = sm[w] & ((1ull << (pos & 63)) - 1ull);
                for (;;) {
                    if (m) { prev = w * 64 + 63 - __clzll((long long)m); break; }
                    if (--w < 0) break;
                    m = sm[w];
                }
            }
            uint64_t y = 0ull;
            if (prev >= 0) {
                const int g = pos - prev;
                if (g <= pw && g <= k) {
                    y = (uint64_t)g;
                    for (int b = 0; b < bits; b++)
                        y |= (read64_bounded(pwd + b * nw, nw, prev) & ((1ull << g) - 1ull)) << (8 + b * pw);
                }
            }
            const uint32_t slot = atomicAdd(&cursor[bucket], 1u);
            ulonglong2 ent;
            ent.x = ((uint64_t)p << 32) | ((uint64_t)pos << 24) | (uint64_t)(h >> 40);
            ent.y = y;
            entries[bucket_off[bucket] + slot] = ent;
        }
    }
}

// copies the packed probe planes into the records (seed mask is filled by seed_expand_kernel)
__global__ void build_precs_kernel(const uint64_t *__restrict__ pwords, int64_t n_probes, int plane_words,
                                   int prec_words, uint64_t *__restrict__ precs)
{
    const int64_t n = n_probes * plane_words;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / plane_words;
        const int w = (int)(i % plane_words);
        precs[p * prec_words + w] = pwords[i];
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
// writes the start (probe coordinate); exact_len is the 0-mismatch extent for the island test.
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
        const int pos = lowest_set<NW>(MR);
        if (pos < 0) break;
        const uint64_t v = (uint64_t)(pos - (s + k));
#pragma unroll
        for (int q = 0; q < 4; q++)
            if ((j >> 3) == q) aft[q] |= v << ((j & 7) * 8);
        clear_bit<NW>(MR, pos);
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
        const int hp = highest_set<NW>(ML);
        const int bef = hp >= 0 ? (s - 1 - hp) : (s - a);
        if (i == 0) bef0 = bef;
        const int tot = bef + k + after(m - i);
        if (tot > best_len) { best_len = tot; best_start = s - bef; }     // strict '>' (:154)
        if (hp < 0) break;       // further i: same `before`, `after` can only shrink
        clear_bit<NW>(ML, hp);
    }
    start = best_start;
    exact_len = bef0 + k + after(0);
    return best_len;
}

// Mismatch mask of probe p aligned at target coordinate d (not yet clipped to the sequence).
template <int NW>
__device__ __forceinline__ void mismatch_mask(const ScanParams &P, const uint64_t *s_tile, int64_t t0, int64_t d,
                                              uint32_t p, uint64_t (&M)[NW])
{
    const int off = (int)(d - t0) + CB_FRONT_PAD;    // bit offset in the staged tile
    const uint64_t *pw = P.precs + (int64_t)p * P.prec_words;
#pragma unroll
    for (int w = 0; w < NW; w++) M[w] = 0ull;
    for (int b = 0; b < P.bits; b++) {
#pragma unroll
        for (int w = 0; w < NW; w++) M[w] |= __ldg(pw + b * NW + w) ^ read64(s_tile + b * TW, off + 64 * w);
    }
}

// ---------------------------------------------------------------------------------------
// Fast path for probes of up to 128 bases (NW == 2, every shipped probe length): masks are two
// scalar 64-bit registers.
// ---------------------------------------------------------------------------------------
struct U128 { uint64_t lo, hi; };

// bits [0, x) of a 128-bit mask, 0 <= x <= 128
__device__ __forceinline__ U128 below128(int x)
{
    U128 r;
    const int nl = x < 64 ? x : 64, nh = x > 64 ? x - 64 : 0;
    r.lo = nl == 64 ? ~0ull : ((1ull << nl) - 1ull);
    r.hi = nh == 64 ? ~0ull : ((1ull << nh) - 1ull);
    return r;
}

__device__ __forceinline__ U128 mismatch_mask_fast(const ScanParams &P, const uint64_t *s_tile, int64_t t0, int64_t d,
                                                   uint32_t p)
{
    const int off = (int)(d - t0) + CB_FRONT_PAD;
    const int idx = off >> 6, sh = off & 63;
    const ulonglong2 *pr = reinterpret_cast<const ulonglong2 *>(P.precs + (int64_t)p * P.prec_words);
    U128 M;
    M.lo = 0ull;
    M.hi = 0ull;
    for (int b = 0; b < P.bits; b++) {
        const ulonglong2 pw = __ldg(pr + b);
        const uint64_t *t = s_tile + b * TW + idx;
        const uint64_t s0 = t[0], s1 = t[1], s2 = t[2];
        // 64-bit funnel shifts; (x << (63 - sh)) << 1 is x << (64 - sh) without the sh == 0 hazard
        const uint64_t w0 = (s0 >> sh) | ((s1 << (63 - sh)) << 1);
        const uint64_t w1 = (s1 >> sh) | ((s2 << (63 - sh)) << 1);
        M.lo |= pw.x ^ w0;
        M.hi |= pw.y ^ w1;
    }
    return M;
}

// anchored extension (utils/longest_common_substring.py:59-159) on scalar 128-bit masks.
// ML / MR: mismatches strictly left of the anchor / at or right of its end, already clipped to [a, b).
__device__ __forceinline__ int anchored_extend_fast(U128 ML, U128 MR, int a, int b, int s, int k, int m,
                                                    int &start, int &exact_len)
{
    // after[j] packed one byte each (distances < 128)
    uint64_t aft[4] = {0, 0, 0, 0};
    const int after_full = b - (s + k);
    int n_right = 0;
    for (int j = 0; j <= m; j++) {
        // lowest mismatch right of the anchor; it is cleared with w & (w - 1) on the word that holds it
        const bool in_lo = MR.lo != 0ull;
        const uint64_t w = in_lo ? MR.lo : MR.hi;
        if (w == 0ull) break;
        const int pos = (in_lo ? 0 : 64) + __ffsll((long long)w) - 1;
        const uint64_t v = (uint64_t)(pos - (s + k));
        if (j < 8) aft[0] |= v << (j * 8);
        else {
#pragma unroll
            for (int q = 1; q < 4; q++)
                if ((j >> 3) == q) aft[q] |= v << ((j & 7) * 8);
        }
        if (in_lo) MR.lo = w & (w - 1ull); else MR.hi = w & (w - 1ull);
        n_right++;
    }
    auto after = [&](int j) -> int {
        if (j >= n_right) return after_full;
        uint64_t word = aft[0];
        if (j >= 8) {
#pragma unroll
            for (int q = 1; q < 4; q++)
                if ((j >> 3) == q) word = aft[q];
        }
        return (int)((word >> ((j & 7) * 8)) & 0xffull);
    };
    int best_len = -1, best_start = -1, bef0 = 0;
    for (int i = 0; i <= m; i++) {
        // highest mismatch left of the anchor
        const bool in_hi = ML.hi != 0ull;
        const uint64_t wl = in_hi ? ML.hi : ML.lo;
        const int lz = __clzll((long long)wl);                           // 64 when wl == 0
        const int hp = wl ? (in_hi ? 127 : 63) - lz : -1;
        const int bef = hp >= 0 ? (s - 1 - hp) : (s - a);
        if (i == 0) bef0 = bef;
        const int tot = bef + k + after(m - i);
        if (tot > best_len) { best_len = tot; best_start = s - bef; }     // strict '>' (:154)
        if (hp < 0) break;
        const uint64_t cleared = wl & ~(0x8000000000000000ull >> lz);
        if (in_hi) ML.hi = cleared; else ML.lo = cleared;
    }
    start = best_start;
    exact_len = bef0 + k + after(0);
    return best_len;
}

// probe.py:1328-1344 (thresholds) and filter/set_cover_filter.py:429-439 (+-cover_extension, clip to
// the sequence, universe offset) for an anchored range [start, start+len) in probe coordinates.
__device__ __forceinline__ bool finish_range(const ScanParams &P, int len, int start, int exact_len, int L, int64_t d,
                                             int64_t qs, int64_t qe, uint32_t q_ubase, uint32_t &us, uint32_t &ue)
{
    const int64_t qlen = qe - qs;
    int thres = P.lcf;                                // probe.py:1332
    if (L < thres) thres = L;
    if (qlen < (int64_t)thres) thres = (int)qlen;
    if (len < thres) return false;
    if (P.island > 0) {                               // probe.py:1335-1342
        const int ex = (P.m == 0) ? len : exact_len;
        if (ex < P.island) return false;
    }
    int64_t rs = d + start - qs, re = rs + len;
    rs -= P.ext;
    re += P.ext;
    if (rs < 0) rs = 0;
    if (re > qlen) re = qlen;
    us = q_ubase + (uint32_t)rs;
    ue = q_ubase + (uint32_t)re;
    return true;
}

// One surviving hit: probe p, its seed at probe position `pos`, aligned on diagonal d of sequence
// [qs, qe).  False when the seed does not match after all (bucket/tag collision) or the range fails
// the thresholds.
template <int NW, bool FAST>
__device__ __forceinline__ bool eval_hit(const ScanParams &P, const uint64_t *s_tile, int64_t t0, int64_t d, int pos,
                                         uint32_t p, int64_t qs, int64_t qe, uint32_t q_ubase, uint32_t &us,
                                         uint32_t &ue)
{
    const int L = P.plen[p];
    // probe.py:1075-1094: the alignment is clipped to the sequence on both sides
    const int a = (int)max((int64_t)0, qs - d);
    const int bnd = (int)min((int64_t)L, qe - d);
    int start, exact_len, len;
    if constexpr (FAST) {
        static_assert(NW == 2, "the scalar 128-bit path is for two-word probes");
        const int k = P.k;
        const U128 M = mismatch_mask_fast(P, s_tile, t0, d, p);
        // the three ranges [pos, pos+k), [a, pos), [pos+k, bnd) from four "bits below x" masks
        const U128 la = below128(a), lp = below128(pos), lk = below128(pos + k), lb = below128(bnd);
        if ((M.lo & lk.lo & ~lp.lo) | (M.hi & lk.hi & ~lp.hi)) return false;
        U128 ML, MR;
        ML.lo = M.lo & lp.lo & ~la.lo; ML.hi = M.hi & lp.hi & ~la.hi;
        MR.lo = M.lo & lb.lo & ~lk.lo; MR.hi = M.hi & lb.hi & ~lk.hi;
        len = anchored_extend_fast(ML, MR, a, bnd, pos, k, P.m, start, exact_len);
    } else {
        uint64_t M[NW];
        mismatch_mask<NW>(P, s_tile, t0, d, p, M);
        if (any_in_range<NW>(M, pos, pos + P.k)) return false;
        len = anchored_extend<NW>(M, a, bnd, pos, P.k, P.m, start, exact_len);
    }
    return finish_range(P, len, start, exact_len, L, d, qs, qe, q_ubase, us, ue);
}

template <int NW, bool FAST, int HITS_PER_THREAD>
__global__ void __launch_bounds__(SCAN_THREADS, SCAN_MIN_BLOCKS)
scan_kernel(const ScanParams P)
{
    constexpr int QUEUE_CAP = QUEUE_FLUSH + SCAN_THREADS * HITS_PER_THREAD;   // 12 KB at 4 hits per thread
    __shared__ __align__(128) uint64_t s_tile[CB_MAX_SYMBOL_BITS * TW];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_begin_lo[CB_TILE];       // bucket begin (entries < 2^32)
    __shared__ uint32_t s_cum[CB_TILE + 1];
    __shared__ uint32_t s_tag[CB_TILE];            // tag24 | min(distance to the sequence start, 255) << 24
    __shared__ uint32_t s_seq[CB_TILE];
    __shared__ uint64_t s_queue[QUEUE_CAP];        // surviving hits: j << 40 | probe << 8 | pos
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    __shared__ uint32_t s_qcount;
    __shared__ long long s_tile_id;
    __shared__ long long s_qrange[2];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_qcount = 0;
    }
    __syncthreads();
    uint32_t phase = 0;
    unsigned long long local_hits = 0, local_lookups = 0, local_surv = 0;
    const int pw = P.pw;

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
                const int64_t qs = P.seq_start[lo], qe = P.seq_start[lo + 1];
                s_seq[j] = (uint32_t)lo;
                if (g + P.k <= qe) {                 // probe.py:1062 i in [0, len-k]
                    const uint64_t h = kmer_hash(P.bits, P.k, [&](int b, int o) {
                        return read64(s_tile + b * TW, j + CB_FRONT_PAD + o);
                    });
                    const uint32_t bucket = (uint32_t)h & P.bucket_mask;
                    const int64_t b0 = P.bucket_off[bucket], b1 = P.bucket_off[bucket + 1];
                    const int64_t dist = g - qs;
                    s_begin_lo[j] = (uint32_t)b0;
                    s_tag[j] = (uint32_t)(h >> 40) | ((uint32_t)(dist < 255 ? dist : 255) << 24);
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
        if (P.count_only) continue;

        // ---- phase 2: candidate hits in chunks; survivors of the cheap test are queued, the queue is
        // drained densely
        for (uint32_t base = 0; base < total; base += SCAN_THREADS * HITS_PER_THREAD) {
            const uint32_t h0 = base + tid * HITS_PER_THREAD;
            // warp-uniform guard: the whole warp enters or skips, lanes past `total` are predicated
            if (base + (uint32_t)(tid & ~31) * HITS_PER_THREAD < total) {
                int j = 0;
                if (h0 < total) {
                    int lo = 0, hi = CB_TILE;            // largest j with s_cum[j] <= h0
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (s_cum[mid] <= h0) lo = mid; else hi = mid;
                    }
                    j = lo;
                }
                bool surv_v[HITS_PER_THREAD];
                uint64_t task_v[HITS_PER_THREAD];
#pragma unroll
                for (int u = 0; u < HITS_PER_THREAD; u++) {
                    const uint32_t h = h0 + u;
                    bool surv = false;
                    uint64_t task = 0;
                    if (h < total) {
                        while (s_cum[j + 1] <= h) j++;
                        const ulonglong2 ent = __ldg(P.entries + (int64_t)s_begin_lo[j] + (h - s_cum[j]));
                        const uint32_t tg = s_tag[j];
                        if ((uint32_t)(ent.x & 0xffffffull) == (tg & 0xffffffu)) {
                            surv = true;
                            // nearest lower selected seed of the probe: g bases before this seed.  If it
                            // starts inside the sequence and those g bases match too, it matches on this
                            // diagonal, shares the mismatch-free run and gives the same range.
                            const uint32_t g = (uint32_t)(ent.y & 0xffull);
                            if (g != 0u && g <= (tg >> 24)) {
                                // g <= 28 bases per plane: 32-bit reads of the tile and one funnel shift each
                                const int off = j + CB_FRONT_PAD - (int)g;
                                const uint32_t *t32 = reinterpret_cast<const uint32_t *>(s_tile) + (off >> 5);
                                const int sh = off & 31;
                                uint32_t diff = 0u;
                                for (int b = 0; b < P.bits; b++)
                                    diff |= __funnelshift_r(t32[b * 2 * TW], t32[b * 2 * TW + 1], sh) ^
                                            (uint32_t)(ent.y >> (8 + b * pw));
                                surv = (diff << (32 - g)) != 0u;
                            }
                            task = ((uint64_t)j << 40) | (ent.x >> 24);      // j | probe << 8 | pos
                        }
                    }
                    surv_v[u] = surv;
                    task_v[u] = task;
                }
                __syncwarp();
                // one warp-aggregated push for all of them
                unsigned ow[HITS_PER_THREAD];
                uint32_t n_push = 0;
#pragma unroll
                for (int u = 0; u < HITS_PER_THREAD; u++) {
                    ow[u] = __ballot_sync(0xffffffffu, surv_v[u]);
                    n_push += __popc(ow[u]);
                }
                if (n_push) {
                    uint32_t qb = 0;
                    if (lane == 0) qb = atomicAdd(&s_qcount, n_push);
                    qb = __shfl_sync(0xffffffffu, qb, 0);
                    uint32_t before = 0;
#pragma unroll
                    for (int u = 0; u < HITS_PER_THREAD; u++) {
                        if (surv_v[u]) s_queue[qb + before + __popc(ow[u] & ((1u << lane) - 1u))] = task_v[u];
                        before += __popc(ow[u]);
                    }
                }
            }
            __syncthreads();
            const uint32_t qn = s_qcount;
            const bool last = base + SCAN_THREADS * HITS_PER_THREAD >= total;
            if (qn > (uint32_t)QUEUE_FLUSH || (last && qn > 0)) {
                for (uint32_t tb = 0; tb < qn; tb += SCAN_THREADS) {
                    const uint32_t ti = tb + tid;
                    bool emit = false;
                    uint32_t p = 0, us = 0, ue = 0, hit = 0;
                    if (ti < qn) {
                        const uint64_t task = s_queue[ti];
                        const int j = (int)(task >> 40);
                        const int pos = (int)(task & 0xffull);
                        p = (uint32_t)(task >> 8);
                        hit = (uint32_t)(t0 + j);                  // target coordinate of the seed hit (probe.py:1062 `i`)
                        const uint32_t q = s_seq[j];
                        emit = eval_hit<NW, FAST>(P, s_tile, t0, t0 + j - pos, pos, p, P.seq_start[q], P.seq_start[q + 1],
                                                P.seq_ubase[q], us, ue);
                        local_surv++;
                    }
                    // warp-aggregated append
                    const unsigned em = __ballot_sync(0xffffffffu, emit);
                    if (em) {
                        unsigned long long wb = 0;
                        if (lane == 0) wb = atomicAdd(P.rec_cursor, (unsigned long long)__popc(em));
                        wb = __shfl_sync(0xffffffffu, wb, 0);
                        if (emit) {
                            const unsigned long long slot = wb + __popc(em & ((1u << lane) - 1u));
                            // the range's number among the probe's ranges rides in the record (unless the caller
                            // wants the hit positions): the bucketing pass then needs no atomics of its own
                            const uint32_t nth = atomicAdd(&P.rec_count[p], 1u);
                            if (slot < P.rec_cap) P.rec[slot] = make_uint4(p, us, ue, P.keep_hit ? hit : nth);
                        }
                    }
                }
                __syncthreads();
                if (tid == 0) s_qcount = 0;
                __syncthreads();
            }
        }
        __syncthreads();                              // tile + scan arrays are reused by the next tile
    }
    if (local_hits) atomicAdd(P.stat_hits, local_hits);
    if (local_lookups) atomicAdd(P.stat_lookups, local_lookups);
    if (local_surv) atomicAdd(P.stat_owners, local_surv);
}

// bucket the global range list by probe: rec_sorted[rec_off[p] + slot] = (start << 32 | end)
__global__ void scatter_by_probe_kernel(const uint4 *__restrict__ rec, unsigned long long n,
                                        const int64_t *__restrict__ rec_off, uint64_t *__restrict__ rec_sorted)
{
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint4 r = rec[i];                      // probe, start, end, number among the probe's ranges
        rec_sorted[rec_off[r.x] + r.w] = ((uint64_t)r.y << 32) | (uint64_t)r.z;
    }
}

// ---------------------------------------------------------------------------------------
// K4: per-probe sort + merge of the emitted ranges.
// One block per probe.  Normalised bitonic network (every compare-exchange orders min -> lower
// index, so virtual +inf padding to a power of two needs no storage) in shared memory when the
// probe's ranges fit, otherwise in place in global memory.  Then warp 0 merges overlapping or
// touching ranges (utils/interval.py:304-314: start <= curr_end) in 32-wide chunks.
// ---------------------------------------------------------------------------------------
constexpr int MERGE_THREADS = 512;
constexpr int MERGE_SMEM_CAP = 12288;    // ranges sorted in (dynamic) shared memory: 96 KB, two CTAs per SM

__device__ __forceinline__ void bitonic_sort(uint64_t *a, uint32_t n)
{
    // all sizes and strides are powers of two: indices come from shifts and masks (ls = log2 size, lt = log2 stride)
    int ln2 = 0;
    while ((1u << ln2) < n) ln2++;
    const uint32_t half = (1u << ln2) >> 1;
    for (int ls = 1; ls <= ln2; ls++) {
        // first step of each merge: compare i with its mirror inside the block of `size`
        const uint32_t size = 1u << ls;
        for (uint32_t t = threadIdx.x; t < half; t += blockDim.x) {
            const uint32_t base = (t >> (ls - 1)) << ls, o = t & ((size >> 1) - 1u);
            const uint32_t i = base + o, j = base + size - 1u - o;
            if (j < n) {
                const uint64_t x = a[i], y = a[j];
                if (x > y) { a[i] = y; a[j] = x; }
            }
        }
        __syncthreads();
        for (int lt = ls - 2; lt >= 0; lt--) {
            const uint32_t stride = 1u << lt;
            for (uint32_t t = threadIdx.x; t < half; t += blockDim.x) {
                const uint32_t i = ((t >> lt) << (lt + 1)) + (t & (stride - 1u)), j = i + stride;
                if (j < n) {
                    const uint64_t x = a[i], y = a[j];
                    if (x > y) { a[i] = y; a[j] = x; }
                }
            }
            __syncthreads();
        }
    }
}

// Warp merge of a sorted list a[0..n): overlapping or touching ranges (utils/interval.py:304-314:
// start <= curr_end) are united; the result is written in place at the front of g (g may alias a).
// Returns the number of merged ranges (valid in every lane); local_max tracks the longest range.
__device__ __forceinline__ uint32_t warp_merge_sorted(const uint64_t *a, uint64_t *g, uint32_t n, int lane,
                                                      uint32_t &local_max)
{
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
    return n_out;
}

// Probes with more than MERGE_WARP_CAP ranges (a probe that hits thousands of genomes: the influenza shape): one block
// per entry of the list the warp kernel left behind; sorted in shared memory up to MERGE_SMEM_CAP ranges, in place in
// global memory beyond.
__global__ void __launch_bounds__(MERGE_THREADS)
merge_kernel(const int64_t *__restrict__ rec_off, uint64_t *__restrict__ rec, uint32_t *__restrict__ n_merged,
             uint32_t *__restrict__ max_len, const uint32_t *__restrict__ big_list, const unsigned int *__restrict__ n_big)
{
    extern __shared__ uint64_t s_rec[];
    const unsigned int todo = *n_big;                // usually 0: the warp kernels took every probe
    uint32_t local_max = 0;
    for (unsigned int i = blockIdx.x; i < todo; i += gridDim.x) {
        const int64_t p = big_list[i];
        const int64_t o0 = rec_off[p];
        const uint32_t n = (uint32_t)(rec_off[p + 1] - o0);
        uint64_t *g = rec + o0;
        uint64_t *a = g;
        const bool in_smem = n <= (uint32_t)MERGE_SMEM_CAP;
        if (in_smem) {
            for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) s_rec[t] = g[t];
            a = s_rec;
        }
        __syncthreads();
        if (n > 1) bitonic_sort(a, n);
        if (threadIdx.x < 32) {
            const uint32_t n_out = warp_merge_sorted(a, g, n, threadIdx.x, local_max);
            if (threadIdx.x == 0) n_merged[p] = n_out;
        }
        __syncthreads();
    }
    if (threadIdx.x < 32) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
        if (threadIdx.x == 0 && local_max) atomicMax(max_len, local_max);
    }
}

// Probes with at most MERGE_WARP_CAP ranges (all of them in the usual one-range-per-genome case): one WARP per
// probe, everything in registers.  Element i of the probe's list lives in register i % EPL of lane i / EPL (padded
// with +inf to 32 * EPL elements): the steps of the bitonic network with a stride below EPL are compare-exchanges
// between registers of the same lane, only the strides of EPL and more (15 of the 45 steps of a 512-element list)
// go through warp shuffles; the merge of the sorted list is a sequential pass of every lane over its own EPL
// consecutive ranges with three warp scans in between.  No shared memory, no barrier.
constexpr int MERGE_WARP_CAP = 1024;     // the scan emits one range per (probe, diagonal, run): ~260 per probe at Zika scale
constexpr int MERGE_WARPS = 4;

// The size loop stays rolled -- fully unrolled, the network of a 512-element list is ~100 KB of straight-line code
// per variant and the warps of an SM, each at a different place in it, starve on instruction fetch (measured: 3.6x
// slower) -- only the loops over a lane's registers are unrolled.
template <int EPL>
__device__ __forceinline__ void warp_bitonic_regs(uint64_t (&v)[EPL], int lane)
{
#pragma unroll 1
    for (int k = 2; k <= 32 * EPL; k <<= 1) {
        // ascending iff bit k of the element index is clear; for k >= EPL that is a property of the lane
#pragma unroll 1
        for (int j = k >> 1; j >= EPL; j >>= 1) {
            const int jl = j / EPL;
            const bool keep_min = ((lane & jl) == 0) == (((lane * EPL) & k) == 0);
#pragma unroll
            for (int r = 0; r < EPL; r++) {
                const uint64_t mine = v[r];
                const uint64_t other = __shfl_xor_sync(0xffffffffu, mine, jl);
                v[r] = (keep_min == (other < mine)) ? other : mine;
            }
        }
#pragma unroll
        for (int J = EPL / 2; J >= 1; J >>= 1) {
            if (J < k) {
#pragma unroll
                for (int r = 0; r < EPL; r++) {
                    if ((r & J) == 0) {
                        const bool up = ((lane * EPL + r) & k) == 0;
                        const uint64_t a = v[r], b = v[r | J];
                        const bool sw = up == (a > b);
                        v[r] = sw ? b : a;
                        v[r | J] = sw ? a : b;
                    }
                }
            }
        }
    }
}

__device__ __forceinline__ uint32_t warp_excl_max(uint32_t x, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x = max(x, t);
    }
    x = __shfl_up_sync(0xffffffffu, x, 1);
    return lane == 0 ? 0u : x;
}

// Sort + merge of one probe's n ranges (n <= 32 * EPL) by one warp; the merged ranges are written to the front of
// g (utils/interval.py:304-314: a range joins the current one iff start <= curr_end).  Returns their number.
template <int EPL>
__device__ __forceinline__ uint32_t warp_sort_merge(uint64_t *__restrict__ g, uint32_t n, int lane, uint32_t &local_max)
{
    uint64_t v[EPL];
#pragma unroll
    for (int r = 0; r < EPL; r++) {                  // any element may start in any slot: coalesced loads
        const uint32_t i = (uint32_t)(r * 32 + lane);
        v[r] = i < n ? g[i] : ~0ull;
    }
    warp_bitonic_regs<EPL>(v, lane);
    // lane L now holds the elements [L * EPL, (L + 1) * EPL) of the sorted list; starts of ranges are < 2^32 - 1
    const uint32_t first = (uint32_t)(lane * EPL);
    // largest end before this lane's elements
    uint32_t lane_max = 0;
#pragma unroll
    for (int r = 0; r < EPL; r++)
        if (first + r < n) lane_max = max(lane_max, (uint32_t)v[r]);
    const uint32_t max_in = warp_excl_max(lane_max, lane);
    // heads (a range that starts beyond everything before it) and the start of the last head in this lane
    uint32_t run = max_in, last_head = 0;
    unsigned heads = 0;
#pragma unroll
    for (int r = 0; r < EPL; r++) {
        const uint32_t s = (uint32_t)(v[r] >> 32);
        if (first + r < n) {
            if (first + r == 0 || s > run) { heads |= 1u << r; last_head = s; }
            run = max(run, (uint32_t)v[r]);
        }
    }
    // start of the group that is open when this lane begins (starts ascend, so the latest head is the largest)
    uint32_t gs = warp_excl_max(last_head, lane);
    // a range closes its group iff the next range of the list is a head (or there is none)
    const unsigned next_heads = __shfl_down_sync(0xffffffffu, heads, 1);
    const bool next_lane_head = lane < 31 && (next_heads & 1u);
    unsigned tails = 0;
#pragma unroll
    for (int r = 0; r < EPL; r++) {
        if (first + r < n) {
            const bool last = first + r + 1 == n;
            const bool nh = r + 1 < EPL ? ((heads >> (r + 1 < EPL ? r + 1 : 0)) & 1u) : next_lane_head;
            if (last || nh) tails |= 1u << r;
        }
    }
    uint32_t before = (uint32_t)__popc(tails);
    uint32_t total = before;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, total, o);
        if (lane >= o) total += t;
    }
    uint32_t k_out = total - before;                 // merged ranges that end in earlier lanes
    const uint32_t n_out = __shfl_sync(0xffffffffu, total, 31);
    run = max_in;
#pragma unroll
    for (int r = 0; r < EPL; r++) {
        if (first + r < n) {
            if ((heads >> r) & 1u) gs = (uint32_t)(v[r] >> 32);
            run = max(run, (uint32_t)v[r]);
            if ((tails >> r) & 1u) {
                g[k_out++] = ((uint64_t)gs << 32) | (uint64_t)run;
                local_max = max(local_max, run - gs);
            }
        }
    }
    return n_out;
}

// WIDE = false: probes with at most 512 ranges (16 registers of ranges per lane: 64 registers per thread, eight
// CTAs per SM); the probes with 513..1024 ranges are appended to wide_list and done by the WIDE = true launch, one
// warp per list entry (few probes, each a long sort: spread over the whole grid instead of wherever they happen to
// fall).  Probes with more than MERGE_WARP_CAP ranges are appended to big_list for the block-per-probe kernel.
template <bool WIDE>
__global__ void __launch_bounds__(MERGE_WARPS * 32, WIDE ? 3 : 8)
merge_warp_kernel(const int64_t *__restrict__ rec_off, uint64_t *__restrict__ rec, int64_t n_probes,
                  uint32_t *__restrict__ n_merged, uint32_t *__restrict__ max_len, uint32_t *__restrict__ big_list,
                  unsigned int *__restrict__ big_count, uint32_t *__restrict__ wide_list,
                  unsigned int *__restrict__ wide_count)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t local_max = 0;
    if (WIDE) {
        const unsigned int n_wide = *wide_count;
        for (unsigned int i = blockIdx.x * MERGE_WARPS + warp; i < n_wide; i += gridDim.x * MERGE_WARPS) {
            const int64_t p = wide_list[i];
            const int64_t o0 = rec_off[p];
            const uint32_t n = (uint32_t)(rec_off[p + 1] - o0);
            const uint32_t n_out = warp_sort_merge<32>(rec + o0, n, lane, local_max);
            if (lane == 0) n_merged[p] = n_out;
        }
    } else {
        for (int64_t p = (int64_t)blockIdx.x * MERGE_WARPS + warp; p < n_probes; p += (int64_t)gridDim.x * MERGE_WARPS) {
            const int64_t o0 = rec_off[p];
            const uint32_t n = (uint32_t)(rec_off[p + 1] - o0);
            uint64_t *g = rec + o0;
            uint32_t n_out;
            if (n == 0) {
                if (lane == 0) n_merged[p] = 0;
                continue;
            }
            if (n > 512u) {
                if (lane == 0) {
                    if (n > (uint32_t)MERGE_WARP_CAP) big_list[atomicAdd(big_count, 1u)] = (uint32_t)p;
                    else wide_list[atomicAdd(wide_count, 1u)] = (uint32_t)p;
                }
                continue;
            }
            if (n <= 32) n_out = warp_sort_merge<1>(g, n, lane, local_max);
            else if (n <= 64) n_out = warp_sort_merge<2>(g, n, lane, local_max);
            else if (n <= 128) n_out = warp_sort_merge<4>(g, n, lane, local_max);
            else if (n <= 256) n_out = warp_sort_merge<8>(g, n, lane, local_max);
            else n_out = warp_sort_merge<16>(g, n, lane, local_max);
            if (lane == 0) n_merged[p] = n_out;
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    if (lane == 0 && local_max) atomicMax(max_len, local_max);
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

// K4 launcher: warp-per-probe kernel first; the block-per-probe kernel only if some probe has
// more than MERGE_WARP_CAP ranges.
int launch_merge(cb_ctx *ctx, const int64_t *d_roff, uint64_t *d_sorted, int64_t P, uint32_t *d_nmerged, uint32_t *d_maxlen)
{
    cudaStream_t st = ctx->stream;
    DevBuf<unsigned int> d_large;            // [0] length of the big list, [1] length of the wide list
    DevBuf<uint32_t> d_wide;                 // [0, P) wide list, [P, 2P) big list
    const size_t Pn = (size_t)(P > 0 ? P : 1);
    CB_CUDA(ctx, d_large.alloc(2));
    CB_CUDA(ctx, d_wide.alloc(2 * Pn));
    CB_CUDA(ctx, cudaMemsetAsync(d_large.p, 0, 2 * sizeof(unsigned int), st));
    int64_t g = (P + MERGE_WARPS - 1) / MERGE_WARPS;
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    merge_warp_kernel<false><<<(unsigned)g, MERGE_WARPS * 32, 0, st>>>(d_roff, d_sorted, P, d_nmerged, d_maxlen, d_wide.p + Pn,
                                                                       d_large.p, d_wide.p, d_large.p + 1);
    merge_warp_kernel<true><<<(unsigned)g, MERGE_WARPS * 32, 0, st>>>(d_roff, d_sorted, P, d_nmerged, d_maxlen, d_wide.p + Pn,
                                                                      d_large.p, d_wide.p, d_large.p + 1);
    ctx->launches += 2;
    CB_CUDA(ctx, cudaGetLastError());
    // probes with more ranges than the warp kernels take: listed on the device (no host round trip in the
    // middle of stage A); the blocks leave at once when the list is empty
    {
        static bool attr_set[64] = {};
        if (ctx->device < 0 || ctx->device >= 64 || !attr_set[ctx->device]) {
            CB_CUDA(ctx, cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)(MERGE_SMEM_CAP * sizeof(uint64_t))));
            if (ctx->device >= 0 && ctx->device < 64) attr_set[ctx->device] = true;
        }
    }
    merge_kernel<<<(unsigned)(ctx->sm_count * 2), MERGE_THREADS, MERGE_SMEM_CAP * sizeof(uint64_t), st>>>(
        d_roff, d_sorted, d_nmerged, d_maxlen, d_wide.p + Pn, d_large.p);
    ctx->launches++;
    CB_CUDA(ctx, cudaGetLastError());
    return CB_OK;
}

template <int NW, bool FAST, int HPT>
int launch_scan(cb_ctx *ctx, const ScanParams &P, int grid)
{
    scan_kernel<NW, FAST, HPT><<<grid, SCAN_THREADS, 0, ctx->stream>>>(P);
    ctx->launches++;
    CB_CUDA(ctx, cudaGetLastError());
    return CB_OK;
}

int launch_scan_nw(cb_ctx *ctx, int nw, const ScanParams &P, int grid)
{
    const bool generic = getenv("CB_SCAN_GENERIC") != nullptr;     // testing: force the multi-word path
    const char *hpt = getenv("CB_SCAN_HPT");                        // tuning: candidate hits per thread and chunk
    switch (nw) {
    case 1: return launch_scan<1, false, 4>(ctx, P, grid);
    case 2:
        if (generic) return launch_scan<2, false, 4>(ctx, P, grid);
        if (hpt && atoi(hpt) == 8) return launch_scan<2, true, 8>(ctx, P, grid);
        if (hpt && atoi(hpt) == 2) return launch_scan<2, true, 2>(ctx, P, grid);
        return launch_scan<2, true, 4>(ctx, P, grid);
    case 3: return launch_scan<3, false, 4>(ctx, P, grid);
    case 4: return launch_scan<4, false, 4>(ctx, P, grid);
    }
    return cb_fail(ctx, CB_ERR_UNSUPPORTED, "probe longer than CB_MAX_PROBE_LEN");
}

}  // namespace

int cb_coverage_impl(cb_ctx *ctx, const cb_probes *probes, const cb_targets *targets,
                     const cb_hyb_params *hp, const int64_t *seed_off, const int32_t *seed_pos,
                     const uint8_t *seed_pos_u8, int32_t seeds_per_probe, int64_t probe_lo, int64_t probe_hi,
                     cb_cover **out, cb_stats *stats, std::vector<uint32_t> *raw_records)
{
    if (!probes || !targets || !hp || !out) return cb_fail(ctx, CB_ERR_ARG, "null argument");
    // [probe_lo, probe_hi): the probes this call scans (a rank's shard); the cover keeps global probe ids
    if (probe_hi < 0) probe_hi = probes->n_probes;
    if (probe_lo < 0 || probe_lo > probe_hi || probe_hi > probes->n_probes) return cb_fail(ctx, CB_ERR_ARG, "bad probe range");
    const bool uniform = seed_pos_u8 != nullptr;        // [n_probes][seeds_per_probe] bytes, no CSR
    if (uniform ? seeds_per_probe < 0 : (probes->n_probes > 0 && (!seed_off || !seed_pos)))
        return cb_fail(ctx, CB_ERR_ARG, "bad seed arguments");
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
    CB_CUDA(ctx, cb_dev_alloc(st, (void **)&cov->d_ubase, sizeof(uint32_t) * (size_t)(targets->n_genomes + 1)));
    CB_CUDA(ctx, cudaMemcpyAsync(cov->d_ubase, targets->d_ubase, sizeof(uint32_t) * (size_t)(targets->n_genomes + 1),
                                 cudaMemcpyDeviceToDevice, st));
    CB_CUDA(ctx, cb_dev_alloc(st, (void **)&cov->d_iv_off, sizeof(int64_t) * (size_t)(P + 1)));

    const int64_t n_raw_seeds = !P ? 0 : uniform ? (probe_hi - probe_lo) * (int64_t)seeds_per_probe : seed_off[P] - seed_off[0];
    const bool empty = (P == 0 || probe_hi == probe_lo || targets->total_bases == 0 || n_raw_seeds == 0);
    if (empty) {
        CB_CUDA(ctx, cudaMemsetAsync(cov->d_iv_off, 0, sizeof(int64_t) * (size_t)(P + 1), st));
        CB_CUDA(ctx, cudaStreamSynchronize(st));
        guard.c = nullptr;
        *out = cov;
        return CB_OK;
    }
    // CSR form: narrow the seed positions to bytes (range-checked here, against the probe length on
    // the device).  Uniform form: the caller's bytes go to the device as they are.
    std::vector<uint8_t> h_spos;
    std::vector<int64_t> h_soff;
    if (!uniform) {
        h_spos.resize((size_t)n_raw_seeds);
        h_soff.resize((size_t)P + 1);
        for (int64_t p = 0; p <= P; p++) {
            h_soff[(size_t)p] = seed_off[p] - seed_off[0];
            if (p && seed_off[p] < seed_off[p - 1]) return cb_fail(ctx, CB_ERR_ARG, "seed_off not monotone");
        }
        const int32_t *sp0 = seed_pos + seed_off[0];
        int bad = 0;
        for (int64_t e = 0; e < n_raw_seeds; e++) {
            const int32_t s = sp0[e];
            bad |= (s < 0) | (s >= CB_MAX_PROBE_LEN);
            h_spos[(size_t)e] = (uint8_t)s;
        }
        if (bad) return cb_fail(ctx, CB_ERR_ARG, "seed position out of range");
    }
    const int nw = probes->nw, bits = probes->bits;
    const int plane_words = bits * nw, prec_words = plane_words + nw;
    DevBuf<uint32_t> d_eprobe, d_bcount, d_bcursor, d_ndist;
    DevBuf<uint8_t> d_spos, d_epos;
    DevBuf<int64_t> d_boff, d_soff, d_eoff;
    DevBuf<ulonglong2> d_entries;
    DevBuf<uint64_t> d_precs;
    DevBuf<int> d_bad;
    DevBuf<unsigned long long> d_ctr;        // [0] tile counter, [1] hits, [2] lookups, [3] owners, [4] range cursor
    if (!uniform) CB_CUDA(ctx, d_soff.alloc((size_t)P + 1));
    CB_CUDA(ctx, d_spos.alloc((size_t)n_raw_seeds));
    CB_CUDA(ctx, d_ndist.alloc((size_t)P));
    CB_CUDA(ctx, d_eoff.alloc((size_t)P + 1));
    CB_CUDA(ctx, d_precs.alloc((size_t)P * (size_t)prec_words));
    CB_CUDA(ctx, d_bad.alloc(1));
    CB_CUDA(ctx, d_ctr.alloc(8));
    CB_CUDA(ctx, cudaMemsetAsync(d_bad.p, 0, sizeof(int), st));
    if (!uniform)
        CB_CUDA(ctx, cudaMemcpyAsync(d_soff.p, h_soff.data(), sizeof(int64_t) * (size_t)(P + 1), cudaMemcpyHostToDevice, st));
    CB_CUDA(ctx, cudaMemcpyAsync(d_spos.p, uniform ? seed_pos_u8 : h_spos.data(), (size_t)n_raw_seeds, cudaMemcpyHostToDevice, st));

    // ---- K2 seed index (+ probe records with their seed masks)
    t_idx.start();
    const int wide = ctx->sm_count * 8;
    build_precs_kernel<<<wide, 256, 0, st>>>(probes->d_words, P, plane_words, prec_words, d_precs.p);
    seed_mask_kernel<<<wide, 256, 0, st>>>(uniform ? nullptr : d_soff.p, (int)seeds_per_probe, d_spos.p, probes->d_len, hp->k, P,
                                           probe_lo, probe_hi, d_precs.p, prec_words, nw, d_ndist.p, d_bad.p);
    ctx->launches += 2;
    CB_CUDA(ctx, cudaGetLastError());
    int64_t n_entries = 0;
    CB_TRY(cb_exclusive_scan_u32_to_i64(ctx, d_ndist.p, d_eoff.p, P, &n_entries));
    {
        int h_bad = 0;
        CB_CUDA(ctx, cudaMemcpyAsync(&h_bad, d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        CB_CUDA(ctx, cudaStreamSynchronize(st));
        if (h_bad) return cb_fail(ctx, CB_ERR_ARG, "seed position + k exceeds the probe length");
    }
    if (n_entries >= (int64_t)0xffffffffll) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "too many seed entries");
    CB_CUDA(ctx, d_eprobe.alloc((size_t)n_entries));
    CB_CUDA(ctx, d_epos.alloc((size_t)n_entries));
    CB_CUDA(ctx, d_entries.alloc((size_t)n_entries));
    int64_t nb = 1024;
    while (nb < 2 * n_entries) nb <<= 1;
    CB_CUDA(ctx, d_bcount.alloc((size_t)nb));
    CB_CUDA(ctx, d_bcursor.alloc((size_t)nb));
    CB_CUDA(ctx, d_boff.alloc((size_t)nb + 1));
    CB_CUDA(ctx, cudaMemsetAsync(d_bcount.p, 0, sizeof(uint32_t) * (size_t)nb, st));
    CB_CUDA(ctx, cudaMemsetAsync(d_bcursor.p, 0, sizeof(uint32_t) * (size_t)nb, st));
    seed_expand_kernel<<<wide, 256, 0, st>>>(d_precs.p, prec_words, nw, P, d_eoff.p, d_eprobe.p, d_epos.p);
    seed_index_kernel<false><<<wide, 256, 0, st>>>(d_eprobe.p, d_epos.p, n_entries, d_precs.p, prec_words, bits, nw,
                                                   hp->k, (uint32_t)(nb - 1), d_bcount.p, nullptr, nullptr, nullptr);
    ctx->launches += 2;
    CB_CUDA(ctx, cudaGetLastError());
    CB_TRY(cb_exclusive_scan_u32_to_i64(ctx, d_bcount.p, d_boff.p, nb, nullptr));
    seed_index_kernel<true><<<wide, 256, 0, st>>>(d_eprobe.p, d_epos.p, n_entries, d_precs.p, prec_words, bits, nw,
                                                  hp->k, (uint32_t)(nb - 1), nullptr, d_boff.p, d_bcursor.p,
                                                  d_entries.p);
    ctx->launches++;
    CB_CUDA(ctx, cudaGetLastError());
    t_idx.stop();

    DevBuf<uint32_t> d_rcount;
    DevBuf<int64_t> d_roff;
    CB_CUDA(ctx, d_rcount.alloc((size_t)P));
    CB_CUDA(ctx, d_roff.alloc((size_t)P + 1));

    ScanParams sp;
    memset(&sp, 0, sizeof sp);
    sp.planes = targets->d_planes;
    sp.plane_words = targets->plane_words;
    sp.bits = bits;
    sp.total_bases = targets->total_bases;
    sp.n_seqs = targets->n_seqs;
    sp.seq_start = targets->d_seq_start;
    sp.seq_ubase = targets->d_seq_ubase;
    sp.precs = d_precs.p;
    sp.prec_words = prec_words;
    sp.plen = probes->d_len;
    sp.bucket_off = d_boff.p;
    sp.entries = d_entries.p;
    sp.pw = 56 / bits;
    sp.bucket_mask = (uint32_t)(nb - 1);
    sp.m = hp->mismatches;
    sp.lcf = hp->lcf_thres;
    sp.island = hp->island_of_exact_match;
    sp.ext = hp->cover_extension;
    sp.k = hp->k;
    sp.rec_count = d_rcount.p;
    sp.keep_hit = raw_records ? 1 : 0;
    sp.n_tiles = (targets->total_bases + CB_TILE - 1) / CB_TILE;
    sp.tile_counter = d_ctr.p;
    sp.stat_hits = d_ctr.p + 1;
    sp.stat_lookups = d_ctr.p + 2;
    sp.stat_owners = d_ctr.p + 3;
    sp.rec_cursor = d_ctr.p + 4;
    int per_sm = 4;
    if (const char *e = getenv("CB_SCAN_BLOCKS_PER_SM")) per_sm = atoi(e) > 0 ? atoi(e) : per_sm;
    int64_t grid64 = sp.n_tiles < (int64_t)ctx->sm_count * per_sm ? sp.n_tiles : (int64_t)ctx->sm_count * per_sm;
    const int grid = (int)grid64;

    // ---- K3 pre-pass: upper bound on the number of candidate hits (sizes the range list).  Skipped when the
    // previous scan of this context ran with the same parameters: then the list gets twice the room its density of
    // ranges per (probe x target base) asks for, and the pre-pass only runs if that overflows.
    unsigned long long h_ctr[8];
    unsigned long long hits_ub = 0;
    bool counted = false;
    auto count_pass = [&]() -> int {
        t_cnt.start();
        CB_CUDA(ctx, cudaMemsetAsync(d_ctr.p, 0, sizeof(unsigned long long) * 8, st));
        sp.count_only = 1;
        CB_TRY(launch_scan_nw(ctx, nw, sp, grid));
        t_cnt.stop();
        CB_CUDA(ctx, cudaMemcpyAsync(h_ctr, d_ctr.p, sizeof h_ctr, cudaMemcpyDeviceToHost, st));
        CB_CUDA(ctx, cudaStreamSynchronize(st));
        hits_ub = h_ctr[1];
        counted = true;
        return CB_OK;
    };
    const int32_t dkey[5] = {hp->mismatches, hp->lcf_thres, hp->island_of_exact_match, hp->k,
                             (int32_t)probes->max_len};
    const double pairs_scanned = (double)(probe_hi - probe_lo) * (double)targets->total_bases;
    const bool hinted = ctx->range_density > 0.0 && memcmp(dkey, ctx->range_density_key, sizeof dkey) == 0 &&
                        !getenv("CB_SCAN_PREPASS");
    unsigned long long cap;
    if (hinted) {
        cap = (unsigned long long)(2.0 * ctx->range_density * pairs_scanned) + 65536ull;
    } else {
        CB_TRY(count_pass());
        // one range at most per surviving hit (one hit per mismatch-free run that holds a seed): start with a
        // third of the candidate hits and fall back to the hard bound if the list overflows
        cap = hits_ub / 3 + 4096;
    }

    // ---- K3 scan
    DevBuf<uint4> d_rec;
    unsigned long long n_raw = 0;
    t_emit.start();
    for (int attempt = 0; attempt < 2; attempt++) {
        CB_CUDA(ctx, d_rec.alloc((size_t)cap));
        CB_CUDA(ctx, cudaMemsetAsync(d_rcount.p, 0, sizeof(uint32_t) * (size_t)P, st));
        CB_CUDA(ctx, cudaMemsetAsync(d_ctr.p, 0, sizeof(unsigned long long) * 8, st));
        sp.count_only = 0;
        sp.rec = d_rec.p;
        sp.rec_cap = cap;
        CB_TRY(launch_scan_nw(ctx, nw, sp, grid));
        CB_CUDA(ctx, cudaMemcpyAsync(h_ctr, d_ctr.p, sizeof h_ctr, cudaMemcpyDeviceToHost, st));
        CB_CUDA(ctx, cudaStreamSynchronize(st));
        n_raw = h_ctr[4];
        if (n_raw <= cap) break;
        if (attempt == 1) return cb_fail(ctx, CB_ERR_STATE, "range list overflow after retry");
        if (!counted) CB_TRY(count_pass());
        cap = hits_ub + 4096;
    }
    if (pairs_scanned > 0) {
        ctx->range_density = (double)n_raw / pairs_scanned;
        memcpy(ctx->range_density_key, dkey, sizeof dkey);
    }
    t_emit.stop();
    if (n_raw >= 0xfffffff0ull * 16ull) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "too many cover ranges");
    if (raw_records) {
        // the caller wants the emitted ranges themselves (probe, start, end, hit position), unmerged: the
        // consumers with merge_overlapping=False semantics (coverage_analysis.py:228-231) and the adapter
        // filter's interval scheduling (adapter_filter.py:191-238) take it from here on the host
        raw_records->resize((size_t)n_raw * 4);
        if (n_raw) CB_CUDA(ctx, cudaMemcpyAsync(raw_records->data(), d_rec.p, sizeof(uint4) * (size_t)n_raw, cudaMemcpyDeviceToHost, st));
        CB_CUDA(ctx, cudaMemsetAsync(cov->d_iv_off, 0, sizeof(int64_t) * (size_t)(P + 1), st));
        CB_CUDA(ctx, cudaStreamSynchronize(st));
        if (stats) {
            stats->ms_seed_index = t_idx.ms();
            stats->ms_scan_count = counted ? t_cnt.ms() : 0.0;
            stats->ms_scan_emit = t_emit.ms();
            stats->n_seed_entries = n_entries;
            stats->n_candidate_hits = (int64_t)h_ctr[1];
            stats->n_raw_ranges = (int64_t)n_raw;
            stats->n_kernel_launches = ctx->launches;
        }
        guard.c = nullptr;
        *out = cov;
        return CB_OK;
    }

    // ---- bucket by probe, K4 merge
    t_merge.start();
    CB_TRY(cb_exclusive_scan_u32_to_i64(ctx, d_rcount.p, d_roff.p, P, nullptr));
    DevBuf<uint64_t> d_sorted;
    DevBuf<uint32_t> d_nmerged, d_maxlen;
    CB_CUDA(ctx, d_sorted.alloc((size_t)n_raw));
    CB_CUDA(ctx, d_nmerged.alloc((size_t)P));
    CB_CUDA(ctx, d_maxlen.alloc(1));
    CB_CUDA(ctx, cudaMemsetAsync(d_maxlen.p, 0, sizeof(uint32_t), st));
    if (n_raw) {
        scatter_by_probe_kernel<<<wide, 256, 0, st>>>(d_rec.p, n_raw, d_roff.p, d_sorted.p);
        ctx->launches++;
    }
    CB_TRY(launch_merge(ctx, d_roff.p, d_sorted.p, P, d_nmerged.p, d_maxlen.p));
    int64_t n_iv = 0;
    CB_TRY(cb_exclusive_scan_u32_to_i64(ctx, d_nmerged.p, cov->d_iv_off, P, &n_iv));
    CB_CUDA(ctx, cb_dev_alloc(st, (void **)&cov->d_iv, sizeof(uint2) * (size_t)(n_iv ? n_iv : 1)));
    compact_kernel<<<wide, 256, 0, st>>>(d_roff.p, d_sorted.p, cov->d_iv_off, P, cov->d_iv);
    ctx->launches++;
    CB_CUDA(ctx, cudaGetLastError());
    t_merge.stop();
    t_all.stop();
    cov->n_intervals = n_iv;

    CB_CUDA(ctx, cudaMemcpyAsync(&cov->max_interval_len, d_maxlen.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    if (stats) {
        stats->ms_seed_index = t_idx.ms();
        stats->ms_scan_count = counted ? t_cnt.ms() : 0.0;
        stats->ms_scan_emit = t_emit.ms();
        stats->ms_merge = t_merge.ms();
        stats->ms_total = t_all.ms();
        stats->n_seed_entries = n_entries;
        stats->n_seed_lookups = (int64_t)h_ctr[2];
        stats->n_candidate_hits = (int64_t)h_ctr[1];
        stats->n_raw_ranges = (int64_t)n_raw;
        stats->n_intervals = n_iv;
        stats->n_kernel_launches = ctx->launches;
        stats->reserved[0] = (int64_t)h_ctr[3];     // hits that survived the cheap test (anchored extensions run)
    }
    guard.c = nullptr;
    *out = cov;
    return CB_OK;
}


// Cover from host intervals: same bucket-by-probe + merge path as the scan output.
int cb_cover_import_impl(cb_ctx *ctx, int64_t P, int32_t NG, const int64_t *genome_len, int64_t n,
                         const int64_t *probe_id, const int32_t *genome, const int64_t *start,
                         const int64_t *end, cb_cover **out)
{
    if (!out || P < 0 || NG < 0 || n < 0) return cb_fail(ctx, CB_ERR_ARG, "bad argument");
    if (NG > 0 && !genome_len) return cb_fail(ctx, CB_ERR_ARG, "null genome_len");
    if (n > 0 && (!probe_id || !genome || !start || !end)) return cb_fail(ctx, CB_ERR_ARG, "null interval array");
    cudaStream_t st = ctx->stream;
    cb_cover *cov = new cb_cover();
    struct Guard { cb_cover *c; ~Guard() { if (c) cb_cover_free(c); } } guard{cov};
    cov->ctx = ctx;
    cov->n_probes = P;
    cov->n_genomes = NG;
    cov->h_ubase.resize((size_t)NG + 1);
    cov->h_genome_len.assign(genome_len, genome_len + NG);
    uint64_t ub = 0;
    for (int32_t g = 0; g < NG; g++) {
        if (genome_len[g] < 0) return cb_fail(ctx, CB_ERR_ARG, "negative genome length");
        cov->h_ubase[(size_t)g] = (uint32_t)ub;
        ub = (ub + (uint64_t)genome_len[g] + 1 + 63) & ~63ull;
        if (ub >= 0xffffff00ull) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "universe exceeds 2^32 bits");
    }
    cov->h_ubase[(size_t)NG] = (uint32_t)ub;
    cov->universe_bits = (int64_t)ub;
    std::vector<uint4> h_rec((size_t)n);
    std::vector<uint32_t> h_count((size_t)P, 0u);
    for (int64_t i = 0; i < n; i++) {
        if (probe_id[i] < 0 || probe_id[i] >= P || genome[i] < 0 || genome[i] >= NG || start[i] < 0 ||
            end[i] < start[i] || end[i] > genome_len[genome[i]])
            return cb_fail(ctx, CB_ERR_ARG, "interval out of range");
        const uint32_t b = cov->h_ubase[(size_t)genome[i]];
        h_rec[(size_t)i] = make_uint4((uint32_t)probe_id[i], b + (uint32_t)start[i], b + (uint32_t)end[i], 0u);
        if (end[i] > start[i]) h_rec[(size_t)i].w = h_count[(size_t)probe_id[i]]++;   // number among the probe's ranges
        else h_rec[(size_t)i].x = 0xffffffffu;            // empty interval: dropped
    }
    // compact away empty intervals on the host
    size_t m = 0;
    for (size_t i = 0; i < (size_t)n; i++) if (h_rec[i].x != 0xffffffffu) h_rec[m++] = h_rec[i];
    CB_CUDA(ctx, cb_dev_alloc(st, (void **)&cov->d_ubase, sizeof(uint32_t) * (size_t)(NG + 1)));
    CB_CUDA(ctx, cudaMemcpyAsync(cov->d_ubase, cov->h_ubase.data(), sizeof(uint32_t) * (size_t)(NG + 1), cudaMemcpyHostToDevice, st));
    CB_CUDA(ctx, cb_dev_alloc(st, (void **)&cov->d_iv_off, sizeof(int64_t) * (size_t)(P + 1)));
    if (P == 0 || m == 0) {
        CB_CUDA(ctx, cudaMemsetAsync(cov->d_iv_off, 0, sizeof(int64_t) * (size_t)(P + 1), st));
        CB_CUDA(ctx, cudaStreamSynchronize(st));
        guard.c = nullptr;
        *out = cov;
        return CB_OK;
    }
    DevBuf<uint4> d_rec;
    DevBuf<uint32_t> d_rcount, d_nmerged, d_maxlen;
    DevBuf<int64_t> d_roff;
    DevBuf<uint64_t> d_sorted;
    CB_CUDA(ctx, d_rec.alloc(m));
    CB_CUDA(ctx, d_rcount.alloc((size_t)P));
    CB_CUDA(ctx, d_roff.alloc((size_t)P + 1));
    CB_CUDA(ctx, d_sorted.alloc(m));
    CB_CUDA(ctx, d_nmerged.alloc((size_t)P));
    CB_CUDA(ctx, d_maxlen.alloc(1));
    CB_CUDA(ctx, cudaMemcpyAsync(d_rec.p, h_rec.data(), sizeof(uint4) * m, cudaMemcpyHostToDevice, st));
    CB_CUDA(ctx, cudaMemcpyAsync(d_rcount.p, h_count.data(), sizeof(uint32_t) * (size_t)P, cudaMemcpyHostToDevice, st));
    CB_CUDA(ctx, cudaMemsetAsync(d_maxlen.p, 0, sizeof(uint32_t), st));
    CB_TRY(cb_exclusive_scan_u32_to_i64(ctx, d_rcount.p, d_roff.p, P, nullptr));
    const int wide = ctx->sm_count * 8;
    scatter_by_probe_kernel<<<wide, 256, 0, st>>>(d_rec.p, (unsigned long long)m, d_roff.p, d_sorted.p);
    ctx->launches++;
    CB_TRY(launch_merge(ctx, d_roff.p, d_sorted.p, P, d_nmerged.p, d_maxlen.p));
    CB_CUDA(ctx, cudaGetLastError());
    int64_t n_iv = 0;
    CB_TRY(cb_exclusive_scan_u32_to_i64(ctx, d_nmerged.p, cov->d_iv_off, P, &n_iv));
    CB_CUDA(ctx, cb_dev_alloc(st, (void **)&cov->d_iv, sizeof(uint2) * (size_t)(n_iv ? n_iv : 1)));
    compact_kernel<<<wide, 256, 0, st>>>(d_roff.p, d_sorted.p, cov->d_iv_off, P, cov->d_iv);
    ctx->launches++;
    CB_CUDA(ctx, cudaGetLastError());
    cov->n_intervals = n_iv;
    CB_CUDA(ctx, cudaMemcpyAsync(&cov->max_interval_len, d_maxlen.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    guard.c = nullptr;
    *out = cov;
    return CB_OK;
}
