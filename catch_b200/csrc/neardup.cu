// Near-duplicate filter on the device (K9 signatures, K10 buckets, K11 exact distance, K12 priority
// independent set).
//
// Replaces (reference paths relative to catch/):
//   utils/lsh.py:74-148    MinHashFamily.make_h / h          -> kmer_prepare_kernel + signature_kernel
//   utils/lsh.py:16-45     HammingDistanceFamily             -> hamming_signature_kernel
//   utils/lsh.py:218-300   HashConcatenation, NearNeighborLookup.add -> bucket_kernel (hash-bucket CSR)
//   utils/lsh.py:302-320   NearNeighborLookup.query + filter/near_duplicate_filter.py:107,148-157
//                          (exact Hamming / Jaccard distance)  -> inside decide_probe
//   filter/near_duplicate_filter.py:81-96  sequential greedy in priority order -> rounds of
//                          decide_rounds_kernel (persistent; all rounds of the fix point)
//
// The sequential loop of the reference keeps probe p iff no EARLIER-priority KEPT probe reports p
// as a neighbour (same key in some table and exact distance <= threshold).  That is the
// lexicographically first maximal independent set of the neighbour graph, computed here as a
// fix point: in every round an undecided probe
//   - is dropped if some already-kept earlier bucket-mate is within the distance threshold,
//   - is kept if every earlier bucket-mate is decided and none of the kept ones is that close,
//   - waits otherwise.
// Decisions are final and each one equals the sequential outcome, so the fix point is the
// reference's result; distances are only ever evaluated against kept probes, like the reference.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "internal.cuh"

namespace {

constexpr uint32_t MERSENNE31 = 2147483647u;
constexpr int ND_THREADS = 256;
constexpr int ND_MAX_LEN = CB_MAX_PROBE_LEN;

// ---- CPython str hash (Python/pyhash.c siphash13) with a zero key, as abs() of the signed value
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }
#define SIPROUND_D(v0, v1, v2, v3)                                   \
    do {                                                             \
        v0 += v1; v1 = rotl64(v1, 13); v1 ^= v0; v0 = rotl64(v0, 32); \
        v2 += v3; v3 = rotl64(v3, 16); v3 ^= v2;                      \
        v0 += v3; v3 = rotl64(v3, 21); v3 ^= v0;                      \
        v2 += v1; v1 = rotl64(v1, 17); v1 ^= v2; v2 = rotl64(v2, 32); \
    } while (0)

__device__ __forceinline__ uint64_t abs_pyhash(const uint8_t *s, int len)
{
    uint64_t v0 = 0x736f6d6570736575ull, v1 = 0x646f72616e646f6dull;
    uint64_t v2 = 0x6c7967656e657261ull, v3 = 0x7465646279746573ull;
    uint64_t b = (uint64_t)len << 56;
    int i = 0;
    for (; i + 8 <= len; i += 8) {
        uint64_t mi = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) mi |= (uint64_t)s[i + j] << (8 * j);
        v3 ^= mi;
        SIPROUND_D(v0, v1, v2, v3);
        v0 ^= mi;
    }
    uint64_t t = 0;
    for (int j = 0; i + j < len; j++) t |= (uint64_t)s[i + j] << (8 * j);
    b |= t;
    v3 ^= b;
    SIPROUND_D(v0, v1, v2, v3);
    v0 ^= b;
    v2 ^= 0xff;
    SIPROUND_D(v0, v1, v2, v3);
    SIPROUND_D(v0, v1, v2, v3);
    SIPROUND_D(v0, v1, v2, v3);
    long long h = (long long)((v0 ^ v1) ^ (v2 ^ v3));
    if (h == -1) h = -2;
    return h < 0 ? (uint64_t)0 - (uint64_t)h : (uint64_t)h;
}

__device__ __forceinline__ uint32_t mod_m31(uint64_t v)
{
    // v < 2^62 + 2^31: two folds bring it under 2^32, then one conditional subtract
    v = (v & MERSENNE31) + (v >> 31);
    v = (v & MERSENNE31) + (v >> 31);
    return v >= MERSENNE31 ? (uint32_t)(v - MERSENNE31) : (uint32_t)v;
}

// ---- K9a: per probe, x_i = |hash(kmer_i)| mod p for every k-mer, and the sorted set of distinct
// k-mers (exact codes) used by the Jaccard distance.  One warp per probe.
__global__ void __launch_bounds__(ND_THREADS)
kmer_prepare_kernel(const uint8_t *__restrict__ ascii, const int64_t *__restrict__ off, int64_t n_probes,
                    const int64_t *__restrict__ koff, int kmer, const uint8_t *__restrict__ lut, int cbits,
                    uint32_t *__restrict__ X, uint64_t *__restrict__ kset, uint32_t *__restrict__ kcnt)
{
    __shared__ uint8_t s_seq[ND_THREADS / 32][ND_MAX_LEN];
    __shared__ uint64_t s_code[ND_THREADS / 32][ND_MAX_LEN];
    __shared__ uint8_t s_lut[256];
    s_lut[threadIdx.x & 255] = lut[threadIdx.x & 255];
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    uint8_t *seq = s_seq[wib];
    uint64_t *code = s_code[wib];
    for (int64_t p = warp; p < n_probes; p += n_warps) {
        const int64_t beg = off[p];
        const int len = (int)(off[p + 1] - beg);
        const int nk = len - kmer + 1;
        __syncwarp();
        for (int i = lane; i < len; i += 32) seq[i] = ascii[beg + i];
        __syncwarp();
        const int64_t ko = koff[p];
        for (int i = lane; i < nk; i += 32) {
            X[ko + i] = (uint32_t)(abs_pyhash(seq + i, kmer) % MERSENNE31);
            uint64_t c = 0;
            for (int j = 0; j < kmer; j++) c = (c << cbits) | s_lut[seq[i + j]];
            code[i] = c;
        }
        __syncwarp();
        // bitonic sort of nk codes (normalised network, virtual +inf padding), then unique
        uint32_t n2 = 1;
        while (n2 < (uint32_t)nk) n2 <<= 1;
        for (uint32_t size = 2; size <= n2; size <<= 1) {
            for (uint32_t t = lane; t < n2 / 2; t += 32) {
                const uint32_t blk = t / (size / 2), o = t % (size / 2);
                const uint32_t i = blk * size + o, j = blk * size + size - 1 - o;
                if (j < (uint32_t)nk) {
                    const uint64_t x = code[i], y = code[j];
                    if (x > y) { code[i] = y; code[j] = x; }
                }
            }
            __syncwarp();
            for (uint32_t stride = size / 4; stride >= 1; stride >>= 1) {
                for (uint32_t t = lane; t < n2 / 2; t += 32) {
                    const uint32_t i = (t / stride) * stride * 2 + (t % stride), j = i + stride;
                    if (j < (uint32_t)nk) {
                        const uint64_t x = code[i], y = code[j];
                        if (x > y) { code[i] = y; code[j] = x; }
                    }
                }
                __syncwarp();
            }
        }
        uint32_t n_out = 0;
        for (int base = 0; base < nk; base += 32) {
            const int i = base + lane;
            const bool head = i < nk && (i == 0 || code[i] != code[i - 1]);
            const unsigned heads = __ballot_sync(0xffffffffu, head);
            if (head) kset[ko + n_out + __popc(heads & ((1u << lane) - 1u))] = code[i];
            n_out += __popc(heads);
        }
        if (lane == 0) kcnt[p] = n_out;
    }
}

// ---- K9b: MinHash signatures.  sig[p][f] = min_i (a_f * x_i + b_f) mod (2^31 - 1)
// (utils/lsh.py:91-147 with N = 1).  One warp per probe, lanes hold the x_i, one redux per function.
__global__ void __launch_bounds__(ND_THREADS)
signature_kernel(const uint32_t *__restrict__ X, const int64_t *__restrict__ koff, int64_t n_probes,
                 const uint32_t *__restrict__ pa, const uint32_t *__restrict__ pb, int n_fn,
                 uint32_t *__restrict__ sig)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < n_probes; p += n_warps) {
        const int64_t ko = koff[p];
        const int nk = (int)(koff[p + 1] - ko);
        uint32_t x[ND_MAX_LEN / 32];
#pragma unroll
        for (int r = 0; r < ND_MAX_LEN / 32; r++) {
            const int i = r * 32 + lane;
            x[r] = i < nk ? X[ko + i] : 0xffffffffu;
        }
        for (int f = 0; f < n_fn; f++) {
            const uint64_t a = pa[f] % MERSENNE31, b = pb[f] % MERSENNE31;
            uint32_t best = 0xffffffffu;
#pragma unroll
            for (int r = 0; r < ND_MAX_LEN / 32; r++)
                if (x[r] != 0xffffffffu) best = min(best, mod_m31(a * (uint64_t)x[r] + b));
            best = __reduce_min_sync(0xffffffffu, best);
            if (lane == 0) sig[p * (int64_t)n_fn + f] = best;
        }
    }
}

// Hamming family (utils/lsh.py:16-45): h(x) = x[i]
__global__ void hamming_signature_kernel(const uint8_t *__restrict__ ascii, const int64_t *__restrict__ off,
                                         int64_t n_probes, const int32_t *__restrict__ positions, int n_fn,
                                         uint32_t *__restrict__ sig)
{
    const int64_t n = n_probes * n_fn;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / n_fn;
        const int f = (int)(i % n_fn);
        sig[i] = ascii[off[p] + positions[f]];
    }
}

__device__ __forceinline__ uint64_t key_hash(const uint32_t *key, int k, int t)
{
    uint64_t h = 0x9E3779B97F4A7C15ull * (uint64_t)(t + 1);
    for (int c = 0; c < k; c++) {
        h ^= key[c];
        h *= 0xBF58476D1CE4E5B9ull;
        h ^= h >> 31;
    }
    return h;
}

// ---- K10: bucket of every (probe, table): global bucket id = t * nb + hash(key) % nb
__global__ void bucket_kernel(const uint32_t *__restrict__ sig, int64_t n_probes, int n_tables, int k_concat,
                              uint32_t nb_mask, uint32_t *__restrict__ pg, uint32_t *__restrict__ bsize)
{
    const int64_t n = n_probes * n_tables;
    const int n_fn = n_tables * k_concat;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / n_tables;
        const int t = (int)(i % n_tables);
        const uint64_t h = key_hash(sig + p * n_fn + (int64_t)t * k_concat, k_concat, t);
        const uint32_t b = (uint32_t)t * (nb_mask + 1u) + ((uint32_t)(h >> 20) & nb_mask);
        pg[i] = b;
        atomicAdd(&bsize[b], 1u);
    }
}

struct DecideParams {
    int64_t n_probes;
    int n_tables, k_concat, family;      // family 0 = MinHash/Jaccard, 1 = Hamming
    double dist_thres;
    uint8_t *state;                      // 0 undecided, 1 kept, 2 dropped, 3 kept this round
    const uint32_t *pg;
    const uint32_t *grp_min;
    const int64_t *boff;
    const uint32_t *inc_count;
    const uint32_t *inc_list;
    uint32_t *checked;                   // [n_probes][n_tables] entries of inc_list already examined
    const uint32_t *sig;
    // Jaccard
    const uint64_t *kset;
    const uint32_t *kcnt;
    const int64_t *koff;
    // Hamming
    const uint8_t *ascii;
    const int64_t *off;
    unsigned long long *n_undecided;
    unsigned long long *n_dist;
};

// exact distance between probes p and q, evaluated by the whole warp
__device__ __forceinline__ bool is_near(const DecideParams &D, int64_t p, int64_t q, int lane)
{
    if (D.family == 0) {
        // filter/near_duplicate_filter.py:148-157: 1 - |A & B| / |A | B| over the SETS of k-mers
        const uint64_t *A = D.kset + D.koff[p], *B = D.kset + D.koff[q];
        const int na = (int)D.kcnt[p], nb = (int)D.kcnt[q];
        int inter = 0;
        for (int i = lane; i < na; i += 32) {
            const uint64_t a = A[i];
            int lo = 0, hi = nb;                       // first index with B[idx] >= a
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (B[mid] < a) lo = mid + 1; else hi = mid;
            }
            inter += (lo < nb && B[lo] == a);
        }
        inter = __reduce_add_sync(0xffffffffu, inter);
        const double sim = (double)inter / (double)(na + nb - inter);
        return (1.0 - sim) <= D.dist_thres;
    }
    // filter/near_duplicate_filter.py:107: Probe.mismatches (probe.py:55-64)
    const uint8_t *a = D.ascii + D.off[p], *b = D.ascii + D.off[q];
    const int len = (int)(D.off[p + 1] - D.off[p]);
    int mm = 0;
    for (int i = lane; i < len; i += 32) mm += a[i] != b[i];
    mm = __reduce_add_sync(0xffffffffu, mm);
    return (double)mm <= D.dist_thres;
}

// ---- K11 + K12, one probe of one round, by one warp: 2 = dropped (a kept earlier bucket-mate is within the
// exact distance), 3 = kept (every earlier bucket-mate is decided), 0 = still undecided
__device__ __forceinline__ int decide_probe(const DecideParams &D, int64_t p, int lane, unsigned long long &dists)
{
    const int n_fn = D.n_tables * D.k_concat;
    bool blocked = false, dropped = false;
    for (int tb = 0; tb < D.n_tables && !dropped; tb += 32) {
        const int t = tb + lane;
        const bool have_t = t < D.n_tables;
        uint32_t b = 0, j = 0, jn = 0;
        if (have_t) {
            b = D.pg[p * D.n_tables + t];
            if (__ldcg(&D.grp_min[b]) < (uint32_t)p) blocked = true;
            j = __ldcg(&D.checked[p * D.n_tables + t]);
            jn = __ldcg(&D.inc_count[b]);
        }
        // walk the kept members of the probe's buckets, 32 tables abreast
        while (!dropped) {
            uint32_t q = 0xffffffffu;
            bool cand = false;
            if (have_t && j < jn) {
                q = __ldcg(&D.inc_list[D.boff[b] + j]);
                j++;
                if (q < (uint32_t)p) {             // kept earlier; same key (not just same bucket)?
                    cand = true;
                    const uint32_t *kp = D.sig + p * n_fn + (int64_t)t * D.k_concat;
                    const uint32_t *kq = D.sig + (int64_t)q * n_fn + (int64_t)t * D.k_concat;
                    for (int c = 0; c < D.k_concat; c++) cand = cand && (kp[c] == kq[c]);
                }
            }
            const unsigned more = __ballot_sync(0xffffffffu, have_t && j < jn);
            unsigned cands = __ballot_sync(0xffffffffu, cand);
            // the same q usually shows up in many tables at once: evaluate it once
            const unsigned same = __match_any_sync(0xffffffffu, cand ? q : 0xffffffffu - lane);
            if (cand && (__ffs(same) - 1) != lane) cand = false;
            cands = __ballot_sync(0xffffffffu, cand);
            while (cands && !dropped) {
                const int src = __ffs(cands) - 1;
                cands &= cands - 1;
                const uint32_t qq = __shfl_sync(0xffffffffu, q, src);
                dists++;
                if (is_near(D, p, (int64_t)qq, lane)) dropped = true;
            }
            if (!more) break;
        }
        if (have_t) __stcg(&D.checked[p * D.n_tables + t], j);
    }
    blocked = __any_sync(0xffffffffu, blocked);
    return dropped ? 2 : (blocked ? 0 : 3);
}

// The sequential keep/drop loop of filter/near_duplicate_filter.py:81-96 as rounds to its fix point, ALL rounds
// in one persistent cooperative kernel.  A round only touches the probes that are still undecided (compact
// list, rebuilt every round), so the cost follows the number of undecided probes, not rounds x probes:
//   A  every undecided probe writes its index into grp_min of its buckets (atomicMin);
//   B  one warp per undecided probe decides: dropped / kept / still undecided (-> next list);
//   C  probes kept this round join the kept lists of their buckets; grp_min of the touched buckets is reset.
struct RoundsCtl {
    uint32_t *list[2];            // undecided probes, ping-pong
    uint32_t *n_list;             // [2]
    uint32_t *kept;               // probes kept in the current round
    uint32_t *n_kept;
    uint32_t *grp_min;
    uint32_t *inc_count;
    uint32_t *inc_list;
    unsigned long long *barrier;
    unsigned long long *rounds;
};

__device__ __forceinline__ void nd_grid_barrier(unsigned long long *counter, unsigned long long &target)
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

__global__ void __launch_bounds__(ND_THREADS)
decide_rounds_kernel(const DecideParams D, const RoundsCtl C)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long bar = 0, dists = 0, rounds = 0;
    int cur = 0;
    uint32_t n_cur = (uint32_t)D.n_probes;
    for (;;) {
        const uint32_t *list = C.list[cur];
        // A: smallest undecided index of every bucket
        if (gtid == 0) *C.n_kept = 0u;
        for (int64_t i = warp; i < (int64_t)n_cur; i += n_warps) {
            const uint32_t p = list[i];
            for (int t = lane; t < D.n_tables; t += 32) atomicMin(&C.grp_min[D.pg[(int64_t)p * D.n_tables + t]], p);
        }
        nd_grid_barrier(C.barrier, bar);
        // B: decide
        for (int64_t i = warp; i < (int64_t)n_cur; i += n_warps) {
            const uint32_t p = list[i];
            const int r = decide_probe(D, (int64_t)p, lane, dists);
            if (lane == 0) {
                if (r == 2) D.state[p] = 2;
                else if (r == 3) C.kept[atomicAdd(C.n_kept, 1u)] = p;
                else C.list[cur ^ 1][atomicAdd(&C.n_list[cur ^ 1], 1u)] = p;
            }
        }
        nd_grid_barrier(C.barrier, bar);
        // C: kept probes join the kept lists of their buckets; grp_min entries used this round are reset
        const uint32_t n_kept = __ldcg(C.n_kept);
        for (int64_t i = warp; i < (int64_t)n_kept; i += n_warps) {
            const uint32_t p = __ldcg(&C.kept[i]);
            for (int t = lane; t < D.n_tables; t += 32) {
                const uint32_t b = D.pg[(int64_t)p * D.n_tables + t];
                const uint32_t slot = atomicAdd(&C.inc_count[b], 1u);
                C.inc_list[D.boff[b] + slot] = p;
            }
            if (lane == 0) D.state[p] = 1;
        }
        for (int64_t i = warp; i < (int64_t)n_cur; i += n_warps) {
            const uint32_t p = list[i];
            for (int t = lane; t < D.n_tables; t += 32) C.grp_min[D.pg[(int64_t)p * D.n_tables + t]] = 0xffffffffu;
        }
        if (gtid == 0) C.n_list[cur] = 0u;
        rounds++;
        nd_grid_barrier(C.barrier, bar);
        cur ^= 1;
        n_cur = __ldcg(&C.n_list[cur]);
        if (n_cur == 0u) break;
    }
    if (lane == 0 && dists) atomicAdd(D.n_dist, dists);
    if (gtid == 0) *C.rounds = rounds;
}

__global__ void iota_kernel(uint32_t *out, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (uint32_t)i;
}

__global__ void finish_kernel(const uint8_t *__restrict__ state, int64_t n, uint8_t *__restrict__ keep)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        keep[i] = state[i] == 1;
}

// Host-side checks shared by the entry points: per-probe lengths -> relative offsets and k-mer
// offsets; symbol table for the exact k-mer codes of the MinHash family.
int neardup_tables(cb_ctx *ctx, int64_t P, int family, int kmer, const int32_t *positions, int n_fn,
                   const std::vector<int64_t> &h_off, std::vector<int64_t> &h_koff, const bool present[256],
                   uint8_t lut[256], int &cbits)
{
    h_koff.assign((size_t)P + 1, 0);
    const int L0 = P ? (int)(h_off[1] - h_off[0]) : 0;
    for (int64_t p = 0; p < P; p++) {
        const int64_t len = h_off[(size_t)p + 1] - h_off[(size_t)p];
        if (len < 0 || len > ND_MAX_LEN) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "probe length outside [0, 256]");
        if (family == 0) {
            if (len < kmer) return cb_fail(ctx, CB_ERR_ARG, "kmer_size exceeds a probe's length (utils/lsh.py:117)");
            h_koff[(size_t)p + 1] = h_koff[(size_t)p] + (len - kmer + 1);
        } else if (len != L0) {
            return cb_fail(ctx, CB_ERR_ARG, "Hamming family needs probes of one length (utils/lsh.py:30)");
        }
    }
    memset(lut, 0, 256);
    cbits = 1;
    if (family == 0) {
        if (kmer < 1) return cb_fail(ctx, CB_ERR_ARG, "kmer_size must be positive");
        int n_sym = 0;
        for (int c = 0; c < 256; c++) if (present[c]) lut[c] = (uint8_t)n_sym++;
        while ((1 << cbits) < n_sym) cbits++;
        if ((int64_t)cbits * kmer > 64)
            return cb_fail(ctx, CB_ERR_UNSUPPORTED, "kmer_size * bits-per-symbol exceeds 64 (exact k-mer codes)");
    } else {
        for (int f = 0; f < n_fn; f++)
            if (positions[f] < 0 || positions[f] >= L0) return cb_fail(ctx, CB_ERR_ARG, "sampled position out of range");
    }
    return CB_OK;
}

// The device pipeline on P distinct probes in priority order whose bytes are already on the device
// (d_ascii, relative offsets h_off).  keep[i] (host) = 1 iff probe i is kept.
int neardup_run(cb_ctx *ctx, const uint8_t *d_ascii_in, const std::vector<int64_t> &h_off,
                const std::vector<int64_t> &h_koff, const uint8_t lut[256], int cbits, int64_t P, int family,
                const uint32_t *pa, const uint32_t *pb, const int32_t *positions, int n_tables, int k_concat,
                int kmer, double dist_thres, uint8_t *keep, cb_stats *stats)
{
    cudaStream_t st = ctx->stream;
    const int n_fn = n_tables * k_concat;
    const int wide = ctx->sm_count * 8;
    EventTimer t_all(st), t_sig(st), t_rounds(st);
    t_all.start();
    struct { const uint8_t *p; } d_ascii{d_ascii_in};
    DevBuf<uint8_t> d_lut, d_state, d_keep;
    DevBuf<int64_t> d_off, d_koff, d_boff;
    DevBuf<uint32_t> d_X, d_kcnt, d_pa, d_pb, d_sig, d_pg, d_bsize, d_grpmin, d_inccount, d_inclist, d_checked;
    DevBuf<int32_t> d_pos;
    DevBuf<uint64_t> d_kset;
    DevBuf<unsigned long long> d_ctr;
    CB_CUDA(ctx, d_off.alloc((size_t)P + 1));
    CB_CUDA(ctx, d_sig.alloc((size_t)P * n_fn));
    CB_CUDA(ctx, d_ctr.alloc(2));
    CB_CUDA(ctx, cudaMemcpyAsync(d_off.p, h_off.data(), sizeof(int64_t) * (size_t)(P + 1), cudaMemcpyHostToDevice, st));

    t_sig.start();
    if (family == 0) {
        const int64_t NK = h_koff[(size_t)P];
        CB_CUDA(ctx, d_koff.alloc((size_t)P + 1));
        CB_CUDA(ctx, d_X.alloc((size_t)NK));
        CB_CUDA(ctx, d_kset.alloc((size_t)NK));
        CB_CUDA(ctx, d_kcnt.alloc((size_t)P));
        CB_CUDA(ctx, d_lut.alloc(256));
        CB_CUDA(ctx, d_pa.alloc((size_t)n_fn));
        CB_CUDA(ctx, d_pb.alloc((size_t)n_fn));
        CB_CUDA(ctx, cudaMemcpyAsync(d_koff.p, h_koff.data(), sizeof(int64_t) * (size_t)(P + 1), cudaMemcpyHostToDevice, st));
        CB_CUDA(ctx, cudaMemcpyAsync(d_lut.p, lut, 256, cudaMemcpyHostToDevice, st));
        CB_CUDA(ctx, cudaMemcpyAsync(d_pa.p, pa, sizeof(uint32_t) * (size_t)n_fn, cudaMemcpyHostToDevice, st));
        CB_CUDA(ctx, cudaMemcpyAsync(d_pb.p, pb, sizeof(uint32_t) * (size_t)n_fn, cudaMemcpyHostToDevice, st));
        kmer_prepare_kernel<<<wide, ND_THREADS, 0, st>>>(d_ascii.p, d_off.p, P, d_koff.p, kmer, d_lut.p, cbits,
                                                         d_X.p, d_kset.p, d_kcnt.p);
        signature_kernel<<<wide, ND_THREADS, 0, st>>>(d_X.p, d_koff.p, P, d_pa.p, d_pb.p, n_fn, d_sig.p);
        ctx->launches += 2;
    } else {
        CB_CUDA(ctx, d_pos.alloc((size_t)n_fn));
        CB_CUDA(ctx, cudaMemcpyAsync(d_pos.p, positions, sizeof(int32_t) * (size_t)n_fn, cudaMemcpyHostToDevice, st));
        hamming_signature_kernel<<<wide, ND_THREADS, 0, st>>>(d_ascii.p, d_off.p, P, d_pos.p, n_fn, d_sig.p);
        ctx->launches++;
    }
    CB_CUDA(ctx, cudaGetLastError());

    // K10 buckets
    int64_t nb = 64;
    while (nb < 2 * P) nb <<= 1;
    const int64_t n_buckets = nb * n_tables;
    if (n_buckets >= 0xffffffffll) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "too many hash buckets");
    CB_CUDA(ctx, d_pg.alloc((size_t)P * n_tables));
    CB_CUDA(ctx, d_bsize.alloc((size_t)n_buckets));
    CB_CUDA(ctx, d_boff.alloc((size_t)n_buckets + 1));
    CB_CUDA(ctx, d_grpmin.alloc((size_t)n_buckets));
    CB_CUDA(ctx, d_inccount.alloc((size_t)n_buckets));
    CB_CUDA(ctx, d_inclist.alloc((size_t)P * n_tables));
    CB_CUDA(ctx, d_checked.alloc((size_t)P * n_tables));
    CB_CUDA(ctx, d_state.alloc((size_t)P));
    CB_CUDA(ctx, d_keep.alloc((size_t)P));
    CB_CUDA(ctx, cudaMemsetAsync(d_bsize.p, 0, sizeof(uint32_t) * (size_t)n_buckets, st));
    CB_CUDA(ctx, cudaMemsetAsync(d_inccount.p, 0, sizeof(uint32_t) * (size_t)n_buckets, st));
    CB_CUDA(ctx, cudaMemsetAsync(d_checked.p, 0, sizeof(uint32_t) * (size_t)P * n_tables, st));
    CB_CUDA(ctx, cudaMemsetAsync(d_state.p, 0, (size_t)P, st));
    CB_CUDA(ctx, cudaMemsetAsync(d_ctr.p, 0, sizeof(unsigned long long) * 2, st));
    bucket_kernel<<<wide, ND_THREADS, 0, st>>>(d_sig.p, P, n_tables, k_concat, (uint32_t)(nb - 1), d_pg.p, d_bsize.p);
    ctx->launches++;
    CB_CUDA(ctx, cudaGetLastError());
    CB_TRY(cb_exclusive_scan_u32_to_i64(ctx, d_bsize.p, d_boff.p, n_buckets, nullptr));
    t_sig.stop();

    DecideParams D;
    memset(&D, 0, sizeof D);
    D.n_probes = P;
    D.n_tables = n_tables;
    D.k_concat = k_concat;
    D.family = family;
    D.dist_thres = dist_thres;
    D.state = d_state.p;
    D.pg = d_pg.p;
    D.grp_min = d_grpmin.p;
    D.boff = d_boff.p;
    D.inc_count = d_inccount.p;
    D.inc_list = d_inclist.p;
    D.checked = d_checked.p;
    D.sig = d_sig.p;
    D.kset = d_kset.p;
    D.kcnt = d_kcnt.p;
    D.koff = d_koff.p;
    D.ascii = d_ascii.p;
    D.off = d_off.p;
    D.n_undecided = d_ctr.p;
    D.n_dist = d_ctr.p + 1;

    // K12 rounds: one persistent cooperative kernel
    t_rounds.start();
    DevBuf<uint32_t> d_lists, d_small;
    DevBuf<unsigned long long> d_bar;
    CB_CUDA(ctx, d_lists.alloc((size_t)P * 3));                 // undecided (ping-pong) + kept-this-round
    CB_CUDA(ctx, d_small.alloc(8));
    CB_CUDA(ctx, d_bar.alloc(2));
    CB_CUDA(ctx, cudaMemsetAsync(d_small.p, 0, sizeof(uint32_t) * 8, st));
    CB_CUDA(ctx, cudaMemsetAsync(d_bar.p, 0, sizeof(unsigned long long) * 2, st));
    CB_CUDA(ctx, cudaMemsetAsync(d_grpmin.p, 0xff, sizeof(uint32_t) * (size_t)n_buckets, st));
    iota_kernel<<<wide, ND_THREADS, 0, st>>>(d_lists.p, P);
    ctx->launches++;
    RoundsCtl C;
    C.list[0] = d_lists.p;
    C.list[1] = d_lists.p + P;
    C.kept = d_lists.p + 2 * P;
    C.n_list = d_small.p;
    C.n_kept = d_small.p + 2;
    C.grp_min = d_grpmin.p;
    C.inc_count = d_inccount.p;
    C.inc_list = d_inclist.p;
    C.barrier = d_bar.p;
    C.rounds = d_bar.p + 1;
    int per_sm = 0;
    CB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decide_rounds_kernel, ND_THREADS, 0));
    if (per_sm < 1) return cb_fail(ctx, CB_ERR_CUDA, "near-duplicate kernel does not fit on an SM");
    if (per_sm > 4) per_sm = 4;
    void *args[] = {(void *)&D, (void *)&C};
    CB_CUDA(ctx, cudaLaunchCooperativeKernel((void *)decide_rounds_kernel, dim3(per_sm * ctx->sm_count), dim3(ND_THREADS),
                                             args, 0, st));
    ctx->launches++;
    unsigned long long h_ctr[2] = {0, 0}, h_rounds = 0;
    CB_CUDA(ctx, cudaMemcpyAsync(h_ctr, d_ctr.p, sizeof h_ctr, cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaMemcpyAsync(&h_rounds, d_bar.p + 1, sizeof h_rounds, cudaMemcpyDeviceToHost, st));
    const int64_t rounds = 0;
    finish_kernel<<<wide, ND_THREADS, 0, st>>>(d_state.p, P, d_keep.p);
    ctx->launches++;
    t_rounds.stop();
    t_all.stop();
    CB_CUDA(ctx, cudaMemcpyAsync(keep, d_keep.p, (size_t)P, cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    if (stats) {
        stats->ms_seed_index += t_sig.ms();      // signatures + buckets
        stats->ms_greedy += t_rounds.ms();       // decision rounds
        stats->ms_total += t_all.ms();
        stats->n_picks = (int64_t)h_rounds + rounds;
        stats->n_candidate_hits = (int64_t)h_ctr[1];   // exact distance evaluations
        stats->n_kernel_launches = ctx->launches;
    }
    return CB_OK;
}

// cb_minhash_neardup / cb_hamming_neardup: the caller passes the DISTINCT probes in priority order.
int neardup_common(cb_ctx *ctx, const uint8_t *ascii, const int64_t *probe_off, int64_t P, int family,
                   const uint32_t *pa, const uint32_t *pb, const int32_t *positions, int n_tables, int k_concat,
                   int kmer, double dist_thres, uint8_t *keep, cb_stats *stats)
{
    if (P < 0 || n_tables < 1 || k_concat < 1 || !keep) return cb_fail(ctx, CB_ERR_ARG, "bad argument");
    if (P == 0) return CB_OK;
    if (!ascii || !probe_off) return cb_fail(ctx, CB_ERR_ARG, "null probe table");
    if (P >= 0xfffffff0ll) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "too many probes");
    const int64_t base = probe_off[0];
    const int64_t total = probe_off[P] - base;
    std::vector<int64_t> h_off((size_t)P + 1), h_koff;
    for (int64_t p = 0; p <= P; p++) h_off[(size_t)p] = probe_off[p] - base;
    bool present[256] = {false};
    if (family == 0)
        for (int64_t i = 0; i < total; i++) present[ascii[base + i]] = true;
    uint8_t lut[256];
    int cbits = 1;
    CB_TRY(neardup_tables(ctx, P, family, kmer, positions, n_tables * k_concat, h_off, h_koff, present, lut, cbits));
    DevBuf<uint8_t> d_ascii;
    CB_CUDA(ctx, d_ascii.alloc((size_t)total));
    CB_CUDA(ctx, cudaMemcpyAsync(d_ascii.p, ascii + base, (size_t)total, cudaMemcpyHostToDevice, ctx->stream));
    return neardup_run(ctx, d_ascii.p, h_off, h_koff, lut, cbits, P, family, pa, pb, positions, n_tables, k_concat,
                       kmer, dist_thres, keep, stats);
}


// ---------------------------------------------------------------------------------------
// Exact-duplicate grouping on the device (filter/near_duplicate_filter.py:61-66: occurrences[p] += 1
// over Probe objects that hash and compare by sequence).  Probes are packed to bit planes first, so
// two probes are the same sequence iff their lengths and packed words are equal.
// An open-addressing table holds, per distinct sequence, the SMALLEST list index carrying it:
// a probe walks its probe sequence until it finds a free slot (claims it) or a slot whose occupant
// is the same sequence (atomicMin of the index).  Slots never change their sequence class, so all
// probes of a class end in the same slot whatever the interleaving.
// ---------------------------------------------------------------------------------------
constexpr uint32_t GRP_EMPTY = 0xffffffffu;

__device__ __forceinline__ uint64_t hash_packed(const uint64_t *w, int wpp, int len)
{
    uint64_t h = 0x9E3779B97F4A7C15ull ^ (uint64_t)len;
    for (int i = 0; i < wpp; i++) {
        h ^= w[i];
        h *= 0xBF58476D1CE4E5B9ull;
        h ^= h >> 31;
    }
    return h;
}

__device__ __forceinline__ bool same_packed(const uint64_t *words, const int32_t *lens, int wpp, int64_t a, int64_t b)
{
    if (lens[a] != lens[b]) return false;
    const uint64_t *x = words + a * wpp, *y = words + b * wpp;
    for (int i = 0; i < wpp; i++)
        if (x[i] != y[i]) return false;
    return true;
}

template <bool INSERT>
__global__ void group_kernel(const uint64_t *__restrict__ words, const int32_t *__restrict__ lens, int64_t n, int wpp,
                             uint32_t *table, uint32_t mask, uint32_t *__restrict__ count, uint32_t *__restrict__ is_rep)
{
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        uint32_t slot = (uint32_t)(hash_packed(words + p * wpp, wpp, lens[p]) >> 17) & mask;
        for (;;) {
            uint32_t cur = __ldcg(&table[slot]);
            if (INSERT && cur == GRP_EMPTY) {
                cur = atomicCAS(&table[slot], GRP_EMPTY, (uint32_t)p);
                if (cur == GRP_EMPTY) break;                       // claimed
            }
            if (cur != GRP_EMPTY && same_packed(words, lens, wpp, p, (int64_t)cur)) {
                if (INSERT) atomicMin(&table[slot], (uint32_t)p);
                else {
                    atomicAdd(&count[cur], 1u);                    // cur is the class's first occurrence now
                    is_rep[p] = cur == (uint32_t)p ? 1u : 0u;
                }
                break;
            }
            slot = (slot + 1) & mask;
        }
    }
}

__global__ void group_compact_kernel(const uint32_t *__restrict__ is_rep, const int64_t *__restrict__ pos,
                                     const uint32_t *__restrict__ count, int64_t n, uint32_t *__restrict__ first_idx,
                                     uint32_t *__restrict__ mult)
{
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x)
        if (is_rep[p]) {
            first_idx[pos[p]] = (uint32_t)p;
            mult[pos[p]] = count[p];
        }
}

// bytes of the chosen probes, back to back in the given order: one warp per probe
__global__ void gather_bytes_kernel(const uint8_t *__restrict__ src, const int64_t *__restrict__ src_off,
                                    const int64_t *__restrict__ dst_off, int64_t n, uint8_t *__restrict__ dst)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < n; p += n_warps) {
        const int64_t a = src_off[p], b = dst_off[p];
        const int len = (int)(dst_off[p + 1] - b);
        for (int i = lane; i < len; i += 32) dst[b + i] = src[a + i];
    }
}

}  // namespace

// Identical sequences of one probe list grouped on the device.  On return `first`/`mult` hold, for every
// distinct sequence in order of first occurrence, the list index of that occurrence and the number
// of occurrences; the bytes stay on the device for the caller.
struct GroupResult {
    DevBuf<uint8_t> d_ascii;
    std::vector<int64_t> rel;              // [n+1] offsets relative to the first probe
    std::vector<uint32_t> first, mult;     // [n_distinct]
    bool present[256];
    double ms = 0;
};

static int group_identical(cb_ctx *ctx, const uint8_t *ascii, const int64_t *probe_off, int64_t n, GroupResult &R)
{
    cudaStream_t st = ctx->stream;
    const int wide = ctx->sm_count * 8;
    const int64_t base = probe_off[0], total = probe_off[n] - base;
    int max_len = 0;
    R.rel.resize((size_t)n + 1);
    for (int64_t i = 0; i <= n; i++) R.rel[(size_t)i] = probe_off[i] - base;
    for (int64_t i = 0; i < n; i++) {
        const int64_t len = R.rel[(size_t)i + 1] - R.rel[(size_t)i];
        if (len < 0 || len > ND_MAX_LEN) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "probe length outside [0, 256]");
        if (len > max_len) max_len = (int)len;
    }
    EventTimer t_grp(st);
    t_grp.start();
    // ---- bytes and offsets to the device, symbol table, bit planes
    DevBuf<uint8_t> d_lut;
    DevBuf<int64_t> d_off, d_pos;
    DevBuf<uint32_t> d_present, d_table, d_count, d_isrep, d_first, d_mult;
    DevBuf<uint64_t> d_words;
    DevBuf<int32_t> d_len;
    CB_CUDA(ctx, R.d_ascii.alloc((size_t)total));
    CB_CUDA(ctx, d_off.alloc((size_t)n + 1));
    CB_CUDA(ctx, d_present.alloc(256));
    CB_CUDA(ctx, cudaMemsetAsync(d_present.p, 0, sizeof(uint32_t) * 256, st));
    if (total) CB_CUDA(ctx, cudaMemcpyAsync(R.d_ascii.p, ascii + base, (size_t)total, cudaMemcpyHostToDevice, st));
    CB_CUDA(ctx, cudaMemcpyAsync(d_off.p, R.rel.data(), sizeof(int64_t) * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
    CB_TRY(cb_launch_byte_presence(ctx, R.d_ascii.p, total, d_present.p));
    uint32_t h_present[256];
    CB_CUDA(ctx, cudaMemcpyAsync(h_present, d_present.p, sizeof h_present, cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    uint8_t pack_lut[256];
    memset(pack_lut, 0, sizeof pack_lut);
    int n_sym = 0;
    for (int c = 0; c < 256; c++) {
        R.present[c] = h_present[c] != 0;
        if (R.present[c]) pack_lut[c] = (uint8_t)n_sym++;
    }
    int bits = 1;
    while ((1 << bits) < n_sym) bits++;
    const int nw = max_len ? (max_len + 63) / 64 : 1, wpp = bits * nw;
    CB_CUDA(ctx, d_lut.alloc(256));
    CB_CUDA(ctx, cudaMemcpyAsync(d_lut.p, pack_lut, 256, cudaMemcpyHostToDevice, st));
    CB_CUDA(ctx, d_words.alloc((size_t)n * (size_t)wpp));
    CB_CUDA(ctx, d_len.alloc((size_t)n));
    CB_TRY(cb_launch_pack_probes(ctx, R.d_ascii.p, d_off.p, 0, n, d_lut.p, bits, nw, d_words.p, d_len.p));
    // ---- grouping
    int64_t cap = 1024;
    while (cap < 2 * n) cap <<= 1;
    CB_CUDA(ctx, d_table.alloc((size_t)cap));
    CB_CUDA(ctx, d_count.alloc((size_t)n));
    CB_CUDA(ctx, d_isrep.alloc((size_t)n));
    CB_CUDA(ctx, d_pos.alloc((size_t)n + 1));
    CB_CUDA(ctx, cudaMemsetAsync(d_table.p, 0xff, sizeof(uint32_t) * (size_t)cap, st));
    CB_CUDA(ctx, cudaMemsetAsync(d_count.p, 0, sizeof(uint32_t) * (size_t)n, st));
    group_kernel<true><<<wide, ND_THREADS, 0, st>>>(d_words.p, d_len.p, n, wpp, d_table.p, (uint32_t)(cap - 1), nullptr, nullptr);
    group_kernel<false><<<wide, ND_THREADS, 0, st>>>(d_words.p, d_len.p, n, wpp, d_table.p, (uint32_t)(cap - 1), d_count.p, d_isrep.p);
    ctx->launches += 2;
    CB_CUDA(ctx, cudaGetLastError());
    int64_t D = 0;
    CB_TRY(cb_exclusive_scan_u32_to_i64(ctx, d_isrep.p, d_pos.p, n, &D));
    CB_CUDA(ctx, d_first.alloc((size_t)D));
    CB_CUDA(ctx, d_mult.alloc((size_t)D));
    group_compact_kernel<<<wide, ND_THREADS, 0, st>>>(d_isrep.p, d_pos.p, d_count.p, n, d_first.p, d_mult.p);
    ctx->launches++;
    CB_CUDA(ctx, cudaGetLastError());
    R.first.resize((size_t)D);
    R.mult.resize((size_t)D);
    CB_CUDA(ctx, cudaMemcpyAsync(R.first.data(), d_first.p, sizeof(uint32_t) * (size_t)D, cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaMemcpyAsync(R.mult.data(), d_mult.p, sizeof(uint32_t) * (size_t)D, cudaMemcpyDeviceToHost, st));
    t_grp.stop();
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    R.ms = t_grp.ms();
    return CB_OK;
}

// DuplicateFilter on the device (filter/duplicate_filter.py:20-26: list(OrderedDict.fromkeys(input))):
// first occurrence of every distinct sequence, in list order, with its multiplicity.
int cb_group_duplicates_impl(cb_ctx *ctx, const uint8_t *ascii, const int64_t *probe_off, int64_t n,
                             int64_t *first_idx, int32_t *count, int64_t *n_distinct, cb_stats *stats)
{
    if (n < 0 || !n_distinct || (n > 0 && !first_idx)) return cb_fail(ctx, CB_ERR_ARG, "bad argument");
    *n_distinct = 0;
    if (n == 0) return CB_OK;
    if (!ascii || !probe_off) return cb_fail(ctx, CB_ERR_ARG, "null probe table");
    if (n >= 0xfffffff0ll) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "too many probes");
    GroupResult R;
    CB_TRY(group_identical(ctx, ascii, probe_off, n, R));
    const int64_t D = (int64_t)R.first.size();
    for (int64_t i = 0; i < D; i++) {
        first_idx[i] = (int64_t)R.first[(size_t)i];
        if (count) count[i] = (int32_t)R.mult[(size_t)i];
    }
    *n_distinct = D;
    if (stats) {
        stats->ms_pack = R.ms;
        stats->ms_total = R.ms;
        stats->n_intervals = D;
        stats->n_kernel_launches = ctx->launches;
    }
    return CB_OK;
}

// Whole near-duplicate filter for one probe list WITH its duplicates, in list order: grouping of
// identical sequences, priority order (multiplicity descending, first occurrence ascending --
// Python's stable sorted(..., reverse=True) over a dict in insertion order), LSH filter.
int cb_neardup_filter_impl(cb_ctx *ctx, const uint8_t *ascii, const int64_t *probe_off, int64_t n, int32_t family,
                           const uint32_t *pa, const uint32_t *pb, const int32_t *positions, int32_t n_tables,
                           int32_t k_concat, int32_t kmer, double dist_thres, int64_t *kept_first_idx,
                           int64_t *n_kept, int64_t *n_distinct_out, cb_stats *stats)
{
    if (n < 0 || n_tables < 1 || k_concat < 1 || !kept_first_idx || !n_kept) return cb_fail(ctx, CB_ERR_ARG, "bad argument");
    *n_kept = 0;
    if (n_distinct_out) *n_distinct_out = 0;
    if (n == 0) return CB_OK;
    if (!ascii || !probe_off) return cb_fail(ctx, CB_ERR_ARG, "null probe table");
    if (n >= 0xfffffff0ll) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "too many probes");
    if (family == 0 ? (!pa || !pb) : !positions) return cb_fail(ctx, CB_ERR_ARG, "null hash parameters");
    cudaStream_t st = ctx->stream;
    const int wide = ctx->sm_count * 8;
    GroupResult R;
    CB_TRY(group_identical(ctx, ascii, probe_off, n, R));
    const int64_t D = (int64_t)R.first.size();
    const std::vector<uint32_t> &h_first = R.first, &h_mult = R.mult;
    const std::vector<int64_t> &rel = R.rel;
    const bool *present = R.present;
    DevBuf<uint8_t> &d_ascii = R.d_ascii;
    EventTimer t_grp(st);
    t_grp.start();
    if (n_distinct_out) *n_distinct_out = D;
    // ---- priority order: multiplicity descending, first occurrence ascending (the distinct entries
    // arrive in first-occurrence order, so a stable counting sort by multiplicity does it)
    uint32_t max_mult = 0;
    for (int64_t i = 0; i < D; i++) max_mult = std::max(max_mult, h_mult[(size_t)i]);
    std::vector<int64_t> start((size_t)max_mult + 2, 0);
    for (int64_t i = 0; i < D; i++) start[(size_t)(max_mult - h_mult[(size_t)i]) + 1]++;
    for (size_t c = 1; c < start.size(); c++) start[c] += start[c - 1];
    std::vector<uint32_t> ordered((size_t)D);                      // list index of the k-th probe in priority order
    for (int64_t i = 0; i < D; i++) ordered[(size_t)start[(size_t)(max_mult - h_mult[(size_t)i])]++] = h_first[(size_t)i];
    std::vector<int64_t> h_src((size_t)D), h_off((size_t)D + 1), h_koff;
    h_off[0] = 0;
    for (int64_t k = 0; k < D; k++) {
        const uint32_t i = ordered[(size_t)k];
        h_src[(size_t)k] = rel[i];
        h_off[(size_t)k + 1] = h_off[(size_t)k] + (rel[(size_t)i + 1] - rel[i]);
    }
    uint8_t lut[256];
    int cbits = 1;
    CB_TRY(neardup_tables(ctx, D, family, kmer, positions, n_tables * k_concat, h_off, h_koff, present, lut, cbits));
    DevBuf<uint8_t> d_ord;
    DevBuf<int64_t> d_src, d_dst;
    CB_CUDA(ctx, d_ord.alloc((size_t)h_off[(size_t)D]));
    CB_CUDA(ctx, d_src.alloc((size_t)D));
    CB_CUDA(ctx, d_dst.alloc((size_t)D + 1));
    CB_CUDA(ctx, cudaMemcpyAsync(d_src.p, h_src.data(), sizeof(int64_t) * (size_t)D, cudaMemcpyHostToDevice, st));
    CB_CUDA(ctx, cudaMemcpyAsync(d_dst.p, h_off.data(), sizeof(int64_t) * (size_t)(D + 1), cudaMemcpyHostToDevice, st));
    gather_bytes_kernel<<<wide, ND_THREADS, 0, st>>>(d_ascii.p, d_src.p, d_dst.p, D, d_ord.p);
    ctx->launches++;
    CB_CUDA(ctx, cudaGetLastError());
    t_grp.stop();
    std::vector<uint8_t> keep((size_t)D, 0);
    CB_TRY(neardup_run(ctx, d_ord.p, h_off, h_koff, lut, cbits, D, family, pa, pb, positions, n_tables, k_concat, kmer,
                       dist_thres, keep.data(), stats));
    int64_t nk = 0;
    for (int64_t k = 0; k < D; k++)
        if (keep[(size_t)k]) kept_first_idx[nk++] = (int64_t)ordered[(size_t)k];
    *n_kept = nk;
    if (stats) {
        stats->ms_pack = R.ms + t_grp.ms();      // upload, packing, grouping; ordering, gather
        stats->ms_total += stats->ms_pack;
        stats->n_intervals = D;                  // distinct sequences
    }
    return CB_OK;
}

int cb_minhash_neardup_impl(cb_ctx *ctx, const uint8_t *ascii, const int64_t *probe_off, int64_t n_probes,
                            const uint32_t *a, const uint32_t *b, int32_t n_tables, int32_t k_concat,
                            int32_t kmer_size, double dist_thres, uint8_t *keep, cb_stats *stats)
{
    if (n_probes > 0 && (!a || !b)) return cb_fail(ctx, CB_ERR_ARG, "null hash parameters");
    return neardup_common(ctx, ascii, probe_off, n_probes, 0, a, b, nullptr, n_tables, k_concat, kmer_size,
                          dist_thres, keep, stats);
}

int cb_hamming_neardup_impl(cb_ctx *ctx, const uint8_t *ascii, const int64_t *probe_off, int64_t n_probes,
                            const int32_t *positions, int32_t n_tables, int32_t k_concat, int32_t dist_thres,
                            uint8_t *keep, cb_stats *stats)
{
    if (n_probes > 0 && !positions) return cb_fail(ctx, CB_ERR_ARG, "null positions");
    return neardup_common(ctx, ascii, probe_off, n_probes, 1, nullptr, nullptr, positions, n_tables, k_concat, 0,
                          (double)dist_thres, keep, stats);
}
