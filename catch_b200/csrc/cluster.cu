// Genome clustering by MinHash sketches: the sketches and every pairwise distance estimate
// on the device.
//
// Replaces (reference paths relative to catch/):
//   utils/cluster.py:29-46   make_signatures_with_minhash            -> sketch_hash_kernel + sketch_select_kernel
//   utils/lsh.py:74-148      MinHashFamily(kmer_size, N).make_h()/h with the md5 inner hash
//                            (use_fast_str_hash=False, :106-111)
//   utils/lsh.py:166-214     MinHashFamily.estimate_jaccard_dist       -> sketch_merge (one thread per pair)
//   utils/cluster.py:103-195 create_condensed_dist_matrix (float32)    -> sketch_condensed_kernel
//   utils/cluster.py:270-290 the distances of one DFS step             -> sketch_rows_kernel
//
// h(s) of the reference: for every k-mer x of s, v = (a * int(md5(x).hexdigest(), 16) + b) mod (2^31 - 1); the
// signature is the N smallest v IN SORTED ORDER, as a multiset (a k-mer that occurs twice contributes twice,
// heapq.nsmallest over a generator), and when s has fewer than N k-mers the k-mer list is run through
// ceil(N / num_kmers) times (:131-139), i.e. every value counts that many times.
//
// Kernels: (1) one thread per k-mer start: single-block MD5 of the k bytes (k <= 55), the 128-bit digest read as
// a big-endian integer and folded mod 2^31 - 1 (2^32 = 2, 2^64 = 4, 2^96 = 8 mod p), then the affine map; 4 bytes
// written per position.  (2) one CTA per sequence: exact multiset selection of the N smallest values by three
// radix-histogram passes (11 + 10 + 10 bits) in shared memory, then a collect + bitonic sort of the < N values
// below the N-th one.  Both are integer-ALU work; the scratch of 4 bytes per base stays in L2 between (1) and (2)
// for batches below ~30 Mbp.
#include <algorithm>
#include <cstring>

#include "internal.cuh"

struct cb_sketches {
    cb_ctx *ctx = nullptr;
    int64_t n = 0;
    int32_t N = 0;
    uint32_t *d_sig = nullptr;       // [n][N], each row sorted ascending
};

namespace {

constexpr uint32_t P31 = 2147483647u;
constexpr int SK_THREADS = 256;
constexpr int SK_MAX_N = 1024;              // values per sketch (the reference's default is 100)
constexpr int SK_MAX_K = 55;                // k-mer bytes that fit one MD5 block with padding and length
constexpr int64_t SK_BATCH_BASES = 1ll << 28;

#define MD5_F(x, y, z) ((z) ^ ((x) & ((y) ^ (z))))
#define MD5_G(x, y, z) ((y) ^ ((z) & ((x) ^ (y))))
#define MD5_H(x, y, z) ((x) ^ (y) ^ (z))
#define MD5_I(x, y, z) ((y) ^ ((x) | ~(z)))
#define MD5_STEP(f, a, b, c, d, x, k, s) \
    do { (a) += f((b), (c), (d)) + (x) + (k); (a) = __funnelshift_l((a), (a), (s)); (a) += (b); } while (0)

// MD5 of one 64-byte block m[16] (message already padded); digest words A, B, C, D.
__device__ __forceinline__ void md5_block(const uint32_t (&m)[16], uint32_t &A, uint32_t &B, uint32_t &C, uint32_t &D)
{
    uint32_t a = 0x67452301u, b = 0xefcdab89u, c = 0x98badcfeu, d = 0x10325476u;
    MD5_STEP(MD5_F, a, b, c, d, m[ 0], 0xd76aa478u,  7);
    MD5_STEP(MD5_F, d, a, b, c, m[ 1], 0xe8c7b756u, 12);
    MD5_STEP(MD5_F, c, d, a, b, m[ 2], 0x242070dbu, 17);
    MD5_STEP(MD5_F, b, c, d, a, m[ 3], 0xc1bdceeeu, 22);
    MD5_STEP(MD5_F, a, b, c, d, m[ 4], 0xf57c0fafu,  7);
    MD5_STEP(MD5_F, d, a, b, c, m[ 5], 0x4787c62au, 12);
    MD5_STEP(MD5_F, c, d, a, b, m[ 6], 0xa8304613u, 17);
    MD5_STEP(MD5_F, b, c, d, a, m[ 7], 0xfd469501u, 22);
    MD5_STEP(MD5_F, a, b, c, d, m[ 8], 0x698098d8u,  7);
    MD5_STEP(MD5_F, d, a, b, c, m[ 9], 0x8b44f7afu, 12);
    MD5_STEP(MD5_F, c, d, a, b, m[10], 0xffff5bb1u, 17);
    MD5_STEP(MD5_F, b, c, d, a, m[11], 0x895cd7beu, 22);
    MD5_STEP(MD5_F, a, b, c, d, m[12], 0x6b901122u,  7);
    MD5_STEP(MD5_F, d, a, b, c, m[13], 0xfd987193u, 12);
    MD5_STEP(MD5_F, c, d, a, b, m[14], 0xa679438eu, 17);
    MD5_STEP(MD5_F, b, c, d, a, m[15], 0x49b40821u, 22);
    MD5_STEP(MD5_G, a, b, c, d, m[ 1], 0xf61e2562u,  5);
    MD5_STEP(MD5_G, d, a, b, c, m[ 6], 0xc040b340u,  9);
    MD5_STEP(MD5_G, c, d, a, b, m[11], 0x265e5a51u, 14);
    MD5_STEP(MD5_G, b, c, d, a, m[ 0], 0xe9b6c7aau, 20);
    MD5_STEP(MD5_G, a, b, c, d, m[ 5], 0xd62f105du,  5);
    MD5_STEP(MD5_G, d, a, b, c, m[10], 0x02441453u,  9);
    MD5_STEP(MD5_G, c, d, a, b, m[15], 0xd8a1e681u, 14);
    MD5_STEP(MD5_G, b, c, d, a, m[ 4], 0xe7d3fbc8u, 20);
    MD5_STEP(MD5_G, a, b, c, d, m[ 9], 0x21e1cde6u,  5);
    MD5_STEP(MD5_G, d, a, b, c, m[14], 0xc33707d6u,  9);
    MD5_STEP(MD5_G, c, d, a, b, m[ 3], 0xf4d50d87u, 14);
    MD5_STEP(MD5_G, b, c, d, a, m[ 8], 0x455a14edu, 20);
    MD5_STEP(MD5_G, a, b, c, d, m[13], 0xa9e3e905u,  5);
    MD5_STEP(MD5_G, d, a, b, c, m[ 2], 0xfcefa3f8u,  9);
    MD5_STEP(MD5_G, c, d, a, b, m[ 7], 0x676f02d9u, 14);
    MD5_STEP(MD5_G, b, c, d, a, m[12], 0x8d2a4c8au, 20);
    MD5_STEP(MD5_H, a, b, c, d, m[ 5], 0xfffa3942u,  4);
    MD5_STEP(MD5_H, d, a, b, c, m[ 8], 0x8771f681u, 11);
    MD5_STEP(MD5_H, c, d, a, b, m[11], 0x6d9d6122u, 16);
    MD5_STEP(MD5_H, b, c, d, a, m[14], 0xfde5380cu, 23);
    MD5_STEP(MD5_H, a, b, c, d, m[ 1], 0xa4beea44u,  4);
    MD5_STEP(MD5_H, d, a, b, c, m[ 4], 0x4bdecfa9u, 11);
    MD5_STEP(MD5_H, c, d, a, b, m[ 7], 0xf6bb4b60u, 16);
    MD5_STEP(MD5_H, b, c, d, a, m[10], 0xbebfbc70u, 23);
    MD5_STEP(MD5_H, a, b, c, d, m[13], 0x289b7ec6u,  4);
    MD5_STEP(MD5_H, d, a, b, c, m[ 0], 0xeaa127fau, 11);
    MD5_STEP(MD5_H, c, d, a, b, m[ 3], 0xd4ef3085u, 16);
    MD5_STEP(MD5_H, b, c, d, a, m[ 6], 0x04881d05u, 23);
    MD5_STEP(MD5_H, a, b, c, d, m[ 9], 0xd9d4d039u,  4);
    MD5_STEP(MD5_H, d, a, b, c, m[12], 0xe6db99e5u, 11);
    MD5_STEP(MD5_H, c, d, a, b, m[15], 0x1fa27cf8u, 16);
    MD5_STEP(MD5_H, b, c, d, a, m[ 2], 0xc4ac5665u, 23);
    MD5_STEP(MD5_I, a, b, c, d, m[ 0], 0xf4292244u,  6);
    MD5_STEP(MD5_I, d, a, b, c, m[ 7], 0x432aff97u, 10);
    MD5_STEP(MD5_I, c, d, a, b, m[14], 0xab9423a7u, 15);
    MD5_STEP(MD5_I, b, c, d, a, m[ 5], 0xfc93a039u, 21);
    MD5_STEP(MD5_I, a, b, c, d, m[12], 0x655b59c3u,  6);
    MD5_STEP(MD5_I, d, a, b, c, m[ 3], 0x8f0ccc92u, 10);
    MD5_STEP(MD5_I, c, d, a, b, m[10], 0xffeff47du, 15);
    MD5_STEP(MD5_I, b, c, d, a, m[ 1], 0x85845dd1u, 21);
    MD5_STEP(MD5_I, a, b, c, d, m[ 8], 0x6fa87e4fu,  6);
    MD5_STEP(MD5_I, d, a, b, c, m[15], 0xfe2ce6e0u, 10);
    MD5_STEP(MD5_I, c, d, a, b, m[ 6], 0xa3014314u, 15);
    MD5_STEP(MD5_I, b, c, d, a, m[13], 0x4e0811a1u, 21);
    MD5_STEP(MD5_I, a, b, c, d, m[ 4], 0xf7537e82u,  6);
    MD5_STEP(MD5_I, d, a, b, c, m[11], 0xbd3af235u, 10);
    MD5_STEP(MD5_I, c, d, a, b, m[ 2], 0x2ad7d2bbu, 15);
    MD5_STEP(MD5_I, b, c, d, a, m[ 9], 0xeb86d391u, 21);
    A = a + 0x67452301u; B = b + 0xefcdab89u; C = c + 0x98badcfeu; D = d + 0x10325476u;
}

// (a * int(hexdigest, 16) + b) mod p.  hexdigest prints the digest bytes in memory order, so the integer is the
// big-endian reading of A|B|C|D as stored little-endian: W0 = bswap(A) is the most significant word.
__device__ __forceinline__ uint32_t affine_of_digest(uint32_t A, uint32_t B, uint32_t C, uint32_t D, uint64_t a, uint64_t b)
{
    uint64_t x = 8ull * __byte_perm(A, 0, 0x0123) + 4ull * __byte_perm(B, 0, 0x0123) +
                 2ull * __byte_perm(C, 0, 0x0123) + (uint64_t)__byte_perm(D, 0, 0x0123);     // < 15 * 2^32
    x = (x & P31) + (x >> 31);
    x = (x & P31) + (x >> 31);
    if (x >= P31) x -= P31;
    uint64_t v = a * x + b;                  // a, b already reduced mod p: < 2^62 + 2^31
    v = (v & P31) + (v >> 31);
    v = (v & P31) + (v >> 31);
    if (v >= P31) v -= P31;
    return (uint32_t)v;
}

// K > 0: k-mer length known at compile time (12 is what the reference clusters with, cluster.py:358); K == 0: run time.
template <int K>
__global__ void __launch_bounds__(SK_THREADS)
sketch_hash_kernel(const uint8_t *__restrict__ ascii, const int64_t *__restrict__ seq_off, int64_t seq_lo, int64_t seq_hi,
                   int64_t base, int64_t total, int k_rt, uint64_t a, uint64_t b, uint32_t *__restrict__ out)
{
    const int k = K > 0 ? K : k_rt;
    __shared__ int64_t s_first;
    int64_t p0 = (int64_t)blockIdx.x * SK_THREADS;
    if (threadIdx.x == 0) {              // sequence that holds the CTA's first position
        int64_t lo = seq_lo, hi = seq_hi;        // seq_off[lo] - base <= p0 < seq_off[hi] - base
        while (hi - lo > 1) {
            int64_t mid = (lo + hi) >> 1;
            if (seq_off[mid] - base <= p0) lo = mid; else hi = mid;
        }
        s_first = lo;
    }
    __syncthreads();
    int64_t p = p0 + threadIdx.x;
    if (p >= total) return;
    int64_t s = s_first;
    while (seq_off[s + 1] - base <= p) s++;
    int64_t end = seq_off[s + 1] - base;
    if (p + k > end) { out[p] = 0xffffffffu; return; }       // not a k-mer start
    const uint8_t *x = ascii + p;
    uint32_t m[16];
#pragma unroll
    for (int w = 0; w < 14; w++) {
        uint32_t v = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int i = 4 * w + j;
            uint32_t byte = 0;
            if (K > 0) byte = i < K ? (uint32_t)__ldg(x + i) : (i == K ? 0x80u : 0u);
            else if (i < k) byte = __ldg(x + i);
            else if (i == k) byte = 0x80u;
            v |= byte << (8 * j);
        }
        m[w] = v;
    }
    m[14] = 8u * (uint32_t)k;
    m[15] = 0;
    uint32_t A, B, C, D;
    md5_block(m, A, B, C, D);
    out[p] = affine_of_digest(A, B, C, D, a, b);
}

// Exclusive block scan of one value per thread (SK_THREADS threads); returns the exclusive prefix, *total the sum.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *s_warp, uint32_t *total)
{
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    uint32_t off = 0, sum = 0;
#pragma unroll
    for (int w = 0; w < SK_THREADS / 32; w++) {
        uint32_t t = s_warp[w];
        if (w < wid) off += t;
        sum += t;
    }
    __syncthreads();
    *total = sum;
    return off + inc - v;
}

// One CTA per sequence: the N smallest values of the multiset {h[i] x mult}, sorted.
__global__ void __launch_bounds__(SK_THREADS)
sketch_select_kernel(const uint32_t *__restrict__ h, const int64_t *__restrict__ seq_off, int64_t seq_lo, int64_t base,
                     int k, int N, uint32_t *__restrict__ sig)
{
    __shared__ uint32_t hist[2048];
    __shared__ uint32_t list[SK_MAX_N];
    __shared__ uint32_t s_warp[SK_THREADS / 32];
    __shared__ uint32_t s_bin, s_rank, s_cnt;
    const int64_t s = seq_lo + blockIdx.x;
    const int64_t beg = seq_off[s] - base;
    const int64_t nk = seq_off[s + 1] - seq_off[s] - k + 1;          // >= 1, checked by the host
    const uint32_t mult = (uint32_t)((N + nk - 1) / nk);             // 1 unless the sequence has fewer than N k-mers
    const uint32_t *v = h + beg;
    const int tid = threadIdx.x;

    uint32_t prefix = 0;                 // the bits of the N-th smallest value found so far (high part)
    uint32_t rank = (uint32_t)N;         // 1-based rank still to be located inside the current prefix class
#pragma unroll 1
    for (int pass = 0; pass < 3; pass++) {
        const int shift = pass == 0 ? 20 : (pass == 1 ? 10 : 0);
        const int bins = pass == 0 ? 2048 : 1024;
        const int hi_shift = pass == 0 ? 31 : (pass == 1 ? 20 : 10);  // bits above the current digit
        for (int i = tid; i < 2048; i += SK_THREADS) hist[i] = 0;
        __syncthreads();
        for (int64_t i = tid; i < nk; i += SK_THREADS) {
            uint32_t x = v[i];
            if ((x >> hi_shift) == prefix) atomicAdd(&hist[(x >> shift) & (bins - 1)], mult);
        }
        __syncthreads();
        const int per = bins / SK_THREADS;           // 8 or 4 bins per thread
        uint32_t local = 0;
        for (int j = 0; j < per; j++) local += hist[tid * per + j];
        uint32_t total;
        uint32_t excl = block_excl_scan(local, s_warp, &total);
        if (excl < rank && rank <= excl + local) {   // exactly one thread
            uint32_t c = excl;
            int j = 0;
            for (; j < per; j++) {
                uint32_t t = hist[tid * per + j];
                if (rank <= c + t) break;
                c += t;
            }
            s_bin = tid * per + j;
            s_rank = rank - c;
        }
        __syncthreads();
        prefix = (prefix << (pass == 0 ? 11 : 10)) | s_bin;
        rank = s_rank;
        __syncthreads();
    }
    // prefix is the N-th smallest value; `rank` copies of it close the sketch, everything smaller comes first
    const uint32_t vN = prefix;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    for (int64_t i = tid; i < nk; i += SK_THREADS) {
        uint32_t x = v[i];
        if (x < vN) {
            uint32_t at = atomicAdd(&s_cnt, mult);
            for (uint32_t c = 0; c < mult; c++) list[at + c] = x;
        }
    }
    __syncthreads();
    const uint32_t below = s_cnt;                    // == N - rank
    for (uint32_t i = below + tid; i < SK_MAX_N; i += SK_THREADS) list[i] = i < (uint32_t)N ? vN : 0xffffffffu;
    __syncthreads();
    int n2 = 32;
    while (n2 < N) n2 <<= 1;
    for (int size = 2; size <= n2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < n2 / 2; i += SK_THREADS) {
                int lo = 2 * i - (i & (stride - 1));
                int hi = lo + stride;
                bool up = (lo & size) == 0;
                uint32_t x = list[lo], y = list[hi];
                if ((x > y) == up) { list[lo] = y; list[hi] = x; }
            }
            __syncthreads();
        }
    for (int i = tid; i < N; i += SK_THREADS) sig[(int64_t)blockIdx.x * N + i] = list[i];
}

// MinHashFamily.estimate_jaccard_dist (utils/lsh.py:188-214), the loop as written: both sketches hold N sorted
// values (repeats allowed); walk them together until N values of the union have been seen.
__device__ __forceinline__ void sketch_merge(const uint32_t *A, const uint32_t *B, int N, int &inter, int &uni)
{
    int ia = 0, ib = 0, in = 0, un = 0;
    uint32_t x = A[0], y = B[0];
    while (ia < N && ib < N && un < N) {
        if (x < y) { ia++; if (ia < N) x = A[ia]; }
        else if (x > y) { ib++; if (ib < N) y = B[ib]; }
        else { in++; ia++; ib++; if (ia < N) x = A[ia]; if (ib < N) y = B[ib]; }
        un++;
    }
    inter = in;
    uni = un;
}

__device__ __forceinline__ double jaccard_dist(int inter, int uni)
{
    double similarity = (double)inter / (double)uni;     // float(intersect_count) / union_count
    return 1.0 - similarity;
}

// out[r][c] = distance between sketch rows[r] and sketch c, as the Python double the reference computes
__global__ void __launch_bounds__(128)
sketch_rows_kernel(const uint32_t *__restrict__ sig, int64_t n, int N, const int64_t *__restrict__ rows, double *__restrict__ out)
{
    __shared__ uint32_t srow[SK_MAX_N];
    const int64_t r = rows[blockIdx.y];
    for (int i = threadIdx.x; i < N; i += blockDim.x) srow[i] = sig[r * N + i];
    __syncthreads();
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    int inter, uni;
    sketch_merge(srow, sig + c * N, N, inter, uni);
    out[(int64_t)blockIdx.y * n + c] = jaccard_dist(inter, uni);
}

// condensed matrix in scipy's layout (cluster.py:95-101: index of (i, j), i < j, is i*n - i*(i+3)/2 + j - 1), stored as
// float32 like the reference's sharedctypes c_float array (:141-142)
__global__ void __launch_bounds__(128)
sketch_condensed_kernel(const uint32_t *__restrict__ sig, int64_t n, int N, float *__restrict__ out)
{
    __shared__ uint32_t srow[SK_MAX_N];
    const int64_t i = blockIdx.y;
    const int64_t j0 = (int64_t)blockIdx.x * blockDim.x;
    if (j0 + blockDim.x <= i + 1) return;                 // whole CTA at or below the diagonal
    for (int t = threadIdx.x; t < N; t += blockDim.x) srow[t] = sig[i * N + t];
    __syncthreads();
    int64_t j = j0 + threadIdx.x;
    if (j <= i || j >= n) return;
    int inter, uni;
    sketch_merge(srow, sig + j * N, N, inter, uni);
    out[i * n - i * (i + 3) / 2 + j - 1] = (float)jaccard_dist(inter, uni);
}

// The sketches within `threshold` of each requested row, as compact (column, distance) lists in ascending column
// order: what one step of the connected-components search needs (cluster.py:270-290 looks at every distance of the
// row, but only those within the threshold have an effect).  Pass 1 counts per row, pass 2 writes at the row's offset.
__global__ void __launch_bounds__(256)
near_count_kernel(const double *__restrict__ d, int64_t n, double threshold, uint32_t *__restrict__ counts)
{
    const double *row = d + (int64_t)blockIdx.x * n;
    uint32_t c = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) c += row[i] <= threshold;
    __shared__ uint32_t s_w[8];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < 8; w++) t += s_w[w];
        counts[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256)
near_compact_kernel(const double *__restrict__ d, int64_t n, double threshold, const int64_t *__restrict__ row_off,
                    uint32_t *__restrict__ idx_out, double *__restrict__ dist_out)
{
    const double *row = d + (int64_t)blockIdx.x * n;
    __shared__ uint32_t s_w[8];
    __shared__ uint32_t s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int64_t out0 = row_off[blockIdx.x];
    for (int64_t i0 = 0; i0 < n; i0 += blockDim.x) {
        const int64_t i = i0 + threadIdx.x;
        const double v = i < n ? row[i] : 2.0;
        const bool hit = i < n && v <= threshold;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_w[warp] = __popc(m);
        __syncthreads();
        uint32_t before = s_base;
        for (int w = 0; w < warp; w++) before += s_w[w];
        if (hit) {
            const int64_t o = out0 + before + __popc(m & ((1u << lane) - 1u));
            idx_out[o] = (uint32_t)i;
            dist_out[o] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int w = 0; w < 8; w++) t += s_w[w];
            s_base += t;
        }
        __syncthreads();
    }
}

}  // namespace

int cb_sketch_near_rows_impl(cb_ctx *ctx, const cb_sketches *sk, const int64_t *rows, int64_t n_rows, double threshold,
                             int64_t *row_off, uint32_t **idx, double **dist)
{
    if (!sk || !row_off || !idx || !dist || (n_rows > 0 && !rows)) return cb_fail(ctx, CB_ERR_ARG, "cb_sketch_near_rows: NULL argument");
    *idx = nullptr;
    *dist = nullptr;
    row_off[0] = 0;
    for (int64_t r = 0; r < n_rows; r++) row_off[r + 1] = 0;
    if (n_rows == 0 || sk->n == 0) return CB_OK;
    for (int64_t r = 0; r < n_rows; r++)
        if (rows[r] < 0 || rows[r] >= sk->n) return cb_fail(ctx, CB_ERR_ARG, "cb_sketch_near_rows: row out of range");
    if (n_rows > 65535) return cb_fail(ctx, CB_ERR_ARG, "cb_sketch_near_rows: at most 65535 rows per call");
    cudaStream_t st = ctx->stream;
    cb_tls_stream = st;
    const int64_t n = sk->n;
    DevBuf<int64_t> d_rows, d_off;
    DevBuf<double> d_full, d_dist;
    DevBuf<uint32_t> d_counts, d_idx;
    if (d_rows.alloc((size_t)n_rows) != cudaSuccess || d_full.alloc((size_t)n_rows * n) != cudaSuccess ||
        d_counts.alloc((size_t)n_rows) != cudaSuccess || d_off.alloc((size_t)n_rows + 1) != cudaSuccess)
        return cb_fail(ctx, CB_ERR_NOMEM, "cb_sketch_near_rows: out of device memory");
    CB_CUDA(ctx, cudaMemcpyAsync(d_rows.p, rows, (size_t)n_rows * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    dim3 grid((unsigned)((n + 127) / 128), (unsigned)n_rows);
    sketch_rows_kernel<<<grid, 128, 0, st>>>(sk->d_sig, n, sk->N, d_rows.p, d_full.p);
    near_count_kernel<<<(unsigned)n_rows, 256, 0, st>>>(d_full.p, n, threshold, d_counts.p);
    CB_CUDA(ctx, cudaGetLastError());
    std::vector<uint32_t> h_counts((size_t)n_rows);
    CB_CUDA(ctx, cudaMemcpyAsync(h_counts.data(), d_counts.p, (size_t)n_rows * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    for (int64_t r = 0; r < n_rows; r++) row_off[r + 1] = row_off[r] + (int64_t)h_counts[(size_t)r];
    const int64_t total = row_off[n_rows];
    ctx->launches += 2;
    if (total == 0) return CB_OK;
    if (d_idx.alloc((size_t)total) != cudaSuccess || d_dist.alloc((size_t)total) != cudaSuccess)
        return cb_fail(ctx, CB_ERR_NOMEM, "cb_sketch_near_rows: out of device memory");
    CB_CUDA(ctx, cudaMemcpyAsync(d_off.p, row_off, (size_t)(n_rows + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    near_compact_kernel<<<(unsigned)n_rows, 256, 0, st>>>(d_full.p, n, threshold, d_off.p, d_idx.p, d_dist.p);
    CB_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    uint32_t *h_idx = (uint32_t *)malloc((size_t)total * sizeof(uint32_t));
    double *h_dist = (double *)malloc((size_t)total * sizeof(double));
    if (!h_idx || !h_dist) { free(h_idx); free(h_dist); return cb_fail(ctx, CB_ERR_NOMEM, "cb_sketch_near_rows: out of host memory"); }
    if (cudaMemcpyAsync(h_idx, d_idx.p, (size_t)total * sizeof(uint32_t), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaMemcpyAsync(h_dist, d_dist.p, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) {
        free(h_idx);
        free(h_dist);
        return cb_fail(ctx, CB_ERR_CUDA, "cb_sketch_near_rows: copy failed");
    }
    *idx = h_idx;
    *dist = h_dist;
    return CB_OK;
}

int cb_sketch_sequences_impl(cb_ctx *ctx, const uint8_t *ascii, const int64_t *seq_off, int64_t n, int32_t k,
                             int32_t N, uint64_t a, uint64_t b, cb_sketches **out, cb_stats *stats)
{
    if (!out) return cb_fail(ctx, CB_ERR_ARG, "cb_sketch_sequences: out is NULL");
    *out = nullptr;
    if (n < 0 || (n > 0 && (!ascii || !seq_off))) return cb_fail(ctx, CB_ERR_ARG, "cb_sketch_sequences: missing input");
    if (k < 1 || k > SK_MAX_K) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "cb_sketch_sequences: kmer_size must be 1..55");
    if (N < 1 || N > SK_MAX_N) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "cb_sketch_sequences: N must be 1..1024");
    for (int64_t s = 0; s < n; s++)
        if (seq_off[s + 1] - seq_off[s] < k)     // the reference asserts kmer_size <= len(s), utils/lsh.py:117
            return cb_fail(ctx, CB_ERR_ARG, "cb_sketch_sequences: a sequence is shorter than kmer_size");
    cudaStream_t st = ctx->stream;
    cb_tls_stream = st;
    cb_sketches *sk = new cb_sketches;
    sk->ctx = ctx;
    sk->n = n;
    sk->N = N;
    cb_stats local;
    memset(&local, 0, sizeof local);
    EventTimer t_all(st), t_hash(st), t_sel(st);
    double ms_hash = 0, ms_sel = 0;
    t_all.start();
    if (cb_dev_alloc(st, (void **)&sk->d_sig, (size_t)std::max<int64_t>(n, 1) * N * sizeof(uint32_t)) != cudaSuccess) {
        delete sk;
        return cb_fail(ctx, CB_ERR_NOMEM, "cb_sketch_sequences: out of device memory");
    }
    const uint64_t am = a % P31, bm = b % P31;   // (a*x + b) mod p only depends on a, b mod p
    DevBuf<int64_t> d_off;
    if (d_off.alloc((size_t)n + 1) != cudaSuccess) { cb_sketches_free(sk); return cb_fail(ctx, CB_ERR_NOMEM, "out of device memory"); }
    int rc = CB_OK;
    auto body = [&]() -> int {
        if (n == 0) return CB_OK;
        CB_CUDA(ctx, cudaMemcpyAsync(d_off.p, seq_off, (size_t)(n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        int64_t lo = 0;
        while (lo < n) {
            int64_t hi = lo + 1;
            while (hi < n && seq_off[hi + 1] - seq_off[lo] <= SK_BATCH_BASES) hi++;
            const int64_t base = seq_off[lo], total = seq_off[hi] - base;
            DevBuf<uint8_t> d_ascii;
            DevBuf<uint32_t> d_h;
            if (d_ascii.alloc((size_t)total) != cudaSuccess || d_h.alloc((size_t)total) != cudaSuccess)
                return cb_fail(ctx, CB_ERR_NOMEM, "cb_sketch_sequences: out of device memory");
            CB_CUDA(ctx, cudaMemcpyAsync(d_ascii.p, ascii + base, (size_t)total, cudaMemcpyHostToDevice, st));
            const unsigned grid = (unsigned)((total + SK_THREADS - 1) / SK_THREADS);
            t_hash.start();
            if (k == 12)
                sketch_hash_kernel<12><<<grid, SK_THREADS, 0, st>>>(d_ascii.p, d_off.p, lo, hi, base, total, k, am, bm, d_h.p);
            else
                sketch_hash_kernel<0><<<grid, SK_THREADS, 0, st>>>(d_ascii.p, d_off.p, lo, hi, base, total, k, am, bm, d_h.p);
            t_hash.stop();
            t_sel.start();
            sketch_select_kernel<<<(unsigned)(hi - lo), SK_THREADS, 0, st>>>(d_h.p, d_off.p, lo, base, k, N,
                                                                             sk->d_sig + lo * N);
            t_sel.stop();
            CB_CUDA(ctx, cudaGetLastError());
            ctx->launches += 2;
            local.n_kernel_launches += 2;
            local.n_seed_lookups += total;
            ms_hash += t_hash.ms();
            ms_sel += t_sel.ms();
            lo = hi;
        }
        return CB_OK;
    };
    rc = body();
    t_all.stop();
    if (rc == CB_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = cb_fail(ctx, CB_ERR_CUDA, "cb_sketch_sequences: kernel failed");
    if (rc != CB_OK) { cb_sketches_free(sk); return rc; }
    local.ms_scan_emit = ms_hash;
    local.ms_merge = ms_sel;
    local.ms_total = t_all.ms();
    if (stats) *stats = local;
    *out = sk;
    return CB_OK;
}

int cb_sketches_export_impl(cb_ctx *ctx, const cb_sketches *sk, uint32_t *sig)
{
    if (!sk || !sig) return cb_fail(ctx, CB_ERR_ARG, "cb_sketches_export: NULL argument");
    CB_CUDA(ctx, cudaMemcpyAsync(sig, sk->d_sig, (size_t)sk->n * sk->N * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CB_OK;
}

int cb_sketch_import_impl(cb_ctx *ctx, const uint32_t *sig, int64_t n, int32_t N, cb_sketches **out)
{
    if (!out || n < 0 || (n > 0 && !sig)) return cb_fail(ctx, CB_ERR_ARG, "cb_sketches_import: bad argument");
    if (N < 1 || N > SK_MAX_N) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "cb_sketches_import: N must be 1..1024");
    cb_sketches *sk = new cb_sketches;
    sk->ctx = ctx; sk->n = n; sk->N = N;
    if (cb_dev_alloc(ctx->stream, (void **)&sk->d_sig, (size_t)std::max<int64_t>(n, 1) * N * sizeof(uint32_t)) != cudaSuccess) {
        delete sk;
        return cb_fail(ctx, CB_ERR_NOMEM, "cb_sketches_import: out of device memory");
    }
    if (n && (cudaMemcpyAsync(sk->d_sig, sig, (size_t)n * N * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
              cudaStreamSynchronize(ctx->stream) != cudaSuccess)) {
        cb_sketches_free(sk);
        return cb_fail(ctx, CB_ERR_CUDA, "cb_sketches_import: copy failed");
    }
    *out = sk;
    return CB_OK;
}

int cb_sketch_dist_rows_impl(cb_ctx *ctx, const cb_sketches *sk, const int64_t *rows, int64_t n_rows, double *out)
{
    if (!sk || (n_rows > 0 && (!rows || !out))) return cb_fail(ctx, CB_ERR_ARG, "cb_sketch_dist_rows: NULL argument");
    if (n_rows == 0 || sk->n == 0) return CB_OK;
    for (int64_t r = 0; r < n_rows; r++)
        if (rows[r] < 0 || rows[r] >= sk->n) return cb_fail(ctx, CB_ERR_ARG, "cb_sketch_dist_rows: row out of range");
    if (n_rows > 65535) return cb_fail(ctx, CB_ERR_ARG, "cb_sketch_dist_rows: at most 65535 rows per call");
    cudaStream_t st = ctx->stream;
    cb_tls_stream = st;
    DevBuf<int64_t> d_rows;
    DevBuf<double> d_out;
    if (d_rows.alloc((size_t)n_rows) != cudaSuccess || d_out.alloc((size_t)n_rows * sk->n) != cudaSuccess)
        return cb_fail(ctx, CB_ERR_NOMEM, "cb_sketch_dist_rows: out of device memory");
    CB_CUDA(ctx, cudaMemcpyAsync(d_rows.p, rows, (size_t)n_rows * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    dim3 grid((unsigned)((sk->n + 127) / 128), (unsigned)n_rows);
    sketch_rows_kernel<<<grid, 128, 0, st>>>(sk->d_sig, sk->n, sk->N, d_rows.p, d_out.p);
    CB_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    CB_CUDA(ctx, cudaMemcpyAsync(out, d_out.p, (size_t)n_rows * sk->n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    return CB_OK;
}

int cb_sketch_dist_condensed_impl(cb_ctx *ctx, const cb_sketches *sk, float *out)
{
    if (!sk) return cb_fail(ctx, CB_ERR_ARG, "cb_sketch_dist_condensed: NULL argument");
    const int64_t n = sk->n;
    if (n < 2) return CB_OK;
    if (!out) return cb_fail(ctx, CB_ERR_ARG, "cb_sketch_dist_condensed: out is NULL");
    if (n > 65535) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "cb_sketch_dist_condensed: at most 65535 sequences");
    const int64_t m = n * (n - 1) / 2;
    cudaStream_t st = ctx->stream;
    cb_tls_stream = st;
    DevBuf<float> d_out;
    if (d_out.alloc((size_t)m) != cudaSuccess) return cb_fail(ctx, CB_ERR_NOMEM, "cb_sketch_dist_condensed: out of device memory");
    dim3 grid((unsigned)((n + 127) / 128), (unsigned)n);
    sketch_condensed_kernel<<<grid, 128, 0, st>>>(sk->d_sig, n, sk->N, d_out.p);
    CB_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    CB_CUDA(ctx, cudaMemcpyAsync(out, d_out.p, (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    return CB_OK;
}

extern "C" void cb_sketches_free(cb_sketches *sk)
{
    if (!sk) return;
    if (sk->d_sig) cudaFreeAsync(sk->d_sig, sk->ctx->stream);
    delete sk;
}
