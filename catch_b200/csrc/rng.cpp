// MT19937 replay (host only): continues numpy's legacy global stream exactly as
// RandomState.randint(0, bound, size=n) / np.random.choice(bound, n) consume it -- masked rejection
// sampling on tempered 32-bit outputs (numpy/random/src/distributions/distributions.c,
// legacy rk_interval) -- so that the seed draws of probe.py:393-396 can be reproduced at memory
// speed.  In a group-sharded multi-GPU run these draws are the one inherently sequential step (the
// stream continues from grouping to grouping), so the inner loop matters: with AVX-512 the state
// update, the tempering and the rejection (compress-store) all run 16 lanes wide; the portable
// scalar loop gives bit-identical output and is used when the CPU lacks AVX-512 F/BW/VL.
#include <immintrin.h>
#include <stdint.h>

#include <thread>

#include "../../include/catch_b200.h"

namespace {

constexpr uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX_A = 0x9908b0dfu;

inline uint32_t twist(uint32_t a, uint32_t b, uint32_t far)
{
    const uint32_t y = (a & UPPER) | (b & LOWER);
    return far ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
}

inline uint32_t temper(uint32_t y)
{
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

void gen_scalar(uint32_t *mt)
{
    int kk;
    for (kk = 0; kk < 624 - 397; kk++) mt[kk] = twist(mt[kk], mt[kk + 1], mt[kk + 397]);
    for (; kk < 623; kk++) mt[kk] = twist(mt[kk], mt[kk + 1], mt[kk + (397 - 624)]);
    mt[623] = twist(mt[623], mt[0], mt[396]);
}

// Portable loop: rejection over the rest of the current block; the accept is branch-free (write, then
// advance only if the value is in range).
template <typename T>
void draw_scalar(uint32_t *key, int &p, uint32_t mask, uint32_t rng, int64_t n, int64_t &i, T *out, bool one_block)
{
    while (i < n) {
        if (p == 624) { gen_scalar(key); p = 0; }
        int j = p;
        for (; j < 624 && i < n; j++) {
            const uint32_t val = temper(key[j]) & mask;
            out[i] = (T)val;
            i += (val <= rng);
        }
        p = j;
        if (one_block) return;
    }
}

#if defined(__x86_64__)
#define CB_AVX512 __attribute__((target("avx512f,avx512bw,avx512vl")))

CB_AVX512 inline __m512i twist16(__m512i a, __m512i b, __m512i far)
{
    const __m512i y = _mm512_or_si512(_mm512_and_si512(a, _mm512_set1_epi32((int)UPPER)),
                                      _mm512_and_si512(b, _mm512_set1_epi32((int)LOWER)));
    const __mmask16 odd = _mm512_test_epi32_mask(y, _mm512_set1_epi32(1));
    const __m512i mag = _mm512_maskz_set1_epi32(odd, (int)MATRIX_A);
    return _mm512_xor_si512(_mm512_xor_si512(far, _mm512_srli_epi32(y, 1)), mag);
}

// 16 lanes at a time: a step reads mt[kk .. kk+16] before it writes mt[kk .. kk+15], and the far
// operand is 227 (first part, still old) or -227 (second part, already new) words away.
CB_AVX512 void gen_avx512(uint32_t *mt)
{
    int kk = 0;
    for (; kk + 16 <= 624 - 397; kk += 16) {
        const __m512i a = _mm512_loadu_si512(mt + kk), b = _mm512_loadu_si512(mt + kk + 1);
        _mm512_storeu_si512(mt + kk, twist16(a, b, _mm512_loadu_si512(mt + kk + 397)));
    }
    for (; kk < 624 - 397; kk++) mt[kk] = twist(mt[kk], mt[kk + 1], mt[kk + 397]);
    for (; kk + 16 <= 623; kk += 16) {
        const __m512i a = _mm512_loadu_si512(mt + kk), b = _mm512_loadu_si512(mt + kk + 1);
        _mm512_storeu_si512(mt + kk, twist16(a, b, _mm512_loadu_si512(mt + kk - 227)));
    }
    for (; kk < 623; kk++) mt[kk] = twist(mt[kk], mt[kk + 1], mt[kk - 227]);
    mt[623] = twist(mt[623], mt[0], mt[396]);
}

CB_AVX512 inline __m512i temper16(__m512i y)
{
    y = _mm512_xor_si512(y, _mm512_srli_epi32(y, 11));
    y = _mm512_xor_si512(y, _mm512_and_si512(_mm512_slli_epi32(y, 7), _mm512_set1_epi32((int)0x9d2c5680u)));
    y = _mm512_xor_si512(y, _mm512_and_si512(_mm512_slli_epi32(y, 15), _mm512_set1_epi32((int)0xefc60000u)));
    return _mm512_xor_si512(y, _mm512_srli_epi32(y, 18));
}

CB_AVX512 inline void store_accepted(int32_t *out, __mmask16 keep, __m512i val)
{
    _mm512_mask_compressstoreu_epi32(out, keep, val);
}
CB_AVX512 inline void store_accepted(uint8_t *out, __mmask16 keep, __m512i val)
{
    const __m128i b = _mm512_cvtepi32_epi8(_mm512_maskz_compress_epi32(keep, val));
    _mm_mask_storeu_epi8(out, (__mmask16)((1u << __builtin_popcount(keep)) - 1u), b);
}

template <typename T>
CB_AVX512 void draw_avx512(uint32_t *key, int &p, uint32_t mask, uint32_t rng, int64_t n, int64_t &i, T *out)
{
    const __m512i vmask = _mm512_set1_epi32((int)mask), vrng = _mm512_set1_epi32((int)rng);
    while (i < n) {
        if (p == 624) { gen_avx512(key); p = 0; }
        // whole 16-word chunks while the block has them and at least 16 more values are wanted
        // (a chunk yields at most 16, so neither `out` nor the stream position can overshoot)
        while (p + 16 <= 624 && i + 16 <= n) {
            const __m512i val = _mm512_and_si512(temper16(_mm512_loadu_si512(key + p)), vmask);
            const __mmask16 keep = _mm512_cmple_epu32_mask(val, vrng);
            store_accepted(out + i, keep, val);
            i += __builtin_popcount(keep);
            p += 16;
        }
        if (i < n && p < 624) {
            // ragged end of the block, or fewer than 16 values left: word by word
            const bool tail = i + 16 > n;
            int j = p;
            for (; j < 624 && i < n && (tail || j < p + 16); j++) {
                const uint32_t val = temper(key[j]) & mask;
                out[i] = (T)val;
                i += (val <= rng);
            }
            p = j;
        }
    }
}

bool have_avx512()
{
    static const bool ok = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
                           __builtin_cpu_supports("avx512vl");
    return ok;
}
#else
bool have_avx512() { return false; }
#endif

template <typename T>
int randint_t(uint32_t *key, int32_t *pos, uint32_t bound, int64_t n, T *out, bool allow_simd)
{
    if (!key || !pos || !out || bound == 0 || n < 0 || *pos < 0 || *pos > 624) return CB_ERR_ARG;
    const uint32_t rng = bound - 1;            // inclusive upper value
    if (rng == 0) {                            // numpy draws nothing when the range is a single value
        for (int64_t i = 0; i < n; i++) out[i] = 0;
        return CB_OK;
    }
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    int p = *pos;
    int64_t i = 0;
#if defined(__x86_64__)
    if (allow_simd && have_avx512()) draw_avx512<T>(key, p, mask, rng, n, i, out);
    else
#endif
        draw_scalar<T>(key, p, mask, rng, n, i, out, false);
    *pos = p;
    return CB_OK;
}

}  // namespace

extern "C" {

int cb_mt19937_randint(uint32_t *key, int32_t *pos, uint32_t bound, int64_t n, int32_t *out)
{
    return randint_t<int32_t>(key, pos, bound, n, out, true);
}

int cb_mt19937_randint_u8(uint32_t *key, int32_t *pos, uint32_t bound, int64_t n, uint8_t *out)
{
    if (bound > 256) return CB_ERR_ARG;
    return randint_t<uint8_t>(key, pos, bound, n, out, true);
}

// Portable loop only (what a CPU without AVX-512 runs): lets the tests compare both paths on one machine.
int cb_mt19937_randint_scalar(uint32_t *key, int32_t *pos, uint32_t bound, int64_t n, int32_t *out)
{
    return randint_t<int32_t>(key, pos, bound, n, out, false);
}

// The same replay on a native thread, so the host can pack and upload sequences meanwhile (no
// Python thread, no GIL hand-over): begin returns at once, end joins and returns the status.
struct cb_rng_job {
    std::thread th;
    int rc = CB_OK;
};

cb_rng_job *cb_mt19937_randint_begin(uint32_t *key, int32_t *pos, uint32_t bound, int64_t n, void *out,
                                     int32_t elem_size)
{
    cb_rng_job *job = new cb_rng_job();
    if (elem_size == 1)
        job->th = std::thread([=]() { job->rc = cb_mt19937_randint_u8(key, pos, bound, n, (uint8_t *)out); });
    else if (elem_size == 4)
        job->th = std::thread([=]() { job->rc = cb_mt19937_randint(key, pos, bound, n, (int32_t *)out); });
    else
        job->rc = CB_ERR_ARG;
    return job;
}

int cb_mt19937_randint_end(cb_rng_job *job)
{
    if (!job) return CB_ERR_ARG;
    if (job->th.joinable()) job->th.join();
    const int rc = job->rc;
    delete job;
    return rc;
}

}  // extern "C"
