// K9-K12 near-duplicate filter (placeholder until the kernels land; fails loudly).
#include "internal.cuh"

int cb_minhash_neardup_impl(cb_ctx *ctx, const uint8_t *, const int64_t *, int64_t, const uint32_t *, const uint32_t *,
                            int32_t, int32_t, int32_t, double, uint8_t *, cb_stats *)
{
    return cb_fail(ctx, CB_ERR_UNSUPPORTED, "cb_minhash_neardup: not built yet");
}
int cb_hamming_neardup_impl(cb_ctx *ctx, const uint8_t *, const int64_t *, int64_t, const int32_t *, int32_t, int32_t,
                            int32_t, uint8_t *, cb_stats *)
{
    return cb_fail(ctx, CB_ERR_UNSUPPORTED, "cb_hamming_neardup: not built yet");
}
