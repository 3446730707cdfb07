// C-ABI dispatch of the sparse-convolution entry points (include/u2mkd.h).
#include <stdlib.h>

#include "u2_common.cuh"

int u2_conv_fwd_simt(const float *X, int64_t n_src, int32_t Cs, const float *W, int32_t w_transposed, const int32_t *table,
                     int64_t ld, int64_t n_dst, int32_t K, int32_t Cd, float *Y, cudaStream_t st);
int u2_conv_wgrad_simt(const float *X, int64_t n_src, int32_t Cs, const float *dY, int64_t n_dst, int32_t Cd,
                       const int32_t *table, int64_t ld, int32_t K, float *dW, cudaStream_t st);

#ifdef U2_WITH_TC
size_t u2_conv_tc_scratch_bytes(int64_t n_dst, int32_t K, int32_t Cs, int32_t Cd, int32_t math);
int u2_conv_tc_supported(int32_t Cs, int32_t Cd, int32_t K, int32_t math);
int u2_conv_fwd_tc(const void *X, int64_t n_src, int32_t Cs, const float *W, int32_t w_transposed, const int32_t *table,
                   const int32_t *perm, int64_t ld, int64_t n_dst, int32_t K, int32_t Cd, float *Y, int32_t math,
                   void *scratch, size_t scratch_bytes, float *tile_stats, const float *Yadd, cudaStream_t st);
size_t u2_conv_pretile_plan_bytes_tc(int32_t n_jobs);
int u2_conv_pretile_plan_tc(int32_t n_jobs, const uint64_t *W, const uint64_t *blob, const int32_t *K, const int32_t *Cs,
                            const int32_t *Cd, const int32_t *w_transposed, int32_t math, void *plan_host, size_t plan_bytes,
                            int64_t *n_blocks);
int u2_conv_pretile_run_tc(const void *plan_dev, int32_t n_jobs, int64_t n_blocks, cudaStream_t st);
int u2_conv_pretile_tc(const float *W, int32_t w_transposed, int32_t K, int32_t Cs, int32_t Cd, int32_t math, void *blob,
                       cudaStream_t st);
int u2_cast_bf16_impl(const float *x, int64_t n, void *y, cudaStream_t st);
int u2_split_bf16x3_impl(const float *x, int64_t n, int32_t C, void *out3, void *hi, void *lo, cudaStream_t st);
int u2_conv_wgrad_tc(const void *X, int32_t Cs, const void *dY, int32_t Cd, const int32_t *nbr, int64_t ld, int64_t n_rows,
                     int32_t K, const int32_t *flat, const int32_t *nbsizes, int32_t swap, float *dW, int32_t math,
                     int32_t dense_k, const int32_t *dense_ok, cudaStream_t st);
#endif

extern "C" int u2_has_tensor_core_path(void) {
#ifdef U2_WITH_TC
    return 1;
#else
    return 0;
#endif
}

extern "C" size_t u2_conv_scratch_bytes(int64_t n_dst, int32_t K, int32_t Cs, int32_t Cd, int32_t math) {
#ifdef U2_WITH_TC
    if (math != U2_MATH_FP32) return u2_conv_tc_scratch_bytes(n_dst, K, Cs, Cd, math);
#endif
    (void)n_dst; (void)K; (void)Cs; (void)Cd; (void)math;
    return 0;
}

static int check_common(const void *X, const void *W, const void *table, const void *Y, int32_t Cs, int32_t Cd, int32_t K,
                        int64_t ld, int64_t n_dst, const char *who) {
    (void)W;  // NULL = the scratch buffer already holds the pre-tiled weights (u2_conv_pretile)
    U2_CHECK_ARG(X && table && Y, "%s: null pointer", who);
    U2_CHECK_ARG(Cs > 0 && Cd > 0 && K > 0, "%s: bad shape Cs=%d Cd=%d K=%d", who, Cs, Cd, K);
    U2_CHECK_ARG(ld >= n_dst, "%s: ld < n_dst", who);
    return 0;
}

extern "C" int u2_conv_fwd(const float *X, int64_t n_src, int32_t Cs, const float *W, int32_t w_transposed,
                           const int32_t *table, int64_t ld, int64_t n_dst, int32_t K, int32_t Cd, float *Y, int32_t math,
                           void *scratch, size_t scratch_bytes, u2_stream_t stream) {
    if (check_common(X, W, table, Y, Cs, Cd, K, ld, n_dst, "u2_conv_fwd")) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (math == U2_MATH_FP32) return u2_conv_fwd_simt(X, n_src, Cs, W, w_transposed, table, ld, n_dst, K, Cd, Y, st);
#ifdef U2_WITH_TC
    if ((math == U2_MATH_TF32 || math == U2_MATH_BF16) && u2_conv_tc_supported(Cs, Cd, K, math))
        return u2_conv_fwd_tc(X, n_src, Cs, W, w_transposed, table, nullptr, ld, n_dst, K, Cd, Y, math, scratch, scratch_bytes, nullptr,
                              nullptr, st);
    U2_CHECK_ARG(math != U2_MATH_BF16, "u2_conv_fwd: shape Cs=%d Cd=%d K=%d has no bf16 kernel (caller must use TF32/FP32)", Cs, Cd, K);
    if (math == U2_MATH_TF32)  // shapes the MMA tiles cannot hold (e.g. the Cs = 4 stem conv)
        return u2_conv_fwd_simt(X, n_src, Cs, W, w_transposed, table, ld, n_dst, K, Cd, Y, st);
#endif
    (void)scratch; (void)scratch_bytes;
    u2_set_error("u2_conv_fwd: math mode %d not available in this build", math);
    return 1;
}

extern "C" int u2_conv_wgrad(const float *X, int64_t n_src, int32_t Cs, const float *dY, int64_t n_dst, int32_t Cd,
                             const int32_t *table, int64_t ld, int32_t K, float *dW, int32_t math, void *scratch,
                             size_t scratch_bytes, u2_stream_t stream) {
    if (check_common(X, dY, table, dW, Cs, Cd, K, ld, n_dst, "u2_conv_wgrad")) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (math == U2_MATH_FP32) return u2_conv_wgrad_simt(X, n_src, Cs, dY, n_dst, Cd, table, ld, K, dW, st);
#ifdef U2_WITH_TC
    if (math == U2_MATH_TF32)  // wgrad still runs the FFMA kernel (fp32-exact) in this build
        return u2_conv_wgrad_simt(X, n_src, Cs, dY, n_dst, Cd, table, ld, K, dW, st);
#endif
    (void)scratch; (void)scratch_bytes;
    u2_set_error("u2_conv_wgrad: math mode %d not available in this build", math);
    return 1;
}

extern "C" int u2_conv_wgrad_pairs_supported(int32_t Cs, int32_t Cd, int32_t K, int32_t math) {
#ifdef U2_WITH_TC
    if (math == U2_MATH_TF32) return u2_conv_tc_supported(32, Cd, K, math) && Cs % 4 == 0;
    if (math == U2_MATH_BF16) return u2_conv_tc_supported(64, Cd, K, math) && Cs % 8 == 0;
    return 0;
#else
    (void)Cs; (void)Cd; (void)K; (void)math;
    return 0;
#endif
}

extern "C" int u2_conv_wgrad_pairs(const float *Xa, int32_t Cs, const float *dYb, int32_t Cd, const int32_t *nbr, int64_t ld,
                                   int64_t n_rows, int32_t K, const int32_t *flat, const int32_t *nbsizes, int32_t swap,
                                   float *dW, int32_t math, u2_stream_t stream) {
#ifdef U2_WITH_TC
    U2_CHECK_ARG(Xa && dYb && nbr && flat && nbsizes && dW, "u2_conv_wgrad_pairs: null pointer");
    U2_CHECK_ARG(u2_conv_wgrad_pairs_supported(Cs, Cd, K, math), "u2_conv_wgrad_pairs: unsupported shape/math");
    return u2_conv_wgrad_tc(Xa, Cs, dYb, Cd, nbr, ld, n_rows, K, flat, nbsizes, swap, dW, math, -1, nullptr, (cudaStream_t)stream);
#else
    u2_set_error("u2_conv_wgrad_pairs: built without the tcgen05 path");
    return 1;
#endif
}

// Same, with a hint: offset `dense_k` (>= 0) of this map pairs row j with row j for EVERY j < n_rows — the centre tap of a
// submanifold map, or the only offset of a 1x1x1 conv / Linear layer (identity map); both matrices then have n_rows rows.
// Those pairs' operand rows are contiguous and are streamed with 2-D TMA tile loads instead of per-row gathers.
// dense_ok (device int32, may be NULL = trusted): 0 makes the kernel take the gather path for that offset after all.
extern "C" int u2_conv_wgrad_pairs_dense(const float *Xa, int32_t Cs, const float *dYb, int32_t Cd, const int32_t *nbr, int64_t ld,
                                         int64_t n_rows, int32_t K, const int32_t *flat, const int32_t *nbsizes, int32_t swap,
                                         float *dW, int32_t math, int32_t dense_k, const int32_t *dense_ok, u2_stream_t stream) {
#ifdef U2_WITH_TC
    U2_CHECK_ARG(Xa && dYb && nbr && flat && nbsizes && dW, "u2_conv_wgrad_pairs_dense: null pointer");
    U2_CHECK_ARG(u2_conv_wgrad_pairs_supported(Cs, Cd, K, math), "u2_conv_wgrad_pairs_dense: unsupported shape/math");
    return u2_conv_wgrad_tc(Xa, Cs, dYb, Cd, nbr, ld, n_rows, K, flat, nbsizes, swap, dW, math, dense_k, dense_ok, (cudaStream_t)stream);
#else
    u2_set_error("u2_conv_wgrad_pairs_dense: built without the tcgen05 path");
    return 1;
#endif
}

extern "C" int u2_conv_fwd_perm(const float *X, int64_t n_src, int32_t Cs, const float *W, int32_t w_transposed,
                                const int32_t *tableP, const int32_t *perm, const float *Yadd, int64_t ld,
                                int64_t n_dst, int32_t K, int32_t Cd, float *Y, int32_t math, void *scratch,
                                size_t scratch_bytes, u2_stream_t stream) {
#ifdef U2_WITH_TC
    if (check_common(X, W, tableP, Y, Cs, Cd, K, ld, n_dst, "u2_conv_fwd_perm")) return 1;
    U2_CHECK_ARG(perm != nullptr, "u2_conv_fwd_perm: null perm");
    U2_CHECK_ARG((math == U2_MATH_TF32 || math == U2_MATH_BF16) && u2_conv_tc_supported(Cs, Cd, K, math),
                 "u2_conv_fwd_perm: unsupported shape/math");
    return u2_conv_fwd_tc(X, n_src, Cs, W, w_transposed, tableP, perm, ld, n_dst, K, Cd, Y, math, scratch, scratch_bytes,
                          nullptr, Yadd, (cudaStream_t)stream);
#else
    u2_set_error("u2_conv_fwd_perm: built without the tcgen05 path");
    return 1;
#endif
}

// Conv forward that also leaves the per-warp column sums of Y behind (fused BatchNorm statistics):
// tile_stats fp32 [4 * ceil(rows / 128)][2][Cd], rows = ld with a row permutation, n_dst without.
extern "C" size_t u2_conv_tile_stats_parts(int64_t rows) { return (size_t)(4 * ((rows + 127) / 128)); }

extern "C" int u2_conv_fwd_stats(const float *X, int64_t n_src, int32_t Cs, const float *W, int32_t w_transposed,
                                 const int32_t *table, const int32_t *perm, int64_t ld, int64_t n_dst, int32_t K, int32_t Cd,
                                 float *Y, int32_t math, void *scratch, size_t scratch_bytes, float *tile_stats,
                                 size_t tile_stats_bytes, u2_stream_t stream) {
#ifdef U2_WITH_TC
    if (check_common(X, W, table, Y, Cs, Cd, K, ld, n_dst, "u2_conv_fwd_stats")) return 1;
    U2_CHECK_ARG((math == U2_MATH_TF32 || math == U2_MATH_BF16) && u2_conv_tc_supported(Cs, Cd, K, math),
                 "u2_conv_fwd_stats: unsupported shape/math");
    U2_CHECK_ARG(tile_stats && tile_stats_bytes >= u2_conv_tile_stats_parts(perm ? ld : n_dst) * 2 * Cd * sizeof(float),
                 "u2_conv_fwd_stats: tile_stats too small");
    return u2_conv_fwd_tc(X, n_src, Cs, W, w_transposed, table, perm, ld, n_dst, K, Cd, Y, math, scratch, scratch_bytes,
                          tile_stats, nullptr, (cudaStream_t)stream);
#else
    u2_set_error("u2_conv_fwd_stats: built without the tcgen05 path");
    return 1;
#endif
}

// Both weight blobs of a layer in one call: `blob_fwd` for u2_conv_fwd* with W[k] (Cs = Cin, Cd = Cout), `blob_dgrad` for the
// input-gradient conv with W[k]^T (Cs = Cout, Cd = Cin); either may be NULL.  Each u2_conv_scratch_bytes(.) large.  The convs
// then take W = NULL and the blob as `scratch`: weights are re-tiled once per optimizer step, not once per launch.
extern "C" int u2_conv_pretile(const float *W, int32_t K, int32_t Cin, int32_t Cout, int32_t math, void *blob_fwd,
                               void *blob_dgrad, u2_stream_t stream) {
#ifdef U2_WITH_TC
    U2_CHECK_ARG(math == U2_MATH_TF32 || math == U2_MATH_BF16, "u2_conv_pretile: math mode %d has no blobs", math);
    if (blob_fwd && u2_conv_pretile_tc(W, 0, K, Cin, Cout, math, blob_fwd, (cudaStream_t)stream)) return 1;
    if (blob_dgrad && u2_conv_pretile_tc(W, 1, K, Cout, Cin, math, blob_dgrad, (cudaStream_t)stream)) return 1;
    return 0;
#else
    (void)W; (void)K; (void)Cin; (void)Cout; (void)math; (void)blob_fwd; (void)blob_dgrad; (void)stream;
    u2_set_error("u2_conv_pretile: built without the tcgen05 path");
    return 1;
#endif
}

extern "C" size_t u2_conv_pretile_plan_bytes(int32_t n_jobs) {
#ifdef U2_WITH_TC
    return u2_conv_pretile_plan_bytes_tc(n_jobs);
#else
    (void)n_jobs;
    return 0;
#endif
}

extern "C" int u2_conv_pretile_plan(int32_t n_jobs, const uint64_t *W, const uint64_t *blob, const int32_t *K, const int32_t *Cs,
                                    const int32_t *Cd, const int32_t *w_transposed, int32_t math, void *plan_host,
                                    size_t plan_bytes, int64_t *n_blocks) {
#ifdef U2_WITH_TC
    return u2_conv_pretile_plan_tc(n_jobs, W, blob, K, Cs, Cd, w_transposed, math, plan_host, plan_bytes, n_blocks);
#else
    (void)n_jobs; (void)W; (void)blob; (void)K; (void)Cs; (void)Cd; (void)w_transposed; (void)math; (void)plan_host; (void)plan_bytes; (void)n_blocks;
    u2_set_error("u2_conv_pretile_plan: built without the tcgen05 path");
    return 1;
#endif
}

extern "C" int u2_conv_pretile_run(const void *plan_dev, int32_t n_jobs, int64_t n_blocks, u2_stream_t stream) {
#ifdef U2_WITH_TC
    return u2_conv_pretile_run_tc(plan_dev, n_jobs, n_blocks, (cudaStream_t)stream);
#else
    (void)plan_dev; (void)n_jobs; (void)n_blocks; (void)stream;
    u2_set_error("u2_conv_pretile_run: built without the tcgen05 path");
    return 1;
#endif
}

extern "C" int u2_conv_tc_shape_supported(int32_t Cs, int32_t Cd, int32_t K, int32_t math) {
#ifdef U2_WITH_TC
    return u2_conv_tc_supported(Cs, Cd, K, math);
#else
    (void)Cs; (void)Cd; (void)K; (void)math;
    return 0;
#endif
}

extern "C" int u2_cast_bf16(const float *x, int64_t n, void *y, u2_stream_t stream) {
#ifdef U2_WITH_TC
    return u2_cast_bf16_impl(x, n, y, (cudaStream_t)stream);
#else
    (void)x; (void)n; (void)y; (void)stream;
    u2_set_error("u2_cast_bf16: built without the tcgen05 path");
    return 1;
#endif
}

extern "C" int u2_split_bf16x3(const float *x, int64_t n, int32_t C, void *out3, void *hi, void *lo, u2_stream_t stream) {
#ifdef U2_WITH_TC
    U2_CHECK_ARG(x && (out3 || hi || lo), "u2_split_bf16x3: null pointer");
    return u2_split_bf16x3_impl(x, n, C, out3, hi, lo, (cudaStream_t)stream);
#else
    (void)x; (void)n; (void)C; (void)out3; (void)hi; (void)lo; (void)stream;
    u2_set_error("u2_split_bf16x3: built without the tcgen05 path");
    return 1;
#endif
}
