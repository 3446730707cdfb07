/*
 * u2mkd.h — C-ABI of the B200-native (sm_100a) LiDAR point-voxel hot path.
 *
 * This is the drop-in boundary: the entry points below are what the
 * reference's native binding for this path — torchsparse v1.4.0 `torchsparse.backend`
 * [TS v1.4.0 torchsparse/backend/pybind_cuda.cpp], pinned at
 * /root/reference/README.md:44-48 and called from core/models/utils.py and
 * core/models/build_blocks.py through torchsparse.nn.functional — would bind instead.
 *
 * Conventions
 *   - extern "C", plain pointers + sizes; no torch / C++ types.
 *   - every pointer is DEVICE memory owned by the caller unless stated otherwise;
 *     the library never allocates: scratch is caller-provided, its size queried first.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*) and
 *     re-entrant; no global mutable state besides the thread-local error string.
 *   - returns 0 on success, non-zero on error; u2_last_error() explains (thread-local).
 *   - coordinates are int32 [N,4] rows (x, y, z, batch) — batch LAST
 *     (core/models/semantickitti/spvcnn.py:90, core/models/utils.py:17).
 *   - "miss" sentinel is -1 everywhere (core/models/utils.py:97-98).
 */
#ifndef U2MKD_H_
#define U2MKD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *u2_stream_t; /* cudaStream_t */

/* arithmetic modes of the conv GEMMs */
enum {
    U2_MATH_FP32 = 0, /* FFMA, fp32-exact parity mode (rel 1e-4 bar)             */
    U2_MATH_TF32 = 1, /* tcgen05 kind::tf32, fp32 storage (rel 2e-2 bar)          */
    U2_MATH_BF16 = 2  /* tcgen05 kind::f16 on bf16 operands, fp32 accumulate (rel 2e-2 bar): the feature /
                         gradient pointers of the conv entry points then address bf16 rows made by
                         u2_cast_bf16; weights, outputs and weight gradients stay fp32             */
};

const char *u2_last_error(void);
int u2_version(void);
/* 1 if the library was built with the tcgen05 (TF32/BF16) conv kernels */
int u2_has_tensor_core_path(void);

/* ---- hashing: replaces backend.hash_cuda / backend.kernel_hash_cuda
 * [TS backend/hash/hash_cuda.cu]; callers core/models/utils.py:19,43,49,86,92.
 * offsets == NULL: out[n].  else offsets int32 [K,3]: out[K,n] = hash(x+ox, y+oy, z+oz, b). */
int u2_hash(const int32_t *coords, int64_t n, const int32_t *offsets, int32_t K, int64_t *out,
            u2_stream_t stream);

/* ---- hash table: replaces backend.hash_query_cuda
 * [TS backend/others/query_cuda.cu + backend/hashmap/hashmap_cuda.cu];
 * callers core/models/utils.py:21,50,93,135 through spf.sphashquery.
 * Open addressing, 16-byte slots {int64 key, uint32 value}, capacity = pow2 >= 2*n_keys.
 * Duplicate keys keep the SMALLEST index (== the CPU backend's "first insert wins").
 * Key 0xFFFFFFFFFFFFFFFF is reserved (never produced by u2_hash: hashes are 60-bit). */
size_t u2_hash_table_bytes(int64_t n_keys);
int u2_hash_table_build(const int64_t *keys, int64_t n_keys, void *table, size_t table_bytes,
                        u2_stream_t stream);
/* out[i] = index of queries[i] among keys, or -1 */
int u2_hash_table_query(const void *table, size_t table_bytes, const int64_t *queries, int64_t nq,
                        int64_t *out, u2_stream_t stream);

/* ---- count: replaces backend.count_cuda [TS backend/others/count_cuda.cu];
 * core/models/utils.py:22,51.  out int32 [n_out] = histogram of idx[i] in [0,n_out). */
int u2_count(const int32_t *idx, int64_t n, int32_t *out, int64_t n_out, u2_stream_t stream);

/* ---- voxelize (scatter-mean): replaces backend.voxelize_{forward,backward}_cuda
 * [TS backend/voxelize/voxelize_cuda.cu]; core/models/utils.py:24,26,58.
 * fwd: out[n_vox,C] = mean over points i with idx[i]==v of feats[i,:]  (idx<0 dropped)
 * bwd: gfeats[i,:] = gout[idx[i],:] / counts[idx[i]]   (0 for dropped points)        */
int u2_voxelize_fwd(const float *feats, int64_t n_pts, int32_t C, const int32_t *idx,
                    const int32_t *counts, float *out, int64_t n_vox, u2_stream_t stream);
int u2_voxelize_bwd(const float *gout, int64_t n_vox, int32_t C, const int32_t *idx,
                    const int32_t *counts, float *gfeats, int64_t n_pts, u2_stream_t stream);

/* ---- trilinear weights: replaces the ~20 torch kernels of spf.calc_ti_weights
 * [TS nn/functional/devoxelize.py]; core/models/utils.py:94-95.
 * coords fp32 [n_pts,4]; idx_kn int64 [8,n_pts]; writes weights fp32 [8,n_pts] (same
 * layout as the reference returns).                                                   */
int u2_ti_weights(const float *coords, const int64_t *idx_kn, int64_t n_pts, float scale,
                  float *weights_kn, u2_stream_t stream);

/* ---- devoxelize: replaces backend.devoxelize_{forward,backward}_cuda
 * [TS backend/devoxelize/devoxelize_cuda.cu]; core/models/utils.py:99,111.
 * fwd: out[i,:] = sum_k w[i,k] * feats[idx[i,k],:]   (idx<0 skipped); idx int32 [n_pts,8]
 * bwd: gfeats[idx[i,k],:] += w[i,k] * gout[i,:]      (gfeats zeroed by the call)       */
int u2_devoxelize_fwd(const float *feats, int64_t n_vox, int32_t C, const int32_t *idx,
                      const float *w, int64_t n_pts, float *out, u2_stream_t stream);
int u2_devoxelize_bwd(const float *gout, int64_t n_pts, int32_t C, const int32_t *idx,
                      const float *w, float *gfeats, int64_t n_vox, u2_stream_t stream);

/* ---- strided-conv output coordinates: replaces F.spdownsample
 * [TS nn/functional/downsample.py] for the stride in {1, kernel_size} case used by every
 * U2MKD strided conv (core/models/build_blocks.py:25-29 with ks=2, stride=2):
 * out = unique rows of (floor(c / s) * s, b), sorted by (b, x, y, z).
 * n_out_dev: device int64 scalar; out_coords has room for n rows. scratch from
 * u2_downsample_scratch_bytes(n).  Requires 0 <= x,y,z < 2^18 and 0 <= b < 2^10.       */
size_t u2_downsample_scratch_bytes(int64_t n);
int u2_downsample_coords(const int32_t *coords, int64_t n, int32_t sx, int32_t sy, int32_t sz,
                         int32_t *out_coords, int64_t *n_out_dev, void *scratch,
                         size_t scratch_bytes, u2_stream_t stream);

/* ---- kernel map: replaces the sphash / sphashquery / nonzero sequence inside F.conv3d
 * [TS nn/functional/conv.py]; SURVEY.md §3.3.
 * nbr  int32 [K, ld_out]: nbr[k][o]  = input row i with coord(i) == coord(o)+offset_k, else -1
 * nbrT int32 [K, ld_in ]: nbrT[k][i] = output row o of that same pair, else -1
 * nbsizes int32 [K]      : pairs per offset (device)
 * scratch: u2_hash_table_bytes(n_in) + 8*n_in bytes.                                    */
size_t u2_kmap_scratch_bytes(int64_t n_in);
int u2_kmap_build(const int32_t *in_coords, int64_t n_in, const int32_t *out_coords,
                  int64_t n_out, const int32_t *offsets, int32_t K, int32_t *nbr, int64_t ld_out,
                  int32_t *nbrT, int64_t ld_in, int32_t *nbsizes, void *scratch,
                  size_t scratch_bytes, u2_stream_t stream);
/* ---- sparse convolution: replaces backend.convolution_{forward,backward}_cuda
 * [TS backend/convolution/convolution_cuda.cu]; core/models/build_blocks.py:25-77.
 * One fused gather-GEMM kernel, output-stationary over the neighbour table:
 *   Y[r,:] = sum_k X[table[k][r],:] @ Wk        Wk = W[k] (Cs x Cd)       if !w_transposed
 *                                               Wk = W[k]^T, W [K,Cd,Cs]   if  w_transposed
 * forward (not transposed): table = nbr,  X = input,  W = kernel
 * forward (transposed conv): table = nbrT, X = input,  W = kernel
 * dgrad: the other table, X = grad_out, w_transposed = 1.
 * scratch: u2_conv_scratch_bytes(...) (0 allowed for FP32 mode).                         */
size_t u2_conv_scratch_bytes(int64_t n_dst, int32_t K, int32_t Cs, int32_t Cd, int32_t math);
int u2_conv_fwd(const float *X, int64_t n_src, int32_t Cs, const float *W, int32_t w_transposed,
                const int32_t *table, int64_t ld, int64_t n_dst, int32_t K, int32_t Cd, float *Y,
                int32_t math, void *scratch, size_t scratch_bytes, u2_stream_t stream);
/* wgrad: dW[k] (Cs x Cd) = sum_r X[table[k][r],:]^T (outer) dY[r,:], with the same table the
 * forward used (nbr for a regular conv, nbrT for a transposed one). dW is zeroed by the call. */
int u2_conv_wgrad(const float *X, int64_t n_src, int32_t Cs, const float *dY, int64_t n_dst,
                  int32_t Cd, const int32_t *table, int64_t ld, int32_t K, float *dW,
                  int32_t math, void *scratch, size_t scratch_bytes, u2_stream_t stream);

/* ---- compacted pair list of a kernel map (the reference's (k, out)-ordered nbmaps
 * [TS nn/functional/conv.py: nonzero(results != -1)], without the host sync):
 * flat[j] = k*ld + out_row of the j-th valid entry of nbr [K, ld], ascending. `flat` needs room
 * for every valid entry (<= total = K*ld). scratch from u2_kmap_pairs_scratch_bytes(total).      */
size_t u2_kmap_pairs_scratch_bytes(int64_t total);
int u2_kmap_pairs(const int32_t *nbr, int64_t total, int32_t *flat, void *scratch, size_t scratch_bytes,
                  u2_stream_t stream);

/* ---- wgrad over the compacted pair list (tcgen05 path; same result as u2_conv_wgrad):
 * dW[k] (Cs x Cd) = sum over pairs (i, o) of offset k of  Xa[a,:]^T (outer) dYb[b,:],
 * (a, b) = (i, o) for a regular conv (swap = 0: Xa = layer input, dYb = grad of the output) and
 * (o, i) for a transposed conv (swap = 1). n_rows = number of rows on the `out` side of nbr.
 * dW is zeroed by the call.                                                                        */
int u2_conv_wgrad_pairs_supported(int32_t Cs, int32_t Cd, int32_t K, int32_t math);
int u2_conv_wgrad_pairs(const float *Xa, int32_t Cs, const float *dYb, int32_t Cd, const int32_t *nbr, int64_t ld,
                        int64_t n_rows, int32_t K, const int32_t *flat, const int32_t *nbsizes, int32_t swap,
                        float *dW, int32_t math, u2_stream_t stream);
/* u2_conv_wgrad_pairs with a hint: offset dense_k (>= 0) pairs row j with row j for every j < n_rows (centre tap of a
 * submanifold map; the single offset of a 1x1x1 conv / Linear layer) — those operand rows are contiguous and go through 2-D
 * TMA tile loads instead of per-row gathers (bf16 mode).  dense_k = -1: identical to u2_conv_wgrad_pairs.  dense_ok: device
 * int32 flag (NULL = trusted), 0 = the property does not hold for this map (duplicate coordinates): gather path.   */
int u2_conv_wgrad_pairs_dense(const float *Xa, int32_t Cs, const float *dYb, int32_t Cd, const int32_t *nbr, int64_t ld,
                              int64_t n_rows, int32_t K, const int32_t *flat, const int32_t *nbsizes, int32_t swap,
                              float *dW, int32_t math, int32_t dense_k, const int32_t *dense_ok, u2_stream_t stream);

/* ---- mask-sorted tile order for the tensor-core conv (no reference counterpart: the reference
 * walks offset-major pair lists; this is the output-stationary equivalent of that compaction).
 * perm_out int32 [ld]: tile row -> original row (-1 in the padding); tableP int32 [K, ld]:
 * table with its rows permuted accordingly.  bitpos_host: HOST array int32 [K], bit position of
 * each offset in the sort key (rare offsets high).  perm_in != NULL: reuse that permutation
 * (no sort), only permute `table`.  scratch from u2_kmap_sort_scratch_bytes(n_rows).            */
size_t u2_kmap_sort_scratch_bytes(int64_t n_rows);
/* tile_mask uint32 [ld/128] (may be NULL): bit k set iff 128-row tile j of tableP uses offset k. */
int u2_kmap_sort_rows(const int32_t *table, int64_t ld, int64_t n_rows, int32_t K, const int32_t *bitpos_host,
                      const int32_t *perm_in, int32_t *perm_out, int32_t *tableP, uint32_t *tile_mask, void *scratch,
                      size_t scratch_bytes, u2_stream_t stream);
/* 1 if (Cs, Cd, K) runs on the tcgen05 kernels in this math mode (else the FFMA kernel is used) */
int u2_conv_tc_shape_supported(int32_t Cs, int32_t Cd, int32_t K, int32_t math);
/* u2_conv_fwd over a permuted table: tile row j reads tableP[k][j] and writes Y[perm[j], :].
 * Yadd fp32 [n_dst, Cd] (may be NULL) is added to the result in the epilogue: when the conv is an input-gradient conv
 * whose input had a second consumer (the shortcut of core/models/build_blocks.py:79-84), the other consumer's gradient
 * rides along instead of a separate accumulation pass.
 * W == NULL in u2_conv_fwd / _perm / _stats: `scratch` already holds the weights re-tiled by u2_conv_pretile.          */
int u2_conv_fwd_perm(const float *X, int64_t n_src, int32_t Cs, const float *W, int32_t w_transposed,
                     const int32_t *tableP, const int32_t *perm, const float *Yadd, int64_t ld, int64_t n_dst,
                     int32_t K, int32_t Cd, float *Y, int32_t math, void *scratch, size_t scratch_bytes,
                     u2_stream_t stream);
/* Re-tile the fp32 parameter W [K, Cin, Cout] into the blobs the tcgen05 kernels stream (replaces the per-launch
 * re-tiling of [TS backend/convolution/convolution_cuda.cu], which reads `kernel` as is): blob_fwd for W[k]
 * (u2_conv_scratch_bytes(., K, Cin, Cout, math) bytes), blob_dgrad for W[k]^T (.., Cout, Cin, ..); either may be NULL. */
int u2_conv_pretile(const float *W, int32_t K, int32_t Cin, int32_t Cout, int32_t math, void *blob_fwd, void *blob_dgrad,
                    u2_stream_t stream);
/* The same for every conv parameter of a model in ONE launch per optimizer step (the reference re-reads `kernel` in every
 * conv launch, [TS backend/convolution/convolution_cuda.cu]; per-layer u2_conv_pretile calls were 2 small launches per layer).
 * u2_conv_pretile_plan writes a job table for n_jobs (parameter, direction) pairs into the HOST buffer plan_host
 * (u2_conv_pretile_plan_bytes(n_jobs) bytes): job i re-tiles the fp32 tensor at device address W[i] — [K, Cs, Cd], or
 * [K, Cd, Cs] when w_transposed[i] (the input-gradient direction of a [K, Cin, Cout] parameter: Cs = Cout, Cd = Cin) — into
 * the blob at device address blob[i].  bf16 blobs only.  The caller copies the table to device memory once and calls
 * u2_conv_pretile_run(plan_dev, n_jobs, *n_blocks) whenever the parameters have changed.                              */
size_t u2_conv_pretile_plan_bytes(int32_t n_jobs);
int u2_conv_pretile_plan(int32_t n_jobs, const uint64_t *W, const uint64_t *blob, const int32_t *K, const int32_t *Cs,
                         const int32_t *Cd, const int32_t *w_transposed, int32_t math, void *plan_host, size_t plan_bytes,
                         int64_t *n_blocks);
int u2_conv_pretile_run(const void *plan_dev, int32_t n_jobs, int64_t n_blocks, u2_stream_t stream);

/* ---- BatchNorm (+ fused ReLU) over feature matrices fp32 [n, C], training mode: what the reference runs as
 * torch BatchNorm1d / SyncBatchNorm + ReLU on SparseTensor.F after every conv
 * (core/models/build_blocks.py:21-84, core/models/utils.py:138-141); SURVEY.md §8(f)-2.
 * sums fp64 [2C+1] = (sum x, sum x^2, row count) — all-reduce it across ranks for SyncBatchNorm, then
 * u2_bn_apply derives mean / invstd (written to save_mean / save_invstd, running stats updated if non-NULL).
 * backward: dsum fp64 [2C] = (sum dz, sum dz*xhat) = (grad beta, grad gamma); all-reduce a copy for dx.    */
int u2_bn_supported(int32_t C);
/* the two reductions leave one fp32 row of partial sums per CTA in `scratch` (u2_bn_scratch_bytes(C), 16-byte aligned)
 * and fold them in a second small kernel: no same-address atomics from every CTA */
size_t u2_bn_scratch_bytes(int32_t C);
int u2_bn_stats(const float *x, int64_t n, int32_t C, double *sums, void *scratch, size_t scratch_bytes,
                u2_stream_t stream);
int u2_bn_apply(const float *x, int64_t n, int32_t C, const double *sums, float eps, float momentum,
                const float *gamma, const float *beta, int32_t relu, float *y, float *save_mean, float *save_invstd,
                float *running_mean, float *running_var, u2_stream_t stream);
int u2_bn_bwd_reduce(const float *dy, const float *x, int64_t n, int32_t C, const float *mean, const float *invstd,
                     const float *gamma, const float *beta, int32_t relu, const float *out_mask, double *dsum,
                     void *scratch, size_t scratch_bytes, u2_stream_t stream);
int u2_bn_bwd_apply(const float *dy, const float *x, int64_t n, int32_t C, const float *mean, const float *invstd,
                    const float *gamma, const float *beta, const double *dsum, const double *count_dev, int32_t relu,
                    float *dx, u2_stream_t stream);

/* ---- conv -> BatchNorm(+ReLU) fusion (the Sequential(Conv3d, BatchNorm, ReLU) triples of
 * core/models/build_blocks.py:21-84).  u2_conv_fwd_stats = u2_conv_fwd / u2_conv_fwd_perm (perm may be NULL) whose epilogue
 * also stores, per warp of each 128-row tile, the column sums and sums of squares of Y:
 * tile_stats fp32 [u2_conv_tile_stats_parts(rows)][2][Cd], rows = ld with perm, n_dst without; needs Cd tiles % 32 == 0.
 * u2_bn_stats_from_tiles folds them into the fp64 [2C+1] sums buffer of u2_bn_apply (no pass over Y).
 * The *_dual variants also emit the bf16 copy the next conv (forward) / the conv backward consumes, so that
 * no separate u2_cast_bf16 pass is needed; in u2_bn_bwd_apply_dual either output may be NULL.
 * residual != NULL: y = relu(bn(x) + residual), the tail of core/models/build_blocks.py ResidualBlock.forward; its
 * backward passes the saved y as out_mask (ReLU mask = y > 0 instead of the recomputed sign of bn(x)) and gets the
 * gradient of the residual input in dresidual (= the masked output gradient).                                   */
size_t u2_conv_tile_stats_parts(int64_t rows);
int u2_conv_fwd_stats(const float *X, int64_t n_src, int32_t Cs, const float *W, int32_t w_transposed,
                      const int32_t *table, const int32_t *perm, int64_t ld, int64_t n_dst, int32_t K, int32_t Cd,
                      float *Y, int32_t math, void *scratch, size_t scratch_bytes, float *tile_stats,
                      size_t tile_stats_bytes, u2_stream_t stream);
int u2_bn_stats_from_tiles(const float *tile_stats, int64_t n_parts, int32_t C, int64_t n, double *sums,
                           u2_stream_t stream);
int u2_bn_apply_dual(const float *x, int64_t n, int32_t C, const double *sums, float eps, float momentum,
                     const float *gamma, const float *beta, int32_t relu, const float *residual, float *y, void *y_bf16,
                     float *save_mean, float *save_invstd, float *running_mean, float *running_var, u2_stream_t stream);
int u2_bn_bwd_apply_dual(const float *dy, const float *x, int64_t n, int32_t C, const float *mean, const float *invstd,
                         const float *gamma, const float *beta, const double *dsum, const double *count_dev,
                         int32_t relu, const float *out_mask, float *dresidual, float *dx, void *dx_bf16,
                         u2_stream_t stream);

/* fp32 -> bf16 (round to nearest even), n % 8 == 0: operand conversion for U2_MATH_BF16 */
int u2_cast_bf16(const float *x, int64_t n, void *y, u2_stream_t stream);

/* Operand split of the "bf16x3" conv mode (fp32-grade results on the bf16 tensor-core kernels; the reference arithmetic is
 * fp32, [TS backend/convolution/convolution_cuda.cu] at::mm): x fp32 [n, C] -> hi = bf16(x), lo = bf16(x - hi).
 * out3 bf16 [n, 3C] = [hi | lo | hi] (rows for u2_conv_fwd* against weights stacked as [Whi; Whi; Wlo], i.e. the three
 * significant terms of (hi + lo)(Whi + Wlo)), hi / lo bf16 [n, C] contiguous (u2_conv_wgrad_pairs operands); any of the
 * three outputs may be NULL.  C % 8 == 0.                                                                              */
int u2_split_bf16x3(const float *x, int64_t n, int32_t C, void *out3, void *hi, void *lo, u2_stream_t stream);

/* ---- coordinate table of one coordinate set, built once per tensor stride and shared by the kernel maps and the
 * point<->voxel index queries of that stride (replaces the per-call table inside [TS nn/functional/query.py]
 * sphashquery, call sites core/models/utils.py:50,93).  table: u2_hash_table_bytes(n) bytes.  u2_coord_table_query:
 * out int64 [K, nq], out[k][i] = row of (q[i] + offsets[k]) or -1; offsets int32 [K,3] device (NULL with K = 1).
 * u2_kmap_build accepts such a table as `scratch` with scratch_bytes = 0.                                               */
int u2_coord_table_build(const int32_t *coords, int64_t n, void *table, size_t table_bytes, u2_stream_t stream);
int u2_coord_table_query(const void *table, size_t table_bytes, const int32_t *qcoords, int64_t nq, const int32_t *offsets,
                         int32_t K, int64_t *out, u2_stream_t stream);

/* ---- index part of initial_voxelize in one call (core/models/utils.py:19-25): replaces sphash -> torch.unique ->
 * sphashquery -> spcount -> round(spvoxelize(coords)).  coords int32 [n,4] (already floored) -> idx_query int64 [n],
 * counts int32 (sized n, first n_vox valid), voxel_coords int32 [n,4] (first n_vox rows valid), n_vox_dev (device int64).
 * Voxel v is the v-th smallest FNV key, i.e. the order torch.unique(pc_hash) gives the reference.                      */
size_t u2_unique_voxelize_scratch_bytes(int64_t n);
int u2_unique_voxelize(const int32_t *coords, int64_t n, int64_t *idx_query, int32_t *counts, int32_t *voxel_coords,
                       int64_t *n_vox_dev, void *scratch, size_t scratch_bytes, u2_stream_t stream);

/* ---- SyncBatchNorm statistics exchange over NVLink peer memory: replaces the NCCL collectives torch SyncBatchNorm issues per
 * layer and pass (core/models/utils.py:138-141, train_spformer.py:77-83) by one single-block kernel that pushes the local
 * fp64 sums into every peer's mapped buffer, waits for the peers' flags and sums in rank order (all-reduce, in place).
 * peer_bufs: HOST array [world] of device pointers, peer_bufs[r] = rank r's buffer (u2_syncbn_buffer_bytes() bytes,
 * zero-filled once) as mapped into this process; seq = 1, 2, 3, ... identical on all ranks for the same exchange.        */
size_t u2_syncbn_buffer_bytes(void);
int32_t u2_syncbn_max_len(void);
int32_t u2_syncbn_max_world(void);
int u2_syncbn_exchange(double *vals, int32_t len, const void *const *peer_bufs, int32_t world, int32_t rank, uint64_t seq,
                       u2_stream_t stream);

/* ---- variable-length window attention of third_party/SparseTransformer (`sptr_cuda`), the
 * SphereFormer blocks between the down stages (core/models/sphereformer/spherical_transformer.py:165-283).
 * Points are sorted by window; window w owns rows win_off[w] .. win_off[w+1]-1 and pairs sq_off[w] + i * n_w + j.
 *
 * u2_window_pairs replaces sptr_cuda.precompute_all_cuda (src/sptr/precompute/precompute_cuda_kernel.cu:4-34, bound at
 * sptr/functional.py:166): index0_offsets / index1_offsets int32 [N], index0 / index1 int32 [M].
 *
 * u2_window_attn_fwd / _bwd replace, as ONE kernel per direction, the chain
 *   dot_prod_with_idx_all_forward_cuda | attention_step1_forward_cuda   (src/sptr/rpe/...cu:116-170, attention/...cu:4-27)
 *   scatter_softmax_csr                                                  (sptr/utils.py:81-95, torch_scatter)
 *   attention_step2_with_rel_pos_value_forward_cuda | attention_step2_forward_cuda
 * and their backward kernels.  q (already scaled), k, v, out: fp32 [N, h, head_dim]; lse fp32 [N, h] (log-sum-exp of
 * every query's scores, saved for the backward); rel_idx int32 [M, 3] and table_{q,k,v} fp32 [L, 3, h, head_dim]
 * (contextual relative position encoding) or all four NULL (pe_type 'none').  head_dim 16 or 32, L <= 64.
 * The backward writes dtable_* (per-block partials in `scratch`, then one fold); dq / dk / dv are written once per row.                              */
int u2_window_attn_supported(int32_t head_dim, int32_t L);
int u2_window_pairs(const int32_t *win_off, const int32_t *sq_off, int32_t n_windows, int32_t *index0_offsets,
                    int32_t *index1_offsets, int32_t *index0, int32_t *index1, u2_stream_t stream);
int u2_window_attn_fwd(const float *q, const float *k, const float *v, const int32_t *win_off, const int32_t *sq_off,
                       int32_t n_windows, int32_t h, int32_t head_dim, const int32_t *rel_idx, const float *table_q,
                       const float *table_k, const float *table_v, int32_t L, float *out, float *lse, u2_stream_t stream);
int u2_window_attn_bwd(const float *q, const float *k, const float *v, const int32_t *win_off, const int32_t *sq_off,
                       int32_t n_windows, int32_t h, int32_t head_dim, const int32_t *rel_idx, const float *table_q,
                       const float *table_k, const float *table_v, int32_t L, const float *out, const float *lse,
                       const float *dout, float *dq, float *dk, float *dv, float *dtable_q, float *dtable_k,
                       float *dtable_v, void *scratch, size_t scratch_bytes, u2_stream_t stream);
/* scratch of the backward: one table-gradient partial per persistent block, folded by a second small kernel (0 without tables) */
size_t u2_window_attn_bwd_scratch_bytes(int32_t n_windows, int32_t h, int32_t head_dim, int32_t L);

/* ---- point <-> pixel transforms of the student's fusion path (pure-torch loops in the reference, no
 * native boundary there; these entries replace the bodies of
 *   Point2Grid            core/models/fusion_blocks.py:217-238 and the per-scale body of
 *                         core/models/nuscenes/spvcnn_swiftnet18_spformer_tsd_full.py:455-473
 *                         (floor pixel, torch.unique(dim=0), scatter_add_, / count, sparse_coo_tensor().to_dense(), permute)
 *   Feature_Gather + masked assignment   fusion_blocks.py:241-278, ...tsd_full.py:482-494 (grid_sample align_corners=True,
 *                         zero padding; for a point seen by several cameras the LAST camera wins).
 * feats fp32 [N, C] (C % 4 == 0), coord fp32 [V, N, 2] = (x, y) in [-1, 1], mask u8 [V, N], grids fp32 [V, C, H, W],
 * counts int32 [V, H, W] (points per pixel; output of the forward, input of the backward).                              */
size_t u2_point2grid_scratch_bytes(int32_t C, int32_t V, int32_t H, int32_t W);
int u2_point2grid_fwd(const float *feats, const float *coord, const uint8_t *mask, int64_t N, int32_t C, int32_t V, int32_t H,
                      int32_t W, float *grid, int32_t *counts, void *scratch, size_t scratch_bytes, u2_stream_t stream);
int u2_point2grid_bwd(const float *dgrid, const float *coord, const uint8_t *mask, const int32_t *counts, int64_t N, int32_t C,
                      int32_t V, int32_t H, int32_t W, float *dfeats, u2_stream_t stream);
int u2_pixel_gather_fwd(const float *img, const float *coord, const uint8_t *mask, int64_t N, int32_t C, int32_t V, int32_t H,
                        int32_t W, float *out, u2_stream_t stream);
int u2_pixel_gather_bwd(const float *dout, const float *coord, const uint8_t *mask, int64_t N, int32_t C, int32_t V, int32_t H,
                        int32_t W, float *dimg, u2_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* U2MKD_H_ */
