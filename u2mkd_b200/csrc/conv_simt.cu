// Sparse convolution, fp32-exact parity mode (U2_MATH_FP32): fused gather-GEMM over the
// neighbour table, output-stationary, FFMA. Replaces the K sequential gather / cuBLAS mm /
// scatter rounds of torchsparse.backend.convolution_{forward,backward}_cuda
// (see include/u2mkd.h).  No staging buffers, no output read-modify-write, no host sync.
//
// fwd/dgrad : block = 64 rows x 64 out-channels, 256 threads, 4x4 register tile per thread,
//             (tile, offset) pairs with no valid neighbour are skipped.
// wgrad     : block = (offset k, chunk of rows, 64x64 tile of dW[k]); the chunk's valid pairs
//             are compacted in shared memory (ordered), then reduced 16 pairs at a time;
//             partial sums leave as fp32 reductions.
#include "u2_common.cuh"

#define BM 64
#define BN 64
#define BK 16

// ------------------------------------------------------------------ fwd / dgrad
template <bool WT>
__global__ void __launch_bounds__(256) conv_fwd_simt_kernel(const float *__restrict__ X, int Cs, const float *__restrict__ W,
                                                            const int *__restrict__ table, int64_t ld, int64_t n_dst, int K,
                                                            int Cd, float *__restrict__ Y) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    __shared__ int s_idx[BM];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t row0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const bool vecA = (Cs & 3) == 0 && ((uintptr_t)X & 15) == 0;
    const bool vecB = WT ? ((Cs & 3) == 0 && ((uintptr_t)W & 15) == 0) : ((Cd & 3) == 0 && ((uintptr_t)W & 15) == 0);

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

    for (int k = 0; k < K; k++) {
        int my = -1;
        if (tid < BM) {
            int64_t r = row0 + tid;
            my = (r < n_dst) ? __ldg(table + (int64_t)k * ld + r) : -1;
            s_idx[tid] = my;
        }
        if (!__syncthreads_or(my >= 0)) continue;  // also orders s_idx writes

        const float *Wk = W + (int64_t)k * Cs * Cd;
        for (int c0 = 0; c0 < Cs; c0 += BK) {
            // ---- A tile: 64 gathered rows x 16 channels, stored channel-major
            {
                const int r = tid >> 2, c4 = (tid & 3) * 4;
                const int src = s_idx[r];
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (src >= 0) {
                    const float *p = X + (int64_t)src * Cs + c0 + c4;
                    if (vecA && c0 + c4 + 3 < Cs) {
                        v = __ldg(reinterpret_cast<const float4 *>(p));
                    } else {
                        if (c0 + c4 + 0 < Cs) v.x = __ldg(p + 0);
                        if (c0 + c4 + 1 < Cs) v.y = __ldg(p + 1);
                        if (c0 + c4 + 2 < Cs) v.z = __ldg(p + 2);
                        if (c0 + c4 + 3 < Cs) v.w = __ldg(p + 3);
                    }
                }
                As[c4 + 0][r] = v.x; As[c4 + 1][r] = v.y; As[c4 + 2][r] = v.z; As[c4 + 3][r] = v.w;
            }
            // ---- B tile: 16 channels x 64 out-channels of Wk (or Wk^T)
            if (!WT) {
                const int kk = tid >> 4, n4 = (tid & 15) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c0 + kk < Cs) {
                    const float *p = Wk + (int64_t)(c0 + kk) * Cd + n0 + n4;
                    if (vecB && n0 + n4 + 3 < Cd) {
                        v = __ldg(reinterpret_cast<const float4 *>(p));
                    } else {
                        if (n0 + n4 + 0 < Cd) v.x = __ldg(p + 0);
                        if (n0 + n4 + 1 < Cd) v.y = __ldg(p + 1);
                        if (n0 + n4 + 2 < Cd) v.z = __ldg(p + 2);
                        if (n0 + n4 + 3 < Cd) v.w = __ldg(p + 3);
                    }
                }
                *reinterpret_cast<float4 *>(&Bs[kk][n4]) = v;
            } else {
                // Wk_eff[cs][cd] = W[k][cd][cs]; contiguous along cs
                const int n = tid >> 2, k4 = (tid & 3) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n0 + n < Cd) {
                    const float *p = Wk + (int64_t)(n0 + n) * Cs + c0 + k4;
                    if (vecB && c0 + k4 + 3 < Cs) {
                        v = __ldg(reinterpret_cast<const float4 *>(p));
                    } else {
                        if (c0 + k4 + 0 < Cs) v.x = __ldg(p + 0);
                        if (c0 + k4 + 1 < Cs) v.y = __ldg(p + 1);
                        if (c0 + k4 + 2 < Cs) v.z = __ldg(p + 2);
                        if (c0 + k4 + 3 < Cs) v.w = __ldg(p + 3);
                    }
                }
                Bs[k4 + 0][n] = v.x; Bs[k4 + 1][n] = v.y; Bs[k4 + 2][n] = v.z; Bs[k4 + 3][n] = v.w;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < BK; kk++) {
                const float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
                const float4 b = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w};
                const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
    const bool vecY = (Cd & 3) == 0 && ((uintptr_t)Y & 15) == 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int64_t r = row0 + ty * 4 + i;
        if (r >= n_dst) continue;
        const int n = n0 + tx * 4;
        float *p = Y + r * Cd + n;
        if (vecY && n + 3 < Cd) {
            *reinterpret_cast<float4 *>(p) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (n + j < Cd) p[j] = acc[i][j];
        }
    }
}

// ------------------------------------------------------------------ fwd, 4 input channels (the stem conv: x, y, z, intensity)
// With Cs = 4 the generic kernel above spends 3/4 of its FMAs on padding (BK = 16) and two barriers per offset.  Here the
// whole im2col row of a voxel — K offsets x 4 channels = 108 values for k = 3 — is gathered into shared memory once
// (one 16-byte load per neighbour) next to the [K*4 x 64] weight slab, then one 64 x 64 x (K*4) register-tiled product:
// same order of summation as the generic kernel (offsets ascending, channels ascending), so the results are bitwise equal.
__global__ void __launch_bounds__(256) conv_fwd_c4_kernel(const float4 *__restrict__ X, const float *__restrict__ W,
                                                          const int *__restrict__ table, int64_t ld, int64_t n_dst, int K, int Cd,
                                                          float *__restrict__ Y) {
    extern __shared__ __align__(16) float smem_c4[];
    const int KC = K * 4;
    float *As = smem_c4;                  // [KC][BM + 4]
    float *Bs = smem_c4 + KC * (BM + 4);  // [KC][BN + 4]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t row0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    for (int e = tid; e < KC * (BN / 4); e += 256) {
        const int kk = e / (BN / 4), n4 = (e % (BN / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n0 + n4 + 3 < Cd) v = __ldg(reinterpret_cast<const float4 *>(W + (int64_t)kk * Cd + n0 + n4));
        *reinterpret_cast<float4 *>(&Bs[kk * (BN + 4) + n4]) = v;
    }
    for (int e = tid; e < K * BM; e += 256) {
        const int r = e % BM, k = e / BM;
        const int64_t row = row0 + r;
        const int src = row < n_dst ? __ldg(table + (int64_t)k * ld + row) : -1;
        const float4 v = src >= 0 ? __ldg(X + src) : make_float4(0.f, 0.f, 0.f, 0.f);
        As[(k * 4 + 0) * (BM + 4) + r] = v.x; As[(k * 4 + 1) * (BM + 4) + r] = v.y;
        As[(k * 4 + 2) * (BM + 4) + r] = v.z; As[(k * 4 + 3) * (BM + 4) + r] = v.w;
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
#pragma unroll 4
    for (int kk = 0; kk < KC; kk++) {
        const float4 a = *reinterpret_cast<const float4 *>(&As[kk * (BM + 4) + ty * 4]);
        const float4 b = *reinterpret_cast<const float4 *>(&Bs[kk * (BN + 4) + tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int64_t r = row0 + ty * 4 + i;
        if (r >= n_dst) continue;
        *reinterpret_cast<float4 *>(Y + r * Cd + n0 + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
}

// ------------------------------------------------------------------ wgrad
#define WG_ROWS 2048

__global__ void __launch_bounds__(256) conv_wgrad_simt_kernel(const float *__restrict__ X, int Cs, const float *__restrict__ dY,
                                                              int64_t n_dst, int Cd, const int *__restrict__ table, int64_t ld,
                                                              float *__restrict__ dW, int tiles_n) {
    __shared__ int s_src[WG_ROWS];
    __shared__ int s_dst[WG_ROWS];
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    __shared__ int s_warp[8];
    __shared__ int s_total;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;
    const int k = blockIdx.z;
    const int64_t r0 = (int64_t)blockIdx.y * WG_ROWS;
    const int m0 = (blockIdx.x / tiles_n) * BM;  // Cs tile
    const int n0 = (blockIdx.x % tiles_n) * BN;  // Cd tile

    // ---- ordered compaction of the chunk's valid (src,dst) pairs
    if (tid == 0) s_total = 0;
    __syncthreads();
    for (int base = 0; base < WG_ROWS; base += 256) {
        const int64_t r = r0 + base + tid;
        const int src = (r < n_dst) ? __ldg(table + (int64_t)k * ld + r) : -1;
        const unsigned b = __ballot_sync(0xffffffffu, src >= 0);
        if (lane == 0) s_warp[warp] = __popc(b);
        __syncthreads();
        int off = s_total;
        for (int w = 0; w < warp; w++) off += s_warp[w];
        if (src >= 0) {
            const int pos = off + __popc(b & ((1u << lane) - 1));
            s_src[pos] = src;
            s_dst[pos] = (int)(r - r0);
        }
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < 8; w++) t += s_warp[w];
            s_total += t;
        }
        __syncthreads();
    }
    const int total = s_total;
    if (total == 0) return;

    const bool vecA = (Cs & 3) == 0 && ((uintptr_t)X & 15) == 0;
    const bool vecB = (Cd & 3) == 0 && ((uintptr_t)dY & 15) == 0;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

    for (int p0 = 0; p0 < total; p0 += BK) {
        const int pr = tid >> 4, c4 = (tid & 15) * 4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p0 + pr < total) {
            const float *pa = X + (int64_t)s_src[p0 + pr] * Cs + m0 + c4;
            if (vecA && m0 + c4 + 3 < Cs) {
                a = __ldg(reinterpret_cast<const float4 *>(pa));
            } else {
                if (m0 + c4 + 0 < Cs) a.x = __ldg(pa + 0);
                if (m0 + c4 + 1 < Cs) a.y = __ldg(pa + 1);
                if (m0 + c4 + 2 < Cs) a.z = __ldg(pa + 2);
                if (m0 + c4 + 3 < Cs) a.w = __ldg(pa + 3);
            }
            const float *pb = dY + (r0 + s_dst[p0 + pr]) * Cd + n0 + c4;
            if (vecB && n0 + c4 + 3 < Cd) {
                b = __ldg(reinterpret_cast<const float4 *>(pb));
            } else {
                if (n0 + c4 + 0 < Cd) b.x = __ldg(pb + 0);
                if (n0 + c4 + 1 < Cd) b.y = __ldg(pb + 1);
                if (n0 + c4 + 2 < Cd) b.z = __ldg(pb + 2);
                if (n0 + c4 + 3 < Cd) b.w = __ldg(pb + 3);
            }
        }
        *reinterpret_cast<float4 *>(&As[pr][c4]) = a;
        *reinterpret_cast<float4 *>(&Bs[pr][c4]) = b;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            const float4 av4 = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            const float4 bv4 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            const float av[4] = {av4.x, av4.y, av4.z, av4.w};
            const float bv[4] = {bv4.x, bv4.y, bv4.z, bv4.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *out = dW + (int64_t)k * Cs * Cd;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int m = m0 + ty * 4 + i;
        if (m >= Cs) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int n = n0 + tx * 4 + j;
            if (n < Cd) atomicAdd(out + (int64_t)m * Cd + n, acc[i][j]);
        }
    }
}

int u2_conv_fwd_simt(const float *X, int64_t n_src, int32_t Cs, const float *W, int32_t w_transposed, const int32_t *table,
                     int64_t ld, int64_t n_dst, int32_t K, int32_t Cd, float *Y, cudaStream_t st) {
    (void)n_src;
    if (n_dst == 0) return 0;
    dim3 grid((unsigned)u2_ceil_div(n_dst, BM), (unsigned)u2_ceil_div(Cd, BN));
    if (!w_transposed && Cs == 4 && Cd % BN == 0 && K * 4 <= 128 && (((uintptr_t)X | (uintptr_t)W | (uintptr_t)Y) & 15) == 0) {
        const size_t smem = (size_t)K * 4 * ((BM + 4) + (BN + 4)) * sizeof(float);
        U2_CUDA_OK(cudaFuncSetAttribute(conv_fwd_c4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_fwd_c4_kernel<<<grid, 256, smem, st>>>((const float4 *)X, W, table, ld, n_dst, K, Cd, Y);
        U2_LAUNCH_OK();
        return 0;
    }
    if (w_transposed)
        conv_fwd_simt_kernel<true><<<grid, 256, 0, st>>>(X, Cs, W, table, ld, n_dst, K, Cd, Y);
    else
        conv_fwd_simt_kernel<false><<<grid, 256, 0, st>>>(X, Cs, W, table, ld, n_dst, K, Cd, Y);
    U2_LAUNCH_OK();
    return 0;
}

int u2_conv_wgrad_simt(const float *X, int64_t n_src, int32_t Cs, const float *dY, int64_t n_dst, int32_t Cd,
                       const int32_t *table, int64_t ld, int32_t K, float *dW, cudaStream_t st) {
    (void)n_src;
    U2_CUDA_OK(cudaMemsetAsync(dW, 0, (size_t)K * Cs * Cd * sizeof(float), st));
    if (n_dst == 0) return 0;
    const int tiles_m = (int)u2_ceil_div(Cs, BM), tiles_n = (int)u2_ceil_div(Cd, BN);
    dim3 grid((unsigned)(tiles_m * tiles_n), (unsigned)u2_ceil_div(n_dst, WG_ROWS), (unsigned)K);
    U2_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "u2_conv_wgrad: grid too large");
    conv_wgrad_simt_kernel<<<grid, 256, 0, st>>>(X, Cs, dY, n_dst, Cd, table, ld, dW, tiles_n);
    U2_LAUNCH_OK();
    return 0;
}
