// BatchNorm (+ optional fused ReLU) over voxel / point feature matrices fp32 [n, C], training mode.
// SURVEY.md §8(f)-2 (epilogue fusion on the operator surface) and §8(a16)/(a5): the reference applies
// torch BatchNorm1d / SyncBatchNorm + ReLU to SparseTensor.F after every conv
// (core/models/build_blocks.py:21-84, core/models/utils.py:138-141).  Pure HBM-bound passes:
//   forward : stats (1 read)  +  normalise[+ReLU] (1 read, 1 write)           vs 5 passes unfused
//   backward: reduce (2 reads) +  dx (2 reads, 1 write), ReLU mask recomputed  vs 8 passes unfused
// Channel sums are accumulated in fp32 per thread (<= 128 rows) and combined in fp64, so that
// (a) var = E[x^2] - E[x]^2 is safe and (b) the [2C+1] fp64 buffer can be all-reduced across ranks
// as is (SyncBatchNorm: one small collective per pass instead of all_gather + several kernels).
#include "u2_common.cuh"

namespace {

constexpr int BN_ROWS_PER_BLOCK = 512;

struct BnGeom {
    int tx;  // float4 columns = C / 4
    int ty;  // rows handled in parallel by one block
};

inline BnGeom bn_geom(int C) {
    BnGeom g;
    g.tx = C / 4;
    g.ty = 256 / g.tx;
    if (g.ty < 1) g.ty = 1;
    return g;
}

// sums[c] += sum_rows x[r][c];  sums[C + c] += sum_rows x[r][c]^2
__global__ void __launch_bounds__(256) bn_stats_kernel(const float4 *__restrict__ x, int64_t n, int tx, int ty,
                                                       double *__restrict__ sums) {
    extern __shared__ float4 s_red[];  // [2][ty][tx]
    const int cx = threadIdx.x % tx, ry = threadIdx.x / tx;
    const int64_t r0 = (int64_t)blockIdx.x * BN_ROWS_PER_BLOCK;
    const int64_t r1 = min(n, r0 + BN_ROWS_PER_BLOCK);
    if (blockIdx.x == 0 && threadIdx.x == 0) sums[2 * tx * 4] = (double)n;  // local row count rides along
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ry < ty) {
        for (int64_t r = r0 + ry; r < r1; r += ty) {
            const float4 v = __ldg(x + r * tx + cx);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            q.x += v.x * v.x; q.y += v.y * v.y; q.z += v.z * v.z; q.w += v.w * v.w;
        }
        s_red[ry * tx + cx] = s;
        s_red[(ty + ry) * tx + cx] = q;
    }
    __syncthreads();
    if (ry == 0) {
        for (int j = 1; j < ty; j++) {
            const float4 a = s_red[j * tx + cx], b = s_red[(ty + j) * tx + cx];
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
            q.x += b.x; q.y += b.y; q.z += b.z; q.w += b.w;
        }
        const int C = tx * 4, c = cx * 4;
        atomicAdd(sums + c + 0, (double)s.x); atomicAdd(sums + c + 1, (double)s.y);
        atomicAdd(sums + c + 2, (double)s.z); atomicAdd(sums + c + 3, (double)s.w);
        atomicAdd(sums + C + c + 0, (double)q.x); atomicAdd(sums + C + c + 1, (double)q.y);
        atomicAdd(sums + C + c + 2, (double)q.z); atomicAdd(sums + C + c + 3, (double)q.w);
    }
}

// per-channel mean / invstd from the (possibly all-reduced) sums; sums[2C] = total row count
__global__ void bn_finalize_kernel(const double *__restrict__ sums, int C, float eps, float momentum,
                                   float *__restrict__ mean, float *__restrict__ invstd, float *__restrict__ running_mean,
                                   float *__restrict__ running_var) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double cnt = sums[2 * C] > 0.0 ? sums[2 * C] : 1.0;
    const double m = sums[c] / cnt;
    double var = sums[C + c] / cnt - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
        const double unbiased = cnt > 1.0 ? var * cnt / (cnt - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

template <bool RELU>
__global__ void __launch_bounds__(256) bn_apply_kernel(const float4 *__restrict__ x, int64_t n4, int tx,
                                                       const float *__restrict__ mean, const float *__restrict__ invstd,
                                                       const float *__restrict__ gamma, const float *__restrict__ beta,
                                                       float4 *__restrict__ y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const int c = (int)(i % tx) * 4;
    const float4 v = __ldg(x + i);
    const float4 mu = __ldg(reinterpret_cast<const float4 *>(mean + c));
    const float4 is = __ldg(reinterpret_cast<const float4 *>(invstd + c));
    const float4 g = __ldg(reinterpret_cast<const float4 *>(gamma + c));
    const float4 b = __ldg(reinterpret_cast<const float4 *>(beta + c));
    float4 o;
    o.x = (v.x - mu.x) * is.x * g.x + b.x;
    o.y = (v.y - mu.y) * is.y * g.y + b.y;
    o.z = (v.z - mu.z) * is.z * g.z + b.z;
    o.w = (v.w - mu.w) * is.w * g.w + b.w;
    if (RELU) {
        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    }
    y[i] = o;
}

// dsum[c] += sum dz ; dsum[C + c] += sum dz * xhat     (dz = dy masked by the recomputed ReLU)
template <bool RELU>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float4 *__restrict__ dy, const float4 *__restrict__ x,
                                                            int64_t n, int tx, int ty, const float *__restrict__ mean,
                                                            const float *__restrict__ invstd, const float *__restrict__ gamma,
                                                            const float *__restrict__ beta, double *__restrict__ dsum) {
    extern __shared__ float4 s_red[];
    const int cx = threadIdx.x % tx, ry = threadIdx.x / tx;
    const int64_t r0 = (int64_t)blockIdx.x * BN_ROWS_PER_BLOCK;
    const int64_t r1 = min(n, r0 + BN_ROWS_PER_BLOCK);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ry < ty) {
        const int c = cx * 4;
        const float4 mu = __ldg(reinterpret_cast<const float4 *>(mean + c));
        const float4 is = __ldg(reinterpret_cast<const float4 *>(invstd + c));
        const float4 g = __ldg(reinterpret_cast<const float4 *>(gamma + c));
        const float4 b = __ldg(reinterpret_cast<const float4 *>(beta + c));
        for (int64_t r = r0 + ry; r < r1; r += ty) {
            const float4 v = __ldg(x + r * tx + cx);
            float4 d = __ldg(dy + r * tx + cx);
            const float4 h = make_float4((v.x - mu.x) * is.x, (v.y - mu.y) * is.y, (v.z - mu.z) * is.z, (v.w - mu.w) * is.w);
            if (RELU) {
                if (h.x * g.x + b.x <= 0.f) d.x = 0.f;
                if (h.y * g.y + b.y <= 0.f) d.y = 0.f;
                if (h.z * g.z + b.z <= 0.f) d.z = 0.f;
                if (h.w * g.w + b.w <= 0.f) d.w = 0.f;
            }
            s.x += d.x; s.y += d.y; s.z += d.z; s.w += d.w;
            q.x += d.x * h.x; q.y += d.y * h.y; q.z += d.z * h.z; q.w += d.w * h.w;
        }
        s_red[ry * tx + cx] = s;
        s_red[(ty + ry) * tx + cx] = q;
    }
    __syncthreads();
    if (ry == 0) {
        for (int j = 1; j < ty; j++) {
            const float4 a = s_red[j * tx + cx], b2 = s_red[(ty + j) * tx + cx];
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
            q.x += b2.x; q.y += b2.y; q.z += b2.z; q.w += b2.w;
        }
        const int C = tx * 4, c = cx * 4;
        atomicAdd(dsum + c + 0, (double)s.x); atomicAdd(dsum + c + 1, (double)s.y);
        atomicAdd(dsum + c + 2, (double)s.z); atomicAdd(dsum + c + 3, (double)s.w);
        atomicAdd(dsum + C + c + 0, (double)q.x); atomicAdd(dsum + C + c + 1, (double)q.y);
        atomicAdd(dsum + C + c + 2, (double)q.z); atomicAdd(dsum + C + c + 3, (double)q.w);
    }
}

// dx = gamma * invstd * (dz - mean(dz) - xhat * mean(dz * xhat)), means over the GLOBAL count
template <bool RELU>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float4 *__restrict__ dy, const float4 *__restrict__ x,
                                                           int64_t n4, int tx, const float *__restrict__ mean,
                                                           const float *__restrict__ invstd, const float *__restrict__ gamma,
                                                           const float *__restrict__ beta, const double *__restrict__ dsum,
                                                           const double *__restrict__ count, float4 *__restrict__ dx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const int C = tx * 4, c = (int)(i % tx) * 4;
    const double inv_count = *count > 0.0 ? 1.0 / *count : 0.0;
    const float4 v = __ldg(x + i);
    float4 d = __ldg(dy + i);
    const float4 mu = __ldg(reinterpret_cast<const float4 *>(mean + c));
    const float4 is = __ldg(reinterpret_cast<const float4 *>(invstd + c));
    const float4 g = __ldg(reinterpret_cast<const float4 *>(gamma + c));
    const float4 h = make_float4((v.x - mu.x) * is.x, (v.y - mu.y) * is.y, (v.z - mu.z) * is.z, (v.w - mu.w) * is.w);
    if (RELU) {
        const float4 b = __ldg(reinterpret_cast<const float4 *>(beta + c));
        if (h.x * g.x + b.x <= 0.f) d.x = 0.f;
        if (h.y * g.y + b.y <= 0.f) d.y = 0.f;
        if (h.z * g.z + b.z <= 0.f) d.z = 0.f;
        if (h.w * g.w + b.w <= 0.f) d.w = 0.f;
    }
    const float a0 = (float)(dsum[c + 0] * inv_count), a1 = (float)(dsum[c + 1] * inv_count);
    const float a2 = (float)(dsum[c + 2] * inv_count), a3 = (float)(dsum[c + 3] * inv_count);
    const float b0 = (float)(dsum[C + c + 0] * inv_count), b1 = (float)(dsum[C + c + 1] * inv_count);
    const float b2 = (float)(dsum[C + c + 2] * inv_count), b3 = (float)(dsum[C + c + 3] * inv_count);
    float4 o;
    o.x = g.x * is.x * (d.x - a0 - h.x * b0);
    o.y = g.y * is.y * (d.y - a1 - h.y * b1);
    o.z = g.z * is.z * (d.z - a2 - h.z * b2);
    o.w = g.w * is.w * (d.w - a3 - h.w * b3);
    dx[i] = o;
}

}  // namespace

extern "C" int u2_bn_supported(int32_t C) { return C > 0 && C % 4 == 0 && C <= 1024; }

// sums: fp64 [2C + 1]; the call zeroes it, accumulates the channel sums and stores n in sums[2C]
extern "C" int u2_bn_stats(const float *x, int64_t n, int32_t C, double *sums, u2_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    U2_CHECK_ARG(u2_bn_supported(C), "u2_bn_stats: C=%d (needs C %% 4 == 0, C <= 1024)", C);
    U2_CHECK_ARG(((uintptr_t)x & 15) == 0, "u2_bn_stats: x must be 16-byte aligned");
    U2_CUDA_OK(cudaMemsetAsync(sums, 0, (size_t)(2 * C + 1) * sizeof(double), st));
    if (n == 0) return 0;
    const BnGeom g = bn_geom(C);
    const int threads = g.tx * g.ty;
    bn_stats_kernel<<<(unsigned)u2_ceil_div(n, BN_ROWS_PER_BLOCK), threads, 2 * threads * sizeof(float4), st>>>(
        (const float4 *)x, n, g.tx, g.ty, sums);
    U2_LAUNCH_OK();
    return 0;
}

extern "C" int u2_bn_apply(const float *x, int64_t n, int32_t C, const double *sums, float eps, float momentum,
                           const float *gamma, const float *beta, int32_t relu, float *y, float *save_mean,
                           float *save_invstd, float *running_mean, float *running_var, u2_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    U2_CHECK_ARG(u2_bn_supported(C), "u2_bn_apply: C=%d", C);
    U2_CHECK_ARG((((uintptr_t)x | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)save_mean |
                   (uintptr_t)save_invstd) & 15) == 0, "u2_bn_apply: pointers must be 16-byte aligned");
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, C, eps, momentum, save_mean, save_invstd, running_mean,
                                                         running_var);
    U2_LAUNCH_OK();
    if (n == 0) return 0;
    const int64_t n4 = n * (C / 4);
    const unsigned grid = (unsigned)u2_ceil_div(n4, 256);
    if (relu)
        bn_apply_kernel<true><<<grid, 256, 0, st>>>((const float4 *)x, n4, C / 4, save_mean, save_invstd, gamma, beta, (float4 *)y);
    else
        bn_apply_kernel<false><<<grid, 256, 0, st>>>((const float4 *)x, n4, C / 4, save_mean, save_invstd, gamma, beta, (float4 *)y);
    U2_LAUNCH_OK();
    return 0;
}

// dsum: fp64 [2C], zeroed by the call: dsum[c] = sum dz (= grad beta), dsum[C+c] = sum dz*xhat (= grad gamma)
extern "C" int u2_bn_bwd_reduce(const float *dy, const float *x, int64_t n, int32_t C, const float *mean,
                                const float *invstd, const float *gamma, const float *beta, int32_t relu, double *dsum,
                                u2_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    U2_CHECK_ARG(u2_bn_supported(C), "u2_bn_bwd_reduce: C=%d", C);
    U2_CHECK_ARG((((uintptr_t)x | (uintptr_t)dy) & 15) == 0, "u2_bn_bwd_reduce: pointers must be 16-byte aligned");
    U2_CUDA_OK(cudaMemsetAsync(dsum, 0, (size_t)(2 * C) * sizeof(double), st));
    if (n == 0) return 0;
    const BnGeom g = bn_geom(C);
    const int threads = g.tx * g.ty;
    const unsigned grid = (unsigned)u2_ceil_div(n, BN_ROWS_PER_BLOCK);
    const size_t smem = 2 * threads * sizeof(float4);
    if (relu)
        bn_bwd_reduce_kernel<true><<<grid, threads, smem, st>>>((const float4 *)dy, (const float4 *)x, n, g.tx, g.ty, mean,
                                                               invstd, gamma, beta, dsum);
    else
        bn_bwd_reduce_kernel<false><<<grid, threads, smem, st>>>((const float4 *)dy, (const float4 *)x, n, g.tx, g.ty, mean,
                                                                invstd, gamma, beta, dsum);
    U2_LAUNCH_OK();
    return 0;
}

extern "C" int u2_bn_bwd_apply(const float *dy, const float *x, int64_t n, int32_t C, const float *mean, const float *invstd,
                               const float *gamma, const float *beta, const double *dsum, const double *count_dev,
                               int32_t relu, float *dx, u2_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    U2_CHECK_ARG(u2_bn_supported(C), "u2_bn_bwd_apply: C=%d", C);
    if (n == 0) return 0;
    const int64_t n4 = n * (C / 4);
    const unsigned grid = (unsigned)u2_ceil_div(n4, 256);
    if (relu)
        bn_bwd_apply_kernel<true><<<grid, 256, 0, st>>>((const float4 *)dy, (const float4 *)x, n4, C / 4, mean, invstd, gamma,
                                                        beta, dsum, count_dev, (float4 *)dx);
    else
        bn_bwd_apply_kernel<false><<<grid, 256, 0, st>>>((const float4 *)dy, (const float4 *)x, n4, C / 4, mean, invstd, gamma,
                                                         beta, dsum, count_dev, (float4 *)dx);
    U2_LAUNCH_OK();
    return 0;
}
