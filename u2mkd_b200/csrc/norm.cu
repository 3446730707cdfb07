// BatchNorm (+ optional fused ReLU) over voxel / point feature matrices fp32 [n, C], training mode.
// SURVEY.md §8(f)-2 (epilogue fusion on the operator surface) and §8(a16)/(a5): the reference applies
// torch BatchNorm1d / SyncBatchNorm + ReLU to SparseTensor.F after every conv
// (core/models/build_blocks.py:21-84, core/models/utils.py:138-141).  Pure HBM-bound passes:
//   forward : stats (1 read)  +  normalise[+ReLU] (1 read, 1 write)           vs 5 passes unfused
//   backward: reduce (2 reads) +  dx (2 reads, 1 write), ReLU mask recomputed  vs 8 passes unfused
// Channel sums are accumulated in fp32 per thread (<= 128 rows) and combined in fp64, so that
// (a) var = E[x^2] - E[x]^2 is safe and (b) the [2C+1] fp64 buffer can be all-reduced across ranks
// as is (SyncBatchNorm: one small collective per pass instead of all_gather + several kernels).
#include <stdlib.h>

#include "u2_common.cuh"

namespace {

// Thread geometry: tx = C / 4 threads across a row (one float4 each), ty = 256 / tx rows side by side; a "row group"
// is ty consecutive rows = one fully coalesced sweep of the block.  CTA b handles row groups b, b + grid, b + 2 grid ..
// (grid-stride, so any n balances over the resident CTAs and there is no tail wave), U groups per iteration: U (or 2U)
// independent 16-byte loads per thread in flight before the first use.
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

inline int bn_env(const char *name, int dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

// grid for n rows: at most `per_sm` CTAs per SM, at least one row group each
inline unsigned bn_grid(int64_t n, const BnGeom &g, int per_sm) {
    const int64_t groups = u2_ceil_div(n, g.ty);
    const int64_t cap = (int64_t)per_sm * U2_NUM_SMS;
    return (unsigned)(groups < cap ? groups : cap);
}

// compiler barrier between the U independent loads of an iteration and their first use: without it the loads are
// interleaved with the math (2-3 in flight instead of U; checked in the SASS)
#define BN_LOADS_FIRST() asm volatile("" ::: "memory")

__device__ __forceinline__ void acc4(float4 &s, const float4 v) { s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
__device__ __forceinline__ void acc4sq(float4 &q, const float4 v) { q.x += v.x * v.x; q.y += v.y * v.y; q.z += v.z * v.z; q.w += v.w * v.w; }

// Block-level combine, then ONE fp32 row of partial sums per CTA: part[blockIdx.x][0][c] = s, [1][c] = q.
// (fp64 atomics from every CTA onto the same 2C addresses serialise in L2: measured ~50 ns per CTA, i.e. the
// reductions ran at 10-40 % of HBM speed and got slower with more CTAs. bn_tiles_reduce_kernel folds the rows.)
__device__ __forceinline__ void block_reduce_to_partials(float4 s, float4 q, int tx, int ty, int cx, int ry, float4 *s_red,
                                                         float *part) {
    s_red[ry * tx + cx] = s;
    s_red[(ty + ry) * tx + cx] = q;
    __syncthreads();
    if (ry == 0) {
        for (int j = 1; j < ty; j++) {
            acc4(s, s_red[j * tx + cx]);
            acc4(q, s_red[(ty + j) * tx + cx]);
        }
        float4 *row = reinterpret_cast<float4 *>(part + (size_t)blockIdx.x * 2 * tx * 4);
        row[cx] = s;
        row[tx + cx] = q;
    }
}

// sums[c] += sum_rows x[r][c];  sums[C + c] += sum_rows x[r][c]^2
template <int U>
__global__ void __launch_bounds__(256) bn_stats_kernel(const float4 *__restrict__ x, int64_t n, int tx, int ty,
                                                       float *__restrict__ part) {
    extern __shared__ float4 s_red[];  // [2][ty][tx]
    const int cx = threadIdx.x % tx, ry = threadIdx.x / tx;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t step = (int64_t)gridDim.x * ty;
    int64_t r = (int64_t)blockIdx.x * ty + ry;
    const float4 *px = x + r * tx + cx;
    const int64_t pstep = step * tx;
    for (; r + (U - 1) * step < n; r += U * step, px += U * pstep) {  // all U rows in range: U loads, then the math
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = __ldg(px + u * pstep);
        BN_LOADS_FIRST();
#pragma unroll
        for (int u = 0; u < U; u++) { acc4(s, v[u]); acc4sq(q, v[u]); }
    }
    for (; r < n; r += step, px += pstep) {
        const float4 v = __ldg(px);
        acc4(s, v); acc4sq(q, v);
    }
    block_reduce_to_partials(s, q, tx, ty, cx, ry, s_red, part);
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

__device__ __forceinline__ float4 ld4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// round-to-nearest-even, same packing as cast_bf16_kernel (conv_tc.cu): low half = first element
__device__ __forceinline__ uint2 bf16x4(const float4 v) {
    uint2 o;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.x) : "f"(v.y), "f"(v.x));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.y) : "f"(v.w), "f"(v.z));
    return o;
}

// y = x * scale + shift (+ReLU); scale = invstd * gamma, shift = beta - mean * scale: computed once per thread
template <bool RELU, int U>
__global__ void __launch_bounds__(256) bn_apply_kernel(const float4 *__restrict__ x, int64_t n, int tx, int ty,
                                                       const double *__restrict__ sums, float eps, float momentum,
                                                       float *__restrict__ mean, float *__restrict__ invstd,
                                                       float *__restrict__ running_mean, float *__restrict__ running_var,
                                                       const float *__restrict__ gamma, const float *__restrict__ beta,
                                                       const float4 *__restrict__ res, float4 *__restrict__ y,
                                                       uint2 *__restrict__ yb) {
    const int cx = threadIdx.x % tx, ry = threadIdx.x / tx;
    const int c = cx * 4, C = tx * 4;
    float4 sc, sh;
    {
        // mean / invstd of this thread's four channels straight from the (possibly all-reduced) fp64 sums — the arithmetic
        // of bn_finalize_kernel, so no separate finalize launch; the first row of threads of CTA 0 also saves them for the
        // backward and updates the running statistics
        const double cnt = sums[2 * C] > 0.0 ? sums[2 * C] : 1.0;
        float mu[4], is[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const double m = sums[c + j] / cnt;
            double var = sums[C + c + j] / cnt - m * m;
            if (var < 0.0) var = 0.0;
            mu[j] = (float)m;
            is[j] = (float)(1.0 / sqrt(var + (double)eps));
            if (blockIdx.x == 0 && ry == 0) {
                mean[c + j] = mu[j];
                invstd[c + j] = is[j];
                if (running_mean) {
                    const double unbiased = cnt > 1.0 ? var * cnt / (cnt - 1.0) : var;
                    running_mean[c + j] = (1.f - momentum) * running_mean[c + j] + momentum * (float)m;
                    running_var[c + j] = (1.f - momentum) * running_var[c + j] + momentum * (float)unbiased;
                }
            }
        }
        const float4 g = ld4(gamma + c), b = ld4(beta + c);
        sc = make_float4(is[0] * g.x, is[1] * g.y, is[2] * g.z, is[3] * g.w);
        sh = make_float4(b.x - mu[0] * sc.x, b.y - mu[1] * sc.y, b.z - mu[2] * sc.z, b.w - mu[3] * sc.w);
    }
    const int64_t step = (int64_t)gridDim.x * ty;
    auto put = [&](int64_t i, const float4 v) {
        float4 o = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
        if (res) {  // residual branch of a ResidualBlock: relu(bn(conv(x)) + shortcut(x)) in one pass
            const float4 rr = __ldg(res + i);
            o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
        }
        if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        y[i] = o;
        if (yb) yb[i] = bf16x4(o);  // the next conv's bf16 operand, written while the row is in registers
    };
    int64_t r = (int64_t)blockIdx.x * ty + ry;
    int64_t i = r * tx + cx;
    const int64_t pstep = step * tx;
    for (; r + (U - 1) * step < n; r += U * step, i += U * pstep) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = __ldg(x + i + u * pstep);
        BN_LOADS_FIRST();
#pragma unroll
        for (int u = 0; u < U; u++) put(i + u * pstep, v[u]);
    }
    for (; r < n; r += step, i += pstep) put(i, __ldg(x + i));
}

// dsum[c] += sum dz ; dsum[C + c] += sum dz * xhat     (dz = dy masked by the recomputed ReLU)
template <bool RELU, int U>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float4 *__restrict__ dy, const float4 *__restrict__ x,
                                                            int64_t n, int tx, int ty, const float *__restrict__ mean,
                                                            const float *__restrict__ invstd, const float *__restrict__ gamma,
                                                            const float *__restrict__ beta, const float4 *__restrict__ zout,
                                                            float *__restrict__ part) {
    extern __shared__ float4 s_red[];
    const int cx = threadIdx.x % tx, ry = threadIdx.x / tx;
    const int c = cx * 4;
    const float4 mu = ld4(mean + c), is = ld4(invstd + c), g = ld4(gamma + c), b = ld4(beta + c);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = make_float4(0.f, 0.f, 0.f, 0.f);
    auto f = [&](float4 v, float4 d, int64_t i) {
        const float4 h = make_float4((v.x - mu.x) * is.x, (v.y - mu.y) * is.y, (v.z - mu.z) * is.z, (v.w - mu.w) * is.w);
        if (RELU) {
            if (zout) {  // a residual was added before the ReLU: the mask is the sign of the saved output
                const float4 z = __ldg(zout + i);
                if (z.x <= 0.f) d.x = 0.f;
                if (z.y <= 0.f) d.y = 0.f;
                if (z.z <= 0.f) d.z = 0.f;
                if (z.w <= 0.f) d.w = 0.f;
            } else {
                if (h.x * g.x + b.x <= 0.f) d.x = 0.f;
                if (h.y * g.y + b.y <= 0.f) d.y = 0.f;
                if (h.z * g.z + b.z <= 0.f) d.z = 0.f;
                if (h.w * g.w + b.w <= 0.f) d.w = 0.f;
            }
        }
        acc4(s, d);
        q.x += d.x * h.x; q.y += d.y * h.y; q.z += d.z * h.z; q.w += d.w * h.w;
    };
    const int64_t step = (int64_t)gridDim.x * ty;
    int64_t r = (int64_t)blockIdx.x * ty + ry;
    int64_t i = r * tx + cx;
    const int64_t pstep = step * tx;
    for (; r + (U - 1) * step < n; r += U * step, i += U * pstep) {
        float4 v[U], d[U];
#pragma unroll
        for (int u = 0; u < U; u++) { v[u] = __ldg(x + i + u * pstep); d[u] = __ldg(dy + i + u * pstep); }
        BN_LOADS_FIRST();
#pragma unroll
        for (int u = 0; u < U; u++) f(v[u], d[u], i + u * pstep);
    }
    for (; r < n; r += step, i += pstep) f(__ldg(x + i), __ldg(dy + i), i);
    block_reduce_to_partials(s, q, tx, ty, cx, ry, s_red, part);
}

// dx = gamma * invstd * (dz - mean(dz) - xhat * mean(dz * xhat)), means over the GLOBAL count
template <bool RELU, int U>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float4 *__restrict__ dy, const float4 *__restrict__ x,
                                                           int64_t n, int tx, int ty, const float *__restrict__ mean,
                                                           const float *__restrict__ invstd, const float *__restrict__ gamma,
                                                           const float *__restrict__ beta, const double *__restrict__ dsum,
                                                           const double *__restrict__ count, const float4 *__restrict__ zout,
                                                           float4 *__restrict__ dres, float4 *__restrict__ dx,
                                                           uint2 *__restrict__ dxb) {
    const int cx = threadIdx.x % tx, ry = threadIdx.x / tx;
    const int C = tx * 4, c = cx * 4;
    const double inv_count = *count > 0.0 ? 1.0 / *count : 0.0;
    const float4 mu = ld4(mean + c), is = ld4(invstd + c), g = ld4(gamma + c), b = ld4(beta + c);
    const float4 a = make_float4((float)(dsum[c] * inv_count), (float)(dsum[c + 1] * inv_count),
                                 (float)(dsum[c + 2] * inv_count), (float)(dsum[c + 3] * inv_count));
    const float4 bb = make_float4((float)(dsum[C + c] * inv_count), (float)(dsum[C + c + 1] * inv_count),
                                  (float)(dsum[C + c + 2] * inv_count), (float)(dsum[C + c + 3] * inv_count));
    const float4 gi = make_float4(g.x * is.x, g.y * is.y, g.z * is.z, g.w * is.w);
    auto f = [&](float4 v, float4 d, int64_t i) {
        const float4 h = make_float4((v.x - mu.x) * is.x, (v.y - mu.y) * is.y, (v.z - mu.z) * is.z, (v.w - mu.w) * is.w);
        if (RELU) {
            if (zout) {
                const float4 z = __ldg(zout + i);
                if (z.x <= 0.f) d.x = 0.f;
                if (z.y <= 0.f) d.y = 0.f;
                if (z.z <= 0.f) d.z = 0.f;
                if (z.w <= 0.f) d.w = 0.f;
            } else {
                if (h.x * g.x + b.x <= 0.f) d.x = 0.f;
                if (h.y * g.y + b.y <= 0.f) d.y = 0.f;
                if (h.z * g.z + b.z <= 0.f) d.z = 0.f;
                if (h.w * g.w + b.w <= 0.f) d.w = 0.f;
            }
        }
        if (dres) dres[i] = d;  // gradient of the residual input = the masked output gradient
        return make_float4(gi.x * (d.x - a.x - h.x * bb.x), gi.y * (d.y - a.y - h.y * bb.y), gi.z * (d.z - a.z - h.z * bb.z),
                           gi.w * (d.w - a.w - h.w * bb.w));
    };
    const int64_t step = (int64_t)gridDim.x * ty;
    auto put = [&](int64_t j, const float4 o) {
        if (dx) dx[j] = o;
        if (dxb) dxb[j] = bf16x4(o);  // the conv backward's bf16 operand (fused conv+BN: the only output)
    };
    int64_t r = (int64_t)blockIdx.x * ty + ry;
    int64_t i = r * tx + cx;
    const int64_t pstep = step * tx;
    for (; r + (U - 1) * step < n; r += U * step, i += U * pstep) {
        float4 v[U], d[U];
#pragma unroll
        for (int u = 0; u < U; u++) { v[u] = __ldg(x + i + u * pstep); d[u] = __ldg(dy + i + u * pstep); }
        BN_LOADS_FIRST();
#pragma unroll
        for (int u = 0; u < U; u++) put(i + u * pstep, f(v[u], d[u], i + u * pstep));
    }
    for (; r < n; r += step, i += pstep) put(i, f(__ldg(x + i), __ldg(dy + i), i));
}

// sums[c] += sum_p part[p][0][c], sums[C + c] += sum_p part[p][1][c]: the per-warp column sums the conv
// epilogue left behind (conv_tc.cu), so that BatchNorm needs no statistics pass over the conv output
// (count >= 0: the local row count rides along in sums[2C])
__global__ void __launch_bounds__(256) bn_tiles_reduce_kernel(const float *__restrict__ part, int64_t P, int C,
                                                              double *__restrict__ sums, double count) {
    __shared__ double s_red[2][8][32];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    if (count >= 0.0 && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) sums[2 * C] = count;
    const int col = blockIdx.x * 32 + cx;
    double s = 0.0, q = 0.0;
    if (col < C)
        for (int64_t p = (int64_t)blockIdx.y * 8 + ry; p < P; p += (int64_t)gridDim.y * 8) {
            s += (double)__ldg(part + (2 * p) * C + col);
            q += (double)__ldg(part + (2 * p + 1) * C + col);
        }
    s_red[0][ry][cx] = s;
    s_red[1][ry][cx] = q;
    __syncthreads();
    if (ry == 0 && col < C) {
        for (int j = 1; j < 8; j++) { s += s_red[0][j][cx]; q += s_red[1][j][cx]; }
        atomicAdd(sums + col, s);
        atomicAdd(sums + C + col, q);
    }
}

}  // namespace

#define BN_DISPATCH_U(u, call)                        \
    do {                                              \
        if ((u) >= 8) { constexpr int UU = 8; call; } \
        else if ((u) >= 4) { constexpr int UU = 4; call; } \
        else if ((u) >= 2) { constexpr int UU = 2; call; } \
        else { constexpr int UU = 1; call; }          \
    } while (0)

extern "C" int u2_bn_supported(int32_t C) { return C > 0 && C % 4 == 0 && C <= 1024; }

constexpr int BN_MAX_CAP = 8;  // most CTAs per SM of the two reduction kernels (sizes their partial-sum scratch)

static int bn_clamp_cap(int cap) { return cap < 1 ? 1 : (cap > BN_MAX_CAP ? BN_MAX_CAP : cap); }

extern "C" size_t u2_bn_scratch_bytes(int32_t C) { return (size_t)BN_MAX_CAP * U2_NUM_SMS * 2 * (size_t)C * sizeof(float); }

// partial rows [P][2][C] fp32 -> sums fp64 [2C] (+ count in sums[2C] if count >= 0); sums zeroed here
static int bn_fold_partials(const float *part, int64_t P, int C, double *sums, double count, cudaStream_t st) {
    U2_CUDA_OK(cudaMemsetAsync(sums, 0, (size_t)(2 * C + (count >= 0.0 ? 1 : 0)) * sizeof(double), st));
    if (P == 0) return 0;
    int64_t chunks = u2_ceil_div(P, 64);  // ~64 partial rows per CTA, <= 64 fp64 atomics per address
    if (chunks > 64) chunks = 64;
    bn_tiles_reduce_kernel<<<dim3((unsigned)((C + 31) / 32), (unsigned)chunks), 256, 0, st>>>(part, P, C, sums, count);
    U2_LAUNCH_OK();
    return 0;
}

// sums: fp64 [2C + 1] = channel sums, sums of squares and n (in sums[2C]); scratch from u2_bn_scratch_bytes(C)
extern "C" int u2_bn_stats(const float *x, int64_t n, int32_t C, double *sums, void *scratch, size_t scratch_bytes,
                           u2_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    U2_CHECK_ARG(u2_bn_supported(C), "u2_bn_stats: C=%d (needs C %% 4 == 0, C <= 1024)", C);
    U2_CHECK_ARG(((uintptr_t)x & 15) == 0, "u2_bn_stats: x must be 16-byte aligned");
    U2_CHECK_ARG(n == 0 || (scratch && scratch_bytes >= u2_bn_scratch_bytes(C) && ((uintptr_t)scratch & 15) == 0),
                 "u2_bn_stats: scratch too small or misaligned");
    unsigned grid = 0;
    if (n > 0) {
        const BnGeom g = bn_geom(C);
        const int threads = g.tx * g.ty;
        static const int cap = bn_clamp_cap(bn_env("U2_BN_CAP_STATS", 6)), unroll = bn_env("U2_BN_U_STATS", 4);
        grid = bn_grid(n, g, cap);
        const size_t smem = 2 * threads * sizeof(float4);
        BN_DISPATCH_U(unroll, (bn_stats_kernel<UU><<<grid, threads, smem, st>>>((const float4 *)x, n, g.tx, g.ty, (float *)scratch)));
        U2_LAUNCH_OK();
    }
    return bn_fold_partials((const float *)scratch, grid, C, sums, (double)n, st);
}

extern "C" int u2_bn_stats_from_tiles(const float *tile_stats, int64_t n_parts, int32_t C, int64_t n, double *sums,
                                      u2_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    U2_CHECK_ARG(tile_stats && sums && C > 0 && n_parts > 0 && n >= 0, "u2_bn_stats_from_tiles: bad arguments");
    return bn_fold_partials(tile_stats, n_parts, C, sums, (double)n, st);
}

extern "C" int u2_bn_apply_dual(const float *x, int64_t n, int32_t C, const double *sums, float eps, float momentum,
                                const float *gamma, const float *beta, int32_t relu, const float *residual, float *y,
                                void *y_bf16,
                                float *save_mean, float *save_invstd, float *running_mean, float *running_var,
                                u2_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    U2_CHECK_ARG(u2_bn_supported(C), "u2_bn_apply: C=%d", C);
    U2_CHECK_ARG((((uintptr_t)x | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)save_mean |
                   (uintptr_t)save_invstd | (uintptr_t)y_bf16 | (uintptr_t)residual) & 15) == 0,
                 "u2_bn_apply: pointers must be 16-byte aligned");
    if (n == 0) {  // nothing to normalise on this rank: only the statistics bookkeeping
        bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, C, eps, momentum, save_mean, save_invstd, running_mean,
                                                             running_var);
        U2_LAUNCH_OK();
        return 0;
    }
    const BnGeom g = bn_geom(C);
    static const int cap = bn_env("U2_BN_CAP_APPLY", 4), unroll = bn_env("U2_BN_U_APPLY", 2);
    const unsigned grid = bn_grid(n, g, cap);
    if (relu)
        BN_DISPATCH_U(unroll, (bn_apply_kernel<true, UU><<<grid, g.tx * g.ty, 0, st>>>(
            (const float4 *)x, n, g.tx, g.ty, sums, eps, momentum, save_mean, save_invstd, running_mean, running_var, gamma, beta,
            (const float4 *)residual, (float4 *)y, (uint2 *)y_bf16)));
    else
        BN_DISPATCH_U(unroll, (bn_apply_kernel<false, UU><<<grid, g.tx * g.ty, 0, st>>>(
            (const float4 *)x, n, g.tx, g.ty, sums, eps, momentum, save_mean, save_invstd, running_mean, running_var, gamma, beta,
            (const float4 *)residual, (float4 *)y, (uint2 *)y_bf16)));
    U2_LAUNCH_OK();
    return 0;
}

extern "C" int u2_bn_apply(const float *x, int64_t n, int32_t C, const double *sums, float eps, float momentum,
                           const float *gamma, const float *beta, int32_t relu, float *y, float *save_mean,
                           float *save_invstd, float *running_mean, float *running_var, u2_stream_t stream) {
    return u2_bn_apply_dual(x, n, C, sums, eps, momentum, gamma, beta, relu, nullptr, y, nullptr, save_mean, save_invstd,
                            running_mean, running_var, stream);
}

// dsum: fp64 [2C], zeroed by the call: dsum[c] = sum dz (= grad beta), dsum[C+c] = sum dz*xhat (= grad gamma)
extern "C" int u2_bn_bwd_reduce(const float *dy, const float *x, int64_t n, int32_t C, const float *mean,
                                const float *invstd, const float *gamma, const float *beta, int32_t relu,
                                const float *out_mask, double *dsum, void *scratch, size_t scratch_bytes,
                                u2_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    U2_CHECK_ARG(u2_bn_supported(C), "u2_bn_bwd_reduce: C=%d", C);
    U2_CHECK_ARG((((uintptr_t)x | (uintptr_t)dy) & 15) == 0, "u2_bn_bwd_reduce: pointers must be 16-byte aligned");
    U2_CHECK_ARG(n == 0 || (scratch && scratch_bytes >= u2_bn_scratch_bytes(C) && ((uintptr_t)scratch & 15) == 0),
                 "u2_bn_bwd_reduce: scratch too small or misaligned");
    unsigned grid = 0;
    if (n > 0) {
        const BnGeom g = bn_geom(C);
        const int threads = g.tx * g.ty;
        static const int cap = bn_clamp_cap(bn_env("U2_BN_CAP_RED", 4)), unroll = bn_env("U2_BN_U_RED", 8);
        grid = bn_grid(n, g, cap);
        const size_t smem = 2 * threads * sizeof(float4);
        if (relu)
            BN_DISPATCH_U(unroll, (bn_bwd_reduce_kernel<true, UU><<<grid, threads, smem, st>>>(
                (const float4 *)dy, (const float4 *)x, n, g.tx, g.ty, mean, invstd, gamma, beta, (const float4 *)out_mask,
                (float *)scratch)));
        else
            BN_DISPATCH_U(unroll, (bn_bwd_reduce_kernel<false, UU><<<grid, threads, smem, st>>>(
                (const float4 *)dy, (const float4 *)x, n, g.tx, g.ty, mean, invstd, gamma, beta, nullptr, (float *)scratch)));
        U2_LAUNCH_OK();
    }
    return bn_fold_partials((const float *)scratch, grid, C, dsum, -1.0, st);
}

extern "C" int u2_bn_bwd_apply_dual(const float *dy, const float *x, int64_t n, int32_t C, const float *mean,
                                    const float *invstd, const float *gamma, const float *beta, const double *dsum,
                                    const double *count_dev, int32_t relu, const float *out_mask, float *dresidual,
                                    float *dx, void *dx_bf16, u2_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    U2_CHECK_ARG(u2_bn_supported(C), "u2_bn_bwd_apply: C=%d", C);
    U2_CHECK_ARG(dx || dx_bf16, "u2_bn_bwd_apply: no output");
    U2_CHECK_ARG((((uintptr_t)dx | (uintptr_t)dx_bf16 | (uintptr_t)out_mask | (uintptr_t)dresidual) & 15) == 0,
                 "u2_bn_bwd_apply: pointers must be 16-byte aligned");
    if (n == 0) return 0;
    const BnGeom g = bn_geom(C);
    static const int cap = bn_env("U2_BN_CAP_BAPPLY", 4), unroll = bn_env("U2_BN_U_BAPPLY", 2);
    const unsigned grid = bn_grid(n, g, cap);
    if (relu)
        BN_DISPATCH_U(unroll, (bn_bwd_apply_kernel<true, UU><<<grid, g.tx * g.ty, 0, st>>>(
            (const float4 *)dy, (const float4 *)x, n, g.tx, g.ty, mean, invstd, gamma, beta, dsum, count_dev,
            (const float4 *)out_mask, (float4 *)dresidual, (float4 *)dx, (uint2 *)dx_bf16)));
    else
        BN_DISPATCH_U(unroll, (bn_bwd_apply_kernel<false, UU><<<grid, g.tx * g.ty, 0, st>>>(
            (const float4 *)dy, (const float4 *)x, n, g.tx, g.ty, mean, invstd, gamma, beta, dsum, count_dev, nullptr,
            (float4 *)dresidual, (float4 *)dx, (uint2 *)dx_bf16)));
    U2_LAUNCH_OK();
    return 0;
}

extern "C" int u2_bn_bwd_apply(const float *dy, const float *x, int64_t n, int32_t C, const float *mean, const float *invstd,
                               const float *gamma, const float *beta, const double *dsum, const double *count_dev,
                               int32_t relu, float *dx, u2_stream_t stream) {
    return u2_bn_bwd_apply_dual(dy, x, n, C, mean, invstd, gamma, beta, dsum, count_dev, relu, nullptr, nullptr, dx, nullptr,
                                stream);
}
