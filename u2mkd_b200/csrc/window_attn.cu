// Variable-length window attention with contextual relative position encoding: the arithmetic of
// third_party/SparseTransformer (`sptr`) that SphereFormer runs between the down stages of the teacher / student
// (core/models/sphereformer/spherical_transformer.py:165-283, core/models/nuscenes/spvcnn_spformer.py:145-149).
//
// Reference: six SIMT kernels + two torch_scatter passes per attention, with the [M, h] score matrix (M = sum of squared
// window sizes) materialised three times in HBM —
//   attn = q.k + q.(Tq[r0,0]+Tq[r1,1]+Tq[r2,2]) + k.(Tk[..])   src/sptr/rpe/relative_pos_encoding_cuda_kernel.cu:116-254
//   softmax over the keys of each query                          sptr/utils.py:81-95 (segment_csr / gather_csr)
//   out  = sum_j attn_ij (v_j + Tv[r0,0]+Tv[r1,1]+Tv[r2,2])      src/sptr/rpe/relative_pos_encoding_cuda_kernel.cu (step 2)
// Here: ONE forward kernel and ONE backward kernel, flash-style — scores live in registers, the softmax is online, the
// backward recomputes the probabilities from the saved log-sum-exp; nothing of size M is written except what the caller
// already owns (rel_idx).  Points are sorted by window (get_indices_params), so a window is a contiguous row range and
// pair m = sq_off[w] + i * n_w + j  <->  (query start_w + i, key start_w + j), exactly the order precompute_all produces
// (src/sptr/precompute/precompute_cuda_kernel.cu:4-22).
//
// Work split: block = (persistent slot, head), 64 threads; a thread owns one query (forward, backward pass A) or one key
// (backward pass B) of the current window; keys / queries of the window stream through shared memory in chunks of 64
// rows; the head's three tables (L x 3 x D floats each) stay in shared memory for the block's lifetime, and so do the
// table-gradient accumulators of the backward (one flush of global atomics per block instead of one per thread as in the
// reference).  fp32 throughout (the reference arithmetic); this path is bound by HBM / shared-memory traffic, not flops.
#include <math.h>

#include "u2_common.cuh"

namespace {

constexpr int WA_THREADS = 64;
constexpr int WA_CHUNK = 64;
constexpr int WA_MAX_L = 64;

struct WaParams {
    const float *q, *k, *v;      // [N, h, D]
    const int *win_off;          // [n_windows + 1] first row of each window
    const int *sq_off;           // [n_windows + 1] first pair of each window
    const int *rel;              // [M, 3] or nullptr
    const float *tq, *tk, *tv;   // [L, 3, h, D] or nullptr
    float *out, *lse;            // [N, h, D], [N, h]
    // backward
    const float *dout;
    float *dq, *dk, *dv, *dtq, *dtk, *dtv;
    int n_windows, h, L;
};

template <int D>
__device__ __forceinline__ void load_row(float (&r)[D], const float *p) {
#pragma unroll
    for (int d = 0; d < D; d += 4) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(p + d));
        r[d] = t.x; r[d + 1] = t.y; r[d + 2] = t.z; r[d + 3] = t.w;
    }
}

// sum of the three table rows selected by (r0, r1, r2) for one head: T[r_a, a, :]
template <int D>
__device__ __forceinline__ void table_sum(float (&t)[D], const float *s_tab, int r0, int r1, int r2) {
    const float *a = s_tab + (r0 * 3 + 0) * D, *b = s_tab + (r1 * 3 + 1) * D, *c = s_tab + (r2 * 3 + 2) * D;
#pragma unroll
    for (int d = 0; d < D; d++) t[d] = a[d] + b[d] + c[d];
}

template <int D, bool REL>
__global__ void __launch_bounds__(WA_THREADS) window_attn_fwd_kernel(const WaParams p) {
    extern __shared__ __align__(16) float smem_f[];
    float *s_k = smem_f;                      // [CHUNK][D]
    float *s_v = s_k + WA_CHUNK * D;          // [CHUNK][D]
    float *s_tq = s_v + WA_CHUNK * D;         // [L][3][D] each
    float *s_tk = s_tq + (REL ? p.L * 3 * D : 0);
    float *s_tv = s_tk + (REL ? p.L * 3 * D : 0);
    const int tid = threadIdx.x, head = blockIdx.y, C = p.h * D;
    if (REL) {
        for (int e = tid; e < p.L * 3 * D; e += WA_THREADS) {
            const int d = e % D, la = e / D;  // la = l * 3 + a
            const size_t g = ((size_t)la * p.h + head) * D + d;
            s_tq[e] = __ldg(p.tq + g);
            s_tk[e] = __ldg(p.tk + g);
            s_tv[e] = __ldg(p.tv + g);
        }
    }
    __syncthreads();
    for (int w = blockIdx.x; w < p.n_windows; w += gridDim.x) {
        const int start = __ldg(p.win_off + w), nw = __ldg(p.win_off + w + 1) - start;
        const long long sq = __ldg(p.sq_off + w);
        for (int i0 = 0; i0 < nw; i0 += WA_THREADS) {
            const int i = i0 + tid;
            const bool live = i < nw;
            float q[D], o[D];
            float m = -INFINITY, l = 0.f;
#pragma unroll
            for (int d = 0; d < D; d++) o[d] = 0.f;
            if (live) load_row<D>(q, p.q + (size_t)(start + i) * C + head * D);
            for (int j0 = 0; j0 < nw; j0 += WA_CHUNK) {
                const int nj = min(WA_CHUNK, nw - j0);
                __syncthreads();
                if (tid < nj) {
                    float r[D];
                    load_row<D>(r, p.k + (size_t)(start + j0 + tid) * C + head * D);
#pragma unroll
                    for (int d = 0; d < D; d++) s_k[tid * D + d] = r[d];
                    load_row<D>(r, p.v + (size_t)(start + j0 + tid) * C + head * D);
#pragma unroll
                    for (int d = 0; d < D; d++) s_v[tid * D + d] = r[d];
                }
                __syncthreads();
                if (!live) continue;
                const int *rel = REL ? p.rel + (sq + (long long)i * nw + j0) * 3 : nullptr;
                for (int j = 0; j < nj; j++) {
                    const float *kj = s_k + j * D, *vj = s_v + j * D;
                    float s = 0.f;
                    float tv[D];
                    if (REL) {
                        const int r0 = __ldg(rel + 3 * j), r1 = __ldg(rel + 3 * j + 1), r2 = __ldg(rel + 3 * j + 2);
                        float tq[D], tk[D];
                        table_sum<D>(tq, s_tq, r0, r1, r2);
                        table_sum<D>(tk, s_tk, r0, r1, r2);
                        table_sum<D>(tv, s_tv, r0, r1, r2);
#pragma unroll
                        for (int d = 0; d < D; d++) s = fmaf(q[d], kj[d] + tq[d], fmaf(kj[d], tk[d], s));
                    } else {
#pragma unroll
                        for (int d = 0; d < D; d++) s = fmaf(q[d], kj[d], s);
                    }
                    const float mn = fmaxf(m, s);
                    const float sc = __expf(m - mn), pj = __expf(s - mn);
                    l = fmaf(l, sc, pj);
#pragma unroll
                    for (int d = 0; d < D; d++) o[d] = fmaf(o[d], sc, pj * (REL ? vj[d] + tv[d] : vj[d]));
                    m = mn;
                }
            }
            if (live) {
                const float inv = 1.f / l;
                float *op = p.out + (size_t)(start + i) * C + head * D;
#pragma unroll
                for (int d = 0; d < D; d += 4)
                    *reinterpret_cast<float4 *>(op + d) = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
                p.lse[(size_t)(start + i) * p.h + head] = m + __logf(l);
            }
        }
    }
}

// Backward.  Pass A (thread = query i): D_i = dO_i.O_i, dQ_i, dTq (+= ds q_i), dTv (+= p dO_i).
//            Pass B (thread = key j)  : dK_j, dV_j, dTk (+= ds k_j).
// Both recompute s_ij and p_ij = exp(s_ij - lse_i); ds_ij = p_ij (dO_i.(v_j + tv_ij) - D_i).
template <int D, bool REL>
__global__ void __launch_bounds__(WA_THREADS) window_attn_bwd_kernel(const WaParams p) {
    extern __shared__ __align__(16) float smem_f[];
    float *s_a = smem_f;                        // [CHUNK][D]  pass A: K chunk      pass B: Q chunk
    float *s_b = s_a + WA_CHUNK * D;            // [CHUNK][D]  pass A: V chunk      pass B: dO chunk
    float *s_c = s_b + WA_CHUNK * D;            // [CHUNK][2]  pass B: lse_i, D_i
    float *s_tq = s_c + WA_CHUNK * 2;
    const int TS = REL ? p.L * 3 * D : 0;
    float *s_tk = s_tq + TS, *s_tv = s_tk + TS;
    float *s_gq = s_tv + TS, *s_gk = s_gq + TS, *s_gv = s_gk + TS;  // table-gradient accumulators
    const int tid = threadIdx.x, head = blockIdx.y, C = p.h * D;
    if (REL) {
        for (int e = tid; e < TS; e += WA_THREADS) {
            const int d = e % D, la = e / D;
            const size_t g = ((size_t)la * p.h + head) * D + d;
            s_tq[e] = __ldg(p.tq + g);
            s_tk[e] = __ldg(p.tk + g);
            s_tv[e] = __ldg(p.tv + g);
            s_gq[e] = 0.f; s_gk[e] = 0.f; s_gv[e] = 0.f;
        }
    }
    __syncthreads();
    for (int w = blockIdx.x; w < p.n_windows; w += gridDim.x) {
        const int start = __ldg(p.win_off + w), nw = __ldg(p.win_off + w + 1) - start;
        const long long sq = __ldg(p.sq_off + w);
        // ---------------- pass A: rows (queries) ----------------
        for (int i0 = 0; i0 < nw; i0 += WA_THREADS) {
            const int i = i0 + tid;
            const bool live = i < nw;
            float q[D], go[D], dq[D];
            float lse = 0.f, Di = 0.f;
#pragma unroll
            for (int d = 0; d < D; d++) dq[d] = 0.f;
            if (live) {
                const size_t row = (size_t)(start + i) * C + head * D;
                load_row<D>(q, p.q + row);
                load_row<D>(go, p.dout + row);
                float o[D];
                load_row<D>(o, p.out + row);
#pragma unroll
                for (int d = 0; d < D; d++) Di = fmaf(go[d], o[d], Di);
                lse = __ldg(p.lse + (size_t)(start + i) * p.h + head);
            }
            for (int j0 = 0; j0 < nw; j0 += WA_CHUNK) {
                const int nj = min(WA_CHUNK, nw - j0);
                __syncthreads();
                if (tid < nj) {
                    float r[D];
                    load_row<D>(r, p.k + (size_t)(start + j0 + tid) * C + head * D);
#pragma unroll
                    for (int d = 0; d < D; d++) s_a[tid * D + d] = r[d];
                    load_row<D>(r, p.v + (size_t)(start + j0 + tid) * C + head * D);
#pragma unroll
                    for (int d = 0; d < D; d++) s_b[tid * D + d] = r[d];
                }
                __syncthreads();
                if (!live) continue;
                const int *rel = REL ? p.rel + (sq + (long long)i * nw + j0) * 3 : nullptr;
                for (int j = 0; j < nj; j++) {
                    const float *kj = s_a + j * D, *vj = s_b + j * D;
                    float s = 0.f, dp = 0.f;
                    float tq[D];
                    int r0 = 0, r1 = 0, r2 = 0;
                    if (REL) {
                        r0 = __ldg(rel + 3 * j); r1 = __ldg(rel + 3 * j + 1); r2 = __ldg(rel + 3 * j + 2);
                        float tk[D], tv[D];
                        table_sum<D>(tq, s_tq, r0, r1, r2);
                        table_sum<D>(tk, s_tk, r0, r1, r2);
                        table_sum<D>(tv, s_tv, r0, r1, r2);
#pragma unroll
                        for (int d = 0; d < D; d++) {
                            s = fmaf(q[d], kj[d] + tq[d], fmaf(kj[d], tk[d], s));
                            dp = fmaf(go[d], vj[d] + tv[d], dp);
                        }
                    } else {
#pragma unroll
                        for (int d = 0; d < D; d++) { s = fmaf(q[d], kj[d], s); dp = fmaf(go[d], vj[d], dp); }
                    }
                    const float pj = __expf(s - lse);
                    const float ds = pj * (dp - Di);
#pragma unroll
                    for (int d = 0; d < D; d++) dq[d] = fmaf(ds, REL ? kj[d] + tq[d] : kj[d], dq[d]);
                    if (REL) {
                        float *gq0 = s_gq + (r0 * 3 + 0) * D, *gq1 = s_gq + (r1 * 3 + 1) * D, *gq2 = s_gq + (r2 * 3 + 2) * D;
                        float *gv0 = s_gv + (r0 * 3 + 0) * D, *gv1 = s_gv + (r1 * 3 + 1) * D, *gv2 = s_gv + (r2 * 3 + 2) * D;
#pragma unroll
                        for (int d = 0; d < D; d++) {
                            const float a = ds * q[d], b = pj * go[d];
                            atomicAdd(gq0 + d, a); atomicAdd(gq1 + d, a); atomicAdd(gq2 + d, a);
                            atomicAdd(gv0 + d, b); atomicAdd(gv1 + d, b); atomicAdd(gv2 + d, b);
                        }
                    }
                }
            }
            if (live) {
                float *op = p.dq + (size_t)(start + i) * C + head * D;
#pragma unroll
                for (int d = 0; d < D; d += 4) *reinterpret_cast<float4 *>(op + d) = make_float4(dq[d], dq[d + 1], dq[d + 2], dq[d + 3]);
            }
        }
        // ---------------- pass B: columns (keys) ----------------
        for (int j0 = 0; j0 < nw; j0 += WA_THREADS) {
            const int j = j0 + tid;
            const bool live = j < nw;
            float kk[D], vv[D], dk[D], dv[D];
#pragma unroll
            for (int d = 0; d < D; d++) { dk[d] = 0.f; dv[d] = 0.f; }
            if (live) {
                load_row<D>(kk, p.k + (size_t)(start + j) * C + head * D);
                load_row<D>(vv, p.v + (size_t)(start + j) * C + head * D);
            }
            for (int i0 = 0; i0 < nw; i0 += WA_CHUNK) {
                const int ni = min(WA_CHUNK, nw - i0);
                __syncthreads();
                if (tid < ni) {
                    const size_t row = (size_t)(start + i0 + tid) * C + head * D;
                    float r[D], g[D], o[D];
                    load_row<D>(r, p.q + row);
                    load_row<D>(g, p.dout + row);
                    load_row<D>(o, p.out + row);
                    float Di = 0.f;
#pragma unroll
                    for (int d = 0; d < D; d++) { s_a[tid * D + d] = r[d]; s_b[tid * D + d] = g[d]; Di = fmaf(g[d], o[d], Di); }
                    s_c[tid * 2] = __ldg(p.lse + (size_t)(start + i0 + tid) * p.h + head);
                    s_c[tid * 2 + 1] = Di;
                }
                __syncthreads();
                if (!live) continue;
                for (int i = 0; i < ni; i++) {
                    const float *qi = s_a + i * D, *gi = s_b + i * D;
                    float s = 0.f, dp = 0.f;
                    float tk[D];
                    int r0 = 0, r1 = 0, r2 = 0;
                    if (REL) {
                        const int *rel = p.rel + (sq + (long long)(i0 + i) * nw + j) * 3;
                        r0 = __ldg(rel); r1 = __ldg(rel + 1); r2 = __ldg(rel + 2);
                        float tq[D], tv[D];
                        table_sum<D>(tq, s_tq, r0, r1, r2);
                        table_sum<D>(tk, s_tk, r0, r1, r2);
                        table_sum<D>(tv, s_tv, r0, r1, r2);
#pragma unroll
                        for (int d = 0; d < D; d++) {
                            s = fmaf(qi[d], kk[d] + tq[d], fmaf(kk[d], tk[d], s));
                            dp = fmaf(gi[d], vv[d] + tv[d], dp);
                        }
                    } else {
#pragma unroll
                        for (int d = 0; d < D; d++) { s = fmaf(qi[d], kk[d], s); dp = fmaf(gi[d], vv[d], dp); }
                    }
                    const float pj = __expf(s - s_c[i * 2]);
                    const float ds = pj * (dp - s_c[i * 2 + 1]);
#pragma unroll
                    for (int d = 0; d < D; d++) {
                        dk[d] = fmaf(ds, REL ? qi[d] + tk[d] : qi[d], dk[d]);
                        dv[d] = fmaf(pj, gi[d], dv[d]);
                    }
                    if (REL) {
                        float *g0 = s_gk + (r0 * 3 + 0) * D, *g1 = s_gk + (r1 * 3 + 1) * D, *g2 = s_gk + (r2 * 3 + 2) * D;
#pragma unroll
                        for (int d = 0; d < D; d++) {
                            const float a = ds * kk[d];
                            atomicAdd(g0 + d, a); atomicAdd(g1 + d, a); atomicAdd(g2 + d, a);
                        }
                    }
                }
            }
            if (live) {
                float *ok = p.dk + (size_t)(start + j) * C + head * D, *ov = p.dv + (size_t)(start + j) * C + head * D;
#pragma unroll
                for (int d = 0; d < D; d += 4) {
                    *reinterpret_cast<float4 *>(ok + d) = make_float4(dk[d], dk[d + 1], dk[d + 2], dk[d + 3]);
                    *reinterpret_cast<float4 *>(ov + d) = make_float4(dv[d], dv[d + 1], dv[d + 2], dv[d + 3]);
                }
            }
        }
    }
    if (REL) {
        __syncthreads();
        for (int e = tid; e < TS; e += WA_THREADS) {
            const int d = e % D, la = e / D;
            const size_t g = ((size_t)la * p.h + head) * D + d;
            if (s_gq[e] != 0.f) atomicAdd(p.dtq + g, s_gq[e]);
            if (s_gk[e] != 0.f) atomicAdd(p.dtk + g, s_gk[e]);
            if (s_gv[e] != 0.f) atomicAdd(p.dtv + g, s_gv[e]);
        }
    }
}

// index_0 / index_1 / offsets of every (query, key) pair, the layout precompute_all defines
// (src/sptr/precompute/precompute_cuda_kernel.cu:4-22): one block per window, threads over its n_w^2 pairs.
__global__ void __launch_bounds__(256) window_pairs_kernel(const int *__restrict__ win_off, const int *__restrict__ sq_off, int n_windows,
                                                           int *__restrict__ index0_offsets, int *__restrict__ index1_offsets,
                                                           int *__restrict__ index0, int *__restrict__ index1) {
    for (int w = blockIdx.x; w < n_windows; w += gridDim.x) {
        const int start = win_off[w], nw = win_off[w + 1] - start, sq = sq_off[w];
        for (int t = threadIdx.x; t < nw; t += blockDim.x) {
            index0_offsets[start + t] = sq + nw * t;
            index1_offsets[start + t] = sq + t;
        }
        for (int e = threadIdx.x; e < nw * nw; e += blockDim.x) {
            index0[sq + e] = start + e / nw;
            index1[sq + e] = start + e % nw;
        }
    }
}

int grid_x(int n_windows, int h) {
    int g = (U2_NUM_SMS * 16 + h - 1) / h;  // ~16 resident 64-thread blocks per SM over all heads
    return g < n_windows ? g : n_windows;
}

}  // namespace

extern "C" int u2_window_pairs(const int32_t *win_off, const int32_t *sq_off, int32_t n_windows, int32_t *index0_offsets,
                               int32_t *index1_offsets, int32_t *index0, int32_t *index1, u2_stream_t stream) {
    U2_CHECK_ARG(win_off && sq_off && index0_offsets && index1_offsets, "u2_window_pairs: null pointer");
    if (n_windows <= 0) return 0;
    const int grid = n_windows < U2_NUM_SMS * 8 ? n_windows : U2_NUM_SMS * 8;
    window_pairs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(win_off, sq_off, n_windows, index0_offsets, index1_offsets, index0, index1);
    U2_LAUNCH_OK();
    return 0;
}

extern "C" int u2_window_attn_supported(int32_t head_dim, int32_t L) { return (head_dim == 16 || head_dim == 32) && L >= 0 && L <= WA_MAX_L; }

template <int D, bool REL>
static int wa_launch_fwd(const WaParams &p, cudaStream_t st) {
    const size_t smem = (size_t)(2 * WA_CHUNK * D + (REL ? 3 * p.L * 3 * D : 0)) * sizeof(float);
    U2_CUDA_OK(cudaFuncSetAttribute(window_attn_fwd_kernel<D, REL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    window_attn_fwd_kernel<D, REL><<<dim3(grid_x(p.n_windows, p.h), p.h), WA_THREADS, smem, st>>>(p);
    U2_LAUNCH_OK();
    return 0;
}

template <int D, bool REL>
static int wa_launch_bwd(const WaParams &p, cudaStream_t st) {
    const size_t smem = (size_t)(2 * WA_CHUNK * D + 2 * WA_CHUNK + (REL ? 6 * p.L * 3 * D : 0)) * sizeof(float);
    U2_CUDA_OK(cudaFuncSetAttribute(window_attn_bwd_kernel<D, REL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    window_attn_bwd_kernel<D, REL><<<dim3(grid_x(p.n_windows, p.h), p.h), WA_THREADS, smem, st>>>(p);
    U2_LAUNCH_OK();
    return 0;
}

static int wa_check(const WaParams &p, int D, const char *who) {
    U2_CHECK_ARG(p.q && p.k && p.v && p.win_off && p.sq_off && p.out && p.lse, "%s: null pointer", who);
    U2_CHECK_ARG(u2_window_attn_supported(D, p.L), "%s: head_dim %d / table length %d not supported (head_dim 16 or 32, L <= %d)", who, D, p.L, WA_MAX_L);
    U2_CHECK_ARG((p.rel != nullptr) == (p.tq != nullptr) && (p.rel != nullptr) == (p.tk != nullptr) && (p.rel != nullptr) == (p.tv != nullptr),
                 "%s: rel_idx and the three tables come together", who);
    U2_CHECK_ARG((((uintptr_t)p.q | (uintptr_t)p.k | (uintptr_t)p.v | (uintptr_t)p.out) & 15) == 0, "%s: 16-byte alignment required", who);
    return 0;
}

extern "C" int u2_window_attn_fwd(const float *q, const float *k, const float *v, const int32_t *win_off, const int32_t *sq_off,
                                  int32_t n_windows, int32_t h, int32_t head_dim, const int32_t *rel_idx, const float *table_q,
                                  const float *table_k, const float *table_v, int32_t L, float *out, float *lse, u2_stream_t stream) {
    WaParams p = {};
    p.q = q; p.k = k; p.v = v; p.win_off = win_off; p.sq_off = sq_off; p.rel = rel_idx; p.tq = table_q; p.tk = table_k; p.tv = table_v;
    p.out = out; p.lse = lse; p.n_windows = n_windows; p.h = h; p.L = rel_idx ? L : 0;
    if (wa_check(p, head_dim, "u2_window_attn_fwd")) return 1;
    if (n_windows <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (head_dim == 16) return rel_idx ? wa_launch_fwd<16, true>(p, st) : wa_launch_fwd<16, false>(p, st);
    return rel_idx ? wa_launch_fwd<32, true>(p, st) : wa_launch_fwd<32, false>(p, st);
}

extern "C" int u2_window_attn_bwd(const float *q, const float *k, const float *v, const int32_t *win_off, const int32_t *sq_off,
                                  int32_t n_windows, int32_t h, int32_t head_dim, const int32_t *rel_idx, const float *table_q,
                                  const float *table_k, const float *table_v, int32_t L, const float *out, const float *lse,
                                  const float *dout, float *dq, float *dk, float *dv, float *dtable_q, float *dtable_k,
                                  float *dtable_v, u2_stream_t stream) {
    WaParams p = {};
    p.q = q; p.k = k; p.v = v; p.win_off = win_off; p.sq_off = sq_off; p.rel = rel_idx; p.tq = table_q; p.tk = table_k; p.tv = table_v;
    p.out = const_cast<float *>(out); p.lse = const_cast<float *>(lse); p.dout = dout; p.dq = dq; p.dk = dk; p.dv = dv;
    p.dtq = dtable_q; p.dtk = dtable_k; p.dtv = dtable_v; p.n_windows = n_windows; p.h = h; p.L = rel_idx ? L : 0;
    if (wa_check(p, head_dim, "u2_window_attn_bwd")) return 1;
    U2_CHECK_ARG(dout && dq && dk && dv && (!rel_idx || (dtable_q && dtable_k && dtable_v)), "u2_window_attn_bwd: null gradient pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (rel_idx) {  // the blocks add their table gradients with atomics
        const size_t tb = (size_t)L * 3 * h * head_dim * sizeof(float);
        U2_CUDA_OK(cudaMemsetAsync(dtable_q, 0, tb, st));
        U2_CUDA_OK(cudaMemsetAsync(dtable_k, 0, tb, st));
        U2_CUDA_OK(cudaMemsetAsync(dtable_v, 0, tb, st));
    }
    if (n_windows <= 0) return 0;
    if (head_dim == 16) return rel_idx ? wa_launch_bwd<16, true>(p, st) : wa_launch_bwd<16, false>(p, st);
    return rel_idx ? wa_launch_bwd<32, true>(p, st) : wa_launch_bwd<32, false>(p, st);
}
