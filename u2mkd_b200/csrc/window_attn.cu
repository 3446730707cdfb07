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
// table-gradient accumulators of the backward (each persistent block writes ONE partial table with plain stores, a second
// small kernel folds the partials — the reference adds 150 values per thread with global atomics).  fp32 throughout (the reference arithmetic); this path is bound by HBM / shared-memory traffic, not flops.
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
    float *dtab_part;            // backward: per-block table-gradient partials [3][gridDim.x][L * 3][h][D]
    int part_smem;               // backward: the block keeps its partial in shared memory and writes it once at the end
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
__device__ __forceinline__ void table_sum(float (&t)[D], const float *s_tab, int r0, int r1, int r2, int stride = D) {
    const float *a = s_tab + (r0 * 3 + 0) * stride, *b = s_tab + (r1 * 3 + 1) * stride, *c = s_tab + (r2 * 3 + 2) * stride;
#pragma unroll
    for (int d = 0; d < D; d++) t[d] = a[d] + b[d] + c[d];
}

template <int D>
__device__ __forceinline__ void store_row(float *p, const float (&r)[D]) {
#pragma unroll
    for (int d = 0; d < D; d += 4) *reinterpret_cast<float4 *>(p + d) = make_float4(r[d], r[d + 1], r[d + 2], r[d + 3]);
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

// Backward.  Pass A (rows = queries i): D_i = dO_i.O_i, dQ_i, and the table gradients that are sums over a query's keys:
//                dTq[r, a, :] += (sum_j ds_ij [r_a(i,j) = r]) q_i,   dTv[r, a, :] += (sum_j p_ij [r_a(i,j) = r]) dO_i.
//            Pass B (rows = keys j)   : dK_j, dV_j, dTk[r, a, :] += (sum_i ds_ij [r_a(i,j) = r]) k_j.
// Both recompute s_ij and p_ij = exp(s_ij - lse_i); ds_ij = p_ij (dO_i.(v_j + tv_ij) - D_i).
// Work layout: a block owns a contiguous range of windows holding ~1/gridDim.x of the PAIRS and walks it in groups of up to
// WB_ROWS (64) rows — several small windows packed together, or one larger window in 64-row pieces.  A row is served by a QUAD
// of lanes, each owning D/4 of the channels (dot products finish with two quad shuffles), so a group keeps 256 threads busy
// whatever the window size (the first version ran one window per 64-thread pass: 5 live threads on the cubic windows).
// Table gradients WITHOUT atomics: the bracketed sums are SCALAR histograms over the 3 L (row, axis) buckets, private to a
// row (shared memory, H[bucket][row]; lane a of the quad updates axis a, `+=` is a plain read-modify-write); when the group's
// pairs are done the block folds  part[bucket][d] += sum_row H[bucket][row] * x_row[d]  — a [3L x 64] x [64 x D] product in
// which every thread owns fixed outputs of the block's partial (shared memory when it fits, else its slice of `dtab_part`).
// (Adding 9 x D floats per pair into shared accumulators with atomicAdd is 80 x slower than the table-free kernel: fp32
// shared atomics are CAS spin loops, ATOMS.CAST.SPIN.)
constexpr int WB_ROWS = 64, WB_THREADS = WB_ROWS * 4;

__device__ __forceinline__ float quad_sum(float x, unsigned mask) {
    x += __shfl_xor_sync(mask, x, 1);
    x += __shfl_xor_sync(mask, x, 2);
    return x;
}

// first w in [0, n] with off[w] >= target (off is non-decreasing, off[n] is the total)
__device__ __forceinline__ int first_window_at(const int *off, int n, long long target) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((long long)__ldg(off + mid) < target) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <int D, bool REL>
__global__ void __launch_bounds__(WB_THREADS) window_attn_bwd_kernel(const WaParams p) {
    constexpr int DL = D / 4;                    // channels of a row owned by one lane of its quad
    extern __shared__ __align__(16) float smem_f[];
    float *s_a = smem_f;                        // [ROWS][D]  pass A: K chunk      pass B: Q chunk
    float *s_b = s_a + WB_ROWS * D;             // [ROWS][D]  pass A: V chunk      pass B: dO chunk
    float *s_c = s_b + WB_ROWS * D;             // [ROWS][2]  pass B: lse_i, D_i
    float *s_r0 = s_c + WB_ROWS * 2;            // [ROWS][D]  the group's own rows for the fold: q_i (A) / k_j (B)
    float *s_r1 = s_r0 + WB_ROWS * D;           // [ROWS][D]  dO_i (A)
    float *s_tq = s_r1 + WB_ROWS * D;
    const int TS = REL ? p.L * 3 * D : 0, LA = REL ? p.L * 3 : 0;
    float *s_tk = s_tq + TS, *s_tv = s_tk + TS;
    float *s_h0 = s_tv + TS;                    // [LA][ROWS] histogram of ds   (A: per query, B: per key)
    float *s_h1 = s_h0 + LA * WB_ROWS;          // [LA][ROWS] histogram of p    (A)
    float *s_part = s_h1 + LA * WB_ROWS;        // [3][TS] the block's table-gradient partial when it fits (p.part_smem)
    const int tid = threadIdx.x, head = blockIdx.y, C = p.h * D;
    const int rs = tid >> 2, c = tid & 3, c0 = c * DL;   // row slot, lane of the quad, first owned channel
    const unsigned qmask = 0xFu << ((tid & 31) & ~3);
    const size_t tab = (size_t)p.L * 3 * p.h * D, slice = (size_t)gridDim.x * tab;
    float *part = REL ? p.dtab_part + (size_t)blockIdx.x * tab : nullptr;   // + t * slice for table t
    if (REL) {
        for (int e = tid; e < TS; e += WB_THREADS) {
            const int d = e % D, la = e / D;
            const size_t g = ((size_t)la * p.h + head) * D + d;
            s_tq[e] = __ldg(p.tq + g);
            s_tk[e] = __ldg(p.tk + g);
            s_tv[e] = __ldg(p.tv + g);
            // the same thread owns these outputs in every fold
            if (p.part_smem) { s_part[e] = 0.f; s_part[TS + e] = 0.f; s_part[2 * TS + e] = 0.f; }
            else { part[g] = 0.f; part[slice + g] = 0.f; part[2 * slice + g] = 0.f; }
        }
        for (int e = tid; e < LA * WB_ROWS; e += WB_THREADS) { s_h0[e] = 0.f; s_h1[e] = 0.f; }
    }
    __syncthreads();
    // part[t][bucket][d] += sum over the group's rows of H[bucket][row] * rows[row][d]; leaves H zeroed
    auto fold = [&](float *H, const float *rows, int n_rows, int t) {
        for (int e = tid; e < TS; e += WB_THREADS) {
            const int d = e % D, la = e / D;
            const float *hrow = H + la * WB_ROWS;
            float acc = 0.f;
            for (int i = 0; i < n_rows; i++) acc = fmaf(hrow[i], rows[i * D + d], acc);
            if (p.part_smem) s_part[t * TS + e] += acc;
            else if (acc != 0.f) part[t * slice + ((size_t)la * p.h + head) * D + d] += acc;
        }
        __syncthreads();
        for (int e = tid; e < LA * WB_ROWS; e += WB_THREADS) H[e] = 0.f;
    };
    // the block's windows: [wb, we) holds the pairs [b, b + 1) * M / gridDim.x
    const long long M = __ldg(p.sq_off + p.n_windows);
    const int wb = blockIdx.x == 0 ? 0 : first_window_at(p.sq_off, p.n_windows, M * blockIdx.x / gridDim.x);
    const int we = blockIdx.x + 1 == gridDim.x ? p.n_windows : first_window_at(p.sq_off, p.n_windows, M * (blockIdx.x + 1) / gridDim.x);
    for (int w0 = wb; w0 < we;) {
        // group = windows [w0, w1): as many as fit WB_ROWS rows (always at least one)
        const int gs = __ldg(p.win_off + w0);
        int w1 = w0 + 1;
        while (w1 < we && __ldg(p.win_off + w1 + 1) - gs <= WB_ROWS) w1++;
        const int grows = __ldg(p.win_off + w1) - gs;    // > WB_ROWS only for a single large window
        // both passes: rows of the group in 64-row pieces (i0), the others of a row's window in 64-row chunks (j0)
        for (int i0 = 0; i0 < grows; i0 += WB_ROWS) {
            const int row = gs + i0 + rs;
            const bool live = row < gs + grows;
            int mw = w0;
            if (live) for (int w = w0 + 1; w < w1 && row >= __ldg(p.win_off + w); w++) mw = w;
            const int ws = __ldg(p.win_off + mw), nw = __ldg(p.win_off + mw + 1) - ws, il = row - ws;
            const long long sq = __ldg(p.sq_off + mw);
            const int n_rows = min(WB_ROWS, grows - i0);
            // ---------------- pass A: the row is a query ----------------
            {
                float q[DL], go[DL], dq[DL];
                float lse = 0.f, Di = 0.f;
#pragma unroll
                for (int d = 0; d < DL; d++) dq[d] = 0.f;
                if (live) {
                    const size_t g = (size_t)row * C + head * D + c0;
                    float o[DL];
                    load_row<DL>(q, p.q + g);
                    load_row<DL>(go, p.dout + g);
                    load_row<DL>(o, p.out + g);
#pragma unroll
                    for (int d = 0; d < DL; d++) Di = fmaf(go[d], o[d], Di);
                    Di = quad_sum(Di, qmask);
                    lse = __ldg(p.lse + (size_t)row * p.h + head);
                    if (REL) {
#pragma unroll
                        for (int d = 0; d < DL; d++) { s_r0[rs * D + c0 + d] = q[d]; s_r1[rs * D + c0 + d] = go[d]; }
                    }
                }
                for (int j0 = 0; j0 < grows; j0 += WB_ROWS) {
                    const int nj = min(WB_ROWS, grows - j0);
                    __syncthreads();
                    if (rs < nj) {
                        const size_t g = (size_t)(gs + j0 + rs) * C + head * D + c0;
                        float r[DL];
                        load_row<DL>(r, p.k + g);
#pragma unroll
                        for (int d = 0; d < DL; d++) s_a[rs * D + c0 + d] = r[d];
                        load_row<DL>(r, p.v + g);
#pragma unroll
                        for (int d = 0; d < DL; d++) s_b[rs * D + c0 + d] = r[d];
                    }
                    __syncthreads();
                    if (!live) continue;
                    // the keys of my window inside this chunk
                    const int jb = max(ws - (gs + j0), 0), je = min(ws + nw - (gs + j0), nj);
                    const int *rel = REL ? p.rel + (sq + (long long)il * nw + (gs + j0 - ws)) * 3 : nullptr;
                    for (int j = jb; j < je; j++) {
                        const float *kj = s_a + j * D + c0, *vj = s_b + j * D + c0;
                        float s = 0.f, dp = 0.f;
                        float tq[DL];
                        int r0 = 0, r1 = 0, r2 = 0;
                        if (REL) {
                            r0 = __ldg(rel + 3 * j); r1 = __ldg(rel + 3 * j + 1); r2 = __ldg(rel + 3 * j + 2);
                            float tk[DL], tv[DL];
                            table_sum<DL>(tq, s_tq + c0, r0, r1, r2, D);
                            table_sum<DL>(tk, s_tk + c0, r0, r1, r2, D);
                            table_sum<DL>(tv, s_tv + c0, r0, r1, r2, D);
#pragma unroll
                            for (int d = 0; d < DL; d++) {
                                s = fmaf(q[d], kj[d] + tq[d], fmaf(kj[d], tk[d], s));
                                dp = fmaf(go[d], vj[d] + tv[d], dp);
                            }
                        } else {
#pragma unroll
                            for (int d = 0; d < DL; d++) { s = fmaf(q[d], kj[d], s); dp = fmaf(go[d], vj[d], dp); }
                        }
                        s = quad_sum(s, qmask);
                        dp = quad_sum(dp, qmask);
                        const float pj = __expf(s - lse);
                        const float ds = pj * (dp - Di);
#pragma unroll
                        for (int d = 0; d < DL; d++) dq[d] = fmaf(ds, REL ? kj[d] + tq[d] : kj[d], dq[d]);
                        if (REL && c < 3) {   // lane a of the quad keeps axis a
                            const int bkt = ((c == 0 ? r0 : c == 1 ? r1 : r2) * 3 + c) * WB_ROWS + rs;
                            s_h0[bkt] += ds;
                            s_h1[bkt] += pj;
                        }
                    }
                }
                if (live) store_row<DL>(p.dq + (size_t)row * C + head * D + c0, dq);
                if (REL) {
                    __syncthreads();
                    fold(s_h0, s_r0, n_rows, 0);   // dTq += Hds x q
                    fold(s_h1, s_r1, n_rows, 2);   // dTv += Hp  x dO
                }
            }
            // ---------------- pass B: the row is a key ----------------
            {
                float kk[DL], vv[DL], dk[DL], dv[DL];
#pragma unroll
                for (int d = 0; d < DL; d++) { dk[d] = 0.f; dv[d] = 0.f; }
                if (live) {
                    const size_t g = (size_t)row * C + head * D + c0;
                    load_row<DL>(kk, p.k + g);
                    load_row<DL>(vv, p.v + g);
                    if (REL) {
#pragma unroll
                        for (int d = 0; d < DL; d++) s_r0[rs * D + c0 + d] = kk[d];
                    }
                }
                for (int j0 = 0; j0 < grows; j0 += WB_ROWS) {
                    const int nj = min(WB_ROWS, grows - j0);
                    __syncthreads();
                    if (rs < nj) {
                        const int qrow = gs + j0 + rs;
                        const size_t g = (size_t)qrow * C + head * D + c0;
                        float r[DL], gg[DL], o[DL];
                        load_row<DL>(r, p.q + g);
                        load_row<DL>(gg, p.dout + g);
                        load_row<DL>(o, p.out + g);
                        float Dq = 0.f;
#pragma unroll
                        for (int d = 0; d < DL; d++) { s_a[rs * D + c0 + d] = r[d]; s_b[rs * D + c0 + d] = gg[d]; Dq = fmaf(gg[d], o[d], Dq); }
                        Dq = quad_sum(Dq, qmask);
                        if (c == 0) { s_c[rs * 2] = __ldg(p.lse + (size_t)qrow * p.h + head); s_c[rs * 2 + 1] = Dq; }
                    }
                    __syncthreads();
                    if (!live) continue;
                    // the queries of my window inside this chunk; pair (i, me) sits at sq + i_local * nw + il
                    const int jb = max(ws - (gs + j0), 0), je = min(ws + nw - (gs + j0), nj);
                    const int *rel = REL ? p.rel + (sq + (long long)(gs + j0 - ws) * nw + il) * 3 : nullptr;
                    for (int j = jb; j < je; j++) {
                        const float *qi = s_a + j * D + c0, *gi = s_b + j * D + c0;
                        float s = 0.f, dp = 0.f;
                        float tk[DL];
                        int r0 = 0, r1 = 0, r2 = 0;
                        if (REL) {
                            const int *rp = rel + (long long)j * nw * 3;
                            r0 = __ldg(rp); r1 = __ldg(rp + 1); r2 = __ldg(rp + 2);
                            float tq[DL], tv[DL];
                            table_sum<DL>(tq, s_tq + c0, r0, r1, r2, D);
                            table_sum<DL>(tk, s_tk + c0, r0, r1, r2, D);
                            table_sum<DL>(tv, s_tv + c0, r0, r1, r2, D);
#pragma unroll
                            for (int d = 0; d < DL; d++) {
                                s = fmaf(qi[d], kk[d] + tq[d], fmaf(kk[d], tk[d], s));
                                dp = fmaf(gi[d], vv[d] + tv[d], dp);
                            }
                        } else {
#pragma unroll
                            for (int d = 0; d < DL; d++) { s = fmaf(qi[d], kk[d], s); dp = fmaf(gi[d], vv[d], dp); }
                        }
                        s = quad_sum(s, qmask);
                        dp = quad_sum(dp, qmask);
                        const float pj = __expf(s - s_c[j * 2]);
                        const float ds = pj * (dp - s_c[j * 2 + 1]);
#pragma unroll
                        for (int d = 0; d < DL; d++) {
                            dk[d] = fmaf(ds, REL ? qi[d] + tk[d] : qi[d], dk[d]);
                            dv[d] = fmaf(pj, gi[d], dv[d]);
                        }
                        if (REL && c < 3) s_h0[((c == 0 ? r0 : c == 1 ? r1 : r2) * 3 + c) * WB_ROWS + rs] += ds;
                    }
                }
                if (live) {
                    store_row<DL>(p.dk + (size_t)row * C + head * D + c0, dk);
                    store_row<DL>(p.dv + (size_t)row * C + head * D + c0, dv);
                }
                if (REL) {
                    __syncthreads();
                    fold(s_h0, s_r0, n_rows, 1);   // dTk += Hds x k
                }
            }
        }
        w0 = w1;
    }
    if (REL && p.part_smem) {
        __syncthreads();
        for (int e = tid; e < TS; e += WB_THREADS) {
            const size_t g = ((size_t)(e / D) * p.h + head) * D + e % D;
            part[g] = s_part[e]; part[slice + g] = s_part[TS + e]; part[2 * slice + g] = s_part[2 * TS + e];
        }
    }
}

// dtable_{q,k,v}[e] = sum over the blocks' partials
__global__ void __launch_bounds__(256) window_attn_tab_reduce_kernel(const float *__restrict__ part, int n_blocks, size_t tab,
                                                                      float *__restrict__ dtq, float *__restrict__ dtk, float *__restrict__ dtv) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= tab) return;
    const int t = blockIdx.y;
    const float *src = part + (size_t)t * n_blocks * tab + e;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int b = 0;
    for (; b + 3 < n_blocks; b += 4) {
        s0 += src[(size_t)b * tab]; s1 += src[(size_t)(b + 1) * tab]; s2 += src[(size_t)(b + 2) * tab]; s3 += src[(size_t)(b + 3) * tab];
    }
    for (; b < n_blocks; b++) s0 += src[(size_t)b * tab];
    (t == 0 ? dtq : (t == 1 ? dtk : dtv))[e] = (s0 + s1) + (s2 + s3);
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

// backward: its shared memory (tables + two per-thread bucket histograms + the block's partial, 100-200 KB) admits one block
// per SM; more blocks than fit only add table-gradient partials to fold
int grid_x_bwd(int n_windows, int h) {
    int g = (U2_NUM_SMS + h - 1) / h;
    if (g < 1) g = 1;
    return g < n_windows ? g : n_windows;
}

// table-free backward: ~17 KB of shared memory, 256 threads: 6 blocks per SM
int grid_x_bwd_plain(int n_windows, int h) {
    int g = (U2_NUM_SMS * 6 + h - 1) / h;
    return g < n_windows ? g : n_windows;
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
    size_t smem = (size_t)(4 * WB_ROWS * D + 2 * WB_ROWS + (REL ? 3 * p.L * 3 * D + 2 * p.L * 3 * WB_ROWS : 0)) * sizeof(float);
    WaParams q = p;
    const size_t part_bytes = REL ? (size_t)3 * p.L * 3 * D * sizeof(float) : 0;
    q.part_smem = REL && smem + part_bytes <= 200 * 1024;   // else the folds add straight into the block's global slice
    if (q.part_smem) smem += part_bytes;
    U2_CUDA_OK(cudaFuncSetAttribute(window_attn_bwd_kernel<D, REL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int gx = REL ? grid_x_bwd(p.n_windows, p.h) : grid_x_bwd_plain(p.n_windows, p.h);
    window_attn_bwd_kernel<D, REL><<<dim3(gx, p.h), WB_THREADS, smem, st>>>(q);
    U2_LAUNCH_OK();
    if (REL) {
        const size_t tab = (size_t)p.L * 3 * p.h * D;
        window_attn_tab_reduce_kernel<<<dim3((unsigned)u2_ceil_div((int64_t)tab, 256), 3), 256, 0, st>>>(p.dtab_part, gx, tab, p.dtq, p.dtk, p.dtv);
        U2_LAUNCH_OK();
    }
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

extern "C" size_t u2_window_attn_bwd_scratch_bytes(int32_t n_windows, int32_t h, int32_t head_dim, int32_t L) {
    if (L <= 0 || n_windows <= 0) return 0;
    return (size_t)3 * grid_x_bwd(n_windows, h) * (size_t)L * 3 * h * head_dim * sizeof(float);
}

extern "C" int u2_window_attn_bwd(const float *q, const float *k, const float *v, const int32_t *win_off, const int32_t *sq_off,
                                  int32_t n_windows, int32_t h, int32_t head_dim, const int32_t *rel_idx, const float *table_q,
                                  const float *table_k, const float *table_v, int32_t L, const float *out, const float *lse,
                                  const float *dout, float *dq, float *dk, float *dv, float *dtable_q, float *dtable_k,
                                  float *dtable_v, void *scratch, size_t scratch_bytes, u2_stream_t stream) {
    WaParams p = {};
    p.q = q; p.k = k; p.v = v; p.win_off = win_off; p.sq_off = sq_off; p.rel = rel_idx; p.tq = table_q; p.tk = table_k; p.tv = table_v;
    p.out = const_cast<float *>(out); p.lse = const_cast<float *>(lse); p.dout = dout; p.dq = dq; p.dk = dk; p.dv = dv;
    p.dtq = dtable_q; p.dtk = dtable_k; p.dtv = dtable_v; p.n_windows = n_windows; p.h = h; p.L = rel_idx ? L : 0;
    if (wa_check(p, head_dim, "u2_window_attn_bwd")) return 1;
    U2_CHECK_ARG(dout && dq && dk && dv && (!rel_idx || (dtable_q && dtable_k && dtable_v)), "u2_window_attn_bwd: null gradient pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (rel_idx) {
        U2_CHECK_ARG(n_windows <= 0 || (scratch && scratch_bytes >= u2_window_attn_bwd_scratch_bytes(n_windows, h, head_dim, L)),
                     "u2_window_attn_bwd: scratch too small (u2_window_attn_bwd_scratch_bytes)");
        p.dtab_part = (float *)scratch;
        if (n_windows <= 0) {  // nothing to fold: the table gradients are zero
            const size_t tb = (size_t)L * 3 * h * head_dim * sizeof(float);
            U2_CUDA_OK(cudaMemsetAsync(dtable_q, 0, tb, st));
            U2_CUDA_OK(cudaMemsetAsync(dtable_k, 0, tb, st));
            U2_CUDA_OK(cudaMemsetAsync(dtable_v, 0, tb, st));
        }
    }
    if (n_windows <= 0) return 0;
    if (head_dim == 16) return rel_idx ? wa_launch_bwd<16, true>(p, st) : wa_launch_bwd<16, false>(p, st);
    return rel_idx ? wa_launch_bwd<32, true>(p, st) : wa_launch_bwd<32, false>(p, st);
}
