/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path.
 *
 * CPU restatement (plain C + OpenMP) of the torchsparse v1.4.0 CPU backend that
 * U2MKD's LiDAR point-voxel path runs on (pinned at /root/reference/README.md:44-48,
 * un-vendored: mit-han-lab/torchsparse@v1.4.0, `torchsparse/backend/**_cpu.cpp`).
 * The arithmetic is restated from the published algorithm
 * and anchored on the reference's call sites:
 *   core/models/utils.py:19-26   (sphash / sphashquery / spcount / spvoxelize)
 *   core/models/utils.py:84-99   (kernel-offset hash, calc_ti_weights, spdevoxelize)
 *   core/models/build_blocks.py:25-77 (spnn.Conv3d -> gather / mm / scatter per offset)
 *
 * PARITY UNPINNED: the reference ships no golden vector, KAT or fixture for this
 * path and torchsparse itself is not installable here; the pins are our own
 * dense-equivalence KATs (tests/test_oracle_kat.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- A.5 hashing: FNV-1a over the four 32-bit words, folded to 60 bits ---- */
static inline int64_t fnv_fold(const int32_t c[4]) {
    uint64_t h = 14695981039346656037ULL;
    for (int j = 0; j < 4; j++) {
        h ^= (uint32_t)c[j];
        h *= 1099511628211ULL;
    }
    h = (h >> 60) ^ (h & 0xFFFFFFFFFFFFFFFULL);
    return (int64_t)h;
}

/* hash_cpu: coords int32 [n,4] = (x,y,z,b) -> int64 [n] */
void u2o_hash(const int32_t *coords, int64_t n, int64_t *out) {
#pragma omp parallel for
    for (int64_t i = 0; i < n; i++) out[i] = fnv_fold(coords + 4 * i);
}

/* kernel_hash_cpu: hash of (x+ox, y+oy, z+oz, b) for K offsets -> int64 [K,n] */
void u2o_kernel_hash(const int32_t *coords, int64_t n, const int32_t *offs, int K, int64_t *out) {
#pragma omp parallel for
    for (int64_t i = 0; i < n; i++) {
        for (int k = 0; k < K; k++) {
            int32_t c[4];
            c[0] = coords[4 * i + 0] + offs[3 * k + 0];
            c[1] = coords[4 * i + 1] + offs[3 * k + 1];
            c[2] = coords[4 * i + 2] + offs[3 * k + 2];
            c[3] = coords[4 * i + 3];
            out[(int64_t)k * n + i] = fnv_fold(c);
        }
    }
}

/* ---- A.6 hash_query_cpu: dense_hash_map<int64,int64>, insert keeps the FIRST
 * duplicate, stored value idx+1, miss -> 0; the Python wrapper subtracts 1. ---- */
static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

void u2o_hash_query(const int64_t *q, int64_t nq, const int64_t *ref, const int64_t *ref_idx,
                    int64_t nref, int64_t *out) {
    uint64_t cap = 16;
    while (cap < (uint64_t)(2 * nref + 2)) cap <<= 1;
    int64_t *keys = (int64_t *)malloc(cap * sizeof(int64_t));
    int64_t *vals = (int64_t *)calloc(cap, sizeof(int64_t));
    uint8_t *used = (uint8_t *)calloc(cap, 1);
    for (int64_t i = 0; i < nref; i++) {            /* sequential: first duplicate wins */
        uint64_t s = mix64((uint64_t)ref[i]) & (cap - 1);
        while (used[s] && keys[s] != ref[i]) s = (s + 1) & (cap - 1);
        if (!used[s]) { used[s] = 1; keys[s] = ref[i]; vals[s] = ref_idx[i] + 1; }
    }
#pragma omp parallel for
    for (int64_t i = 0; i < nq; i++) {
        uint64_t s = mix64((uint64_t)q[i]) & (cap - 1);
        int64_t v = 0;
        while (used[s]) {
            if (keys[s] == q[i]) { v = vals[s]; break; }
            s = (s + 1) & (cap - 1);
        }
        out[i] = v;
    }
    free(keys); free(vals); free(used);
}

/* ---- A.7 count_cpu: histogram of non-negative indices ---- */
void u2o_count(const int32_t *idx, int64_t n, int32_t *out, int64_t s) {
    memset(out, 0, (size_t)s * sizeof(int32_t));
    for (int64_t i = 0; i < n; i++) {
        int32_t v = idx[i];
        if (v >= 0 && v < s) out[v]++;
    }
}

#define DEFINE_FP_OPS(T, SUF)                                                                     \
/* A.8 voxelize_forward_cpu: out[idx[i],:] += feat[i,:] / counts[idx[i]]  (divide first) */      \
void u2o_voxelize_fwd_##SUF(const T *feat, int64_t N, int64_t c, const int32_t *idx,             \
                            const int32_t *counts, T *out, int64_t s) {                           \
    memset(out, 0, (size_t)(s * c) * sizeof(T));                                                  \
    for (int64_t i = 0; i < N; i++) {                                                             \
        int32_t p = idx[i];                                                                       \
        if (p < 0 || p >= s) continue;                                                            \
        if (counts[p] <= 0) continue;                                                             \
        T cnt = (T)counts[p];                                                                     \
        for (int64_t j = 0; j < c; j++) out[(int64_t)p * c + j] += feat[i * c + j] / cnt;         \
    }                                                                                             \
}                                                                                                 \
/* voxelize_backward_cpu: grad_feat[i,:] = grad_out[idx[i],:] / counts[idx[i]] */                 \
void u2o_voxelize_bwd_##SUF(const T *gout, int64_t N, int64_t c, const int32_t *idx,             \
                            const int32_t *counts, T *gin, int64_t s) {                           \
    _Pragma("omp parallel for")                                                                   \
    for (int64_t i = 0; i < N; i++) {                                                             \
        int32_t p = idx[i];                                                                       \
        if (p < 0 || p >= s || counts[p] <= 0) {                                                  \
            for (int64_t j = 0; j < c; j++) gin[i * c + j] = (T)0;                                \
            continue;                                                                             \
        }                                                                                         \
        T cnt = (T)counts[p];                                                                     \
        for (int64_t j = 0; j < c; j++) gin[i * c + j] = gout[(int64_t)p * c + j] / cnt;          \
    }                                                                                             \
}                                                                                                 \
/* A.9 devoxelize_forward_cpu: out[i,:] = sum_k w[i,k] * feat[idx[i,k],:]  (idx<0 skipped) */     \
void u2o_devoxelize_fwd_##SUF(const T *feat, int64_t n, int64_t c, const int32_t *idx,            \
                              const T *w, int64_t N, T *out) {                                    \
    (void)n;                                                                                      \
    _Pragma("omp parallel for")                                                                   \
    for (int64_t i = 0; i < N; i++) {                                                             \
        for (int64_t j = 0; j < c; j++) out[i * c + j] = (T)0;                                    \
        for (int k = 0; k < 8; k++) {                                                             \
            int32_t p = idx[i * 8 + k];                                                           \
            if (p < 0) continue;                                                                  \
            T wk = w[i * 8 + k];                                                                  \
            for (int64_t j = 0; j < c; j++) out[i * c + j] += wk * feat[(int64_t)p * c + j];      \
        }                                                                                         \
    }                                                                                             \
}                                                                                                 \
/* devoxelize_backward_cpu: grad_feat[idx[i,k],:] += w[i,k] * grad_out[i,:] */                    \
void u2o_devoxelize_bwd_##SUF(const T *gout, int64_t N, int64_t c, const int32_t *idx,            \
                              const T *w, int64_t n, T *gfeat) {                                  \
    memset(gfeat, 0, (size_t)(n * c) * sizeof(T));                                                \
    for (int64_t i = 0; i < N; i++) {                                                             \
        for (int k = 0; k < 8; k++) {                                                             \
            int32_t p = idx[i * 8 + k];                                                           \
            if (p < 0) continue;                                                                  \
            T wk = w[i * 8 + k];                                                                  \
            for (int64_t j = 0; j < c; j++) gfeat[(int64_t)p * c + j] += wk * gout[i * c + j];    \
        }                                                                                         \
    }                                                                                             \
}                                                                                                 \
/* A.12 gather: buf[i,:] = in[nbmap[2i + t],:] */                                                 \
void u2o_gather_##SUF(const T *in, int64_t c, const int32_t *nbmap, int64_t n_active, int t,     \
                      T *buf) {                                                                   \
    _Pragma("omp parallel for")                                                                   \
    for (int64_t i = 0; i < n_active; i++) {                                                      \
        int32_t p = nbmap[2 * i + t];                                                             \
        if (p < 0) { memset(buf + i * c, 0, (size_t)c * sizeof(T)); continue; }                   \
        memcpy(buf + i * c, in + (int64_t)p * c, (size_t)c * sizeof(T));                          \
    }                                                                                             \
}                                                                                                 \
/* scatter: out[nbmap[2i + 1 - t],:] += buf[i,:]  (indices unique within one offset) */           \
void u2o_scatter_##SUF(const T *buf, int64_t c, const int32_t *nbmap, int64_t n_active, int t,    \
                       T *out) {                                                                  \
    _Pragma("omp parallel for")                                                                   \
    for (int64_t i = 0; i < n_active; i++) {                                                      \
        int32_t p = nbmap[2 * i + 1 - t];                                                         \
        if (p < 0) continue;                                                                      \
        T *o = out + (int64_t)p * c;                                                              \
        const T *b = buf + i * c;                                                                 \
        for (int64_t j = 0; j < c; j++) o[j] += b[j];                                             \
    }                                                                                             \
}

DEFINE_FP_OPS(float, f32)
DEFINE_FP_OPS(double, f64)

int u2o_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
