// Coordinate hashing, open-addressing hash table, histogram.
// Replaces torchsparse.backend.{hash,kernel_hash,hash_query,count}_cuda (see include/u2mkd.h).
// HBM-bound integer work: 16-byte coalesced coordinate loads, one 16-byte slot per probe.
#include <stdarg.h>
#include <string.h>

#include "u2_common.cuh"

static thread_local char g_err[512] = "";

void u2_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *u2_last_error(void) { return g_err; }
extern "C" int u2_version(void) { return 100; }

// ------------------------------------------------------------------ hash
__global__ void __launch_bounds__(256) hash_kernel(const int4 *__restrict__ coords, int64_t n,
                                                   int64_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int4 c = __ldg(coords + i);
    out[i] = u2_fnv4(c.x, c.y, c.z, c.w);
}

// out[k*n + i]; one thread per point, K hashes each -> K coalesced 8-byte store streams.
__global__ void __launch_bounds__(256) kernel_hash_kernel(const int4 *__restrict__ coords, int64_t n,
                                                          const int *__restrict__ offsets, int K,
                                                          int64_t *__restrict__ out) {
    extern __shared__ int s_off[];
    for (int t = threadIdx.x; t < 3 * K; t += blockDim.x) s_off[t] = offsets[t];
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int4 c = __ldg(coords + i);
    for (int k = 0; k < K; k++)
        out[(int64_t)k * n + i] = u2_fnv4(c.x + s_off[3 * k], c.y + s_off[3 * k + 1], c.z + s_off[3 * k + 2], c.w);
}

extern "C" int u2_hash(const int32_t *coords, int64_t n, const int32_t *offsets, int32_t K, int64_t *out,
                       u2_stream_t stream) {
    if (n == 0) return 0;
    U2_CHECK_ARG(coords && out, "u2_hash: null pointer");
    U2_CHECK_ARG(((uintptr_t)coords & 15) == 0, "u2_hash: coords must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned grid = (unsigned)u2_ceil_div(n, 256);
    if (offsets == nullptr) {
        hash_kernel<<<grid, 256, 0, st>>>((const int4 *)coords, n, out);
    } else {
        U2_CHECK_ARG(K > 0 && K <= 1024, "u2_hash: bad K=%d", K);
        kernel_hash_kernel<<<grid, 256, 3 * K * sizeof(int), st>>>((const int4 *)coords, n, offsets, K, out);
    }
    U2_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------ table
extern "C" size_t u2_hash_table_bytes(int64_t n_keys) { return (size_t)u2_table_capacity(n_keys) * sizeof(U2Slot); }

__global__ void __launch_bounds__(256) table_insert_kernel(const unsigned long long *__restrict__ keys, int64_t n,
                                                           U2Slot *table, unsigned long long mask) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u2_table_insert(table, mask, __ldg(keys + i), (unsigned int)i);
}

__global__ void __launch_bounds__(256) table_query_kernel(const U2Slot *__restrict__ table, unsigned long long mask,
                                                          const unsigned long long *__restrict__ q, int64_t nq,
                                                          int64_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    out[i] = (int64_t)u2_table_lookup(table, mask, __ldg(q + i));
}

extern "C" int u2_hash_table_build(const int64_t *keys, int64_t n_keys, void *table, size_t table_bytes,
                                   u2_stream_t stream) {
    U2_CHECK_ARG(table != nullptr, "u2_hash_table_build: null table");
    U2_CHECK_ARG(table_bytes >= u2_hash_table_bytes(n_keys), "u2_hash_table_build: table too small (%zu < %zu)",
                 table_bytes, u2_hash_table_bytes(n_keys));
    U2_CHECK_ARG(n_keys < 0xFFFFFFFFLL, "u2_hash_table_build: too many keys");
    cudaStream_t st = (cudaStream_t)stream;
    size_t bytes = u2_hash_table_bytes(n_keys);
    U2_CUDA_OK(cudaMemsetAsync(table, 0xFF, bytes, st));
    if (n_keys == 0) return 0;
    unsigned long long mask = bytes / sizeof(U2Slot) - 1;
    table_insert_kernel<<<(unsigned)u2_ceil_div(n_keys, 256), 256, 0, st>>>((const unsigned long long *)keys, n_keys,
                                                                           (U2Slot *)table, mask);
    U2_LAUNCH_OK();
    return 0;
}

extern "C" int u2_hash_table_query(const void *table, size_t table_bytes, const int64_t *queries, int64_t nq,
                                   int64_t *out, u2_stream_t stream) {
    if (nq == 0) return 0;
    U2_CHECK_ARG(table && queries && out, "u2_hash_table_query: null pointer");
    size_t cap = table_bytes / sizeof(U2Slot);
    U2_CHECK_ARG(cap >= 1024 && (cap & (cap - 1)) == 0, "u2_hash_table_query: table_bytes %zu is not a table size",
                 table_bytes);
    table_query_kernel<<<(unsigned)u2_ceil_div(nq, 256), 256, 0, (cudaStream_t)stream>>>(
        (const U2Slot *)table, cap - 1, (const unsigned long long *)queries, nq, out);
    U2_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------ count
// Warp-aggregated histogram: lanes holding the same voxel elect one leader that adds the
// group size (consecutive points usually share a voxel at coarse strides).
__global__ void __launch_bounds__(256) count_kernel(const int *__restrict__ idx, int64_t n, int *__restrict__ out,
                                                    int64_t n_out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int v = (i < n) ? __ldg(idx + i) : -1;
    bool ok = v >= 0 && v < n_out;
    unsigned active = __ballot_sync(0xffffffffu, ok);
    if (!ok) return;
    unsigned peers = __match_any_sync(active, v);
    int leader = __ffs(peers) - 1;
    if ((threadIdx.x & 31) == leader) atomicAdd(out + v, __popc(peers));
}

extern "C" int u2_count(const int32_t *idx, int64_t n, int32_t *out, int64_t n_out, u2_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_out > 0) U2_CUDA_OK(cudaMemsetAsync(out, 0, n_out * sizeof(int), st));
    if (n == 0 || n_out == 0) return 0;
    count_kernel<<<(unsigned)u2_ceil_div(n, 256), 256, 0, st>>>(idx, n, out, n_out);
    U2_LAUNCH_OK();
    return 0;
}
