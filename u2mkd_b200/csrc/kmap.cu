// Strided-conv output coordinates and kernel-map construction.
// Replaces F.spdownsample and the sphash -> sphashquery -> nonzero sequence inside
// torchsparse F.conv3d (see include/u2mkd.h). Integer, HBM/L2-bound: 16-byte coordinate
// loads, 16-byte table slots, neighbour tables written as K coalesced int32 streams.
#include <cub/cub.cuh>

#include "u2_common.cuh"

// ------------------------------------------------------------------ downsample
#define U2_COORD_BITS 18
#define U2_COORD_MAX (1 << U2_COORD_BITS)

__device__ __forceinline__ int floor_div(int a, int b) {
    int q = a / b;
    return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

__global__ void __launch_bounds__(256) downsample_key_kernel(const int4 *__restrict__ coords, int64_t n, int sx, int sy,
                                                             int sz, unsigned long long *__restrict__ keys,
                                                             int *__restrict__ err) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int4 c = __ldg(coords + i);
    int x = floor_div(c.x, sx) * sx, y = floor_div(c.y, sy) * sy, z = floor_div(c.z, sz) * sz;
    if ((unsigned)x >= U2_COORD_MAX || (unsigned)y >= U2_COORD_MAX || (unsigned)z >= U2_COORD_MAX || (unsigned)c.w >= 1024u)
        atomicExch(err, 1);
    keys[i] = ((unsigned long long)(unsigned)c.w << (3 * U2_COORD_BITS)) | ((unsigned long long)(unsigned)x << (2 * U2_COORD_BITS)) |
              ((unsigned long long)(unsigned)y << U2_COORD_BITS) | (unsigned long long)(unsigned)z;
}

__global__ void __launch_bounds__(256) downsample_unpack_kernel(const unsigned long long *__restrict__ keys,
                                                                const int64_t *__restrict__ n_out, int4 *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_out) return;
    unsigned long long k = keys[i];
    const unsigned m = U2_COORD_MAX - 1;
    out[i] = make_int4((int)((k >> (2 * U2_COORD_BITS)) & m), (int)((k >> U2_COORD_BITS) & m), (int)(k & m),
                       (int)(k >> (3 * U2_COORD_BITS)));
}

__global__ void downsample_poison_kernel(const int *err, int64_t *n_out) {
    if (*err) *n_out = -1;
}

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t downsample_cub_bytes(int64_t n) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, a, (const unsigned long long *)nullptr, (unsigned long long *)nullptr, n, 0, 64);
    cub::DeviceSelect::Unique(nullptr, b, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                              (int64_t *)nullptr, n);
    return align_up(a > b ? a : b);
}

extern "C" size_t u2_downsample_scratch_bytes(int64_t n) {
    if (n <= 0) n = 1;
    return 2 * align_up((size_t)n * 8) + 256 + downsample_cub_bytes(n);
}

extern "C" int u2_downsample_coords(const int32_t *coords, int64_t n, int32_t sx, int32_t sy, int32_t sz,
                                    int32_t *out_coords, int64_t *n_out_dev, void *scratch, size_t scratch_bytes,
                                    u2_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    U2_CHECK_ARG(sx > 0 && sy > 0 && sz > 0, "u2_downsample_coords: bad stride");
    U2_CHECK_ARG(scratch_bytes >= u2_downsample_scratch_bytes(n), "u2_downsample_coords: scratch too small");
    U2_CHECK_ARG((((uintptr_t)coords | (uintptr_t)out_coords | (uintptr_t)scratch) & 15) == 0,
                 "u2_downsample_coords: pointers must be 16-byte aligned");
    if (n == 0) {
        U2_CUDA_OK(cudaMemsetAsync(n_out_dev, 0, 8, st));
        return 0;
    }
    char *p = (char *)scratch;
    unsigned long long *keys_a = (unsigned long long *)p; p += align_up((size_t)n * 8);
    unsigned long long *keys_b = (unsigned long long *)p; p += align_up((size_t)n * 8);
    int *err = (int *)p; p += 256;
    size_t cub_bytes = downsample_cub_bytes(n);
    U2_CUDA_OK(cudaMemsetAsync(err, 0, 4, st));
    const unsigned grid = (unsigned)u2_ceil_div(n, 256);
    downsample_key_kernel<<<grid, 256, 0, st>>>((const int4 *)coords, n, sx, sy, sz, keys_a, err);
    U2_LAUNCH_OK();
    U2_CUDA_OK(cub::DeviceRadixSort::SortKeys(p, cub_bytes, keys_a, keys_b, n, 0, 64, st));
    U2_CUDA_OK(cub::DeviceSelect::Unique(p, cub_bytes, keys_b, keys_a, n_out_dev, n, st));
    downsample_unpack_kernel<<<grid, 256, 0, st>>>(keys_a, n_out_dev, (int4 *)out_coords);
    U2_LAUNCH_OK();
    // range violations poison the count so the caller cannot miss them
    downsample_poison_kernel<<<1, 1, 0, st>>>(err, n_out_dev);
    U2_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------ initial voxelisation (index part)
// core/models/utils.py:19-25: pc_hash = sphash(floor(coords)); sparse_hash = torch.unique(pc_hash);
// idx_query = sphashquery(pc_hash, sparse_hash); counts = spcount(idx_query, len(sparse_hash));
// coords = round(spvoxelize(floored, idx_query, counts)) — five operators, a sort inside torch.unique, a hash table
// built only to be queried once, and a scatter-mean of integers whose result is the integer itself.  One call here:
// FNV key per point -> radix sort of (key, point) -> run heads -> exclusive scan = voxel id (ascending key: the
// reference's voxel order) -> idx_query / counts / voxel coordinates scattered from the sorted runs.
__global__ void __launch_bounds__(256) uv_key_kernel(const int4 *__restrict__ coords, int64_t n, unsigned long long *__restrict__ keys,
                                                     int *__restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = __ldg(coords + i);
    keys[i] = (unsigned long long)u2_fnv4(c.x, c.y, c.z, c.w);
    idx[i] = (int)i;
}

__global__ void __launch_bounds__(256) uv_head_kernel(const unsigned long long *__restrict__ keys, int64_t n, int *__restrict__ head) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// vid = inclusive scan of the run heads (1-based voxel id of every sorted position)
__global__ void __launch_bounds__(256) uv_scatter_kernel(const int4 *__restrict__ coords, const int *__restrict__ idx_sorted,
                                                         const int *__restrict__ head, const int *__restrict__ vid, int64_t n,
                                                         int64_t *__restrict__ idx_query, int *__restrict__ counts,
                                                         int4 *__restrict__ voxel_coords, int64_t *__restrict__ n_vox) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int v = vid[j] - 1, pt = idx_sorted[j];
    idx_query[pt] = v;
    if (head[j]) {
        voxel_coords[v] = __ldg(coords + pt);
        // run length = distance to the next head: the sorted positions of a voxel are contiguous
        int64_t e = j + 1;
        while (e < n && !head[e]) e++;
        counts[v] = (int)(e - j);
    }
    if (j == n - 1) *n_vox = (int64_t)v + 1;
}

static size_t uv_cub_bytes(int64_t n) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const unsigned long long *)nullptr, (unsigned long long *)nullptr, (const int *)nullptr,
                                    (int *)nullptr, n, 0, 60);
    cub::DeviceScan::InclusiveSum(nullptr, b, (const int *)nullptr, (int *)nullptr, n);
    return align_up(a > b ? a : b);
}

extern "C" size_t u2_unique_voxelize_scratch_bytes(int64_t n) {
    if (n <= 0) n = 1;
    return 2 * align_up((size_t)n * 8) + 4 * align_up((size_t)n * 4) + uv_cub_bytes(n);
}

// coords int32 [n,4] (already floored) -> idx_query int64 [n], counts int32 [>= n_vox] (caller sizes it n), voxel_coords
// int32 [>= n_vox, 4], n_vox (device int64).  Voxel v = the v-th smallest FNV key, as torch.unique orders them.
extern "C" int u2_unique_voxelize(const int32_t *coords, int64_t n, int64_t *idx_query, int32_t *counts, int32_t *voxel_coords,
                                  int64_t *n_vox_dev, void *scratch, size_t scratch_bytes, u2_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    U2_CHECK_ARG(scratch_bytes >= u2_unique_voxelize_scratch_bytes(n), "u2_unique_voxelize: scratch too small");
    U2_CHECK_ARG((((uintptr_t)coords | (uintptr_t)voxel_coords | (uintptr_t)scratch) & 15) == 0,
                 "u2_unique_voxelize: pointers must be 16-byte aligned");
    U2_CHECK_ARG(n < 0x7FFFFFFFLL, "u2_unique_voxelize: too many points for int32 row indices");
    if (n == 0) {
        U2_CUDA_OK(cudaMemsetAsync(n_vox_dev, 0, 8, st));
        return 0;
    }
    char *p = (char *)scratch;
    unsigned long long *keys_a = (unsigned long long *)p; p += align_up((size_t)n * 8);
    unsigned long long *keys_b = (unsigned long long *)p; p += align_up((size_t)n * 8);
    int *idx_a = (int *)p; p += align_up((size_t)n * 4);
    int *idx_b = (int *)p; p += align_up((size_t)n * 4);
    int *head = (int *)p; p += align_up((size_t)n * 4);
    int *vid = (int *)p; p += align_up((size_t)n * 4);
    size_t cub_bytes = uv_cub_bytes(n);
    const unsigned grid = (unsigned)u2_ceil_div(n, 256);
    uv_key_kernel<<<grid, 256, 0, st>>>((const int4 *)coords, n, keys_a, idx_a);
    U2_LAUNCH_OK();
    U2_CUDA_OK(cub::DeviceRadixSort::SortPairs(p, cub_bytes, keys_a, keys_b, idx_a, idx_b, n, 0, 60, st));
    uv_head_kernel<<<grid, 256, 0, st>>>(keys_b, n, head);
    U2_LAUNCH_OK();
    U2_CUDA_OK(cub::DeviceScan::InclusiveSum(p, cub_bytes, head, vid, n, st));
    uv_scatter_kernel<<<grid, 256, 0, st>>>((const int4 *)coords, idx_b, head, vid, n, idx_query, counts, (int4 *)voxel_coords,
                                            n_vox_dev);
    U2_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------ kernel map
__global__ void __launch_bounds__(256) coord_insert_kernel(const int4 *__restrict__ coords, int64_t n, U2Slot *table,
                                                           unsigned long long mask) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int4 c = __ldg(coords + i);
    u2_table_insert(table, mask, (unsigned long long)u2_fnv4(c.x, c.y, c.z, c.w), (unsigned int)i);
}

// grid.x over output rows (padded to ld_out), grid.y over slices of the K offsets.
__global__ void __launch_bounds__(256) kmap_query_kernel(const U2Slot *__restrict__ table, unsigned long long mask,
                                                         const int4 *__restrict__ out_coords, int64_t n_out, int64_t ld_out,
                                                         const int *__restrict__ offsets, int K, int k_per_slice,
                                                         int *__restrict__ nbr, int *__restrict__ nbrT, int64_t ld_in,
                                                         int *__restrict__ nbsizes) {
    extern __shared__ int smem[];
    int *s_off = smem;           // [3*K]
    int *s_cnt = smem + 3 * K;   // [K]
    for (int t = threadIdx.x; t < 3 * K; t += blockDim.x) s_off[t] = offsets[t];
    for (int t = threadIdx.x; t < K; t += blockDim.x) s_cnt[t] = 0;
    __syncthreads();
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = row < n_out;
    const int k0 = blockIdx.y * k_per_slice;
    const int k1 = min(K, k0 + k_per_slice);
    int4 c = make_int4(0, 0, 0, 0);
    if (live) c = __ldg(out_coords + row);
    for (int k = k0; k < k1; k++) {
        int hit = -1;
        if (live)
            hit = u2_table_lookup(table, mask,
                                  (unsigned long long)u2_fnv4(c.x + s_off[3 * k], c.y + s_off[3 * k + 1], c.z + s_off[3 * k + 2], c.w));
        if (row < ld_out) nbr[(int64_t)k * ld_out + row] = hit;
        if (hit >= 0) nbrT[(int64_t)k * ld_in + hit] = (int)row;
        const unsigned b = __ballot_sync(0xffffffffu, hit >= 0);
        if ((threadIdx.x & 31) == 0 && b) atomicAdd(s_cnt + k, __popc(b));
    }
    __syncthreads();
    for (int k = k0 + threadIdx.x; k < k1; k += blockDim.x)
        if (s_cnt[k]) atomicAdd(nbsizes + k, s_cnt[k]);
}

extern "C" size_t u2_kmap_scratch_bytes(int64_t n_in) { return u2_hash_table_bytes(n_in); }

// Coordinate table of one coordinate set (key = the reference's FNV hash of the row, value = row index): built ONCE per
// tensor stride and shared by the kernel maps, point_to_voxel and voxel_to_point of that stride (the reference rebuilds a
// cuckoo table inside every sphashquery call, 16 x per forward).
extern "C" int u2_coord_table_build(const int32_t *coords, int64_t n, void *table, size_t table_bytes, u2_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    U2_CHECK_ARG(table && table_bytes >= u2_hash_table_bytes(n), "u2_coord_table_build: table too small");
    U2_CHECK_ARG((((uintptr_t)coords | (uintptr_t)table) & 15) == 0, "u2_coord_table_build: pointers must be 16-byte aligned");
    U2_CHECK_ARG(n < 0x7FFFFFFFLL, "u2_coord_table_build: too many rows");
    const size_t tbytes = u2_hash_table_bytes(n);
    U2_CUDA_OK(cudaMemsetAsync(table, 0xFF, tbytes, st));
    if (n > 0) {
        coord_insert_kernel<<<(unsigned)u2_ceil_div(n, 256), 256, 0, st>>>((const int4 *)coords, n, (U2Slot *)table,
                                                                          tbytes / sizeof(U2Slot) - 1);
        U2_LAUNCH_OK();
    }
    return 0;
}

// out[k * nq + i] = row of (q[i].xyz + offsets[k], q[i].b) in the table's coordinate set, -1 if absent (K = 1, offsets =
// NULL: the plain lookup of point_to_voxel; K = 8: the corner lookup of voxel_to_point, core/models/utils.py:84-93 —
// sphash with offsets + sphashquery without the [K, N] int64 hash tensor in between).
__global__ void __launch_bounds__(256) coord_query_kernel(const U2Slot *__restrict__ table, unsigned long long mask,
                                                          const int4 *__restrict__ q, int64_t nq, const int *__restrict__ offsets,
                                                          int K, int64_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const int4 c = __ldg(q + i);
    for (int k = 0; k < K; k++) {
        const int ox = offsets ? __ldg(offsets + 3 * k) : 0, oy = offsets ? __ldg(offsets + 3 * k + 1) : 0,
                  oz = offsets ? __ldg(offsets + 3 * k + 2) : 0;
        out[(int64_t)k * nq + i] = (int64_t)u2_table_lookup(table, mask, (unsigned long long)u2_fnv4(c.x + ox, c.y + oy, c.z + oz, c.w));
    }
}

extern "C" int u2_coord_table_query(const void *table, size_t table_bytes, const int32_t *qcoords, int64_t nq,
                                    const int32_t *offsets, int32_t K, int64_t *out, u2_stream_t stream) {
    if (nq == 0) return 0;
    U2_CHECK_ARG(table && qcoords && out && K >= 1, "u2_coord_table_query: bad arguments");
    const size_t cap = table_bytes / sizeof(U2Slot);
    U2_CHECK_ARG(cap >= 1024 && (cap & (cap - 1)) == 0, "u2_coord_table_query: table_bytes %zu is not a table size", table_bytes);
    U2_CHECK_ARG(((uintptr_t)qcoords & 15) == 0, "u2_coord_table_query: coordinates must be 16-byte aligned");
    coord_query_kernel<<<(unsigned)u2_ceil_div(nq, 256), 256, 0, (cudaStream_t)stream>>>((const U2Slot *)table, cap - 1,
                                                                                         (const int4 *)qcoords, nq, offsets, K, out);
    U2_LAUNCH_OK();
    return 0;
}

// scratch_bytes == 0: `scratch` is a coordinate table of in_coords that u2_coord_table_build already filled
// (u2_hash_table_bytes(n_in) bytes); otherwise it is built here.
extern "C" int u2_kmap_build(const int32_t *in_coords, int64_t n_in, const int32_t *out_coords, int64_t n_out,
                             const int32_t *offsets, int32_t K, int32_t *nbr, int64_t ld_out, int32_t *nbrT, int64_t ld_in,
                             int32_t *nbsizes, void *scratch, size_t scratch_bytes, u2_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const bool prebuilt = scratch_bytes == 0 && scratch != nullptr;
    if (prebuilt) scratch_bytes = u2_kmap_scratch_bytes(n_in);
    U2_CHECK_ARG(K > 0 && K <= 1024, "u2_kmap_build: bad K=%d", K);
    U2_CHECK_ARG(ld_out >= n_out && ld_in >= n_in, "u2_kmap_build: leading dimensions too small");
    U2_CHECK_ARG(scratch_bytes >= u2_kmap_scratch_bytes(n_in), "u2_kmap_build: scratch too small");
    U2_CHECK_ARG((((uintptr_t)in_coords | (uintptr_t)out_coords | (uintptr_t)scratch) & 15) == 0,
                 "u2_kmap_build: pointers must be 16-byte aligned");
    U2_CHECK_ARG(n_in < 0x7FFFFFFFLL && n_out < 0x7FFFFFFFLL, "u2_kmap_build: too many rows");
    U2_CUDA_OK(cudaMemsetAsync(nbsizes, 0, K * sizeof(int), st));
    if (ld_in > 0) U2_CUDA_OK(cudaMemsetAsync(nbrT, 0xFF, (size_t)K * ld_in * sizeof(int), st));
    if (ld_out == 0) return 0;
    const size_t tbytes = u2_hash_table_bytes(n_in);
    const unsigned long long mask = tbytes / sizeof(U2Slot) - 1;
    if (!prebuilt) {
        U2_CUDA_OK(cudaMemsetAsync(scratch, 0xFF, tbytes, st));
        if (n_in > 0) {
            coord_insert_kernel<<<(unsigned)u2_ceil_div(n_in, 256), 256, 0, st>>>((const int4 *)in_coords, n_in,
                                                                                (U2Slot *)scratch, mask);
            U2_LAUNCH_OK();
        }
    }
    const int64_t row_blocks = u2_ceil_div(ld_out, 256);
    int slices = 1;
    while (slices < K && row_blocks * slices < (int64_t)U2_NUM_SMS * 16) slices++;
    const int k_per_slice = (K + slices - 1) / slices;
    slices = (K + k_per_slice - 1) / k_per_slice;
    dim3 grid((unsigned)row_blocks, (unsigned)slices);
    kmap_query_kernel<<<grid, 256, 4 * K * sizeof(int), st>>>((const U2Slot *)scratch, mask, (const int4 *)out_coords, n_out,
                                                              ld_out, offsets, K, k_per_slice, nbr, nbrT, ld_in, nbsizes);
    U2_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------ compacted pair list
// flat[j] = k*ld + out_row of the j-th valid entry of nbr, ascending: the reference's
// (k, out)-ordered nbmaps without the host-side nonzero(); consumed by the wgrad kernels.
struct U2ValidEntry {
    const int *nbr;
    __device__ __forceinline__ bool operator()(const int &f) const { return nbr[f] >= 0; }
};

extern "C" size_t u2_kmap_pairs_scratch_bytes(int64_t total) {
    size_t b = 0;
    cub::DeviceSelect::If(nullptr, b, cub::CountingInputIterator<int>(0), (int *)nullptr, (int *)nullptr, (int)total,
                          U2ValidEntry{nullptr});
    return align_up(b) + 256;
}

extern "C" int u2_kmap_pairs(const int32_t *nbr, int64_t total, int32_t *flat, void *scratch, size_t scratch_bytes,
                             u2_stream_t stream) {
    U2_CHECK_ARG(total < 0x7FFFFFFFLL, "u2_kmap_pairs: table too large");
    U2_CHECK_ARG(scratch_bytes >= u2_kmap_pairs_scratch_bytes(total), "u2_kmap_pairs: scratch too small");
    if (total == 0) return 0;
    int *n_sel = (int *)scratch;
    size_t cub_bytes = scratch_bytes - 256;
    U2_CUDA_OK(cub::DeviceSelect::If((char *)scratch + 256, cub_bytes, cub::CountingInputIterator<int>(0), flat, n_sel,
                                     (int)total, U2ValidEntry{nbr}, (cudaStream_t)stream));
    return 0;
}

// ------------------------------------------------------------------ mask-sorted tile order
// The tensor-core conv walks 128-row tiles and, per tile, only the offsets that have at least
// one neighbour in the tile.  Sorting the rows by their neighbour-presence mask (rare offsets
// in the high bits) makes the rows of a tile agree on which offsets they use: ~2.7x fewer
// (tile, offset) work items on LiDAR scans.  Rows keep their identity through `perm`
// (tile row -> original row); results are written back to the original rows, so nothing
// observable changes order.
struct U2BitPos { int pos[32]; };

__global__ void __launch_bounds__(256) rowkey_kernel(const int *__restrict__ tab, int64_t ld, int64_t n, int K, U2BitPos bp,
                                                     unsigned long long *__restrict__ keys) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    unsigned int m = 0;
    for (int k = 0; k < K; k++)
        if (__ldg(tab + (int64_t)k * ld + r) >= 0) m |= 1u << bp.pos[k];
    keys[r] = ((unsigned long long)m << 32) | (unsigned long long)r;
}

__global__ void __launch_bounds__(256) perm_from_keys_kernel(const unsigned long long *__restrict__ keys, int64_t n, int64_t ld,
                                                             int *__restrict__ perm) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ld) return;
    perm[j] = j < n ? (int)(keys[j] & 0xFFFFFFFFull) : -1;
}

__global__ void __launch_bounds__(256) permute_table_kernel(const int *__restrict__ tab, int64_t ld, int K,
                                                            const int *__restrict__ perm, int *__restrict__ tabP) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ld) return;
    const int r = __ldg(perm + j);
    const int k = blockIdx.y;
    tabP[(int64_t)k * ld + j] = r >= 0 ? __ldg(tab + (int64_t)k * ld + r) : -1;
}

// one warp per 128-row tile: bit k set iff some row of the tile has a neighbour at offset k
__global__ void __launch_bounds__(256) tile_mask_kernel(const int *__restrict__ tabP, int64_t ld, int K,
                                                        unsigned int *__restrict__ tile_mask) {
    const int64_t tile = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (tile >= ld / 128) return;
    unsigned int m = 0;
    for (int k = 0; k < K; k++) {
        const int4 v = __ldg(reinterpret_cast<const int4 *>(tabP + (int64_t)k * ld + tile * 128) + lane);
        const bool any = v.x >= 0 || v.y >= 0 || v.z >= 0 || v.w >= 0;
        if (__any_sync(0xffffffffu, any)) m |= 1u << k;
    }
    if (lane == 0) tile_mask[tile] = m;
}

static size_t sort_cub_bytes(int64_t n) {
    size_t b = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, b, (const unsigned long long *)nullptr, (unsigned long long *)nullptr, n, 32, 64);
    return align_up(b);
}

extern "C" size_t u2_kmap_sort_scratch_bytes(int64_t n_rows) {
    if (n_rows <= 0) n_rows = 1;
    return 2 * align_up((size_t)n_rows * 8) + sort_cub_bytes(n_rows);
}

extern "C" int u2_kmap_sort_rows(const int32_t *table, int64_t ld, int64_t n_rows, int32_t K, const int32_t *bitpos_host,
                                 const int32_t *perm_in, int32_t *perm_out, int32_t *tableP, uint32_t *tile_mask,
                                 void *scratch, size_t scratch_bytes, u2_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    U2_CHECK_ARG(K > 0 && K <= 32, "u2_kmap_sort_rows: K=%d (needs 1..32)", K);
    U2_CHECK_ARG(ld >= n_rows && n_rows < 0x7FFFFFFFLL, "u2_kmap_sort_rows: bad sizes");
    if (ld == 0) return 0;
    const unsigned grid_ld = (unsigned)u2_ceil_div(ld, 256);
    const int *perm = perm_in;
    if (perm_in == nullptr) {
        U2_CHECK_ARG(scratch_bytes >= u2_kmap_sort_scratch_bytes(n_rows), "u2_kmap_sort_rows: scratch too small");
        U2BitPos bp;
        for (int k = 0; k < 32; k++) bp.pos[k] = k < K ? bitpos_host[k] : 0;
        for (int k = 0; k < K; k++) U2_CHECK_ARG(bp.pos[k] >= 0 && bp.pos[k] < 32, "u2_kmap_sort_rows: bad bit position");
        char *p = (char *)scratch;
        unsigned long long *keys_a = (unsigned long long *)p; p += align_up((size_t)(n_rows > 0 ? n_rows : 1) * 8);
        unsigned long long *keys_b = (unsigned long long *)p; p += align_up((size_t)(n_rows > 0 ? n_rows : 1) * 8);
        if (n_rows > 0) {
            rowkey_kernel<<<(unsigned)u2_ceil_div(n_rows, 256), 256, 0, st>>>(table, ld, n_rows, K, bp, keys_a);
            U2_LAUNCH_OK();
            size_t cub_bytes = sort_cub_bytes(n_rows);
            U2_CUDA_OK(cub::DeviceRadixSort::SortKeys(p, cub_bytes, keys_a, keys_b, n_rows, 32, 32 + K, st));
        }
        perm_from_keys_kernel<<<grid_ld, 256, 0, st>>>(keys_b, n_rows, ld, perm_out);
        U2_LAUNCH_OK();
        perm = perm_out;
    }
    permute_table_kernel<<<dim3(grid_ld, (unsigned)K), 256, 0, st>>>(table, ld, K, perm, tableP);
    U2_LAUNCH_OK();
    if (tile_mask) {
        U2_CHECK_ARG(ld % 128 == 0 && ((uintptr_t)tableP & 15) == 0, "u2_kmap_sort_rows: tile masks need ld %% 128 == 0");
        tile_mask_kernel<<<(unsigned)u2_ceil_div(ld / 128 * 32, 256), 256, 0, st>>>(tableP, ld, K, tile_mask);
        U2_LAUNCH_OK();
    }
    return 0;
}
