// Microbenchmark: how fast can one SM gather 128-byte row slices from an L2-resident matrix into shared memory?
// This is the A-operand path of the sparse conv kernels (fwd / dgrad / wgrad all live on it).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gather_bench gather_bench.cu -lcuda
//   ./gather_bench            (prints one JSON line per configuration)
//
// One "item" = 128 rows x 128 bytes = 16 KB into one pipeline stage (the conv kernel's stage geometry).  A consumer
// thread frees the stage as soon as it is full (no MMA), so the number measures the producers alone.
// methods: 0 LDGSTS.ca 16 B (8 lanes per row, 4 rows per warp instruction — the product's scheme)
//          1 LDGSTS.cg
//          2 LDGSTS.ca, one lane per row (32 rows per instruction, 8 instructions per row slice)
//          3 LDG.128 -> STS.128 through registers (8 lanes per row)
//          4 cp.async.bulk 128 B per row, one thread per row (TMA engine, no tensor map; linear smem rows)
//          5 TMA tile::gather4 (4 rows per instruction, SWIZZLE_128B), 32 per item spread over the W warps
//          6 LDGSTS.ca only for present rows, a zeroing STS.128 for absent ones (divergent)
//          7 LDGSTS.ca only for present rows, nothing for absent ones (lower bound of 6)
//          8 hybrid: rows 0-63 by LDGSTS.ca (all W warps), rows 64-127 by 16 gather4 (lanes 0-3 of warps 0-3)
//         10 LDGSTS.ca free-running: no barriers, no consumer (pure issue throughput; cp.async.wait_all at the end)
//         11 LDGSTS.ca, completion by cp.async.commit_group / wait_group 1 per thread + one mbarrier arrive per warp
//         12 COMPACTED LDGSTS.ca: only the present rows of an item are fetched, 4 per warp instruction whatever their slots
//            (ceil(m / 4) instructions instead of 32), absent slots zeroed by STS.128 (upper bound: every absent row, every item)
//         13 = 12 without the zeroing stores
//         14 LDGSTS.ca with lane -> (row = lane % 8, chunk = lane / 8 + 4 h): one instruction writes FULL 128-byte
//            shared-memory lines (8 rows x 16 B of one chunk plane) and reads 64 B of 8 rows (W = 4 only)
//          9 cp.async.bulk of the WHOLE row (pitch bytes) per present row, one thread per row (item = 128 whole rows)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > 2000000000LL) __trap();
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes, int cg) {
    if (cg) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap *tmap, int col, int4 rows, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
                 " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)),
                 "r"(col), "r"(rows.x), "r"(rows.y), "r"(rows.z), "r"(rows.w) : "memory");
}

struct Params {
    const uint8_t *X;      // [n][pitch] bytes
    const int *idx;        // [items][128]
    int n_items, pitch, slices;  // slices = 128-byte slices per row walked per item index (items are (index set, slice))
    int method, W, S;
    unsigned long long *sink;
};

constexpr int STAGE = 128 * 128 + 128 * 16;  // room for the 16-byte skew of the product's no-swizzle layout

__global__ void __launch_bounds__(544) gather_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)p.S * STAGE);
    uint64_t *empty = full + 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_prod = p.W * 32;
    if (tid == 0) {
        for (int s = 0; s < p.S; s++) {
            // LDGSTS methods: one noinc arrival per producer thread; bulk / gather4: one arrive.expect_tx per issuing thread
            int cnt = (p.method <= 2) ? n_prod : (p.method == 3 ? p.W : (p.method == 4 ? 128 : p.W));
            if (p.method == 6 || p.method == 7) cnt = n_prod;
            if (p.method == 8) cnt = n_prod + 16;
            if (p.method == 9) cnt = 128;
            if (p.method == 11) cnt = p.W;
            if (p.method == 12 || p.method == 13 || p.method == 14) cnt = n_prod;
            mbar_init(full + s, cnt);
            mbar_init(empty + s, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // items of this CTA: blockIdx.x, + gridDim.x, ...
    if (warp < p.W) {
        int s = 0;
        uint32_t ph = 0;
        // indices of the NEXT item are fetched one item ahead (the product kernels read them from shared memory: no
        // dependent global load in front of a gather)
        const int pr0 = tid >> 3, prstep = n_prod >> 3;
        int nx[8];
        int4 nx4 = make_int4(-1, -1, -1, -1);
        auto prefetch = [&](int it) {
            const int *rows = p.idx + (size_t)(it / p.slices) * 128;
#pragma unroll
            for (int i = 0; i < 8; i++) nx[i] = (it < p.n_items && pr0 + i * prstep < 128) ? __ldg(rows + pr0 + i * prstep) : -1;
            if (p.method == 5 && it < p.n_items) {
                const int per_warp = 32 / p.W;
                if (lane < per_warp) nx4 = __ldg(reinterpret_cast<const int4 *>(rows) + warp * per_warp + lane);
            }
            if (p.method == 8 && it < p.n_items && warp < 4 && lane < 4) nx4 = __ldg(reinterpret_cast<const int4 *>(rows) + 16 + warp * 4 + lane);
        };
        prefetch(blockIdx.x);
        for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
            const int *rows = p.idx + (size_t)(it / p.slices) * 128;
            const int slice = it % p.slices;
            const uint8_t *xs = p.X + (size_t)slice * 128;
            int cur[8];
#pragma unroll
            for (int i = 0; i < 8; i++) cur[i] = nx[i];
            const int4 cur4 = nx4;
            prefetch(it + gridDim.x);
            if (p.method != 10) mbar_wait(empty + s, ph ^ 1u);
            const uint32_t base = smem_u32(smem + (size_t)s * STAGE);
            if (p.method == 0 || p.method == 1) {
                // thread -> (row = tid / 8 + i * n_prod / 8, chunk = tid % 8)
                const int chunk = tid & 7;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int r = pr0 + i * prstep;
                    if (r < 128) {
                        const int src = cur[i];
                        const bool ok = src >= 0;
                        cp_async16(base + chunk * (2048 + 16) + r * 16, xs + (ok ? (size_t)src * p.pitch + chunk * 16 : 0), ok ? 16u : 0u,
                                   p.method);
                    }
                }
                cp_async_arrive_noinc(full + s);
            } else if (p.method == 2) {
                for (int r = tid; r < 128; r += n_prod) {
                    const int src = __ldg(rows + r);
                    const bool ok = src >= 0;
#pragma unroll
                    for (int c = 0; c < 8; c++)
                        cp_async16(base + c * (2048 + 16) + r * 16, xs + (ok ? (size_t)src * p.pitch + c * 16 : 0), ok ? 16u : 0u, 0);
                }
                cp_async_arrive_noinc(full + s);
            } else if (p.method == 3) {
                const int chunk = tid & 7, r0 = tid >> 3, rstep = n_prod >> 3;
                uint4 v[16];
                int k = 0;
                for (int r = r0; r < 128; r += rstep, k++) {
                    const int src = __ldg(rows + r);
                    v[k] = src >= 0 ? __ldg(reinterpret_cast<const uint4 *>(xs + (size_t)src * p.pitch + chunk * 16)) : make_uint4(0, 0, 0, 0);
                }
                k = 0;
                for (int r = r0; r < 128; r += rstep, k++)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + chunk * (2048 + 16) + r * 16), "r"(v[k].x),
                                 "r"(v[k].y), "r"(v[k].z), "r"(v[k].w) : "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(full + s);
            } else if (p.method == 14) {
                // warp w owns rows 32 w .. 32 w + 31 (W == 4); instruction (g, h): rows 8 g + lane % 8, chunk 4 h + lane / 8
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    const int r = warp * 32 + g * 8 + (lane & 7);
                    const int src = __shfl_sync(0xFFFFFFFFu, cur[g], (lane & 7) * 1);  // placeholder, replaced below
                    (void)src;
                }
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    const int r = warp * 32 + g * 8 + (lane & 7);
                    const int srcr = __ldg(rows + r);
                    const bool ok = srcr >= 0;
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int chunk = 4 * h + (lane >> 3);
                        cp_async16(base + chunk * (2048 + 16) + r * 16, xs + (ok ? (size_t)srcr * p.pitch + chunk * 16 : 0), ok ? 16u : 0u, 0);
                    }
                }
                cp_async_arrive_noinc(full + s);
            } else if (p.method == 12 || p.method == 13) {
                // compact the present rows of the item (ballot over the 128 indices, 32 per warp-sized group), then fetch
                // present row j with the 8-lane group j % (n_prod / 8)
                __shared__ unsigned char s_list[8][128];  // one list per pipeline stage (<= 8)
                __shared__ int s_cnt[8];
                // every warp builds the same list redundantly from the prefetched indices? -> cheaper: warp 0 builds it
                if (warp == 0) {
                    int m = 0;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int v = __ldg(rows + q * 32 + lane);
                        const unsigned b = __ballot_sync(0xFFFFFFFFu, v >= 0);
                        if (v >= 0) s_list[s][m + __popc(b & ((1u << lane) - 1u))] = (unsigned char)(q * 32 + lane);
                        m += __popc(b);
                    }
                    if (lane == 0) s_cnt[s] = m;
                }
                asm volatile("bar.sync 1, %0;" ::"r"(n_prod) : "memory");
                const int m = s_cnt[s];
                const int chunk = tid & 7, g = tid >> 3, ng = n_prod >> 3;
                for (int j = g; j < m; j += ng) {
                    const int slot = s_list[s][j];
                    const int src = __ldg(rows + slot);
                    cp_async16(base + chunk * (2048 + 16) + slot * 16, xs + (size_t)src * p.pitch + chunk * 16, 16u, 0);
                }
                if (p.method == 12) {
                    for (int r = tid >> 3; r < 128; r += ng) {
                        if (__ldg(rows + r) < 0)
                            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + chunk * (2048 + 16) + r * 16), "r"(0) : "memory");
                    }
                }
                cp_async_arrive_noinc(full + s);
            } else if (p.method == 10 || p.method == 11) {
                const int chunk = tid & 7;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int r = pr0 + i * prstep;
                    if (r < 128) {
                        const int src = cur[i];
                        const bool ok = src >= 0;
                        cp_async16(base + chunk * (2048 + 16) + r * 16, xs + (ok ? (size_t)src * p.pitch + chunk * 16 : 0), ok ? 16u : 0u, 0);
                    }
                }
                if (p.method == 11) {
                    // signal the PREVIOUS item of this warp: its group has had one item's worth of time to land
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    asm volatile("cp.async.wait_group 1;" ::: "memory");
                    __syncwarp();
                    if (it != (int)blockIdx.x && lane == 0) mbar_arrive(full + (s == 0 ? p.S - 1 : s - 1));
                }
            } else if (p.method == 6 || p.method == 7) {
                const int chunk = tid & 7;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int r = pr0 + i * prstep;
                    if (r < 128) {
                        const int src = cur[i];
                        const uint32_t dst = base + chunk * (2048 + 16) + r * 16;
                        if (src >= 0) cp_async16(dst, xs + (size_t)src * p.pitch + chunk * 16, 16u, 0);
                        else if (p.method == 6) asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0) : "memory");
                    }
                }
                cp_async_arrive_noinc(full + s);
            } else if (p.method == 8) {
                // rows 0-63: LDGSTS into the SWIZZLE_128B layout (row pitch 128 B, chunk ^ (row & 7)); rows 64-127: gather4
                const int chunk = tid & 7;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int r = pr0 + i * prstep;
                    if (r < 64) {
                        const int src = cur[i];
                        const bool ok = src >= 0;
                        cp_async16(base + r * 128 + ((chunk ^ (r & 7)) << 4), xs + (ok ? (size_t)src * p.pitch + chunk * 16 : 0), ok ? 16u : 0u, 0);
                    }
                }
                cp_async_arrive_noinc(full + s);
                if (warp < 4 && lane < 4) {
                    const int g = 16 + warp * 4 + lane;
                    mbar_arrive_expect_tx(full + s, 512u);
                    tma_gather4(base + g * 512, &tmap, slice * 64, cur4, full + s);
                }
            } else if (p.method == 9) {
                if (tid < 128) {
                    const int src = __ldg(rows + tid);
                    if (src >= 0 && slice == 0) {
                        mbar_arrive_expect_tx(full + s, (uint32_t)p.pitch);
                        bulk_g2s(base + (tid % 48) * 384, p.X + (size_t)src * p.pitch, (uint32_t)p.pitch, full + s);
                    } else {
                        mbar_arrive(full + s);
                    }
                }
            } else if (p.method == 4) {
                // one thread per row (the first 128 producer threads); rows land linearly (128 B pitch)
                if (tid < 128) {
                    const int src = __ldg(rows + tid);
                    if (src >= 0) {
                        mbar_arrive_expect_tx(full + s, 128u);
                        bulk_g2s(base + tid * 128, xs + (size_t)src * p.pitch, 128u, full + s);
                    } else {
                        mbar_arrive(full + s);
                    }
                }
            } else {
                // gather4: 32 instructions per item spread over W warps (lanes 0 .. 32/W-1 of each warp issue one each),
                // then lane 0 of the warp posts the expected bytes of its share
                const int per_warp = 32 / p.W;
                if (lane == 0) mbar_arrive_expect_tx(full + s, (uint32_t)per_warp * 512u);
                __syncwarp();
                if (lane < per_warp) {
                    const int g = warp * per_warp + lane;
                    tma_gather4(base + g * 512, &tmap, slice * 64, cur4, full + s);
                }
            }
            if (++s == p.S) { s = 0; ph ^= 1u; }
        }
        if (p.method == 10 || p.method == 11) {
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
            if (p.method == 11 && lane == 0 && blockIdx.x < p.n_items) mbar_arrive(full + (s == 0 ? p.S - 1 : s - 1));
        }
    } else if (warp == p.W) {
        // consumer: frees a stage as soon as it is full (reads one word so that the data dependency is real)
        int s = 0;
        uint32_t ph = 0;
        unsigned long long acc = 0;
        for (int it = blockIdx.x; it < p.n_items && p.method != 10; it += gridDim.x) {
            mbar_wait(full + s, ph);
            if (lane == 0) {
                acc += *reinterpret_cast<volatile unsigned int *>(smem + (size_t)s * STAGE + 64);
                mbar_arrive(empty + s);
            }
            __syncwarp();
            if (++s == p.S) { s = 0; ph ^= 1u; }
        }
        if (lane == 0 && acc == 0x1234567ULL) *p.sink = acc;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
    const int n = 237144;
    int sm_count = 0, clock_khz = 0;
    CK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
    std::mt19937 rng(1);
    EncodeTiledFn encode = nullptr;
    {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            encode = (EncodeTiledFn)ptr;
    }
    unsigned long long *sink;
    CK(cudaMalloc(&sink, 8));
    // argv[1]: comma-separated methods (default 0,4,5,9,12,13,14); argv[2]: comma-separated row pitches in bytes (default 128,384)
    bool want[16] = {};
    std::vector<int> pitches;
    {
        const char *ms = argc > 1 ? argv[1] : "0,4,5,9,12,13,14";
        for (const char *c = ms; *c;) { int m = atoi(c); if (m >= 0 && m < 16) want[m] = true; while (*c && *c != ',') c++; if (*c) c++; }
        const char *ps = argc > 2 ? argv[2] : "128,384";
        for (const char *c = ps; *c;) { pitches.push_back(atoi(c)); while (*c && *c != ',') c++; if (*c) c++; }
    }
    for (int pitch : pitches) {
        const int slices = pitch / 128;
        uint8_t *X;
        CK(cudaMalloc(&X, (size_t)n * pitch));
        CK(cudaMemset(X, 1, (size_t)n * pitch));
        alignas(64) CUtensorMap tmap;
        memset(&tmap, 0, sizeof tmap);
        if (encode) {
            const cuuint64_t gdim[2] = {(cuuint64_t)(pitch / 2), (cuuint64_t)n};
            const cuuint64_t gstride[1] = {(cuuint64_t)pitch};
            const cuuint32_t box[2] = {64u, 1u};
            const cuuint32_t estr[2] = {1u, 1u};
            CUresult rc = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, X, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (rc != CUDA_SUCCESS) { fprintf(stderr, "encode failed %d\n", (int)rc); encode = nullptr; }
        }
        const int n_sets = 16384;  // index sets of 128 rows; items = n_sets * slices
        for (const char *pattern : {"random"}) {
            for (double zf : {0.0, 0.47}) {
                std::vector<int> idx((size_t)n_sets * 128);
                std::uniform_int_distribution<int> uni(0, n - 1), win(-256, 256);
                std::uniform_real_distribution<double> u01(0, 1);
                for (int s = 0; s < n_sets; s++) {
                    const int base = (int)((long long)s * (n - 1024) / n_sets) + 512;
                    for (int r = 0; r < 128; r++) {
                        int v = !strcmp(pattern, "random") ? uni(rng) : (!strcmp(pattern, "local") ? base + win(rng) : base + r);
                        if (u01(rng) < zf) v = -1;
                        idx[(size_t)s * 128 + r] = v;
                    }
                }
                int *d_idx;
                CK(cudaMalloc(&d_idx, idx.size() * 4));
                CK(cudaMemcpy(d_idx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice));
                long long real_rows = 0;
                for (int v : idx) real_rows += v >= 0;
                for (int method = 0; method <= 14; method++) {
                    if (!want[method]) continue;
                    if ((method == 5 || method == 8) && !encode) continue;
                    if (method == 9 && pitch == 128) continue;
                    for (int W : {4, 8}) {
                        if (method == 5 && W < 4) continue;
                        if ((method == 4 || method == 9 || method == 14) && W != 4) continue;
                        for (int ctas : {1, 2, 3, 4}) {
                            for (int S : {2, 4, 6}) {
                                if (ctas * (W + 1) > 64 || (size_t)ctas * (S * STAGE + 3072) > 226 * 1024) continue;
                                if (method == 3 && W == 4) continue;  // 32 rows per thread would not fit the register array
                                Params p;
                                p.X = X; p.idx = d_idx; p.n_items = n_sets * slices; p.pitch = pitch; p.slices = slices;
                                p.method = method; p.W = W; p.S = S; p.sink = sink;
                                const size_t smem = (size_t)S * STAGE + 1024 + 512;
                                CK(cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                                CK(cudaFuncSetAttribute(gather_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
                                const int threads = (W + 1) * 32, grid = sm_count * ctas;
                                cudaEvent_t e0, e1;
                                CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
                                gather_kernel<<<grid, threads, smem>>>(tmap, p);
                                CK(cudaDeviceSynchronize());
                                float best = 1e9f;
                                for (int rep = 0; rep < 3; rep++) {
                                    CK(cudaEventRecord(e0));
                                    gather_kernel<<<grid, threads, smem>>>(tmap, p);
                                    CK(cudaEventRecord(e1));
                                    CK(cudaEventSynchronize(e1));
                                    float ms;
                                    CK(cudaEventElapsedTime(&ms, e0, e1));
                                    best = ms < best ? ms : best;
                                }
                                const double slot_bytes = (double)p.n_items * 16384.0, real_bytes = (double)real_rows * slices * 128.0;
                                const double clk = (double)clock_khz * 1e3;
                                printf("{\"pitch\": %d, \"pattern\": \"%s\", \"zero_fill\": %.2f, \"method\": %d, \"W\": %d, \"ctas\": %d, \"S\": %d, "
                                       "\"ms\": %.4f, \"slot_B_per_clk_sm\": %.1f, \"real_B_per_clk_sm\": %.1f, \"real_TBps\": %.2f, "
                                       "\"cycles_per_item_sm\": %.0f}\n",
                                       pitch, pattern, zf, method, W, ctas, S, best, slot_bytes / (best * 1e-3 * clk * sm_count),
                                       real_bytes / (best * 1e-3 * clk * sm_count), real_bytes / (best * 1e-3) / 1e12,
                                       best * 1e-3 * clk * sm_count / p.n_items);
                                fflush(stdout);
                                CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
                            }
                        }
                    }
                }
                CK(cudaFree(d_idx));
            }
        }
        CK(cudaFree(X));
    }
    return 0;
}
