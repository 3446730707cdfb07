// Sparse convolution on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a only.
// U2_MATH_TF32: fp32 storage, kind::tf32 MMA, fp32 accumulation in TMEM.
//
// fwd / dgrad  (u2_conv_fwd_tc): output-stationary implicit GEMM.
//   CTA  = 128 destination rows x NT output channels (UMMA M=128, N=NT<=256, K=8/instr).
//   item = (kernel offset k with >=1 valid neighbour in the tile) x (32-channel slice of Cs).
//   warps 0-3 : A producers — gather the 128 neighbour rows of the slice with 16-byte
//               cp.async (zero-fill for missing neighbours) straight into the canonical
//               K-major no-swizzle UMMA layout; completion lands on the stage's mbarrier.
//               Afterwards the same warps are the epilogue (TMEM -> registers -> HBM).
//   warp 4    : B producer — one cp.async.bulk (TMA engine, UBLKCP) per item from the
//               pre-tiled weight blob (already in UMMA layout), complete_tx on the mbarrier.
//   warp 5    : TMEM allocator + single-thread tcgen05.mma issuer; tcgen05.commit frees the
//               stage / publishes the accumulator.
//   No gather buffer, no output read-modify-write, no atomics: every output row is written
//   once; (tile, offset) pairs without neighbours are skipped.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "u2_common.cuh"

namespace {

constexpr int TILE_M = 128;
constexpr int NUM_THREADS = 192;
constexpr int MAX_STAGES = 4;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}

// ---------------------------------------------------------------- async copies
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// experiment variants (U2_CPASYNC_MODE): 1 = .ca (allocate in L1), 2 = .cg + L2::128B prefetch, 3 = .cg + L2::256B
__device__ __forceinline__ void cp_async16_mode(uint32_t dst, const void *src, uint32_t src_bytes, int mode) {
    if (mode == 1) asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
    else if (mode == 2) asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
    else if (mode == 3) asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
    else asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16_ca(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
template <bool BF16>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if (BF16) umma_bf16(d_tmem, a_desc, b_desc, idesc, accumulate);
    else umma_tf32(d_tmem, a_desc, b_desc, idesc, accumulate);
}
// Warp-converged variants: ALL 32 lanes execute the statement with identical operands and elect.sync picks the issuing
// lane inside the same asm block.  With `if (lane == 0) tcgen05.mma ...` the compiler treats the descriptors as per-thread
// values and wraps every UTCHMMA / UTCBAR in a "waterfall" loop (ELECT + 4 R2UR.BROADCAST + BRA.U.ANY): measured ~75 cycles
// per MMA and ~300 per commit on the issuing thread, i.e. ~600 cycles per 4-MMA work item before any tensor work — the
// MMA thread, not the tensor pipe, paced the kernels (profiles/r2_conv_phase_clocks.md).
template <bool BF16>
__device__ __forceinline__ void umma_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if (BF16)
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "elect.sync _|q, 0xffffffff;\n\t"
            "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "elect.sync _|q, 0xffffffff;\n\t"
            "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t *bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleave") canonical layouts
// (cute/arch/mma_sm100_desc.hpp SmemDescriptor; version_=1 for Blackwell):
//   K-major : core matrix = 8 rows x 16 B contiguous; SBO = stride between 8-row groups (M/N),
//             LBO = stride between the two 16-byte K chunks of one MMA.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// Instruction descriptor (cute InstrDescriptor): F32 accumulate, TF32 x TF32, both K-major.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// BF16 x BF16 -> F32 (kind::f16), both K-major
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int A_ROW_GROUP = TILE_M * 16;      // 2048: bytes of one 16-byte K chunk over 128 rows
constexpr int A_LBO = A_ROW_GROUP + 16;       // +16: bank-conflict-free 16-byte gather writes

struct FwdParams {
    const uint8_t *X;   // fp32 (tf32 math) or bf16 rows, Cs channels each
    const uint8_t *Wt;  // pre-tiled weights, same element type
    const int *table;  // [K, ld] source row per (offset, tile row), -1 = none
    const int *perm;   // tile row -> destination row (mask-sorted tiles), or nullptr = identity
    float *Y;
    int64_t ld, n_dst;
    int Cs, Cd, K, NT, stages, tmem_cols;
    long long *dbg;  // optional per-CTA phase timestamps (U2_DEBUG_CONV_TIMING)
    int cp_mode;
    float *tile_stats;  // optional [tiles * 4][2][Cd]: per-warp column sums / sums of squares of Y (fused BatchNorm)
    const float *Yadd;  // optional [n_dst, Cd]: added to the result in the epilogue (gradient of a second consumer of the input)
    int64_t n_tiles;  // persistent kernel: 128-row tiles to walk
    int diag;  // U2_CONV_DIAG (timing diagnostics, results invalid): 1 = no weight loads, 2 = no gathers, 4 = gathers hit 128 hot rows
};

// ROWB = bytes of one gathered row per pipeline stage (128 or 64): 32/16 fp32 or 64/32 bf16 channels.
template <int ROWB, bool BF16, bool YADD>
__global__ void __launch_bounds__(NUM_THREADS) conv_fwd_tc_kernel(const FwdParams p) {
    constexpr int ES = BF16 ? 2 : 4;          // element size
    constexpr int CHUNKS = ROWB / 16;         // 16-byte chunks per row per stage
    constexpr int ROWS_PER_IT = TILE_M / CHUNKS;  // rows covered by one cp.async round of the 128 producers
    constexpr int A_BYTES = CHUNKS * A_LBO;
    extern __shared__ __align__(128) uint8_t smem[];
    const int NT = p.NT;
    const int B_LBO = NT * 16;
    const int B_BYTES = CHUNKS * B_LBO;
    const int stage_bytes = A_BYTES + B_BYTES;
    uint8_t *s_stage = smem;
    int *s_tab = reinterpret_cast<int *>(smem + (size_t)p.stages * stage_bytes);
    uint64_t *s_full = reinterpret_cast<uint64_t *>(s_tab + p.K * TILE_M);
    uint64_t *s_empty = s_full + MAX_STAGES;
    uint64_t *s_accum = s_empty + MAX_STAGES;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_accum + 1);
    uint32_t *s_mask = s_tmem + 1;
    uint32_t *s_pm = s_mask + 1;                                   // [K][4]  presence mask of every 32-row quarter
    uint32_t *s_dirty = s_pm + p.K * 4;                            // [MAX_STAGES][4] rows of a stage that hold data
    int *s_cnt = reinterpret_cast<int *>(s_dirty + MAX_STAGES * 4);  // [K]     rows present at this offset
    uint8_t *s_list = reinterpret_cast<uint8_t *>(s_cnt + p.K);    // [K][128] their tile rows, ascending

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t row0 = (int64_t)blockIdx.x * TILE_M;
    const int nt = blockIdx.y;
    long long *dbg = p.dbg ? p.dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16 : nullptr;
    if (dbg && tid == 0) dbg[0] = clock64();
    const int n_cc = p.Cs * ES / ROWB;
    const int n_nt = p.Cd / NT;

    if (tid < MAX_STAGES * 4) s_dirty[tid] = 0xFFFFFFFFu;  // a stage starts with stale shared memory in every row
    if (tid == 0) {
        *s_mask = 0;
        for (int s = 0; s < p.stages; s++) {
            mbar_init(s_full + s, TILE_M + 1);
            mbar_init(s_empty + s, 1);
        }
        mbar_init(s_accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        tmem_alloc(s_tmem, (uint32_t)p.tmem_cols);
        tmem_relinquish();
    }
    __syncthreads();
    // neighbour table of this tile -> smem, and the set of offsets that have any neighbour.
    // All loads of a thread are issued before the first use (one L2 round trip, not K/6 of them).
    {
        constexpr int MAXJ = 6;  // ceil(32 / 6 warps)
        int v[MAXJ][TILE_M / 32];
#pragma unroll
        for (int j = 0; j < MAXJ; j++) {
            const int k = warp + j * (NUM_THREADS / 32);
#pragma unroll
            for (int i = 0; i < TILE_M / 32; i++)
                v[j][i] = k < p.K ? __ldg(p.table + (int64_t)k * p.ld + row0 + lane + 32 * i) : -1;
        }
#pragma unroll
        for (int j = 0; j < MAXJ; j++) {
            const int k = warp + j * (NUM_THREADS / 32);
            if (k < p.K) {
                // compact list of the rows that have a neighbour at this offset: the producers fetch those and only those
                int cnt = 0;
#pragma unroll
                for (int i = 0; i < TILE_M / 32; i++) {
                    s_tab[k * TILE_M + lane + 32 * i] = v[j][i];
                    const uint32_t b = __ballot_sync(0xffffffffu, v[j][i] >= 0);
                    if (v[j][i] >= 0) s_list[k * TILE_M + cnt + __popc(b & ((1u << lane) - 1u))] = (uint8_t)(lane + 32 * i);
                    if (lane == 0) s_pm[k * 4 + i] = b;
                    cnt += __popc(b);
                }
                if (lane == 0) {
                    s_cnt[k] = cnt;
                    if (cnt) atomicOr(s_mask, 1u << k);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t kmask = *s_mask;
    const uint32_t tmem_base = *s_tmem;
    const int n_items = __popc(kmask) * n_cc;
    if (dbg && tid == 0) { dbg[1] = clock64(); dbg[6] = n_items; }

    if (warp < 4) {
        // ============================ A producers ============================
        // Per (offset, channel slice) item only the PRESENT rows are fetched: present row j of the offset's compact list goes
        // to lane group j % ROWS_PER_IT (CHUNKS lanes, one 16-byte piece each), so an item costs ceil(m / (32 / CHUNKS))
        // warp-level LDGSTS instead of TILE_M / (32 / CHUNKS).  That matters because the gather is bound by the LDGSTS
        // instruction rate of the SM (~20-25 cycles per warp instruction whatever it fetches: a zero-fill copy costs as
        // much as a real one; scripts/microbench/gather_bench.cu, profiles/r2_gather_microbench.md) and ~47 % of the
        // row slots of the mask-sorted tiles are empty.  Rows that are absent now but still hold data of the stage's
        // previous item are cleared with plain shared-memory stores (warp w owns rows 32 w .. 32 w + 31).
        const int chunk = tid % CHUNKS, grp = tid / CHUNKS;
        int s = 0;
        uint32_t ph = 0;
        long long dbg_wait = 0, dbg_ph[3] = {0, 0, 0};
        for (uint32_t m = kmask; m; m &= m - 1) {
            const int k = __ffs(m) - 1;
            const int cnt = s_cnt[k];
            const uint32_t pm = s_pm[k * 4 + warp];
            // per-offset setup: global byte offset and shared-memory offset of this thread's piece in each of its rounds
            uint32_t off[CHUNKS], dsto[CHUNKS];
#pragma unroll
            for (int i = 0; i < CHUNKS; i++) {
                const int j = grp + i * ROWS_PER_IT;
                if (j < cnt) {
                    const int slot = s_list[k * TILE_M + j];
                    const int src = (p.diag & 4) ? slot : s_tab[k * TILE_M + slot];
                    off[i] = (uint32_t)src * (uint32_t)(p.Cs * ES) + (uint32_t)(chunk * 16);
                    dsto[i] = (uint32_t)(chunk * A_LBO + slot * 16);
                } else {
                    off[i] = 0xFFFFFFFFu;
                    dsto[i] = 0u;
                }
            }
            const uint8_t *xc = p.X;
            for (int cc = 0; cc < n_cc; cc++, xc += ROWB) {
                if (dbg && tid == 0) {
                    const long long t0 = clock64();
                    mbar_wait(s_empty + s, ph ^ 1u);
                    dbg_wait += clock64() - t0;
                } else {
                    mbar_wait(s_empty + s, ph ^ 1u);
                }
                const long long tp0 = (dbg && tid == 0) ? clock64() : 0;
                const uint32_t a_base = smem_u32(s_stage + (size_t)s * stage_bytes);
                const uint32_t stale = s_dirty[s * 4 + warp] & ~pm;
                if (stale) {  // warp-uniform
                    if ((stale >> lane) & 1u) {
#pragma unroll
                        for (int c = 0; c < CHUNKS; c++)
                            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a_base + c * A_LBO + (warp * 32 + lane) * 16), "r"(0)
                                         : "memory");
                    }
                    proxy_fence_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
                }
                __syncwarp();
                if (lane == 0) s_dirty[s * 4 + warp] = pm;
                const long long tp1 = (dbg && tid == 0) ? clock64() : 0;
#pragma unroll
                for (int i = 0; i < CHUNKS; i++) {
                    if (i * ROWS_PER_IT + (warp * 32) / CHUNKS < cnt) {  // warp-uniform: this round has rows for this warp
                        if (off[i] != 0xFFFFFFFFu && !(p.diag & 2)) cp_async16_ca(a_base + dsto[i], xc + off[i]);
                    }
                }
                const long long tp2 = (dbg && tid == 0) ? clock64() : 0;
                cp_async_mbar_arrive_noinc(s_full + s);
                if (dbg && tid == 0) { const long long tp3 = clock64(); dbg_ph[0] += tp1 - tp0; dbg_ph[1] += tp2 - tp1; dbg_ph[2] += tp3 - tp2; }
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
        }
        // ============================ epilogue ============================
        // TMEM -> registers (lane = row) -> shared memory (the idle pipeline stages, XOR-swizzled 16-byte chunks) ->
        // coalesced global stores: 8 lanes write the 128 contiguous bytes of one row, 4 rows per instruction; the fused
        // BatchNorm column sums are read from the same staging tile.  (The fp32 output write is ~25 % of the kernel on the
        // stride-1 layers, U2_CONV_DIAG=16, and it is the HBM traffic itself: storing straight from the TMEM registers,
        // 32 rows per instruction, takes the same time.)
        if (dbg && tid == 0) { dbg[2] = clock64(); dbg[8] = dbg_wait; dbg[10] = dbg_ph[0]; dbg[11] = dbg_ph[1]; dbg[12] = dbg_ph[2]; }
        const int64_t trow = row0 + warp * 32 + lane;
        int64_t row = trow;
        if (p.perm) row = __ldg(p.perm + trow);
        const int my_dst = (row >= 0 && row < p.n_dst && !(p.diag & 16)) ? (int)row : -1;  // diag 16: no output stores
        if (n_items > 0) {
            mbar_wait(s_accum, 0);
            tc_fence_after();
            if (dbg && tid == 0) dbg[3] = clock64();
            const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
            // rows without a destination (padding of the sorted table) have no neighbours: their accumulators are 0
            float *ts = (p.tile_stats && !(p.diag & 32)) ? p.tile_stats + ((size_t)blockIdx.x * 4 + warp) * 2 * p.Cd + nt * NT : nullptr;
            uint8_t *stg = smem + warp * 4096;            // this warp's 32 rows x 128 B staging tile (inside stage 0)
            const uint32_t stg_u32 = smem_u32(stg);
            const int rsub = lane >> 3, cj = lane & 7;    // store phase: row within a group of 4, 16-byte chunk of the row
            for (int c0 = 0; c0 < NT; c0 += 32) {
                const int ncol = min(32, NT - c0);        // 32, or 16 for the last chunk when NT % 32 == 16
                // the addend rows of this chunk: all 8 loads of a lane in flight before the TMEM read and the staging
                float4 ya[YADD ? 8 : 1];
                if (YADD) {
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int dst = __shfl_sync(0xFFFFFFFFu, my_dst, 4 * i + rsub);
                        ya[i] = (dst >= 0 && cj * 4 < ncol) ? __ldg(reinterpret_cast<const float4 *>(p.Yadd + (int64_t)dst * p.Cd + nt * NT + c0 + cj * 4))
                                                             : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                uint32_t v[32];
                {
                    uint32_t lo[16], hi[16];
                    tmem_ld16(t_row + (uint32_t)c0, lo);
                    if (ncol == 32) tmem_ld16(t_row + (uint32_t)c0 + 16, hi);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; j++) { v[j] = lo[j]; v[16 + j] = ncol == 32 ? hi[j] : 0u; }
                }
                __syncwarp();  // the previous chunk's readers are done with the staging tile
#pragma unroll
                for (int j = 0; j < 8; j++)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_u32 + lane * 128 + ((j ^ (lane & 7)) << 4)),
                                 "r"(v[4 * j]), "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                                 : "memory");
                __syncwarp();
                if (ts && ncol == 32) {
                    // column `lane` over the warp's 32 rows: one conflict-free 4-byte read per row
                    float sum = 0.f, sq = 0.f;
#pragma unroll 8
                    for (int r = 0; r < 32; r++) {
                        const float x = *reinterpret_cast<const float *>(stg + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + ((lane & 3) << 2));
                        sum += x;
                        sq = fmaf(x, x, sq);
                    }
                    ts[c0 + lane] = sum;
                    ts[p.Cd + c0 + lane] = sq;
                }
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int r = 4 * i + rsub;
                    const int dst = __shfl_sync(0xFFFFFFFFu, my_dst, r);  // all lanes take part, whatever ncol is
                    if (dst >= 0 && cj * 4 < ncol) {
                        uint4 q;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                                     : "r"(stg_u32 + r * 128 + ((cj ^ (r & 7)) << 4)));
                        const int64_t o = (int64_t)dst * p.Cd + nt * NT + c0 + cj * 4;
                        if (YADD) {
                            q.x = __float_as_uint(__uint_as_float(q.x) + ya[i].x);
                            q.y = __float_as_uint(__uint_as_float(q.y) + ya[i].y);
                            q.z = __float_as_uint(__uint_as_float(q.z) + ya[i].z);
                            q.w = __float_as_uint(__uint_as_float(q.w) + ya[i].w);
                        }
                        *reinterpret_cast<uint4 *>(p.Y + o) = q;
                    }
                }
            }
        } else {
            if (my_dst >= 0) {
                float *yrow = p.Y + (int64_t)my_dst * p.Cd + nt * NT;
                const float *arow = YADD ? p.Yadd + (int64_t)my_dst * p.Cd + nt * NT : nullptr;
                for (int c0 = 0; c0 < NT; c0 += 4)
                    *reinterpret_cast<float4 *>(yrow + c0) = arow ? __ldg(reinterpret_cast<const float4 *>(arow + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (p.tile_stats) {
                float *ts = p.tile_stats + ((size_t)blockIdx.x * 4 + warp) * 2 * p.Cd + nt * NT;
                for (int c = lane; c < NT; c += 32) { ts[c] = 0.f; ts[p.Cd + c] = 0.f; }
            }
        }
    } else if (warp == 4) {
        // ============================ B producer ============================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (uint32_t m = kmask; m; m &= m - 1) {
                const int k = __ffs(m) - 1;
                for (int cc = 0; cc < n_cc; cc++) {
                    mbar_wait(s_empty + s, ph ^ 1u);
                    const uint32_t b_base = smem_u32(s_stage + (size_t)s * stage_bytes + A_BYTES);
                    const uint8_t *blob = p.Wt + ((size_t)(k * n_cc + cc) * n_nt + nt) * (size_t)(ROWB * NT);
                    if (p.diag & 1) {
                        mbar_arrive(s_full + s);
                    } else {
                        mbar_arrive_expect_tx(s_full + s, (uint32_t)B_BYTES);
                        bulk_g2s(b_base, blob, (uint32_t)B_BYTES, s_full + s);
                    }
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else {
        // ============================ MMA issuer ============================
        const uint32_t idesc = BF16 ? make_idesc_bf16(TILE_M, NT) : make_idesc_tf32(TILE_M, NT);
        int s = 0;
        uint32_t ph = 0;
        long long dbg_wait_full = 0, dbg_mma[3] = {0, 0, 0};
        for (int it = 0; it < n_items; it++) {
            if (dbg && lane == 0) {
                const long long t0 = clock64();
                mbar_wait(s_full + s, ph);
                dbg_wait_full += clock64() - t0;
            } else {
                mbar_wait(s_full + s, ph);
            }
            const long long tm0 = (dbg && lane == 0) ? clock64() : 0;
            tc_fence_after();
            proxy_fence_async();
            if (dbg && lane == 0) dbg_mma[0] += clock64() - tm0;
            {   // all 32 lanes, converged: uniform descriptors, elect.sync inside the asm (see umma_elect)
                const long long tm1 = dbg ? clock64() : 0;
                const uint32_t a_base = smem_u32(s_stage + (size_t)s * stage_bytes);
                const uint32_t b_base = a_base + A_BYTES;
#pragma unroll
                for (int kk = 0; kk < ROWB / 32; kk++) {  // one MMA consumes 32 bytes of K: 8 tf32 or 16 bf16
                    const uint64_t ad = make_smem_desc(a_base + kk * 2 * A_LBO, A_LBO, 128);
                    const uint64_t bd = make_smem_desc(b_base + kk * 2 * B_LBO, B_LBO, 128);
                    umma_elect<BF16>(tmem_base, ad, bd, idesc, (it > 0 || kk > 0) ? 1u : 0u);
                }
                const long long tm2 = dbg ? clock64() : 0;
                umma_commit_elect(s_empty + s);
                if (it == n_items - 1) umma_commit_elect(s_accum);
                if (dbg && lane == 0) { dbg_mma[1] += tm2 - tm1; dbg_mma[2] += clock64() - tm2; }
            }
            if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
        if (dbg && lane == 0) { dbg[9] = dbg_wait_full; dbg[13] = dbg_mma[0]; dbg[14] = dbg_mma[1]; dbg[15] = dbg_mma[2]; }
    }
    if (dbg && tid == 0) dbg[4] = clock64();
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    if (dbg && tid == 160) { dbg[5] = clock64(); unsigned sm; asm("mov.u32 %0, %%smid;" : "=r"(sm)); dbg[7] = sm; }
}

// ------------------------------------------------------------------ fwd / dgrad, persistent warp-specialised kernel
// Same math and shared-memory operand layouts as conv_fwd_tc_kernel above, different schedule.  What the phase clocks
// of that kernel showed (profiles/r2_conv_phase_clocks.md): with 2-3 co-resident CTAs of 2 stages each, a work item
// costs ~1400 cycles per CTA even with every global load switched off — all four producer warps walk through every
// item (wait, clear stale rows, proxy fence, issue, arrive: ~900 dependent cycles each), the single MMA thread spends
// ~650 cycles per item in fences / issue / commit, and with two stages neither side has slack.  Here:
//   * ONE CTA per SM, persistent over (row tile, channel tile) work units, all of shared memory in one deep ring
//     (3-8 stages) instead of 2-3 shallow ones;
//   * a work item belongs to ONE producer warp (item g -> stage g % S -> warp g % S): the per-item fixed cost is paid by
//     one warp while seven others are busy with the neighbouring items;
//   * the accumulator is double-buffered in TMEM (2 x NT columns): four dedicated warps drain tile t while the
//     producers and the MMA thread are already in tile t + 1; the neighbour table / row lists of tile t + 1 are built by
//     two dedicated warps into the second table buffer meanwhile.
constexpr int PS_MAX_STAGES = 8;
constexpr int PS_PRODUCERS = 8;       // warps 8..15
constexpr int PS_THREADS = 512;       // warps 0-3 epilogue, 4 MMA, 5 weights, 6-7 tables, 8-15 gather

struct PsTab {   // byte offsets inside one table buffer
    int tab, list, pm, cnt, kmask, bytes;
};
__host__ __device__ inline PsTab ps_tab_layout(int K) {
    PsTab t;
    t.tab = 0;
    t.list = t.tab + K * TILE_M * 4;
    t.pm = t.list + K * TILE_M;
    t.cnt = t.pm + K * 16;
    t.kmask = t.cnt + K * 4;
    t.bytes = (t.kmask + 16 + 127) / 128 * 128;
    return t;
}

template <int ROWB, bool BF16, bool YADD>
__global__ void __launch_bounds__(PS_THREADS, 1) conv_fwd_ps_kernel(const FwdParams p) {
    constexpr int ES = BF16 ? 2 : 4;
    constexpr int CHUNKS = ROWB / 16;             // 16-byte chunks per row per stage
    constexpr int RPI = 32 / CHUNKS;              // rows per warp-level LDGSTS
    constexpr int A_BYTES = CHUNKS * A_LBO;
    extern __shared__ __align__(128) uint8_t smem[];
    const int NT = p.NT, S = p.stages;
    const int B_LBO = NT * 16;
    const int B_BYTES = CHUNKS * B_LBO;
    const int stage_bytes = A_BYTES + B_BYTES;
    const PsTab TL = ps_tab_layout(p.K);
    uint8_t *s_stage = smem;
    uint8_t *s_stg = smem + (size_t)S * stage_bytes;                 // 4 x 4 KB epilogue staging
    uint8_t *s_tabs = s_stg + 4 * 4096;                              // 2 table buffers
    uint64_t *s_full = reinterpret_cast<uint64_t *>(s_tabs + 2 * TL.bytes);
    uint64_t *s_empty = s_full + PS_MAX_STAGES;
    uint64_t *s_tab_full = s_empty + PS_MAX_STAGES;                  // [2]
    uint64_t *s_tab_empty = s_tab_full + 2;                          // [2]
    uint64_t *s_acc_full = s_tab_empty + 2;                          // [2]
    uint64_t *s_acc_empty = s_acc_full + 2;                          // [2]
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_acc_empty + 2);
    int *s_acc_items = reinterpret_cast<int *>(s_tmem + 1);          // [2] items accumulated into buffer a (0: nothing)
    uint32_t *s_dirty = reinterpret_cast<uint32_t *>(s_acc_items + 2);  // [PS_MAX_STAGES][4] rows of a stage that hold data

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_cc = p.Cs * ES / ROWB;
    const int n_nt = p.Cd / NT;
    const int64_t n_work = p.n_tiles * n_nt;
    // producer warp w owns stage w: the items of one stage are filled by one warp in program order, so a parity wait on the
    // stage's barriers can never be satisfied by a phase two uses back (a warp running ahead on a shared stage could)
    const int n_prod = S < PS_PRODUCERS ? S : PS_PRODUCERS;

    if (tid < PS_MAX_STAGES * 4) s_dirty[tid] = 0xFFFFFFFFu;  // a stage starts with stale shared memory in every row
    if (tid == 0) {
        for (int s = 0; s < S; s++) {
            mbar_init(s_full + s, 32 + 2);    // 32 cp.async completions of the owning warp + its release arrive + the weight thread
            mbar_init(s_empty + s, 1);        // tcgen05.commit
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(s_tab_full + b, 2);                    // the two table warps
            mbar_init(s_tab_empty + b, n_prod + 2);          // active producer warps + weight thread + MMA warp
            mbar_init(s_acc_full + b, 1);                    // tcgen05.commit (or a plain arrive for an empty tile)
            mbar_init(s_acc_empty + b, 4);                   // the four epilogue warps
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        tmem_alloc(s_tmem, (uint32_t)p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    // optional phase clocks (U2_DEBUG_CONV_TIMING=2): 32 slots per CTA
    //  0 start, 1 end | MMA thread: 2 items, 3 wait full, 4 fences, 5 issue, 6 commit, 7 wait tab/acc
    //  producer warp 0: 8 items, 9 wait empty, 10 zero + fence, 11 gather issue, 12 arrive, 13 wait table
    //  epilogue warp 0: 14 wait acc_full, 15 work | table warp 0: 16 wait, 17 work | 18 tiles of this CTA
    long long *dbg = p.dbg ? p.dbg + (size_t)blockIdx.x * 32 : nullptr;
    if (dbg && tid == 0) dbg[0] = clock64();
#define DBG_T() (dbg ? clock64() : 0LL)

    if (warp >= 8) {
        // ============================ A producers: one warp per work item ============================
        const int pw = warp - 8;
        const int chunk = lane % CHUNKS, grp = lane / CHUNKS;
        uint32_t g = 0;  // running item number of this CTA (same sequence in every role)
        int i = 0;
        for (int64_t w = pw < n_prod ? (int64_t)blockIdx.x : n_work; w < n_work; w += gridDim.x, i++) {
            const int b = i & 1;
            const uint8_t *tb = s_tabs + b * TL.bytes;
            const long long tq0 = DBG_T();
            mbar_wait(s_tab_full + b, (uint32_t)(i >> 1) & 1u);
            if (dbg && warp == 8 && lane == 0) dbg[13] += clock64() - tq0;
            const int *s_tab = reinterpret_cast<const int *>(tb + TL.tab);
            const uint8_t *s_list = tb + TL.list;
            const uint32_t *s_pm = reinterpret_cast<const uint32_t *>(tb + TL.pm);
            const int *s_cnt = reinterpret_cast<const int *>(tb + TL.cnt);
            const uint32_t *km = reinterpret_cast<const uint32_t *>(tb + TL.kmask);
            const uint32_t kmask = km[0] | km[1];
            for (uint32_t m = kmask; m; m &= m - 1) {
                const int k = __ffs(m) - 1;
                for (int cc = 0; cc < n_cc; cc++, g++) {
                    const int s = (int)(g % (uint32_t)S);
                    if (s != pw) continue;
                    const uint32_t ph = (g / (uint32_t)S) & 1u;
                    const long long tp0 = DBG_T();
                    mbar_wait(s_empty + s, ph ^ 1u);
                    const long long tp1 = DBG_T();
                    const uint32_t a_base = smem_u32(s_stage + (size_t)s * stage_bytes);
                    // rows that still hold the previous item's data but have no neighbour at this offset -> zero
                    uint32_t any_stale = 0;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const uint32_t pmq = s_pm[k * 4 + q];
                        const uint32_t stale = s_dirty[s * 4 + q] & ~pmq;
                        any_stale |= stale;
                        if ((stale >> lane) & 1u) {
#pragma unroll
                            for (int c = 0; c < CHUNKS; c++)
                                asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a_base + c * A_LBO + (q * 32 + lane) * 16), "r"(0)
                                             : "memory");
                        }
                    }
                    if (any_stale) proxy_fence_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
                    __syncwarp();
                    if (lane < 4) s_dirty[s * 4 + lane] = s_pm[k * 4 + lane];
                    const long long tp2 = DBG_T();
                    // gather the present rows, RPI per instruction (lane group `grp` takes present row r * RPI + grp)
                    const int cnt = s_cnt[k];
                    const uint8_t *xc = p.X + (size_t)cc * ROWB + chunk * 16;
                    const uint32_t dst0 = a_base + chunk * A_LBO;
                    if (!(p.diag & 2)) {
                        for (int j0 = 0; j0 < cnt; j0 += 4 * RPI) {
                            int slot[4], src[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const int j = j0 + u * RPI + grp;
                                slot[u] = j < cnt ? (int)s_list[k * TILE_M + j] : -1;
                            }
#pragma unroll
                            for (int u = 0; u < 4; u++) src[u] = slot[u] >= 0 ? s_tab[k * TILE_M + slot[u]] : 0;
#pragma unroll
                            for (int u = 0; u < 4; u++)
                                if (slot[u] >= 0) cp_async16_ca(dst0 + slot[u] * 16, xc + (size_t)(uint32_t)src[u] * (uint32_t)(p.Cs * ES));
                        }
                    }
                    const long long tp3 = DBG_T();
                    cp_async_mbar_arrive_noinc(s_full + s);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_full + s);  // release: the s_dirty update and the zeroing stores of this warp
                    if (dbg && warp == 8 && lane == 0) {
                        dbg[8] += 1; dbg[9] += tp1 - tp0; dbg[10] += tp2 - tp1; dbg[11] += tp3 - tp2; dbg[12] += clock64() - tp3;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(s_tab_empty + b);
        }
    } else if (warp == 6 || warp == 7) {
        // ============================ table warps: neighbour table of the next tile -> shared memory ============================
        const int tw = warp - 6;
        int i = 0;
        for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x, i++) {
            const int b = i & 1;
            uint8_t *tb = s_tabs + b * TL.bytes;
            const long long tt0 = DBG_T();
            if (i >= 2) mbar_wait(s_tab_empty + b, (uint32_t)((i >> 1) - 1) & 1u);
            const long long tt1 = DBG_T();
            int *s_tab = reinterpret_cast<int *>(tb + TL.tab);
            uint8_t *s_list = tb + TL.list;
            uint32_t *s_pm = reinterpret_cast<uint32_t *>(tb + TL.pm);
            int *s_cnt = reinterpret_cast<int *>(tb + TL.cnt);
            uint32_t *km = reinterpret_cast<uint32_t *>(tb + TL.kmask);
            const int64_t row0 = (w / n_nt) * TILE_M;
            uint32_t kmask = 0;
            // offsets tw, tw + 2, ...; 7 offsets (28 loads per lane) in flight at a time
            for (int k0 = tw; k0 < p.K; k0 += 14) {
                int v[7][4];
#pragma unroll
                for (int j = 0; j < 7; j++) {
                    const int k = k0 + 2 * j;
#pragma unroll
                    for (int q = 0; q < 4; q++) v[j][q] = k < p.K ? __ldg(p.table + (int64_t)k * p.ld + row0 + lane + 32 * q) : -1;
                }
#pragma unroll
                for (int j = 0; j < 7; j++) {
                    const int k = k0 + 2 * j;
                    if (k < p.K) {
                        int cnt = 0;
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            s_tab[k * TILE_M + lane + 32 * q] = v[j][q];
                            const uint32_t bal = __ballot_sync(0xffffffffu, v[j][q] >= 0);
                            if (v[j][q] >= 0) s_list[k * TILE_M + cnt + __popc(bal & ((1u << lane) - 1u))] = (uint8_t)(lane + 32 * q);
                            if (lane == 0) s_pm[k * 4 + q] = bal;
                            cnt += __popc(bal);
                        }
                        if (lane == 0) s_cnt[k] = cnt;
                        if (cnt) kmask |= 1u << k;
                    }
                }
            }
            if (lane == 0) km[tw] = kmask;
            __syncwarp();
            if (lane == 0) mbar_arrive(s_tab_full + b);
            if (dbg && tw == 0 && lane == 0) { dbg[16] += tt1 - tt0; dbg[17] += clock64() - tt1; dbg[18] += 1; }
        }
    } else if (warp == 5) {
        // ============================ B producer: pre-tiled weight blobs, one bulk copy per item ============================
        if (lane == 0) {
            uint32_t g = 0;
            int i = 0;
            for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x, i++) {
                const int b = i & 1;
                const int nt = (int)(w % n_nt);
                mbar_wait(s_tab_full + b, (uint32_t)(i >> 1) & 1u);
                const uint32_t *km = reinterpret_cast<const uint32_t *>(s_tabs + b * TL.bytes + TL.kmask);
                const uint32_t kmask = km[0] | km[1];
                for (uint32_t m = kmask; m; m &= m - 1) {
                    const int k = __ffs(m) - 1;
                    for (int cc = 0; cc < n_cc; cc++, g++) {
                        const int s = (int)(g % (uint32_t)S);
                        const uint32_t ph = (g / (uint32_t)S) & 1u;
                        mbar_wait(s_empty + s, ph ^ 1u);
                        const uint32_t b_base = smem_u32(s_stage + (size_t)s * stage_bytes + A_BYTES);
                        const uint8_t *blob = p.Wt + ((size_t)(k * n_cc + cc) * n_nt + nt) * (size_t)(ROWB * NT);
                        if (p.diag & 1) {
                            mbar_arrive(s_full + s);
                        } else {
                            mbar_arrive_expect_tx(s_full + s, (uint32_t)B_BYTES);
                            bulk_g2s(b_base, blob, (uint32_t)B_BYTES, s_full + s);
                        }
                    }
                }
                mbar_arrive(s_tab_empty + b);
            }
        }
    } else if (warp == 4) {
        // ============================ MMA issuer ============================
        const uint32_t idesc = BF16 ? make_idesc_bf16(TILE_M, NT) : make_idesc_tf32(TILE_M, NT);
        uint32_t g = 0;
        int i = 0;
        for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x, i++) {
            const int b = i & 1, a = i & 1;
            const long long tw0 = DBG_T();
            mbar_wait(s_tab_full + b, (uint32_t)(i >> 1) & 1u);
            const uint32_t *km = reinterpret_cast<const uint32_t *>(s_tabs + b * TL.bytes + TL.kmask);
            const int n_items = __popc(km[0] | km[1]) * n_cc;
            __syncwarp();
            if (lane == 0) mbar_arrive(s_tab_empty + b);  // only the item count is needed from the table buffer
            if (i >= 2) mbar_wait(s_acc_empty + a, (uint32_t)((i >> 1) - 1) & 1u);  // the epilogue drained this accumulator
            tc_fence_after();
            if (dbg && lane == 0) dbg[7] += clock64() - tw0;
            const uint32_t d_tmem = tmem_base + (uint32_t)(a * NT);
            for (int it = 0; it < n_items; it++, g++) {
                const int s = (int)(g % (uint32_t)S);
                const uint32_t ph = (g / (uint32_t)S) & 1u;
                const long long tm0 = DBG_T();
                mbar_wait(s_full + s, ph);
                const long long tm1 = DBG_T();
                tc_fence_after();
                proxy_fence_async();
                const long long tm2 = DBG_T();
                {
                    const uint32_t a_base = smem_u32(s_stage + (size_t)s * stage_bytes);
                    const uint32_t b_base = a_base + A_BYTES;
#pragma unroll
                    for (int kk = 0; kk < ROWB / 32; kk++) {  // one MMA consumes 32 bytes of K: 8 tf32 or 16 bf16
                        const uint64_t ad = make_smem_desc(a_base + kk * 2 * A_LBO, A_LBO, 128);
                        const uint64_t bd = make_smem_desc(b_base + kk * 2 * B_LBO, B_LBO, 128);
                        umma_elect<BF16>(d_tmem, ad, bd, idesc, (it > 0 || kk > 0) ? 1u : 0u);
                    }
                    const long long tm3 = DBG_T();
                    umma_commit_elect(s_empty + s);
                    if (dbg && lane == 0) { dbg[2] += 1; dbg[3] += tm1 - tm0; dbg[4] += tm2 - tm1; dbg[5] += tm3 - tm2; dbg[6] += clock64() - tm3; }
                }
            }
            if (lane == 0) s_acc_items[a] = n_items;
            __syncwarp();
            if (n_items > 0) umma_commit_elect(s_acc_full + a);
            else if (lane == 0) mbar_arrive(s_acc_full + a);
            __syncwarp();
        }
    } else {
        // ============================ epilogue warps 0-3 ============================
        // TMEM -> registers (lane = row) -> this warp's staging tile (XOR-swizzled 16-byte chunks) -> coalesced global
        // stores (8 lanes write the 128 contiguous bytes of one row); the fused BatchNorm column sums come from the same
        // staging tile.
        uint8_t *stg = s_stg + warp * 4096;
        const uint32_t stg_u32 = smem_u32(stg);
        const int rsub = lane >> 3, cj = lane & 7;
        int i = 0;
        for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x, i++) {
            const int a = i & 1;
            const int64_t rt = w / n_nt;
            const int nt = (int)(w % n_nt);
            const int64_t trow = rt * TILE_M + warp * 32 + lane;
            int64_t row = trow;
            if (p.perm) row = __ldg(p.perm + trow);
            const int my_dst = (row >= 0 && row < p.n_dst && !(p.diag & 16)) ? (int)row : -1;
            const long long te0 = DBG_T();
            mbar_wait(s_acc_full + a, (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            const long long te1 = DBG_T();
            const int n_items = s_acc_items[a];
            float *ts = p.tile_stats ? p.tile_stats + ((size_t)rt * 4 + warp) * 2 * p.Cd + nt * NT : nullptr;
            if (n_items > 0) {
                const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * NT);
                for (int c0 = 0; c0 < NT; c0 += 32) {
                    const int ncol = min(32, NT - c0);
                    float4 ya[YADD ? 8 : 1];
                    if (YADD) {
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            const int dst = __shfl_sync(0xFFFFFFFFu, my_dst, 4 * u + rsub);
                            ya[u] = (dst >= 0 && cj * 4 < ncol) ? __ldg(reinterpret_cast<const float4 *>(p.Yadd + (int64_t)dst * p.Cd + nt * NT + c0 + cj * 4))
                                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                    uint32_t v[32];
                    {
                        uint32_t lo[16], hi[16];
                        tmem_ld16(t_row + (uint32_t)c0, lo);
                        if (ncol == 32) tmem_ld16(t_row + (uint32_t)c0 + 16, hi);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; j++) { v[j] = lo[j]; v[16 + j] = ncol == 32 ? hi[j] : 0u; }
                    }
                    if (c0 + 32 >= NT) {  // last read of this accumulator: hand it back before the stores
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(s_acc_empty + a);
                    }
                    __syncwarp();  // the previous chunk's readers are done with the staging tile
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_u32 + lane * 128 + ((j ^ (lane & 7)) << 4)),
                                     "r"(v[4 * j]), "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                                     : "memory");
                    __syncwarp();
                    if (ts && ncol == 32) {
                        float sum = 0.f, sq = 0.f;
#pragma unroll 8
                        for (int r = 0; r < 32; r++) {
                            const float x = *reinterpret_cast<const float *>(stg + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + ((lane & 3) << 2));
                            sum += x;
                            sq = fmaf(x, x, sq);
                        }
                        ts[c0 + lane] = sum;
                        ts[p.Cd + c0 + lane] = sq;
                    }
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int r = 4 * u + rsub;
                        const int dst = __shfl_sync(0xFFFFFFFFu, my_dst, r);
                        if (dst >= 0 && cj * 4 < ncol) {
                            uint4 q;
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                                         : "r"(stg_u32 + r * 128 + ((cj ^ (r & 7)) << 4)));
                            const int64_t o = (int64_t)dst * p.Cd + nt * NT + c0 + cj * 4;
                            if (YADD) {
                                q.x = __float_as_uint(__uint_as_float(q.x) + ya[u].x);
                                q.y = __float_as_uint(__uint_as_float(q.y) + ya[u].y);
                                q.z = __float_as_uint(__uint_as_float(q.z) + ya[u].z);
                                q.w = __float_as_uint(__uint_as_float(q.w) + ya[u].w);
                            }
                            *reinterpret_cast<uint4 *>(p.Y + o) = q;
                        }
                    }
                }
            } else {
                __syncwarp();
                if (lane == 0) mbar_arrive(s_acc_empty + a);
                if (my_dst >= 0) {
                    float *yrow = p.Y + (int64_t)my_dst * p.Cd + nt * NT;
                    const float *arow = YADD ? p.Yadd + (int64_t)my_dst * p.Cd + nt * NT : nullptr;
                    for (int c0 = 0; c0 < NT; c0 += 4)
                        *reinterpret_cast<float4 *>(yrow + c0) = arow ? __ldg(reinterpret_cast<const float4 *>(arow + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (ts)
                    for (int c = lane; c < NT; c += 32) { ts[c] = 0.f; ts[p.Cd + c] = 0.f; }
            }
            if (dbg && tid == 0) { dbg[14] += te1 - te0; dbg[15] += clock64() - te1; }
        }
    }
#undef DBG_T
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    if (dbg && tid == 0) dbg[1] = clock64();
}

template <bool WT>
__global__ void __launch_bounds__(256) pretile_weights_kernel(const float *__restrict__ W, float4 *__restrict__ out, int K,
                                                              int Cs, int Cd, int KC, int NT) {
    const int64_t total = (int64_t)K * Cs * Cd / 4;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int chunks = KC / 4, n_cc = Cs / KC, n_nt = Cd / NT;
    int64_t r = t;
    const int n = (int)(r % NT); r /= NT;
    const int chunk = (int)(r % chunks); r /= chunks;
    const int nti = (int)(r % n_nt); r /= n_nt;
    const int cc = (int)(r % n_cc); r /= n_cc;
    const int k = (int)r;
    const int cs = cc * KC + chunk * 4, cd = nti * NT + n;
    float4 v;
    if (WT) {
        v = __ldg(reinterpret_cast<const float4 *>(W + ((int64_t)k * Cd + cd) * Cs + cs));
    } else {
        const float *src = W + ((int64_t)k * Cs + cs) * Cd + cd;
        v = make_float4(__ldg(src), __ldg(src + Cd), __ldg(src + 2 * (int64_t)Cd), __ldg(src + 3 * (int64_t)Cd));
    }
    out[t] = v;
}

// ------------------------------------------------------------------ wgrad
// dW[k] (Cs x Cd) = sum over the pairs of offset k of  Xrow^T (outer) dYrow.
// MMA view: D[M = 128 input channels][N = NT output channels] += A[M x 8 pairs] * B[8 pairs x N].
// Both operands are MN-major (a gathered row IS contiguous along M resp. N).  For 32-bit
// operands the only MN-major shared-memory layout the tensor core accepts is
// SWIZZLE_128B_BASE32B (cutlass sm100_common.inl: "for mn-major tf32 operands, SW128_32B is
// the only available smem layout"): atoms of 32 channels (128 B) x 4 pairs, Swizzle<2,5,2> on
// the byte address = the 32-byte chunk index XORed with the pair index mod 4.  16-byte cp.async
// pieces stay whole, only their destination moves:
//   addr(pair r, 16-B chunk c) = (c / 8) * 4096 + r * 128 + ((((c % 8) >> 1) ^ (r & 3)) << 1 | (c & 1)) * 16
//   (LBO = 4096 between channel atoms, SBO = 512 between groups of 4 pairs)
// CTA = (offset k, chunk of WG_PAIRS pairs, 128-channel slice of Cs, NT-channel slice of Cd);
// the pair list is the kernel map's compacted list (ascending k, then output row), so no MMA
// cycle and no gathered byte is spent on missing neighbours.  Partial tiles are combined with
// 16-byte fp32 reductions into dW (zeroed by the caller).
constexpr int WG_PAIRS = 4096;  // most pairs per CTA (the launcher takes fewer for small maps, see u2_conv_wgrad_tc)

__host__ __device__ constexpr uint32_t make_idesc_mn(bool bf16, int M, int N) {
    return (bf16 ? make_idesc_bf16(M, N) : make_idesc_tf32(M, N)) | (1u << 15) | (1u << 16);  // a_major = b_major = MN
}

// Geometry of one pipeline stage of the wgrad kernel.
//   tf32: SWIZZLE_128B_BASE32B, atoms of 32 channels x 4 pairs, MMA K = 8 pairs, 32 pairs / stage
//   bf16: SWIZZLE_128B,         atoms of 64 channels x 8 pairs, MMA K = 16 pairs, 64 pairs / stage
template <bool BF16>
struct WgGeom {
    static constexpr int ES = BF16 ? 2 : 4;
    static constexpr int KR = BF16 ? 64 : 32;          // pairs per stage (4 MMAs)
    static constexpr int CPA = 128 / ES;               // channels per 128-byte atom row
    static constexpr int ATOM = KR * 128;              // bytes: KR pair-rows x 128 B  (= LBO between channel atoms)
    static constexpr int A_ATOMS = TILE_M / CPA;       // 4 / 2
    static constexpr int A_BYTES = A_ATOMS * ATOM;     // 16 KB
    static constexpr int SBO = BF16 ? 1024 : 512;      // bytes between swizzle atoms along the pair axis
    static constexpr int MMA_ADV = BF16 ? 2048 : 1024; // bytes of pair rows one MMA consumes
    static constexpr int RB = KR / 32;                 // 32-row blocks per stage
    static constexpr uint64_t LAYOUT = BF16 ? 2 : 1;   // SWIZZLE_128B / SWIZZLE_128B_BASE32B
    // physical 16-byte chunk of logical chunk c (0..7) in pair row r
    __device__ static __forceinline__ int swz(int c, int r) {
        return BF16 ? (c ^ (r & 7)) : ((((c >> 1) ^ (r & 3)) << 1) | (c & 1));
    }
};

struct WgradParams {
    const uint8_t *X;   // rows indexed by the "a" side of a pair (fp32 or bf16)
    const uint8_t *dY;  // rows indexed by the "b" side
    const int *nbr;     // [K, ld] neighbour table the pair list was compacted from
    const int *flat;    // compacted flat indices k*ld + out_row of the valid entries
    const int *nbsizes;
    float *dW;
    int64_t ld;
    int Cs, Cd, K, NT, stages, tmem_cols, swap, n_mt, n_nt, cp_mode, TM;
    int dense_k;   // offset whose pairs are (row j, row j) for every row (centre tap of a submanifold map, k = 1 layers), or -1
    const int *dense_ok;  // optional device flag: 0 = the map does not have that property after all (duplicate coordinates) -> gather path
    int wg_pairs;  // pairs per CTA, multiple of 128, <= WG_PAIRS
    long long *dbg;  // optional per-CTA phase clocks (U2_DEBUG_CONV_TIMING)
    int npw;         // producer warps: 4 (warps 0-3) or 8 (+ warps 6-9) when only one CTA fits an SM
};

constexpr int WG_MAX_THREADS = NUM_THREADS + 4 * 32;  // + 4 optional extra producer warps (warps 6-9)

// 2-D tile load through the TMA engine: box {64 channels, 64 rows} of a row-major bf16 matrix, SWIZZLE_128B — the byte
// pattern r * 128 + ((chunk ^ (r & 7)) << 4) that the gather producers write (WgGeom<true>::swz); rows past the end of the
// matrix are zero-filled by the hardware.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tmap, int c0, int r0, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(r0)
                 : "memory");
}

template <bool BF16>
__global__ void __launch_bounds__(WG_MAX_THREADS) conv_wgrad_tc_kernel(const WgradParams p, const __grid_constant__ CUtensorMap tmapA,
                                                                       const __grid_constant__ CUtensorMap tmapB) {
    using G = WgGeom<BF16>;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int NT = p.NT;
    const int n_batoms = (NT + G::CPA - 1) / G::CPA;
    const int B_BYTES = n_batoms * G::ATOM;
    const int TM = p.TM;                          // 128-channel input slices per CTA (accumulators side by side in TMEM)
    const int A_BYTES = TM * G::A_BYTES;
    const int stage_bytes = A_BYTES + B_BYTES;
    int2 *s_pairs = reinterpret_cast<int2 *>(smem + (size_t)p.stages * stage_bytes);
    uint64_t *s_full = reinterpret_cast<uint64_t *>(s_pairs + p.wg_pairs);
    uint64_t *s_empty = s_full + MAX_STAGES;
    uint64_t *s_accum = s_empty + MAX_STAGES;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_accum + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k = blockIdx.y;
    const int mt = (blockIdx.z / p.n_nt) * p.TM, nt = blockIdx.z % p.n_nt;  // first input slice of this CTA
    // pair range of this CTA (all threads compute the same scalar prefix over <= 32 sizes)
    int koff = (lane < k) ? __ldg(p.nbsizes + lane) : 0;  // K <= 32: one load per lane, then a warp sum
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) koff += __shfl_xor_sync(0xFFFFFFFFu, koff, o);
    const int n_k = __ldg(p.nbsizes + k);
    const int c0 = blockIdx.x * p.wg_pairs;
    if (c0 >= n_k) return;
    const int n_pairs = min(p.wg_pairs, n_k - c0);
    const int n_items = (n_pairs + G::KR - 1) / G::KR;
    long long *dbg = p.dbg ? p.dbg + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 : nullptr;
    if (dbg && tid == 0) { dbg[0] = clock64(); dbg[6] = n_items; }

    // Offset `dense_k` of a submanifold / identity map pairs every row with itself (the centre tap; all of a 1x1x1 conv or
    // Linear layer): its operand rows are CONTIGUOUS, so one thread streams them with 2-D TMA tile loads — 5-7 instructions
    // per 64-pair stage instead of ~100 warp-level LDGSTS, and no pair-list prologue.  ~20 % of the pairs of a k = 3 map.
    const bool dense = BF16 && k == p.dense_k && (p.dense_ok == nullptr || __ldg(p.dense_ok) != 0);
    if (tid == 0) {
        for (int s = 0; s < p.stages; s++) {
            mbar_init(s_full + s, dense ? 1 : p.npw * 32);
            mbar_init(s_empty + s, 1);
        }
        mbar_init(s_accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        tmem_alloc(s_tmem, (uint32_t)p.tmem_cols);
        tmem_relinquish();
    }
    // (a-row, b-row) of every pair of the chunk; padded to a whole stage with -1.  Two dependent global loads per
    // pair (flat index -> neighbour table): 8 pairs per thread are in flight at a time, otherwise this prologue is a
    // chain of ~2 x 11 exposed L2 latencies (measured 20 k cycles per CTA before batching).
    if (!dense) {
        const int *flat = p.flat + koff + c0;
        const int padded = n_items * G::KR;
        const int nthr = blockDim.x;
        for (int pr0 = tid; pr0 < padded; pr0 += 8 * nthr) {
            int f[8], in[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int pr = pr0 + u * nthr;
                f[u] = pr < n_pairs ? __ldg(flat + pr) : -1;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) in[u] = f[u] >= 0 ? __ldg(p.nbr + f[u]) : -1;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int pr = pr0 + u * nthr;
                if (pr >= padded) break;
                const int o = f[u] >= 0 ? f[u] - k * (int)p.ld : -1;
                s_pairs[pr] = p.swap ? make_int2(o, in[u]) : make_int2(in[u], o);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    const int m0 = mt * TILE_M, n0 = nt * NT;
    if (dbg && tid == 0) dbg[1] = clock64();

    if (warp < 4 || warp >= 6) {
        // ============================ producers: gather both operands ============================
        const int pw = warp < 4 ? warp : warp - 2;  // producer index 0 .. npw-1
        const int npw = p.npw;
        long long dbg_wait = 0;
        const int cc = lane & 7;     // 16-byte chunk inside a 128-byte atom row
        const int rsub = lane >> 3;  // pair inside a group of 4
        const size_t a_pitch = (size_t)p.Cs * G::ES, b_pitch = (size_t)p.Cd * G::ES;
        if (dense) {
            if (tid == 0) {
                int n_a = 0;  // channel atoms of the A side that exist (an atom beyond Cs is never fetched, see below)
                for (int a = 0; a < TM * G::A_ATOMS; a++) n_a += (m0 + a * G::CPA < p.Cs) ? 1 : 0;
                const uint32_t tx = (uint32_t)(n_a + n_batoms) * (uint32_t)G::ATOM;
                for (int it = 0; it < n_items; it++) {
                    const int s = it % p.stages;
                    const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                    mbar_wait(s_empty + s, ph ^ 1u);
                    const uint32_t a_base = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint32_t b_base = a_base + A_BYTES;
                    const int r0 = c0 + it * G::KR;  // pair j of the offset = (row j, row j)
                    mbar_arrive_expect_tx(s_full + s, tx);
                    for (int a = 0; a < TM * G::A_ATOMS; a++)
                        if (m0 + a * G::CPA < p.Cs) tma_load_2d(a_base + a * G::ATOM, &tmapA, m0 + a * G::CPA, r0, s_full + s);
                    for (int b = 0; b < n_batoms; b++) tma_load_2d(b_base + b * G::ATOM, &tmapB, n0 + b * G::CPA, r0, s_full + s);
                }
            }
        } else
        for (int it = 0; it < n_items; it++) {
            const int s = it % p.stages;
            const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
            if (dbg && tid == 0) {
                const long long t0 = clock64();
                mbar_wait(s_empty + s, ph ^ 1u);
                dbg_wait += clock64() - t0;
            } else {
                mbar_wait(s_empty + s, ph ^ 1u);
            }
            const uint32_t a_base = smem_u32(smem + (size_t)s * stage_bytes);
            const uint32_t b_base = a_base + A_BYTES;
            const int2 *pairs = s_pairs + it * G::KR;
            // A: units (input slice, channel atom, 32-row block) w, w+4, ...
            for (int u = pw; u < TM * G::A_ATOMS * G::RB; u += npw) {
                const int a_atom = u / G::RB, a_rb = u % G::RB;  // a_atom counts atoms across the TM slices
                // an atom entirely beyond Cs (Cs = 64 or 192: half of a 128-channel slice) is not fetched at all: its
                // shared memory stays stale, which only reaches accumulator rows >= Cs, and those are never stored
                if (m0 + a_atom * G::CPA >= p.Cs) continue;
                const int a_ch = m0 + a_atom * G::CPA + cc * (16 / G::ES);
                const bool a_ch_ok = a_ch < p.Cs;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int r = a_rb * 32 + 4 * i + rsub;
                    const int arow = pairs[r].x;
                    const bool ok = arow >= 0 && a_ch_ok;
                    const uint8_t *src = p.X + (ok ? (size_t)arow * a_pitch + (size_t)a_ch * G::ES : 0);
                    cp_async16_mode(a_base + a_atom * G::ATOM + r * 128 + G::swz(cc, r) * 16, src, ok ? 16u : 0u, p.cp_mode);
                }
            }
            // B: units (channel atom, 32-row block) w, w+4, ...
            for (int u = pw; u < n_batoms * G::RB; u += npw) {
                const int atom = u / G::RB, rb = u % G::RB;
                const int ch = atom * G::CPA + cc * (16 / G::ES);
                const bool ch_ok = ch < NT;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int r = rb * 32 + 4 * i + rsub;
                    const int brow = pairs[r].y;
                    const bool ok = brow >= 0 && ch_ok;
                    const uint8_t *src = p.dY + (ok ? (size_t)brow * b_pitch + (size_t)(n0 + ch) * G::ES : 0);
                    cp_async16_mode(b_base + atom * G::ATOM + r * 128 + G::swz(cc, r) * 16, src, ok ? 16u : 0u, p.cp_mode);
                }
            }
            cp_async_mbar_arrive_noinc(s_full + s);
        }
        // ============================ epilogue: reduce the tile into dW[k] ============================
        if (dbg && tid == 0) { dbg[2] = clock64(); dbg[8] = dbg_wait; }
        mbar_wait(s_accum, 0);
        tc_fence_after();
        // a warp may only touch TMEM lanes 32 * (warp % 4) ..; with 8 producer warps the second set takes the upper
        // half of the columns
        const int q = warp & 3, half = warp >= 6 ? 1 : 0, nh = npw / 4;
        const int c_lo = (NT / 16 * half / nh) * 16, c_hi = (NT / 16 * (half + 1) / nh) * 16;
        for (int tm = 0; tm < TM; tm++) {
            const int ci = m0 + tm * TILE_M + q * 32 + lane;  // accumulator row (TMEM lane) = input channel
            if (m0 + tm * TILE_M >= p.Cs) break;
            float *out = p.dW + ((int64_t)k * p.Cs + ci) * p.Cd + n0;
            for (int c = c_lo; c < c_hi; c += 16) {
                uint32_t v[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tm * NT + c), v);
                tmem_ld_wait();
                if (ci < p.Cs) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + c + j), "f"(__uint_as_float(v[j])),
                                     "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                                     : "memory");
                }
            }
        }
    } else if (warp == 5) {
        // ============================ MMA issuer ============================
        const uint32_t idesc = make_idesc_mn(BF16, TILE_M, NT);
        long long dbg_wait_full = 0;
        for (int it = 0; it < n_items; it++) {
            const int s = it % p.stages;
            const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
            if (dbg && lane == 0) {
                const long long t0 = clock64();
                mbar_wait(s_full + s, ph);
                dbg_wait_full += clock64() - t0;
            } else {
                mbar_wait(s_full + s, ph);
            }
            tc_fence_after();
            proxy_fence_async();
            {   // all 32 lanes, converged (see umma_elect)
                const uint32_t a_base = smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t b_base = a_base + A_BYTES;
                for (int tm = 0; tm < TM; tm++) {
                    if (m0 + tm * TILE_M >= p.Cs) break;
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        const uint64_t ad = make_smem_desc(a_base + tm * G::A_BYTES + g * G::MMA_ADV, G::ATOM, G::SBO) | (G::LAYOUT << 61);
                        const uint64_t bd = make_smem_desc(b_base + g * G::MMA_ADV, G::ATOM, G::SBO) | (G::LAYOUT << 61);
                        umma_elect<BF16>(tmem_base + (uint32_t)(tm * NT), ad, bd, idesc, (it > 0 || g > 0) ? 1u : 0u);
                    }
                }
                umma_commit_elect(s_empty + s);
                if (it == n_items - 1) umma_commit_elect(s_accum);
            }
        }
        if (dbg && lane == 0) dbg[9] = dbg_wait_full;
    }
    if (dbg && tid == 0) dbg[4] = clock64();
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    if (dbg && tid == 160) { dbg[5] = clock64(); unsigned sm; asm("mov.u32 %0, %%smid;" : "=r"(sm)); dbg[7] = sm; }
}

__device__ __forceinline__ unsigned int pack_bf16x2(float lo, float hi) {
    unsigned int r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// fp32 weights -> bf16 blobs [k][cc][nt][chunk][n][8] (K-major UMMA layout, 16-byte chunk = 8 channels)
template <bool WT>
__device__ __forceinline__ void pretile_bf16_piece(const float *__restrict__ W, uint4 *__restrict__ out, int K, int Cs, int Cd, int KC,
                                                   int NT, int64_t t) {
    const int64_t total = (int64_t)K * Cs * Cd / 8;
    if (t >= total) return;
    const int chunks = KC / 8, n_cc = Cs / KC, n_nt = Cd / NT;
    int64_t r = t;
    const int n = (int)(r % NT); r /= NT;
    const int chunk = (int)(r % chunks); r /= chunks;
    const int nti = (int)(r % n_nt); r /= n_nt;
    const int cc = (int)(r % n_cc); r /= n_cc;
    const int k = (int)r;
    const int cs = cc * KC + chunk * 8, cd = nti * NT + n;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; e++)
        v[e] = WT ? __ldg(W + ((int64_t)k * Cd + cd) * Cs + cs + e) : __ldg(W + ((int64_t)k * Cs + cs + e) * Cd + cd);
    out[t] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

template <bool WT>
__global__ void __launch_bounds__(256) pretile_weights_bf16_kernel(const float *__restrict__ W, uint4 *__restrict__ out, int K,
                                                                   int Cs, int Cd, int KC, int NT) {
    pretile_bf16_piece<WT>(W, out, K, Cs, Cd, KC, NT, (int64_t)blockIdx.x * blockDim.x + threadIdx.x);
}

// every (parameter, direction) of a model in ONE launch: the job table says which blocks belong to which re-tiling
struct PretileJob {
    const float *W;
    uint4 *out;
    int64_t block_start;  // first thread block of this job
    int K, Cs, Cd, KC, NT, WT;
};

__global__ void __launch_bounds__(256) pretile_multi_bf16_kernel(const PretileJob *__restrict__ jobs, int n_jobs) {
    int lo = 0, hi = n_jobs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].block_start <= (int64_t)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const PretileJob j = jobs[lo];
    const int64_t t = ((int64_t)blockIdx.x - j.block_start) * blockDim.x + threadIdx.x;
    if (j.WT) pretile_bf16_piece<true>(j.W, j.out, j.K, j.Cs, j.Cd, j.KC, j.NT, t);
    else pretile_bf16_piece<false>(j.W, j.out, j.K, j.Cs, j.Cd, j.KC, j.NT, t);
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float4 *__restrict__ x, int64_t n8, uint4 *__restrict__ y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    const float4 a = __ldg(x + 2 * i), b = __ldg(x + 2 * i + 1);
    y[i] = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
}

// fp32 rows -> bf16 "hi" (round to nearest) and "lo" = bf16(x - hi): x = hi + lo to ~2^-17 relative.  Writes the
// concatenated operand rows [hi | lo | hi] (3C wide, conv forward / dgrad of the bf16x3 mode) and, optionally, contiguous
// hi / lo matrices (wgrad operands).  One thread per 8 channels.
__global__ void __launch_bounds__(256) split_bf16x3_kernel(const float4 *__restrict__ x, int64_t n, int C8, uint4 *__restrict__ out3,
                                                           uint4 *__restrict__ hi, uint4 *__restrict__ lo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * C8) return;
    const int64_t row = i / C8;
    const int c = (int)(i % C8);
    const float4 a = __ldg(x + 2 * i), b = __ldg(x + 2 * i + 1);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float h[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
        h[e] = __bfloat162float(__float2bfloat16_rn(v[e]));
        l[e] = v[e] - h[e];  // exact in fp32
    }
    const uint4 H = make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
    const uint4 L = make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]), pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
    if (out3) {
        uint4 *o = out3 + row * 3 * C8 + c;
        o[0] = H;
        o[C8] = L;
        o[2 * C8] = H;
    }
    if (hi) hi[i] = H;
    if (lo) lo[i] = L;
}

int pick_nt(int Cd) {
    const int tiles = (Cd + 255) / 256;
    if (Cd % tiles) return 0;
    const int nt = Cd / tiles;
    return (nt % 16 == 0 && nt >= 16) ? nt : 0;
}

// bytes of a gathered row per stage (128 preferred, 64 otherwise); 0 = shape not supported
int pick_rowb(int Cs, int es) { return (Cs * es) % 128 == 0 ? 128 : ((Cs * es) % 64 == 0 ? 64 : 0); }
int pick_kc(int Cs) { return pick_rowb(Cs, 4) / 4; }

}  // namespace

int u2_conv_tc_supported(int32_t Cs, int32_t Cd, int32_t K, int32_t math) {
    if (math == U2_MATH_TF32) return pick_rowb(Cs, 4) && pick_nt(Cd) && K <= 32;
    if (math == U2_MATH_BF16) return pick_rowb(Cs, 2) && pick_nt(Cd) && K <= 32 && Cs % 8 == 0;
    return 0;
}

size_t u2_conv_tc_scratch_bytes(int64_t n_dst, int32_t K, int32_t Cs, int32_t Cd, int32_t math) {
    (void)n_dst;
    if (!u2_conv_tc_supported(Cs, Cd, K, math)) return 0;
    return (size_t)K * Cs * Cd * (math == U2_MATH_BF16 ? 2 : 4);
}

template <int ROWB, bool BF16, bool YADD>
static int launch_fwd_v2(const FwdParams &p, dim3 grid, size_t smem, cudaStream_t st) {
    U2_CUDA_OK(cudaFuncSetAttribute(conv_fwd_tc_kernel<ROWB, BF16, YADD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // the launcher sizes the stages for 2-3 co-resident CTAs: ask for the full shared-memory carve-out, otherwise the
    // driver may keep a larger L1 and fewer CTAs fit than planned
    static const int carve = getenv("U2_NO_CARVEOUT") ? -1 : (int)cudaSharedmemCarveoutMaxShared;
    U2_CUDA_OK(cudaFuncSetAttribute(conv_fwd_tc_kernel<ROWB, BF16, YADD>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    if (getenv("U2_DEBUG_CONV_TIMING")) {
        int nb = 0, nb0 = 0, nb32 = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, conv_fwd_tc_kernel<ROWB, BF16, YADD>, NUM_THREADS, smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb0, conv_fwd_tc_kernel<ROWB, BF16, YADD>, NUM_THREADS, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb32, conv_fwd_tc_kernel<ROWB, BF16, YADD>, NUM_THREADS, 32768);
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, conv_fwd_tc_kernel<ROWB, BF16, YADD>);
        int smem_sm = 0, smem_blk = 0, regs_sm = 0;
        cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, 0);
        cudaDeviceGetAttribute(&smem_blk, cudaDevAttrMaxSharedMemoryPerBlockOptin, 0);
        cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, 0);
        fprintf(stderr, "[conv dbg] occupancy: %d CTAs/SM at %zu B dynamic smem (%d at 0 B, %d at 32 KB); regs %d, static smem %zu, "
                        "local %zu, maxDyn %d, carveout %d; device: smem/SM %d, smem/block optin %d, regs/SM %d\n",
                nb, smem, nb0, nb32, fa.numRegs, fa.sharedSizeBytes, fa.localSizeBytes, fa.maxDynamicSharedSizeBytes,
                fa.preferredShmemCarveout, smem_sm, smem_blk, regs_sm);
    }
    conv_fwd_tc_kernel<ROWB, BF16, YADD><<<grid, NUM_THREADS, smem, st>>>(p);
    U2_LAUNCH_OK();
    return 0;
}

template <int ROWB, bool BF16>
static int launch_fwd_v1(const FwdParams &p, dim3 grid, size_t smem, cudaStream_t st) {
    return p.Yadd ? launch_fwd_v2<ROWB, BF16, true>(p, grid, smem, st) : launch_fwd_v2<ROWB, BF16, false>(p, grid, smem, st);
}

// Host side of U2_DEBUG_CONV_TIMING: per-CTA clock64 stamps (16 slots per CTA) -> one summary line.
// slots: 0 start, 1 setup done, 2 producers issued their last item, 4 epilogue done, 5 CTA end, 6 items, 7 SM id,
//        8 producer cycles waiting for a free stage, 9 MMA-thread cycles waiting for data
static void conv_dbg_report(const char *tag, const long long *d_dbg, size_t n_cta) {
    long long *h = (long long *)malloc(n_cta * 16 * sizeof(long long));
    if (cudaMemcpy(h, d_dbg, n_cta * 16 * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) { free(h); return; }
    double items = 0, main = 0, epi = 0, wait_e = 0, wait_f = 0, setup = 0, ph[6] = {0, 0, 0, 0, 0, 0};
    size_t live = 0;
    for (size_t i = 0; i < n_cta; i++) {
        const long long *d = h + i * 16;
        if (d[6] <= 0) continue;
        live++;
        items += (double)d[6];
        setup += (double)(d[1] - d[0]);
        main += (double)(d[2] - d[1]);
        epi += (double)(d[4] - d[2]);
        wait_e += (double)d[8];
        wait_f += (double)d[9];
        for (int j = 0; j < 6; j++) ph[j] += (double)d[10 + j];
    }
    // co-residency actually reached: CTAs of one SM share its clock64, so count overlapping lifetimes per SM
    int max_conc = 0;
    double busy = 0, area = 0;
    for (int sm = 0; sm < 256; sm++) {
        long long lo = 0x7FFFFFFFFFFFFFFFLL, hi = 0;
        for (size_t i = 0; i < n_cta; i++) {
            const long long *d = h + i * 16;
            if (d[6] <= 0 || d[7] != sm) continue;
            if (d[0] < lo) lo = d[0];
            if (d[5] > hi) hi = d[5];
            area += (double)(d[5] - d[0]);
            int conc = 0;
            for (size_t j = 0; j < n_cta; j++) {
                const long long *e = h + j * 16;
                if (e[6] > 0 && e[7] == sm && e[0] <= d[0] && e[5] > d[0]) conc++;
            }
            if (conc > max_conc) max_conc = conc;
        }
        if (hi > lo) busy += (double)(hi - lo);
    }
    if (live)
        fprintf(stderr, "[conv dbg] %s ctas=%zu | items/cta %.1f | per item: issue-span %.0f (producer waits empty %.0f, mma waits "
                        "full %.0f) | setup %.0f  tail+epilogue %.0f cycles/cta | co-resident CTAs/SM max %d avg %.2f\n",
                tag, live, items / live, main / items, wait_e / items, wait_f / items, setup / live, epi / live, max_conc,
                busy > 0 ? area / busy : 0.0);
    if (live && ph[0] + ph[1] + ph[2] > 0)
        fprintf(stderr, "[conv dbg]   per item, producer thread 0: stale-zero+fence %.0f, LDGSTS issue %.0f, noinc arrive %.0f | MMA thread: fences "
                        "%.0f, mma issue %.0f, commit %.0f\n", ph[0] / items, ph[1] / items, ph[2] / items, ph[3] / items, ph[4] / items,
                ph[5] / items);
    free(h);
}

// fp32 weights [K][Cs][Cd] (or [K][Cd][Cs] when w_transposed) -> the blob layout conv_fwd_tc_kernel streams
int u2_conv_pretile_tc(const float *W, int32_t w_transposed, int32_t K, int32_t Cs, int32_t Cd, int32_t math, void *blob,
                       cudaStream_t st) {
    const bool bf16 = math == U2_MATH_BF16;
    const int es = bf16 ? 2 : 4;
    const int ROWB = pick_rowb(Cs, es), NT = pick_nt(Cd);
    const int KC = ROWB ? ROWB / es : 0;
    U2_CHECK_ARG(ROWB && NT && K <= 32, "u2_conv_pretile: unsupported shape Cs=%d Cd=%d K=%d", Cs, Cd, K);
    U2_CHECK_ARG(W && blob && (((uintptr_t)W | (uintptr_t)blob) & 15) == 0, "u2_conv_pretile: null or misaligned pointer");
    if (bf16) {
        const int64_t total8 = (int64_t)K * Cs * Cd / 8;
        if (w_transposed)
            pretile_weights_bf16_kernel<true><<<(unsigned)u2_ceil_div(total8, 256), 256, 0, st>>>(W, (uint4 *)blob, K, Cs, Cd, KC, NT);
        else
            pretile_weights_bf16_kernel<false><<<(unsigned)u2_ceil_div(total8, 256), 256, 0, st>>>(W, (uint4 *)blob, K, Cs, Cd, KC, NT);
    } else {
        const int64_t total4 = (int64_t)K * Cs * Cd / 4;
        if (w_transposed)
            pretile_weights_kernel<true><<<(unsigned)u2_ceil_div(total4, 256), 256, 0, st>>>(W, (float4 *)blob, K, Cs, Cd, KC, NT);
        else
            pretile_weights_kernel<false><<<(unsigned)u2_ceil_div(total4, 256), 256, 0, st>>>(W, (float4 *)blob, K, Cs, Cd, KC, NT);
    }
    U2_LAUNCH_OK();
    return 0;
}

size_t u2_conv_pretile_plan_bytes_tc(int32_t n_jobs) { return (size_t)(n_jobs > 0 ? n_jobs : 0) * sizeof(PretileJob); }

// host side of the one-launch re-tiling: job table for n (parameter, direction) pairs, in the caller's (host) buffer
int u2_conv_pretile_plan_tc(int32_t n_jobs, const uint64_t *W, const uint64_t *blob, const int32_t *K, const int32_t *Cs,
                            const int32_t *Cd, const int32_t *w_transposed, int32_t math, void *plan_host, size_t plan_bytes,
                            int64_t *n_blocks) {
    U2_CHECK_ARG(math == U2_MATH_BF16, "u2_conv_pretile_plan: bf16 blobs only (math %d)", math);
    U2_CHECK_ARG(n_jobs >= 0 && n_blocks && (n_jobs == 0 || (W && blob && K && Cs && Cd && w_transposed && plan_host)),
                 "u2_conv_pretile_plan: null pointer");
    U2_CHECK_ARG(plan_bytes >= u2_conv_pretile_plan_bytes_tc(n_jobs), "u2_conv_pretile_plan: plan buffer too small");
    PretileJob *jobs = (PretileJob *)plan_host;
    int64_t blocks = 0;
    for (int i = 0; i < n_jobs; i++) {
        const int ROWB = pick_rowb(Cs[i], 2), NT = pick_nt(Cd[i]);
        U2_CHECK_ARG(ROWB && NT && K[i] > 0 && K[i] <= 32, "u2_conv_pretile_plan: job %d: unsupported shape Cs=%d Cd=%d K=%d", i, Cs[i],
                     Cd[i], K[i]);
        U2_CHECK_ARG(W[i] && blob[i] && ((W[i] | blob[i]) & 15) == 0, "u2_conv_pretile_plan: job %d: null or misaligned pointer", i);
        jobs[i].W = (const float *)(uintptr_t)W[i];
        jobs[i].out = (uint4 *)(uintptr_t)blob[i];
        jobs[i].block_start = blocks;
        jobs[i].K = K[i]; jobs[i].Cs = Cs[i]; jobs[i].Cd = Cd[i]; jobs[i].KC = ROWB / 2; jobs[i].NT = NT; jobs[i].WT = w_transposed[i] ? 1 : 0;
        blocks += u2_ceil_div((int64_t)K[i] * Cs[i] * Cd[i] / 8, 256);
    }
    U2_CHECK_ARG(blocks < (int64_t)1 << 31, "u2_conv_pretile_plan: too many thread blocks");
    *n_blocks = blocks;
    return 0;
}

int u2_conv_pretile_run_tc(const void *plan_dev, int32_t n_jobs, int64_t n_blocks, cudaStream_t st) {
    U2_CHECK_ARG(n_jobs >= 0 && n_blocks >= 0 && (n_jobs == 0 || plan_dev), "u2_conv_pretile_run: bad arguments");
    if (n_jobs == 0 || n_blocks == 0) return 0;
    pretile_multi_bf16_kernel<<<(unsigned)n_blocks, 256, 0, st>>>((const PretileJob *)plan_dev, n_jobs);
    U2_LAUNCH_OK();
    return 0;
}

// X: fp32 rows (math TF32) or bf16 rows (math BF16); W the fp32 parameter tensor, or NULL when `scratch` already holds the
// blob u2_conv_pretile_tc wrote for this (shape, direction, math).
int u2_conv_fwd_tc(const void *X, int64_t n_src, int32_t Cs, const float *W, int32_t w_transposed, const int32_t *table,
                   const int32_t *perm, int64_t ld, int64_t n_dst, int32_t K, int32_t Cd, float *Y, int32_t math,
                   void *scratch, size_t scratch_bytes, float *tile_stats, const float *Yadd, cudaStream_t st) {
    const bool bf16 = math == U2_MATH_BF16;
    const int es = bf16 ? 2 : 4;
    U2_CHECK_ARG(!tile_stats || pick_nt(Cd) % 32 == 0, "u2_conv_fwd_tc: fused column statistics need Cd tiles of 32 (Cd=%d)", Cd);
    U2_CHECK_ARG(n_src * (int64_t)Cs * es < 0xFFFFFFFFLL, "u2_conv_fwd_tc: source tensor too large for 32-bit byte offsets");
    if (n_dst == 0) return 0;
    const int ROWB = pick_rowb(Cs, es), NT = pick_nt(Cd);
    const int KC = ROWB / es;
    U2_CHECK_ARG(ROWB && NT && K <= 32, "u2_conv_fwd_tc: unsupported shape Cs=%d Cd=%d K=%d", Cs, Cd, K);
    U2_CHECK_ARG(ld % TILE_M == 0, "u2_conv_fwd_tc: table leading dimension must be a multiple of 128");
    U2_CHECK_ARG(scratch && scratch_bytes >= (size_t)K * Cs * Cd * es, "u2_conv_fwd_tc: scratch too small");
    U2_CHECK_ARG((((uintptr_t)X | (uintptr_t)W | (uintptr_t)Y | (uintptr_t)scratch | (uintptr_t)Yadd) & 15) == 0,
                 "u2_conv_fwd_tc: pointers must be 16-byte aligned");
    if (W && u2_conv_pretile_tc(W, w_transposed, K, Cs, Cd, math, scratch, st)) return 1;
    (void)KC;

    FwdParams p;
    p.X = (const uint8_t *)X; p.Wt = (const uint8_t *)scratch; p.table = table; p.perm = perm; p.Y = Y;
    p.dbg = nullptr;
    p.tile_stats = tile_stats;
    p.Yadd = Yadd;
    static const int cp_mode = getenv("U2_CPASYNC_MODE") ? atoi(getenv("U2_CPASYNC_MODE")) : 1;  // .ca measured 15-25 % faster
    p.cp_mode = cp_mode;
    const char *diag_env = getenv("U2_CONV_DIAG");  // read per call: scripts/diag_conv.py flips it between launches
    p.diag = diag_env ? atoi(diag_env) : 0;
    p.ld = ld; p.n_dst = n_dst; p.Cs = Cs; p.Cd = Cd; p.K = K; p.NT = NT;
    int cols = 32;
    while (cols < NT) cols <<= 1;
    p.tmem_cols = cols;
    const int chunks = ROWB / 16;
    const size_t stage_bytes = (size_t)chunks * A_LBO + (size_t)chunks * NT * 16;
    // U2_CONV_KERNEL=ps selects the persistent warp-specialised kernel (measured slower than the co-resident-CTA kernel on
    // every SPVCNN layer; read per call so that tests can exercise both)
    const char *kern_env = getenv("U2_CONV_KERNEL");
    const bool use_ps = (kern_env && !strcmp(kern_env, "ps")) || (getenv("U2_DEBUG_CONV_TIMING") && atoi(getenv("U2_DEBUG_CONV_TIMING")) == 2);
    if (use_ps) {
        // persistent warp-specialised kernel: one CTA per SM, the whole shared memory as one ring of stages
        const PsTab TL = ps_tab_layout(K);
        const size_t ps_fixed = 4 * 4096 + 2 * (size_t)TL.bytes + 1024;
        int S = (int)((227 * 1024 - ps_fixed) / stage_bytes);
        static const int s_cap = getenv("U2_CONV_PS_STAGES") ? atoi(getenv("U2_CONV_PS_STAGES")) : PS_MAX_STAGES;
        if (S > s_cap) S = s_cap;
        if (S > PS_MAX_STAGES) S = PS_MAX_STAGES;
        U2_CHECK_ARG(S >= 2, "u2_conv_fwd_tc: tile does not fit shared memory");
        p.stages = S;
        int cols2 = 32;
        while (cols2 < 2 * NT) cols2 <<= 1;
        p.tmem_cols = cols2;
        p.ld = perm ? ld : u2_ceil_div(n_dst, TILE_M) * TILE_M;  // rows walked by the tiles (the table is padded to ld >= this)
        const size_t ps_smem = (size_t)S * stage_bytes + ps_fixed;
        const int64_t n_work = (p.ld / TILE_M) * (Cd / NT);
        static const int ps_ctas = getenv("U2_CONV_PS_CTAS") ? atoi(getenv("U2_CONV_PS_CTAS")) : U2_NUM_SMS;
        dim3 pgrid((unsigned)(n_work < ps_ctas ? n_work : ps_ctas));
        p.ld = ld;
        p.n_tiles = perm ? ld / TILE_M : u2_ceil_div(n_dst, TILE_M);
#define U2_PS_LAUNCH(RB, BF, YA)                                                                                                  \
    do {                                                                                                                          \
        U2_CUDA_OK(cudaFuncSetAttribute(conv_fwd_ps_kernel<RB, BF, YA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ps_smem)); \
        conv_fwd_ps_kernel<RB, BF, YA><<<pgrid, PS_THREADS, ps_smem, st>>>(p);                                                    \
    } while (0)
        static const int ps_dbg = getenv("U2_DEBUG_CONV_TIMING") && atoi(getenv("U2_DEBUG_CONV_TIMING")) == 2;
        if (ps_dbg) {
            U2_CUDA_OK(cudaMalloc(&p.dbg, (size_t)pgrid.x * 32 * sizeof(long long)));
            U2_CUDA_OK(cudaMemsetAsync(p.dbg, 0, (size_t)pgrid.x * 32 * sizeof(long long), st));
        }
        if (bf16) {
            if (ROWB == 128) { if (Yadd) U2_PS_LAUNCH(128, true, true); else U2_PS_LAUNCH(128, true, false); }
            else { if (Yadd) U2_PS_LAUNCH(64, true, true); else U2_PS_LAUNCH(64, true, false); }
        } else {
            if (ROWB == 128) { if (Yadd) U2_PS_LAUNCH(128, false, true); else U2_PS_LAUNCH(128, false, false); }
            else { if (Yadd) U2_PS_LAUNCH(64, false, true); else U2_PS_LAUNCH(64, false, false); }
        }
#undef U2_PS_LAUNCH
        U2_LAUNCH_OK();
        if (ps_dbg) {
            U2_CUDA_OK(cudaStreamSynchronize(st));
            const size_t nc = pgrid.x;
            long long *h = (long long *)malloc(nc * 32 * sizeof(long long));
            U2_CUDA_OK(cudaMemcpy(h, p.dbg, nc * 32 * sizeof(long long), cudaMemcpyDeviceToHost));
            double a[32] = {0};
            for (size_t c = 0; c < nc; c++) {
                for (int j = 2; j < 32; j++) a[j] += (double)h[c * 32 + j];
                a[0] += (double)(h[c * 32 + 1] - h[c * 32]);
            }
            const double it = a[2] > 0 ? a[2] : 1, pit = a[8] > 0 ? a[8] : 1, tl = a[18] > 0 ? a[18] : 1;
            fprintf(stderr, "[conv ps dbg] Cs=%d Cd=%d NT=%d rows=%lld stages=%d ctas=%zu | cta cycles %.0f, tiles/cta %.1f, items/cta %.1f (%.0f cycles/item)\n"
                            "[conv ps dbg]   MMA thread per item: wait full %.0f, fences %.0f, issue %.0f, commit %.0f | per tile: wait table/acc %.0f\n"
                            "[conv ps dbg]   producer warp 0 per own item: wait empty %.0f, zero+fence %.0f, gather issue %.0f, arrive %.0f | per tile: wait table %.0f\n"
                            "[conv ps dbg]   epilogue warp 0 per tile: wait acc %.0f, work %.0f | table warp 0 per tile: wait %.0f, work %.0f\n",
                    Cs, Cd, NT, (long long)n_dst, S, nc, a[0] / nc, tl / nc, it / nc, a[0] / it,
                    a[3] / it, a[4] / it, a[5] / it, a[6] / it, a[7] / tl,
                    a[9] / pit, a[10] / pit, a[11] / pit, a[12] / pit, a[13] / tl,
                    a[14] / tl, a[15] / tl, a[16] / tl, a[17] / tl);
            free(h);
            cudaFree(p.dbg);
        }
        return 0;
    }
    // neighbour table + per-offset compact row lists / presence masks / counts + per-stage dirty masks + barriers
    const size_t fixed = (size_t)K * TILE_M * sizeof(int) + (size_t)K * (TILE_M + 16 + 4) + MAX_STAGES * 16 +
                         (2 * MAX_STAGES + 1) * sizeof(uint64_t) + 16;
    // prefer two CTAs per SM (one gathers while the other drains its accumulator) if that
    // still leaves >= 3 stages each; otherwise one CTA with as many stages as fit
    const size_t budget = 227 * 1024;
    // co-resident CTAs hide each other's prologue / epilogue: take the most CTAs per SM that still leave
    // two pipeline stages each (measured: 2 CTAs x 2 stages beat 1 CTA x 4 stages by 9 % on 192->192)
    static const int max_ctas = getenv("U2_CONV_MAX_CTAS") ? atoi(getenv("U2_CONV_MAX_CTAS")) : 3;
    int stages = 0;
    for (int c = max_ctas; c >= 1 && stages < 2; c--) {
        if (c * cols > 512 || budget / c < fixed + 1024 + 2 * stage_bytes) continue;
        stages = (int)((budget / c - fixed - 1024) / stage_bytes);
    }
    U2_CHECK_ARG(stages >= 2, "u2_conv_fwd_tc: tile does not fit shared memory");
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    p.stages = stages;
    const size_t smem = stages * stage_bytes + fixed;
    // with a row permutation the tiles run over the (padded) table rows, destination rows come from perm
    dim3 grid((unsigned)u2_ceil_div(perm ? ld : n_dst, TILE_M), (unsigned)(Cd / NT));
    static const int debug_timing = getenv("U2_DEBUG_CONV_TIMING") ? 1 : 0;
    if (debug_timing) {  // per-CTA phase clocks: where does a work item's time go (synchronises; diagnostics only)
        const size_t n_cta = (size_t)grid.x * grid.y;
        long long *d_dbg = nullptr;
        U2_CUDA_OK(cudaMalloc(&d_dbg, n_cta * 16 * sizeof(long long)));
        U2_CUDA_OK(cudaMemsetAsync(d_dbg, 0, n_cta * 16 * sizeof(long long), st));
        p.dbg = d_dbg;
        int rc = bf16 ? (ROWB == 128 ? launch_fwd_v1<128, true>(p, grid, smem, st) : launch_fwd_v1<64, true>(p, grid, smem, st))
                      : (ROWB == 128 ? launch_fwd_v1<128, false>(p, grid, smem, st) : launch_fwd_v1<64, false>(p, grid, smem, st));
        if (rc) return rc;
        U2_CUDA_OK(cudaStreamSynchronize(st));
        char tag[128];
        snprintf(tag, sizeof tag, "fwd Cs=%d Cd=%d NT=%d rows=%lld stages=%d", Cs, Cd, NT, (long long)n_dst, stages);
        conv_dbg_report(tag, d_dbg, n_cta);
        cudaFree(d_dbg);
        return 0;
    }
    if (bf16) return ROWB == 128 ? launch_fwd_v1<128, true>(p, grid, smem, st) : launch_fwd_v1<64, true>(p, grid, smem, st);
    return ROWB == 128 ? launch_fwd_v1<128, false>(p, grid, smem, st) : launch_fwd_v1<64, false>(p, grid, smem, st);
}

int u2_split_bf16x3_impl(const float *x, int64_t n, int32_t C, void *out3, void *hi, void *lo, cudaStream_t st) {
    U2_CHECK_ARG(C > 0 && C % 8 == 0, "u2_split_bf16x3: C %% 8 == 0 required (C=%d)", C);
    U2_CHECK_ARG((((uintptr_t)x | (uintptr_t)out3 | (uintptr_t)hi | (uintptr_t)lo) & 15) == 0, "u2_split_bf16x3: 16-byte alignment required");
    if (n == 0) return 0;
    const int64_t total = n * (C / 8);
    split_bf16x3_kernel<<<(unsigned)u2_ceil_div(total, 256), 256, 0, st>>>((const float4 *)x, n, C / 8, (uint4 *)out3, (uint4 *)hi, (uint4 *)lo);
    U2_LAUNCH_OK();
    return 0;
}

int u2_cast_bf16_impl(const float *x, int64_t n, void *y, cudaStream_t st) {
    U2_CHECK_ARG(n % 8 == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0, "u2_cast_bf16: n %% 8 == 0 and 16-byte alignment required");
    if (n == 0) return 0;
    cast_bf16_kernel<<<(unsigned)u2_ceil_div(n / 8, 256), 256, 0, st>>>((const float4 *)x, n / 8, (uint4 *)y);
    U2_LAUNCH_OK();
    return 0;
}

typedef CUresult (*U2EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// tensor map over a row-major bf16 matrix [rows, C]: box = 64 channels x 64 rows, SWIZZLE_128B, zero fill outside
static bool u2_make_rows_tmap(CUtensorMap *tm, const void *base, int64_t rows, int C) {
    static U2EncodeTiledFn encode = []() -> U2EncodeTiledFn {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return (U2EncodeTiledFn)ptr;
    }();
    if (!encode || rows <= 0 || ((uintptr_t)base & 15) || (C * 2) % 16) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)C * 2};
    const cuuint32_t box[2] = {64u, 64u};
    const cuuint32_t estr[2] = {1u, 1u};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// X / dY: fp32 rows (math TF32) or bf16 rows (math BF16).  dense_k >= 0: offset dense_k pairs row j with row j for every
// j < n_rows (both matrices have n_rows rows) — its operands are then streamed with TMA tile loads (bf16 only).
int u2_conv_wgrad_tc(const void *X, int32_t Cs, const void *dY, int32_t Cd, const int32_t *nbr, int64_t ld, int64_t n_rows,
                     int32_t K, const int32_t *flat, const int32_t *nbsizes, int32_t swap, float *dW, int32_t math,
                     int32_t dense_k, const int32_t *dense_ok, cudaStream_t st) {
    const bool bf16 = math == U2_MATH_BF16;
    const int NT = pick_nt(Cd);
    U2_CHECK_ARG(NT && Cs % (bf16 ? 8 : 4) == 0 && K <= 32, "u2_conv_wgrad_tc: unsupported shape Cs=%d Cd=%d K=%d", Cs, Cd, K);
    U2_CHECK_ARG((((uintptr_t)X | (uintptr_t)dY | (uintptr_t)dW) & 15) == 0, "u2_conv_wgrad_tc: pointers must be 16-byte aligned");
    U2_CHECK_ARG((int64_t)K * ld < 0x7FFFFFFFLL, "u2_conv_wgrad_tc: table too large for int32 flat indices");
    U2_CUDA_OK(cudaMemsetAsync(dW, 0, (size_t)K * Cs * Cd * sizeof(float), st));
    if (n_rows == 0) return 0;
    WgradParams p;
    p.X = (const uint8_t *)X; p.dY = (const uint8_t *)dY; p.nbr = nbr; p.flat = flat; p.nbsizes = nbsizes; p.dW = dW;
    p.ld = ld; p.Cs = Cs; p.Cd = Cd; p.K = K; p.NT = NT; p.swap = swap;
    static const int cp_mode_w = getenv("U2_CPASYNC_MODE_W") ? atoi(getenv("U2_CPASYNC_MODE_W")) : 1;
    p.cp_mode = cp_mode_w;
    p.n_mt = (Cs + TILE_M - 1) / TILE_M;
    p.n_nt = Cd / NT;
    const int cpa = bf16 ? 64 : 32, atom = bf16 ? 8192 : 4096;
    // Pairs per CTA.  A CTA walks its pairs serially (64 per stage), so a small map cut into 4096-pair chunks leaves
    // most SMs idle behind a few long chains (measured: every layer below ~100k rows took the same 0.16 ms).  Aim
    // at `target` working CTAs, assuming ~40 % of the K * n_rows table entries are present; the price of smaller
    // chunks is one more fp32 reduction of the Cs x Cd tile into dW per chunk, hence the floor.
    static const int wg_fixed = getenv("U2_WGRAD_PAIRS") ? atoi(getenv("U2_WGRAD_PAIRS")) : 0;
    static const int wg_min = getenv("U2_WGRAD_PAIRS_MIN") ? atoi(getenv("U2_WGRAD_PAIRS_MIN")) : 2048;
    static const int wg_target = getenv("U2_WGRAD_TARGET_CTAS") ? atoi(getenv("U2_WGRAD_TARGET_CTAS")) : 8 * U2_NUM_SMS;
    {
        const int tm0 = (p.n_mt < 4 ? p.n_mt : 4);
        const int64_t slices = (int64_t)u2_ceil_div(p.n_mt, tm0) * p.n_nt;
        int64_t want = (int64_t)(0.4 * (double)n_rows * K) * slices / wg_target;
        want = (want + 127) / 128 * 128;
        if (want < wg_min) want = wg_min;
        if (want > WG_PAIRS) want = WG_PAIRS;
        if (wg_fixed > 0) want = wg_fixed;
        p.wg_pairs = (int)want;
    }
    const size_t fixed = (size_t)p.wg_pairs * sizeof(int2) + (2 * MAX_STAGES + 1) * sizeof(uint64_t) + 16 + 1024;
    const size_t budget = 227 * 1024;
    // input slices per CTA: every slice re-gathers the dY rows, so take as many as TMEM (512 columns)
    // and shared memory (>= 3 stages) allow
    static const int tm_max = getenv("U2_WGRAD_TM") ? atoi(getenv("U2_WGRAD_TM")) : 4;
    int TM = p.n_mt < tm_max ? p.n_mt : tm_max;
    while (TM > 1 && (TM * NT > 512 ||
                      (budget - fixed) / ((size_t)(TM * (TILE_M / cpa) + (NT + cpa - 1) / cpa) * atom) < 3))
        TM--;
    p.TM = TM;
    int cols = 32;
    while (cols < TM * NT) cols <<= 1;
    p.tmem_cols = cols;
    const size_t stage_bytes = (size_t)(TM * (TILE_M / cpa) + (NT + cpa - 1) / cpa) * atom;
    static const int max_ctas_w = getenv("U2_WGRAD_MAX_CTAS") ? atoi(getenv("U2_WGRAD_MAX_CTAS")) : 2;
    static const int min_st_w = getenv("U2_WGRAD_MIN_STAGES") ? atoi(getenv("U2_WGRAD_MIN_STAGES")) : 3;
    int stages = 0;
    for (int c = max_ctas_w; c >= 1 && stages < (c > 1 ? min_st_w : 2); c--) {
        if (c * cols > 512 || budget / c < fixed + 1024 + 2 * stage_bytes) { stages = 0; continue; }
        stages = (int)((budget / c - fixed - 1024) / stage_bytes);
    }
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    U2_CHECK_ARG(stages >= 2, "u2_conv_wgrad_tc: tile does not fit shared memory");
    p.stages = stages;
    const size_t smem = stages * stage_bytes + fixed;
    // One warp keeps issuing a 16-byte LDGSTS only every ~120 cycles; two co-resident CTAs (8 producer warps) come close
    // to what the LSU takes, a lone CTA (wide layers: its accumulators fill TMEM) does not -> give it 8 producer warps.
    static const int npw_env = getenv("U2_WGRAD_NPW") ? atoi(getenv("U2_WGRAD_NPW")) : 0;
    const bool lone = 2 * cols > 512 || 2 * (smem + 1024) > budget;
    p.npw = npw_env == 4 || npw_env == 8 ? npw_env : (lone ? 8 : 4);
    const int threads = NUM_THREADS + (p.npw == 8 ? 4 * 32 : 0);
    // a single offset has at most n_rows pairs (one per row of the table's row side)
    dim3 grid((unsigned)u2_ceil_div(n_rows, p.wg_pairs), (unsigned)K, (unsigned)(u2_ceil_div(p.n_mt, TM) * p.n_nt));
    p.dbg = nullptr;
    static const int debug_timing = getenv("U2_DEBUG_CONV_TIMING") ? 1 : 0;
    const size_t n_cta = (size_t)grid.x * grid.y * grid.z;
    if (debug_timing) {
        U2_CUDA_OK(cudaMalloc(&p.dbg, n_cta * 16 * sizeof(long long)));
        U2_CUDA_OK(cudaMemsetAsync(p.dbg, 0, n_cta * 16 * sizeof(long long), st));
    }
    alignas(64) CUtensorMap tmA, tmB;
    memset(&tmA, 0, sizeof tmA);
    memset(&tmB, 0, sizeof tmB);
    const int tma_on = getenv("U2_WGRAD_TMA") ? atoi(getenv("U2_WGRAD_TMA")) : 1;  // read per call: tests compare both paths
    p.dense_k = -1;
    if (bf16 && tma_on && dense_k >= 0 && dense_k < K && u2_make_rows_tmap(&tmA, X, n_rows, Cs) && u2_make_rows_tmap(&tmB, dY, n_rows, Cd))
        p.dense_k = dense_k;
    p.dense_ok = dense_ok;
    if (bf16) {
        U2_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        U2_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_tc_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        getenv("U2_NO_CARVEOUT") ? -1 : (int)cudaSharedmemCarveoutMaxShared));
        conv_wgrad_tc_kernel<true><<<grid, threads, smem, st>>>(p, tmA, tmB);
    } else {
        U2_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        U2_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_tc_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        getenv("U2_NO_CARVEOUT") ? -1 : (int)cudaSharedmemCarveoutMaxShared));
        conv_wgrad_tc_kernel<false><<<grid, threads, smem, st>>>(p, tmA, tmB);
    }
    U2_LAUNCH_OK();
    if (debug_timing) {
        U2_CUDA_OK(cudaStreamSynchronize(st));
        char tag[160];
        snprintf(tag, sizeof tag, "wgrad Cs=%d Cd=%d NT=%d TM=%d rows=%lld pairs/cta=%d stages=%d npw=%d", Cs, Cd, NT, TM,
                 (long long)n_rows, p.wg_pairs, stages, p.npw);
        conv_dbg_report(tag, p.dbg, n_cta);
        cudaFree(p.dbg);
    }
    return 0;
}
