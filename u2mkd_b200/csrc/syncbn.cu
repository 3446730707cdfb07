// SyncBatchNorm statistics exchange over NVLink peer memory (one tiny kernel, no NCCL launch).
//
// The reference synchronises BatchNorm across data-parallel ranks with torch SyncBatchNorm
// (core/models/utils.py:138-141, train_spformer.py:77-83): per layer an all_gather of (mean, invstd, count) forward and
// an all_reduce of (sum dy, sum dy*xhat) backward — 2 NCCL collectives x 52 layers per step, each latency-bound and on
// the critical path.  Here every rank owns a symmetric buffer that all peers have mapped (torch symmetric memory:
// plumbing only); an exchange is a PUSH all-reduce of a short fp64 vector:
//   1. store my vector into slot [seq % SLOTS][my rank] of EVERY peer's buffer (NVLink stores, fire and forget),
//   2. __threadfence_system(), then release-store `seq` into the peer's flag [slot][my rank],
//   3. spin (acquire loads of LOCAL memory) until all flags of the slot carry `seq`,
//   4. sum the `world` vectors of the slot in rank order (bit-identical on every rank) in place of the input.
// Exchanges of one group are totally ordered (every rank runs the same layers in the same order), so a slot is reused
// only after every rank has published a LATER sequence number, i.e. after it finished reading: SLOTS = 4 is ample.
// A rank that never arrives makes the others trap after a bounded (~30 s) spin instead of hanging the GPU.
#include <stdlib.h>

#include "u2_common.cuh"

namespace {

constexpr int SB_SLOTS = 4;
constexpr int SB_MAX_WORLD = 16;
constexpr int SB_MAX_LEN = 2 * 1024 + 8;  // 2C + 1 doubles for C <= 1024, padded

// layout of one rank's symmetric buffer
//   flags : uint64 [SB_SLOTS][SB_MAX_WORLD]
//   data  : double [SB_SLOTS][SB_MAX_WORLD][SB_MAX_LEN]
constexpr size_t SB_FLAG_BYTES = sizeof(unsigned long long) * SB_SLOTS * SB_MAX_WORLD;
constexpr size_t SB_BYTES = SB_FLAG_BYTES + sizeof(double) * SB_SLOTS * SB_MAX_WORLD * SB_MAX_LEN;

struct PeerPtrs {
    unsigned char *p[SB_MAX_WORLD];
};

__device__ __forceinline__ void st_release_sys(unsigned long long *addr, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *addr) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(addr) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) syncbn_exchange_kernel(double *__restrict__ vals, int len, const PeerPtrs peers, int world,
                                                              int rank, unsigned long long seq) {
    const int slot = (int)(seq % SB_SLOTS);
    const int tid = threadIdx.x;
    // 1. push my vector to every rank (my own buffer included: the reader treats all ranks alike)
    for (int r = 0; r < world; r++) {
        double *dst = reinterpret_cast<double *>(peers.p[r] + SB_FLAG_BYTES) + ((size_t)slot * SB_MAX_WORLD + rank) * SB_MAX_LEN;
        for (int i = tid; i < len; i += blockDim.x) dst[i] = vals[i];
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish
    if (tid < world) {
        unsigned long long *flag = reinterpret_cast<unsigned long long *>(peers.p[tid]) + slot * SB_MAX_WORLD + rank;
        st_release_sys(flag, seq);
    }
    // 3. wait for everybody (flags live in MY buffer)
    if (tid < world) {
        const unsigned long long *flag = reinterpret_cast<const unsigned long long *>(peers.p[rank]) + slot * SB_MAX_WORLD + tid;
        const long long t0 = clock64();
        while (ld_acquire_sys(flag) != seq) {
            if (clock64() - t0 > 60000000000LL) __trap();  // ~30 s: a rank left the lockstep (start-up skew between ranks can reach seconds)
        }
    }
    __syncthreads();
    // 4. sum in rank order
    const double *src = reinterpret_cast<const double *>(peers.p[rank] + SB_FLAG_BYTES) + (size_t)slot * SB_MAX_WORLD * SB_MAX_LEN;
    for (int i = tid; i < len; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < world; r++) s += __ldcg(src + (size_t)r * SB_MAX_LEN + i);  // L2: written by the peers
        vals[i] = s;
    }
}

}  // namespace

extern "C" size_t u2_syncbn_buffer_bytes(void) { return SB_BYTES; }
extern "C" int32_t u2_syncbn_max_len(void) { return SB_MAX_LEN; }
extern "C" int32_t u2_syncbn_max_world(void) { return SB_MAX_WORLD; }

// vals fp64 [len] (device, local): replaced by the sum over the `world` ranks.  peer_bufs: HOST array of `world` device
// pointers, peer_bufs[r] = rank r's symmetric buffer (u2_syncbn_buffer_bytes(), zero-filled once before the first
// exchange) as mapped into THIS process.  seq: 1, 2, 3, ... identical on every rank for the same exchange.
extern "C" int u2_syncbn_exchange(double *vals, int32_t len, const void *const *peer_bufs, int32_t world, int32_t rank,
                                  uint64_t seq, u2_stream_t stream) {
    U2_CHECK_ARG(vals && peer_bufs && len > 0 && len <= SB_MAX_LEN, "u2_syncbn_exchange: bad vector (len=%d, max %d)", len, SB_MAX_LEN);
    U2_CHECK_ARG(world >= 1 && world <= SB_MAX_WORLD && rank >= 0 && rank < world && seq > 0, "u2_syncbn_exchange: bad world/rank/seq");
    PeerPtrs pp;
    for (int r = 0; r < SB_MAX_WORLD; r++) pp.p[r] = r < world ? (unsigned char *)const_cast<void *>(peer_bufs[r]) : nullptr;
    for (int r = 0; r < world; r++) U2_CHECK_ARG(pp.p[r] && ((uintptr_t)pp.p[r] & 15) == 0, "u2_syncbn_exchange: null / misaligned peer buffer %d", r);
    syncbn_exchange_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(vals, len, pp, world, rank, (unsigned long long)seq);
    U2_LAUNCH_OK();
    return 0;
}
