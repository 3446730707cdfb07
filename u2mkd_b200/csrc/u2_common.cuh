// Shared helpers for the sm_100a kernels behind include/u2mkd.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/u2mkd.h"

#define U2_NUM_SMS 148

void u2_set_error(const char *fmt, ...);

#define U2_CHECK_ARG(cond, ...)          \
    do {                                 \
        if (!(cond)) {                   \
            u2_set_error(__VA_ARGS__);   \
            return 1;                    \
        }                                \
    } while (0)

#define U2_CUDA_OK(expr)                                                              \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            u2_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                         __FILE__, __LINE__);                                         \
            return 2;                                                                 \
        }                                                                             \
    } while (0)

#define U2_LAUNCH_OK()                                                                \
    do {                                                                              \
        cudaError_t _e = cudaGetLastError();                                          \
        if (_e != cudaSuccess) {                                                      \
            u2_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),  \
                         __FILE__, __LINE__);                                         \
            return 3;                                                                 \
        }                                                                             \
    } while (0)

static inline int64_t u2_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// FNV-1a over (x, y, z, b) folded to 60 bits — the reference's key function
// [TS v1.4.0 backend/hash/hash_cuda.cu]; SURVEY.md A.5.
__host__ __device__ __forceinline__ int64_t u2_fnv4(int x, int y, int z, int b) {
    unsigned long long h = 14695981039346656037ULL;
    h ^= (unsigned int)x; h *= 1099511628211ULL;
    h ^= (unsigned int)y; h *= 1099511628211ULL;
    h ^= (unsigned int)z; h *= 1099511628211ULL;
    h ^= (unsigned int)b; h *= 1099511628211ULL;
    h = (h >> 60) ^ (h & 0xFFFFFFFFFFFFFFFULL);
    return (int64_t)h;
}

// ---- open-addressing table: 16-byte slots {key, value, pad} ----
struct __align__(16) U2Slot {
    unsigned long long key;
    unsigned int val;
    unsigned int pad;
};
#define U2_EMPTY_KEY 0xFFFFFFFFFFFFFFFFULL

__host__ __device__ __forceinline__ unsigned long long u2_mix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

static inline uint64_t u2_table_capacity(int64_t n_keys) {
    uint64_t cap = 1024;
    while (cap < (uint64_t)(2 * n_keys)) cap <<= 1;
    return cap;
}

__device__ __forceinline__ void u2_table_insert(U2Slot *table, unsigned long long mask,
                                                unsigned long long key, unsigned int val) {
    unsigned long long s = u2_mix64(key) & mask;
    while (true) {
        unsigned long long prev = atomicCAS(&table[s].key, U2_EMPTY_KEY, key);
        if (prev == U2_EMPTY_KEY || prev == key) {
            atomicMin(&table[s].val, val);  // duplicates: smallest index wins (deterministic)
            return;
        }
        s = (s + 1) & mask;
    }
}

// returns value or -1
__device__ __forceinline__ int u2_table_lookup(const U2Slot *__restrict__ table, unsigned long long mask,
                                               unsigned long long key) {
    unsigned long long s = u2_mix64(key) & mask;
    while (true) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(table + s));
        const unsigned long long k = ((unsigned long long)raw.y << 32) | raw.x;
        if (k == key) return (int)raw.z;
        if (k == U2_EMPTY_KEY) return -1;
        s = (s + 1) & mask;
    }
}
