// Shared device/host helpers for libanncur_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <float.h>

#include "../../include/anncur_b200.h"

namespace anncur {

// ---- host-side error plumbing (thread-local message, ABI error codes) ----------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define ANNCUR_CUDA_OK(expr)                                                                   \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            anncur::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                              __LINE__);                                                        \
            return ANNCUR_E_CUDA;                                                               \
        }                                                                                       \
    } while (0)

#define ANNCUR_LAUNCH_OK(name)                                                                  \
    do {                                                                                        \
        cudaError_t e__ = cudaGetLastError();                                                   \
        if (e__ != cudaSuccess) {                                                               \
            anncur::set_error("launch of %s failed: %s (%s:%d)", name, cudaGetErrorString(e__), \
                              __FILE__, __LINE__);                                              \
            return ANNCUR_E_CUDA;                                                               \
        }                                                                                       \
        anncur::count_launch();                                                                 \
    } while (0)

#define ANNCUR_REQUIRE(cond, ...)                \
    do {                                         \
        if (!(cond)) {                           \
            anncur::set_error(__VA_ARGS__);      \
            return ANNCUR_E_INVALID;             \
        }                                        \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
int sm_count();

// ---- candidate keys ----------------------------------------------------------------------------
// A candidate is one 64-bit key: high word = score mapped to an order-preserving unsigned,
// low word = ~index.  Larger key <=> (higher score) or (equal score and LOWER index), so one
// unsigned compare implements the library's total order, keys of distinct items are distinct,
// and key 0 is below every real candidate (used as padding).
__host__ __device__ __forceinline__ uint32_t float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t idx) {
    return (uint64_t(float_to_ordered(score)) << 32) | uint64_t(~idx);
}
__host__ __device__ __forceinline__ float key_score(uint64_t key) { return ordered_to_float(uint32_t(key >> 32)); }
__host__ __device__ __forceinline__ uint32_t key_index(uint64_t key) { return ~uint32_t(key); }

// Survivor lists of the fused kernel hold RAW entries -- low word = fp32 score bits, high word = item index -- so that a
// push is a plain 64-bit store of two registers; the readers (list compaction, select kernels) turn them into keys.
__host__ __device__ __forceinline__ uint64_t raw_to_key(uint64_t raw) {
    return (uint64_t(float_to_ordered(
#ifdef __CUDA_ARCH__
        __uint_as_float(uint32_t(raw))
#else
        [](uint32_t u) { union { float f; uint32_t u; } c; c.u = u; return c.f; }(uint32_t(raw))
#endif
        )) << 32) | uint64_t(~uint32_t(raw >> 32));
}
__host__ __device__ __forceinline__ uint64_t key_to_raw(uint64_t key) {
    const float score = key_score(key);
#ifdef __CUDA_ARCH__
    const uint32_t bits = __float_as_uint(score);
#else
    union { float f; uint32_t u; } c; c.f = score; const uint32_t bits = c.u;
#endif
    return (uint64_t(key_index(key)) << 32) | uint64_t(bits);
}

#define ANNCUR_PAD_VAL (-FLT_MAX)

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        uint64_t t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t < v ? t : v;
    }
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// In-place bitonic sort, DESCENDING, of n_pow2 keys in shared memory by all threads of the CTA.
__device__ __forceinline__ void block_bitonic_sort_desc(uint64_t* keys, int n_pow2) {
    for (int size = 2; size <= n_pow2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (n_pow2 >> 1); t += blockDim.x) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool desc = ((lo & size) == 0);
                uint64_t a = keys[lo], b = keys[hi];
                if ((a < b) == desc) { keys[lo] = b; keys[hi] = a; }
            }
        }
    }
    __syncthreads();
}
#endif  // __CUDACC__

}  // namespace anncur
