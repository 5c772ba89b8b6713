// Warp-level selection primitives over 64-bit candidate keys held in shared memory (see common.cuh for the key
// order).  Every function is called by all 32 lanes of one warp; no block-wide barriers.
#pragma once
#include "common.cuh"

namespace anncur {

// MSD radix select of the k-th largest of keys[0 .. total) (total > k >= 1).  On return the winners are the keys with
// (key & mask) >= prefix -- exactly k of them (keys of distinct items are distinct).  The candidates of one row usually
// passed the same threshold, so their leading bytes coincide: the select starts at the first byte in which the row's
// largest and smallest key differ.  hist: 256 words of shared memory owned by this warp.
__device__ __forceinline__ void warp_radix_kth(const uint64_t* keys, uint32_t total, uint32_t k, uint32_t* hist,
                                               uint64_t& prefix_out, uint64_t& mask_out) {
    const uint32_t lane = lane_id();
    uint64_t kmax = 0ull, kmin = ~0ull;
    for (uint32_t t = lane; t < total; t += 32) {
        const uint64_t key = keys[t];
        kmax = key > kmax ? key : kmax;
        kmin = key < kmin ? key : kmin;
    }
    kmin = warp_min_u64(kmin);
    kmax = ~warp_min_u64(~kmax);
    const int top_bit = 63 - __clzll((long long)((kmax ^ kmin) | 1ull));
    const int first_shift = top_bit & ~7;
    uint64_t mask = first_shift >= 56 ? 0ull : ~((1ull << (first_shift + 8)) - 1ull);
    uint64_t prefix = kmax & mask;
    uint32_t need = k;
    for (int shift = first_shift; shift >= 0; shift -= 8) {
#pragma unroll
        for (int b = 0; b < 8; ++b) hist[lane * 8 + b] = 0;
        __syncwarp();
        for (uint32_t t = lane; t < total; t += 32) {
            const uint64_t key = keys[t];
            if ((key & mask) == prefix) atomicAdd(&hist[uint32_t(key >> shift) & 255u], 1u);
        }
        __syncwarp();
        uint32_t c[8], lane_sum = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) { c[b] = hist[lane * 8 + b]; lane_sum += c[b]; }
        uint32_t suf = lane_sum;                 // inclusive suffix sum towards the higher digits
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t t = __shfl_down_sync(0xffffffffu, suf, off);
            if (lane + off < 32) suf += t;
        }
        uint32_t running = suf - lane_sum;
        bool found = false;
        uint32_t d = 0, new_need = 0, bucket = 0;
#pragma unroll
        for (int b = 7; b >= 0; --b) {
            if (!found && running + c[b] >= need) { found = true; d = lane * 8 + b; new_need = need - running; bucket = c[b]; }
            running += c[b];
        }
        const uint32_t ballot = __ballot_sync(0xffffffffu, found);
        __syncwarp();
        if (ballot == 0) break;
        const int srcl = 31 - __clz(int(ballot));
        d = __shfl_sync(0xffffffffu, d, srcl);
        need = __shfl_sync(0xffffffffu, new_need, srcl);
        bucket = __shfl_sync(0xffffffffu, bucket, srcl);
        prefix |= uint64_t(d) << shift;
        mask |= 0xffull << shift;
        if (bucket == need) break;               // digit bucket taken whole
    }
    prefix_out = prefix;
    mask_out = mask;
}

// In-place stable compaction of the keys with key != 0 and (key & mask) >= prefix to the front of keys[0 .. total)
// (the write index never passes the read index).  Returns how many were kept.
__device__ __forceinline__ uint32_t warp_compact_ge(uint64_t* keys, uint32_t total, uint64_t prefix, uint64_t mask) {
    const uint32_t lane = lane_id();
    uint32_t running = 0;
    for (uint32_t t0 = 0; t0 < total; t0 += 32) {
        const uint32_t t = t0 + lane;
        const uint64_t key = t < total ? keys[t] : 0ull;
        const bool keep = key != 0ull && (key & mask) >= prefix;
        const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) keys[running + __popc(ballot & ((1u << lane) - 1u))] = key;
        running += __popc(ballot);
        __syncwarp();
    }
    return running;
}

// Bitonic sort, descending, of keys[0 .. n_sort) (n_sort a power of two >= 2) by one warp.
__device__ __forceinline__ void warp_bitonic_sort_desc(uint64_t* keys, uint32_t n_sort) {
    const uint32_t lane = lane_id();
    for (uint32_t size = 2; size <= n_sort; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = lane; t < (n_sort >> 1); t += 32) {
                const uint32_t lo = 2 * t - (t & (stride - 1));
                const uint32_t hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const uint64_t a = keys[lo], b = keys[hi];
                if ((a < b) == desc) { keys[lo] = b; keys[hi] = a; }
            }
            __syncwarp();
        }
    }
}

}  // namespace anncur
