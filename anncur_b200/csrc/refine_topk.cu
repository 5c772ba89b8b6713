// F32R, second half: exact fp32 re-scoring of the survivors of the one-pass filter, then the row top-k.
//
// The fused kernel (score_topk_umma.cu, kind F32R) leaves, per query row, the items whose score UPPER bound
// UB_n = A_n + b_n reached the row's push threshold T (a lower bound of the k-th best fp32 score).  One warp per row:
//   1. gather the row's candidates (keys ordered by UB) into shared memory;
//   2. re-score the k candidates with the largest UB in fp32 from the item-major copy of E in the packed index
//      (S_n = sum_i q_i e_in, the reference's torch.matmul arithmetic, eval/matrix_approx_zeshel.py:109-119);
//      the smallest of these k scores, s_min, is a lower bound of the true k-th best score;
//   3. re-score every other candidate whose UB reaches s_min (an item with UB < s_min cannot be in the top-k);
//   4. top-k of the re-scored candidates by (S, lower index), sorted best-first (torch.topk, :121-126);
//   5. certificate: every item that was NOT pushed has S <= UB <= T, so the result is the exact top-k if the k-th
//      re-scored value is > T.  Rows that fail it, rows with fewer than min(k, N) candidates and rows with a list
//      that filled up are flagged for the REDO pass (3-pass kind, streaming from -inf).
// On the bench workloads a row has ~400 candidates, of which k + ~6 are re-scored (2 KB of E each, one coalesced
// 512-byte request per warp load).
#include "common.cuh"
#include "kernels.h"
#include "warp_select.cuh"

#include <cstdlib>
#include <type_traits>

namespace anncur {

constexpr int kRefMaxLists = 1024;
constexpr int kRefBatch = 4;           // candidates re-scored together (their loads are all in flight at once)

struct RefineParams {
    const uint64_t* cand;     // [row][n_lists][cap] RAW entries (common.cuh: raw_to_key); score word = UB in scaled units
    const uint32_t* counts;   // [row][n_lists]; top bit = the list filled up
    int n_lists, cap, n_rows, k, n_sort;
    int64_t n_items, idx_offset;
    const float* Q;           // [row][ldq] fp32 queries as the caller passed them
    int ldq, k_dim;
    const float* ET;          // [item][ld] fp32 item-major copy of E
    int ld;
    const float* row_inv_scale;   // UB (scaled units) * row_inv_scale[row] = UB in the units of S
    float* out_vals;
    int64_t* out_idx;
    uint32_t* thr_shared;     // in: the row's push threshold (exclusive, ordered image); out: -inf = REDO this row, +inf = done
    uint32_t* mtile_flags;    // [m_tiles] query-tile flags, then [1] number of flagged rows, then the flagged rows
    int m_tiles;
};

// dot(q, ET[item]) for kRefBatch = 4 items at once; q is held in registers as U float4 per lane (k_dim <= 128 U).
// All 4 x U loads of a lane are issued before the first use.  Returns, in every lane, the score of item[lane >> 3]
// (the four partial sums are reduced together: 6 shuffles instead of 20).
template <int U>
__device__ __forceinline__ float rescore_batch(const float4 (&qv)[U], const float* __restrict__ ET, int ld, int ld4,
                                               const uint32_t (&item)[kRefBatch]) {
    static_assert(kRefBatch == 4, "the reduction below is written for 4 candidates");
    const int lane = int(lane_id());
    // The loads are volatile asm and are followed by one (empty) volatile asm per loaded vector, so that neither
    // compiler stage can sink a load down to its first use: all 4 x U requests are in flight before the first FMA
    // (without this the compiler interleaved load / use and every load paid its full latency: 134 -> 313 us at C2).
    float4 e[kRefBatch][U];
#pragma unroll
    for (int c = 0; c < kRefBatch; ++c) {
        const float4* rowp = reinterpret_cast<const float4*>(ET + int64_t(item[c]) * ld);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = lane + 32 * u;
            const float4* ptr = rowp + (j < ld4 ? j : 0);                  // out-of-range lanes re-read element 0; their q is 0
            asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(e[c][u].x), "=f"(e[c][u].y), "=f"(e[c][u].z), "=f"(e[c][u].w) : "l"(ptr));
        }
    }
#pragma unroll
    for (int c = 0; c < kRefBatch; ++c) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            asm volatile("" : "+f"(e[c][u].x), "+f"(e[c][u].y), "+f"(e[c][u].z), "+f"(e[c][u].w));
        }
    }
    // packed fp32x2 FMAs (sm_100: fma.rn.f32x2): the (x, y) and (z, w) halves of a float4 accumulate side by side
    float s[kRefBatch];
#pragma unroll
    for (int c = 0; c < kRefBatch; ++c) {
        uint64_t acc2 = 0ull;                                            // (0.f, 0.f)
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint64_t q_lo, q_hi, e_lo, e_hi;
            asm("mov.b64 %0, {%1, %2};" : "=l"(q_lo) : "f"(qv[u].x), "f"(qv[u].y));
            asm("mov.b64 %0, {%1, %2};" : "=l"(q_hi) : "f"(qv[u].z), "f"(qv[u].w));
            asm("mov.b64 %0, {%1, %2};" : "=l"(e_lo) : "f"(e[c][u].x), "f"(e[c][u].y));
            asm("mov.b64 %0, {%1, %2};" : "=l"(e_hi) : "f"(e[c][u].z), "f"(e[c][u].w));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2) : "l"(q_lo), "l"(e_lo));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2) : "l"(q_hi), "l"(e_hi));
        }
        float a_lo, a_hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a_lo), "=f"(a_hi) : "l"(acc2));
        s[c] = a_lo + a_hi;
    }
    const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
    float a = (hi16 ? s[2] : s[0]) + __shfl_xor_sync(0xffffffffu, hi16 ? s[0] : s[2], 16);
    float b = (hi16 ? s[3] : s[1]) + __shfl_xor_sync(0xffffffffu, hi16 ? s[1] : s[3], 16);
    float v = (hi8 ? b : a) + __shfl_xor_sync(0xffffffffu, hi8 ? a : b, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// any k_dim: q re-read from global (L1) per 128-float chunk
__device__ __forceinline__ float rescore_long(const float* __restrict__ q, int k_dim, const float* __restrict__ ET, int ld,
                                              int ld4, uint32_t item) {
    const int lane = int(lane_id());
    const float4* rowp = reinterpret_cast<const float4*>(ET + int64_t(item) * ld);
    float acc = 0.f;
    for (int j = lane; j < ld4; j += 32) {
        const float4 e = __ldg(rowp + j);
        const int i = 4 * j;
        acc = fmaf(i + 0 < k_dim ? __ldg(q + i + 0) : 0.f, e.x, acc);
        acc = fmaf(i + 1 < k_dim ? __ldg(q + i + 1) : 0.f, e.y, acc);
        acc = fmaf(i + 2 < k_dim ? __ldg(q + i + 2) : 0.f, e.z, acc);
        acc = fmaf(i + 3 < k_dim ? __ldg(q + i + 3) : 0.f, e.w, acc);
    }
    return warp_sum(acc);
}

// One CTA of W warps per row.  W = 1 for large batches (every row is one warp's private chain, ~16 rows in flight per
// SM); W = 4 for small batches, where the kernel time IS the per-row latency: the gather and the two re-scoring phases are
// split over the warps (the selects stay on warp 0).
// Register budget 128: with less, ptxas sinks the 16 loads of a re-scoring batch down to their uses.
template <int W> __device__ __forceinline__ void cta_sync() { if (W > 1) __syncthreads(); else __syncwarp(); }

// CAP = candidates of one row held in shared memory (1024; 2048 / 4096 for large k)
template <int U, int W, int CAP>
__global__ void __launch_bounds__(W * 32, W == 1 ? (CAP > 2048 ? 6 : CAP > 1024 ? 11 : 16) : W == 4 ? 4 : 2)
refine_topk_kernel(const RefineParams p) {
    constexpr int kRefCap = CAP;
    extern __shared__ __align__(16) uint64_t ref_smem[];
    const int warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    uint64_t* keys = ref_smem;                                            // [kRefCap]
    uint64_t* sh64 = ref_smem + kRefCap;                                  // [2]: prefix, mask of the phase-A select
    uint32_t* hist = reinterpret_cast<uint32_t*>(sh64 + 2);               // [256]
    float* sh_min = reinterpret_cast<float*>(hist + 256);                 // [W]
    uint32_t* sh_total = reinterpret_cast<uint32_t*>(sh_min + W);         // [2]: total, marked
    uint16_t* work = reinterpret_cast<uint16_t*>(sh_total + 2) + warp * (kRefCap / W);   // [kRefCap / W] per warp: positions to re-score
    uint32_t* offs = sh_total + 2 + kRefCap / 2;                          // [n_lists + 1]
    const uint32_t k = uint32_t(p.k);
    const int ld4 = p.ld >> 2;
    const uint32_t lowest = float_to_ordered(-INFINITY);

    for (int row = blockIdx.x; row < p.n_rows; row += gridDim.x) {
        // ---- list lengths -> offsets; a marked list gives the row up (warp 0) ----------------------------
        if (warp == 0) {
            bool marked = false;
            for (int l = int(lane); l < p.n_lists; l += 32) {
                const uint32_t c = __ldcg(p.counts + int64_t(row) * p.n_lists + l);
                marked |= (c & 0x80000000u) != 0u;
                offs[l + 1] = min(c & 0x7fffffffu, uint32_t(p.cap));
            }
            marked = __any_sync(0xffffffffu, marked);
            __syncwarp();
            uint32_t carry = 0;
            for (int base = 0; base < p.n_lists; base += 32) {
                const int l = base + int(lane);
                const uint32_t v = l < p.n_lists ? offs[l + 1] : 0u;
                uint32_t incl = v;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= uint32_t(off)) incl += t;
                }
                if (l < p.n_lists) offs[l + 1] = carry + incl;
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) { offs[0] = 0; sh_total[0] = carry; sh_total[1] = marked ? 1u : 0u; }
        }
        cta_sync<W>();
        const uint32_t total = sh_total[0];
        const int64_t need = p.n_items < int64_t(k) ? p.n_items : int64_t(k);
        bool redo = sh_total[1] != 0u || int64_t(total) < need || total > uint32_t(kRefCap);
        const float to_scaled = 1.0f / p.row_inv_scale[row];               // power of two: exact
        const uint32_t thr_ord = __ldcg(p.thr_shared + row);

        if (!redo) {
            // ---- gather: list by list (a list is a contiguous run of <= cap entries), the loads of 8 lists in flight ------
            for (int l0 = warp * 8; l0 < p.n_lists; l0 += 8 * W) {
                uint64_t reg[8];
                uint32_t base[8], cnt[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int l = l0 + u;
                    base[u] = l < p.n_lists ? offs[l] : 0u;
                    cnt[u] = l < p.n_lists ? offs[l + 1] - base[u] : 0u;
                    reg[u] = lane < cnt[u] ? __ldcg(p.cand + (int64_t(row) * p.n_lists + l) * p.cap + lane) : 0ull;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (lane < cnt[u]) keys[base[u] + lane] = raw_to_key(reg[u]);
                    for (uint32_t t = lane + 32u; t < cnt[u]; t += 32u)      // the rare list with more than 32 entries
                        keys[base[u] + t] = raw_to_key(__ldcg(p.cand + (int64_t(row) * p.n_lists + l0 + u) * p.cap + t));
                }
            }
            // ---- the query row, in registers (zero beyond k_dim: those lanes re-read element 0 of the item) ---
            const float* q = p.Q + int64_t(row) * p.ldq;
            float4 qv[U > 0 ? U : 1];
            if (U > 0) {
#pragma unroll
                for (int u = 0; u < (U > 0 ? U : 1); ++u) {
                    const int i = 4 * (int(lane) + 32 * u);
                    qv[u].x = i + 0 < p.k_dim ? __ldg(q + i + 0) : 0.f;
                    qv[u].y = i + 1 < p.k_dim ? __ldg(q + i + 1) : 0.f;
                    qv[u].z = i + 2 < p.k_dim ? __ldg(q + i + 2) : 0.f;
                    qv[u].w = i + 3 < p.k_dim ? __ldg(q + i + 3) : 0.f;
                }
            }
            cta_sync<W>();
            // re-score the candidates at positions work[0 .. n_work), four at a time: keys[pos] becomes (S, item);
            // s_min collects this lane's share of the minimum
            float s_min = INFINITY;
            auto rescore = [&](uint32_t n_work) {
                for (uint32_t i0 = 0; i0 < n_work; i0 += kRefBatch) {
                    uint32_t pos[kRefBatch], item[kRefBatch];
                    const int n = int(min(uint32_t(kRefBatch), n_work - i0));
#pragma unroll
                    for (int c = 0; c < kRefBatch; ++c) {
                        pos[c] = work[min(i0 + uint32_t(c), n_work - 1u)];
                        item[c] = key_index(keys[pos[c]]);
                    }
                    const int c_own = int(lane >> 3);
                    float s_own;
                    if (U > 0) {
                        s_own = rescore_batch<(U > 0 ? U : 1)>(qv, p.ET, p.ld, ld4, item);
                    } else {
                        float s4[kRefBatch];
#pragma unroll
                        for (int c = 0; c < kRefBatch; ++c) s4[c] = rescore_long(q, p.k_dim, p.ET, p.ld, ld4, item[c]);
                        s_own = c_own == 0 ? s4[0] : c_own == 1 ? s4[1] : c_own == 2 ? s4[2] : s4[3];
                    }
                    const uint32_t pos_own = c_own == 0 ? pos[0] : c_own == 1 ? pos[1] : c_own == 2 ? pos[2] : pos[3];
                    const uint32_t item_own = c_own == 0 ? item[0] : c_own == 1 ? item[1] : c_own == 2 ? item[2] : item[3];
                    __syncwarp();
                    if (c_own < n) {
                        s_min = fminf(s_min, s_own);
                        if ((lane & 7u) == 0u) keys[pos_own] = make_key(s_own, item_own);
                    }
                    __syncwarp();
                }
            };
            // ---- phase A: the k candidates with the largest upper bound ------------------------------------
            if (warp == 0) {
                uint64_t prefix = 0ull, mask = 0ull;
                if (total > k) warp_radix_kth(keys, total, k, hist, prefix, mask);
                if (lane == 0) { sh64[0] = prefix; sh64[1] = mask; }
            }
            cta_sync<W>();
            {
                // this warp's winners -> work list (so that every batch of the re-scoring is full), then the re-scoring.
                // (an L2 prefetch of the next candidates was tried here: no gain -- the row's time goes to issue latency
                // of the many small steps at 16 warps per SM, not to waiting for DRAM)
                const uint64_t prefix = sh64[0], mask = sh64[1];
                uint32_t n_work = 0;
                for (uint32_t t0 = uint32_t(warp) * 32u; t0 < total; t0 += 32u * W) {
                    const uint32_t t = t0 + lane;
                    const bool win = t < total && (keys[t] & mask) >= prefix;
                    const uint32_t ballot = __ballot_sync(0xffffffffu, win);
                    if (win) work[n_work + __popc(ballot & ((1u << lane) - 1u))] = uint16_t(t);
                    if (lane == 0) hist[t0 >> 5] = ballot;                 // who has been re-scored (hist is free here)
                    n_work += __popc(ballot);
                }
                __syncwarp();
                rescore(n_work);
            }
            s_min = -warp_max_f(-s_min);
            if (W > 1) {
                if (lane == 0) sh_min[warp] = s_min;
                __syncthreads();
#pragma unroll
                for (int w = 0; w < W; ++w) s_min = fminf(s_min, sh_min[w]);
            }
            // ---- phase B: every other candidate whose upper bound reaches s_min; the rest is dropped --------
            if (total > k) {
                const float s_min_scaled = s_min * to_scaled;
                __syncwarp();
                uint32_t n_work = 0;
                for (uint32_t t0 = uint32_t(warp) * 32u; t0 < total; t0 += 32u * W) {
                    const uint32_t t = t0 + lane;
                    const bool done = (hist[t0 >> 5] >> lane) & 1u;
                    const bool open = t < total && !done;
                    const bool more = open && key_score(keys[t]) >= s_min_scaled;
                    if (open && !more) keys[t] = 0ull;
                    const uint32_t ballot = __ballot_sync(0xffffffffu, more);
                    if (more) work[n_work + __popc(ballot & ((1u << lane) - 1u))] = uint16_t(t);
                    n_work += __popc(ballot);
                }
                __syncwarp();
                rescore(n_work);
            }
            cta_sync<W>();
            // ---- top-k of the re-scored candidates (warp 0) --------------------------------------------------
            if (warp == 0) {
                uint64_t prefix, mask;
                uint32_t n_live = warp_compact_ge(keys, total, 0ull, 0ull);   // drops the zeroed entries
                if (n_live > uint32_t(p.n_sort)) {                          // otherwise the sort below orders them all
                    warp_radix_kth(keys, n_live, k, hist, prefix, mask);
                    n_live = warp_compact_ge(keys, n_live, prefix, mask);
                }
                for (uint32_t t = n_live + lane; t < uint32_t(p.n_sort); t += 32) keys[t] = 0ull;
                __syncwarp();
                warp_bitonic_sort_desc(keys, uint32_t(p.n_sort));
                // ---- certificate: items that were never pushed have S <= UB <= T (all in scaled units) ---------
                if (int64_t(total) < p.n_items && thr_ord > lowest) {
                    const float kth_scaled = key_score(keys[need - 1]) * to_scaled;
                    if (!(kth_scaled > ordered_to_float(thr_ord))) redo = true;
                }
                if (!redo) {
                    for (uint32_t t = lane; t < k; t += 32) {
                        const uint64_t key = keys[t];
                        const bool ok = key != 0ull;
                        p.out_vals[int64_t(row) * k + t] = ok ? key_score(key) : ANNCUR_PAD_VAL;
                        p.out_idx[int64_t(row) * k + t] = ok ? int64_t(key_index(key)) + p.idx_offset : int64_t(-1);
                    }
                }
            }
        }
        if (warp == 0 && lane == 0) {
            p.thr_shared[row] = float_to_ordered(redo ? -INFINITY : INFINITY);
            if (redo) {
                atomicOr(p.mtile_flags + row / 128, 1u);
                uint32_t* n_flagged = p.mtile_flags + p.m_tiles;
                n_flagged[1 + atomicAdd(n_flagged, 1u)] = uint32_t(row);
            }
        }
        cta_sync<W>();
    }
}

int refine_topk_keylists(const uint64_t* cand, const uint32_t* counts, int n_lists, int cap, int n_rows, int k,
                         int64_t idx_offset, const float* Q, int ldq, int k_dim, const float* ET, int ld,
                         const float* row_inv_scale, float* out_vals, int64_t* out_idx, uint32_t* thr_shared,
                         uint32_t* mtile_flags, int m_tiles, int64_t n_items, int row_cap, cudaStream_t stream) {
    if (row_cap != 1024 && row_cap != 2048 && row_cap != 4096) { set_error("refine_topk_keylists: row_cap %d is not 1024, 2048 or 4096", row_cap); return ANNCUR_E_INVALID; }
    if (n_lists > kRefMaxLists) { set_error("refine_topk_keylists: %d lists per row > %d", n_lists, kRefMaxLists); return ANNCUR_E_UNSUPPORTED; }
    if (k > row_cap) { set_error("refine_topk_keylists: k = %d > %d", k, row_cap); return ANNCUR_E_UNSUPPORTED; }
    if (n_rows == 0) return ANNCUR_OK;
    int n_sort = 2;
    while (n_sort < k) n_sort <<= 1;
    RefineParams p{cand, counts, n_lists, cap, n_rows, k, n_sort, n_items, idx_offset, Q, ldq, k_dim, ET, ld,
                   row_inv_scale, out_vals, out_idx, thr_shared, mtile_flags, m_tiles};
    const bool wide = n_rows <= 4 * sm_count();               // few rows: 4 warps per row
    static const int forced_w = [] { const char* e = getenv("ANNCUR_REFINE_WARPS"); return e ? atoi(e) : 0; }();
    const bool wide8 = forced_w == 8 || (forced_w == 0 && n_rows <= sm_count());   // a row per SM at most: 8 warps per row
    const int W = wide8 ? 8 : wide ? 4 : 1;
    const size_t smem = (size_t(row_cap) + 2) * sizeof(uint64_t) + (256 + size_t(W) + 2 + size_t(row_cap) / 2 + size_t(n_lists) + 1) * sizeof(uint32_t);
    const int grid = n_rows < 32 * sm_count() ? n_rows : 32 * sm_count();
    const int ld4 = ld >> 2;
    auto launch = [&](auto kernel) {
        ANNCUR_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        kernel<<<grid, W * 32, smem, stream>>>(p);
        ANNCUR_LAUNCH_OK("refine_topk_kernel");
        return ANNCUR_OK;
    };
    auto pick = [&](auto u_tag) {
        constexpr int U = decltype(u_tag)::value;
        if (wide8) return row_cap > 2048 ? launch(refine_topk_kernel<U, 8, 4096>) : row_cap > 1024 ? launch(refine_topk_kernel<U, 8, 2048>) : launch(refine_topk_kernel<U, 8, 1024>);
        if (wide) return row_cap > 2048 ? launch(refine_topk_kernel<U, 4, 4096>) : row_cap > 1024 ? launch(refine_topk_kernel<U, 4, 2048>) : launch(refine_topk_kernel<U, 4, 1024>);
        return row_cap > 2048 ? launch(refine_topk_kernel<U, 1, 4096>) : row_cap > 1024 ? launch(refine_topk_kernel<U, 1, 2048>) : launch(refine_topk_kernel<U, 1, 1024>);
    };
    if (ld4 <= 32) return pick(std::integral_constant<int, 1>{});
    if (ld4 <= 64) return pick(std::integral_constant<int, 2>{});
    if (ld4 <= 128) return pick(std::integral_constant<int, 4>{});
    return pick(std::integral_constant<int, 0>{});
}

}  // namespace anncur
