// Per-row top-k selection over 64-bit candidate keys (see common.cuh for the key order).
//
// One CTA per row.  Three key sources share the kernel:
//   DenseRow   : a row of fp32 scores            (torch.topk(row, k) of the eval loops; the
//                 FFMA fallback of score+top-k)   -- HBM/L2-bound streaming, 8-bit radix select
//   KeyLists   : per-(row, chunk) survivor lists written by the fused tcgen05 kernel
//   PairLists  : (val, idx) candidate lists       (merge of all-gathered per-shard top-k)
// Selection = MSD radix select on the key (early exit as soon as a digit bucket is taken whole),
// then the <= k winners are sorted best-first in shared memory.  When a row's candidates fit in
// shared memory they are sorted there directly and the radix passes are skipped.
#include "common.cuh"
#include "kernels.h"
#include "warp_select.cuh"

namespace anncur {

constexpr int kSelectThreads = 512;
constexpr int kSmemSortCap = 4096;  // keys sorted directly in smem when a row has at most this many
constexpr int kMaxLists = 1024;     // survivor lists per row (= item chunks of the fused kernel)

struct DenseRow {
    static constexpr bool kIsLists = false;
    const float* S;
    int64_t lds;
    int64_t n_cols;
    __device__ int64_t size(int) const { return n_cols; }
    __device__ uint64_t key(int row, int64_t j) const {
        return make_key(__ldg(S + int64_t(row) * lds + j), uint32_t(j));
    }
};

struct KeyLists {
    static constexpr bool kIsLists = true;
    const uint64_t* keys;     // [row][n_lists][cap] RAW entries (common.cuh: raw_to_key)
    const uint32_t* counts;   // [row][n_lists]
    int n_lists;
    int cap;
    // flag_mode 1: rows whose lists hold fewer than min(k, n_items) candidates are flagged for the REDO pass
    // (thr_shared[row] = -inf, query-tile flag set), every other row gets thr_shared[row] = +inf;
    // flag_mode 2: only rows flagged by mode 1 are processed (thr_shared[row] != +inf).
    int flag_mode;
    uint32_t* thr_shared;
    uint32_t* mtile_flags;    // [m_tiles] query-tile flags, then [1] number of flagged rows, then [n_rows] flagged rows
    int m_tiles;
    int64_t n_items;
    const uint32_t* row_list; // optional: [0] = count, then the rows to process (CTA kernel, after the warp kernel)
    // addressing: entry (row, list, pos) lives at keys[row * row_stride + list * list_stride + pos].  Fused-kernel lists:
    // list_stride = cap, row_stride = n_lists * cap, RAW entries, lengths in `counts`.  All-gathered per-shard top-k
    // lists ([shard][row][k] keys): list_stride = n_rows * k, row_stride = k, entries are keys, every list holds `cap`.
    int64_t list_stride, row_stride;
    int raw;
    __device__ uint32_t count(int row, int list) const {
        return counts ? min(__ldcg(counts + int64_t(row) * n_lists + list), uint32_t(cap)) : uint32_t(cap);
    }
    __device__ const uint64_t* list_ptr(int row, int list) const { return keys + int64_t(row) * row_stride + int64_t(list) * list_stride; }
    __device__ uint64_t load(const uint64_t* p) const { const uint64_t v = __ldcg(p); return raw ? raw_to_key(v) : v; }
    __device__ uint32_t* n_flagged() const { return mtile_flags + m_tiles; }
    __device__ uint32_t* flagged_rows() const { return mtile_flags + m_tiles + 1; }
    __device__ int64_t size(int) const { return int64_t(n_lists) * cap; }
    __device__ uint64_t key(int row, int64_t j) const {
        int list = int(j / cap), pos = int(j % cap);
        uint32_t c = count(row, list);
        return pos < int(c) ? load(list_ptr(row, list) + pos) : 0ull;
    }
};

struct PairLists {
    static constexpr bool kIsLists = false;
    const float* vals;        // [row][n_cand]
    const int64_t* idx;       // [row][n_cand], < 0 = padding
    int n_cand;
    __device__ int64_t size(int) const { return n_cand; }
    __device__ uint64_t key(int row, int64_t j) const {
        int64_t i = __ldg(idx + int64_t(row) * n_cand + j);
        return i < 0 ? 0ull : make_key(__ldg(vals + int64_t(row) * n_cand + j), uint32_t(i));
    }
};

struct SelectOut {
    float* vals;              // [row][k]
    int64_t* idx;             // [row][k]
    int64_t idx_offset;
    const float* row_scale;   // optional: value = score * row_scale[row]
    int k;
    int n_sort;               // next_pow2(k) (radix path) -- smem holds max(n_sort, kSmemSortCap) keys
};

__device__ __forceinline__ void write_sorted(const uint64_t* sel, int row, const SelectOut& o) {
    float scale = o.row_scale ? o.row_scale[row] : 1.0f;
    for (int t = threadIdx.x; t < o.k; t += blockDim.x) {
        uint64_t key = sel[t];
        bool ok = key != 0ull;
        o.vals[int64_t(row) * o.k + t] = ok ? key_score(key) * scale : ANNCUR_PAD_VAL;
        o.idx[int64_t(row) * o.k + t] = ok ? int64_t(key_index(key)) + o.idx_offset : int64_t(-1);
    }
}

// MSD radix select of the k-th largest of M keys (get(j), j < M; 0 = padding), then the <= k winners are
// collected into dest[0 .. n_sort) and sorted best-first.  All threads of the CTA call this.
template <class Get>
__device__ __forceinline__ void radix_select_collect_sort(Get get, const int64_t M, const SelectOut& o, uint64_t* dest) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_cnt;
    __shared__ uint64_t s_prefix, s_mask;
    __shared__ uint32_t s_need, s_done;
    const int tid = threadIdx.x;
    __syncthreads();
    // ---- radix select of the k-th largest key ---------------------------------------------------
    if (tid == 0) { s_prefix = 0; s_mask = 0; s_need = uint32_t(o.k); s_done = 0; s_cnt = 0; }
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int t = tid; t < 256; t += blockDim.x) hist[t] = 0;
        __syncthreads();
        if (s_done) break;
        const uint64_t prefix = s_prefix, mask = s_mask;
        for (int64_t j = tid; j < M; j += blockDim.x) {
            uint64_t key = get(j);
            if ((key & mask) == prefix) atomicAdd(&hist[uint32_t(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid < 32) {
            const uint32_t need = s_need;
            uint32_t c[8], lane_sum = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) { c[b] = hist[tid * 8 + b]; lane_sum += c[b]; }
            uint32_t incl = lane_sum;   // inclusive suffix sum over lanes (towards higher digits)
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                uint32_t t = __shfl_down_sync(0xffffffffu, incl, off);
                if (tid + off < 32) incl += t;
            }
            uint32_t running = incl - lane_sum;
            bool found = false;
            uint32_t d = 0, new_need = 0, bucket = 0;
#pragma unroll
            for (int b = 7; b >= 0; --b) {
                if (!found && running + c[b] >= need) {
                    found = true; d = uint32_t(tid * 8 + b); new_need = need - running; bucket = c[b];
                }
                running += c[b];
            }
            uint32_t ballot = __ballot_sync(0xffffffffu, found);
            int src_lane = ballot ? 31 - __clz(int(ballot)) : 0;
            d = __shfl_sync(0xffffffffu, d, src_lane);
            new_need = __shfl_sync(0xffffffffu, new_need, src_lane);
            bucket = __shfl_sync(0xffffffffu, bucket, src_lane);
            if (tid == 0) {
                s_prefix = prefix | (uint64_t(d) << shift);
                s_mask = mask | (0xffull << shift);
                s_need = new_need;
                if (bucket == new_need || ballot == 0) s_done = 1;   // bucket taken whole
            }
        }
        __syncthreads();
    }
    __syncthreads();

    // ---- collect the winners, sort them, write ----------------------------------------------------
    const uint64_t prefix = s_prefix, mask = s_mask;
    for (int t = tid; t < o.n_sort; t += blockDim.x) dest[t] = 0ull;
    __syncthreads();
    for (int64_t j = tid; j < M; j += blockDim.x) {
        uint64_t key = get(j);
        if (key != 0ull && (key & mask) >= prefix) {
            uint32_t pos = atomicAdd(&s_cnt, 1u);
            if (pos < uint32_t(o.n_sort)) dest[pos] = key;
        }
    }
    __syncthreads();
    block_bitonic_sort_desc(dest, o.n_sort);
}


template <class Src>
__device__ __forceinline__ void select_one_row(const Src& src, const SelectOut& o, const int row, uint64_t* sel) {
    const int64_t M = src.size(row);
    const int tid = threadIdx.x;

    // ---- survivor lists: usually a few thousand live keys in a much larger slot space -> gather
    //      the live ones into shared memory and sort there ------------------------------------
    if constexpr (Src::kIsLists) {
        __shared__ uint32_t offs[kMaxLists + 1];
        __shared__ uint32_t warp_tot[32];
        {   // exclusive prefix sum of the list lengths: blocked over the threads, then warp + block scan
            const int per = (src.n_lists + int(blockDim.x) - 1) / int(blockDim.x);
            const int l0 = tid * per, l1 = min(l0 + per, src.n_lists);
            uint32_t local = 0;
            for (int l = l0; l < l1; ++l) {
                const uint32_t c = src.count(row, l);
                offs[l] = local;                       // exclusive within this thread's block, fixed up below
                local += c;
            }
            uint32_t incl = local;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
                if ((tid & 31) >= off) incl += t;
            }
            if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
            __syncthreads();
            uint32_t base = incl - local;
            for (int w = 0; w < (tid >> 5); ++w) base += warp_tot[w];
            for (int l = l0; l < l1; ++l) offs[l] += base;
            if (tid == int(blockDim.x) - 1) {
                const uint32_t run = base + local;
                offs[src.n_lists] = run;
                if (src.flag_mode == 1) {
                    const int64_t need = src.n_items < int64_t(o.k) ? src.n_items : int64_t(o.k);
                    const bool short_row = int64_t(run) < need;
                    src.thr_shared[row] = float_to_ordered(short_row ? -INFINITY : INFINITY);
                    if (short_row) {
                        atomicOr(src.mtile_flags + row / 128, 1u);
                        src.flagged_rows()[atomicAdd(src.n_flagged(), 1u)] = uint32_t(row);
                    }
                }
            }
        }
        __syncthreads();
        const uint32_t total = offs[src.n_lists];
        if (total <= uint32_t(kSmemSortCap)) {
            int n_pow2 = 2;
            while (n_pow2 < int(total) || n_pow2 < o.k) n_pow2 <<= 1;
            // many more candidates than k: select in shared memory first, sort only the winners
            const bool select_first = n_pow2 > 2 * o.n_sort && total <= uint32_t(kSmemSortCap / 2) && o.n_sort <= kSmemSortCap / 2;
            if (!select_first)
                for (int t = tid + int(total); t < n_pow2; t += blockDim.x) sel[t] = 0ull;
            const int warp = tid >> 5, lane = tid & 31, n_warps = blockDim.x >> 5;
            for (int l = warp; l < src.n_lists; l += n_warps) {
                const uint32_t base = offs[l], cnt = offs[l + 1] - offs[l];
                const uint64_t* p = src.list_ptr(row, l);
                for (uint32_t t = lane; t < cnt; t += 32) sel[base + t] = src.load(p + t);
            }
            if (select_first) {
                uint64_t* dest = sel + kSmemSortCap / 2;
                radix_select_collect_sort([&](int64_t j) { return sel[j]; }, int64_t(total), o, dest);
                write_sorted(dest, row, o);
            } else {
                block_bitonic_sort_desc(sel, n_pow2);
                write_sorted(sel, row, o);
            }
            return;
        }
    }

    // ---- small rows: everything fits in shared memory -> one sort, no radix passes -----------
    if (M <= kSmemSortCap) {
        int n_pow2 = 1;
        while (n_pow2 < M || n_pow2 < o.k) n_pow2 <<= 1;
        if (n_pow2 < 2) n_pow2 = 2;
        for (int t = tid; t < n_pow2; t += blockDim.x) sel[t] = t < M ? src.key(row, t) : 0ull;
        block_bitonic_sort_desc(sel, n_pow2);
        write_sorted(sel, row, o);
        return;
    }

    // ---- large rows: radix select from the source, then sort the winners --------------------------
    radix_select_collect_sort([&](int64_t j) { return src.key(row, j); }, M, o, sel);
    write_sorted(sel, row, o);
}

template <class Src>
__global__ void __launch_bounds__(kSelectThreads) select_topk_kernel(Src src, SelectOut o) {
    extern __shared__ __align__(16) uint64_t sel[];
    if constexpr (Src::kIsLists) {
        if (src.flag_mode == 2) {                      // persistent over the (normally empty) list of flagged rows
            const uint32_t n = __ldcg(src.n_flagged());
            for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
                select_one_row(src, o, int(__ldcg(src.flagged_rows() + i)), sel);
                __syncthreads();
            }
            return;
        }
        if (src.row_list != nullptr) {                 // persistent over the rows the warp kernel passed on
            const uint32_t n = __ldcg(src.row_list);
            for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
                select_one_row(src, o, int(__ldcg(src.row_list + 1 + i)), sel);
                __syncthreads();
            }
            return;
        }
    }
    select_one_row(src, o, int(blockIdx.x), sel);
}

// ---- survivor lists, one WARP per row --------------------------------------------------------------
// The fused kernel leaves a few hundred live keys per row spread over its lists.  A warp gathers them into
// its slice of shared memory (coalesced per list), radix-selects the k-th largest, compacts the winners in
// place, bitonic-sorts them and writes the row -- no block-wide barriers, so the many rows in flight on an SM
// hide each other's memory latency.  Rows with more than kWarpSelCap live keys (streaming-mode overflow,
// adversarial inputs) are appended to a list and handled by the CTA-per-row kernel afterwards.
constexpr int kWarpSelCap = 1024;
constexpr int kWarpSelWarps = 2;        // small CTAs (18 KB smem): they fit beside a resident fused-kernel CTA of another stream

__global__ void __launch_bounds__(kWarpSelWarps * 32)
select_lists_warp_kernel(KeyLists src, SelectOut o, int n_rows, uint32_t* __restrict__ big_rows /* [0] = count, then rows */) {
    extern __shared__ __align__(16) uint64_t wsel_smem[];
    const int warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    uint64_t* keys = wsel_smem + size_t(warp) * kWarpSelCap;
    uint32_t* hist = reinterpret_cast<uint32_t*>(wsel_smem + size_t(kWarpSelWarps) * kWarpSelCap) + warp * 256;
    uint32_t* offs = reinterpret_cast<uint32_t*>(wsel_smem + size_t(kWarpSelWarps) * kWarpSelCap) + kWarpSelWarps * 256 +
                     warp * (src.n_lists + 1);
    const uint32_t n_sort = uint32_t(o.n_sort);

    for (int row = blockIdx.x * kWarpSelWarps + warp; row < n_rows; row += gridDim.x * kWarpSelWarps) {
        // list lengths -> exclusive offsets in shared memory (one parallel read of the counts, then a warp scan)
        for (int l = int(lane); l < src.n_lists; l += 32)
            offs[l + 1] = src.count(row, l);
        __syncwarp();
        uint32_t carry = 0;
        for (int base = 0; base < src.n_lists; base += 32) {
            const int l = base + int(lane);
            const uint32_t v = l < src.n_lists ? offs[l + 1] : 0u;
            uint32_t incl = v;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= uint32_t(off)) incl += t;
            }
            if (l < src.n_lists) offs[l + 1] = carry + incl;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) offs[0] = 0;
        const uint32_t total = carry;
        __syncwarp();
        if (src.flag_mode == 1 && lane == 0) {
            const int64_t need = src.n_items < int64_t(o.k) ? src.n_items : int64_t(o.k);
            const bool short_row = int64_t(total) < need;
            src.thr_shared[row] = float_to_ordered(short_row ? -INFINITY : INFINITY);
            if (short_row) {
                atomicOr(src.mtile_flags + row / 128, 1u);
                src.flagged_rows()[atomicAdd(src.n_flagged(), 1u)] = uint32_t(row);
            }
        }
        if (total > uint32_t(kWarpSelCap)) {
            if (lane == 0) big_rows[1 + atomicAdd(big_rows, 1u)] = uint32_t(row);
            __syncwarp();
            continue;
        }
        // gather: flat element e -> (list, position) by binary search in the offsets; four loads in flight per lane

        for (uint32_t e0 = 0; e0 < total; e0 += 128) {
            uint64_t reg[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t e = e0 + uint32_t(u) * 32u + lane;
                reg[u] = 0ull;
                if (e < total) {
                    int lo = 0, hi = src.n_lists;
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (offs[mid] <= e) lo = mid; else hi = mid;
                    }
                    reg[u] = src.load(src.list_ptr(row, lo) + (e - offs[lo]));
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t e = e0 + uint32_t(u) * 32u + lane;
                if (e < total) keys[e] = reg[u];
            }
        }
        __syncwarp();
        uint32_t n_win = total;
        if (total > uint32_t(o.k)) {
            uint64_t prefix, mask;
            warp_radix_kth(keys, total, uint32_t(o.k), hist, prefix, mask);
            n_win = warp_compact_ge(keys, total, prefix, mask);
        }
        for (uint32_t t = n_win + lane; t < n_sort; t += 32) keys[t] = 0ull;
        __syncwarp();
        warp_bitonic_sort_desc(keys, n_sort);
        const float scale = o.row_scale ? o.row_scale[row] : 1.0f;
        for (int t = int(lane); t < o.k; t += 32) {
            const uint64_t key = keys[t];
            const bool ok = key != 0ull;
            o.vals[int64_t(row) * o.k + t] = ok ? key_score(key) * scale : ANNCUR_PAD_VAL;
            o.idx[int64_t(row) * o.k + t] = ok ? int64_t(key_index(key)) + o.idx_offset : int64_t(-1);
        }
        __syncwarp();
    }
}

static int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }

template <class Src>
static int launch_select(const Src& src, int n_rows, int k, int64_t idx_offset, const float* row_scale,
                         float* out_vals, int64_t* out_idx, cudaStream_t stream, int threads = kSelectThreads,
                         int grid = 0) {
    if (n_rows == 0) return ANNCUR_OK;
    SelectOut o{out_vals, out_idx, idx_offset, row_scale, k, next_pow2(k < 2 ? 2 : k)};
    size_t smem = sizeof(uint64_t) * size_t(o.n_sort > kSmemSortCap ? o.n_sort : kSmemSortCap);
    ANNCUR_CUDA_OK(cudaFuncSetAttribute(select_topk_kernel<Src>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        int(smem)));
    select_topk_kernel<Src><<<grid > 0 ? grid : n_rows, threads, smem, stream>>>(src, o);
    ANNCUR_LAUNCH_OK("select_topk_kernel");
    return ANNCUR_OK;
}

// ---- dense rows, sampled threshold: ~1.06 passes over the row instead of one pass per radix digit -----------------
// Pass A reads one float4 out of every G (a 1/G sample), keeps the maxima of <= kDsGroups contiguous runs of the sample
// and takes their j-th largest as threshold (<= the j-th best sampled value; j = binomial_tail_rank(k-1, 1/G), so at
// least k elements of the row reach it except with probability ~1e-6).  Pass B streams the row once and collects
// the elements >= threshold in shared memory, then selects / sorts them.  A row that collects fewer than
// min(k, n) or more than kDsCap elements falls back to the radix path, so the result never depends on the sample.
constexpr int kDsThreads = 256;
constexpr int kDsGroups = 512;
constexpr int kDsCap = 4096;        // >= kSmemSortCap: the region doubles as the fallback path's scratch
constexpr int kDsStride = 16;          // G

__global__ void __launch_bounds__(kDsThreads)
topk_dense_sampled_kernel(DenseRow src, SelectOut o, int rank_j, int n_sort_j) {
    extern __shared__ __align__(16) uint64_t ds_smem[];
    uint64_t* list = ds_smem;                         // kDsCap keys (also the fallback's scratch)
    uint64_t* grp = ds_smem + kDsCap;                 // kDsGroups keys
    uint64_t* dest = grp + kDsGroups;                 // max(n_sort, n_sort_j) keys
    __shared__ uint32_t s_count;
    __shared__ float s_thr;
    const int row = blockIdx.x;
    const int tid = threadIdx.x;
    const float* rp = src.S + int64_t(row) * src.lds;
    const int64_t n = src.n_cols;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(rp) & 15) == 0;
    const int64_t n4 = vec_ok ? n / 4 : 0;                         // float4 units of the row (tail handled separately)
    const int64_t n_s4 = n4 / kDsStride;                           // sampled float4 units: unit u -> float4 index u * G
    bool fallback = n_s4 < int64_t(4 * rank_j);
    if (!fallback) {
        // ---- pass A: group maxima of the sample ------------------------------------------------------------------
        const int n_groups = int(n_s4 < kDsGroups ? n_s4 : kDsGroups);
        const int64_t per = (n_s4 + n_groups - 1) / n_groups;
        for (int g = tid; g < n_groups; g += kDsThreads) {
            float m = -INFINITY;
            const int64_t u1 = (int64_t(g) + 1) * per < n_s4 ? (int64_t(g) + 1) * per : n_s4;
            for (int64_t u = int64_t(g) * per; u < u1; ++u) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(rp) + u * kDsStride);
                m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
            }
            grp[g] = make_key(m, uint32_t(g));
        }
        SelectOut oj = o;
        oj.k = rank_j;
        oj.n_sort = n_sort_j;
        radix_select_collect_sort([&](int64_t j) { return grp[j]; }, int64_t(n_groups), oj, dest);
        if (tid == 0) { s_thr = key_score(dest[rank_j - 1]); s_count = 0; }
        __syncthreads();
        const float thr = s_thr;
        // ---- pass B: one streaming pass, survivors to shared memory -----------------------------------------------
        auto push = [&](float v, int64_t j) {
            if (v >= thr) {
                const uint32_t pos = atomicAdd(&s_count, 1u);
                if (pos < uint32_t(kDsCap)) list[pos] = make_key(v, uint32_t(j));
            }
        };
        constexpr int UN = 8;                         // independent 16-byte loads in flight per thread
        for (int64_t u0 = tid; u0 < n4; u0 += int64_t(kDsThreads) * UN) {
            float4 v[UN];
#pragma unroll
            for (int i = 0; i < UN; ++i) {
                const int64_t u = u0 + int64_t(i) * kDsThreads;
                v[i] = u < n4 ? __ldg(reinterpret_cast<const float4*>(rp) + u) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            }
#pragma unroll
            for (int i = 0; i < UN; ++i) {
                const int64_t u = u0 + int64_t(i) * kDsThreads;
                if (fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)) >= thr && u < n4) {
                    push(v[i].x, 4 * u); push(v[i].y, 4 * u + 1); push(v[i].z, 4 * u + 2); push(v[i].w, 4 * u + 3);
                }
            }
        }
        for (int64_t j = 4 * n4 + tid; j < n; j += kDsThreads) push(__ldg(rp + j), j);
        __syncthreads();
        const uint32_t cnt = s_count;
        const int64_t need = n < int64_t(o.k) ? n : int64_t(o.k);
        fallback = cnt > uint32_t(kDsCap) || int64_t(cnt) < need;
        if (!fallback) {
            radix_select_collect_sort([&](int64_t j) { return list[j]; }, int64_t(cnt), o, dest);
            write_sorted(dest, row, o);
            return;
        }
    }
    __syncthreads();
    select_one_row(src, o, row, ds_smem);           // exact radix path (reads the row once per digit)
}

int select_topk_dense(const float* S, int64_t lds, int n_rows, int64_t n_cols, int k, int64_t idx_offset,
                      float* out_vals, int64_t* out_idx, cudaStream_t stream) {
    if (n_rows == 0) return ANNCUR_OK;
    if (n_cols >= 32768 && k <= 2048) {
        const int rank_j = binomial_tail_rank(k - 1, 1.0 / kDsStride, 1e-6);
        if (4 * rank_j <= kDsGroups) {
            SelectOut o{out_vals, out_idx, idx_offset, nullptr, k, next_pow2(k < 2 ? 2 : k)};
            const int n_sort_j = next_pow2(rank_j < 2 ? 2 : rank_j);
            const int n_dest = o.n_sort > n_sort_j ? o.n_sort : n_sort_j;
            const size_t smem = sizeof(uint64_t) * size_t(kDsCap + kDsGroups + n_dest);
            ANNCUR_CUDA_OK(cudaFuncSetAttribute(topk_dense_sampled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
            topk_dense_sampled_kernel<<<n_rows, kDsThreads, smem, stream>>>(DenseRow{S, lds, n_cols}, o, rank_j, n_sort_j);
            ANNCUR_LAUNCH_OK("topk_dense_sampled_kernel");
            return ANNCUR_OK;
        }
    }
    return launch_select(DenseRow{S, lds, n_cols}, n_rows, k, idx_offset, nullptr, out_vals, out_idx, stream);
}

int select_topk_keylists(const uint64_t* keys, const uint32_t* counts, int n_lists, int cap, int n_rows, int k,
                         int64_t idx_offset, const float* row_scale, float* out_vals, int64_t* out_idx,
                         int flag_mode, uint32_t* thr_shared, uint32_t* mtile_flags, int m_tiles, int64_t n_items,
                         uint32_t* big_rows, cudaStream_t stream) {
    if (n_lists > kMaxLists) { set_error("select_topk_keylists: %d lists per row > %d", n_lists, kMaxLists); return ANNCUR_E_UNSUPPORTED; }
    if (n_rows == 0) return ANNCUR_OK;
    KeyLists src{keys, counts, n_lists, cap, flag_mode, thr_shared, mtile_flags, m_tiles, n_items, nullptr,
                 int64_t(cap), int64_t(n_lists) * cap, 1};
    if (flag_mode == 2) {
        // REDO select: persistent CTA kernel over the (normally empty) list of flagged rows
        const int grid = n_rows < 2 * sm_count() ? n_rows : 2 * sm_count();
        return launch_select(src, n_rows, k, idx_offset, row_scale, out_vals, out_idx, stream, 128, grid);
    }
    // one warp per row; rows with too many live keys are passed on to the CTA kernel through big_rows
    SelectOut o{out_vals, out_idx, idx_offset, row_scale, k, next_pow2(k < 2 ? 2 : k)};
    ANNCUR_CUDA_OK(cudaMemsetAsync(big_rows, 0, sizeof(uint32_t), stream));
    const size_t smem = size_t(kWarpSelWarps) * (kWarpSelCap * sizeof(uint64_t) + (256 + size_t(n_lists) + 1) * sizeof(uint32_t));
    ANNCUR_CUDA_OK(cudaFuncSetAttribute(select_lists_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int grid = (n_rows + kWarpSelWarps - 1) / kWarpSelWarps;
    if (grid > 12 * sm_count()) grid = 12 * sm_count();
    select_lists_warp_kernel<<<grid, kWarpSelWarps * 32, smem, stream>>>(src, o, n_rows, big_rows);
    ANNCUR_LAUNCH_OK("select_lists_warp_kernel");
    src.flag_mode = 0;                                  // flags were written by the warp kernel
    src.row_list = big_rows;
    const int grid2 = n_rows < 2 * sm_count() ? n_rows : 2 * sm_count();
    return launch_select(src, n_rows, k, idx_offset, row_scale, out_vals, out_idx, stream, 128, grid2);
}

// (val, idx) top-k lists -> 64-bit keys (padding idx < 0 -> key 0): the exchange format of the item-sharded search
__global__ void topk_to_keys_kernel(const float* __restrict__ vals, const int64_t* __restrict__ idx, int64_t n, uint64_t* __restrict__ keys) {
    for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < n; t += int64_t(gridDim.x) * blockDim.x) {
        const int64_t i = idx[t];
        keys[t] = i < 0 ? 0ull : make_key(vals[t], uint32_t(i));
    }
}

int topk_to_keys(const float* vals, const int64_t* idx, int n_rows, int k, uint64_t* keys, cudaStream_t stream) {
    const int64_t n = int64_t(n_rows) * k;
    if (n == 0) return ANNCUR_OK;
    topk_to_keys_kernel<<<unsigned((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096), 256, 0, stream>>>(vals, idx, n, keys);
    ANNCUR_LAUNCH_OK("topk_to_keys_kernel");
    return ANNCUR_OK;
}

// Best k_out of n_shards sorted key lists per row, keys laid out [shard][row][k_in] (the all-gather's output as is).
int merge_topk_keys(const uint64_t* keys, int n_shards, int n_rows, int k_in, int k_out, float* out_vals, int64_t* out_idx,
                    uint32_t* scratch_rows, cudaStream_t stream) {
    return merge_topk_keys_strided(keys, n_shards, n_rows, k_in, int64_t(n_rows) * k_in, k_out, out_vals, out_idx, scratch_rows, stream);
}

// The same with an explicit distance (in keys) between the lists of consecutive shards: the receive buffer of the peer
// exchange is [shard][rows_cap][k_in] with rows_cap >= n_rows (peer_exchange.cu).
int merge_topk_keys_strided(const uint64_t* keys, int n_shards, int n_rows, int k_in, int64_t shard_stride, int k_out,
                            float* out_vals, int64_t* out_idx, uint32_t* scratch_rows, cudaStream_t stream) {
    if (n_rows == 0) return ANNCUR_OK;
    if (n_shards > kMaxLists) { set_error("merge_topk_keys: %d shards > %d", n_shards, kMaxLists); return ANNCUR_E_UNSUPPORTED; }
    KeyLists src{keys, nullptr, n_shards, k_in, 0, nullptr, nullptr, 0, 0, nullptr, shard_stride, int64_t(k_in), 0};
    SelectOut o{out_vals, out_idx, 0, nullptr, k_out, next_pow2(k_out < 2 ? 2 : k_out)};
    ANNCUR_CUDA_OK(cudaMemsetAsync(scratch_rows, 0, sizeof(uint32_t), stream));
    const size_t smem = size_t(kWarpSelWarps) * (kWarpSelCap * sizeof(uint64_t) + (256 + size_t(n_shards) + 1) * sizeof(uint32_t));
    ANNCUR_CUDA_OK(cudaFuncSetAttribute(select_lists_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int grid = (n_rows + kWarpSelWarps - 1) / kWarpSelWarps;
    if (grid > 12 * sm_count()) grid = 12 * sm_count();
    select_lists_warp_kernel<<<grid, kWarpSelWarps * 32, smem, stream>>>(src, o, n_rows, scratch_rows);
    ANNCUR_LAUNCH_OK("select_lists_warp_kernel");
    src.row_list = scratch_rows;                        // rows with more than kWarpSelCap candidates (n_shards * k_in large)
    const int grid2 = n_rows < 2 * sm_count() ? n_rows : 2 * sm_count();
    return launch_select(src, n_rows, k_out, 0, nullptr, out_vals, out_idx, stream, 128, grid2);
}

int select_topk_pairs(const float* vals, const int64_t* idx, int n_rows, int n_cand, int k, float* out_vals,
                      int64_t* out_idx, cudaStream_t stream) {
    return launch_select(PairLists{vals, idx, n_cand}, n_rows, k, 0, nullptr, out_vals, out_idx, stream);
}

}  // namespace anncur
