// K5 + K6: exact-score rerank of the retrieved items and overlap with the exact top-k.
//
// Replaces, per query, the reference's  temp = zeros(N) - 1e14; temp[idx] = exact[idx]; temp.topk(k)
// (eval/run_retrieval_eval_wrt_exact_crossenc.py:108-113, ..._w_fixed_train_test_splits.py:91-96)
// without the N-long temporary: gather k_retr exact scores, sort them in shared memory, and count
// |exact[:k] & reranked[:k]| (eval/eval_utils.py:139-150) for every requested k in one pass.
// One CTA per query row.
#include "common.cuh"
#include "kernels.h"

namespace anncur {

constexpr int RR_THREADS = 256;
constexpr int RR_MAX_K_LIST = 16;

struct KList {
    int n;
    int k[RR_MAX_K_LIST];
};

__global__ void __launch_bounds__(RR_THREADS)
rerank_overlap_kernel(const float* __restrict__ exact, int64_t lds, int64_t n_cols,
                      const int64_t* __restrict__ retr_idx, int k_retr, int n_pow2,
                      const int64_t* __restrict__ exact_idx, int k_max, KList kl,
                      int64_t* __restrict__ out_rr_idx, float* __restrict__ out_rr_vals,
                      int32_t* __restrict__ out_common) {
    extern __shared__ __align__(16) uint64_t keys[];          // n_pow2 keys, then k_max ints (positions)
    int* pos = reinterpret_cast<int*>(keys + n_pow2);
    __shared__ int red[RR_THREADS / 32];
    const int row = blockIdx.x, tid = threadIdx.x;
    const float* erow = exact + int64_t(row) * lds;

    for (int t = tid; t < n_pow2; t += RR_THREADS) {
        uint64_t key = 0ull;
        if (t < k_retr) {
            int64_t i = retr_idx[int64_t(row) * k_retr + t];
            if (i >= 0 && i < n_cols) key = make_key(__ldg(erow + i), uint32_t(i));
        }
        keys[t] = key;
    }
    block_bitonic_sort_desc(keys, n_pow2);

    for (int t = tid; t < k_max; t += RR_THREADS) {
        uint64_t key = keys[t];
        bool ok = key != 0ull;
        out_rr_idx[int64_t(row) * k_max + t] = ok ? int64_t(key_index(key)) : int64_t(-1);
        out_rr_vals[int64_t(row) * k_max + t] = ok ? key_score(key) : ANNCUR_PAD_VAL;
        pos[t] = 0x7fffffff;
    }
    __syncthreads();
    // pos[t] = rank of reranked item t inside the exact top-k_max list (or "absent")
    const int64_t* ex = exact_idx + int64_t(row) * k_max;
    for (int64_t pair = tid; pair < int64_t(k_max) * k_max; pair += RR_THREADS) {
        int t = int(pair / k_max), e = int(pair % k_max);
        uint64_t key = keys[t];
        if (key != 0ull && int64_t(key_index(key)) == __ldg(ex + e)) pos[t] = e;
    }
    __syncthreads();
    for (int q = 0; q < kl.n; ++q) {
        const int k = kl.k[q];
        int c = 0;
        for (int t = tid; t < k; t += RR_THREADS) c += (pos[t] < k) ? 1 : 0;
        c = warp_sum(c);
        if ((tid & 31) == 0) red[tid >> 5] = c;
        __syncthreads();
        if (tid == 0) {
            int s = 0;
            for (int w = 0; w < RR_THREADS / 32; ++w) s += red[w];
            out_common[int64_t(row) * kl.n + q] = s;
        }
        __syncthreads();
    }
}

int rerank_overlap(const float* exact, int64_t lds, int n_rows, int64_t n_cols, const int64_t* retr_idx,
                   int k_retr, const int64_t* exact_idx, int k_max, const int* k_list_host, int n_k,
                   int64_t* out_rr_idx, float* out_rr_vals, int32_t* out_common, cudaStream_t stream) {
    if (n_rows <= 0) return ANNCUR_OK;
    if (k_retr < 1 || k_retr > ANNCUR_MAX_K || k_max < 1 || k_max > k_retr) {
        set_error("rerank_overlap: need 1 <= k_max (%d) <= k_retr (%d) <= %d", k_max, k_retr, ANNCUR_MAX_K);
        return ANNCUR_E_INVALID;
    }
    if (n_k < 1 || n_k > RR_MAX_K_LIST) {
        set_error("rerank_overlap: n_k = %d outside [1, %d]", n_k, RR_MAX_K_LIST);
        return ANNCUR_E_INVALID;
    }
    KList kl;
    kl.n = n_k;
    for (int q = 0; q < n_k; ++q) {
        if (k_list_host[q] < 1 || k_list_host[q] > k_max) {
            set_error("rerank_overlap: k_list[%d] = %d outside [1, k_max = %d]", q, k_list_host[q], k_max);
            return ANNCUR_E_INVALID;
        }
        kl.k[q] = k_list_host[q];
    }
    int n_pow2 = 2;
    while (n_pow2 < k_retr) n_pow2 <<= 1;
    size_t smem = sizeof(uint64_t) * size_t(n_pow2) + sizeof(int) * size_t(k_max);
    rerank_overlap_kernel<<<n_rows, RR_THREADS, smem, stream>>>(exact, lds, n_cols, retr_idx, k_retr, n_pow2,
                                                                exact_idx, k_max, kl, out_rr_idx, out_rr_vals,
                                                                out_common);
    ANNCUR_LAUNCH_OK("rerank_overlap_kernel");
    return ANNCUR_OK;
}

// K6 on its own: |set(a[row]) & set(b[row])| for two index lists per row (eval/eval_utils.py:139-150,
// len(set(indices1).intersection(set(indices2)))).  One CTA per row: b is sorted in shared memory, duplicates inside a
// row are counted once on either side (set semantics), every distinct a is looked up by binary search.
__global__ void __launch_bounds__(RR_THREADS)
overlap_counts_kernel(const int64_t* __restrict__ a, const int64_t* __restrict__ b, int k, int n_pow2, int32_t* __restrict__ out) {
    extern __shared__ __align__(16) uint64_t sb[];            // n_pow2 sorted b keys, then n_pow2 sorted a keys
    uint64_t* sa = sb + n_pow2;
    __shared__ int red[RR_THREADS / 32];
    const int row = blockIdx.x, tid = threadIdx.x;
    // indices are mapped to unsigned keys that keep their order (offset by 2^63) so that -1 padding stays valid input
    for (int t = tid; t < n_pow2; t += RR_THREADS) {
        sb[t] = t < k ? uint64_t(b[int64_t(row) * k + t]) ^ 0x8000000000000000ull : 0ull;
        sa[t] = t < k ? uint64_t(a[int64_t(row) * k + t]) ^ 0x8000000000000000ull : 0ull;
    }
    block_bitonic_sort_desc(sb, n_pow2);
    block_bitonic_sort_desc(sa, n_pow2);                       // descending; real entries first (padding 0 is the minimum)
    int c = 0;
    for (int t = tid; t < k; t += RR_THREADS) {
        const uint64_t key = sa[t];
        if (t > 0 && sa[t - 1] == key) continue;               // duplicate inside a
        int lo = 0, hi = k;                                    // first position in sb[0..k) with sb[pos] <= key (descending)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (sb[mid] > key) lo = mid + 1; else hi = mid;
        }
        c += (lo < k && sb[lo] == key) ? 1 : 0;
    }
    c = warp_sum(c);
    if ((tid & 31) == 0) red[tid >> 5] = c;
    __syncthreads();
    if (tid == 0) {
        int s = 0;
        for (int w = 0; w < RR_THREADS / 32; ++w) s += red[w];
        out[row] = s;
    }
}

int overlap_counts(const int64_t* a, const int64_t* b, int n_rows, int k, int32_t* out_common, cudaStream_t stream) {
    if (n_rows <= 0) return ANNCUR_OK;
    if (k < 1 || k > 4096) { set_error("overlap_counts: k = %d outside [1, 4096]", k); return ANNCUR_E_INVALID; }
    int n_pow2 = 2;
    while (n_pow2 < k) n_pow2 <<= 1;
    const size_t smem = sizeof(uint64_t) * 2 * size_t(n_pow2);
    ANNCUR_CUDA_OK(cudaFuncSetAttribute(overlap_counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    overlap_counts_kernel<<<n_rows, RR_THREADS, smem, stream>>>(a, b, k, n_pow2, out_common);
    ANNCUR_LAUNCH_OK("overlap_counts_kernel");
    return ANNCUR_OK;
}

}  // namespace anncur
