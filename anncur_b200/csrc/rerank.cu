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

}  // namespace anncur
