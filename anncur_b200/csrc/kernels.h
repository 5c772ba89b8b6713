// Internal C++ entry points of the kernel translation units (wrapped by capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace anncur {

// select_topk.cu
int select_topk_dense(const float* S, int64_t lds, int n_rows, int64_t n_cols, int k, int64_t idx_offset,
                      float* out_vals, int64_t* out_idx, cudaStream_t stream);
int select_topk_keylists(const uint64_t* keys, const uint32_t* counts, int n_lists, int cap, int n_rows, int k,
                         int64_t idx_offset, const float* row_scale, float* out_vals, int64_t* out_idx,
                         int flag_mode, uint32_t* thr_shared, uint32_t* mtile_flags, int m_tiles, int64_t n_items,
                         uint32_t* big_rows, cudaStream_t stream);
int topk_to_keys(const float* vals, const int64_t* idx, int n_rows, int k, uint64_t* keys, cudaStream_t stream);
int merge_topk_keys(const uint64_t* keys, int n_shards, int n_rows, int k_in, int k_out, float* out_vals, int64_t* out_idx,
                    uint32_t* scratch_rows, cudaStream_t stream);
int merge_topk_keys_strided(const uint64_t* keys, int n_shards, int n_rows, int k_in, int64_t shard_stride, int k_out,
                            float* out_vals, int64_t* out_idx, uint32_t* scratch_rows, cudaStream_t stream);
int select_topk_pairs(const float* vals, const int64_t* idx, int n_rows, int n_cand, int k, float* out_vals,
                      int64_t* out_idx, cudaStream_t stream);

// refine_topk.cu
int refine_topk_keylists(const uint64_t* cand, const uint32_t* counts, int n_lists, int cap, int n_rows, int k,
                         int64_t idx_offset, const float* Q, int ldq, int k_dim, const float* ET, int ld,
                         const float* row_inv_scale, float* out_vals, int64_t* out_idx, uint32_t* thr_shared,
                         uint32_t* mtile_flags, int m_tiles, int64_t n_items, int row_cap, cudaStream_t stream);

// sgemm.cu
int sgemm_rowmajor(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int m,
                   int64_t n, int k, cudaStream_t stream);
int recon_error(const float* Q, int64_t ldq, const float* E, int64_t lde, const float* A, int64_t lda, int n_rows,
                int64_t n_items, int k_dim, double* out_err2, double* out_norm2, cudaStream_t stream);

// pinv.cu
size_t pinv_workspace_bytes(int m, int n);
int pinv_f32(const float* A, int m, int n, int lda, double rcond, float* out, int ldo, double* cond_out,
             void* workspace, size_t workspace_bytes, cudaStream_t stream);
int orthonormalize_f32(const float* A, int m, int n, int lda, float* Q, int ldq, double* sigma_out, void* workspace,
                       size_t workspace_bytes, cudaStream_t stream);
int jacobi_status(const void* workspace, double* status4_out, cudaStream_t stream);
int singular_values_f32(const float* A, int m, int n, int lda, double* sigma_out, void* workspace, size_t workspace_bytes,
                        cudaStream_t stream);

// rerank.cu
int rerank_overlap(const float* exact, int64_t lds, int n_rows, int64_t n_cols, const int64_t* retr_idx,
                   int k_retr, const int64_t* exact_idx, int k_max, const int* k_list_host, int n_k,
                   int64_t* out_rr_idx, float* out_rr_vals, int32_t* out_common, cudaStream_t stream);
int overlap_counts(const int64_t* a, const int64_t* b, int n_rows, int k, int32_t* out_common, cudaStream_t stream);

// score_topk_umma.cu (tcgen05 / TMEM / TMA)
size_t packed_items_bytes(int64_t n_items, int k_dim, int kind);
int pack_items(const float* E, int64_t lde, int64_t n_items, int k_dim, int kind, void* packed,
               float* e_scale_out, cudaStream_t stream);
size_t score_topk_workspace_bytes(int n_queries, int64_t n_items, int k_dim, int k, int kind);
int score_topk_fused(const float* Q, int ldq, int n_queries, const void* packed_items, const float* e_scale,
                     int64_t n_items, int k_dim, int kind, int k, int64_t idx_offset, float* out_vals,
                     int64_t* out_idx, void* workspace, size_t workspace_bytes, cudaStream_t stream);

size_t score_dense_workspace_bytes(int n_queries, int64_t n_items, int k_dim, int kind);
int score_dense(const float* Q, int ldq, int n_queries, const void* packed_items, const float* e_scale, int64_t n_items,
                int k_dim, int kind, float* out, int64_t ldo, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int score_bounds_dense(const float* Q, int ldq, int n_queries, const void* packed_items, const float* e_scale, int64_t n_items,
                       int k_dim, int sign, float* out, int64_t ldo, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int recon_error_packed(const float* Q, int ldq, int n_queries, const void* packed_items, const float* e_scale, int64_t n_items,
                       int k_dim, int kind, const float* A, int64_t lda, double* out_err2, double* out_norm2, void* workspace,
                       size_t workspace_bytes, cudaStream_t stream);
int score_topk_redo_rows(const void* workspace, int n_queries, int64_t n_items, int k_dim, int k, int kind, int* redo_rows_host,
                         cudaStream_t stream);

// peer_exchange.cu (candidate exchange of the item-sharded search over NVLink peer memory)
size_t peer_channel_bytes(int world, int rows_cap, int k_cap);
int peer_alloc(size_t bytes, void** out);
int peer_free(void* p);
int peer_export(const void* base, void* handle64);
int peer_open(const void* handle64, void** out);
int peer_close(void* mapped);
int peer_scatter_keys(const float* vals, const int64_t* idx, int n_rows, int k, int rank, int world, int rows_cap, int k_cap,
                      uint32_t epoch, void* const* peer_bases, cudaStream_t stream);
int peer_merge_owned(void* local_base, int rank, int world, int rows_owned, int rows_cap, int k_cap, int k_out, uint32_t epoch,
                     float* out_vals, int64_t* out_idx, uint32_t* scratch_rows, cudaStream_t stream);
int peer_cert_failures(void* local_base, int world, int rows_cap, int k_cap, int reset, unsigned* count_host, cudaStream_t stream);
int peer_error(void* local_base, int world, int rows_cap, int k_cap, int* err_host, cudaStream_t stream);

// smallest j with P[Binomial(n, p) >= j] <= eps (rank used by the sampled thresholds)
int binomial_tail_rank(int n, double p, double eps);

int profile_enable(int on);
int profile_read(double* ms_sum, int* launches);

// adaptive.cu
size_t adaptive_round_workspace_bytes(int n_queries, int k_q, int m, int64_t n_items, int n_next);
int adaptive_round(const float* R_anc, int64_t ldr, int k_q, int64_t n_items, const int64_t* anchors,
                   const float* c, int n_queries, int m, double rcond, int n_next, int64_t* next_idx,
                   float* next_val, void* workspace, size_t workspace_bytes, cudaStream_t stream);

size_t adaptive_solve_workspace_bytes(int n_queries, int k_q, int m, int64_t n_items);
int adaptive_solve(const float* R_anc, int64_t ldr, int k_q, int64_t n_items, const float* Rt_cached, const int64_t* anchors,
                   const float* c, int n_queries, int m, double rcond, float* e_out, void* workspace, size_t workspace_bytes,
                   cudaStream_t stream);
int filter_excluded(const float* cand_vals, const int64_t* cand_idx, int n_rows, int k_in, const int64_t* excl, int m, int n_out,
                    float* out_vals, int64_t* out_idx, cudaStream_t stream);
int transpose_rows(const float* in, int64_t ld_in, int rows, int64_t cols, float* out, cudaStream_t stream);

// adaptive_inc.cu
size_t adaptive_shared_bytes(int k_q, int64_t n_items, int m_shared);
size_t adaptive_prepare_workspace_bytes(int k_q, int64_t n_items, int m_shared);
int adaptive_prepare(const float* Rt, int k_q, int64_t n_items, const int64_t* shared_anchors, int s, double rcond,
                     void* shared, size_t shared_bytes, void* workspace, size_t workspace_bytes, cudaStream_t stream);
size_t adaptive_state_bytes(int n_queries, int k_q, int m_shared, int n_new, int m_max);
int adaptive_begin(const float* Rt, int k_q, int64_t n_items, const void* shared, int s, const float* c, int n_queries, int n_new,
                   int m_max, float* e_out, void* state, size_t state_bytes, cudaStream_t stream);
int adaptive_extend(const float* Rt, int k_q, int64_t n_items, const void* shared, int s, const int64_t* new_anchors,
                    const float* c_new, int n_queries, int n, int m_max, int m_cur, double rcond, float* e_out, void* state,
                    size_t state_bytes, cudaStream_t stream);

}  // namespace anncur
