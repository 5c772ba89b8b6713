/*
 * anncur_b200.h -- C ABI of libanncur_b200.so: the B200 (sm_100a) engine behind the ANNCUR
 * test-time search path of iesl/anncur.
 *
 * The reference has no FFI of its own: its boundary is the Python call surface of
 *   eval/matrix_approx_zeshel.py   (class CURApprox, :19-126)
 *   models/nearest_nbr.py          (build_flat_or_ivff_index(...).search, :24-55)
 *   eval/run_retrieval_eval_wrt_exact_crossenc*.py (per-query retrieve/rerank/overlap loops)
 * Each entry point below names the reference lines whose arithmetic it replaces.  The host-side
 * mirror of those Python surfaces lives in the anncur_b200 package (its .py files) and binds this ABI with ctypes
 * (INTEGRATION.md shows the binding).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in _host; all matrices are row-major,
 *    dense unless a leading dimension (ld*, in elements) is given;
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream and
 *    never synchronises the device;
 *  - the library never allocates device memory: scratch comes from the caller, sized by the
 *    matching *_workspace_bytes() query (workspaces must be 256-byte aligned);
 *  - return value 0 = success, negative = ANNCUR_E_* ; anncur_last_error() returns a thread-local
 *    message for the last failure on this host thread;
 *  - top-k results are best-first; ties are broken towards the LOWER item index; indices are
 *    int64 like torch.topk / faiss; unused slots (k larger than the candidate set) hold
 *    idx = -1, val = -FLT_MAX (faiss' inner-product padding).
 */
#ifndef ANNCUR_B200_H_
#define ANNCUR_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define ANNCUR_API __attribute__((visibility("default")))
#else
#define ANNCUR_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define ANNCUR_ABI_VERSION 1

#define ANNCUR_OK            0
#define ANNCUR_E_INVALID    -1  /* bad argument (shape, k, null pointer, alignment)            */
#define ANNCUR_E_WORKSPACE  -2  /* workspace too small                                          */
#define ANNCUR_E_CUDA       -3  /* a CUDA runtime / driver call failed (message has the code)   */
#define ANNCUR_E_UNSUPPORTED -4 /* shape outside what this kernel family handles                */

/* precision of the fused score + top-k tensor-core path */
#define ANNCUR_KIND_F32X3 0  /* fp32-grade: operands split into two fp16 terms, 3 tcgen05 passes */
#define ANNCUR_KIND_BF16  1  /* single bf16 pass (reported separately as recall@k)             */
#define ANNCUR_KIND_F32R  2  /* fp32 results at the one-pass rate: ONE f16 tcgen05 pass computes a rigorous upper bound of
                                every score (per-item error bound folded into the contraction), the ~k items whose bound
                                reaches the k-th best score are re-scored in plain fp32 (FFMA) from an item-major fp32 copy
                                of E kept in the packed index (8 instead of 4 bytes per element of E) */
#define ANNCUR_MAX_K_DIM_F32R 8192  /* largest k_dim of kind F32R                                */

#define ANNCUR_MAX_K        2048  /* largest k of any top-k entry point                         */
#define ANNCUR_MAX_K_FUSED  1024  /* largest k of the fused tensor-core path                    */

ANNCUR_API int         anncur_abi_version(void);
ANNCUR_API const char* anncur_last_error(void);

/* ---- K1: pseudo-inverse -------------------------------------------------------------------
 * Replaces np.linalg.pinv at eval/matrix_approx_zeshel.py:47,49 (U = pinv(C[row_idxs,:])).
 * A is m x n fp32 (lda), out is n x m fp32 (ldo).  One-sided Jacobi SVD in fp64; singular values
 * <= max(rcond, 8 tol) * s_max are dropped (numpy's rule with a floor at what fp64 sweeps resolve, see
 * anncur_jacobi_status; the reference uses numpy's default rcond = 1e-15).
 * cond_out (optional, device double[2]) receives {s_max, s_min_kept}.
 * Route: with cond_out == NULL and rcond <= 1e-10 a full-rank, well-conditioned input is served by the fp64 normal equations
 * (Gram, cooperative blocked Cholesky, two triangular solves per column; ~25x faster than the SVD); a pivot breakdown or a
 * diagonal ratio of the factor above 3e3 (a lower bound of the condition number) falls through to the Jacobi SVD.  The
 * decision is taken on the device (kernels of the route not taken return at once): the call stays asynchronous.
 * ANNCUR_PINV_JACOBI_ONLY=1 in the environment disables the first route. */
ANNCUR_API size_t anncur_pinv_workspace_bytes(int m, int n);
ANNCUR_API int anncur_pinv_f32(const float* A, int m, int n, int lda, double rcond, float* out, int ldo,
                    double* cond_out, void* workspace, size_t workspace_bytes, void* stream);

/* Singular values of A (m x n fp32) as min(m, n) UNSORTED fp64 numbers (device): the same Jacobi factorisation
 * without the inverse.  Replaces the SVD inside np.linalg.matrix_rank at eval/compute_m2e_matrix_ranks.py:44-53
 * (rank = #{sigma > max(m, n) * eps_fp32 * sigma_max}, counted by the caller).  Workspace: anncur_pinv_workspace_bytes. */
ANNCUR_API int anncur_singular_values_f32(const float* A, int m, int n, int lda, double* sigma_out,
                               void* workspace, size_t workspace_bytes, void* stream);

/* Orthonormal basis of the column space of a TALL matrix A (m x n fp32, m >= n): Q (m x n fp32, ldq) = its left singular
 * vectors (unsorted; zero columns where A is rank-deficient), sigma_out (optional, device double[n]) the matching
 * singular values.  The range-finder step of the randomised rank analysis that replaces np.linalg.matrix_rank
 * (eval/compute_m2e_matrix_ranks.py:44-53) where a full SVD is out of reach (BASELINE configs[4]: 10k x 1M).
 * Workspace: anncur_pinv_workspace_bytes(m, n). */
ANNCUR_API int anncur_orthonormalize_f32(const float* A, int m, int n, int lda, float* Q, int ldq, double* sigma_out,
                              void* workspace, size_t workspace_bytes, void* stream);

/* Outcome of the last anncur_pinv_f32 / anncur_singular_values_f32 that ran on `workspace` (enqueued on `stream`):
 * status4_out (device double[4]) = {s_max, s_min_kept (pinv only), converged (1 = every column pair orthogonal to the
 * tolerance within the sweep limit, 0 = the factorisation stopped at the limit and the result must not be trusted),
 * sweeps done}.  A rank-deficient input is not an error: singular values below max(rcond, 8 tol) * s_max are dropped,
 * tol = the sweeps' orthogonality tolerance (~4 sqrt(len) eps_fp64) -- fp64 noise values are never inverted. */
ANNCUR_API int anncur_jacobi_status(const void* workspace, double* status4_out, void* stream);

/* ---- K2 / dense products -------------------------------------------------------------------
 * C[m x n] = A[m x k] . B[k x n], fp32 FFMA accumulate.  Replaces `U @ R`
 * (eval/matrix_approx_zeshel.py:65), `C @ U` (:61) and the dense getters get_rows/get_cols/get/
 * get_complete_row/get_complete_col (:71-119) when the caller wants the full matrix. */
ANNCUR_API int anncur_gemm_f32(const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                    int m, int n, int k, void* stream);

/* ---- item-embedding index for the tensor-core path ------------------------------------------
 * Packs E = latent_cols (k_i x N fp32, row-major, N contiguous as the reference stores it,
 * eval/matrix_approx_zeshel.py:65) into the k-block-major fp16-pair / bf16 planes the fused
 * kernel streams with TMA.  `packed` must hold anncur_packed_items_bytes().  e_scale_out
 * (device float[1]) receives the power-of-two scale applied before the fp16 split. */
ANNCUR_API size_t anncur_packed_items_bytes(int64_t n_items, int k_dim, int kind);
ANNCUR_API int anncur_pack_items(const float* E, int64_t lde, int64_t n_items, int k_dim, int kind,
                      void* packed, float* e_scale_out, void* stream);

/* ---- K3+K4: fused approximate score + per-row top-k -----------------------------------------
 * Replaces CURApprox.topk_in_row = torch.topk(sparse_rows @ latent_cols, k, dim=1)
 * (eval/matrix_approx_zeshel.py:109-126) and IndexFlatIP.search (models/nearest_nbr.py:36-38):
 * the B x N score matrix is never written.  Q is B x k_dim fp32 (ldq).  out_idx = local item
 * index + idx_offset (the shard's first global item).  Requires 1 <= k <=
 * ANNCUR_MAX_K_FUSED. */
ANNCUR_API size_t anncur_score_topk_workspace_bytes(int n_queries, int64_t n_items, int k_dim, int k, int kind);
ANNCUR_API int anncur_score_topk(const float* Q, int ldq, int n_queries, const void* packed_items,
                      const float* e_scale, int64_t n_items, int k_dim, int kind, int k,
                      int64_t idx_offset, float* out_vals, int64_t* out_idx,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Masked search (the "excluded-index list per row" of the adaptive rounds): out[b] = the k best items of
 * row b that are NOT among excluded[b][0 .. m_excl) (GLOBAL indices, i.e. including idx_offset; order free; entries < 0 are
 * ignored), best first, padded with (-FLT_MAX, -1).  Exact: the fused search runs for k + m_excl candidates -- the excluded
 * items of a row can displace at most m_excl of them -- and anncur_filter_excluded keeps the first k that survive.
 * k + m_excl <= ANNCUR_MAX_K_FUSED. */
ANNCUR_API size_t anncur_score_topk_excluding_workspace_bytes(int n_queries, int64_t n_items, int k_dim, int k, int m_excl, int kind);
ANNCUR_API int anncur_score_topk_excluding(const float* Q, int ldq, int n_queries, const void* packed_items, const float* e_scale,
                                int64_t n_items, int k_dim, int kind, int k, const int64_t* excluded, int m_excl,
                                int64_t idx_offset, float* out_vals, int64_t* out_idx, void* workspace, size_t workspace_bytes,
                                void* stream);

/* ---- dense products on the tensor-core pipeline ---------------------------------------------------
 * out[B x N] = Q . E with the fp32-grade 3-pass arithmetic on a packed index of kind F32X3 or F32R: the tcgen05 form of
 * CURApprox.get_complete_row / get_rows / get (eval/matrix_approx_zeshel.py:71-119) and of the item-embedding build
 * latent_cols = U @ R (:65 -- pack R with anncur_pack_items, pass U as Q).  ~10x the FFMA rate of anncur_gemm_f32;
 * operands carry 22-23 significant bits (relative error ~1e-6 of sum_i |q_i e_in|). */
ANNCUR_API size_t anncur_score_dense_workspace_bytes(int n_queries, int64_t n_items, int k_dim, int kind);
ANNCUR_API int anncur_score_dense(const float* Q, int ldq, int n_queries, const void* packed_items, const float* e_scale,
                       int64_t n_items, int k_dim, int kind, float* out, int64_t ldo,
                       void* workspace, size_t workspace_bytes, void* stream);
/* The bounds kind F32R works with, as a dense matrix (verification / diagnostics): out[b][n] = the one-pass f16 score of
 * (query b, item n) plus (sign >= 0) or minus (sign < 0) its error bound b_n -- exactly what the MAIN (upper bounds) and
 * SAMPLE (lower bounds) launches of anncur_score_topk compare with their thresholds.  The guarantee the kind rests on is
 * lower <= fp32 score <= upper for every pair; tests/test_gpu_fused.py checks it against fp64, also on operands built to
 * make the fp16 rounding errors add up coherently.  packed_items must be of kind F32R. */
ANNCUR_API int anncur_score_bounds_dense(const float* Q, int ldq, int n_queries, const void* packed_items, const float* e_scale,
                              int64_t n_items, int k_dim, int sign, float* out, int64_t ldo,
                              void* workspace, size_t workspace_bytes, void* stream);

/* K7 on the same pipeline (replaces torch.norm((approx - A)[rows,:]), torch.norm(A[rows,:]),
 * eval/run_retrieval_eval_wrt_exact_crossenc.py:146-147): out_err2[r] = sum_j (Q[r,:].E[:,j] - A[r,j])^2,
 * out_norm2[r] = sum_j A[r,j]^2 in fp64; Q . E is never written.  Workspace: anncur_score_dense_workspace_bytes. */
ANNCUR_API int anncur_recon_error_packed(const float* Q, int ldq, int n_queries, const void* packed_items, const float* e_scale,
                              int64_t n_items, int k_dim, int kind, const float* A, int64_t lda,
                              double* out_err2, double* out_norm2, void* workspace, size_t workspace_bytes, void* stream);

/* Introspection: how many query rows of the LAST anncur_score_topk call that used `workspace` (same shape arguments)
 * were recomputed by the fallback pass (sampled threshold missed; kind F32R: a candidate list filled up or the exactness
 * certificate failed).  Results are correct either way -- this is the number that tells whether the fast path served
 * the call.  Blocks until `stream` is idle. */
ANNCUR_API int anncur_score_topk_redo_rows(const void* workspace, int n_queries, int64_t n_items, int k_dim, int k, int kind,
                                int* redo_rows_host, void* stream);

/* Host-buffer form of the same call -- what a CPU-resident caller such as the reference's eval
 * scripts (CPU torch tensors, eval/run_retrieval_eval_wrt_exact_crossenc_w_fixed_train_test_splits.py
 * :298-301) binds: Q_host / out_*_host are HOST pointers (pinned for the copies to be asynchronous);
 * the call enqueues H2D(Q) -> fused score + top-k -> D2H(vals, idx) on `stream` and returns; the
 * results are valid after the stream is synchronised.  The workspace (device memory) also holds
 * the staging copies. */
ANNCUR_API size_t anncur_search_host_workspace_bytes(int n_queries, int64_t n_items, int k_dim, int k, int kind);
ANNCUR_API int anncur_search_host(const float* Q_host, int ldq, int n_queries, const void* packed_items,
                       const float* e_scale, int64_t n_items, int k_dim, int kind, int k,
                       int64_t idx_offset, float* out_vals_host, int64_t* out_idx_host,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Same contract on plain fp32 E (k_dim x N, lde) with FFMA arithmetic: scores are materialised
 * in row blocks inside the workspace and reduced by the row top-k below.  Any k <= ANNCUR_MAX_K. */
ANNCUR_API size_t anncur_score_topk_f32_workspace_bytes(int n_queries, int64_t n_items, int k_dim, int k);
ANNCUR_API int anncur_score_topk_f32(const float* Q, int ldq, int n_queries, const float* E, int64_t lde,
                          int64_t n_items, int k_dim, int k, int64_t idx_offset,
                          float* out_vals, int64_t* out_idx,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- row top-k over dense scores -------------------------------------------------------------
 * torch.topk(row, k) for every row of S (n_rows x n_cols fp32, lds): the exact-score top-k of
 * the eval loops (eval/run_retrieval_eval_wrt_exact_crossenc.py:103,
 * ..._w_fixed_train_test_splits.py:86). */
ANNCUR_API int anncur_topk_rows_f32(const float* S, int64_t lds, int n_rows, int64_t n_cols, int k,
                         int64_t idx_offset, float* out_vals, int64_t* out_idx, void* stream);

/* ---- K9: merge of per-shard candidate lists ---------------------------------------------------
 * cand_vals/cand_idx: n_rows x n_cand (row-major; e.g. the all-gathered [P x k] lists laid out
 * per row).  Entries with idx < 0 are padding.  Output: best k per row. */
ANNCUR_API int anncur_merge_topk(const float* cand_vals, const int64_t* cand_idx, int n_rows, int n_cand,
                      int k, float* out_vals, int64_t* out_idx, void* stream);

/* Key form of the same merge, the exchange format of the item-sharded search: a candidate is one
 * 64-bit key = (order-preserving image of the fp32 score) << 32 | ~(uint32 global item index), so one unsigned compare
 * is the library's total order and a rank ships 8 bytes per candidate.  anncur_topk_to_keys converts a (vals, idx)
 * top-k result (global indices < 2^32; idx < 0 = padding -> key 0); anncur_merge_topk_keys takes the all-gathered
 * buffer AS IS -- keys[shard][row][k_in] -- and writes the best k_out per row (k_out <= 1024). */
ANNCUR_API int anncur_topk_to_keys(const float* vals, const int64_t* idx, int n_rows, int k, uint64_t* keys, void* stream);
ANNCUR_API size_t anncur_merge_topk_keys_workspace_bytes(int n_rows);
ANNCUR_API int anncur_merge_topk_keys(const uint64_t* keys, int n_shards, int n_rows, int k_in, int k_out,
                           float* out_vals, int64_t* out_idx, void* workspace, size_t workspace_bytes, void* stream);

/* ---- candidate exchange of the item-sharded search over NVLink peer memory (north_star: "candidates are
 * merged after ... over NVLink"; no counterpart in the reference, which is single-process) -------------------------------
 * Query rows are owned in contiguous blocks (owner o merges rows [o*B/P, (o+1)*B/P)).  anncur_peer_scatter_keys turns this
 * rank's local top-k (values + global indices) into 64-bit keys and stores every row straight into its OWNER's receive
 * buffer through the peer mapping, then publishes an epoch flag in every owner; anncur_peer_merge_owned waits for the P
 * flags of `epoch` and writes the best k_out of the P lists of each owned row.  One "channel" = one buffer of
 * anncur_peer_channel_bytes per rank, allocated with anncur_peer_alloc (plain cudaMalloc -- the one place the library
 * allocates: CUDA IPC can only export such memory), exported with anncur_peer_export (64-byte handle, exchanged by the
 * host, e.g. torch.distributed.all_gather_object) and mapped with anncur_peer_open.  peer_bases[p] = base of rank p's
 * buffer as mapped in THIS process (own rank: the local pointer).  Calls on one channel must be stream-ordered on every
 * rank, epochs 1, 2, 3, ...; all ranks pass the same n_rows / k / rows_cap / k_cap. */
ANNCUR_API size_t anncur_peer_channel_bytes(int world, int rows_cap, int k_cap);
ANNCUR_API int anncur_peer_alloc(size_t bytes, void** base_out);
ANNCUR_API int anncur_peer_free(void* base);
ANNCUR_API int anncur_peer_export(const void* base, void* handle64_out);
ANNCUR_API int anncur_peer_open(const void* handle64, void** mapped_out);
ANNCUR_API int anncur_peer_close(void* mapped);
ANNCUR_API int anncur_peer_scatter_keys(const float* vals, const int64_t* idx, int n_rows, int k, int rank, int world,
                             int rows_cap, int k_cap, uint32_t epoch, void* const* peer_bases, void* stream);
ANNCUR_API int anncur_peer_merge_owned(void* local_base, int rank, int world, int rows_owned, int rows_cap, int k_cap,
                            int k_out, uint32_t epoch, float* out_vals, int64_t* out_idx, void* workspace,
                            size_t workspace_bytes, void* stream);
/* Rank-budgeted form: when the senders ship only their best k_cap < k_out candidates per row (a row's global top-k takes
 * ~k/P items from each of P shards, so a sender need not re-score k), anncur_peer_merge_owned also runs the exactness
 * certificate of every merged row -- exact iff no full sender list was consumed entirely -- and counts the rows that fail
 * (cumulative per channel).  The caller reads the count (anncur_peer_cert_failures; reset != 0 clears it) and recomputes
 * with k_cap = k_out when it is not 0: results are never silently short, whatever the placement of the items. */
ANNCUR_API int anncur_peer_cert_failures(void* local_base, int world, int rows_cap, int k_cap, int reset,
                              unsigned* count_host, void* stream);
/* 0 = every wait so far was served; 1 + s = the wait for sender s timed out (~10 s) and the merged rows are invalid. */
ANNCUR_API int anncur_peer_error(void* local_base, int world, int rows_cap, int k_cap, int* err_host, void* stream);

/* ---- K5+K6: rerank the retrieved items by exact score, then overlap with the exact top-k -----
 * Replaces the per-query loop body at eval/run_retrieval_eval_wrt_exact_crossenc.py:108-113 /
 * ..._w_fixed_train_test_splits.py:91-96 (temp = -1e14; temp[idx] = exact[idx]; temp.topk) and
 * _compute_overlap_helper (eval/eval_utils.py:139-150).
 *   exact      n_rows x n_cols fp32 (lds)            exact CE scores
 *   retr_idx   n_rows x k_retr int64                 approx top-k_retr (item indices)
 *   exact_idx  n_rows x k_max  int64                 exact top-k_max (best first)
 *   k_list     n_k ints (host), each <= k_max <= k_retr
 *   out_rr_idx/out_rr_vals  n_rows x k_max           reranked lists
 *   out_common n_rows x n_k int32                    |exact[:k] & reranked[:k]| per row and k   */
ANNCUR_API int anncur_rerank_overlap(const float* exact, int64_t lds, int n_rows, int64_t n_cols,
                          const int64_t* retr_idx, int k_retr, const int64_t* exact_idx, int k_max,
                          const int* k_list_host, int n_k, int64_t* out_rr_idx, float* out_rr_vals,
                          int32_t* out_common, void* stream);

/* K6 on its own: out_common[row] = |set(a_idx[row, :]) & set(b_idx[row, :])| for two n_rows x k index lists
 * (_compute_overlap_helper, eval/eval_utils.py:139-150).  k <= 4096. */
ANNCUR_API int anncur_overlap_counts(const int64_t* a_idx, const int64_t* b_idx, int n_rows, int k,
                          int32_t* out_common, void* stream);

/* ---- K7: reconstruction error without materialising the approximation -------------------------
 * Replaces torch.norm((approx - A)[rows,:]) and torch.norm(A[rows,:])
 * (eval/run_retrieval_eval_wrt_exact_crossenc.py:146-147): per row r of Q (n x k_dim) and
 * A (n x N): out_err2[r] = sum_j (Q[r,:].E[:,j] - A[r,j])^2, out_norm2[r] = sum_j A[r,j]^2, fp64.
 * The caller sums rows over its subsets (anchor / non-anchor / all) and takes square roots. */
ANNCUR_API int anncur_recon_error_f32(const float* Q, int ldq, const float* E, int64_t lde, const float* A,
                           int64_t lda, int n_rows, int64_t n_items, int k_dim,
                           double* out_err2, double* out_norm2, void* stream);

/* ---- K8: adaptive multi-round ANNCUR (not in the reference) --------------
 * One round for a batch of queries that all hold m anchors:
 *   M_b = R_anc[:, anchors[b,:]] (k_q x m);  e_b = c_b . pinv(M_b);  s_b = e_b . R_anc with the
 *   anchors masked;  next[b,:] = top-n_next(s_b).
 * solved through the regularisation-free normal equations in fp64 (Cholesky with diagonal
 * pivot guard; rank-deficient systems fall back to the min-norm rule of pinv by dropping
 * pivots <= rcond * max pivot).
 *   R_anc    k_q x N fp32 (ldr)        anchors  B x m int64        c  B x m fp32 (exact scores)
 *   next_idx B x n_next int64          next_val B x n_next fp32    (approx scores of the picks) */
ANNCUR_API size_t anncur_adaptive_round_workspace_bytes(int n_queries, int k_q, int m, int64_t n_items, int n_next);
ANNCUR_API int anncur_adaptive_round(const float* R_anc, int64_t ldr, int k_q, int64_t n_items,
                          const int64_t* anchors, const float* c, int n_queries, int m,
                          double rcond, int n_next, int64_t* next_idx, float* next_val,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Split form of the same round ("optional mask / excluded-index list per row"): anncur_adaptive_solve leaves
 * e_b = c_b . pinv(R_anc[:, I_b]) (n_queries x k_q fp32); the caller re-scores with the fused tensor-core kernel on the
 * PACKED R_anc -- anncur_score_topk(e, packed R_anc, k = n_next + m): no B x N block is ever formed, and the index can be
 * item-sharded with one exchange per round -- and anncur_filter_excluded drops the row's anchors from the candidate list
 * (a row's m anchors can displace at most m of its best n_next + m candidates), keeping the first n_out in order, padded
 * with (-FLT_MAX, -1).  Rt_cached (optional): R_anc^T as n_items x k_q fp32 (anncur_transpose_f32), which does not change
 * between rounds; NULL = rebuilt in the workspace every call. */
ANNCUR_API size_t anncur_adaptive_solve_workspace_bytes(int n_queries, int k_q, int m, int64_t n_items);
ANNCUR_API int anncur_adaptive_solve(const float* R_anc, int64_t ldr, int k_q, int64_t n_items, const float* Rt_cached,
                          const int64_t* anchors, const float* c, int n_queries, int m, double rcond, float* e_out,
                          void* workspace, size_t workspace_bytes, void* stream);
ANNCUR_API int anncur_transpose_f32(const float* in, int64_t ld_in, int rows, int64_t cols, float* out, void* stream);
ANNCUR_API int anncur_filter_excluded(const float* cand_vals, const int64_t* cand_idx, int n_rows, int k_in,
                           const int64_t* excluded, int m, int n_out, float* out_vals, int64_t* out_idx, void* stream);

/* Incremental form of K8: the per-query factorisation is CARRIED across rounds (I_{t+1} = I_t U new_t) and the first
 * m_shared anchors are shared by all queries, so a round costs the Gram rows and the Cholesky rows of its n_new NEW
 * anchors only (45 MFLOP per query over BASELINE configs[2]'s rounds instead of 241; csrc/adaptive_inc.cu).
 *   Rt       R_anc^T, n_items x k_q fp32 (anncur_transpose_f32)
 *   shared   blob of anncur_adaptive_shared_bytes: built ONCE per (R_anc, shared anchors) by anncur_adaptive_prepare
 *            (Cholesky factor of the shared Gram matrix, its inverse, and W_1^T = (L_1^-1 M_1^T R_anc)^T, n_items x
 *            m_shared fp64, from which every later round gathers its shared columns)
 *   state    blob of anncur_adaptive_state_bytes: per-query factor rows, block inverses, z = L^-1 c, anchor lists
 *   anncur_adaptive_begin   round 1: e_b = c_b . pinv(R_anc[:, shared])            (c: n_queries x m_shared)
 *   anncur_adaptive_extend  every later round: n_new more anchors per query (m_cur = anchors held before the call, =
 *                           m_shared + r n_new), their exact scores c_new, e_b for the grown set (n_queries x k_q fp32)
 * m_shared and n_new <= ANNCUR_ADAPTIVE_MAX_BLOCK, m_max = m_shared + (rounds) n_new <= k_q.  Same pivot rule as
 * anncur_adaptive_solve (pivots <= rcond^2 * largest Gram diagonal are dropped: that anchor's coordinate is 0).  An anchor
 * index outside [0, n_items) -- the -1 padding of a short candidate list -- counts as a zero column (dropped the same way). */
#define ANNCUR_ADAPTIVE_MAX_BLOCK 128
ANNCUR_API size_t anncur_adaptive_shared_bytes(int k_q, int64_t n_items, int m_shared);
ANNCUR_API size_t anncur_adaptive_prepare_workspace_bytes(int k_q, int64_t n_items, int m_shared);
ANNCUR_API int anncur_adaptive_prepare(const float* Rt, int k_q, int64_t n_items, const int64_t* shared_anchors, int m_shared,
                            double rcond, void* shared, size_t shared_bytes, void* workspace, size_t workspace_bytes,
                            void* stream);
ANNCUR_API size_t anncur_adaptive_state_bytes(int n_queries, int k_q, int m_shared, int n_new, int m_max);
ANNCUR_API int anncur_adaptive_begin(const float* Rt, int k_q, int64_t n_items, const void* shared, int m_shared, const float* c,
                          int n_queries, int n_new, int m_max, float* e_out, void* state, size_t state_bytes, void* stream);
ANNCUR_API int anncur_adaptive_extend(const float* Rt, int k_q, int64_t n_items, const void* shared, int m_shared,
                           const int64_t* new_anchors, const float* c_new, int n_queries, int n_new, int m_max, int m_cur,
                           double rcond, float* e_out, void* state, size_t state_bytes, void* stream);

/* ---- per-kernel device timing (bench.py's roofline) ---------------------------------------------
 * While enabled, every anncur_score_topk call records a CUDA-event pair around its fused tcgen05
 * kernel on the call's stream.  anncur_profile_read waits for the recorded events (it is the one
 * call that blocks), returns the summed kernel milliseconds and the number of launches, and clears
 * the record. */
ANNCUR_API int anncur_profile_enable(int on);
ANNCUR_API int anncur_profile_read(double* fused_ms_sum, int* fused_launches);

/* ---- introspection used by bench.py ("gpu_launches") ------------------------------------------
 * Number of kernels this library has launched on the calling host thread since the last reset. */
ANNCUR_API uint64_t anncur_kernel_launch_count(void);
ANNCUR_API void     anncur_reset_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* ANNCUR_B200_H_ */
