// extern "C" surface of libanncur_b200.so (include/anncur_b200.h): argument checking, workspace
// carving and dispatch into the kernel translation units.  No torch types, no allocation.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace anncur {

static thread_local char g_err[512] = "";
static thread_local uint64_t g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches += uint64_t(n); }

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// rows of the score matrix materialised at once by the FFMA path (scratch = rows x N fp32 <= ~1 GiB)
static int f32_row_block(int n_queries, int64_t n_items) {
    int64_t rows = (int64_t(1) << 28) / (n_items > 0 ? n_items : 1);
    rows = rows / 128 * 128;
    if (rows < 128) rows = 128;
    if (rows > n_queries) rows = n_queries;
    return int(rows);
}

}  // namespace anncur

using namespace anncur;

extern "C" {

int anncur_abi_version(void) { return ANNCUR_ABI_VERSION; }
const char* anncur_last_error(void) { return g_err; }
uint64_t anncur_kernel_launch_count(void) { return g_launches; }
void anncur_reset_kernel_launch_count(void) { g_launches = 0; }

size_t anncur_pinv_workspace_bytes(int m, int n) { return pinv_workspace_bytes(m, n); }

int anncur_pinv_f32(const float* A, int m, int n, int lda, double rcond, float* out, int ldo, double* cond_out,
                    void* workspace, size_t workspace_bytes, void* stream) {
    ANNCUR_REQUIRE(m >= 0 && n >= 0, "pinv: negative shape %d x %d", m, n);
    if (m == 0 || n == 0) return ANNCUR_OK;
    ANNCUR_REQUIRE(A && out && workspace, "pinv: null pointer");
    ANNCUR_REQUIRE(lda >= n && ldo >= m, "pinv: lda %d < n %d or ldo %d < m %d", lda, n, ldo, m);
    return pinv_f32(A, m, n, lda, rcond, out, ldo, cond_out, workspace, workspace_bytes, cudaStream_t(stream));
}

int anncur_singular_values_f32(const float* A, int m, int n, int lda, double* sigma_out, void* workspace,
                               size_t workspace_bytes, void* stream) {
    ANNCUR_REQUIRE(m >= 0 && n >= 0, "singular_values: negative shape %d x %d", m, n);
    if (m == 0 || n == 0) return ANNCUR_OK;
    ANNCUR_REQUIRE(A && sigma_out && workspace, "singular_values: null pointer");
    ANNCUR_REQUIRE(lda >= n, "singular_values: lda %d < n %d", lda, n);
    return singular_values_f32(A, m, n, lda, sigma_out, workspace, workspace_bytes, cudaStream_t(stream));
}

int anncur_orthonormalize_f32(const float* A, int m, int n, int lda, float* Q, int ldq, double* sigma_out, void* workspace,
                              size_t workspace_bytes, void* stream) {
    ANNCUR_REQUIRE(m >= 0 && n >= 0, "orthonormalize: negative shape %d x %d", m, n);
    if (m == 0 || n == 0) return ANNCUR_OK;
    ANNCUR_REQUIRE(A && Q && workspace, "orthonormalize: null pointer");
    ANNCUR_REQUIRE(lda >= n && ldq >= n, "orthonormalize: lda %d or ldq %d < n %d", lda, ldq, n);
    return orthonormalize_f32(A, m, n, lda, Q, ldq, sigma_out, workspace, workspace_bytes, cudaStream_t(stream));
}

int anncur_jacobi_status(const void* workspace, double* status4_out, void* stream) {
    ANNCUR_REQUIRE(workspace && status4_out, "jacobi_status: null pointer");
    return jacobi_status(workspace, status4_out, cudaStream_t(stream));
}

int anncur_gemm_f32(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int m, int n, int k,
                    void* stream) {
    ANNCUR_REQUIRE(m >= 0 && n >= 0 && k >= 0, "gemm: negative shape");
    if (m == 0 || n == 0) return ANNCUR_OK;
    ANNCUR_REQUIRE(C && (k == 0 || (A && B)), "gemm: null pointer");
    ANNCUR_REQUIRE(ldc >= n && (k == 0 || (lda >= k && ldb >= n)), "gemm: leading dimension too small");
    return sgemm_rowmajor(A, lda, B, ldb, C, ldc, m, n, k, cudaStream_t(stream));
}

size_t anncur_packed_items_bytes(int64_t n_items, int k_dim, int kind) { return packed_items_bytes(n_items, k_dim, kind); }

int anncur_pack_items(const float* E, int64_t lde, int64_t n_items, int k_dim, int kind, void* packed,
                      float* e_scale_out, void* stream) {
    ANNCUR_REQUIRE(n_items >= 0 && k_dim >= 0, "pack_items: negative shape");
    ANNCUR_REQUIRE(packed && e_scale_out, "pack_items: null pointer");
    ANNCUR_REQUIRE(n_items == 0 || k_dim == 0 || (E && lde >= n_items), "pack_items: bad E / lde");
    ANNCUR_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 255) == 0, "pack_items: packed must be 256-byte aligned");
    return pack_items(E, lde, n_items, k_dim, kind, packed, e_scale_out, cudaStream_t(stream));
}

size_t anncur_score_topk_workspace_bytes(int n_queries, int64_t n_items, int k_dim, int k, int kind) {
    return score_topk_workspace_bytes(n_queries, n_items, k_dim, k, kind);
}

int anncur_score_topk(const float* Q, int ldq, int n_queries, const void* packed_items, const float* e_scale,
                      int64_t n_items, int k_dim, int kind, int k, int64_t idx_offset, float* out_vals,
                      int64_t* out_idx, void* workspace, size_t workspace_bytes, void* stream) {
    ANNCUR_REQUIRE(n_queries >= 0 && n_items >= 0 && k_dim >= 0, "score_topk: negative shape");
    if (n_queries == 0) return ANNCUR_OK;
    ANNCUR_REQUIRE(out_vals && out_idx, "score_topk: null output");
    ANNCUR_REQUIRE(n_items == 0 || k_dim == 0 || (Q && packed_items && e_scale && workspace && ldq >= k_dim),
                   "score_topk: null input or ldq < k_dim");
    return score_topk_fused(Q, ldq, n_queries, packed_items, e_scale, n_items, k_dim, kind, k, idx_offset, out_vals,
                            out_idx, workspace, workspace_bytes, cudaStream_t(stream));
}

// Masked search ("optional mask / excluded-index list per row"): top-k of every row among the items NOT in the row's
// excluded list.  Composition of the two kernels above it: the fused search for k + m_excl candidates (a row's m_excl excluded
// items can displace at most m_excl of them) into the head of the workspace, then the order-keeping filter.
static size_t excl_lists_bytes(int n_queries, int k_all) {
    return align_up(sizeof(float) * size_t(n_queries) * k_all, 256) + align_up(sizeof(int64_t) * size_t(n_queries) * k_all, 256);
}
size_t anncur_score_topk_excluding_workspace_bytes(int n_queries, int64_t n_items, int k_dim, int k, int m_excl, int kind) {
    if (n_queries <= 0 || k < 1 || m_excl < 0) return 256;
    return excl_lists_bytes(n_queries, k + m_excl) + score_topk_workspace_bytes(n_queries, n_items, k_dim, k + m_excl, kind);
}
int anncur_score_topk_excluding(const float* Q, int ldq, int n_queries, const void* packed_items, const float* e_scale,
                                int64_t n_items, int k_dim, int kind, int k, const int64_t* excluded, int m_excl,
                                int64_t idx_offset, float* out_vals, int64_t* out_idx, void* workspace, size_t workspace_bytes,
                                void* stream) {
    ANNCUR_REQUIRE(n_queries >= 0 && n_items >= 0 && k_dim >= 0 && m_excl >= 0 && k >= 1, "score_topk_excluding: bad shape");
    if (n_queries == 0) return ANNCUR_OK;
    if (m_excl == 0)
        return anncur_score_topk(Q, ldq, n_queries, packed_items, e_scale, n_items, k_dim, kind, k, idx_offset, out_vals, out_idx,
                                 workspace, workspace_bytes, stream);
    const int k_all = k + m_excl;
    ANNCUR_REQUIRE(k_all <= ANNCUR_MAX_K_FUSED, "score_topk_excluding: k + m_excl = %d > %d", k_all, ANNCUR_MAX_K_FUSED);
    ANNCUR_REQUIRE(out_vals && out_idx && excluded && workspace, "score_topk_excluding: null pointer");
    const size_t head = excl_lists_bytes(n_queries, k_all);
    ANNCUR_REQUIRE(workspace_bytes >= anncur_score_topk_excluding_workspace_bytes(n_queries, n_items, k_dim, k, m_excl, kind),
                   "score_topk_excluding: workspace too small");
    char* ws = reinterpret_cast<char*>(workspace);
    float* cand_v = reinterpret_cast<float*>(ws);
    int64_t* cand_i = reinterpret_cast<int64_t*>(ws + align_up(sizeof(float) * size_t(n_queries) * k_all, 256));
    int rc = anncur_score_topk(Q, ldq, n_queries, packed_items, e_scale, n_items, k_dim, kind, k_all, idx_offset, cand_v, cand_i,
                               ws + head, workspace_bytes - head, stream);
    if (rc != ANNCUR_OK) return rc;
    return filter_excluded(cand_v, cand_i, n_queries, k_all, excluded, m_excl, k, out_vals, out_idx, cudaStream_t(stream));
}

size_t anncur_score_dense_workspace_bytes(int n_queries, int64_t n_items, int k_dim, int kind) {
    return score_dense_workspace_bytes(n_queries, n_items, k_dim, kind);
}

int anncur_score_dense(const float* Q, int ldq, int n_queries, const void* packed_items, const float* e_scale, int64_t n_items,
                       int k_dim, int kind, float* out, int64_t ldo, void* workspace, size_t workspace_bytes, void* stream) {
    ANNCUR_REQUIRE(n_queries >= 0 && n_items >= 0 && k_dim > 0, "score_dense: bad shape");
    if (n_queries == 0 || n_items == 0) return ANNCUR_OK;
    ANNCUR_REQUIRE(Q && packed_items && e_scale && out && workspace && ldq >= k_dim && ldo >= n_items, "score_dense: null pointer or short leading dimension");
    return score_dense(Q, ldq, n_queries, packed_items, e_scale, n_items, k_dim, kind, out, ldo, workspace, workspace_bytes, cudaStream_t(stream));
}

int anncur_score_bounds_dense(const float* Q, int ldq, int n_queries, const void* packed_items, const float* e_scale, int64_t n_items,
                              int k_dim, int sign, float* out, int64_t ldo, void* workspace, size_t workspace_bytes, void* stream) {
    ANNCUR_REQUIRE(n_queries >= 0 && n_items >= 0 && k_dim > 0, "score_bounds_dense: bad shape");
    if (n_queries == 0 || n_items == 0) return ANNCUR_OK;
    ANNCUR_REQUIRE(Q && packed_items && e_scale && out && workspace && ldq >= k_dim && ldo >= n_items, "score_bounds_dense: null pointer or short leading dimension");
    return score_bounds_dense(Q, ldq, n_queries, packed_items, e_scale, n_items, k_dim, sign, out, ldo, workspace, workspace_bytes, cudaStream_t(stream));
}

int anncur_recon_error_packed(const float* Q, int ldq, int n_queries, const void* packed_items, const float* e_scale,
                              int64_t n_items, int k_dim, int kind, const float* A, int64_t lda, double* out_err2,
                              double* out_norm2, void* workspace, size_t workspace_bytes, void* stream) {
    ANNCUR_REQUIRE(n_queries >= 0 && n_items >= 0 && k_dim > 0, "recon_error_packed: bad shape");
    if (n_queries == 0) return ANNCUR_OK;
    ANNCUR_REQUIRE(out_err2 && out_norm2, "recon_error_packed: null output");
    if (n_items == 0) {
        ANNCUR_CUDA_OK(cudaMemsetAsync(out_err2, 0, sizeof(double) * size_t(n_queries), cudaStream_t(stream)));
        ANNCUR_CUDA_OK(cudaMemsetAsync(out_norm2, 0, sizeof(double) * size_t(n_queries), cudaStream_t(stream)));
        return ANNCUR_OK;
    }
    ANNCUR_REQUIRE(Q && packed_items && e_scale && A && workspace && ldq >= k_dim && lda >= n_items, "recon_error_packed: null pointer or short leading dimension");
    return recon_error_packed(Q, ldq, n_queries, packed_items, e_scale, n_items, k_dim, kind, A, lda, out_err2, out_norm2, workspace,
                              workspace_bytes, cudaStream_t(stream));
}

int anncur_score_topk_redo_rows(const void* workspace, int n_queries, int64_t n_items, int k_dim, int k, int kind,
                                int* redo_rows_host, void* stream) {
    ANNCUR_REQUIRE(workspace && redo_rows_host, "score_topk_redo_rows: null pointer");
    return score_topk_redo_rows(workspace, n_queries, n_items, k_dim, k, kind, redo_rows_host, cudaStream_t(stream));
}

// ---- host-buffer search: H2D of the query batch, fused score + top-k, D2H of the result ----------
static size_t host_stage_q_bytes(int n_queries, int k_dim) { return align_up(sizeof(float) * size_t(n_queries) * size_t(k_dim > 0 ? k_dim : 1), 256); }
static size_t host_stage_v_bytes(int n_queries, int k) { return align_up(sizeof(float) * size_t(n_queries) * size_t(k), 256); }
static size_t host_stage_i_bytes(int n_queries, int k) { return align_up(sizeof(int64_t) * size_t(n_queries) * size_t(k), 256); }

size_t anncur_search_host_workspace_bytes(int n_queries, int64_t n_items, int k_dim, int k, int kind) {
    if (n_queries <= 0 || k <= 0) return 256;
    return host_stage_q_bytes(n_queries, k_dim) + host_stage_v_bytes(n_queries, k) + host_stage_i_bytes(n_queries, k) +
           score_topk_workspace_bytes(n_queries, n_items, k_dim, k, kind);
}

int anncur_search_host(const float* Q_host, int ldq, int n_queries, const void* packed_items, const float* e_scale,
                       int64_t n_items, int k_dim, int kind, int k, int64_t idx_offset, float* out_vals_host,
                       int64_t* out_idx_host, void* workspace, size_t workspace_bytes, void* stream) {
    ANNCUR_REQUIRE(n_queries >= 0 && n_items >= 0 && k_dim >= 0 && k >= 1, "search_host: bad shape");
    if (n_queries == 0) return ANNCUR_OK;
    ANNCUR_REQUIRE(out_vals_host && out_idx_host && workspace, "search_host: null pointer");
    ANNCUR_REQUIRE(k_dim == 0 || (Q_host && ldq >= k_dim), "search_host: null Q or ldq < k_dim");
    ANNCUR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "search_host: workspace must be 256-byte aligned");
    const size_t need = anncur_search_host_workspace_bytes(n_queries, n_items, k_dim, k, kind);
    if (workspace_bytes < need) { set_error("search_host workspace too small: %zu < %zu", workspace_bytes, need); return ANNCUR_E_WORKSPACE; }
    cudaStream_t s = cudaStream_t(stream);
    char* ws = reinterpret_cast<char*>(workspace);
    float* q_dev = reinterpret_cast<float*>(ws);                      ws += host_stage_q_bytes(n_queries, k_dim);
    float* v_dev = reinterpret_cast<float*>(ws);                      ws += host_stage_v_bytes(n_queries, k);
    int64_t* i_dev = reinterpret_cast<int64_t*>(ws);                  ws += host_stage_i_bytes(n_queries, k);
    const size_t inner = workspace_bytes - size_t(ws - reinterpret_cast<char*>(workspace));
    if (k_dim > 0 && ldq == k_dim)
        ANNCUR_CUDA_OK(cudaMemcpyAsync(q_dev, Q_host, sizeof(float) * size_t(k_dim) * size_t(n_queries), cudaMemcpyHostToDevice, s));
    else if (k_dim > 0)
        ANNCUR_CUDA_OK(cudaMemcpy2DAsync(q_dev, sizeof(float) * size_t(k_dim), Q_host, sizeof(float) * size_t(ldq),
                                         sizeof(float) * size_t(k_dim), size_t(n_queries), cudaMemcpyHostToDevice, s));
    int rc = anncur_score_topk(q_dev, k_dim, n_queries, packed_items, e_scale, n_items, k_dim, kind, k, idx_offset, v_dev,
                               i_dev, ws, inner, stream);
    if (rc != ANNCUR_OK) return rc;
    ANNCUR_CUDA_OK(cudaMemcpyAsync(out_vals_host, v_dev, sizeof(float) * size_t(n_queries) * size_t(k), cudaMemcpyDeviceToHost, s));
    ANNCUR_CUDA_OK(cudaMemcpyAsync(out_idx_host, i_dev, sizeof(int64_t) * size_t(n_queries) * size_t(k), cudaMemcpyDeviceToHost, s));
    return ANNCUR_OK;
}

size_t anncur_score_topk_f32_workspace_bytes(int n_queries, int64_t n_items, int k_dim, int k) {
    (void)k_dim; (void)k;
    if (n_queries <= 0 || n_items <= 0) return 256;
    return align_up(sizeof(float) * size_t(f32_row_block(n_queries, n_items)) * size_t(n_items), 256);
}

int anncur_score_topk_f32(const float* Q, int ldq, int n_queries, const float* E, int64_t lde, int64_t n_items,
                          int k_dim, int k, int64_t idx_offset, float* out_vals, int64_t* out_idx, void* workspace,
                          size_t workspace_bytes, void* stream) {
    ANNCUR_REQUIRE(n_queries >= 0 && n_items >= 0 && k_dim >= 0, "score_topk_f32: negative shape");
    ANNCUR_REQUIRE(k >= 1 && k <= ANNCUR_MAX_K, "score_topk_f32: k = %d outside [1, %d]", k, ANNCUR_MAX_K);
    if (n_queries == 0) return ANNCUR_OK;
    ANNCUR_REQUIRE(out_vals && out_idx && workspace, "score_topk_f32: null pointer");
    ANNCUR_REQUIRE(k_dim == 0 || n_items == 0 || (Q && E && ldq >= k_dim && lde >= n_items), "score_topk_f32: bad input");
    if (workspace_bytes < anncur_score_topk_f32_workspace_bytes(n_queries, n_items, k_dim, k)) {
        set_error("score_topk_f32 workspace too small");
        return ANNCUR_E_WORKSPACE;
    }
    cudaStream_t s = cudaStream_t(stream);
    float* scratch = reinterpret_cast<float*>(workspace);
    const int rb = f32_row_block(n_queries, n_items > 0 ? n_items : 1);
    for (int r0 = 0; r0 < n_queries; r0 += rb) {
        const int rows = n_queries - r0 < rb ? n_queries - r0 : rb;
        int rc = sgemm_rowmajor(Q + int64_t(r0) * ldq, ldq, E, lde, scratch, n_items, rows, n_items, k_dim, s);
        if (rc != ANNCUR_OK) return rc;
        rc = select_topk_dense(scratch, n_items, rows, n_items, k, idx_offset, out_vals + int64_t(r0) * k,
                               out_idx + int64_t(r0) * k, s);
        if (rc != ANNCUR_OK) return rc;
    }
    return ANNCUR_OK;
}

int anncur_topk_rows_f32(const float* S, int64_t lds, int n_rows, int64_t n_cols, int k, int64_t idx_offset,
                         float* out_vals, int64_t* out_idx, void* stream) {
    ANNCUR_REQUIRE(n_rows >= 0 && n_cols >= 0, "topk_rows: negative shape");
    ANNCUR_REQUIRE(k >= 1 && k <= ANNCUR_MAX_K, "topk_rows: k = %d outside [1, %d]", k, ANNCUR_MAX_K);
    if (n_rows == 0) return ANNCUR_OK;
    ANNCUR_REQUIRE(out_vals && out_idx && (n_cols == 0 || (S && lds >= n_cols)), "topk_rows: bad pointers / lds");
    ANNCUR_REQUIRE(n_cols < (int64_t(1) << 32) - 1, "topk_rows: n_cols too large");
    return select_topk_dense(S, lds, n_rows, n_cols, k, idx_offset, out_vals, out_idx, cudaStream_t(stream));
}

int anncur_merge_topk(const float* cand_vals, const int64_t* cand_idx, int n_rows, int n_cand, int k, float* out_vals,
                      int64_t* out_idx, void* stream) {
    ANNCUR_REQUIRE(n_rows >= 0 && n_cand >= 0, "merge_topk: negative shape");
    ANNCUR_REQUIRE(k >= 1 && k <= ANNCUR_MAX_K, "merge_topk: k = %d outside [1, %d]", k, ANNCUR_MAX_K);
    if (n_rows == 0) return ANNCUR_OK;
    ANNCUR_REQUIRE(out_vals && out_idx && (n_cand == 0 || (cand_vals && cand_idx)), "merge_topk: null pointer");
    return select_topk_pairs(cand_vals, cand_idx, n_rows, n_cand, k, out_vals, out_idx, cudaStream_t(stream));
}

int anncur_topk_to_keys(const float* vals, const int64_t* idx, int n_rows, int k, uint64_t* keys, void* stream) {
    ANNCUR_REQUIRE(n_rows >= 0 && k >= 1, "topk_to_keys: bad shape");
    if (n_rows == 0) return ANNCUR_OK;
    ANNCUR_REQUIRE(vals && idx && keys, "topk_to_keys: null pointer");
    return topk_to_keys(vals, idx, n_rows, k, keys, cudaStream_t(stream));
}

size_t anncur_merge_topk_keys_workspace_bytes(int n_rows) { return align_up(sizeof(uint32_t) * (size_t(n_rows > 0 ? n_rows : 0) + 1), 256); }

int anncur_merge_topk_keys(const uint64_t* keys, int n_shards, int n_rows, int k_in, int k_out, float* out_vals,
                           int64_t* out_idx, void* workspace, size_t workspace_bytes, void* stream) {
    ANNCUR_REQUIRE(n_shards >= 1 && n_rows >= 0 && k_in >= 1, "merge_topk_keys: bad shape");
    ANNCUR_REQUIRE(k_out >= 1 && k_out <= ANNCUR_MAX_K, "merge_topk_keys: k = %d outside [1, %d]", k_out, ANNCUR_MAX_K);
    if (n_rows == 0) return ANNCUR_OK;
    ANNCUR_REQUIRE(keys && out_vals && out_idx && workspace, "merge_topk_keys: null pointer");
    ANNCUR_REQUIRE(k_out <= 1024, "merge_topk_keys: k_out = %d > 1024", k_out);
    if (workspace_bytes < anncur_merge_topk_keys_workspace_bytes(n_rows)) { set_error("merge_topk_keys workspace too small"); return ANNCUR_E_WORKSPACE; }
    return merge_topk_keys(keys, n_shards, n_rows, k_in, k_out, out_vals, out_idx, reinterpret_cast<uint32_t*>(workspace),
                           cudaStream_t(stream));
}

size_t anncur_peer_channel_bytes(int world, int rows_cap, int k_cap) { return peer_channel_bytes(world, rows_cap, k_cap); }
int anncur_peer_alloc(size_t bytes, void** base_out) {
    ANNCUR_REQUIRE(base_out && bytes > 0, "peer_alloc: null pointer or zero size");
    return peer_alloc(bytes, base_out);
}
int anncur_peer_free(void* base) { return peer_free(base); }
int anncur_peer_export(const void* base, void* handle64_out) {
    ANNCUR_REQUIRE(base && handle64_out, "peer_export: null pointer");
    return peer_export(base, handle64_out);
}
int anncur_peer_open(const void* handle64, void** mapped_out) {
    ANNCUR_REQUIRE(handle64 && mapped_out, "peer_open: null pointer");
    return peer_open(handle64, mapped_out);
}
int anncur_peer_close(void* mapped) { return peer_close(mapped); }
int anncur_peer_scatter_keys(const float* vals, const int64_t* idx, int n_rows, int k, int rank, int world, int rows_cap,
                             int k_cap, uint32_t epoch, void* const* peer_bases, void* stream) {
    ANNCUR_REQUIRE(peer_bases && (n_rows == 0 || (vals && idx)), "peer_scatter_keys: null pointer");
    ANNCUR_REQUIRE(rows_cap >= 1 && k_cap >= 1, "peer_scatter_keys: bad capacities");
    return peer_scatter_keys(vals, idx, n_rows, k, rank, world, rows_cap, k_cap, epoch, peer_bases, cudaStream_t(stream));
}
int anncur_peer_merge_owned(void* local_base, int rank, int world, int rows_owned, int rows_cap, int k_cap, int k_out,
                            uint32_t epoch, float* out_vals, int64_t* out_idx, void* workspace, size_t workspace_bytes,
                            void* stream) {
    ANNCUR_REQUIRE(local_base && workspace && (rows_owned == 0 || (out_vals && out_idx)), "peer_merge_owned: null pointer");
    ANNCUR_REQUIRE(k_out >= 1 && k_out <= 1024, "peer_merge_owned: k_out = %d outside [1, 1024]", k_out);
    ANNCUR_REQUIRE(k_out <= world * k_cap, "peer_merge_owned: k_out = %d > world * k_cap = %d", k_out, world * k_cap);
    if (workspace_bytes < anncur_merge_topk_keys_workspace_bytes(rows_owned)) { set_error("peer_merge_owned workspace too small"); return ANNCUR_E_WORKSPACE; }
    return peer_merge_owned(local_base, rank, world, rows_owned, rows_cap, k_cap, k_out, epoch, out_vals, out_idx,
                            reinterpret_cast<uint32_t*>(workspace), cudaStream_t(stream));
}
int anncur_peer_cert_failures(void* local_base, int world, int rows_cap, int k_cap, int reset, unsigned* count_host, void* stream) {
    ANNCUR_REQUIRE(local_base && count_host, "peer_cert_failures: null pointer");
    return peer_cert_failures(local_base, world, rows_cap, k_cap, reset, count_host, cudaStream_t(stream));
}
int anncur_peer_error(void* local_base, int world, int rows_cap, int k_cap, int* err_host, void* stream) {
    ANNCUR_REQUIRE(local_base && err_host, "peer_error: null pointer");
    return peer_error(local_base, world, rows_cap, k_cap, err_host, cudaStream_t(stream));
}

int anncur_rerank_overlap(const float* exact, int64_t lds, int n_rows, int64_t n_cols, const int64_t* retr_idx,
                          int k_retr, const int64_t* exact_idx, int k_max, const int* k_list_host, int n_k,
                          int64_t* out_rr_idx, float* out_rr_vals, int32_t* out_common, void* stream) {
    ANNCUR_REQUIRE(n_rows >= 0, "rerank_overlap: negative n_rows");
    if (n_rows == 0) return ANNCUR_OK;
    ANNCUR_REQUIRE(exact && retr_idx && exact_idx && k_list_host && out_rr_idx && out_rr_vals && out_common,
                   "rerank_overlap: null pointer");
    ANNCUR_REQUIRE(lds >= n_cols, "rerank_overlap: lds < n_cols");
    return rerank_overlap(exact, lds, n_rows, n_cols, retr_idx, k_retr, exact_idx, k_max, k_list_host, n_k,
                          out_rr_idx, out_rr_vals, out_common, cudaStream_t(stream));
}

int anncur_overlap_counts(const int64_t* a_idx, const int64_t* b_idx, int n_rows, int k, int32_t* out_common, void* stream) {
    ANNCUR_REQUIRE(n_rows >= 0, "overlap_counts: negative n_rows");
    if (n_rows == 0) return ANNCUR_OK;
    ANNCUR_REQUIRE(a_idx && b_idx && out_common, "overlap_counts: null pointer");
    return overlap_counts(a_idx, b_idx, n_rows, k, out_common, cudaStream_t(stream));
}

int anncur_recon_error_f32(const float* Q, int ldq, const float* E, int64_t lde, const float* A, int64_t lda,
                           int n_rows, int64_t n_items, int k_dim, double* out_err2, double* out_norm2,
                           void* stream) {
    ANNCUR_REQUIRE(n_rows >= 0 && n_items >= 0 && k_dim >= 0, "recon_error: negative shape");
    if (n_rows == 0) return ANNCUR_OK;
    ANNCUR_REQUIRE(out_err2 && out_norm2, "recon_error: null output");
    ANNCUR_REQUIRE(n_items == 0 || (A && lda >= n_items && (k_dim == 0 || (Q && E && ldq >= k_dim && lde >= n_items))),
                   "recon_error: bad input");
    return recon_error(Q, ldq, E, lde, A, lda, n_rows, n_items, k_dim, out_err2, out_norm2, cudaStream_t(stream));
}

int anncur_profile_enable(int on) { return profile_enable(on); }
int anncur_profile_read(double* fused_ms_sum, int* fused_launches) { return profile_read(fused_ms_sum, fused_launches); }

size_t anncur_adaptive_round_workspace_bytes(int n_queries, int k_q, int m, int64_t n_items, int n_next) {
    return adaptive_round_workspace_bytes(n_queries, k_q, m, n_items, n_next);
}

int anncur_adaptive_round(const float* R_anc, int64_t ldr, int k_q, int64_t n_items, const int64_t* anchors,
                          const float* c, int n_queries, int m, double rcond, int n_next, int64_t* next_idx,
                          float* next_val, void* workspace, size_t workspace_bytes, void* stream) {
    ANNCUR_REQUIRE(n_queries >= 0 && k_q > 0 && m > 0 && n_items > 0, "adaptive_round: bad shape");
    if (n_queries == 0) return ANNCUR_OK;
    ANNCUR_REQUIRE(R_anc && anchors && c && next_idx && next_val && workspace, "adaptive_round: null pointer");
    ANNCUR_REQUIRE(ldr >= n_items, "adaptive_round: ldr < n_items");
    ANNCUR_REQUIRE(n_next >= 1 && n_next <= ANNCUR_MAX_K, "adaptive_round: n_next = %d outside [1, %d]", n_next, ANNCUR_MAX_K);
    return adaptive_round(R_anc, ldr, k_q, n_items, anchors, c, n_queries, m, rcond, n_next, next_idx, next_val,
                          workspace, workspace_bytes, cudaStream_t(stream));
}

size_t anncur_adaptive_solve_workspace_bytes(int n_queries, int k_q, int m, int64_t n_items) {
    return adaptive_solve_workspace_bytes(n_queries, k_q, m, n_items);
}

int anncur_adaptive_solve(const float* R_anc, int64_t ldr, int k_q, int64_t n_items, const float* Rt_cached,
                          const int64_t* anchors, const float* c, int n_queries, int m, double rcond, float* e_out,
                          void* workspace, size_t workspace_bytes, void* stream) {
    ANNCUR_REQUIRE(n_queries >= 0 && k_q > 0 && m > 0 && n_items > 0, "adaptive_solve: bad shape");
    if (n_queries == 0) return ANNCUR_OK;
    ANNCUR_REQUIRE((R_anc || Rt_cached) && anchors && c && e_out && workspace, "adaptive_solve: null pointer");
    ANNCUR_REQUIRE(Rt_cached || ldr >= n_items, "adaptive_solve: ldr < n_items");
    return adaptive_solve(R_anc, ldr, k_q, n_items, Rt_cached, anchors, c, n_queries, m, rcond, e_out, workspace, workspace_bytes,
                          cudaStream_t(stream));
}

int anncur_transpose_f32(const float* in, int64_t ld_in, int rows, int64_t cols, float* out, void* stream) {
    ANNCUR_REQUIRE(rows >= 0 && cols >= 0, "transpose: negative shape");
    if (rows == 0 || cols == 0) return ANNCUR_OK;
    ANNCUR_REQUIRE(in && out && ld_in >= cols, "transpose: null pointer or ld_in < cols");
    return transpose_rows(in, ld_in, rows, cols, out, cudaStream_t(stream));
}

int anncur_filter_excluded(const float* cand_vals, const int64_t* cand_idx, int n_rows, int k_in, const int64_t* excluded,
                           int m, int n_out, float* out_vals, int64_t* out_idx, void* stream) {
    ANNCUR_REQUIRE(n_rows >= 0 && k_in >= 1 && m >= 0 && n_out >= 1, "filter_excluded: bad shape");
    if (n_rows == 0) return ANNCUR_OK;
    ANNCUR_REQUIRE(cand_vals && cand_idx && out_vals && out_idx && (m == 0 || excluded), "filter_excluded: null pointer");
    return filter_excluded(cand_vals, cand_idx, n_rows, k_in, excluded, m, n_out, out_vals, out_idx, cudaStream_t(stream));
}

size_t anncur_adaptive_shared_bytes(int k_q, int64_t n_items, int m_shared) { return adaptive_shared_bytes(k_q, n_items, m_shared); }
size_t anncur_adaptive_prepare_workspace_bytes(int k_q, int64_t n_items, int m_shared) {
    return adaptive_prepare_workspace_bytes(k_q, n_items, m_shared);
}
int anncur_adaptive_prepare(const float* Rt, int k_q, int64_t n_items, const int64_t* shared_anchors, int m_shared, double rcond,
                            void* shared, size_t shared_bytes, void* workspace, size_t workspace_bytes, void* stream) {
    ANNCUR_REQUIRE(k_q > 0 && n_items > 0 && m_shared >= 0, "adaptive_prepare: bad shape");
    ANNCUR_REQUIRE(Rt && shared && (m_shared == 0 || (shared_anchors && workspace)), "adaptive_prepare: null pointer");
    return adaptive_prepare(Rt, k_q, n_items, shared_anchors, m_shared, rcond, shared, shared_bytes, workspace, workspace_bytes,
                            cudaStream_t(stream));
}
size_t anncur_adaptive_state_bytes(int n_queries, int k_q, int m_shared, int n_new, int m_max) {
    return adaptive_state_bytes(n_queries, k_q, m_shared, n_new, m_max);
}
int anncur_adaptive_begin(const float* Rt, int k_q, int64_t n_items, const void* shared, int m_shared, const float* c, int n_queries,
                          int n_new, int m_max, float* e_out, void* state, size_t state_bytes, void* stream) {
    ANNCUR_REQUIRE(n_queries >= 0 && k_q > 0 && n_items > 0 && m_shared >= 0, "adaptive_begin: bad shape");
    ANNCUR_REQUIRE(Rt && shared && state && (n_queries == 0 || (e_out && (m_shared == 0 || c))), "adaptive_begin: null pointer");
    return adaptive_begin(Rt, k_q, n_items, shared, m_shared, c, n_queries, n_new, m_max, e_out, state, state_bytes, cudaStream_t(stream));
}
int anncur_adaptive_extend(const float* Rt, int k_q, int64_t n_items, const void* shared, int m_shared, const int64_t* new_anchors,
                           const float* c_new, int n_queries, int n_new, int m_max, int m_cur, double rcond, float* e_out,
                           void* state, size_t state_bytes, void* stream) {
    ANNCUR_REQUIRE(n_queries >= 0 && k_q > 0 && n_items > 0 && m_shared >= 0, "adaptive_extend: bad shape");
    ANNCUR_REQUIRE(Rt && shared && state && (n_queries == 0 || (new_anchors && c_new && e_out)), "adaptive_extend: null pointer");
    return adaptive_extend(Rt, k_q, n_items, shared, m_shared, new_anchors, c_new, n_queries, n_new, m_max, m_cur, rcond, e_out,
                           state, state_bytes, cudaStream_t(stream));
}

}  // extern "C"
