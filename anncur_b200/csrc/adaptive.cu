// K8: one round of adaptive multi-round ANNCUR for a batch of queries (NOT in the reference; spec =
// SURVEY.md section 8a row A8, parity unpinned -- checked against oracle.cur_oracle.adaptive_anncur).
//
// For query b with anchor items I_b (|I_b| = m <= k_q) and exact scores c_b on them:
//     M_b = R_anc[:, I_b]  (k_q x m)        e_b = c_b . pinv(M_b)         s_b = e_b . R_anc, I_b masked
// pinv through the normal equations in fp64:  G_b = M_b^T M_b,  G_b y = c_b (Cholesky, pivots that
// fall below rcond * max pivot are dropped = that coordinate of y is 0),  e_b = (M_b y)^T.
// Queries are processed in blocks so that the scratch (gathered anchors, Gram matrices, the block of
// approximate scores) stays bounded; the re-score is the FFMA GEMM and the pick is the row top-k.
#include "common.cuh"
#include "kernels.h"

namespace anncur {

constexpr int AD_QB = 128;          // queries per block

__global__ void transpose_kernel(const float* __restrict__ in, int64_t ld_in, int rows, int64_t cols,
                                 float* __restrict__ out /* cols x rows */) {
    __shared__ float tile[32][33];
    const int64_t c0 = int64_t(blockIdx.x) * 32;
    const int r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i; int64_t c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? in[int64_t(r) * ld_in + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int64_t c = c0 + i; int r = r0 + threadIdx.x;
        if (c < cols && r < rows) out[c * rows + r] = tile[threadIdx.x][i];
    }
}

// Mt[b][j][:] = Rt[anchors[b][j]][:]   (one warp per gathered item row)
__global__ void gather_anchor_rows_kernel(const float* __restrict__ Rt, int k_q, const int64_t* __restrict__ anchors,
                                          int m, int n_q, float* __restrict__ Mt) {
    const int64_t w = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
    if (w >= int64_t(n_q) * m) return;
    const int64_t item = anchors[w];
    const float* src = Rt + item * k_q;
    float* dst = Mt + w * k_q;
    for (int t = lane_id(); t < k_q; t += 32) dst[t] = src[t];
}

// G[b] = Mt[b] . Mt[b]^T  (m x m, fp64 accumulate), 32 x 32 output tile per CTA
__global__ void __launch_bounds__(256)
gram_kernel(const float* __restrict__ Mt, int m, int k_q, double* __restrict__ G) {
    __shared__ float As[32][33], Bs[32][33];
    const int b = blockIdx.z;
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    if (j0 > i0) return;                                   // lower triangle only
    const float* M = Mt + int64_t(b) * m * k_q;
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;   // 32 x 8
    double acc[4] = {0, 0, 0, 0};
    for (int k0 = 0; k0 < k_q; k0 += 32) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            int r = ty + s * 8;
            As[r][tx] = (i0 + r < m && k0 + tx < k_q) ? M[int64_t(i0 + r) * k_q + k0 + tx] : 0.f;
            Bs[r][tx] = (j0 + r < m && k0 + tx < k_q) ? M[int64_t(j0 + r) * k_q + k0 + tx] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < 32; ++kk) {
            double bv = double(Bs[tx][kk]);
#pragma unroll
            for (int s = 0; s < 4; ++s) acc[s] = fma(double(As[ty + s * 8][kk]), bv, acc[s]);
        }
        __syncthreads();
    }
    double* Gb = G + int64_t(b) * m * m;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        int i = i0 + ty + s * 8, j = j0 + tx;
        if (i < m && j < m) { Gb[int64_t(i) * m + j] = acc[s]; Gb[int64_t(j) * m + i] = acc[s]; }
    }
}

// One CTA per query: in-place Cholesky G = L L^T (lower), forward/back substitution for y, then
// e = y^T Mt  (1 x k_q).  Dropped pivots zero the matching coordinate of y.
__global__ void __launch_bounds__(256)
cholesky_solve_kernel(double* __restrict__ G, const float* __restrict__ c, const float* __restrict__ Mt, int m, int k_q,
                      double rcond, double* __restrict__ ybuf, float* __restrict__ e_out) {
    const int b = blockIdx.x, tid = threadIdx.x;
    double* L = G + int64_t(b) * m * m;
    double* y = ybuf + int64_t(b) * m;
    __shared__ double s_piv, s_maxd;
    __shared__ double red[8];
    // max diagonal -> drop tolerance
    double md = 0.0;
    for (int i = tid; i < m; i += 256) md = fmax(md, L[int64_t(i) * m + i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) md = fmax(md, __shfl_xor_sync(0xffffffffu, md, o));
    if ((tid & 31) == 0) red[tid >> 5] = md;
    __syncthreads();
    if (tid == 0) { double v = 0; for (int w = 0; w < 8; ++w) v = fmax(v, red[w]); s_maxd = v; }
    __syncthreads();
    const double drop = fmax(rcond * rcond, 1e-28) * s_maxd;   // pivots are squared singular-value scale
    for (int j = 0; j < m; ++j) {
        if (tid == 0) {
            double d = L[int64_t(j) * m + j];
            s_piv = d > drop ? sqrt(d) : 0.0;
            L[int64_t(j) * m + j] = s_piv;
        }
        __syncthreads();
        const double piv = s_piv;
        if (piv == 0.0) {
            for (int i = j + 1 + tid; i < m; i += 256) L[int64_t(i) * m + j] = 0.0;
            __syncthreads();
            continue;
        }
        const double inv = 1.0 / piv;
        for (int i = j + 1 + tid; i < m; i += 256) L[int64_t(i) * m + j] *= inv;
        __syncthreads();
        // trailing update of the lower triangle: L[i][t] -= L[i][j] * L[t][j], j < t <= i
        const int rem = m - j - 1;
        for (int e = tid; e < rem * rem; e += 256) {
            int i = j + 1 + e / rem, t = j + 1 + e % rem;
            if (t <= i) L[int64_t(i) * m + t] -= L[int64_t(i) * m + j] * L[int64_t(t) * m + j];
        }
        __syncthreads();
    }
    // forward substitution L z = c (thread 0 pivots, all threads update), z stored in y
    for (int i = tid; i < m; i += 256) y[i] = double(c[int64_t(b) * m + i]);
    __syncthreads();
    for (int j = 0; j < m; ++j) {
        if (tid == 0) { double p = L[int64_t(j) * m + j]; y[j] = p > 0.0 ? y[j] / p : 0.0; }
        __syncthreads();
        const double yj = y[j];
        for (int i = j + 1 + tid; i < m; i += 256) y[i] -= L[int64_t(i) * m + j] * yj;
        __syncthreads();
    }
    // back substitution L^T y = z
    for (int j = m - 1; j >= 0; --j) {
        if (tid == 0) { double p = L[int64_t(j) * m + j]; y[j] = p > 0.0 ? y[j] / p : 0.0; }
        __syncthreads();
        const double yj = y[j];
        for (int i = tid; i < j; i += 256) y[i] -= L[int64_t(j) * m + i] * yj;
        __syncthreads();
    }
    // e = y^T Mt
    const float* M = Mt + int64_t(b) * m * k_q;
    for (int t = tid; t < k_q; t += 256) {
        double acc = 0.0;
        for (int i = 0; i < m; ++i) acc = fma(y[i], double(M[int64_t(i) * k_q + t]), acc);
        e_out[int64_t(b) * k_q + t] = float(acc);
    }
}

__global__ void mask_anchors_kernel(float* __restrict__ S, int64_t lds, const int64_t* __restrict__ anchors, int m, int n_q) {
    const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (t >= int64_t(n_q) * m) return;
    S[(t / m) * lds + anchors[t]] = -INFINITY;
}

struct AdaptivePlan { size_t off_rt, off_mt, off_g, off_y, off_e, off_s, total; };
static AdaptivePlan adaptive_plan(int k_q, int m, int64_t n_items) {
    AdaptivePlan p{};
    size_t off = 0;
    p.off_rt = off; off += align_up(sizeof(float) * size_t(n_items) * k_q, 256);
    p.off_mt = off; off += align_up(sizeof(float) * size_t(AD_QB) * m * k_q, 256);
    p.off_g = off; off += align_up(sizeof(double) * size_t(AD_QB) * m * m, 256);
    p.off_y = off; off += align_up(sizeof(double) * size_t(AD_QB) * m, 256);
    p.off_e = off; off += align_up(sizeof(float) * size_t(AD_QB) * k_q, 256);
    p.off_s = off; off += align_up(sizeof(float) * size_t(AD_QB) * n_items, 256);
    p.total = off;
    return p;
}

size_t adaptive_round_workspace_bytes(int n_queries, int k_q, int m, int64_t n_items, int n_next) {
    (void)n_queries; (void)n_next;
    if (k_q <= 0 || m <= 0 || n_items <= 0) return 256;
    return adaptive_plan(k_q, m, n_items).total;
}

int adaptive_round(const float* R_anc, int64_t ldr, int k_q, int64_t n_items, const int64_t* anchors, const float* c,
                   int n_queries, int m, double rcond, int n_next, int64_t* next_idx, float* next_val,
                   void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (m > k_q) { set_error("adaptive_round: m = %d anchors > k_q = %d anchor queries is not supported", m, k_q); return ANNCUR_E_UNSUPPORTED; }
    const AdaptivePlan pl = adaptive_plan(k_q, m, n_items);
    if (workspace_bytes < pl.total) { set_error("adaptive_round workspace too small: %zu < %zu", workspace_bytes, pl.total); return ANNCUR_E_WORKSPACE; }
    char* ws = reinterpret_cast<char*>(workspace);
    float* Rt = reinterpret_cast<float*>(ws + pl.off_rt);
    float* Mt = reinterpret_cast<float*>(ws + pl.off_mt);
    double* G = reinterpret_cast<double*>(ws + pl.off_g);
    double* y = reinterpret_cast<double*>(ws + pl.off_y);
    float* e = reinterpret_cast<float*>(ws + pl.off_e);
    float* S = reinterpret_cast<float*>(ws + pl.off_s);

    dim3 tgrid(unsigned((n_items + 31) / 32), unsigned((k_q + 31) / 32));
    transpose_kernel<<<tgrid, dim3(32, 8), 0, stream>>>(R_anc, ldr, k_q, n_items, Rt);
    ANNCUR_LAUNCH_OK("transpose_kernel");
    for (int q0 = 0; q0 < n_queries; q0 += AD_QB) {
        const int nq = n_queries - q0 < AD_QB ? n_queries - q0 : AD_QB;
        const int64_t* anc = anchors + int64_t(q0) * m;
        const int64_t warps = int64_t(nq) * m;
        gather_anchor_rows_kernel<<<unsigned((warps * 32 + 255) / 256), 256, 0, stream>>>(Rt, k_q, anc, m, nq, Mt);
        ANNCUR_LAUNCH_OK("gather_anchor_rows_kernel");
        dim3 ggrid(unsigned((m + 31) / 32), unsigned((m + 31) / 32), unsigned(nq));
        gram_kernel<<<ggrid, 256, 0, stream>>>(Mt, m, k_q, G);
        ANNCUR_LAUNCH_OK("gram_kernel");
        cholesky_solve_kernel<<<nq, 256, 0, stream>>>(G, c + int64_t(q0) * m, Mt, m, k_q, rcond, y, e);
        ANNCUR_LAUNCH_OK("cholesky_solve_kernel");
        int rc = sgemm_rowmajor(e, k_q, R_anc, ldr, S, n_items, nq, n_items, k_q, stream);
        if (rc != ANNCUR_OK) return rc;
        mask_anchors_kernel<<<unsigned((warps + 255) / 256), 256, 0, stream>>>(S, n_items, anc, m, nq);
        ANNCUR_LAUNCH_OK("mask_anchors_kernel");
        rc = select_topk_dense(S, n_items, nq, n_items, n_next, 0, next_val + int64_t(q0) * n_next,
                               next_idx + int64_t(q0) * n_next, stream);
        if (rc != ANNCUR_OK) return rc;
    }
    return ANNCUR_OK;
}

}  // namespace anncur
