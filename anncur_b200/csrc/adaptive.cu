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

constexpr int AD_QB = 384;          // queries per block (the Cholesky kernel runs one CTA per query: ~3 per SM keep the SMs busy)

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

// G[b] = Mt[b] . Mt[b]^T  (m x m lower triangle, fp64), 64 x 64 output tile per CTA on the fp64 tensor cores
// (mma.sync m8n8k4 f64: 256 FMAs per warp instruction; a warp owns a 32 x 16 block = 4 x 2 MMA tiles, so one 4-wide k step
// is 6 shared-memory loads for 8 MMAs).  Operands are converted to fp64 once, on their way into shared memory; the row
// stride of 20 doubles makes the fragment loads conflict-free per half warp.
// G has row stride m and (m + 1) rows per query: row m is the right-hand side c of the solve (see below).
constexpr int GR_T = 64, GR_K = 16, GR_LD = GR_K + 4;
__global__ void __launch_bounds__(256)
gram_kernel(const float* __restrict__ Mt, int m, int k_q, double* __restrict__ G) {
    __shared__ __align__(16) double As[GR_T][GR_LD], Bs[GR_T][GR_LD];
    const int b = blockIdx.z;
    const int i0 = blockIdx.y * GR_T, j0 = blockIdx.x * GR_T;
    if (j0 > i0) return;                                   // lower triangle only
    const float* M = Mt + int64_t(b) * m * k_q;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wr = (warp >> 2) * 32, wc = (warp & 3) * 16; // this warp's block inside the tile
    const int fr = lane >> 2, fc = lane & 3;               // fragment row / k index (A), fragment column / k index (B)
    double acc[4][2][2];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 2; ++c) { acc[r][c][0] = 0.0; acc[r][c][1] = 0.0; }
    for (int k0 = 0; k0 < k_q; k0 += GR_K) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int e = threadIdx.x + s * 256;           // 0 .. 1023: 64 rows x 16 k, k fastest (as in global memory)
            const int r = e >> 4, kk = e & 15;
            As[r][kk] = (i0 + r < m && k0 + kk < k_q) ? double(M[int64_t(i0 + r) * k_q + k0 + kk]) : 0.0;
            Bs[r][kk] = (j0 + r < m && k0 + kk < k_q) ? double(M[int64_t(j0 + r) * k_q + k0 + kk]) : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k4 = 0; k4 < GR_K; k4 += 4) {
            double a[4], bb[2];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = As[wr + r * 8 + fr][k4 + fc];
#pragma unroll
            for (int c = 0; c < 2; ++c) bb[c] = Bs[wc + c * 8 + fr][k4 + fc];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                 : "+d"(acc[r][c][0]), "+d"(acc[r][c][1]) : "d"(a[r]), "d"(bb[c]));
        }
        __syncthreads();
    }
    double* Gb = G + int64_t(b) * (m + 1) * m;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = i0 + wr + r * 8 + fr, j = j0 + wc + c * 8 + fc * 2 + h;
                if (i < m && j <= i) Gb[int64_t(i) * m + j] = acc[r][c][h];
            }
}

// One CTA per query: blocked right-looking Cholesky G = L L^T of the lower triangle (panels of 32 columns held in shared
// memory), the solve G y = c, then e = y^T Mt (1 x k_q).
//  * The right-hand side rides along as row m of the matrix: after the factorisation that row holds z = L^-1 c (forward
//    substitution for free -- it takes part in the panel factorisations and is skipped by the trailing updates' diagonal).
//  * Back substitution L^T y = z walks the panels in reverse, again from shared memory.
//  * A pivot at or below rcond^2 * max diagonal is dropped: its column is zeroed and y_j = 0 (the min-norm rule of pinv
//    restricted to the kept coordinates).
// Per query: one pass over the trailing matrix per panel (m^3/6 fp64 FMAs from shared-memory operands), instead of one
// pass per COLUMN over global memory with strided accesses as in the first version (7.7 ms per 256 queries at m = 250).
constexpr int CH_NB = 32;
constexpr int CH_LD = CH_NB + 1;          // panel row stride in doubles (bank-conflict padding)

__global__ void __launch_bounds__(256)
cholesky_solve_kernel(double* __restrict__ G, const float* __restrict__ c, const float* __restrict__ Mt, int m, int k_q,
                      double rcond, double* __restrict__ ybuf, float* __restrict__ e_out) {
    extern __shared__ __align__(16) double ch_smem[];
    double* P = ch_smem;                                   // [(m + 1) rows][CH_LD]
    double* ys = ch_smem + size_t(m + 1) * CH_LD;          // [m] solution vector (back substitution)
    __shared__ double s_piv, s_maxd;
    __shared__ double red[8];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;
    double* L = G + int64_t(b) * (m + 1) * m;
    double* y = ybuf + int64_t(b) * m;
    // row m <- c
    for (int i = tid; i < m; i += 256) L[int64_t(m) * m + i] = double(c[int64_t(b) * m + i]);
    // max diagonal -> drop tolerance
    double md = 0.0;
    for (int i = tid; i < m; i += 256) md = fmax(md, L[int64_t(i) * m + i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) md = fmax(md, __shfl_xor_sync(0xffffffffu, md, o));
    if (tx == 0) red[ty] = md;
    __syncthreads();
    if (tid == 0) { double v = 0; for (int w = 0; w < 8; ++w) v = fmax(v, red[w]); s_maxd = v; }
    __syncthreads();
    const double drop = fmax(rcond * rcond, 1e-28) * s_maxd;   // pivots are squared singular-value scale

    for (int j0 = 0; j0 < m; j0 += CH_NB) {
        const int nb = m - j0 < CH_NB ? m - j0 : CH_NB;
        const int rows = m + 1 - j0;                        // panel rows: matrix rows j0 .. m (row m = right-hand side)
        // 1. panel -> shared memory (8 independent loads per thread in flight)
        for (int r0 = 0; r0 < rows; r0 += 64) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = r0 + ty + 8 * u;
                v[u] = (r < rows && tx < nb && (r >= tx || r >= nb)) ? __ldcg(L + int64_t(j0 + r) * m + j0 + tx) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = r0 + ty + 8 * u;
                if (r < rows && tx < nb) P[r * CH_LD + tx] = v[u];
            }
        }
        __syncthreads();
        // 2a. the nb x nb diagonal block, column by column (two barriers per column; every thread derives the pivot itself)
        for (int jj = 0; jj < nb; ++jj) {
            const double d = P[jj * CH_LD + jj];
            const double piv = d > drop ? sqrt(d) : 0.0;
            const double inv = piv > 0.0 ? 1.0 / piv : 0.0;        // dropped pivot: the column becomes 0
            __syncthreads();                                        // everyone has read d
            if (tid == 0) P[jj * CH_LD + jj] = piv;
            if (tid > jj && tid < nb) P[tid * CH_LD + jj] *= inv;
            __syncthreads();
            // P[r][t] -= P[r][jj] * P[t][jj] for jj < t <= r < nb
            const int t = jj + 1 + tx;
            if (t < nb) {
                const double ptj = P[t * CH_LD + jj];
                for (int r = t + ty; r < nb; r += 8) P[r * CH_LD + t] -= P[r * CH_LD + jj] * ptj;
            }
            __syncthreads();                                        // the next pivot was written by this update
        }
        __syncthreads();
        // 2b. the rows below it (incl. the right-hand side row): X L_d^T = P by forward substitution, one thread per row
        for (int r = nb + tid; r < rows; r += 256) {
            double* pr = P + r * CH_LD;
            for (int jj = 0; jj < nb; ++jj) {
                double acc = pr[jj];
                for (int k = 0; k < jj; ++k) acc = fma(-pr[k], P[jj * CH_LD + k], acc);
                const double piv = P[jj * CH_LD + jj];
                pr[jj] = piv > 0.0 ? acc / piv : 0.0;
            }
        }
        __syncthreads();
        // 3. panel back to global memory
        for (int r = ty; r < rows; r += 8)
            if (tx < nb && (r >= tx || r >= nb)) __stcg(L + int64_t(j0 + r) * m + j0 + tx, P[r * CH_LD + tx]);
        // 4. trailing update: L[i][t] -= sum_k P[i][k] P[t][k] for j0 + nb <= t <= i <= m (t < m).  64 x 64 tiles, each thread a
        //    4 x 4 block with rows / columns 16 apart (conflict-free panel reads at row stride 33): 8 LDS feed 16 DFMA per k.
        const int base = j0 + nb;
        const int n_tr = m + 1 - base;                      // trailing rows (incl. the right-hand side row)
        const int n_tc = m - base;                          // trailing columns
        const int sx = tid & 15, sy = tid >> 4;
        for (int ti = 0; ti < n_tr; ti += 64) {
            for (int tt = 0; tt <= ti && tt < n_tc; tt += 64) {
                const double* pr[4];
                const double* pc[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int ri = ti + sy + 16 * q, ci = tt + sx + 16 * q;
                    pr[q] = P + (nb + (ri < n_tr ? ri : 0)) * CH_LD;       // out-of-range rows / columns compute garbage that is never stored
                    pc[q] = P + (nb + (ci < n_tc ? ci : 0)) * CH_LD;
                }
                // the 16 old values are fetched before the k loop (their latency hides behind it; as read-modify-writes in one
                // statement each, the compiler has to serialise them: 16 L2 round trips per tile)
                double acc[4][4], old[4][4];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int ri = ti + sy + 16 * r, ci = tt + sx + 16 * cc;
                        acc[r][cc] = 0.0;
                        old[r][cc] = (ri < n_tr && ci < n_tc && ci <= ri) ? __ldcg(L + int64_t(base + ri) * m + base + ci) : 0.0;
                    }
                for (int k = 0; k < nb; ++k) {
                    double a[4], bb[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) { a[q] = pr[q][k]; bb[q] = pc[q][k]; }
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) acc[r][cc] = fma(a[r], bb[cc], acc[r][cc]);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int ri = ti + sy + 16 * r;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int ci = tt + sx + 16 * cc;
                        if (ri < n_tr && ci < n_tc && ci <= ri) __stcg(L + int64_t(base + ri) * m + base + ci, old[r][cc] - acc[r][cc]);
                    }
                }
            }
        }
        __syncthreads();
    }
    // back substitution L^T y = z (z = row m), panels in reverse
    for (int i = tid; i < m; i += 256) ys[i] = 0.0;
    __syncthreads();
    const int n_panels = (m + CH_NB - 1) / CH_NB;
    for (int pi = n_panels - 1; pi >= 0; --pi) {
        const int j0 = pi * CH_NB;
        const int nb = m - j0 < CH_NB ? m - j0 : CH_NB;
        const int rows = m + 1 - j0;
        for (int r0 = 0; r0 < rows; r0 += 64) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = r0 + ty + 8 * u;
                v[u] = (r < rows && tx < nb && (r >= tx || r >= nb)) ? __ldcg(L + int64_t(j0 + r) * m + j0 + tx) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = r0 + ty + 8 * u;
                if (r < rows && tx < nb) P[r * CH_LD + tx] = v[u];
            }
        }
        __syncthreads();
        // rhs_k = z_k - sum_{i >= j0 + nb} L[i][k] y_i : one warp per column k (stride 8), lanes over the rows
        for (int k = ty; k < nb; k += 8) {
            double part = 0.0;
            for (int r = nb + tx; r < rows - 1; r += 32) part = fma(P[r * CH_LD + k], ys[j0 + r], part);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (tx == 0) P[(rows - 1) * CH_LD + k] -= part;           // z_k (row m of the panel) becomes the block's right-hand side
        }
        __syncthreads();
        // the nb x nb triangle, backwards, by one warp
        if (ty == 0) {
            for (int jj = nb - 1; jj >= 0; --jj) {
                const double piv = P[jj * CH_LD + jj];
                const double yj = piv > 0.0 ? P[(rows - 1) * CH_LD + jj] / piv : 0.0;
                __syncwarp();
                if (tx == 0) ys[j0 + jj] = yj;
                if (tx < jj) P[(rows - 1) * CH_LD + tx] -= P[jj * CH_LD + tx] * yj;
                __syncwarp();
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < m; i += 256) y[i] = ys[i];
    // e = y^T Mt
    const float* M = Mt + int64_t(b) * m * k_q;
    for (int t = tid; t < k_q; t += 256) {
        double acc = 0.0;
        for (int i = 0; i < m; ++i) acc = fma(ys[i], double(M[int64_t(i) * k_q + t]), acc);
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
    p.off_g = off; off += align_up(sizeof(double) * size_t(AD_QB) * (m + 1) * m, 256);
    p.off_y = off; off += align_up(sizeof(double) * size_t(AD_QB) * m, 256);
    p.off_e = off; off += align_up(sizeof(float) * size_t(AD_QB) * k_q, 256);
    p.off_s = off; off += align_up(sizeof(float) * size_t(AD_QB) * n_items, 256);
    p.total = off;
    return p;
}

// ---- split form: the per-query solve alone, and the anchor filter of a fused re-score --------------------------------------
// adaptive_solve leaves e_b = c_b . pinv(R_anc[:, I_b]) (n_queries x k_q fp32); the caller re-scores with the fused tensor-core
// kernel on the PACKED R_anc (anncur_score_topk with k = n_next + m: no B x N block, shardable) and drops the anchors from the
// candidate lists with filter_excluded -- "masked re-score" without a mask in the inner loop: a row's m anchors can displace
// at most m of the best n_next + m candidates.
static AdaptivePlan adaptive_solve_plan(int k_q, int m, int64_t n_items) {
    AdaptivePlan p{};
    size_t off = 0;
    p.off_rt = off; off += align_up(sizeof(float) * size_t(n_items) * k_q, 256);
    p.off_mt = off; off += align_up(sizeof(float) * size_t(AD_QB) * m * k_q, 256);
    p.off_g = off; off += align_up(sizeof(double) * size_t(AD_QB) * (m + 1) * m, 256);
    p.off_y = off; off += align_up(sizeof(double) * size_t(AD_QB) * m, 256);
    p.off_e = off; p.off_s = off;
    p.total = off;
    return p;
}

size_t adaptive_solve_workspace_bytes(int n_queries, int k_q, int m, int64_t n_items) {
    (void)n_queries;
    if (k_q <= 0 || m <= 0 || n_items <= 0) return 256;
    return adaptive_solve_plan(k_q, m, n_items).total;
}

// Rt_cached != nullptr: the item-major copy R_anc^T (n_items x k_q fp32) kept by the caller across rounds (it does not
// change); otherwise it is rebuilt in the workspace.
int adaptive_solve(const float* R_anc, int64_t ldr, int k_q, int64_t n_items, const float* Rt_cached, const int64_t* anchors,
                   const float* c, int n_queries, int m, double rcond, float* e_out, void* workspace, size_t workspace_bytes,
                   cudaStream_t stream) {
    if (m > k_q) { set_error("adaptive_solve: m = %d anchors > k_q = %d anchor queries is not supported", m, k_q); return ANNCUR_E_UNSUPPORTED; }
    const AdaptivePlan pl = adaptive_solve_plan(k_q, m, n_items);
    if (workspace_bytes < pl.total) { set_error("adaptive_solve workspace too small: %zu < %zu", workspace_bytes, pl.total); return ANNCUR_E_WORKSPACE; }
    char* ws = reinterpret_cast<char*>(workspace);
    float* Rt = reinterpret_cast<float*>(ws + pl.off_rt);
    float* Mt = reinterpret_cast<float*>(ws + pl.off_mt);
    double* G = reinterpret_cast<double*>(ws + pl.off_g);
    double* y = reinterpret_cast<double*>(ws + pl.off_y);
    const size_t ch_smem = sizeof(double) * (size_t(m + 1) * CH_LD + size_t(m));
    if (ch_smem > 200 * 1024) { set_error("adaptive_solve: m = %d anchors need %zu bytes of shared memory per query", m, ch_smem); return ANNCUR_E_UNSUPPORTED; }
    ANNCUR_CUDA_OK(cudaFuncSetAttribute(cholesky_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(ch_smem)));
    if (Rt_cached == nullptr) {
        dim3 tgrid(unsigned((n_items + 31) / 32), unsigned((k_q + 31) / 32));
        transpose_kernel<<<tgrid, dim3(32, 8), 0, stream>>>(R_anc, ldr, k_q, n_items, Rt);
        ANNCUR_LAUNCH_OK("transpose_kernel");
    }
    const float* Rt_use = Rt_cached ? Rt_cached : Rt;
    for (int q0 = 0; q0 < n_queries; q0 += AD_QB) {
        const int nq = n_queries - q0 < AD_QB ? n_queries - q0 : AD_QB;
        const int64_t* anc = anchors + int64_t(q0) * m;
        const int64_t warps = int64_t(nq) * m;
        gather_anchor_rows_kernel<<<unsigned((warps * 32 + 255) / 256), 256, 0, stream>>>(Rt_use, k_q, anc, m, nq, Mt);
        ANNCUR_LAUNCH_OK("gather_anchor_rows_kernel");
        dim3 ggrid(unsigned((m + GR_T - 1) / GR_T), unsigned((m + GR_T - 1) / GR_T), unsigned(nq));
        gram_kernel<<<ggrid, 256, 0, stream>>>(Mt, m, k_q, G);
        ANNCUR_LAUNCH_OK("gram_kernel");
        cholesky_solve_kernel<<<nq, 256, ch_smem, stream>>>(G, c + int64_t(q0) * m, Mt, m, k_q, rcond, y, e_out + int64_t(q0) * k_q);
        ANNCUR_LAUNCH_OK("cholesky_solve_kernel");
    }
    return ANNCUR_OK;
}

// out[row] = the first n_out entries of cand[row] (best first; idx < 0 = padding) whose index is not in excl[row]; short rows
// are padded with (ANNCUR_PAD_VAL, -1).  One warp per row, order kept.  The row's excluded indices go into an open-addressing
// hash table in shared memory (>= 2 m slots, linear probing), so a candidate costs one or two probes instead of a scan of
// all m entries (m = 375 at BASELINE configs[2]: 0.145 -> 0.0x ms per 1024 rows).
__global__ void __launch_bounds__(256)
filter_excluded_kernel(const float* __restrict__ cand_vals, const int64_t* __restrict__ cand_idx, int n_rows, int k_in,
                       const int64_t* __restrict__ excl, int m, int n_out, int slots, float* __restrict__ out_vals,
                       int64_t* __restrict__ out_idx) {
    extern __shared__ unsigned long long ex_tab[];             // [warps][slots], empty = ~0
    const int warp = threadIdx.x >> 5, lane = int(lane_id());
    const int row = blockIdx.x * (blockDim.x >> 5) + warp;
    if (row >= n_rows) return;
    unsigned long long* tab = ex_tab + size_t(warp) * slots;
    constexpr unsigned long long EMPTY = ~0ull;
    const unsigned mask = unsigned(slots - 1);
    for (int t = lane; t < slots; t += 32) tab[t] = EMPTY;
    __syncwarp();
    for (int t = lane; t < m; t += 32) {
        const unsigned long long key = (unsigned long long)excl[int64_t(row) * m + t];
        unsigned h = (unsigned(key) * 2654435761u >> 7) & mask;
        while (true) {
            const unsigned long long old = atomicCAS(&tab[h], EMPTY, key);
            if (old == EMPTY || old == key) break;
            h = (h + 1) & mask;
        }
    }
    __syncwarp();
    int n_done = 0;
    for (int t0 = 0; t0 < k_in && n_done < n_out; t0 += 32) {
        const int t = t0 + lane;
        const int64_t i = t < k_in ? cand_idx[int64_t(row) * k_in + t] : -1;
        bool keep = i >= 0;
        if (keep) {
            unsigned h = (unsigned(uint64_t(i)) * 2654435761u >> 7) & mask;
            while (true) {
                const unsigned long long v = tab[h];
                if (v == EMPTY) break;
                if (v == (unsigned long long)i) { keep = false; break; }
                h = (h + 1) & mask;
            }
        }
        const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
        const int pos = n_done + __popc(ballot & ((1u << lane) - 1u));
        if (keep && pos < n_out) {
            out_vals[int64_t(row) * n_out + pos] = cand_vals[int64_t(row) * k_in + t];
            out_idx[int64_t(row) * n_out + pos] = i;
        }
        n_done += __popc(ballot);
    }
    for (int t = (n_done < n_out ? n_done : n_out) + lane; t < n_out; t += 32) {
        out_vals[int64_t(row) * n_out + t] = ANNCUR_PAD_VAL;
        out_idx[int64_t(row) * n_out + t] = -1;
    }
}

int filter_excluded(const float* cand_vals, const int64_t* cand_idx, int n_rows, int k_in, const int64_t* excl, int m, int n_out,
                    float* out_vals, int64_t* out_idx, cudaStream_t stream) {
    if (n_rows == 0) return ANNCUR_OK;
    int slots = 64;
    while (slots < 2 * m) slots <<= 1;
    if (size_t(slots) * 8 > 200 * 1024) { set_error("filter_excluded: %d excluded indices per row do not fit shared memory", m); return ANNCUR_E_UNSUPPORTED; }
    int warps = 8;
    while (warps > 1 && size_t(warps) * slots * 8 > 96 * 1024) warps >>= 1;
    const size_t smem = sizeof(unsigned long long) * size_t(warps) * slots;
    ANNCUR_CUDA_OK(cudaFuncSetAttribute(filter_excluded_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    filter_excluded_kernel<<<(n_rows + warps - 1) / warps, warps * 32, smem, stream>>>(cand_vals, cand_idx, n_rows, k_in, excl, m, n_out, slots,
                                                                                     out_vals, out_idx);
    ANNCUR_LAUNCH_OK("filter_excluded_kernel");
    return ANNCUR_OK;
}

int transpose_rows(const float* in, int64_t ld_in, int rows, int64_t cols, float* out, cudaStream_t stream) {
    dim3 tgrid(unsigned((cols + 31) / 32), unsigned((rows + 31) / 32));
    transpose_kernel<<<tgrid, dim3(32, 8), 0, stream>>>(in, ld_in, rows, cols, out);
    ANNCUR_LAUNCH_OK("transpose_kernel");
    return ANNCUR_OK;
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

    const size_t ch_smem = sizeof(double) * (size_t(m + 1) * CH_LD + size_t(m));
    if (ch_smem > 200 * 1024) { set_error("adaptive_round: m = %d anchors need %zu bytes of shared memory per query", m, ch_smem); return ANNCUR_E_UNSUPPORTED; }
    ANNCUR_CUDA_OK(cudaFuncSetAttribute(cholesky_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(ch_smem)));
    dim3 tgrid(unsigned((n_items + 31) / 32), unsigned((k_q + 31) / 32));
    transpose_kernel<<<tgrid, dim3(32, 8), 0, stream>>>(R_anc, ldr, k_q, n_items, Rt);
    ANNCUR_LAUNCH_OK("transpose_kernel");
    for (int q0 = 0; q0 < n_queries; q0 += AD_QB) {
        const int nq = n_queries - q0 < AD_QB ? n_queries - q0 : AD_QB;
        const int64_t* anc = anchors + int64_t(q0) * m;
        const int64_t warps = int64_t(nq) * m;
        gather_anchor_rows_kernel<<<unsigned((warps * 32 + 255) / 256), 256, 0, stream>>>(Rt, k_q, anc, m, nq, Mt);
        ANNCUR_LAUNCH_OK("gather_anchor_rows_kernel");
        dim3 ggrid(unsigned((m + GR_T - 1) / GR_T), unsigned((m + GR_T - 1) / GR_T), unsigned(nq));
        gram_kernel<<<ggrid, 256, 0, stream>>>(Mt, m, k_q, G);
        ANNCUR_LAUNCH_OK("gram_kernel");
        cholesky_solve_kernel<<<nq, 256, ch_smem, stream>>>(G, c + int64_t(q0) * m, Mt, m, k_q, rcond, y, e);
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
