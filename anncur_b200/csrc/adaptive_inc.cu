// K8 (incremental form): adaptive multi-round ANNCUR with the per-query factorisation CARRIED ACROSS ROUNDS.
// NOT in the reference (spec = SURVEY.md section 8a row A8, parity unpinned -- checked against
// oracle.cur_oracle.adaptive_anncur).
//
// Round t solves e_b = c_b . pinv(M_b), M_b = R_anc[:, I_b] (k_q x m), through the normal equations G_b = M_b^T M_b = L L^T
// in fp64.  adaptive.cu rebuilds G_b and L from scratch every round (m = 125, 250, 375: 218 MFLOP of Gram + 23 MFLOP of
// Cholesky per query at BASELINE configs[2]).  Here the anchor set only GROWS, I_{t+1} = I_t U new_t, and round 1 is shared
// by all queries, so with L_t = [[L_{t-1}, 0], [X, L_D]]:
//     X   = B^T L_{t-1}^-T        B = M_{t-1}^T N   (N = the n new columns)
//     L_D = chol(N^T N - X X^T)   z_t = [z_{t-1}, L_D^-1 (c_new - X z_{t-1})]        (forward substitution is incremental)
//     y   = L_t^-T z_t            e = (M y)^T
//  * the columns of X that belong to the SHARED first anchors are a gather: X_a = W_1[:, new], W_1 = L_1^-1 M_1^T R_anc
//    (m_1 x N, built once per index by anncur_adaptive_prepare);
//  * the n x n diagonal block is factorised right-looking in SUB-BLOCKS of <= 32 columns and the 32 x 32 inverse of every
//    diagonal sub-block is kept, so every other step -- block substitution against earlier rounds, Schur complements,
//    panel and trailing updates -- is a batched fp64 GEMM C -= A B^T on the fp64 tensor cores (DMMA 8x8x4): one generic
//    kernel, operands either fp64 panels or fp32 rows of R_anc^T gathered by item index;
//  * two small per-query kernels complete a round: a warp-level Cholesky + triangular inverse of a 32 x 32 sub-block
//    (registers + shared-memory broadcasts), and the substitutions z, y and e = (M y)^T;
//  * per query and round: 45 MFLOP in total instead of 241.  The CTA-per-query Cholesky + inverse of a whole block
//    (chol_inv_kernel, n <= 128) is used once, for the shared anchors.
// Pivots at or below rcond^2 * (largest Gram diagonal) are dropped exactly as in adaptive.cu (that anchor's coordinate of y
// is 0): zero pivot, zero column of L, zero row and column of the block inverse.
#include "common.cuh"
#include "kernels.h"

namespace anncur {

// ---------------------------------------------------------------------------------------------------------------------------
// Generic batched C_out = C_in + sign * A B^T on the fp64 tensor cores.  64 x 64 tile per CTA (8 warps, each 32 x 16 = 4 x 2
// DMMA tiles), 32-wide k stages: the next stage's global loads are issued before the current stage's MMAs (registers), so the
// L2 latency hides behind 64 DMMAs per warp.  Row stride 36 doubles: conflict-free fragment loads per half warp.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int DG_T = 64, DG_K = 32, DG_LD = DG_K + 4;
constexpr int DG_ALL = 1 << 29;                           // diag_off: "no triangular skipping"

struct DgArgs {
    // GATHER mode: row i of A = fp32 row idxA[i] of Rt (row length K, row stride ld_rt); idx == nullptr: row i itself
    const float* Rt; int64_t ld_rt; int64_t n_rt;           // n_rt rows: an index outside [0, n_rt) reads as a zero row
    const int64_t* idxA; int64_t bs_idxA;
    const int64_t* idxB; int64_t bs_idxB;
    // fp64 mode
    const double* A; int64_t lda, bsA;
    const double* B; int64_t ldb, bsB;
    const double* Cin; int64_t ldcin, bsCin;               // nullptr: zero
    double* Cout; int64_t ldc, bsC;
    int M, N, K;
    double sign;
    int diag_off;                                          // entries with j > i + diag_off are neither computed nor stored
    int vec4;                                              // GATHER: rows are 16-byte aligned and K % 4 == 0
};

// TN = 64: 256 threads, 8 warps (2 x 4).  TN = 32 (fp64 operands only): the narrow launches of the block algorithms (N <= 32:
// a 32-wide sub-block) on 64 x 32 tiles with 128 threads, 4 warps (2 x 2) -- no half-empty tile, twice the CTAs per SM.
template <bool GATHER, int TN>
__global__ void __launch_bounds__(TN == 64 ? 256 : 128, TN == 64 ? (GATHER ? 3 : 2) : 4)
dgemm_nt_kernel(const DgArgs p) {
    static_assert(TN == 64 || (TN == 32 && !GATHER), "tile width");
    constexpr int NT = TN == 64 ? 256 : 128;               // threads
    constexpr int TA = NT / DG_T, KA = DG_K / TA;          // loader threads per row of A, k values per thread (8 or 16)
    __shared__ __align__(16) double As[DG_T][DG_LD], Bs[TN][DG_LD];
    const int b = blockIdx.z;
    const int i0 = blockIdx.x * DG_T, j0 = blockIdx.y * TN;          // rows on grid.x: no 65535-tile limit on M
    if (int64_t(j0) > int64_t(i0) + DG_T - 1 + p.diag_off) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wr = TN == 64 ? (warp >> 2) * 32 : (warp >> 1) * 32, wc = TN == 64 ? (warp & 3) * 16 : (warp & 1) * 16;
    const int fr = lane >> 2, fc = lane & 3;
    const int lrow = tid / TA, lq = tid % TA;              // loader, A: tile row lrow, k = lq + TA u
    const int lrowb = tid >> 2, lqb = tid & 3;             // loader, B: tile row lrowb (< TN), k = lqb + 4 u
    bool va = i0 + lrow < p.M, vb = j0 + lrowb < p.N;

    const float* fa = nullptr; const float* fb = nullptr;
    const double* da = nullptr; const double* db = nullptr;
    if (GATHER) {
        int64_t ia = va ? (p.idxA ? p.idxA[int64_t(b) * p.bs_idxA + i0 + lrow] : int64_t(i0 + lrow)) : 0;
        int64_t ib = vb ? (p.idxB ? p.idxB[int64_t(b) * p.bs_idxB + j0 + lrowb] : int64_t(j0 + lrowb)) : 0;
        if (ia < 0 || ia >= p.n_rt) { va = false; ia = 0; }          // invalid anchor: a zero column (its pivot is dropped)
        if (ib < 0 || ib >= p.n_rt) { vb = false; ib = 0; }
        fa = p.Rt + ia * p.ld_rt; fb = p.Rt + ib * p.ld_rt;
    } else {
        da = p.A + int64_t(b) * p.bsA + int64_t(va ? i0 + lrow : 0) * p.lda;
        db = p.B + int64_t(b) * p.bsB + int64_t(vb ? j0 + lrowb : 0) * p.ldb;
    }
    float4 ga[2], gb[2];                                   // GATHER stage registers: k = 4 lq + {0, 16} .. +3
    double ra[KA], rb[8];                                  // fp64 stage registers:   k = lq + TA u (A), lqb + 4 u (B)
    auto load_stage = [&](int k0) {
        if (GATHER) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = k0 + 4 * lq + 16 * h;
                float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
                if (p.vec4) {
                    if (va && k < p.K) x = __ldg(reinterpret_cast<const float4*>(fa + k));
                    if (vb && k < p.K) y = __ldg(reinterpret_cast<const float4*>(fb + k));
                } else {
                    if (va) { if (k < p.K) x.x = __ldg(fa + k); if (k + 1 < p.K) x.y = __ldg(fa + k + 1);
                              if (k + 2 < p.K) x.z = __ldg(fa + k + 2); if (k + 3 < p.K) x.w = __ldg(fa + k + 3); }
                    if (vb) { if (k < p.K) y.x = __ldg(fb + k); if (k + 1 < p.K) y.y = __ldg(fb + k + 1);
                              if (k + 2 < p.K) y.z = __ldg(fb + k + 2); if (k + 3 < p.K) y.w = __ldg(fb + k + 3); }
                }
                ga[h] = x; gb[h] = y;
            }
        } else {
#pragma unroll
            for (int u = 0; u < KA; ++u) {
                const int k = k0 + lq + TA * u;
                ra[u] = (va && k < p.K) ? da[k] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int k = k0 + lqb + 4 * u;
                rb[u] = (vb && k < p.K) ? db[k] : 0.0;
            }
        }
    };
    auto store_stage = [&]() {
        if (GATHER) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double* pa = &As[lrow][4 * lq + 16 * h];
                double* pb = &Bs[lrow][4 * lq + 16 * h];
                pa[0] = double(ga[h].x); pa[1] = double(ga[h].y); pa[2] = double(ga[h].z); pa[3] = double(ga[h].w);
                pb[0] = double(gb[h].x); pb[1] = double(gb[h].y); pb[2] = double(gb[h].z); pb[3] = double(gb[h].w);
            }
        } else {
#pragma unroll
            for (int u = 0; u < KA; ++u) As[lrow][lq + TA * u] = ra[u];
#pragma unroll
            for (int u = 0; u < 8; ++u) Bs[lrowb][lqb + 4 * u] = rb[u];
        }
    };

    double acc[4][2][2];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 2; ++c) { acc[r][c][0] = 0.0; acc[r][c][1] = 0.0; }

    if (p.K > 0) load_stage(0);
    for (int k0 = 0; k0 < p.K; k0 += DG_K) {
        store_stage();
        __syncthreads();
        if (k0 + DG_K < p.K) load_stage(k0 + DG_K);
#pragma unroll
        for (int k4 = 0; k4 < DG_K; k4 += 4) {
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
    // epilogue: all old values are fetched before the first store (C_in may alias C_out, so the compiler would otherwise
    // keep every load behind the previous store: 16 L2 round trips in a row)
    const double* Cin = p.Cin ? p.Cin + int64_t(b) * p.bsCin : nullptr;
    double* Cout = p.Cout + int64_t(b) * p.bsC;
    double old[4][2][2];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = i0 + wr + r * 8 + fr, j = j0 + wc + c * 8 + fc * 2 + h;
                const bool ok = i < p.M && j < p.N && int64_t(j) <= int64_t(i) + p.diag_off;
                old[r][c][h] = (Cin && ok) ? __ldcg(Cin + int64_t(i) * p.ldcin + j) : 0.0;
            }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = i0 + wr + r * 8 + fr, j = j0 + wc + c * 8 + fc * 2 + h;
                if (i < p.M && j < p.N && int64_t(j) <= int64_t(i) + p.diag_off)
                    Cout[int64_t(i) * p.ldc + j] = old[r][c][h] + p.sign * acc[r][c][h];
            }
}

static int dgemm_launch(const DgArgs& a, bool gather, int batch, cudaStream_t stream) {
    if (a.M <= 0 || a.N <= 0 || batch <= 0) return ANNCUR_OK;
    const bool narrow = !gather && a.N <= 32;
    const int tn = narrow ? 32 : DG_T;
    dim3 grid(unsigned((a.M + DG_T - 1) / DG_T), unsigned((a.N + tn - 1) / tn), unsigned(batch));
    if (gather) dgemm_nt_kernel<true, 64><<<grid, 256, 0, stream>>>(a);
    else if (narrow) dgemm_nt_kernel<false, 32><<<grid, 128, 0, stream>>>(a);
    else dgemm_nt_kernel<false, 64><<<grid, 256, 0, stream>>>(a);
    ANNCUR_LAUNCH_OK("dgemm_nt_kernel");
    return ANNCUR_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Per query: in-place Cholesky of the n x n Schur complement S (lower triangle, n <= 128), the inverse of its factor, and the
// new block of z.  Both triangles live PACKED in shared memory (index i (i + 1) / 2 + j: the accesses of both loops below
// are conflict-free), 2 * n (n + 1) / 2 doubles = 126 KB at n = 125.
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }

__global__ void __launch_bounds__(256)
chol_inv_kernel(double* __restrict__ Sbase, int64_t bsS, int64_t lds,                   // in: S (lower), out: L_D (lower)
                const double* __restrict__ rawdiag, int64_t bs_rd, int64_t ld_rd,       // raw Gram block (its diagonal = |column|^2)
                double* __restrict__ maxd,                                              // [batch] running max of the Gram diagonal (in / out)
                double* __restrict__ Linv, int64_t bsLinv,                              // out: L_D^-1, n x n row-major, upper triangle zeroed
                int n, double rcond) {
    extern __shared__ __align__(16) double ci_smem[];
    const int ntri = n * (n + 1) / 2;
    double* Ls = ci_smem;                    // [ntri]
    double* Li = Ls + ntri;                  // [ntri]
    double* pivs = Li + ntri;                // [n]
    double* invp = pivs + n;                 // [n]
    __shared__ double red[8];
    __shared__ double s_maxd;
    const int b = blockIdx.x, tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    double* S = Sbase + int64_t(b) * bsS;
    for (int i = ty; i < n; i += 8)
        for (int j = tx; j <= i; j += 32) Ls[tri(i, j)] = S[int64_t(i) * lds + j];
    double md = 0.0;
    for (int i = tid; i < n; i += 256) md = fmax(md, rawdiag[int64_t(b) * bs_rd + int64_t(i) * ld_rd + i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) md = fmax(md, __shfl_xor_sync(0xffffffffu, md, o));
    if (tx == 0) red[ty] = md;
    __syncthreads();
    if (tid == 0) {
        double v = maxd[b];
        for (int w = 0; w < 8; ++w) v = fmax(v, red[w]);
        s_maxd = v; maxd[b] = v;
    }
    __syncthreads();
    const double drop = fmax(rcond * rcond, 1e-28) * s_maxd;

    // Cholesky, column by column; the diagonal keeps its raw value until the end (pivs holds the pivots)
    for (int j = 0; j < n; ++j) {
        __syncthreads();                                              // trailing update of column j - 1 is complete
        const double d = Ls[tri(j, j)];
        const double piv = d > drop ? sqrt(d) : 0.0;
        const double inv = piv > 0.0 ? 1.0 / piv : 0.0;              // dropped pivot: the column becomes 0
        for (int i = j + 1 + tid; i < n; i += 256) Ls[tri(i, j)] *= inv;
        if (tid == 0) { pivs[j] = piv; invp[j] = inv; }
        __syncthreads();
        for (int t = j + 1 + tx; t < n; t += 32) {
            const double ptj = Ls[tri(t, j)];
            for (int r = t + ty; r < n; r += 8) Ls[tri(r, t)] -= Ls[tri(r, j)] * ptj;
        }
    }
    __syncthreads();
    for (int j = tid; j < n; j += 256) Ls[tri(j, j)] = pivs[j];
    __syncthreads();
    // inverse of the factor, one thread per column: x_j = 1 / L_jj, x_i = -(sum_{k = j}^{i-1} L_ik x_k) / L_ii
    if (tid < n) {
        const int j = tid;
        Li[tri(j, j)] = invp[j];
        for (int i = j + 1; i < n; ++i) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int k = j;
            for (; k + 3 < i; k += 4) {
                a0 = fma(Ls[tri(i, k)], Li[tri(k, j)], a0);
                a1 = fma(Ls[tri(i, k + 1)], Li[tri(k + 1, j)], a1);
                a2 = fma(Ls[tri(i, k + 2)], Li[tri(k + 2, j)], a2);
                a3 = fma(Ls[tri(i, k + 3)], Li[tri(k + 3, j)], a3);
            }
            for (; k < i; ++k) a0 = fma(Ls[tri(i, k)], Li[tri(k, j)], a0);
            Li[tri(i, j)] = -((a0 + a1) + (a2 + a3)) * invp[i];
        }
    }
    __syncthreads();
    double* Lo = Linv + int64_t(b) * bsLinv;
    for (int i = ty; i < n; i += 8) {
        for (int j = tx; j < n; j += 32) {
            if (j <= i) S[int64_t(i) * lds + j] = Ls[tri(i, j)];
            Lo[int64_t(i) * n + j] = j <= i ? Li[tri(i, j)] : 0.0;
        }
    }
}

// The same for a block of w <= 32 columns, ONE WARP per query: lane i holds row i of the block in registers, finished rows of
// L sit in the warp's slice of shared memory and are read back as broadcasts (one LDS per operand instead of two shuffles),
// all loops are unrolled (rows / columns >= w are padded with the identity).  Left-looking Cholesky (column j = one dot
// product of length j per lane), then the inverse column by column (lane = column).  ~1000 LDS + ~1000 DFMA per block on the
// dependency chain (the shuffle version: 8300 instructions per warp at 11 cycles each = 50 us per launch).
constexpr int CB = 32;
constexpr int CB_LD = CB + 1;
__global__ void __launch_bounds__(128)
chol_inv32_kernel(double* __restrict__ Sbase, int64_t bsS, int64_t lds, const double* __restrict__ rawdiag, int64_t bs_rd, int64_t ld_rd,
                  double* __restrict__ maxd, double* __restrict__ Linv, int64_t bsLinv, int w, double rcond, int n_queries) {
    __shared__ double Ls_all[4][CB][CB_LD];
    __shared__ double inv_all[4][CB];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int b = blockIdx.x * 4 + wp;
    if (b >= n_queries) return;
    constexpr uint32_t FULL = 0xffffffffu;
    double (*Ls)[CB_LD] = Ls_all[wp];
    double* invd = inv_all[wp];
    double* S = Sbase + int64_t(b) * bsS;
    // coalesced load of the lower triangle, transposed through shared memory into "lane = row" registers
#pragma unroll 4
    for (int r = 0; r < CB; ++r) Ls[r][lane] = (r < w && lane <= r) ? S[int64_t(r) * lds + lane] : (r == lane ? 1.0 : 0.0);
    double md = lane < w ? rawdiag[int64_t(b) * bs_rd + int64_t(lane) * ld_rd + lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) md = fmax(md, __shfl_xor_sync(FULL, md, o));
    md = fmax(md, maxd[b]);
    __syncwarp();
    if (lane == 0) maxd[b] = md;
    const double drop = fmax(rcond * rcond, 1e-28) * md;
    double a[CB];
#pragma unroll
    for (int j = 0; j < CB; ++j) a[j] = Ls[lane][j];
    __syncwarp();
    // Cholesky, left-looking: L[i][j] = (A[i][j] - sum_{k<j} L[i][k] L[j][k]) / L[j][j]
#pragma unroll
    for (int j = 0; j < CB; ++j) {
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
        for (int k = 0; k < j; ++k) {
            const double ljk = Ls[j][k];                          // broadcast: row j is final up to column j - 1
            if (k & 1) acc1 = fma(a[k], ljk, acc1); else acc0 = fma(a[k], ljk, acc0);
        }
        const double v = a[j] - (acc0 + acc1);
        const double d = __shfl_sync(FULL, v, j);
        const double piv = j < w ? (d > drop ? sqrt(d) : 0.0) : 1.0;
        const double inv = piv > 0.0 ? 1.0 / piv : 0.0;          // dropped pivot: the column becomes 0
        double l = v * inv;
        if (lane == j) { l = piv; invd[j] = inv; }
        if (lane < j) l = 0.0;
        a[j] = l;
        Ls[lane][j] = l;
        __syncwarp();
    }
    // inverse, lane = column c: x_c = 1 / L_cc, x_i = -(sum_{k=c}^{i-1} L_ik x_k) / L_ii   (x_k = 0 for k < c)
    double x[CB];
#pragma unroll
    for (int i = 0; i < CB; ++i) {
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
        for (int k = 0; k < i; ++k) {
            const double lik = Ls[i][k];
            if (k & 1) acc1 = fma(lik, x[k], acc1); else acc0 = fma(lik, x[k], acc0);
        }
        const double invi = invd[i];
        x[i] = i == lane ? invi : (i > lane ? -(acc0 + acc1) * invi : 0.0);
    }
    double* Lo = Linv + int64_t(b) * bsLinv;
#pragma unroll
    for (int i = 0; i < CB; ++i)
        if (i < w && lane < w) Lo[i * CB + lane] = x[i];
#pragma unroll 4
    for (int r = 0; r < CB; ++r)
        if (r < w && lane <= r) S[int64_t(r) * lds + lane] = Ls[r][lane];
}

// z_0 = L_1^-1 c_1 for every query (the shared first block), and the running Gram-diagonal maximum starts at the shared one.
__global__ void __launch_bounds__(128)
z_shared_kernel(const double* __restrict__ L1inv, int s, const double* __restrict__ maxd_shared, const float* __restrict__ c,
                double* __restrict__ z, int64_t bsz, double* __restrict__ maxd) {
    extern __shared__ double zc[];
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < s; i += blockDim.x) zc[i] = double(c[int64_t(b) * s + i]);
    if (threadIdx.x == 0) maxd[b] = s > 0 ? *maxd_shared : 0.0;
    __syncthreads();
    for (int i = threadIdx.x; i < s; i += blockDim.x) {
        double a = 0.0;
        for (int j = 0; j <= i; ++j) a = fma(L1inv[int64_t(i) * s + j], zc[j], a);
        z[int64_t(b) * bsz + i] = a;
    }
}

// The new anchors join the query's list, and the columns of their rows of L that belong to the shared anchors are gathered
// from W_1^T (one warp per new anchor).
__global__ void __launch_bounds__(256)
append_anchors_kernel(const int64_t* __restrict__ new_anchors, int n, int n_queries, int r_done, int64_t n_items,
                      int64_t* __restrict__ anc_q, int64_t bs_anc, const double* __restrict__ W1t, int s,
                      double* __restrict__ Lq, int64_t bsLq, int ldl) {
    const int64_t w = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
    if (w >= int64_t(n_queries) * n) return;
    const int b = int(w / n), i = int(w % n);
    const int64_t item = new_anchors[w];
    if (lane_id() == 0) anc_q[int64_t(b) * bs_anc + int64_t(r_done) * n + i] = item;
    const bool valid = item >= 0 && item < n_items;                  // invalid index (e.g. the -1 padding of a short list): zero row
    const double* src = W1t + (valid ? item : 0) * s;
    double* dst = Lq + int64_t(b) * bsLq + (int64_t(r_done) * n + i) * ldl;
    for (int t = lane_id(); t < s; t += 32) dst[t] = valid ? src[t] : 0.0;
}

// Per query, after the round's rows of L are complete: the new blocks of z (forward substitution through the new rows, one
// sub-block of <= 32 after the other), then y = L^-T z by blocks in reverse (the block inverses make every step a transposed
// matrix-vector product), then e = (M y)^T.  Per-query rows come in rounds of n, each cut into sub-blocks of <= 32 columns
// whose 32 x 32 inverses sit at Linvq + (local row) * 32; the shared block (s columns) has its s x s inverse L1inv.
__global__ void __launch_bounds__(256, 4)
backsolve_e_kernel(const float* __restrict__ Rt, int k_q, int64_t n_items, const int64_t* __restrict__ shared_anc, const double* __restrict__ L1inv, int s,
                   const int64_t* __restrict__ anc_q, int64_t bs_anc, const double* __restrict__ Lq, int64_t bsLq, int ldl,
                   const double* __restrict__ Linvq, int64_t bsLinv, double* __restrict__ z, int64_t bsz, int n, int rounds_done,
                   int new_round, const float* __restrict__ c_new, float* __restrict__ e_out) {
    // rounds_done: per-query rounds whose z is already in the state; new_round = 1: one more round (rows rounds_done * n ..)
    // is forward-substituted here from c_new.  The kernel is a chain of dependent steps per query, so every step keeps
    // many independent loads in flight (its time is L2 round trips, not bytes).
    extern __shared__ __align__(16) double bs_smem[];
    const int rounds = rounds_done + new_round;
    const int m = s + rounds * n;
    double* zs = bs_smem;                                   // [m] z
    double* ys = zs + m;                                    // [m] right-hand side, becomes y block by block
    double* tv = ys + m;                                    // [32]
    double* Ivs = tv + CB;                                  // [32 x 32] the sub-block's inverse
    int64_t* items = reinterpret_cast<int64_t*>(Ivs + CB * CB);   // [m]
    const int b = blockIdx.x, tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int m_old = s + rounds_done * n;
    for (int i = tid; i < m; i += 256) {
        zs[i] = i < m_old ? z[int64_t(b) * bsz + i] : 0.0;
        const int64_t it = i < s ? shared_anc[i] : anc_q[int64_t(b) * bs_anc + (i - s)];
        items[i] = (it >= 0 && it < n_items) ? it : -1;                  // invalid anchors were zero columns: y_i = 0, row skipped
    }
    __syncthreads();
    const double* L = Lq + int64_t(b) * bsLq;
    const double* Iq = Linvq + int64_t(b) * bsLinv;
    const int nsub = (n + CB - 1) / CB;
    if (new_round) {
        for (int k = 0; k < nsub; ++k) {
            const int rho = rounds_done * n + k * CB, w = n - k * CB < CB ? n - k * CB : CB, c0 = s + rho;
            double iv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) iv[u] = (tid + 256 * u) < w * CB ? Iq[int64_t(rho) * CB + tid + 256 * u] : 0.0;
            // t_i = c_i - sum_{col < c0} L[row i][col] z[col]: a warp takes rows ty, ty + 8, ty + 16, ty + 24 together
            double p[4] = {0.0, 0.0, 0.0, 0.0};
            const double* Lr[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) Lr[u] = L + int64_t(rho + (ty + 8 * u < w ? ty + 8 * u : 0)) * ldl;
            double p2[4] = {0.0, 0.0, 0.0, 0.0};
            int col = tx;
            for (; col + 32 < c0; col += 64) {              // eight loads in flight per lane
                const double zc = zs[col], zd = zs[col + 32];
#pragma unroll
                for (int u = 0; u < 4; ++u) { p[u] = fma(Lr[u][col], zc, p[u]); p2[u] = fma(Lr[u][col + 32], zd, p2[u]); }
            }
            if (col < c0) {
                const double zc = zs[col];
#pragma unroll
                for (int u = 0; u < 4; ++u) p[u] = fma(Lr[u][col], zc, p[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) p[u] += p2[u];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) p[u] += __shfl_xor_sync(0xffffffffu, p[u], o);
                const int i = ty + 8 * u;
                if (tx == 0 && i < w) tv[i] = double(c_new[int64_t(b) * n + k * CB + i]) - p[u];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) Ivs[tid + 256 * u] = iv[u];
            __syncthreads();
            if (tid < w) {
                double acc = 0.0;
                for (int j = 0; j <= tid; ++j) acc = fma(Ivs[tid * CB + j], tv[j], acc);
                zs[c0 + tid] = acc;
                z[int64_t(b) * bsz + c0 + tid] = acc;
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < m; i += 256) ys[i] = zs[i];
    __syncthreads();
    for (int rr = rounds - 1; rr >= 0; --rr) {
        for (int k = nsub - 1; k >= 0; --k) {
            const int rho = rr * n + k * CB, w = n - k * CB < CB ? n - k * CB : CB, c0 = s + rho;
#pragma unroll
            for (int u = 0; u < 4; ++u) Ivs[tid + 256 * u] = (tid + 256 * u) < w * CB ? Iq[int64_t(rho) * CB + tid + 256 * u] : 0.0;
            __syncthreads();
            double yj = 0.0;
            if (tid < w)                                    // y_j = sum_{i >= j} Linv[i][j] rhs[i]
                for (int i = tid; i < w; ++i) yj = fma(Ivs[i * CB + tid], ys[c0 + i], yj);
            __syncthreads();
            if (tid < w) ys[c0 + tid] = yj;
            __syncthreads();
            // rhs[col] -= sum_i L[row i of the block][col] y_i for col < c0: columns tid and tid + 256, eight rows at a time
            const double* Lr = L + int64_t(rho) * ldl;
            const int colA = tid, colB = tid + 256;
            const bool okA = colA < c0, okB = colB < c0;
            double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
            for (int i = 0; i < w; i += 8) {
                double va[8], vb[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const bool in = i + u < w;
                    va[u] = (okA && in) ? Lr[int64_t(i + u) * ldl + colA] : 0.0;
                    vb[u] = (okB && in) ? Lr[int64_t(i + u) * ldl + colB] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 8; u += 2) {
                    const double y0 = i + u < w ? ys[c0 + i + u] : 0.0, y1 = i + u + 1 < w ? ys[c0 + i + u + 1] : 0.0;
                    a0 = fma(va[u], y0, a0); a1 = fma(va[u + 1], y1, a1);
                    b0 = fma(vb[u], y0, b0); b1 = fma(vb[u + 1], y1, b1);
                }
            }
            if (okA) ys[colA] -= a0 + a1;
            if (okB) ys[colB] -= b0 + b1;
            for (int col = tid + 512; col < c0; col += 256) {     // m > 512 anchors: the plain loop
                double acc = 0.0;
                for (int i = 0; i < w; ++i) acc = fma(Lr[int64_t(i) * ldl + col], ys[c0 + i], acc);
                ys[col] -= acc;
            }
            __syncthreads();
        }
    }
    if (s > 0) {                                            // the shared block: y_0 = L_1^-T rhs_0
        double yj = 0.0;
        if (tid < s) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int i = tid;
            for (; i + 15 < s; i += 16) {                   // sixteen loads in flight
                double v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) v[u] = __ldg(L1inv + int64_t(i + u) * s + tid);
#pragma unroll
                for (int u = 0; u < 16; u += 4) {
                    a0 = fma(v[u], ys[i + u], a0); a1 = fma(v[u + 1], ys[i + u + 1], a1);
                    a2 = fma(v[u + 2], ys[i + u + 2], a2); a3 = fma(v[u + 3], ys[i + u + 3], a3);
                }
            }
            {                                               // tail: up to 15 rows, predicated, all loads issued together
                double v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) v[u] = i + u < s ? __ldg(L1inv + int64_t(i + u) * s + tid) : 0.0;
#pragma unroll
                for (int u = 0; u < 16; u += 4) {
                    a0 = fma(v[u], i + u < s ? ys[i + u] : 0.0, a0); a1 = fma(v[u + 1], i + u + 1 < s ? ys[i + u + 1] : 0.0, a1);
                    a2 = fma(v[u + 2], i + u + 2 < s ? ys[i + u + 2] : 0.0, a2); a3 = fma(v[u + 3], i + u + 3 < s ? ys[i + u + 3] : 0.0, a3);
                }
            }
            yj = (a0 + a1) + (a2 + a3);
        }
        __syncthreads();
        if (tid < s) ys[tid] = yj;
        __syncthreads();
    }
    // e = y^T M.  k_q even and 8-byte aligned rows: one 64-bit load per row brings columns 2t, 2t + 1, sixteen rows in flight
    // (the loop is a chain of L2 / DRAM round trips: 24 instead of 47 at m = 375); otherwise columns t and t + 256, eight rows.
    if ((k_q & 1) == 0 && (reinterpret_cast<uintptr_t>(Rt) & 7) == 0) {
        for (int t = tid; 2 * t < k_q; t += 256) {
            double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
            for (int i = 0; i < m; i += 16) {
                float2 v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const bool in = i + u < m && items[i + u] >= 0;
                    const float2* row = reinterpret_cast<const float2*>(Rt + (in ? items[i + u] : 0) * k_q);
                    v[u] = in ? __ldg(row + t) : make_float2(0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 16; u += 2) {
                    const double y0 = i + u < m ? ys[i + u] : 0.0, y1 = i + u + 1 < m ? ys[i + u + 1] : 0.0;
                    a0 = fma(y0, double(v[u].x), a0); a1 = fma(y1, double(v[u + 1].x), a1);
                    b0 = fma(y0, double(v[u].y), b0); b1 = fma(y1, double(v[u + 1].y), b1);
                }
            }
            *reinterpret_cast<float2*>(e_out + int64_t(b) * k_q + 2 * t) = make_float2(float(a0 + a1), float(b0 + b1));
        }
        return;
    }
    for (int t = tid; t < k_q; t += 512) {
        const bool ok2 = t + 256 < k_q;
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
        for (int i = 0; i < m; i += 8) {
            float va[8], vb[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool in = i + u < m && items[i + u] >= 0;
                const float* row = Rt + (in ? items[i + u] : 0) * k_q;
                va[u] = in ? __ldg(row + t) : 0.f;
                vb[u] = (in && ok2) ? __ldg(row + t + 256) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                const double y0 = i + u < m ? ys[i + u] : 0.0, y1 = i + u + 1 < m ? ys[i + u + 1] : 0.0;
                a0 = fma(y0, double(va[u]), a0); a1 = fma(y1, double(va[u + 1]), a1);
                b0 = fma(y0, double(vb[u]), b0); b1 = fma(y1, double(vb[u + 1]), b1);
            }
        }
        e_out[int64_t(b) * k_q + t] = float(a0 + a1);
        if (ok2) e_out[int64_t(b) * k_q + t + 256] = float(b0 + b1);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// layouts
// ---------------------------------------------------------------------------------------------------------------------------
struct IncShared { size_t off_anc, off_maxd, off_l1, off_l1inv, off_w1t, total; };
static IncShared inc_shared_layout(int64_t n_items, int s) {
    IncShared p{};
    size_t off = 0;
    p.off_anc = off; off += align_up(sizeof(int64_t) * size_t(s > 0 ? s : 1), 256);
    p.off_maxd = off; off += 256;
    p.off_l1 = off; off += align_up(sizeof(double) * size_t(s) * s, 256);
    p.off_l1inv = off; off += align_up(sizeof(double) * size_t(s) * s, 256);
    p.off_w1t = off; off += align_up(sizeof(double) * size_t(n_items) * s, 256);
    p.total = off > 0 ? off : 256;
    return p;
}

struct IncState { int mq, ldl; size_t off_anc, off_maxd, off_z, off_lq, off_linv, off_t, total; int64_t bs_anc, bsz, bsLq, bsLinv, bsT; };
static IncState inc_state_layout(int n_queries, int s, int n, int m_max) {
    IncState p{};
    p.mq = m_max - s;
    p.ldl = (m_max + 1) & ~1;
    const size_t B = size_t(n_queries > 0 ? n_queries : 1);
    size_t off = 0;
    p.bs_anc = p.mq; p.bsz = p.ldl; p.bsLq = int64_t(p.mq) * p.ldl; p.bsLinv = int64_t(p.mq) * CB; p.bsT = int64_t(n) * p.ldl;
    p.off_anc = off; off += align_up(sizeof(int64_t) * B * size_t(p.mq > 0 ? p.mq : 1), 256);
    p.off_maxd = off; off += align_up(sizeof(double) * B, 256);
    p.off_z = off; off += align_up(sizeof(double) * B * p.ldl, 256);
    p.off_lq = off; off += align_up(sizeof(double) * B * size_t(p.bsLq > 0 ? p.bsLq : 1), 256);
    p.off_linv = off; off += align_up(sizeof(double) * B * size_t(p.bsLinv > 0 ? p.bsLinv : 1), 256);
    p.off_t = off; off += align_up(sizeof(double) * B * size_t(p.bsT), 256);
    p.total = off;
    return p;
}

static size_t chol_inv_smem(int n) { return sizeof(double) * (size_t(n) * (n + 1) + 2 * size_t(n)); }

size_t adaptive_shared_bytes(int k_q, int64_t n_items, int m_shared) {
    (void)k_q;
    if (m_shared < 0 || n_items <= 0) return 256;
    return inc_shared_layout(n_items, m_shared).total;
}
size_t adaptive_prepare_workspace_bytes(int k_q, int64_t n_items, int m_shared) {
    (void)k_q;
    if (m_shared <= 0 || n_items <= 0) return 256;
    return align_up(sizeof(double) * size_t(n_items) * m_shared, 256);
}
size_t adaptive_state_bytes(int n_queries, int k_q, int m_shared, int n_new, int m_max) {
    (void)k_q;
    if (n_new <= 0 || m_max < m_shared || m_shared < 0) return 256;
    return inc_state_layout(n_queries, m_shared, n_new, m_max).total;
}

static int check_inc_shape(const char* who, int k_q, int s, int n, int m_max) {
    if (s < 0 || s > ANNCUR_ADAPTIVE_MAX_BLOCK || n < 1 || n > ANNCUR_ADAPTIVE_MAX_BLOCK) {
        set_error("%s: m_shared = %d / n_new = %d outside [0 | 1, %d]", who, s, n, ANNCUR_ADAPTIVE_MAX_BLOCK);
        return ANNCUR_E_UNSUPPORTED;
    }
    if (m_max < s || (m_max - s) % n != 0) { set_error("%s: m_max = %d is not m_shared + a multiple of n_new", who, m_max); return ANNCUR_E_INVALID; }
    if (m_max > k_q) { set_error("%s: m_max = %d anchors > k_q = %d anchor queries is not supported", who, m_max, k_q); return ANNCUR_E_UNSUPPORTED; }
    if (m_max > 2048) { set_error("%s: m_max = %d > 2048", who, m_max); return ANNCUR_E_UNSUPPORTED; }
    return ANNCUR_OK;
}

// Once per (index, first anchors): L_1, L_1^-1 and W_1^T = (L_1^-1 M_1^T R_anc)^T (n_items x m_shared fp64).
int adaptive_prepare(const float* Rt, int k_q, int64_t n_items, const int64_t* shared_anchors, int s, double rcond,
                     void* shared, size_t shared_bytes, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    int rc = check_inc_shape("adaptive_prepare", k_q, s, 1, s);
    if (rc != ANNCUR_OK) return rc;
    const IncShared sl = inc_shared_layout(n_items, s);
    if (shared_bytes < sl.total) { set_error("adaptive_prepare: shared blob too small: %zu < %zu", shared_bytes, sl.total); return ANNCUR_E_WORKSPACE; }
    if (s == 0) return ANNCUR_OK;
    if (workspace_bytes < adaptive_prepare_workspace_bytes(k_q, n_items, s)) { set_error("adaptive_prepare: workspace too small"); return ANNCUR_E_WORKSPACE; }
    if (n_items >= (int64_t(1) << 31) - DG_T) { set_error("adaptive_prepare: n_items = %lld too large", (long long)n_items); return ANNCUR_E_UNSUPPORTED; }
    char* sb = reinterpret_cast<char*>(shared);
    int64_t* anc = reinterpret_cast<int64_t*>(sb + sl.off_anc);
    double* maxd = reinterpret_cast<double*>(sb + sl.off_maxd);
    double* L1 = reinterpret_cast<double*>(sb + sl.off_l1);
    double* L1inv = reinterpret_cast<double*>(sb + sl.off_l1inv);
    double* W1t = reinterpret_cast<double*>(sb + sl.off_w1t);
    double* V = reinterpret_cast<double*>(workspace);
    ANNCUR_CUDA_OK(cudaMemcpyAsync(anc, shared_anchors, sizeof(int64_t) * s, cudaMemcpyDeviceToDevice, stream));
    ANNCUR_CUDA_OK(cudaMemsetAsync(maxd, 0, sizeof(double), stream));
    const int vec4 = (k_q % 4 == 0) && (reinterpret_cast<uintptr_t>(Rt) % 16 == 0);
    DgArgs g{};
    g.Rt = Rt; g.ld_rt = k_q; g.n_rt = n_items; g.idxA = anc; g.bs_idxA = 0; g.idxB = anc; g.bs_idxB = 0;
    g.Cout = L1; g.ldc = s; g.bsC = 0; g.M = s; g.N = s; g.K = k_q; g.sign = 1.0; g.diag_off = 0; g.vec4 = vec4;
    rc = dgemm_launch(g, true, 1, stream);
    if (rc != ANNCUR_OK) return rc;
    // V (workspace) first holds a copy of the raw Gram diagonal block (the Cholesky runs in place on L1)
    ANNCUR_CUDA_OK(cudaMemcpyAsync(V, L1, sizeof(double) * size_t(s) * s, cudaMemcpyDeviceToDevice, stream));
    const size_t smem = chol_inv_smem(s);
    ANNCUR_CUDA_OK(cudaFuncSetAttribute(chol_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    chol_inv_kernel<<<1, 256, smem, stream>>>(L1, 0, s, V, 0, s, maxd, L1inv, 0, s, rcond);
    ANNCUR_LAUNCH_OK("chol_inv_kernel");
    // V = R_anc^T M_1 (n_items x s), W_1^T = V L_1^-T
    DgArgs v{};
    v.Rt = Rt; v.ld_rt = k_q; v.n_rt = n_items; v.idxA = nullptr; v.idxB = anc; v.bs_idxB = 0;
    v.Cout = V; v.ldc = s; v.bsC = 0; v.M = int(n_items); v.N = s; v.K = k_q; v.sign = 1.0; v.diag_off = DG_ALL; v.vec4 = vec4;
    rc = dgemm_launch(v, true, 1, stream);
    if (rc != ANNCUR_OK) return rc;
    DgArgs w{};
    w.A = V; w.lda = s; w.bsA = 0; w.B = L1inv; w.ldb = s; w.bsB = 0;
    w.Cout = W1t; w.ldc = s; w.bsC = 0; w.M = int(n_items); w.N = s; w.K = s; w.sign = 1.0; w.diag_off = DG_ALL;
    return dgemm_launch(w, false, 1, stream);
}

static int backsolve_launch(const float* Rt, int k_q, int64_t n_items, const IncShared& sl, const char* sb, int s, const IncState& st, char* stb,
                            int n, int rounds_done, int new_round, const float* c_new, int n_queries, float* e_out, cudaStream_t stream) {
    const int m = s + (rounds_done + new_round) * n;
    const size_t smem = sizeof(double) * (2 * size_t(m) + CB + CB * CB) + sizeof(int64_t) * size_t(m > 0 ? m : 1);
    ANNCUR_CUDA_OK(cudaFuncSetAttribute(backsolve_e_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    backsolve_e_kernel<<<n_queries, 256, smem, stream>>>(
        Rt, k_q, n_items, reinterpret_cast<const int64_t*>(sb + sl.off_anc), reinterpret_cast<const double*>(sb + sl.off_l1inv), s,
        reinterpret_cast<const int64_t*>(stb + st.off_anc), st.bs_anc, reinterpret_cast<const double*>(stb + st.off_lq), st.bsLq, st.ldl,
        reinterpret_cast<const double*>(stb + st.off_linv), st.bsLinv, reinterpret_cast<double*>(stb + st.off_z), st.bsz, n, rounds_done,
        new_round, c_new, e_out);
    ANNCUR_LAUNCH_OK("backsolve_e_kernel");
    return ANNCUR_OK;
}

// Round 1: the shared anchors.  z_0 = L_1^-1 c_1 per query, e = c_1 . pinv(M_1).
int adaptive_begin(const float* Rt, int k_q, int64_t n_items, const void* shared, int s, const float* c, int n_queries, int n_new,
                   int m_max, float* e_out, void* state, size_t state_bytes, cudaStream_t stream) {
    int rc = check_inc_shape("adaptive_begin", k_q, s, n_new, m_max);
    if (rc != ANNCUR_OK) return rc;
    const IncShared sl = inc_shared_layout(n_items, s);
    const IncState st = inc_state_layout(n_queries, s, n_new, m_max);
    if (state_bytes < st.total) { set_error("adaptive_begin: state too small: %zu < %zu", state_bytes, st.total); return ANNCUR_E_WORKSPACE; }
    if (n_queries == 0) return ANNCUR_OK;
    const char* sb = reinterpret_cast<const char*>(shared);
    char* stb = reinterpret_cast<char*>(state);
    z_shared_kernel<<<n_queries, 128, sizeof(double) * size_t(s > 0 ? s : 1), stream>>>(
        reinterpret_cast<const double*>(sb + sl.off_l1inv), s, reinterpret_cast<const double*>(sb + sl.off_maxd), c,
        reinterpret_cast<double*>(stb + st.off_z), st.bsz, reinterpret_cast<double*>(stb + st.off_maxd));
    ANNCUR_LAUNCH_OK("z_shared_kernel");
    if (s == 0) { ANNCUR_CUDA_OK(cudaMemsetAsync(e_out, 0, sizeof(float) * size_t(n_queries) * k_q, stream)); return ANNCUR_OK; }
    return backsolve_launch(Rt, k_q, n_items, sl, sb, s, st, stb, n_new, 0, 0, nullptr, n_queries, e_out, stream);
}

// Round t + 1: every query gets n_new more anchors (m_cur = anchors so far = m_shared + r * n_new).
int adaptive_extend(const float* Rt, int k_q, int64_t n_items, const void* shared, int s, const int64_t* new_anchors,
                    const float* c_new, int n_queries, int n, int m_max, int m_cur, double rcond, float* e_out, void* state,
                    size_t state_bytes, cudaStream_t stream) {
    int rc = check_inc_shape("adaptive_extend", k_q, s, n, m_max);
    if (rc != ANNCUR_OK) return rc;
    if (m_cur < s || (m_cur - s) % n != 0 || m_cur + n > m_max) { set_error("adaptive_extend: m_cur = %d does not fit m_shared = %d, n_new = %d, m_max = %d", m_cur, s, n, m_max); return ANNCUR_E_INVALID; }
    const IncShared sl = inc_shared_layout(n_items, s);
    const IncState st = inc_state_layout(n_queries, s, n, m_max);
    if (state_bytes < st.total) { set_error("adaptive_extend: state too small: %zu < %zu", state_bytes, st.total); return ANNCUR_E_WORKSPACE; }
    if (n_queries == 0) return ANNCUR_OK;
    if (n_queries > 65535) { set_error("adaptive_extend: n_queries = %d > 65535 per call", n_queries); return ANNCUR_E_UNSUPPORTED; }
    const char* sb = reinterpret_cast<const char*>(shared);
    char* stb = reinterpret_cast<char*>(state);
    int64_t* anc_q = reinterpret_cast<int64_t*>(stb + st.off_anc);
    double* maxd = reinterpret_cast<double*>(stb + st.off_maxd);
    double* z = reinterpret_cast<double*>(stb + st.off_z);
    double* Lq = reinterpret_cast<double*>(stb + st.off_lq);
    double* Linvq = reinterpret_cast<double*>(stb + st.off_linv);
    double* T = reinterpret_cast<double*>(stb + st.off_t);
    const double* W1t = reinterpret_cast<const double*>(sb + sl.off_w1t);
    const int r = (m_cur - s) / n;                         // per-query blocks so far
    const int64_t new_row0 = int64_t(r) * n * st.ldl;      // offset of the new rows inside a query's Lq
    const int vec4 = (k_q % 4 == 0) && (reinterpret_cast<uintptr_t>(Rt) % 16 == 0);

    // a. anchors join the lists; shared columns of the new rows = gather from W_1^T
    {
        const int64_t warps = int64_t(n_queries) * n;
        append_anchors_kernel<<<unsigned((warps * 32 + 255) / 256), 256, 0, stream>>>(new_anchors, n, n_queries, r, n_items, anc_q, st.bs_anc, W1t, s,
                                                                                    Lq, st.bsLq, st.ldl);
        ANNCUR_LAUNCH_OK("append_anchors_kernel");
    }
    // b. raw Gram rows of the new anchors against the per-query anchors (old, then new: lower triangle of the last block)
    {
        DgArgs g{};
        g.Rt = Rt; g.ld_rt = k_q; g.n_rt = n_items; g.idxA = anc_q + int64_t(r) * n; g.bs_idxA = st.bs_anc; g.idxB = anc_q; g.bs_idxB = st.bs_anc;
        g.Cout = T + s; g.ldc = st.ldl; g.bsC = st.bsT; g.M = n; g.N = (r + 1) * n; g.K = k_q; g.sign = 1.0; g.diag_off = r * n; g.vec4 = vec4;
        rc = dgemm_launch(g, true, n_queries, stream);
        if (rc != ANNCUR_OK) return rc;
    }
    // c. the columns of the earlier per-query rounds.  Per round block: everything to its left in ONE GEMM with full tiles
    //    (T_blk -= X[:, 0:cb] L_blk[:, 0:cb]^T), then sub-block by sub-block the part inside the block (K = 32 k) and
    //    X_k = T_k L_kk^-T.
    const int nsub = (n + CB - 1) / CB;
    for (int rr = 0; rr < r; ++rr) {
        const int cb = s + rr * n;
        if (cb > 0) {
            DgArgs u{};
            u.A = Lq + new_row0; u.lda = st.ldl; u.bsA = st.bsLq; u.B = Lq + int64_t(rr) * n * st.ldl; u.ldb = st.ldl; u.bsB = st.bsLq;
            u.Cin = T + cb; u.ldcin = st.ldl; u.bsCin = st.bsT; u.Cout = T + cb; u.ldc = st.ldl; u.bsC = st.bsT;
            u.M = n; u.N = n; u.K = cb; u.sign = -1.0; u.diag_off = DG_ALL;
            rc = dgemm_launch(u, false, n_queries, stream);
            if (rc != ANNCUR_OK) return rc;
        }
        for (int k = 0; k < nsub; ++k) {
            const int rho = rr * n + k * CB, w = n - k * CB < CB ? n - k * CB : CB, c0 = s + rho;
            if (k > 0) {
                DgArgs u{};
                u.A = Lq + new_row0 + cb; u.lda = st.ldl; u.bsA = st.bsLq; u.B = Lq + int64_t(rho) * st.ldl + cb; u.ldb = st.ldl; u.bsB = st.bsLq;
                u.Cin = T + c0; u.ldcin = st.ldl; u.bsCin = st.bsT; u.Cout = T + c0; u.ldc = st.ldl; u.bsC = st.bsT;
                u.M = n; u.N = w; u.K = k * CB; u.sign = -1.0; u.diag_off = DG_ALL;
                rc = dgemm_launch(u, false, n_queries, stream);
                if (rc != ANNCUR_OK) return rc;
            }
            DgArgs x{};
            x.A = T + c0; x.lda = st.ldl; x.bsA = st.bsT; x.B = Linvq + int64_t(rho) * CB; x.ldb = CB; x.bsB = st.bsLinv;
            x.Cout = Lq + new_row0 + c0; x.ldc = st.ldl; x.bsC = st.bsLq; x.M = n; x.N = w; x.K = w; x.sign = 1.0; x.diag_off = DG_ALL;
            rc = dgemm_launch(x, false, n_queries, stream);
            if (rc != ANNCUR_OK) return rc;
        }
    }
    // d. Schur complement of everything before this round: S = N^T N - X X^T (lower) into the diagonal block
    {
        DgArgs d{};
        d.A = Lq + new_row0; d.lda = st.ldl; d.bsA = st.bsLq; d.B = Lq + new_row0; d.ldb = st.ldl; d.bsB = st.bsLq;
        d.Cin = T + m_cur; d.ldcin = st.ldl; d.bsCin = st.bsT; d.Cout = Lq + new_row0 + m_cur; d.ldc = st.ldl; d.bsC = st.bsLq;
        d.M = n; d.N = n; d.K = m_cur; d.sign = -1.0; d.diag_off = 0;
        rc = dgemm_launch(d, false, n_queries, stream);
        if (rc != ANNCUR_OK) return rc;
    }
    // e. right-looking factorisation of the n x n block in sub-blocks of <= 32: warp-level Cholesky + inverse of the diagonal
    //    sub-block, the rows below it times the inverse (in place: one column tile, every CTA reads only the rows it writes),
    //    rank-w update of what is left
    for (int k = 0; k < nsub; ++k) {
        const int rho = r * n + k * CB, w = n - k * CB < CB ? n - k * CB : CB, ck = m_cur + k * CB;
        double* blk = Lq + int64_t(rho) * st.ldl + ck;
        chol_inv32_kernel<<<(n_queries + 3) / 4, 128, 0, stream>>>(blk, st.bsLq, st.ldl, T + int64_t(k) * CB * st.ldl + ck, st.bsT, st.ldl, maxd,
                                                                 Linvq + int64_t(rho) * CB, st.bsLinv, w, rcond, n_queries);
        ANNCUR_LAUNCH_OK("chol_inv32_kernel");
        const int mb = n - k * CB - w;
        if (mb <= 0) continue;
        double* below = Lq + int64_t(rho + w) * st.ldl + ck;
        DgArgs x{};
        x.A = below; x.lda = st.ldl; x.bsA = st.bsLq; x.B = Linvq + int64_t(rho) * CB; x.ldb = CB; x.bsB = st.bsLinv;
        x.Cout = below; x.ldc = st.ldl; x.bsC = st.bsLq; x.M = mb; x.N = w; x.K = w; x.sign = 1.0; x.diag_off = DG_ALL;
        rc = dgemm_launch(x, false, n_queries, stream);
        if (rc != ANNCUR_OK) return rc;
        DgArgs t{};
        t.A = below; t.lda = st.ldl; t.bsA = st.bsLq; t.B = below; t.ldb = st.ldl; t.bsB = st.bsLq;
        t.Cin = below + w; t.ldcin = st.ldl; t.bsCin = st.bsLq; t.Cout = below + w; t.ldc = st.ldl; t.bsC = st.bsLq;
        t.M = mb; t.N = mb; t.K = w; t.sign = -1.0; t.diag_off = 0;
        rc = dgemm_launch(t, false, n_queries, stream);
        if (rc != ANNCUR_OK) return rc;
    }
    // f. new blocks of z, y = L^-T z, e = (M y)^T
    return backsolve_launch(Rt, k_q, n_items, sl, sb, s, st, stb, n, r, 1, c_new, n_queries, e_out, stream);
}

}  // namespace anncur
