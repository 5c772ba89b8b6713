// K1: Moore-Penrose pseudo-inverse of an fp32 matrix, replacing np.linalg.pinv at
// eval/matrix_approx_zeshel.py:47,49 (LAPACK SVD, singular values <= rcond*s_max dropped).
//
// One-sided (Hestenes) Jacobi SVD in fp64 on the tall orientation T (len x p, p = min(m, n)):
// columns are rotated pairwise until mutually orthogonal, T.V = G, sigma_j = |g_j|, and
//     pinv(T) = V . diag(1/sigma_j^2 [sigma_j > rcond*sigma_max]) . G^T            (p x len)
// Columns of G and V are stored contiguously (k-major), one CTA per column pair per round of a
// round-robin tournament; everything stays on the device and on the caller's stream (a converged
// flag turns the remaining round launches into no-ops instead of synchronising with the host).
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace anncur {

constexpr int JAC_THREADS = 256;
constexpr int JAC_MAX_SWEEPS = 30;

struct JacobiState {
    unsigned long long max_off_bits;   // max |g_i.g_j| / (|g_i||g_j|) seen in this sweep (double bits)
    int converged;
    int sweeps_done;
    double s_max, s_min_kept;
    // Cholesky fast path of pinv_f32 (see below): 1 = the normal-equations route produced the result and every kernel of
    // the Jacobi route returns at once; 0 = not tried / refused (pivot breakdown or ill-conditioned) -> Jacobi runs
    int fast_ok;
    int chol_fail;
    double diag_min, diag_max;         // extreme diagonal entries of the Cholesky factor: diag_max / diag_min <= cond(T)
};

__global__ void jacobi_init_kernel(const float* __restrict__ A, int m, int n, int lda, int len, int p,
                                   bool tall, double* __restrict__ G, double* __restrict__ V, JacobiState* st) {
    if (st->fast_ok) return;
    // G[j][r] = T[r][j];  tall: T = A (len = m, p = n);  wide: T = A^T (len = n, p = m)
    const int64_t total = int64_t(len) * p;
    for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
        int j = int(t / len), r = int(t % len);
        G[t] = tall ? double(A[int64_t(r) * lda + j]) : double(A[int64_t(j) * lda + r]);
    }
    const int64_t vt = int64_t(p) * p;
    for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < vt; t += int64_t(gridDim.x) * blockDim.x)
        V[t] = (t / p == t % p) ? 1.0 : 0.0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->max_off_bits = 0ull; st->converged = 0; st->sweeps_done = 0; st->s_max = 0.0; st->s_min_kept = 0.0;
    }
}

__global__ void jacobi_state_reset_kernel(JacobiState* st) {
    st->max_off_bits = 0ull; st->converged = 0; st->sweeps_done = 0; st->s_max = 0.0; st->s_min_kept = 0.0;
    st->fast_ok = 0; st->chol_fail = 0; st->diag_min = 0.0; st->diag_max = 0.0;
}

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < JAC_THREADS / 32; ++w) s += red[w];
    return s;
}

// one column pair of one tournament round (the whole CTA works on it)
__device__ __forceinline__ void jacobi_pair(double* __restrict__ G, double* __restrict__ V, int len, int p, int p_even,
                                            int round, double tol, JacobiState* st) {
    __shared__ double red[JAC_THREADS / 32];
    __shared__ double cs[2];
    // round-robin tournament (circle method): player p_even-1 is fixed, the rest rotate
    int i, j;
    const int q = p_even - 1;
    if (blockIdx.x == 0) { i = q; j = round % q; }
    else { i = (round + int(blockIdx.x)) % q; j = (round - int(blockIdx.x) + q) % q; }
    if (i >= p || j >= p) return;                      // bye (odd p)
    if (i > j) { int t = i; i = j; j = t; }
    double* gi = G + int64_t(i) * len;
    double* gj = G + int64_t(j) * len;
    double a = 0.0, b = 0.0, g = 0.0;
    for (int r = threadIdx.x; r < len; r += JAC_THREADS) {
        double x = gi[r], y = gj[r];
        a = fma(x, x, a); b = fma(y, y, b); g = fma(x, y, g);
    }
    // one block reduction for the three sums
    __shared__ double red3[3][JAC_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        g += __shfl_xor_sync(0xffffffffu, g, o);
    }
    if ((threadIdx.x & 31) == 0) { red3[0][threadIdx.x >> 5] = a; red3[1][threadIdx.x >> 5] = b; red3[2][threadIdx.x >> 5] = g; }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = b = g = 0.0;
        for (int w = 0; w < JAC_THREADS / 32; ++w) { a += red3[0][w]; b += red3[1][w]; g += red3[2][w]; }
        double c = 1.0, s = 0.0;
        if (a > 0.0 && b > 0.0) {
            double off = fabs(g) / sqrt(a * b);
            atomicMax(&st->max_off_bits, (unsigned long long)__double_as_longlong(off));
            if (off > tol) {
                double zeta = (b - a) / (2.0 * g);
                double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                c = 1.0 / sqrt(1.0 + t * t);
                s = c * t;
            }
        }
        cs[0] = c; cs[1] = s;
    }
    __syncthreads();
    const double c = cs[0], s = cs[1];
    if (s == 0.0) return;
    for (int r = threadIdx.x; r < len; r += JAC_THREADS) {
        double x = gi[r], y = gj[r];
        gi[r] = c * x - s * y;
        gj[r] = s * x + c * y;
    }
    double* vi = V + int64_t(i) * p;
    double* vj = V + int64_t(j) * p;
    for (int r = threadIdx.x; r < p; r += JAC_THREADS) {
        double x = vi[r], y = vj[r];
        vi[r] = c * x - s * y;
        vj[r] = s * x + c * y;
    }
}

__global__ void __launch_bounds__(JAC_THREADS)
jacobi_round_kernel(double* __restrict__ G, double* __restrict__ V, int len, int p, int p_even, int round,
                    double tol, JacobiState* st) {
    if (st->converged || st->fast_ok) return;
    jacobi_pair(G, V, len, p, p_even, round, tol, st);
}

__device__ __forceinline__ void jacobi_sweep_end(JacobiState* st, double tol) {
    double off = __longlong_as_double((long long)st->max_off_bits);
    st->sweeps_done += 1;
    if (off <= tol) st->converged = 1;
    st->max_off_bits = 0ull;
}

__global__ void jacobi_sweep_end_kernel(JacobiState* st, double tol) {
    if (st->converged || st->fast_ok) return;
    jacobi_sweep_end(st, tol);
}

// All sweeps in ONE cooperative launch: one CTA per column pair, a grid-wide barrier between the rounds of the
// tournament instead of a kernel boundary (the factorisation is launch-latency-bound: ~10 sweeps x (p - 1) rounds of
// microsecond-sized work).  Used when p_even / 2 CTAs are co-resident; the multi-launch path above is the fallback.
__global__ void __launch_bounds__(JAC_THREADS)
jacobi_coop_kernel(double* __restrict__ G, double* __restrict__ V, int len, int p, int p_even, double tol, JacobiState* st) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    if (st->fast_ok) return;                               // uniform over the grid: nobody reaches a grid.sync
    for (int sweep = 0; sweep < JAC_MAX_SWEEPS; ++sweep) {
        for (int round = 0; round < p_even - 1; ++round) {
            jacobi_pair(G, V, len, p, p_even, round, tol, st);
            grid.sync();
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) jacobi_sweep_end(st, tol);
        grid.sync();
        if (*reinterpret_cast<volatile int*>(&st->converged)) break;
    }
}

// sigma_j^2 -> weights 1/sigma_j^2 with numpy's cutoff; one CTA, p threads strided
__global__ void __launch_bounds__(JAC_THREADS)
jacobi_weights_kernel(const double* __restrict__ G, int len, int p, double rcond, double tol, double* __restrict__ w,
                      JacobiState* st) {
    __shared__ double red[JAC_THREADS / 32];
    __shared__ double smax;
    if (st->fast_ok) return;
    // pass 1: squared norms (one column per loop trip, whole CTA reduces)
    double local_max = 0.0;
    for (int j = 0; j < p; ++j) {
        double a = 0.0;
        const double* g = G + int64_t(j) * len;
        for (int r = threadIdx.x; r < len; r += JAC_THREADS) a = fma(g[r], g[r], a);
        a = block_sum(a, red);
        if (threadIdx.x == 0) w[j] = a;
        local_max = fmax(local_max, a);
    }
    if (threadIdx.x == 0) smax = sqrt(local_max);
    __syncthreads();
    // numpy's cutoff is rcond * s_max on singular values that LAPACK computed in fp32.  Ours are fp64 Jacobi values: a
    // column that is an exact combination of others comes out at ~1e-16 s_max (not ~1e-7 as in fp32) and would be inverted
    // to 1e32 under rcond = 1e-15.  Floor the cutoff at the level the sweeps resolve (8 x the orthogonality tolerance).
    const double cutoff = fmax(rcond, 8.0 * tol) * smax;
    double min_kept = smax;
    for (int j = threadIdx.x; j < p; j += JAC_THREADS) {
        double sig = sqrt(w[j]);
        bool keep = sig > cutoff && sig > 0.0;
        w[j] = keep ? 1.0 / w[j] : 0.0;
        if (keep) min_kept = fmin(min_kept, sig);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) min_kept = fmin(min_kept, __shfl_xor_sync(0xffffffffu, min_kept, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = min_kept;
    __syncthreads();
    if (threadIdx.x == 0) {
        double mk = red[0];
        for (int i = 1; i < JAC_THREADS / 32; ++i) mk = fmin(mk, red[i]);
        st->s_max = smax; st->s_min_kept = mk;
    }
}

// P[i][r] = sum_j V[j][i] * w[j] * G[j][r]   (p x len), written as fp32 to out (optionally transposed)
constexpr int PT = 32;
__global__ void __launch_bounds__(PT * 8)
pinv_compose_kernel(const double* __restrict__ V, const double* __restrict__ G, const double* __restrict__ w,
                    int len, int p, bool transpose_out, float* __restrict__ out, int ldo, const JacobiState* st) {
    __shared__ double Vs[PT][PT + 1];
    if (st->fast_ok) return;
    __shared__ double Gs[PT][PT + 1];
    const int tx = threadIdx.x % PT, ty = threadIdx.x / PT;       // 32 x 8 threads, 4 rows each
    const int i0 = blockIdx.y * PT, r0 = blockIdx.x * PT;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int j0 = 0; j0 < p; j0 += PT) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            int jj = ty + s * 8, j = j0 + jj;
            Vs[jj][tx] = (j < p && i0 + tx < p) ? V[int64_t(j) * p + i0 + tx] * w[j] : 0.0;
            Gs[jj][tx] = (j < p && r0 + tx < len) ? G[int64_t(j) * len + r0 + tx] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int jj = 0; jj < PT; ++jj) {
            double g = Gs[jj][tx];
#pragma unroll
            for (int s = 0; s < 4; ++s) acc[s] = fma(Vs[jj][ty + s * 8], g, acc[s]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        int i = i0 + ty + s * 8, r = r0 + tx;
        if (i < p && r < len) {
            if (transpose_out) out[int64_t(r) * ldo + i] = float(acc[s]);
            else out[int64_t(i) * ldo + r] = float(acc[s]);
        }
    }
}

// sigma_j = |g_j| (unsorted), one warp per column
__global__ void jacobi_sigma_kernel(const double* __restrict__ G, int len, int p, double* __restrict__ sigma) {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= p) return;
    const double* g = G + int64_t(j) * len;
    double a = 0.0;
    for (int r = lane_id(); r < len; r += 32) a = fma(g[r], g[r], a);
    a = warp_sum(a);
    if (lane_id() == 0) sigma[j] = sqrt(a);
}

__global__ void pinv_export_cond_kernel(const JacobiState* st, double* cond_out) {
    cond_out[0] = st->s_max;
    cond_out[1] = st->s_min_kept;
}

// ---- fast path: pinv of a full-rank, well-conditioned T (len x p, len >= p) through the normal equations in fp64 ------------
//   S = T^T T  (p x p),  S = L L^T  (cooperative blocked Cholesky),  pinv(T) = S^-1 T^T: two triangular solves per column.
// The squared condition number is affordable in fp64 for cond(T) <~ 1e5 (error ~ cond^2 eps_64 ~ 1e-6 of the result; the
// intersections the CUR build sees have cond 20 .. 3.4e3, SURVEY 8c, i.e. ~1e-9, far below the fp32 output rounding).  The
// route is taken only when nobody asked for singular values (cond_out == NULL) and rcond <= 1e-10 (no truncation wanted);
// a pivot breakdown (rank deficiency) or diag_max / diag_min > CHOL_MAX_RATIO -- the ratio of the extreme diagonal entries
// of L is a lower bound of cond(T) -- hands the matrix to the Jacobi SVD above, decided ON THE DEVICE: every kernel of the
// route not taken returns at once, the host never synchronises.
constexpr int CH_NB = 32;
constexpr double CHOL_MAX_RATIO = 3.0e3;

// S[i][j] = sum_r T[r][i] T[r][j] (lower triangle incl. diagonal is what the factorisation reads; the full tile is written)
__global__ void __launch_bounds__(256)
gram_cols_kernel(const float* __restrict__ A, int lda, int len, int p, bool tall, double* __restrict__ S) {
    __shared__ double Ti[32][33], Tj[32][33];
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    if (j0 > i0) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 32 x 8 threads, 4 outputs each
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    auto at = [&](int r, int c) -> double {                           // T[r][c]
        if (r >= len || c >= p) return 0.0;
        return tall ? double(A[int64_t(r) * lda + c]) : double(A[int64_t(c) * lda + r]);
    };
    for (int r0 = 0; r0 < len; r0 += 32) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int rr = ty + 8 * s;
            // tall: consecutive threads read consecutive columns of one row (coalesced); wide: consecutive rows of one column
            if (tall) { Ti[rr][tx] = at(r0 + rr, i0 + tx); Tj[rr][tx] = at(r0 + rr, j0 + tx); }
            else      { Ti[tx][rr] = at(r0 + tx, i0 + rr); Tj[tx][rr] = at(r0 + tx, j0 + rr); }
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            const double b = Tj[r][tx];
#pragma unroll
            for (int s = 0; s < 4; ++s) acc[s] = fma(Ti[r][ty + 8 * s], b, acc[s]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int i = i0 + ty + 8 * s, j = j0 + tx;
        if (i < p && j < p) S[int64_t(i) * p + j] = acc[s];
    }
}

// In-place blocked right-looking Cholesky of the lower triangle of S (p x p, row-major), ONE cooperative launch.
__global__ void __launch_bounds__(256)
chol_coop_kernel(double* __restrict__ S, int p, JacobiState* st) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    __shared__ double D[CH_NB][CH_NB + 1];
    __shared__ double P2[CH_NB][CH_NB + 1];
    __shared__ int s_fail;
    const int tid = threadIdx.x;
    for (int k0 = 0; k0 < p; k0 += CH_NB) {
        const int kb = min(CH_NB, p - k0);
        // (1) diagonal block: unblocked Cholesky by CTA 0
        if (blockIdx.x == 0) {
            if (tid == 0) s_fail = 0;
            for (int e = tid; e < kb * kb; e += blockDim.x) D[e / kb][e % kb] = S[int64_t(k0 + e / kb) * p + k0 + e % kb];
            __syncthreads();
            for (int c = 0; c < kb; ++c) {
                if (tid == 0) {
                    const double d = D[c][c];
                    if (!(d > 0.0)) { s_fail = 1; D[c][c] = 1.0; } else D[c][c] = sqrt(d);
                }
                __syncthreads();
                const double piv = D[c][c];
                for (int r = c + 1 + tid; r < kb; r += blockDim.x) D[r][c] /= piv;
                __syncthreads();
                for (int e = tid; e < (kb - c - 1) * (kb - c - 1); e += blockDim.x) {
                    const int r = c + 1 + e / (kb - c - 1), q = c + 1 + e % (kb - c - 1);
                    if (q <= r) D[r][q] -= D[r][c] * D[q][c];
                }
                __syncthreads();
            }
            for (int e = tid; e < kb * kb; e += blockDim.x) {
                const int r = e / kb, q = e % kb;
                S[int64_t(k0 + r) * p + k0 + q] = q <= r ? D[r][q] : 0.0;
            }
            if (tid == 0) {
                double mn = st->diag_min, mx = st->diag_max;
                for (int c = 0; c < kb; ++c) { const double d = D[c][c]; mn = (k0 == 0 && c == 0) ? d : fmin(mn, d); mx = fmax(mx, d); }
                st->diag_min = mn; st->diag_max = mx;
                if (s_fail) st->chol_fail = 1;
            }
        }
        __threadfence();
        grid.sync();
        // (2) panel below the diagonal block: L_ik = S_ik . L_kk^-T, one thread per row
        const int below = p - (k0 + kb);
        if (below > 0) {
            for (int e = tid; e < kb * kb; e += blockDim.x) D[e / kb][e % kb] = S[int64_t(k0 + e / kb) * p + k0 + e % kb];
            __syncthreads();
            for (int r = blockIdx.x * blockDim.x + tid; r < below; r += gridDim.x * blockDim.x) {
                double* row = S + int64_t(k0 + kb + r) * p + k0;
                double x[CH_NB];
#pragma unroll
                for (int c = 0; c < CH_NB; ++c) x[c] = c < kb ? row[c] : 0.0;
#pragma unroll
                for (int c = 0; c < CH_NB; ++c) {
                    if (c < kb) {
                        double v = x[c];
#pragma unroll
                        for (int q = 0; q < CH_NB; ++q) if (q < c) v -= x[q] * D[c][q];
                        x[c] = v / D[c][c];
                    }
                }
#pragma unroll
                for (int c = 0; c < CH_NB; ++c) if (c < kb) row[c] = x[c];
            }
        }
        __threadfence();
        grid.sync();
        // (3) trailing update of the lower triangle: S_ij -= L_ik L_jk^T, 32 x 32 tiles round-robin over the CTAs
        if (below > 0) {
            const int nt = (below + CH_NB - 1) / CH_NB;
            for (int t = blockIdx.x; t < nt * nt; t += gridDim.x) {
                const int ti = t / nt, tj = t % nt;
                if (tj > ti) continue;
                const int i0 = k0 + kb + ti * CH_NB, j0 = k0 + kb + tj * CH_NB;
                __syncthreads();
                for (int e = tid; e < CH_NB * CH_NB; e += blockDim.x) {
                    const int r = e / CH_NB, c = e % CH_NB;
                    D[r][c] = (i0 + r < p && c < kb) ? S[int64_t(i0 + r) * p + k0 + c] : 0.0;
                    P2[r][c] = (j0 + r < p && c < kb) ? S[int64_t(j0 + r) * p + k0 + c] : 0.0;
                }
                __syncthreads();
                for (int e = tid; e < CH_NB * CH_NB; e += blockDim.x) {
                    const int r = e / CH_NB, c = e % CH_NB;
                    if (i0 + r < p && j0 + c < p && j0 + c <= i0 + r) {
                        double acc = 0.0;
#pragma unroll 8
                        for (int q = 0; q < CH_NB; ++q) acc = fma(D[r][q], P2[c][q], acc);
                        S[int64_t(i0 + r) * p + j0 + c] -= acc;
                    }
                }
            }
        }
        __threadfence();
        grid.sync();
    }
    if (blockIdx.x == 0 && tid == 0) {
        const bool ok = !st->chol_fail && st->diag_min > 0.0 && st->diag_max <= CHOL_MAX_RATIO * st->diag_min;
        st->fast_ok = ok ? 1 : 0;
        if (ok) { st->converged = 1; st->sweeps_done = 0; st->s_max = st->diag_max; st->s_min_kept = st->diag_min; }
    }
}

// Lt = L^T (upper triangle, row-major) so that the backward substitution also reads rows
__global__ void chol_transpose_kernel(const double* __restrict__ L, int p, double* __restrict__ Lt, const JacobiState* st) {
    if (!st->fast_ok) return;
    const int64_t total = int64_t(p) * p;
    for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
        const int i = int(t / p), j = int(t % p);
        Lt[t] = j >= i ? L[int64_t(j) * p + i] : 0.0;
    }
}

// X = (L L^T)^-1 T^T for a group of CH_RHS right-hand sides per CTA (column c of X belongs to row c of T); each group is
// independent, so there is no grid-wide synchronisation.  Blocked substitution: a 32 x 32 triangular solve by one warp per
// right-hand side, then every thread folds the solved block into the rows below (forward) / above (backward).
constexpr int CH_RHS = 8;
__global__ void __launch_bounds__(256)
chol_solve_kernel(const float* __restrict__ A, int lda, int len, int p, bool tall, const double* __restrict__ L,
                  const double* __restrict__ Lt, float* __restrict__ out, int ldo, const JacobiState* st) {
    extern __shared__ double xs[];                          // [CH_RHS][p]
    __shared__ double Dk[CH_NB][CH_NB + 1];
    if (!st->fast_ok) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int c0 = blockIdx.x * CH_RHS; c0 < len; c0 += gridDim.x * CH_RHS) {
        const int nr = min(CH_RHS, len - c0);
        __syncthreads();
        for (int e = tid; e < CH_RHS * p; e += blockDim.x) {
            const int q = e / p, i = e % p, c = c0 + q;
            xs[e] = q < nr ? (tall ? double(A[int64_t(c) * lda + i]) : double(A[int64_t(i) * lda + c])) : 0.0;
        }
        __syncthreads();
        for (int pass = 0; pass < 2; ++pass) {              // 0: L y = b (forward), 1: L^T x = y (backward)
            const double* M = pass == 0 ? L : Lt;           // row i of M holds the coefficients of equation i
            const int nblk = (p + CH_NB - 1) / CH_NB;
            for (int bi = 0; bi < nblk; ++bi) {
                const int blk = pass == 0 ? bi : nblk - 1 - bi;
                const int k0 = blk * CH_NB, kb = min(CH_NB, p - k0);
                for (int e = tid; e < kb * kb; e += blockDim.x) Dk[e / kb][e % kb] = M[int64_t(k0 + e / kb) * p + k0 + e % kb];
                __syncthreads();
                // triangular solve inside the block: warp q handles right-hand side q, lane = unknown
                if (warp < nr) {
                    double* x = xs + warp * p + k0;
                    double v = lane < kb ? x[lane] : 0.0;
                    if (pass == 0) {
                        for (int c = 0; c < kb; ++c) {
                            const double xc = __shfl_sync(0xffffffffu, v, c) / Dk[c][c];
                            if (lane == c) v = xc;
                            else if (lane > c && lane < kb) v -= Dk[lane][c] * xc;
                        }
                    } else {
                        for (int c = kb - 1; c >= 0; --c) {
                            const double xc = __shfl_sync(0xffffffffu, v, c) / Dk[c][c];
                            if (lane == c) v = xc;
                            else if (lane < c) v -= Dk[lane][c] * xc;
                        }
                    }
                    if (lane < kb) x[lane] = v;
                }
                __syncthreads();
                // fold the solved block into the remaining equations: rows below (forward) / above (backward)
                const int lo = pass == 0 ? k0 + kb : 0, hi = pass == 0 ? p : k0;
                for (int r = lo + tid; r < hi; r += blockDim.x) {
                    const double* mrow = M + int64_t(r) * p + k0;
                    double acc[CH_RHS];
#pragma unroll
                    for (int q = 0; q < CH_RHS; ++q) acc[q] = 0.0;
                    for (int c = 0; c < kb; ++c) {
                        const double m = mrow[c];
#pragma unroll
                        for (int q = 0; q < CH_RHS; ++q) acc[q] = fma(m, xs[q * p + k0 + c], acc[q]);
                    }
#pragma unroll
                    for (int q = 0; q < CH_RHS; ++q) xs[q * p + r] -= acc[q];
                }
                __syncthreads();
            }
        }
        // pinv(T)[i][c] = x_c[i]; tall: out is p x len (row i, column c); wide: out = pinv(T)^T is len x p
        for (int e = tid; e < nr * p; e += blockDim.x) {
            const int q = e / p, i = e % p, c = c0 + q;
            if (tall) out[int64_t(i) * ldo + c] = float(xs[q * p + i]);
            else out[int64_t(c) * ldo + i] = float(xs[q * p + i]);
        }
    }
}

// column-orthogonality tolerance of the sweeps: |g_i . g_j| <= tol |g_i| |g_j|
static double jacobi_tol(int len) { return fmax(1e-15, 4.0 * sqrt(double(len)) * 1.1102230246251565e-16); }

__global__ void jacobi_export_status_kernel(const JacobiState* st, double* out) {
    out[0] = st->s_max; out[1] = st->s_min_kept; out[2] = double(st->converged); out[3] = double(st->sweeps_done);
}

// {s_max, s_min_kept (pinv only), converged (0 / 1), sweeps done} of the last factorisation run on `workspace`
int jacobi_status(const void* workspace, double* status4_out, cudaStream_t stream) {
    jacobi_export_status_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<const JacobiState*>(workspace), status4_out);
    ANNCUR_LAUNCH_OK("jacobi_export_status_kernel");
    return ANNCUR_OK;
}

size_t pinv_workspace_bytes(int m, int n) {
    if (m <= 0 || n <= 0) return 256;
    size_t len = size_t(m > n ? m : n), p = size_t(m > n ? n : m);
    return align_up(sizeof(double) * (len * p + p * p + p), 256) + 256;
}

static int jacobi_factor(const float* A, int m, int n, int lda, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                         bool reset_state = false) {
    if (workspace_bytes < pinv_workspace_bytes(m, n)) {
        set_error("pinv workspace too small: %zu < %zu", workspace_bytes, pinv_workspace_bytes(m, n));
        return ANNCUR_E_WORKSPACE;
    }
    const bool tall = m >= n;
    const int len = tall ? m : n, p = tall ? n : m;
    JacobiState* st = reinterpret_cast<JacobiState*>(workspace);
    double* G = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 256);
    double* V = G + size_t(len) * p;

    if (reset_state) {
        jacobi_state_reset_kernel<<<1, 1, 0, stream>>>(st);
        ANNCUR_LAUNCH_OK("jacobi_state_reset_kernel");
    }
    jacobi_init_kernel<<<sm_count() * 4, 256, 0, stream>>>(A, m, n, lda, len, p, tall, G, V, st);
    ANNCUR_LAUNCH_OK("jacobi_init_kernel");
    const double tol = jacobi_tol(len);
    bool coop_done = false;
    if (p <= 1) {                                       // a single column is its own factorisation
        jacobi_sweep_end_kernel<<<1, 1, 0, stream>>>(st, tol);
        ANNCUR_LAUNCH_OK("jacobi_sweep_end_kernel");
    }
    if (p > 1) {
        const int p_even = (p + 1) & ~1;
        int dev = 0, coop = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jacobi_coop_kernel, JAC_THREADS, 0);
        if (coop && int64_t(per_sm) * sm_count() >= p_even / 2) {
            int len_ = len, p_ = p, pe_ = p_even;
            double tol_ = tol;
            void* args[] = {&G, &V, &len_, &p_, &pe_, &tol_, &st};
            cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(jacobi_coop_kernel), dim3(p_even / 2),
                                                        dim3(JAC_THREADS), args, 0, stream);
            if (e == cudaSuccess) { count_launch(1); coop_done = true; }
            else cudaGetLastError();                    // not launchable here: fall back to one launch per round
        }
    }
    if (p > 1 && !coop_done) {
        const int p_even = (p + 1) & ~1;
        for (int sweep = 0; sweep < JAC_MAX_SWEEPS; ++sweep) {
            for (int round = 0; round < p_even - 1; ++round) {
                jacobi_round_kernel<<<p_even / 2, JAC_THREADS, 0, stream>>>(G, V, len, p, p_even, round, tol, st);
                ANNCUR_LAUNCH_OK("jacobi_round_kernel");
            }
            jacobi_sweep_end_kernel<<<1, 1, 0, stream>>>(st, tol);
            ANNCUR_LAUNCH_OK("jacobi_sweep_end_kernel");
        }
    }
    return ANNCUR_OK;
}

int pinv_f32(const float* A, int m, int n, int lda, double rcond, float* out, int ldo, double* cond_out,
             void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (m <= 0 || n <= 0) return ANNCUR_OK;   // empty anchor set: pinv is the empty n x m matrix
    if (workspace_bytes < pinv_workspace_bytes(m, n)) {
        set_error("pinv workspace too small: %zu < %zu", workspace_bytes, pinv_workspace_bytes(m, n));
        return ANNCUR_E_WORKSPACE;
    }
    const bool tall = m >= n;
    const int len = tall ? m : n, p = tall ? n : m;
    JacobiState* st = reinterpret_cast<JacobiState*>(workspace);
    double* G = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 256);
    double* V = G + size_t(len) * p;
    double* w = V + size_t(p) * p;
    jacobi_state_reset_kernel<<<1, 1, 0, stream>>>(st);
    ANNCUR_LAUNCH_OK("jacobi_state_reset_kernel");
    // ---- Cholesky route (normal equations in fp64) for full-rank, well-conditioned inputs; see the kernels above ----------
    static const bool no_fast = getenv("ANNCUR_PINV_JACOBI_ONLY") != nullptr;
    if (!no_fast && cond_out == nullptr && rcond <= 1e-10 && p >= 2) {
        int dev = 0, coop = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chol_coop_kernel, 256, 0);
        const size_t solve_smem = sizeof(double) * size_t(CH_RHS) * p;
        if (coop && per_sm >= 1 && solve_smem <= 200 * 1024) {
            double* S = V;                  // p x p: Gram, then L in place;  L^T goes to the (free) Jacobi column area
            double* Lt = G;
            const dim3 ggrid((p + 31) / 32, (p + 31) / 32);
            gram_cols_kernel<<<ggrid, 256, 0, stream>>>(A, lda, len, p, tall, S);
            ANNCUR_LAUNCH_OK("gram_cols_kernel");
            int grid = per_sm * sm_count();
            const int tiles = ((p + CH_NB - 1) / CH_NB) * ((p + CH_NB - 1) / CH_NB);
            if (grid > tiles) grid = tiles;
            if (grid < 1) grid = 1;
            int p_ = p;
            void* args[] = {&S, &p_, &st};
            cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(chol_coop_kernel), dim3(grid), dim3(256), args, 0, stream);
            if (e == cudaSuccess) {
                count_launch(1);
                chol_transpose_kernel<<<sm_count() * 2, 256, 0, stream>>>(S, p, Lt, st);
                ANNCUR_LAUNCH_OK("chol_transpose_kernel");
                ANNCUR_CUDA_OK(cudaFuncSetAttribute(chol_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(solve_smem)));
                int sgrid = (len + CH_RHS - 1) / CH_RHS;
                if (sgrid > 2 * sm_count()) sgrid = 2 * sm_count();
                chol_solve_kernel<<<sgrid, 256, solve_smem, stream>>>(A, lda, len, p, tall, S, Lt, out, ldo, st);
                ANNCUR_LAUNCH_OK("chol_solve_kernel");
            } else {
                cudaGetLastError();         // not launchable here: the Jacobi route below serves the call
            }
        }
    }
    int rc = jacobi_factor(A, m, n, lda, workspace, workspace_bytes, stream);
    if (rc != ANNCUR_OK) return rc;
    jacobi_weights_kernel<<<1, JAC_THREADS, 0, stream>>>(G, len, p, rcond, jacobi_tol(len), w, st);
    ANNCUR_LAUNCH_OK("jacobi_weights_kernel");
    dim3 grid((len + PT - 1) / PT, (p + PT - 1) / PT);
    // tall: pinv(A) = P (n x m = p x len);  wide: pinv(A) = P^T (n x m = len x p)
    pinv_compose_kernel<<<grid, PT * 8, 0, stream>>>(V, G, w, len, p, !tall, out, ldo, st);
    ANNCUR_LAUNCH_OK("pinv_compose_kernel");
    if (cond_out) {
        pinv_export_cond_kernel<<<1, 1, 0, stream>>>(st, cond_out);
        ANNCUR_LAUNCH_OK("pinv_export_cond_kernel");
    }
    return ANNCUR_OK;
}

__global__ void jacobi_sigma_kernel(const double* __restrict__ G, int len, int p, double* __restrict__ sigma);

// Q[r][j] = G[j][r] / |g_j| (0 for columns below the cutoff): after the sweeps the columns of G = T V are mutually
// orthogonal, so these are the left singular vectors of T -- an orthonormal basis of its column space.  sigma[j] = |g_j|
// was written by jacobi_sigma_kernel; columns below cutoff_rel * max_j sigma[j] (null directions) come out as zero.
__global__ void __launch_bounds__(256)
orth_export_kernel(const double* __restrict__ G, int len, int p, double cutoff_rel, const double* __restrict__ sigma,
                   float* __restrict__ Q, int ldq) {
    __shared__ double red[8];
    __shared__ double s_max;
    double m = 0.0;
    for (int t = threadIdx.x; t < p; t += blockDim.x) m = fmax(m, sigma[t]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = fmax(m, red[w]);
        s_max = m;
    }
    __syncthreads();
    const int j = blockIdx.x;
    const double* g = G + int64_t(j) * len;
    const double nrm = sigma[j];
    const double inv = (nrm > cutoff_rel * s_max && nrm > 0.0) ? 1.0 / nrm : 0.0;
    for (int r = threadIdx.x; r < len; r += blockDim.x) Q[int64_t(r) * ldq + j] = float(g[r] * inv);
}

// Orthonormal basis of the column space of a TALL matrix A (m x n, m >= n): Q (m x n fp32, ldq), columns = left singular
// vectors (unsorted), zero columns for singular values below ~1e-12 of the largest; sigma_out (optional, n fp64).  The
// range-finder step of the randomised rank analysis (eval/compute_m2e_matrix_ranks.py:44-53 at sizes where a full SVD is
// out of reach).
int orthonormalize_f32(const float* A, int m, int n, int lda, float* Q, int ldq, double* sigma_out, void* workspace,
                       size_t workspace_bytes, cudaStream_t stream) {
    if (m <= 0 || n <= 0) return ANNCUR_OK;
    if (m < n) { set_error("orthonormalize: matrix must be tall (m = %d < n = %d)", m, n); return ANNCUR_E_INVALID; }
    int rc = jacobi_factor(A, m, n, lda, workspace, workspace_bytes, stream, true);
    if (rc != ANNCUR_OK) return rc;
    double* G = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 256);
    double* sig = sigma_out ? sigma_out : G + size_t(m) * n + size_t(n) * n;          // the weights slot of the workspace
    jacobi_sigma_kernel<<<(n + 7) / 8, 256, 0, stream>>>(G, m, n, sig);
    ANNCUR_LAUNCH_OK("jacobi_sigma_kernel");
    orth_export_kernel<<<n, 256, 0, stream>>>(G, m, n, 8.0 * jacobi_tol(m), sig, Q, ldq);
    ANNCUR_LAUNCH_OK("orth_export_kernel");
    return ANNCUR_OK;
}

// Singular values of A (m x n fp32) in fp64, unsorted: the same Jacobi factorisation without the inverse.
// Replaces the SVD inside np.linalg.matrix_rank (eval/compute_m2e_matrix_ranks.py:44-53).
int singular_values_f32(const float* A, int m, int n, int lda, double* sigma_out, void* workspace, size_t workspace_bytes,
                        cudaStream_t stream) {
    if (m <= 0 || n <= 0) return ANNCUR_OK;
    int rc = jacobi_factor(A, m, n, lda, workspace, workspace_bytes, stream, true);
    if (rc != ANNCUR_OK) return rc;
    const int len = m >= n ? m : n, p = m >= n ? n : m;
    const double* G = reinterpret_cast<const double*>(reinterpret_cast<const char*>(workspace) + 256);
    jacobi_sigma_kernel<<<(p + 7) / 8, 256, 0, stream>>>(G, len, p, sigma_out);
    ANNCUR_LAUNCH_OK("jacobi_sigma_kernel");
    return ANNCUR_OK;
}

}  // namespace anncur
