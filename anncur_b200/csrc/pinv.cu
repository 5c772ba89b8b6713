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
};

__global__ void jacobi_init_kernel(const float* __restrict__ A, int m, int n, int lda, int len, int p,
                                   bool tall, double* __restrict__ G, double* __restrict__ V, JacobiState* st) {
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
    if (st->converged) return;
    jacobi_pair(G, V, len, p, p_even, round, tol, st);
}

__device__ __forceinline__ void jacobi_sweep_end(JacobiState* st, double tol) {
    double off = __longlong_as_double((long long)st->max_off_bits);
    st->sweeps_done += 1;
    if (off <= tol) st->converged = 1;
    st->max_off_bits = 0ull;
}

__global__ void jacobi_sweep_end_kernel(JacobiState* st, double tol) {
    if (st->converged) return;
    jacobi_sweep_end(st, tol);
}

// All sweeps in ONE cooperative launch: one CTA per column pair, a grid-wide barrier between the rounds of the
// tournament instead of a kernel boundary (the factorisation is launch-latency-bound: ~10 sweeps x (p - 1) rounds of
// microsecond-sized work).  Used when p_even / 2 CTAs are co-resident; the multi-launch path above is the fallback.
__global__ void __launch_bounds__(JAC_THREADS)
jacobi_coop_kernel(double* __restrict__ G, double* __restrict__ V, int len, int p, int p_even, double tol, JacobiState* st) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
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
                    int len, int p, bool transpose_out, float* __restrict__ out, int ldo) {
    __shared__ double Vs[PT][PT + 1];
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

static int jacobi_factor(const float* A, int m, int n, int lda, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (workspace_bytes < pinv_workspace_bytes(m, n)) {
        set_error("pinv workspace too small: %zu < %zu", workspace_bytes, pinv_workspace_bytes(m, n));
        return ANNCUR_E_WORKSPACE;
    }
    const bool tall = m >= n;
    const int len = tall ? m : n, p = tall ? n : m;
    JacobiState* st = reinterpret_cast<JacobiState*>(workspace);
    double* G = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 256);
    double* V = G + size_t(len) * p;

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
    int rc = jacobi_factor(A, m, n, lda, workspace, workspace_bytes, stream);
    if (rc != ANNCUR_OK) return rc;
    const bool tall = m >= n;
    const int len = tall ? m : n, p = tall ? n : m;
    JacobiState* st = reinterpret_cast<JacobiState*>(workspace);
    double* G = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 256);
    double* V = G + size_t(len) * p;
    double* w = V + size_t(p) * p;
    jacobi_weights_kernel<<<1, JAC_THREADS, 0, stream>>>(G, len, p, rcond, jacobi_tol(len), w, st);
    ANNCUR_LAUNCH_OK("jacobi_weights_kernel");
    dim3 grid((len + PT - 1) / PT, (p + PT - 1) / PT);
    // tall: pinv(A) = P (n x m = p x len);  wide: pinv(A) = P^T (n x m = len x p)
    pinv_compose_kernel<<<grid, PT * 8, 0, stream>>>(V, G, w, len, p, !tall, out, ldo);
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
    int rc = jacobi_factor(A, m, n, lda, workspace, workspace_bytes, stream);
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
    int rc = jacobi_factor(A, m, n, lda, workspace, workspace_bytes, stream);
    if (rc != ANNCUR_OK) return rc;
    const int len = m >= n ? m : n, p = m >= n ? n : m;
    const double* G = reinterpret_cast<const double*>(reinterpret_cast<const char*>(workspace) + 256);
    jacobi_sigma_kernel<<<(p + 7) / 8, 256, 0, stream>>>(G, len, p, sigma_out);
    ANNCUR_LAUNCH_OK("jacobi_sigma_kernel");
    return ANNCUR_OK;
}

}  // namespace anncur
