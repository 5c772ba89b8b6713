// fp32 FFMA GEMM on row-major operands, C[m x n] = A[m x k] . B[k x n], with two epilogues:
//   Store : write C                    (item-embedding build U.R, dense getters, FFMA score path)
//   Recon : never write C; accumulate per-row sum (C - T)^2 and sum T^2 against a target T
//           (reconstruction error of eval/run_retrieval_eval_wrt_exact_crossenc.py:146-147)
// This is the exact-fp32 path (plain FFMA, k accumulated in order); the tensor-core path is in
// score_topk_umma.cu.  128x128x16 CTA tile, 256 threads, 8x8 register tile (two 4-wide strips 64 apart in each
// direction, so that the 128-bit shared-memory reads of a quarter warp are conflict-free), double-buffered smem.
#include "common.cuh"
#include "kernels.h"

namespace anncur {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8, GEMM_THREADS = 256;

struct StoreEpilogue {
    float* C;
    int64_t ldc;
};
struct ReconEpilogue {
    const float* T;   // target matrix [m x n]
    int64_t ldt;
    double* err2;     // [m]
    double* norm2;    // [m]
};

template <class Epi>
__global__ void __launch_bounds__(GEMM_THREADS)
sgemm_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb, int m,
             int64_t n, int k, Epi epi) {
    __shared__ float As[2][BK][BM + 4];
    __shared__ float Bs[2][BK][BN + 4];
    __shared__ double rowacc[2][BM];

    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;                 // 16 x 16 threads, each 8 x 8 outputs:
    // rows  ty*4 + {0..3} and 64 + ty*4 + {0..3},  columns  tx*4 + {0..3} and 64 + tx*4 + {0..3}
    auto row_of = [&](int i) { return (i < 4 ? 0 : 64) + ty * 4 + (i & 3); };
    auto col_of = [&](int j) { return (j < 4 ? 0 : 64) + tx * 4 + (j & 3); };
    const int64_t col0 = int64_t(blockIdx.x) * BN;
    const int row0 = blockIdx.y * BM;

    // global -> smem assignment: A tile 128 x 16 (each thread 8 elements), B tile 16 x 128 (8 elements)
    const int a_r = tid / 2, a_c = (tid % 2) * 8;          // row in tile, first k column
    const int b_r = tid / 16, b_c = (tid % 16) * 8;        // k row, first n column

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float a_reg[8], b_reg[8];
    auto load_tiles = [&](int k0) {
        const int gr = row0 + a_r;
        const float* ap = A + int64_t(gr) * lda + k0 + a_c;
        if (gr < m && k0 + a_c + 7 < k && ((reinterpret_cast<uintptr_t>(ap) & 15) == 0)) {
            float4 v0 = __ldg(reinterpret_cast<const float4*>(ap));
            float4 v1 = __ldg(reinterpret_cast<const float4*>(ap + 4));
            a_reg[0] = v0.x; a_reg[1] = v0.y; a_reg[2] = v0.z; a_reg[3] = v0.w;
            a_reg[4] = v1.x; a_reg[5] = v1.y; a_reg[6] = v1.z; a_reg[7] = v1.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int gk = k0 + a_c + i;
                a_reg[i] = (gr < m && gk < k) ? __ldg(A + int64_t(gr) * lda + gk) : 0.f;
            }
        }
        const int gk = k0 + b_r;
        const int64_t gc = col0 + b_c;
        if (gk < k && gc + 7 < n && ((reinterpret_cast<uintptr_t>(B + int64_t(gk) * ldb + gc) & 15) == 0)) {
            float4 v0 = __ldg(reinterpret_cast<const float4*>(B + int64_t(gk) * ldb + gc));
            float4 v1 = __ldg(reinterpret_cast<const float4*>(B + int64_t(gk) * ldb + gc + 4));
            b_reg[0] = v0.x; b_reg[1] = v0.y; b_reg[2] = v0.z; b_reg[3] = v0.w;
            b_reg[4] = v1.x; b_reg[5] = v1.y; b_reg[6] = v1.z; b_reg[7] = v1.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                b_reg[i] = (gk < k && gc + i < n) ? __ldg(B + int64_t(gk) * ldb + gc + i) : 0.f;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; ++i) As[buf][a_c + i][a_r] = a_reg[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) Bs[buf][b_r][b_c + i] = b_reg[i];
    };

    const int n_kt = (k + BK - 1) / BK;
    if (n_kt > 0) { load_tiles(0); store_tiles(0); }
    __syncthreads();
    for (int kt = 0; kt < n_kt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < n_kt) load_tiles((kt + 1) * BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                float4 v = *reinterpret_cast<const float4*>(&As[buf][kk][row_of(i)]);
                a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
                float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][col_of(j)]);
                b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < n_kt) store_tiles(buf ^ 1);
        __syncthreads();
    }

    if constexpr (sizeof(Epi) == sizeof(StoreEpilogue)) {
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int gr = row0 + row_of(i);
            if (gr >= m) continue;
#pragma unroll
            for (int jh = 0; jh < TN; jh += 4) {
                const int64_t gc = col0 + col_of(jh);
                float* dst = epi.C + int64_t(gr) * epi.ldc + gc;
                if (gc + 3 < n && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                    *reinterpret_cast<float4*>(dst) = make_float4(acc[i][jh], acc[i][jh + 1], acc[i][jh + 2], acc[i][jh + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (gc + j < n) dst[j] = acc[i][jh + j];
                }
            }
        }
    } else {
        for (int t = tid; t < 2 * BM; t += GEMM_THREADS) (&rowacc[0][0])[t] = 0.0;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int gr = row0 + row_of(i);
            double e2 = 0.0, n2 = 0.0;
            if (gr < m) {
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    const int64_t gc = col0 + col_of(j);
                    if (gc < n) {
                        float t = __ldg(epi.T + int64_t(gr) * epi.ldt + gc);
                        float d = acc[i][j] - t;
                        e2 += double(d) * double(d);
                        n2 += double(t) * double(t);
                    }
                }
            }
            // the 16 threads sharing this row are the 16 lanes of a half warp (tx = lane % 16)
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                e2 += __shfl_xor_sync(0xffffffffu, e2, o);
                n2 += __shfl_xor_sync(0xffffffffu, n2, o);
            }
            if (tx == 0) { rowacc[0][row_of(i)] = e2; rowacc[1][row_of(i)] = n2; }
        }
        __syncthreads();
        if (tid < BM && row0 + tid < m) {
            atomicAdd(epi.err2 + row0 + tid, rowacc[0][tid]);
            atomicAdd(epi.norm2 + row0 + tid, rowacc[1][tid]);
        }
    }
}

int sgemm_rowmajor(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int m,
                   int64_t n, int k, cudaStream_t stream) {
    if (m <= 0 || n <= 0) return ANNCUR_OK;
    if (k <= 0) {   // empty anchor set: every score is 0 (reference grid point k_i = 0)
        ANNCUR_CUDA_OK(cudaMemset2DAsync(C, size_t(ldc) * 4, 0, size_t(n) * 4, size_t(m), stream));
        return ANNCUR_OK;
    }
    dim3 grid(unsigned((n + BN - 1) / BN), unsigned((m + BM - 1) / BM));
    sgemm_kernel<StoreEpilogue><<<grid, GEMM_THREADS, 0, stream>>>(A, lda, B, ldb, m, n, k, StoreEpilogue{C, ldc});
    ANNCUR_LAUNCH_OK("sgemm_kernel<store>");
    return ANNCUR_OK;
}

int recon_error(const float* Q, int64_t ldq, const float* E, int64_t lde, const float* A, int64_t lda, int n_rows,
                int64_t n_items, int k_dim, double* out_err2, double* out_norm2, cudaStream_t stream) {
    if (n_rows <= 0) return ANNCUR_OK;
    ANNCUR_CUDA_OK(cudaMemsetAsync(out_err2, 0, sizeof(double) * size_t(n_rows), stream));
    ANNCUR_CUDA_OK(cudaMemsetAsync(out_norm2, 0, sizeof(double) * size_t(n_rows), stream));
    if (n_items <= 0) return ANNCUR_OK;
    dim3 grid(unsigned((n_items + BN - 1) / BN), unsigned((n_rows + BM - 1) / BM));
    sgemm_kernel<ReconEpilogue><<<grid, GEMM_THREADS, 0, stream>>>(
        Q, ldq, E, lde, n_rows, n_items, k_dim > 0 ? k_dim : 0, ReconEpilogue{A, lda, out_err2, out_norm2});
    ANNCUR_LAUNCH_OK("sgemm_kernel<recon>");
    return ANNCUR_OK;
}

}  // namespace anncur
