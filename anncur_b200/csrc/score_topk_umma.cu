// K3 + K4: fused approximate-score GEMM and streaming per-row top-k on tcgen05 / TMEM / TMA.
//
//   scores = Q (B x K) . E (K x N)            eval/matrix_approx_zeshel.py:109-119 (get_complete_row)
//   torch.topk(scores, k, dim=1)               eval/matrix_approx_zeshel.py:121-126 (topk_in_row)
// without ever writing the B x N score matrix.
//
// Precision kinds
//   F32X3 : every fp32 operand x (scaled by a power of two into fp16 range) is split into two
//           fp16 terms x = h + l (22-23 significant bits).  Three tcgen05 kind::f16 passes
//           h.h + h.l + l.h accumulate in one fp32 TMEM tile -- fp32-grade scores at 1/3 of the
//           f16 tensor rate, and each operand still costs 4 bytes per element like plain fp32.
//   BF16  : single bf16 pass.
//   F32R  : fp32 results at the one-pass rate ("filter and refine").  MAIN is ONE f16 pass on the high halves, but
//           the contraction carries one extra anchor slot: the item side holds ||e_n|| 2^-5, the query side
//           +-f(K) ||q|| 2^-5, so the accumulator is A_n +- b_n with b_n = f(K) 2^-10 ||q|| ||e_n|| >= |S_n - A_n|, f(K) =
//           1.04 + K-proportional terms for the fp32 accumulation of the tensor pipe and of the re-scoring (f32r_slot_factor)
//           (Cauchy-Schwarz over the per-element fp16 rounding errors; S_n = the fp32 score).  SAMPLE runs with
//           the minus sign (lower bounds -> threshold T <= k-th best S), MAIN with the plus sign (upper bounds:
//           every item that can be in the top-k is pushed), and refine_topk_keylists (refine_topk.cu) re-scores
//           the ~k candidates whose upper bound reaches the k-th best fp32 score from an item-major fp32 copy of E
//           kept in the packed index.  No step depends on a statistical error model: rows that end short, overflow
//           a list or fail the final certificate (k-th refined score > push threshold) go to REDO, which is the
//           3-pass kind on the same planes.
//
// Layout in HBM (built once per index by pack_items, per batch for Q by pack_queries):
//   plane[kb][row][64]  16-bit elements, kb = k / 64; a TMA box {64, rows, 1} is one contiguous
//   rows*128-byte span and lands in shared memory in the canonical K-major SWIZZLE_128B layout.
//
// Kernel: persistent, one CTA pair per SM pair (cta_group::2; single CTAs for one query tile), 320 threads, warp-specialised
//   warp 0      TMA producer  (ring of 3 stages x 1 k-block of 64, 3-pass kind; 3 stages x 2 k-blocks, one-pass kinds)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer, accumulator 128 x 256 fp32,
//               double-buffered in the 512 TMEM columns
//   warps 2-9   epilogue: thread = one query row x one half of the tile's columns (two warps per TMEM
//               lane quarter, so every scheduler has two epilogue warps to interleave); tcgen05.ld 32
//               columns at a time (the next load is
//               in flight while the current one is processed), compare with the row's threshold,
//               push the rare survivors as 64-bit keys to the row's candidate list (L2-resident);
//               if a list fills up the warp radix-selects it back to k entries and tightens the
//               threshold (streaming fallback -- stays correct for adversarial inputs).
// The same pipeline also serves the dense products (EPI_DENSE: store the scores; EPI_ERR: squared reconstruction error).
// Work item = (chunk of item tiles, 128-query tile or pair of them), ordered chunk-major so that co-resident CTAs
// stream the same slice of E and it is read from HBM once.
//
// Thresholds.  Pushing is only cheap when the threshold is tight from the first tile on, so a call
// runs the kernel up to three times:
//   SAMPLE  every G-th item (a strided TMA view of the same packed planes, no copy) is scored and
//           the epilogue only records the maximum of each 32-column group; the j-th largest group
//           maximum of a row (sample_threshold_kernel) is <= the j-th best sampled score, so at
//           least j sampled items reach it.  j is chosen so that P[Binomial(k-1, 1/G) >= j] <= 1e-6:
//           then at least k items of the full row reach the threshold except with that probability,
//           and about j*G items do (a few hundred).
//   MAIN    all items, thresholds preloaded; survivors >= threshold are pushed.
//   REDO    rows that ended with fewer than k candidates (the rare miss, found by the select
//           kernel) are re-run with the threshold reset to -inf and streaming compaction; query
//           tiles without such a row are skipped, so normally this launch exits at once.
// Small item sets (too few sampled groups) skip SAMPLE/REDO and run MAIN in streaming mode.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>
#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace anncur {

// ------------------------------------------------------------------------------------------------
constexpr int BLOCK_M = 128;           // queries per tile (TMEM lanes)
constexpr int BLOCK_N = 256;           // items per tile (TMEM columns per accumulator buffer)
constexpr int BLOCK_K = 64;            // 16-bit elements per k-block = 128 bytes = one SWIZZLE_128B row
constexpr int K32_PER_KB = BLOCK_K / 32;   // the packing kernels work in 32-wide chunks (one element per lane)
constexpr int UMMA_K = 16;
constexpr int A_PLANE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB
constexpr int B_PLANE_BYTES = BLOCK_N * BLOCK_K * 2;   // 32 KB

// element (anchor index kidx, row) of a plane[kb][row][BLOCK_K] with `rows` rows per k-block
__host__ __device__ __forceinline__ int64_t plane_off(int kidx, int64_t row, int64_t rows) {
    return (int64_t(kidx / BLOCK_K) * rows + row) * BLOCK_K + kidx % BLOCK_K;
}
constexpr int NUM_EPI_WARPS = 8;           // two warps per TMEM lane quarter, each takes half of a tile's columns
constexpr int EPI_HALVES = NUM_EPI_WARPS / 4;
constexpr int FUSED_THREADS = 32 * (2 + NUM_EPI_WARPS);
constexpr int TMEM_COLS = 512;
constexpr int DENSE_TBUF_FLOATS = 16 * 36;  // per epilogue warp: 16 rows x (32 + 4 pad) floats of the dense-store transposition
// mbarrier watchdog: a wait that outlasts this many SM cycles (~3 s) sets the workspace's error flag and traps instead of
// hanging the GPU for good.  A trap takes the whole CUDA context with it, so the limit is a run-time knob:
// ANNCUR_WAIT_TIMEOUT_CYCLES=<n> (0 = no watchdog, e.g. under cuda-gdb / compute-sanitizer, where clock64 keeps running
// while the kernel is stopped).
constexpr unsigned long long WAIT_TIMEOUT_CYCLES = 6000000000ull;

// CG = CTAs per MMA (1, or 2 = CTA pair: each CTA stages its own 128 queries and HALF of the 256-item tile)
template <int PASSES, int CG> struct StageCfg {
    static constexpr int kBBytes = B_PLANE_BYTES / CG;
    // one k-block (32 wide) of operands: A (+ its low half) and this CTA's share of B (+ its low half)
    static constexpr int kSubBytes = PASSES == 3 ? 2 * (A_PLANE_BYTES + kBBytes) : (A_PLANE_BYTES + kBBytes);
    // k-blocks per pipeline stage.  The one-pass kinds issue only two MMAs per k-block, and the single issuing thread's
    // barrier wait + commit per stage then costs as much as the tensor work it feeds (measured: tensor pipe 69 % busy).
    // Four k-blocks per stage (8 MMAs per wait) took bf16 MAIN from 0.357 to 0.277 ms at C2; six per stage with two
    // stages is slower again (3.29 vs 2.81 ms at N = 1M), and the fp32-grade kind (6 MMAs per k-block) gains nothing.
    // With 64-wide k-blocks: pairs 2 k-blocks x 3 stages (1 x 6: 0.295 ms, 3 x 2: 0.292 ms against 0.265 ms at C2); single CTAs
    // (one query tile, HBM-bound) 1 x 4 (0.159 against 0.163 ms at N = 1M, B = 64).
    static constexpr int kKbPerStage = PASSES == 3 ? 1 : (CG == 2 ? 2 : 1);
    static constexpr int kStageBytes = kKbPerStage * kSubBytes;
    static constexpr int kStages = PASSES == 3 ? (CG == 2 ? 3 : 2) : (CG == 2 ? 3 : 4);
};

enum : int { MODE_MAIN = 0, MODE_SAMPLE = 1 };
// epilogue family (compile time): per-row top-k (MAIN / SAMPLE), dense store of the scores, or squared reconstruction error
enum : int { EPI_TOPK = 0, EPI_DENSE = 1, EPI_ERR = 2 };

struct FusedParams {
    int mode;                // MODE_MAIN: push survivors;  MODE_SAMPLE: record 32-column group maxima
    int n_queries;
    int n_items;             // valid columns of this pass (all items, or the sampled ones)
    int num_kb;
    int k;
    int m_tiles;
    int n_tiles;
    int n_chunks;
    uint64_t* cand;          // [n_queries][n_chunks * EPI_HALVES][cap]
    uint32_t* counts;        // [n_queries][n_chunks * EPI_HALVES]
    uint32_t* thr_shared;    // [n_queries] ordered-uint lower bound (exclusive) on useful scores
    const uint32_t* mtile_flags;   // optional [m_tiles]: work items of query tiles whose flag is 0 are skipped
    float* smax;             // MODE_SAMPLE: [n_queries][n_smax] group maxima
    int n_smax;
    int smax_wide;           // MODE_SAMPLE: one maximum per (tile, column half) = 128 sampled items instead of one per 32
    int close_compact;       // MODE_MAIN: cut lists back to k at the end of a work item (streaming mode: tightens the shared bound)
    int a_last_kb;           // k-block coordinate of the query high plane used for the LAST k-block (F32R: the +b / -b / 0 variants)
    int filter;              // MODE_MAIN, F32R: scores are upper bounds -> lists are never cut back; a list that fills up is
                             // marked (top bit of its count) and the row goes to REDO
    int* error_flag;
    unsigned long long wait_timeout;   // mbarrier watchdog in SM cycles, 0 = off (see WAIT_TIMEOUT_CYCLES)
    unsigned wait_sleep_ns;            // back-off between polls of the long waits (epilogue: accumulator ready; producer: stage free)
    // EPI_DENSE: out[row][col] = score; EPI_ERR: err2[row] += (score - exact[row][col])^2, norm2[row] += exact[row][col]^2
    const float* row_inv_scale;     // accumulator (scaled units) * row_inv_scale[row] = score
    float* dense_out;  int64_t ldo;
    int dense_vec_ok;               // dense_out is 16-byte aligned and ldo a multiple of 4: rows can be stored as float4
    const float* exact; int64_t lda;
    double* err2; double* norm2;
};

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
struct WaitCtx { int* error_flag; unsigned long long timeout; unsigned sleep_ns; };
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, const WaitCtx& w) {
    if (mbar_try_wait(bar, parity)) return;
    const unsigned long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (w.timeout != 0ull && clock64() - t0 > w.timeout) {
            if (w.error_flag) atomicExch(w.error_flag, 1);
            __trap();
        }
    }
}
// The waits that are LONG by construction (the epilogue warps wait most of a tile's MMA time for the accumulator, the
// producer for a free stage) can back off between polls (ANNCUR_WAIT_SLEEP_NS, default 0 = plain polling).  Measured as a way
// to save power in the tensor-bound, power-capped MAIN kernel -- two thirds of its issued instructions are these polls --
// and found neutral (20 / 100 / 400 ns: +-0.5 % at N = 1M and at C2, profiles/r2x_wait_sleep.txt): mbarrier.try_wait already
// suspends the thread in hardware.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, const WaitCtx& w) {
    if (mbar_try_wait(bar, parity)) return;
    const unsigned long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (w.sleep_ns != 0u) __nanosleep(w.sleep_ns);
        if (w.timeout != 0ull && clock64() - t0 > w.timeout) {
            if (w.error_flag) atomicExch(w.error_flag, 1);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// ---- CTA-pair (cta_group::2) forms.  Inside a pair a shared::cluster address with bit 24 cleared names the
//      same offset in the even (leader) CTA's shared memory.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {          // arrive on the leader CTA's copy of `bar`
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & PEER_BIT_MASK) : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    // data lands in THIS CTA's shared memory, the bytes are counted on the LEADER's barrier
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {           // arrives on `bar` in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(uint16_t(3)) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, SWIZZLE_128B canonical layout: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= uint64_t((saddr >> 4) & 0x3fffu);        // start address
    d |= uint64_t(1) << 16;                       // leading byte offset (unused for swizzled K-major)
    d |= uint64_t(1024 >> 4) << 32;               // stride byte offset: 8 rows x 128 B
    d |= uint64_t(1) << 46;                       // descriptor version (sm_100)
    d |= uint64_t(2) << 61;                       // layout type SWIZZLE_128B
    return d;
}
// The same descriptor split into its address word and its constant word: inside a pipeline stage the operands differ only by
// multiples of 16 bytes, so the issuer derives each descriptor from the stage's address word with ONE add (the sum stays
// inside the 14-bit field: shared-memory offsets are below 256 KB).
constexpr uint32_t SMEM_DESC_HI = uint32_t(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3fffu) | (1u << 16); }
__device__ __forceinline__ uint64_t smem_desc_at(uint32_t lo, uint32_t byte_offset) {
    return (uint64_t(SMEM_DESC_HI) << 32) | uint64_t(lo + (byte_offset >> 4));
}
// tcgen05.ld of 32 consecutive fp32 columns of this thread's TMEM lane: issue only (asynchronous) ...
__device__ __forceinline__ void tmem_ld_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// ... and the wait for every outstanding load.  The registers are in/out operands so that no use of
// them can be scheduled above the wait.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.wait::ld.sync.aligned;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
          "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
          "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
          "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
        :: "memory");
}

// ---- warp-cooperative compaction of one row's candidate list back to its best k -----------------
// Called by all 32 lanes.  list: n (> k) keys in global memory.  Returns the k-th best key.
template <int CPL>
__device__ __forceinline__ uint64_t warp_compact_list(uint64_t* list, uint32_t n, uint32_t k, uint32_t* hist) {
    const uint32_t lane = lane_id();
    uint64_t key[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
        uint32_t t = uint32_t(i) * 32u + lane;
        key[i] = t < n ? raw_to_key(__ldcg(list + t)) : 0ull;
    }
    uint64_t prefix = 0, mask = 0;
    uint32_t need = k;
    for (int shift = 56; shift >= 0; shift -= 8) {
#pragma unroll
        for (int b = 0; b < 8; ++b) hist[lane * 8 + b] = 0;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < CPL; ++i)
            if (key[i] != 0ull && (key[i] & mask) == prefix) atomicAdd(&hist[uint32_t(key[i] >> shift) & 255u], 1u);
        __syncwarp();
        uint32_t c[8], lane_sum = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) { c[b] = hist[lane * 8 + b]; lane_sum += c[b]; }
        uint32_t incl = lane_sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t t = __shfl_down_sync(0xffffffffu, incl, off);
            if (lane + off < 32) incl += t;
        }
        uint32_t running = incl - lane_sum;
        bool found = false;
        uint32_t d = 0, new_need = 0, bucket = 0;
#pragma unroll
        for (int b = 7; b >= 0; --b) {
            if (!found && running + c[b] >= need) { found = true; d = lane * 8 + b; new_need = need - running; bucket = c[b]; }
            running += c[b];
        }
        const uint32_t ballot = __ballot_sync(0xffffffffu, found);
        const int src = ballot ? 31 - __clz(int(ballot)) : 0;
        d = __shfl_sync(0xffffffffu, d, src);
        new_need = __shfl_sync(0xffffffffu, new_need, src);
        bucket = __shfl_sync(0xffffffffu, bucket, src);
        __syncwarp();
        if (ballot == 0) break;                          // fewer than `need` keys left: keep them all
        prefix |= uint64_t(d) << shift;
        mask |= 0xffull << shift;
        need = new_need;
        if (bucket == need) break;                       // digit bucket taken whole
    }
    // keep keys >= the selected prefix; write them back densely; find the smallest kept key
    uint32_t running = 0;
    uint64_t kth = ~0ull;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
        const bool keep = key[i] != 0ull && (key[i] & mask) >= prefix;
        const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            list[running + __popc(ballot & ((1u << lane) - 1u))] = key_to_raw(key[i]);
            kth = key[i] < kth ? key[i] : kth;
        }
        running += __popc(ballot);
    }
    return warp_min_u64(kth);
}

// ------------------------------------------------------------------------------------------------
template <int PASSES, bool BF16, int CPL, int CG, int EPI>
__global__ void __launch_bounds__(FUSED_THREADS, 1)
fused_score_topk_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                        const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
                        const FusedParams p) {
    using Cfg = StageCfg<PASSES, CG>;
    constexpr int NS = Cfg::kStages;
    constexpr uint32_t CAP = uint32_t(CPL) * 32u;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NS * Cfg::kStageBytes);
    // bars: full[NS] | empty[NS] | tmem_full[2] | tmem_empty[2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * NS + 4);
    uint32_t* hist_all = tmem_ptr_smem + 4;                      // NUM_EPI_WARPS x 256
    float* dense_tbuf = reinterpret_cast<float*>(hist_all + NUM_EPI_WARPS * 256);   // EPI_DENSE: NUM_EPI_WARPS x DENSE_TBUF_FLOATS

    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * uint32_t(s); };
    auto empty_bar = [&](int s) { return bar_base + 8u * uint32_t(NS + s); };
    auto tfull_bar = [&](int b) { return bar_base + 8u * uint32_t(2 * NS + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * uint32_t(2 * NS + 2 + b); };

    const int warp = threadIdx.x >> 5;
    const uint32_t lane = lane_id();
    const WaitCtx wctx{p.error_flag, p.wait_timeout, p.wait_sleep_ns};

    // REDO launch without a flagged row (the normal case): nothing to do, leave before any TMEM / barrier set-up
    if (p.mtile_flags != nullptr && __ldg(p.mtile_flags + p.m_tiles) == 0u) return;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB0) : "memory");
        if (PASSES == 3) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB1) : "memory");
        }
        // full: the leader's copy collects one arrival per CTA of the pair plus all TMA bytes; empty / tmem_full:
        // one tcgen05.commit (multicast to both CTAs); tmem_empty: the leader's copy collects the epilogue warps
        // of both CTAs
        for (int s = 0; s < NS; ++s) { mbar_init(full_bar(s), CG); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), CG * NUM_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tcgen05_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr_smem);

    // A CTA pair works on two adjacent query tiles (one per CTA) against the same item tiles.
    const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    const int unit = int(blockIdx.x) / CG, n_units = int(gridDim.x) / CG;          // CTA (pair) index / count
    const int m_groups = (p.m_tiles + CG - 1) / CG;
    const int total_items = p.n_chunks * m_groups;
    // a work item is skipped by all roles (of both CTAs) when none of its query tiles is flagged (REDO launch only)
    auto item_live = [&](int m_group) {
        if (p.mtile_flags == nullptr) return true;
        bool live = false;
        for (int r = 0; r < CG; ++r) {
            const int mt = m_group * CG + r;
            if (mt < p.m_tiles && __ldg(p.mtile_flags + mt) != 0u) live = true;
        }
        return live;
    };

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            auto load = [&](uint32_t dst, const CUtensorMap* map, uint32_t bar, int c1, int kb) {
                if (CG == 2) tma_load_3d_pair(dst, map, bar, 0, c1, kb); else tma_load_3d(dst, map, bar, 0, c1, kb);
            };
            for (int w = unit; w < total_items; w += n_units) {
                const int chunk = w / m_groups, m_tile = (w % m_groups) * CG + int(cta_rank);
                if (!item_live(w % m_groups)) continue;
                const int t0 = int((int64_t(chunk) * p.n_tiles) / p.n_chunks);
                const int t1 = int((int64_t(chunk + 1) * p.n_tiles) / p.n_chunks);
                for (int tile = t0; tile < t1; ++tile) {
                    const int item0 = tile * BLOCK_N + int(cta_rank) * (BLOCK_N / CG);     // this CTA's share of the item tile
                    for (int kb0 = 0; kb0 < p.num_kb; kb0 += Cfg::kKbPerStage) {
                        const int n_sub = p.num_kb - kb0 < Cfg::kKbPerStage ? p.num_kb - kb0 : Cfg::kKbPerStage;
                        mbar_wait_relaxed(empty_bar(stage), phase ^ 1u, wctx);
                        if (leader) mbar_expect_tx(full_bar(stage), uint32_t(CG * n_sub * Cfg::kSubBytes));
                        for (int u = 0; u < n_sub; ++u) {
                            const int kb = kb0 + u;
                            const int akb = kb == p.num_kb - 1 ? p.a_last_kb : kb;
                            const uint32_t sb = smem_base + uint32_t(stage) * Cfg::kStageBytes + uint32_t(u) * Cfg::kSubBytes;
                            if (PASSES == 3) {
                                load(sb, &tmA0, full_bar(stage), m_tile * BLOCK_M, akb);
                                load(sb + A_PLANE_BYTES, &tmA1, full_bar(stage), m_tile * BLOCK_M, kb);
                                load(sb + 2 * A_PLANE_BYTES, &tmB0, full_bar(stage), item0, kb);
                                load(sb + 2 * A_PLANE_BYTES + Cfg::kBBytes, &tmB1, full_bar(stage), item0, kb);
                            } else {
                                load(sb, &tmA0, full_bar(stage), m_tile * BLOCK_M, akb);
                                load(sb + A_PLANE_BYTES, &tmB0, full_bar(stage), item0, kb);
                            }
                        }
                        if (!leader) mbar_arrive_leader(full_bar(stage));
                        if (++stage == NS) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =======================================
        // One thread of the leader CTA.  (A warp-uniform variant -- all 32 lanes walk the loops, an elected lane issues, as
        // CUTLASS does -- was measured: 14 instead of 21 instructions per MMA and the same kernel time, so the issuing thread
        // is not what the tensor pipe waits for at 8 MMAs per barrier; the simpler form stayed.)
        if (lane == 0 && leader) {
            // instruction descriptor: D fp32, A/B f16 or bf16, both K-major, N = 256, M = 128 per CTA
            const uint32_t fmt = BF16 ? 1u : 0u;
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (uint32_t(BLOCK_N >> 3) << 17) | (uint32_t((BLOCK_M * CG) >> 4) << 24);
            auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t acc) {
                if (CG == 2) umma_f16_pair(d, da, db, idesc, acc); else umma_f16(d, da, db, idesc, acc);
            };
            auto commit = [&](uint32_t bar) {
                if (CG == 2) umma_commit_pair(bar); else umma_commit(bar);
            };
            int stage = 0; uint32_t phase = 0;
            int buf = 0; uint32_t acc_phase = 0;
            for (int w = unit; w < total_items; w += n_units) {
                const int chunk = w / m_groups;
                if (!item_live(w % m_groups)) continue;
                const int t0 = int((int64_t(chunk) * p.n_tiles) / p.n_chunks);
                const int t1 = int((int64_t(chunk + 1) * p.n_tiles) / p.n_chunks);
                for (int tile = t0; tile < t1; ++tile) {
                    mbar_wait(tempty_bar(buf), acc_phase ^ 1u, wctx);
                    tcgen05_fence_after();
                    const uint32_t d_tmem = tmem_base + uint32_t(buf) * BLOCK_N;
                    for (int kb0 = 0; kb0 < p.num_kb; kb0 += Cfg::kKbPerStage) {
                        const int n_sub = p.num_kb - kb0 < Cfg::kKbPerStage ? p.num_kb - kb0 : Cfg::kKbPerStage;
                        mbar_wait(full_bar(stage), phase, wctx);
                        tcgen05_fence_after();
                        const uint32_t lo0 = smem_desc_lo(smem_base + uint32_t(stage) * Cfg::kStageBytes);
#pragma unroll
                        for (int u = 0; u < Cfg::kKbPerStage; ++u) {
                            if (u < n_sub) {
#pragma unroll
                                for (int ks = 0; ks < BLOCK_K / UMMA_K; ++ks) {
                                    const uint32_t off = uint32_t(u) * Cfg::kSubBytes + uint32_t(ks) * UMMA_K * 2;   // sub-block + bytes inside the 128 B row
                                    const uint32_t accum = (kb0 | u | ks) != 0 ? 1u : 0u;
                                    if (PASSES == 3) {
                                        const uint64_t a_h = smem_desc_at(lo0, off);
                                        const uint64_t a_l = smem_desc_at(lo0, off + A_PLANE_BYTES);
                                        const uint64_t b_h = smem_desc_at(lo0, off + 2 * A_PLANE_BYTES);
                                        const uint64_t b_l = smem_desc_at(lo0, off + 2 * A_PLANE_BYTES + Cfg::kBBytes);
                                        mma(d_tmem, a_h, b_h, accum);
                                        mma(d_tmem, a_h, b_l, 1u);
                                        mma(d_tmem, a_l, b_h, 1u);
                                    } else {
                                        mma(d_tmem, smem_desc_at(lo0, off), smem_desc_at(lo0, off + A_PLANE_BYTES), accum);
                                    }
                                }
                            }
                        }
                        commit(empty_bar(stage));                      // smem slot free (in both CTAs) once these MMAs retire
                        if (++stage == NS) { stage = 0; phase ^= 1u; }
                    }
                    commit(tfull_bar(buf));                            // accumulator ready for the epilogue (of both CTAs)
                    if (++buf == 2) { buf = 0; acc_phase ^= 1u; }
                }
            }
        }
    } else {
        // ===================================== epilogue =========================================
        const int q = warp & 3;                                   // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;                         // which half of a tile's 32-column groups
        constexpr int GROUPS_PER_HALF = BLOCK_N / 32 / EPI_HALVES;
        const int n_lists = p.n_chunks * EPI_HALVES;
        uint32_t* hist = hist_all + (warp - 2) * 256;
        int buf = 0; uint32_t acc_phase = 0;
        const uint32_t k = uint32_t(p.k);
        const bool sample = p.mode == MODE_SAMPLE;
        for (int w = unit; w < total_items; w += n_units) {
            const int chunk = w / m_groups, m_tile = (w % m_groups) * CG + int(cta_rank);
            if (!item_live(w % m_groups)) continue;
            const int t0 = int((int64_t(chunk) * p.n_tiles) / p.n_chunks);
            const int t1 = int((int64_t(chunk + 1) * p.n_tiles) / p.n_chunks);
            const int row = m_tile * BLOCK_M + q * 32 + int(lane);
            const bool row_ok = row < p.n_queries;
            const int row_c = row_ok ? row : 0;
            uint64_t* list = p.cand + (int64_t(row_c) * n_lists + chunk * EPI_HALVES + half) * int64_t(CAP);
            float* smax_row = p.smax + int64_t(row_c) * p.n_smax;
            uint32_t cnt = 0, overflow = 0;
            float thr_own = -INFINITY;
            float thr = INFINITY;
            const float row_scale = (EPI != EPI_TOPK && row_ok) ? __ldg(p.row_inv_scale + row) : 0.f;
            double err_acc = 0.0, norm_acc = 0.0;
            float smax_acc = -INFINITY;

            // one 32-column group of this thread's row, already in registers
            auto process = [&](const uint32_t (&r)[32], int tile, int c) {
                const int col0 = tile * BLOCK_N + c * 32;
                if constexpr (EPI == EPI_DENSE) {
                    // Each thread holds 32 consecutive scores of ONE row.  Stored as they are, one warp store instruction
                    // touches 32 different rows with 16 bytes each (measured 0.8 - 1.4 TB/s).  So the 32 x 32 block is turned
                    // through shared memory, 16 rows at a time (row stride 36 floats: conflict-free 16-byte writes and
                    // reads), and every store instruction writes 4 complete 128-byte row segments.
                    if (col0 + 32 <= p.n_items && p.dense_vec_ok) {
                        float* tb = dense_tbuf + (warp - 2) * DENSE_TBUF_FLOATS;
                        const int row0 = m_tile * BLOCK_M + q * 32;
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            if (int(lane >> 4) == hh) {
                                float* wr = tb + (lane & 15) * 36;
#pragma unroll
                                for (int j = 0; j < 32; j += 4)
                                    *reinterpret_cast<float4*>(wr + j) =
                                        make_float4(__uint_as_float(r[j]) * row_scale, __uint_as_float(r[j + 1]) * row_scale,
                                                    __uint_as_float(r[j + 2]) * row_scale, __uint_as_float(r[j + 3]) * row_scale);
                            }
                            __syncwarp();
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int rr = int(lane >> 3) + 4 * i;                 // row inside this half
                                const int r_abs = row0 + 16 * hh + rr;
                                const float4 v = *reinterpret_cast<const float4*>(tb + rr * 36 + 4 * int(lane & 7));
                                if (r_abs < p.n_queries)
                                    __stcs(reinterpret_cast<float4*>(p.dense_out + int64_t(r_abs) * p.ldo + col0 + 4 * int(lane & 7)), v);
                            }
                            __syncwarp();
                        }
                        return;
                    }
                    // ragged last tile / unaligned output: this thread's 32 scores straight to its row
                    if (!row_ok) return;
                    float* dst = p.dense_out + int64_t(row) * p.ldo + col0;
                    if (col0 + 32 <= p.n_items && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            __stcs(reinterpret_cast<float4*>(dst + j),
                                   make_float4(__uint_as_float(r[j]) * row_scale, __uint_as_float(r[j + 1]) * row_scale,
                                               __uint_as_float(r[j + 2]) * row_scale, __uint_as_float(r[j + 3]) * row_scale));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < p.n_items) dst[j] = __uint_as_float(r[j]) * row_scale;
                    }
                    return;
                } else if constexpr (EPI == EPI_ERR) {
                    if (!row_ok) return;
                    const float* a = p.exact + int64_t(row) * p.lda + col0;
                    float e2 = 0.f, n2 = 0.f;
                    if (col0 + 32 <= p.n_items && (reinterpret_cast<uintptr_t>(a) & 15) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 v = __ldcs(reinterpret_cast<const float4*>(a + j));
                            const float d0 = __uint_as_float(r[j]) * row_scale - v.x, d1 = __uint_as_float(r[j + 1]) * row_scale - v.y;
                            const float d2 = __uint_as_float(r[j + 2]) * row_scale - v.z, d3 = __uint_as_float(r[j + 3]) * row_scale - v.w;
                            e2 = fmaf(d0, d0, e2); e2 = fmaf(d1, d1, e2); e2 = fmaf(d2, d2, e2); e2 = fmaf(d3, d3, e2);
                            n2 = fmaf(v.x, v.x, n2); n2 = fmaf(v.y, v.y, n2); n2 = fmaf(v.z, v.z, n2); n2 = fmaf(v.w, v.w, n2);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < p.n_items) {
                                const float v = a[j], d = __uint_as_float(r[j]) * row_scale - v;
                                e2 = fmaf(d, d, e2); n2 = fmaf(v, v, n2);
                            }
                    }
                    err_acc += double(e2); norm_acc += double(n2);       // fp32 inside 32 columns, fp64 across them
                    return;
                }
                float g[4], g4[8];
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) {
                    float m01 = fmaxf(__uint_as_float(r[gi * 8 + 0]), __uint_as_float(r[gi * 8 + 1]));
                    float m23 = fmaxf(__uint_as_float(r[gi * 8 + 2]), __uint_as_float(r[gi * 8 + 3]));
                    float m45 = fmaxf(__uint_as_float(r[gi * 8 + 4]), __uint_as_float(r[gi * 8 + 5]));
                    float m67 = fmaxf(__uint_as_float(r[gi * 8 + 6]), __uint_as_float(r[gi * 8 + 7]));
                    g4[gi * 2] = fmaxf(m01, m23);
                    g4[gi * 2 + 1] = fmaxf(m45, m67);
                    g[gi] = fmaxf(g4[gi * 2], g4[gi * 2 + 1]);
                }
                if (sample) {
                    float m = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
                    if (col0 + 32 > p.n_items) {                      // ragged last group: ignore the zero-filled tail
                        m = -INFINITY;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < p.n_items) m = fmaxf(m, __uint_as_float(r[j]));
                    }
                    if (p.smax_wide) {
                        smax_acc = fmaxf(smax_acc, m);
                        if ((c % GROUPS_PER_HALF) == GROUPS_PER_HALF - 1) {
                            if (row_ok) smax_row[tile * EPI_HALVES + half] = smax_acc;
                            smax_acc = -INFINITY;
                        }
                    } else if (row_ok) {
                        smax_row[tile * (BLOCK_N / 32) + c] = m;
                    }
                    return;
                }
                // survivors: test 4-column sub-groups (their maxima fall out of the tree above), push RAW entries
                const bool ragged = col0 + 32 > p.n_items;                  // only the last tile of the item set
#pragma unroll
                for (int h4 = 0; h4 < 8; ++h4) {
                    if (g4[h4] > thr) {
                        // branch-free inside: slot of element j = cnt + (survivors among elements < j)
                        bool keep[4];
                        uint32_t slot[4];
                        uint32_t run = cnt;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            keep[j] = __uint_as_float(r[h4 * 4 + j]) > thr && (!ragged || col0 + h4 * 4 + j < p.n_items);
                            slot[j] = run;
                            run += keep[j] ? 1u : 0u;
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (keep[j]) list[slot[j]] = (uint64_t(uint32_t(col0 + h4 * 4 + j)) << 32) | uint64_t(r[h4 * 4 + j]);
                        cnt = run;
                    }
                }
            };
            // make room: a 32-column group can add at most 32 survivors to a list
            auto make_room = [&]() {
                if (p.filter) {                                  // upper-bound scores cannot be cut back to k: give the row up
                    if (cnt > CAP - 32u) { overflow = 0x80000000u; thr = INFINITY; }
                    return;
                }
                uint32_t full_mask = __ballot_sync(0xffffffffu, cnt > CAP - 32u);
                while (full_mask) {
                    const int src = __ffs(int(full_mask)) - 1;
                    full_mask &= full_mask - 1;
                    const uint32_t n_src = __shfl_sync(0xffffffffu, cnt, src);
                    const uint64_t lp = __shfl_sync(0xffffffffu, reinterpret_cast<uint64_t>(list), src);
                    __syncwarp();
                    const uint64_t kth = warp_compact_list<CPL>(reinterpret_cast<uint64_t*>(lp), n_src, k, hist);
                    __syncwarp();
                    if (int(lane) == src) {
                        cnt = k;
                        thr_own = key_score(kth);
                        thr = fmaxf(thr, thr_own);
                        atomicMax(p.thr_shared + row, float_to_ordered(thr_own) - 1u);
                    }
                }
            };

            for (int tile = t0; tile < t1; ++tile) {
                // the cross-chunk bound (exclusive): fixed for the whole pass when it was sampled, refreshed per tile
                // in streaming mode (other chunks of the row tighten it as they compact)
                if (EPI == EPI_TOPK && !sample && (tile == t0 || p.close_compact != 0)) {
                    thr = INFINITY;
                    if (row_ok && overflow == 0u) thr = fmaxf(thr_own, ordered_to_float(__ldcg(p.thr_shared + row)));
                }
                mbar_wait_relaxed(tfull_bar(buf), acc_phase, wctx);
                tcgen05_fence_after();
                const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(buf) * BLOCK_N;
                uint32_t ra[32], rb[32];
                const int c_begin = half * GROUPS_PER_HALF, c_end = c_begin + GROUPS_PER_HALF;
                tmem_ld_issue(taddr + uint32_t(c_begin * 32), ra);
#pragma unroll 1
                for (int c = c_begin; c < c_end; c += 2) {
                    tmem_ld_wait(ra);
                    tmem_ld_issue(taddr + uint32_t((c + 1) * 32), rb);
                    if (EPI == EPI_TOPK && !sample) make_room();
                    process(ra, tile, c);
                    tmem_ld_wait(rb);
                    if (c + 2 < c_end) {
                        tmem_ld_issue(taddr + uint32_t((c + 2) * 32), ra);
                    } else {
                        // the whole accumulator is in registers: hand the TMEM buffer back before the last group
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) { if (CG == 2) mbar_arrive_leader(tempty_bar(buf)); else mbar_arrive(tempty_bar(buf)); }
                    }
                    if (EPI == EPI_TOPK && !sample) make_room();
                    process(rb, tile, c + 1);
                }
                if (++buf == 2) { buf = 0; acc_phase ^= 1u; }
            }
            if (EPI == EPI_ERR && row_ok) {
                atomicAdd(p.err2 + row, err_acc);
                atomicAdd(p.norm2 + row, norm_acc);
            }
            if (EPI != EPI_TOPK || sample) continue;
            // close the work item (streaming mode): lists longer than k are cut back, which publishes a tighter bound
            uint32_t over_mask = __ballot_sync(0xffffffffu, p.close_compact != 0 && row_ok && cnt > k);
            while (over_mask) {
                const int src = __ffs(int(over_mask)) - 1;
                over_mask &= over_mask - 1;
                const uint32_t n_src = __shfl_sync(0xffffffffu, cnt, src);
                const uint64_t lp = __shfl_sync(0xffffffffu, reinterpret_cast<uint64_t>(list), src);
                __syncwarp();
                const uint64_t kth = warp_compact_list<CPL>(reinterpret_cast<uint64_t*>(lp), n_src, k, hist);
                __syncwarp();
                if (int(lane) == src) {
                    cnt = k;
                    atomicMax(p.thr_shared + row, float_to_ordered(key_score(kth)) - 1u);
                }
            }
            if (row_ok) p.counts[int64_t(row) * n_lists + chunk * EPI_HALVES + half] = cnt | overflow;
        }
    }

    tcgen05_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---- operand packing ------------------------------------------------------------------------------
__global__ void absmax_kernel(const float* __restrict__ X, int64_t ld, int rows, int64_t cols, uint32_t* out_bits) {
    float m = 0.f;
    const int64_t total = int64_t(rows) * cols;
    for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
        float x = fabsf(X[(t / cols) * ld + (t % cols)]);
        if (x <= FLT_MAX) m = fmaxf(m, x);      // ignore inf / NaN when choosing the scale
    }
    m = warp_max_f(m);
    if (lane_id() == 0) atomicMax(out_bits, __float_as_uint(m));
}

// power-of-two scale that maps [0, maxabs] into [0, 2^14): fp16 then holds h and the residual l
__device__ __forceinline__ float pow2_scale_for(float maxabs) {
    if (!(maxabs > 0.f)) return 1.f;
    int e;
    frexpf(maxabs, &e);                          // maxabs = f * 2^e, f in [0.5, 1)
    e = 14 - e;
    e = e > 100 ? 100 : (e < -100 ? -100 : e);
    return ldexpf(1.f, e);
}

__global__ void finish_scale_kernel(const uint32_t* maxabs_bits, float* scale_out, bool bf16) {
    scale_out[0] = bf16 ? 1.f : pow2_scale_for(__uint_as_float(maxabs_bits[0]));
}

// rowmax[i] = scale * max_n |E[i, n]| for every anchor dimension i (0 for the zero-padded ones): bounds the error
// of the single-pass SAMPLE scores, see pack_queries_kernel.  One CTA per row of E.
__global__ void __launch_bounds__(256)
row_absmax_kernel(const float* __restrict__ E, int64_t lde, int64_t n_items, int k_dim, const float* __restrict__ scale_p,
                  float* __restrict__ rowmax) {
    __shared__ float red[8];
    const int i = blockIdx.x;
    float m = 0.f;
    if (i < k_dim)
        for (int64_t n = threadIdx.x; n < n_items; n += blockDim.x) {
            float x = fabsf(E[int64_t(i) * lde + n]);
            if (x <= FLT_MAX) m = fmaxf(m, x);
        }
    m = warp_max_f(m);
    if (lane_id() == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
        rowmax[i] = m * scale_p[0];
    }
}

template <bool BF16>
__device__ __forceinline__ void split_store(float x, uint16_t* plane_h, uint16_t* plane_l, int64_t off) {
    if (BF16) {
        plane_h[off] = __bfloat16_as_ushort(__float2bfloat16_rn(x));
    } else {
        __half h = __float2half_rn(x);
        __half l = __float2half_rn(x - __half2float(h));
        plane_h[off] = __half_as_ushort(h);
        plane_l[off] = __half_as_ushort(l);
    }
}

// E (k_dim x N, N contiguous) -> plane[kb][n][32].  CTA = one k-block x 64 items, transposed via smem.
template <bool BF16>
__global__ void __launch_bounds__(256)
pack_items_kernel(const float* __restrict__ E, int64_t lde, int64_t n_items, int k_dim, const float* __restrict__ scale_p,
                  uint16_t* __restrict__ plane_h, uint16_t* __restrict__ plane_l) {
    __shared__ float tile[32][65];
    const float scale = scale_p[0];
    const int kb = blockIdx.y;                              // 32-wide chunk of anchors
    const int64_t n0 = int64_t(blockIdx.x) * 64;
    for (int e = threadIdx.x; e < 32 * 64; e += 256) {
        int kk = e / 64, j = e % 64;
        int kidx = kb * 32 + kk;
        int64_t n = n0 + j;
        tile[kk][j] = (kidx < k_dim && n < n_items) ? E[int64_t(kidx) * lde + n] * scale : 0.f;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 64 * 32; e += 256) {
        int j = e / 32, kk = e % 32;
        int64_t n = n0 + j;
        if (n < n_items) split_store<BF16>(tile[kk][j], plane_h, plane_l, plane_off(kb * 32 + kk, n, n_items));
    }
}

// F32R: factor on the query-side bound slot, see pack_queries_kernel.  Relative to the fp16-rounding bound 2^-10 |q'| |e'|:
// tensor-pipe accumulation (K/16 + 1) 2^-22 / 2^-10, re-scoring accumulation (K/32 + 8) 2^-24 / 2^-10, and a flat 4 %.
__host__ __device__ __forceinline__ float f32r_slot_factor(int k_dim) {
    return 1.04f + (float(k_dim) * (1.0f / 16.0f) + 1.0f) * 0x1p-12f + (float(k_dim) * (1.0f / 32.0f) + 8.0f) * 0x1p-14f;
}

// Q (B x k_dim, k contiguous) -> plane[kb][b][32], one warp per query row, per-row power-of-two scale.
// F32R: slot k_dim of the last k-block of the high plane holds +f(K) ||q'|| 2^-5 (rounded up); k-blocks num_kb and
// num_kb + 1 are copies of that last k-block with the slot set to -f(K) ||q'|| 2^-5 and to 0 (see FusedParams::a_last_kb).
template <int KIND>
__global__ void __launch_bounds__(256)
pack_queries_kernel(const float* __restrict__ Q, int64_t ldq, int n_queries, int plane_rows, int k_dim, int num_kb,
                    const float* __restrict__ e_scale, uint16_t* __restrict__ plane_h, uint16_t* __restrict__ plane_l,
                    float* __restrict__ row_inv_scale, uint32_t* __restrict__ thr_shared,
                    uint32_t* __restrict__ mtile_flags, const float* __restrict__ e_rowmax, float* __restrict__ row_delta) {
    constexpr bool BF16 = KIND == ANNCUR_KIND_BF16;
    constexpr bool X3 = KIND == ANNCUR_KIND_F32X3;
    constexpr bool R = KIND == ANNCUR_KIND_F32R;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= plane_rows) return;
    const uint32_t lane = lane_id();
    if (row >= n_queries) {                                // padding rows of the last query tile read as zero
        const int n_chunks = (num_kb + (R ? 2 : 0)) * K32_PER_KB;
        for (int u = 0; u < n_chunks; ++u) {
            plane_h[plane_off(u * 32 + int(lane), row, plane_rows)] = 0;
            if (!BF16) plane_l[plane_off(u * 32 + int(lane), row, plane_rows)] = 0;
        }
        return;
    }
    const float* q = Q + int64_t(row) * ldq;
    float scale = 1.f;
    float bound = 0.f, norm2 = 0.f;
    constexpr int RC = 16;                       // rows of up to 512 anchors are read once and kept in registers
    const int n32 = num_kb * K32_PER_KB;         // 32-wide chunks (one element per lane)
    if (n32 <= RC) {
        float xr[RC];
#pragma unroll
        for (int u = 0; u < RC; ++u) {
            const int kidx = u * 32 + int(lane);
            xr[u] = (u < n32 && kidx < k_dim) ? q[kidx] : 0.f;
        }
        if (!BF16) {
            float m = 0.f;
#pragma unroll
            for (int u = 0; u < RC; ++u) { const float x = fabsf(xr[u]); if (x <= FLT_MAX) m = fmaxf(m, x); }
            scale = pow2_scale_for(warp_max_f(m));
        }
#pragma unroll
        for (int u = 0; u < RC; ++u) {
            if (u < n32) {
                const int kidx = u * 32 + int(lane);
                const float x = xr[u] * scale;
                split_store<BF16>(x, plane_h, plane_l, plane_off(kidx, row, plane_rows));
                if (X3 && kidx < k_dim) { const float t = x * e_rowmax[kidx]; bound = fmaf(t, t, bound); }
                if (R) norm2 = fmaf(x, x, norm2);
            }
        }
    } else {
        if (!BF16) {
            float m = 0.f;
            for (int kidx = lane; kidx < k_dim; kidx += 32) {
                float x = fabsf(q[kidx]);
                if (x <= FLT_MAX) m = fmaxf(m, x);
            }
            scale = pow2_scale_for(warp_max_f(m));
        }
        for (int kb = 0; kb < n32; ++kb) {
            int kidx = kb * 32 + int(lane);
            float x = kidx < k_dim ? q[kidx] * scale : 0.f;
            split_store<BF16>(x, plane_h, plane_l, plane_off(kidx, row, plane_rows));
            if (X3 && kidx < k_dim) { const float t = x * e_rowmax[kidx]; bound = fmaf(t, t, bound); }
            if (R) norm2 = fmaf(x, x, norm2);
        }
    }
    if (X3) {
        // SAMPLE scores use the high fp16 halves only.  Each product q_i e_i is then off by q_i e_i (eps_q + eps_e)
        // with rounding errors |eps| <= 2^-12 (std 2^-12 / sqrt 3), so a sampled score is off by a sum of K such
        // terms: std <= 0.82 * 2^-12 * sqrt(sum_i q_i^2 max_n e_in^2) (scaled units of the accumulator).  The
        // threshold is lowered by 2^-9 * sqrt(...) ~ 10 sigma; should that ever not be enough for a row, it ends
        // MAIN with fewer than k candidates and the REDO pass recomputes it, so the margin is a speed knob only.
        bound = warp_sum(bound);
        if (lane == 0) row_delta[row] = sqrtf(bound) * (1.0f / 512.0f);
    } else if (lane == 0) {
        row_delta[row] = 0.f;
    }
    if (R) {
        // Error bound of the one-pass score A = sum_i h(q'_i) h(e'_i) against S = sum_i q'_i e'_i (scaled operands):
        // |x - h(x)| <= 2^-11 |x| + 2^-25 (fp16 rounding, subnormals included), so by Cauchy-Schwarz
        //   |S - A| <= (2^-10 + 2^-22) sqrt(|q'|^2 + K 2^-28) sqrt(|e'_n|^2 + K 2^-28).
        // The slot product is f(K) x that (rounded up on both sides), f(K) = f32r_slot_factor(k_dim): a reserve that GROWS
        // with K for what is not fp16 operand rounding -- the fp32 accumulation of the tensor pipe (K / 16 + 1 additions into
        // the TMEM accumulator, each allowed a full 2 ulp = 2^-22 of sum_i |q'_i e'_i| <= |q'| |e'_n|: truncating adders,
        // coherent signs), the accumulation of the re-scoring kernel (K / 32 + 8 roundings of 2^-24 per lane chain and
        // shuffle tree) and 4 % for the fp32 rounding of the norms and the slot product.  f(500) = 1.049, f(8192) = 1.181.
        norm2 = warp_sum(norm2);
        const float slot = sqrtf(norm2 + float(k_dim) * 0x1p-28f) * (f32r_slot_factor(k_dim) * 1.001f * 0x1p-5f);
        const uint16_t slot_p = __half_as_ushort(__float2half_ru(slot));
        const int kb_s = num_kb - 1;                                         // the slot lives at anchor index k_dim, in the last k-block
#pragma unroll
        for (int u = 0; u < K32_PER_KB; ++u) {
            const int kidx = kb_s * BLOCK_K + u * 32 + int(lane);
            const float x = kidx < k_dim ? q[kidx] * scale : 0.f;
            const uint16_t h = __half_as_ushort(__float2half_rn(x));
            const bool is_slot = kidx == k_dim;
            plane_h[plane_off(kidx, row, plane_rows)] = is_slot ? slot_p : h;                                  // +b: upper bounds
            plane_h[plane_off(kidx + BLOCK_K, row, plane_rows)] = is_slot ? uint16_t(slot_p | 0x8000u) : h;    // -b: lower bounds
            plane_h[plane_off(kidx + 2 * BLOCK_K, row, plane_rows)] = h;                                       //  0: plain scores
        }
    }
    if (lane == 0) {
        row_inv_scale[row] = 1.f / (scale * e_scale[0]);
        thr_shared[row] = float_to_ordered(-INFINITY);
        if ((row % BLOCK_M) == 0) mtile_flags[row / BLOCK_M] = 0u;
        if (row == 0) mtile_flags[(n_queries + BLOCK_M - 1) / BLOCK_M] = 0u;      // number of flagged rows
    }
}

// F32R: slot k_dim of every item = ||e'_n|| 2^-5 rounded up (see pack_queries_kernel), one thread per item
__global__ void __launch_bounds__(256)
item_bound_slot_kernel(const float* __restrict__ E, int64_t lde, int64_t n_items, int k_dim, const float* __restrict__ scale_p,
                       uint16_t* __restrict__ plane_h) {
    const float scale = scale_p[0];
    for (int64_t n = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; n < n_items; n += int64_t(gridDim.x) * blockDim.x) {
        float s = 0.f;
        for (int i = 0; i < k_dim; ++i) { const float x = E[int64_t(i) * lde + n] * scale; s = fmaf(x, x, s); }
        const float ne = sqrtf(s + float(k_dim) * 0x1p-28f) * (1.001f * 0x1p-5f);
        plane_h[plane_off(k_dim, n, n_items)] = __half_as_ushort(__float2half_ru(ne));
    }
}

// F32R: item-major fp32 copy ET[n][ld] of E (k_dim x N), values untouched; 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256)
transpose_items_kernel(const float* __restrict__ E, int64_t lde, int64_t n_items, int k_dim, int ld, float* __restrict__ ET) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t n0 = int64_t(blockIdx.x) * 32;
    const int k0 = blockIdx.y * 32;
    for (int j = ty; j < 32; j += 8) {
        const int kk = k0 + j;
        const int64_t n = n0 + tx;
        tile[j][tx] = (kk < k_dim && n < n_items) ? E[int64_t(kk) * lde + n] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int64_t n = n0 + j;
        const int kk = k0 + tx;
        if (n < n_items && kk < ld) ET[n * ld + kk] = tile[tx][j];
    }
}

__global__ void fill_zero_scores_kernel(int n_queries, int k, int64_t n_items, int64_t idx_offset, float* out_vals,
                                        int64_t* out_idx) {
    const int64_t total = int64_t(n_queries) * k;
    for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
        int j = int(t % k);
        bool ok = j < n_items;
        out_vals[t] = ok ? 0.f : ANNCUR_PAD_VAL;
        out_idx[t] = ok ? j + idx_offset : -1;
    }
}

// ---- SAMPLE -> per-row threshold -------------------------------------------------------------------
// One CTA per query row: j-th largest of the row's n_smax group maxima (MSD radix select on the order-preserving
// 32-bit image of the floats, staged in shared memory when they fit), published as the exclusive bound of MAIN.
constexpr int THR_THREADS = 128;
constexpr int THR_SMEM_KEYS = 8192;

__global__ void __launch_bounds__(THR_THREADS)
sample_threshold_kernel(const float* __restrict__ smax, int n_smax, int n_queries, int j, const float* __restrict__ row_delta,
                        uint32_t* __restrict__ thr_shared) {
    __shared__ uint32_t keys_s[THR_SMEM_KEYS];
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_red[2][THR_THREADS / 32];
    __shared__ uint32_t s_prefix, s_mask, s_need, s_stop;
    const int row = blockIdx.x, tid = threadIdx.x;
    const float* v = smax + int64_t(row) * n_smax;
    const bool staged = n_smax <= THR_SMEM_KEYS;
    auto key_at = [&](int t) { return staged ? keys_s[t] : float_to_ordered(__ldcg(v + t)); };
    // stage + min / max: group maxima of one row lie in a narrow band, the select starts at the first byte in which
    // the largest and the smallest differ
    uint32_t kmax = 0u, kmin = 0xffffffffu;
    for (int t = tid; t < n_smax; t += THR_THREADS) {
        const uint32_t key = float_to_ordered(__ldcg(v + t));
        if (staged) keys_s[t] = key;
        kmax = max(kmax, key);
        kmin = min(kmin, key);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    }
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = kmax; s_red[1][tid >> 5] = kmin; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < THR_THREADS / 32; ++w) { kmax = max(kmax, s_red[0][w]); kmin = min(kmin, s_red[1][w]); }
        const int first_shift = (31 - __clz(int((kmax ^ kmin) | 1u))) & ~7;
        s_mask = first_shift >= 24 ? 0u : ~((1u << (first_shift + 8)) - 1u);
        s_prefix = kmax & s_mask;
        s_need = uint32_t(j);
        s_stop = uint32_t(first_shift);                 // reused as the current shift
    }
    __syncthreads();
    const bool have = n_smax >= j;
    bool ok = have;
    for (int shift = int(s_stop); ok && shift >= 0; shift -= 8) {
        for (int t = tid; t < 256; t += THR_THREADS) hist[t] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix, mask = s_mask;
        for (int t = tid; t < n_smax; t += THR_THREADS) {
            const uint32_t key = key_at(t);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid < 32) {
            const uint32_t need = s_need;
            uint32_t c[8], lane_sum = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) { c[b] = hist[tid * 8 + b]; lane_sum += c[b]; }
            uint32_t incl = lane_sum;                   // inclusive suffix sum towards the higher digits
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                uint32_t t = __shfl_down_sync(0xffffffffu, incl, off);
                if (tid + off < 32) incl += t;
            }
            uint32_t running = incl - lane_sum;
            bool found = false;
            uint32_t d = 0, new_need = 0;
#pragma unroll
            for (int b = 7; b >= 0; --b) {
                if (!found && running + c[b] >= need) { found = true; d = uint32_t(tid) * 8 + b; new_need = need - running; }
                running += c[b];
            }
            const uint32_t ballot = __ballot_sync(0xffffffffu, found);
            if (ballot != 0) {
                const int src = 31 - __clz(int(ballot));
                d = __shfl_sync(0xffffffffu, d, src);
                new_need = __shfl_sync(0xffffffffu, new_need, src);
            }
            if (tid == 0) {
                if (ballot == 0) s_need = 0xffffffffu;  // fewer than j values: no threshold
                else { s_prefix = prefix | (d << shift); s_mask = mask | (0xffu << shift); s_need = new_need; }
            }
        }
        __syncthreads();
        ok = s_need != 0xffffffffu;
    }
    if (tid == 0) {
        // s_prefix = ordered image of the j-th largest maximum; scores >= it are kept (exclusive bound = it - 1)
        const uint32_t lowest = float_to_ordered(-INFINITY);
        uint32_t bound = lowest;
        if (ok && s_prefix > lowest) {
            const float t = ordered_to_float(s_prefix) - row_delta[row];      // see pack_queries_kernel
            bound = float_to_ordered(t);
            bound = bound > lowest ? bound - 1u : lowest;
        }
        thr_shared[row] = bound;
    }
}

// Same result, one WARP per row with the row's group maxima in registers (n_smax <= 32 VPL): the j-th largest key is
// found bit by bit (largest v with #{keys >= v} >= j; 32 rounds of VPL compares + one warp sum) -- no shared memory,
// no block barriers.  Large batches: 23 -> ~5 us at C2 (4096 rows x 200 maxima).
template <int VPL>
__global__ void __launch_bounds__(128)
sample_threshold_warp_kernel(const float* __restrict__ smax, int n_smax, int n_queries, int j, const float* __restrict__ row_delta,
                             uint32_t* __restrict__ thr_shared) {
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= n_queries) return;
    const int lane = int(lane_id());
    const float* v = smax + int64_t(row) * n_smax;
    uint32_t key[VPL];
#pragma unroll
    for (int u = 0; u < VPL; ++u) {
        const int t = u * 32 + lane;
        key[u] = t < n_smax ? float_to_ordered(__ldcg(v + t)) : 0u;
    }
    uint32_t result = 0u;
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t trial = result | (1u << bit);
        int cnt = 0;
#pragma unroll
        for (int u = 0; u < VPL; ++u) cnt += key[u] >= trial ? 1 : 0;
        cnt = warp_sum(cnt);
        if (cnt >= j) result = trial;
    }
    if (lane == 0) {
        const uint32_t lowest = float_to_ordered(-INFINITY);
        uint32_t bound = lowest;
        if (n_smax >= j && result > lowest) {
            const float t = ordered_to_float(result) - row_delta[row];      // see pack_queries_kernel
            bound = float_to_ordered(t);
            bound = bound > lowest ? bound - 1u : lowest;
        }
        thr_shared[row] = bound;
    }
}

static int launch_sample_threshold(const float* smax, int n_smax, int n_queries, int j, const float* row_delta, uint32_t* thr,
                                   cudaStream_t stream) {
    const int wgrid = (n_queries + 3) / 4;
    if (n_smax <= 256) sample_threshold_warp_kernel<8><<<wgrid, 128, 0, stream>>>(smax, n_smax, n_queries, j, row_delta, thr);
    else if (n_smax <= 512) sample_threshold_warp_kernel<16><<<wgrid, 128, 0, stream>>>(smax, n_smax, n_queries, j, row_delta, thr);
    else if (n_smax <= 1024) sample_threshold_warp_kernel<32><<<wgrid, 128, 0, stream>>>(smax, n_smax, n_queries, j, row_delta, thr);
    else if (n_smax <= 2048 && n_queries >= 256) sample_threshold_warp_kernel<64><<<wgrid, 128, 0, stream>>>(smax, n_smax, n_queries, j, row_delta, thr);
    else sample_threshold_kernel<<<n_queries, THR_THREADS, 0, stream>>>(smax, n_smax, n_queries, j, row_delta, thr);
    ANNCUR_LAUNCH_OK("sample_threshold_kernel");
    return ANNCUR_OK;
}

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

// rows = rows of the plane; row_stride > 1 makes a view of every row_stride-th row (the SAMPLE pass)
static int make_plane_map(CUtensorMap* map, const void* base, int64_t rows, int num_kb, int box_rows, bool bf16,
                          int row_stride = 1) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return ANNCUR_E_CUDA; }
    const int64_t view_rows = (rows + row_stride - 1) / row_stride;
    cuuint64_t dims[3] = {cuuint64_t(BLOCK_K), cuuint64_t(view_rows), cuuint64_t(num_kb)};
    cuuint64_t strides[2] = {cuuint64_t(BLOCK_K * 2) * cuuint64_t(row_stride), cuuint64_t(rows) * BLOCK_K * 2};
    cuuint32_t box[3] = {cuuint32_t(BLOCK_K), cuuint32_t(box_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld kb=%d)", int(r), (long long)rows, num_kb); return ANNCUR_E_CUDA; }
    return ANNCUR_OK;
}

// F32R carries one extra anchor slot (index k_dim) for the error-bound term
static int num_kb_for(int k_dim, int kind) { return (k_dim + (kind == ANNCUR_KIND_F32R ? 1 : 0) + BLOCK_K - 1) / BLOCK_K; }
static int planes_for(int kind) { return kind == ANNCUR_KIND_BF16 ? 1 : 2; }
static size_t plane_bytes(int64_t rows, int num_kb) { return align_up(size_t(num_kb) * size_t(rows) * BLOCK_K * 2, 256); }
// query planes: F32R keeps two more copies of the last k-block of the high plane (bound slot = -b and 0)
static int q_plane_kb(int num_kb, int kind) { return num_kb + (kind == ANNCUR_KIND_F32R ? 2 : 0); }
// Query planes are allocated in whole 256-row units (rows past n_queries zeroed): the TMA box of a query tile then never
// leaves the tensor.  Measured at N = 1M: with the box hanging out of a 1-row tensor MAIN took 0.267 ms, at 64 rows 0.172 ms.
// (256 = the two query tiles of a CTA pair: the second CTA of the last pair may own a tile past the batch)
static int q_plane_rows(int n_queries) { return (n_queries + 2 * BLOCK_M - 1) / (2 * BLOCK_M) * (2 * BLOCK_M); }
static int et_ld(int k_dim) { return (k_dim + 3) & ~3; }                 // row stride (floats) of the item-major fp32 copy
static bool valid_kind(int kind) { return kind == ANNCUR_KIND_F32X3 || kind == ANNCUR_KIND_BF16 || kind == ANNCUR_KIND_F32R; }

// The SAMPLE pass scores every G-th item.  For the common stride (16) the packed index keeps a CONTIGUOUS copy of those rows
// of the high plane (+1/16 of one plane): read through a strided view of the full plane, the sampled rows are 128-byte
// pieces 2 KB apart, which streams poorly from DRAM -- that matters for small batches, where SAMPLE is memory-bound
// (N = 1M, B = 64: 38 -> ~12 us of a 0.25 ms step).  Other strides keep using the strided view.
constexpr int SAMPLE_PLANE_STRIDE = 16;
static int64_t sample_plane_rows(int64_t n_items) { return (n_items + SAMPLE_PLANE_STRIDE - 1) / SAMPLE_PLANE_STRIDE; }

struct PackedLayout {               // byte offsets inside a packed item index
    int num_kb;
    size_t plane, off_scratch, off_rowmax, off_et, off_sample, total;
};
static PackedLayout packed_layout(int64_t n_items, int k_dim, int kind) {
    PackedLayout L{};
    L.num_kb = num_kb_for(k_dim, kind);
    L.plane = plane_bytes(n_items, L.num_kb);
    L.off_scratch = size_t(planes_for(kind)) * L.plane;
    L.off_rowmax = L.off_scratch + 256;
    L.off_et = L.off_rowmax + align_up(sizeof(float) * size_t(L.num_kb) * BLOCK_K, 256);
    L.off_sample = L.off_et + (kind == ANNCUR_KIND_F32R ? align_up(sizeof(float) * size_t(n_items) * et_ld(k_dim), 256) : 0);
    L.total = L.off_sample + plane_bytes(sample_plane_rows(n_items), L.num_kb);
    return L;
}

// sample[kb][t][:] = plane_h[kb][t * SAMPLE_PLANE_STRIDE][:]  (one 128-byte row per 8 threads, 16 bytes each)
__global__ void __launch_bounds__(256)
copy_sample_rows_kernel(const uint16_t* __restrict__ plane_h, int64_t n_items, int64_t s_rows, int num_kb, uint16_t* __restrict__ sample) {
    const int64_t total = int64_t(num_kb) * s_rows * 8;
    for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
        const int seg = int(t & 7);
        const int64_t row = (t >> 3) % s_rows;
        const int kb = int((t >> 3) / s_rows);
        const uint4 v = *reinterpret_cast<const uint4*>(plane_h + (int64_t(kb) * n_items + row * SAMPLE_PLANE_STRIDE) * BLOCK_K + seg * 8);
        *reinterpret_cast<uint4*>(sample + (int64_t(kb) * s_rows + row) * BLOCK_K + seg * 8) = v;
    }
}

static uint32_t pow2_at_least(uint32_t want) {
    uint32_t cap = 256;
    while (cap < want) cap <<= 1;
    return cap;
}

// Number of item chunks: pick the split of the item tiles that balances (chunks x query tiles)
// work items over the SMs; `warmup_tiles` is what a chunk pays before its threshold is useful
// (about two tiles when streaming from -inf, next to nothing with sampled thresholds).
static int choose_chunks(int m_tiles, int n_tiles, int sms, double warmup_tiles, int c_min = 1) {
    const int c_max = n_tiles < 256 ? n_tiles : 256;
    if (c_min > c_max) c_min = c_max;
    int best_c = c_min;
    double best_cost = 1e300;
    for (int c = c_min; c <= c_max; ++c) {
        const long long items = 1ll * m_tiles * c;
        const long long waves = (items + sms - 1) / sms;
        const int tiles_per = (n_tiles + c - 1) / c;
        const double cost = double(waves) * (tiles_per + warmup_tiles);
        if (cost < best_cost * 0.999) { best_cost = cost; best_c = c; }
    }
    return best_c;
}

// smallest j with P[Binomial(n, p) >= j] <= eps
int binomial_tail_rank(int n, double p, double eps) {
    if (n <= 0) return 1;
    // pmf by recurrence from the mode outwards is overkill here: n <= 1023, plain forward recurrence in log space
    std::vector<double> pmf(size_t(n) + 1);
    double logq = log1p(-p), logp = log(p);
    double lc = 0.0;                                           // log C(n, i)
    for (int i = 0; i <= n; ++i) {
        if (i > 0) lc += log(double(n - i + 1)) - log(double(i));
        pmf[size_t(i)] = exp(lc + i * logp + (n - i) * logq);
    }
    double tail = 0.0;
    int j = n + 1;
    for (int i = n; i >= 1; --i) {
        tail += pmf[size_t(i)];
        if (tail > eps) break;
        j = i;
    }
    return j < 1 ? 1 : j;
}

static int cta_group_for(int m_tiles);

static unsigned wait_sleep_ns() {
    static const unsigned v = [] { const char* e = getenv("ANNCUR_WAIT_SLEEP_NS"); return e ? unsigned(atoi(e)) : 0u; }();
    return v;
}
static unsigned long long wait_timeout_cycles() {
    static unsigned long long v = [] {
        const char* e = getenv("ANNCUR_WAIT_TIMEOUT_CYCLES");
        return e ? strtoull(e, nullptr, 10) : WAIT_TIMEOUT_CYCLES;
    }();
    return v;
}

struct FusedPlan {
    int num_kb, m_tiles, n_tiles, n_chunks;
    uint32_t cap;
    // SAMPLE pass (sample_stride == 0: not used, MAIN streams from -inf)
    int sample_stride, sample_rank, s_items, s_tiles, s_chunks, n_smax, smax_wide;
    size_t off_qplanes, off_inv_scale, off_delta, off_thr, off_flags, off_big, off_counts, off_cand, off_smax, off_err, total;
};

static FusedPlan make_plan(int n_queries, int64_t n_items, int k_dim, int k, int kind) {
    FusedPlan pl{};
    const int sms = sm_count();
    pl.num_kb = num_kb_for(k_dim, kind);
    pl.m_tiles = (n_queries + BLOCK_M - 1) / BLOCK_M;
    pl.n_tiles = int((n_items + BLOCK_N - 1) / BLOCK_N);
    const int cg = cta_group_for(pl.m_tiles);                 // work is scheduled over CTA pairs when cg == 2
    const int units = sms / cg > 0 ? sms / cg : 1;
    const int m_groups = pl.m_tiles > 0 ? (pl.m_tiles + cg - 1) / cg : 1;
    // sampling stride G: the coarsest of 16 / 8 / 4 that still leaves >= 4 j group maxima per row.  (Coarser strides
    // were tried for large item sets: G = 64 makes SAMPLE 4x cheaper but leaves ~800 instead of ~460 survivors per
    // row at k = 100, which costs more in the select than it saves -- measured slower at N = 1M for B = 64 and 4096.)
    // Second round, only when no stride gives 4 j maxima: settle for 2 j.  The j-th largest of n group maxima then sits
    // at up to the median of the maxima (~1.4 j G survivors per row instead of ~1.25 j G) -- still a one-pass threshold
    // where the call would otherwise stream from -inf: k = 1000 at N = 100k (the reference's largest k_r,
    // ..._w_fixed_train_test_splits.py:238-247) took 17 ms per 4096 queries unsampled.
    static const int forced_stride = [] { const char* e = getenv("ANNCUR_SAMPLE_STRIDE"); return e ? atoi(e) : 0; }();
    static const int forced_wide = [] { const char* e = getenv("ANNCUR_SMAX_WIDE"); return e ? atoi(e) : -1; }();
    for (int need : {4, 2}) {
        for (int G : {forced_stride > 0 ? forced_stride : 16, 16, 8, 4}) {
            const int j = binomial_tail_rank(k - 1, 1.0 / G, 1e-6);
            const int64_t s_items = (n_items + G - 1) / G;
            const int64_t n_smax = (s_items + 31) / 32;
            if (n_smax >= int64_t(need) * j && s_items >= 4 * BLOCK_N) {
                pl.sample_stride = G; pl.sample_rank = j; pl.s_items = int(s_items);
                pl.s_tiles = int((s_items + BLOCK_N - 1) / BLOCK_N);
                // Group maxima: one per 32 sampled items, or -- when that leaves far more maxima than the rank needs -- one
                // per 128 (tile, column half): the j-th largest of n maxima keeps the threshold equally tight as long as
                // n >> j, and the threshold kernel reads a quarter of the values (47 -> ~12 us at N = 1M, B = 4096).
                pl.smax_wide = (forced_wide >= 0 ? forced_wide != 0 : (n_smax / 4 >= 16ll * j)) ? 1 : 0;
                pl.n_smax = pl.smax_wide ? pl.s_tiles * EPI_HALVES : pl.s_tiles * (BLOCK_N / 32);
                pl.s_chunks = choose_chunks(m_groups, pl.s_tiles, units, 0.25);
                break;
            }
        }
        if (pl.sample_stride != 0) break;
    }
    const bool sampled = pl.sample_stride != 0;
    // with sampled thresholds a row keeps ~1.25 j G survivors: enough chunks that one list holds twice its share
    const int c_min = sampled ? int((2.5 * pl.sample_rank * pl.sample_stride) / (1984.0 * EPI_HALVES)) + 1 : 1;
    pl.n_chunks = choose_chunks(m_groups, pl.n_tiles > 0 ? pl.n_tiles : 1, units, sampled ? 0.5 : 2.0, c_min);
    // list capacity: room for 2k (streaming compaction keeps k) and for twice the expected survivors of a chunk
    uint32_t want = uint32_t(2 * k);
    if (sampled) {
        const double expect = 1.25 * pl.sample_rank * pl.sample_stride / double(pl.n_chunks * EPI_HALVES);
        const uint32_t w2 = uint32_t(2.0 * expect) + 64u;
        if (w2 > want) want = w2;
    }
    pl.cap = pow2_at_least(want);
    if (pl.cap > 2048u) pl.cap = 2048u;
    size_t off = 0;
    pl.off_qplanes = off; off += size_t(planes_for(kind)) * plane_bytes(q_plane_rows(n_queries), q_plane_kb(pl.num_kb, kind));
    pl.off_inv_scale = off; off += align_up(sizeof(float) * size_t(n_queries), 256);
    pl.off_delta = off; off += align_up(sizeof(float) * size_t(n_queries), 256);
    pl.off_thr = off; off += align_up(sizeof(uint32_t) * size_t(n_queries), 256);
    pl.off_flags = off; off += align_up(sizeof(uint32_t) * (size_t(pl.m_tiles > 0 ? pl.m_tiles : 1) + 1 + size_t(n_queries)), 256);
    pl.off_big = off; off += align_up(sizeof(uint32_t) * (1 + size_t(n_queries)), 256);
    pl.off_counts = off; off += align_up(sizeof(uint32_t) * size_t(n_queries) * pl.n_chunks * EPI_HALVES, 256);
    pl.off_cand = off; off += align_up(sizeof(uint64_t) * size_t(n_queries) * pl.n_chunks * EPI_HALVES * pl.cap, 256);
    pl.off_smax = off; off += align_up(sizeof(float) * size_t(n_queries) * size_t(pl.n_smax > 0 ? pl.n_smax : 1), 256);
    pl.off_err = off; off += 256;
    pl.total = off;
    return pl;
}

size_t packed_items_bytes(int64_t n_items, int k_dim, int kind) {
    if (n_items <= 0 || k_dim <= 0 || !valid_kind(kind)) return 256;
    // planes | 256 B max-abs scratch | per-anchor-dimension max |E| (num_kb * 32 floats) | F32R: E^T (N x et_ld fp32) | sample rows
    return packed_layout(n_items, k_dim, kind).total;
}

int pack_items(const float* E, int64_t lde, int64_t n_items, int k_dim, int kind, void* packed, float* e_scale_out,
               cudaStream_t stream) {
    if (!valid_kind(kind)) { set_error("pack_items: unknown kind %d", kind); return ANNCUR_E_INVALID; }
    const bool bf16 = kind == ANNCUR_KIND_BF16;
    if (n_items <= 0 || k_dim <= 0) {
        finish_scale_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<uint32_t*>(packed), e_scale_out, true);
        ANNCUR_LAUNCH_OK("finish_scale_kernel");
        return ANNCUR_OK;
    }
    if (n_items >= (int64_t(1) << 31) - BLOCK_N) { set_error("pack_items: n_items %lld too large for one shard", (long long)n_items); return ANNCUR_E_UNSUPPORTED; }
    if (kind == ANNCUR_KIND_F32R && k_dim > ANNCUR_MAX_K_DIM_F32R) { set_error("pack_items: k_dim %d > %d for kind F32R", k_dim, ANNCUR_MAX_K_DIM_F32R); return ANNCUR_E_UNSUPPORTED; }
    const PackedLayout L = packed_layout(n_items, k_dim, kind);
    char* base = reinterpret_cast<char*>(packed);
    uint16_t* plane_h = reinterpret_cast<uint16_t*>(base);
    uint16_t* plane_l = bf16 ? nullptr : reinterpret_cast<uint16_t*>(base + L.plane);
    uint32_t* maxabs = reinterpret_cast<uint32_t*>(base + L.off_scratch);
    ANNCUR_CUDA_OK(cudaMemsetAsync(maxabs, 0, 4, stream));
    if (!bf16) {
        absmax_kernel<<<sm_count() * 8, 256, 0, stream>>>(E, lde, k_dim, n_items, maxabs);
        ANNCUR_LAUNCH_OK("absmax_kernel");
    }
    finish_scale_kernel<<<1, 1, 0, stream>>>(maxabs, e_scale_out, bf16);
    ANNCUR_LAUNCH_OK("finish_scale_kernel");
    float* rowmax = reinterpret_cast<float*>(base + L.off_rowmax);
    row_absmax_kernel<<<L.num_kb * BLOCK_K, 256, 0, stream>>>(E, lde, n_items, k_dim, e_scale_out, rowmax);
    ANNCUR_LAUNCH_OK("row_absmax_kernel");
    dim3 grid(unsigned((n_items + 63) / 64), unsigned(L.num_kb * K32_PER_KB));
    if (bf16) pack_items_kernel<true><<<grid, 256, 0, stream>>>(E, lde, n_items, k_dim, e_scale_out, plane_h, plane_l);
    else pack_items_kernel<false><<<grid, 256, 0, stream>>>(E, lde, n_items, k_dim, e_scale_out, plane_h, plane_l);
    ANNCUR_LAUNCH_OK("pack_items_kernel");
    if (kind == ANNCUR_KIND_F32R) {
        const int64_t blocks = (n_items + 255) / 256;
        item_bound_slot_kernel<<<unsigned(blocks < 16 * sm_count() ? blocks : 16 * sm_count()), 256, 0, stream>>>(E, lde, n_items, k_dim, e_scale_out, plane_h);
        ANNCUR_LAUNCH_OK("item_bound_slot_kernel");
        const int ld = et_ld(k_dim);
        dim3 tgrid(unsigned((n_items + 31) / 32), unsigned((ld + 31) / 32));
        transpose_items_kernel<<<tgrid, 256, 0, stream>>>(E, lde, n_items, k_dim, ld, reinterpret_cast<float*>(base + L.off_et));
        ANNCUR_LAUNCH_OK("transpose_items_kernel");
    }
    {   // contiguous copy of every 16th row of the finished high plane (bound slot included) for the SAMPLE pass
        const int64_t s_rows = sample_plane_rows(n_items);
        const int64_t work = int64_t(L.num_kb) * s_rows * 8;
        const int64_t blocks = (work + 255) / 256;
        copy_sample_rows_kernel<<<unsigned(blocks < 16 * sm_count() ? blocks : 16 * sm_count()), 256, 0, stream>>>(
            plane_h, n_items, s_rows, L.num_kb, reinterpret_cast<uint16_t*>(base + L.off_sample));
        ANNCUR_LAUNCH_OK("copy_sample_rows_kernel");
    }
    return ANNCUR_OK;
}

size_t score_topk_workspace_bytes(int n_queries, int64_t n_items, int k_dim, int k, int kind) {
    if (n_queries <= 0 || n_items <= 0 || k_dim <= 0 || k <= 0) return 256;
    return make_plan(n_queries, n_items, k_dim, k, kind).total;
}

// rows of the last call on `workspace` that the sampled path had to hand to the REDO pass (blocks on `stream`)
int score_topk_redo_rows(const void* workspace, int n_queries, int64_t n_items, int k_dim, int k, int kind, int* redo_rows_host,
                         cudaStream_t stream) {
    *redo_rows_host = 0;
    if (n_queries <= 0 || n_items <= 0 || k_dim <= 0 || k <= 0 || !valid_kind(kind)) return ANNCUR_OK;
    const FusedPlan pl = make_plan(n_queries, n_items, k_dim, k, kind);
    if (pl.sample_stride == 0) return ANNCUR_OK;                 // unsampled calls stream from -inf, there is no REDO
    const uint32_t* flags = reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(workspace) + pl.off_flags);
    uint32_t n = 0;
    ANNCUR_CUDA_OK(cudaMemcpyAsync(&n, flags + pl.m_tiles, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    ANNCUR_CUDA_OK(cudaStreamSynchronize(stream));
    *redo_rows_host = int(n);
    return ANNCUR_OK;
}

// ---- optional event timing of the fused kernel (anncur_profile_*) ---------------------------------
struct ProfileState {
    bool on = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
};
static thread_local ProfileState g_prof;

int profile_enable(int on) {
    g_prof.on = on != 0;
    if (!g_prof.on) {
        for (auto& e : g_prof.events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
        g_prof.events.clear();
    }
    return ANNCUR_OK;
}

int profile_read(double* ms_sum, int* launches) {
    double sum = 0.0;
    int n = 0;
    for (auto& e : g_prof.events) {
        ANNCUR_CUDA_OK(cudaEventSynchronize(e.second));
        float ms = 0.f;
        ANNCUR_CUDA_OK(cudaEventElapsedTime(&ms, e.first, e.second));
        sum += ms;
        ++n;
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    g_prof.events.clear();
    if (ms_sum) *ms_sum = sum;
    if (launches) *launches = n;
    return ANNCUR_OK;
}

template <int PASSES, bool BF16, int CPL, int CG, int EPI = EPI_TOPK>
static int launch_fused(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0, const CUtensorMap& b1,
                        const FusedParams& fp, bool timed, cudaStream_t stream) {
    using Cfg = StageCfg<PASSES, CG>;
    const int smem = Cfg::kStages * Cfg::kStageBytes + 1024 /*align slack*/ + 8 * (2 * Cfg::kStages + 4) + 16 + NUM_EPI_WARPS * 256 * 4 +
                     (EPI == EPI_DENSE ? NUM_EPI_WARPS * DENSE_TBUF_FLOATS * 4 : 0);
    auto kernel = fused_score_topk_kernel<PASSES, BF16, CPL, CG, EPI>;
    {   // once per instantiation and device (the value is a compile-time constant): a call issues three of these launches,
        // and at small batches the host side of a call is what the GPU waits for between its short kernels
        static bool attr_set[64] = {};
        int dev = 0;
        ANNCUR_CUDA_OK(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            ANNCUR_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
    }
    // persistent: one CTA (pair) per SM (pair), never more CTAs than work items
    const long long units = 1ll * ((fp.m_tiles + CG - 1) / CG) * fp.n_chunks;
    const long long max_units = sm_count() / CG;
    const int grid = int(units < max_units ? units : max_units) * CG;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    timed = timed && g_prof.on;
    if (timed) {
        ANNCUR_CUDA_OK(cudaEventCreate(&ev0));
        ANNCUR_CUDA_OK(cudaEventCreate(&ev1));
        ANNCUR_CUDA_OK(cudaEventRecord(ev0, stream));
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(grid));
    cfg.blockDim = dim3(FUSED_THREADS);
    cfg.dynamicSmemBytes = size_t(smem);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ANNCUR_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, a0, a1, b0, b1, fp));
    ANNCUR_LAUNCH_OK("fused_score_topk_kernel");
    if (timed) {
        ANNCUR_CUDA_OK(cudaEventRecord(ev1, stream));
        g_prof.events.emplace_back(ev0, ev1);
    }
    return ANNCUR_OK;
}

template <int PASSES, bool BF16, int CG>
static int dispatch_cap(uint32_t cap, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0,
                        const CUtensorMap& b1, const FusedParams& fp, bool timed, cudaStream_t stream) {
    switch (cap) {
        case 256: return launch_fused<PASSES, BF16, 8, CG>(a0, a1, b0, b1, fp, timed, stream);
        case 512: return launch_fused<PASSES, BF16, 16, CG>(a0, a1, b0, b1, fp, timed, stream);
        case 1024: return launch_fused<PASSES, BF16, 32, CG>(a0, a1, b0, b1, fp, timed, stream);
        case 2048: return launch_fused<PASSES, BF16, 64, CG>(a0, a1, b0, b1, fp, timed, stream);
    }
    set_error("score_topk: no kernel for candidate capacity %u", cap);
    return ANNCUR_E_UNSUPPORTED;
}

// CTAs per MMA: pairs (cta_group::2) whenever there are at least two query tiles to pair up.
// ANNCUR_CTA_GROUP=1|2 overrides (testing).
static int cta_group_for(int m_tiles) {
    static int forced = [] {
        const char* e = getenv("ANNCUR_CTA_GROUP");
        return e ? atoi(e) : 0;
    }();
    if (forced == 1 || forced == 2) return forced;
    return m_tiles >= 2 ? 2 : 1;
}

int score_topk_fused(const float* Q, int ldq, int n_queries, const void* packed_items, const float* e_scale,
                     int64_t n_items, int k_dim, int kind, int k, int64_t idx_offset, float* out_vals,
                     int64_t* out_idx, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!valid_kind(kind)) { set_error("score_topk: unknown kind %d", kind); return ANNCUR_E_INVALID; }
    if (k < 1 || k > ANNCUR_MAX_K_FUSED) { set_error("score_topk: k = %d outside [1, %d]", k, ANNCUR_MAX_K_FUSED); return ANNCUR_E_INVALID; }
    if (n_queries <= 0) return ANNCUR_OK;
    if (n_items <= 0 || k_dim <= 0) {
        // empty anchor set (k_i = 0 is in the reference's grid): all approximate scores are 0
        fill_zero_scores_kernel<<<sm_count(), 256, 0, stream>>>(n_queries, k, n_items, idx_offset, out_vals, out_idx);
        ANNCUR_LAUNCH_OK("fill_zero_scores_kernel");
        return ANNCUR_OK;
    }
    if (n_items >= (int64_t(1) << 31) - BLOCK_N) { set_error("score_topk: n_items %lld too large for one shard", (long long)n_items); return ANNCUR_E_UNSUPPORTED; }
    const FusedPlan pl = make_plan(n_queries, n_items, k_dim, k, kind);
    if (workspace_bytes < pl.total) { set_error("score_topk workspace too small: %zu < %zu", workspace_bytes, pl.total); return ANNCUR_E_WORKSPACE; }
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) || (reinterpret_cast<uintptr_t>(packed_items) & 255)) {
        set_error("score_topk: workspace and packed_items must be 256-byte aligned");
        return ANNCUR_E_INVALID;
    }
    const bool bf16 = kind == ANNCUR_KIND_BF16;
    const bool refine = kind == ANNCUR_KIND_F32R;
    const bool sampled = pl.sample_stride != 0;
    char* ws = reinterpret_cast<char*>(workspace);
    const int q_kb = q_plane_kb(pl.num_kb, kind);
    const int q_rows = q_plane_rows(n_queries);
    const size_t qpb = plane_bytes(q_rows, q_kb);
    uint16_t* q_h = reinterpret_cast<uint16_t*>(ws + pl.off_qplanes);
    uint16_t* q_l = bf16 ? nullptr : reinterpret_cast<uint16_t*>(ws + pl.off_qplanes + qpb);
    float* inv_scale = reinterpret_cast<float*>(ws + pl.off_inv_scale);
    float* delta = reinterpret_cast<float*>(ws + pl.off_delta);
    uint32_t* thr = reinterpret_cast<uint32_t*>(ws + pl.off_thr);
    uint32_t* flags = reinterpret_cast<uint32_t*>(ws + pl.off_flags);
    uint32_t* big_rows = reinterpret_cast<uint32_t*>(ws + pl.off_big);
    uint32_t* counts = reinterpret_cast<uint32_t*>(ws + pl.off_counts);
    uint64_t* cand = reinterpret_cast<uint64_t*>(ws + pl.off_cand);
    float* smax = reinterpret_cast<float*>(ws + pl.off_smax);
    int* err = reinterpret_cast<int*>(ws + pl.off_err);

    const int qgrid = (q_rows + 7) / 8;                    // the kernel also zeroes the padding rows of the last query tile
    const PackedLayout L = packed_layout(n_items, k_dim, kind);
    const char* items = reinterpret_cast<const char*>(packed_items);
    const float* e_rowmax = reinterpret_cast<const float*>(items + L.off_rowmax);
    if (bf16) pack_queries_kernel<ANNCUR_KIND_BF16><<<qgrid, 256, 0, stream>>>(Q, ldq, n_queries, q_rows, k_dim, pl.num_kb, e_scale, q_h, q_l, inv_scale, thr, flags, e_rowmax, delta);
    else if (refine) pack_queries_kernel<ANNCUR_KIND_F32R><<<qgrid, 256, 0, stream>>>(Q, ldq, n_queries, q_rows, k_dim, pl.num_kb, e_scale, q_h, q_l, inv_scale, thr, flags, e_rowmax, delta);
    else pack_queries_kernel<ANNCUR_KIND_F32X3><<<qgrid, 256, 0, stream>>>(Q, ldq, n_queries, q_rows, k_dim, pl.num_kb, e_scale, q_h, q_l, inv_scale, thr, flags, e_rowmax, delta);
    ANNCUR_LAUNCH_OK("pack_queries_kernel");

    const int cg = cta_group_for(pl.m_tiles);
    const int b_box = BLOCK_N / cg;                       // a CTA of a pair stages half of each item tile
    CUtensorMap a0, a1, b0, b1;
    int rc;
    if ((rc = make_plane_map(&a0, q_h, q_rows, q_kb, BLOCK_M, bf16)) != ANNCUR_OK) return rc;
    if ((rc = make_plane_map(&b0, items, n_items, pl.num_kb, b_box, bf16)) != ANNCUR_OK) return rc;
    if (bf16) { a1 = a0; b1 = b0; }
    else {
        if ((rc = make_plane_map(&a1, q_l, q_rows, q_kb, BLOCK_M, false)) != ANNCUR_OK) return rc;
        if ((rc = make_plane_map(&b1, items + L.plane, n_items, pl.num_kb, b_box, false)) != ANNCUR_OK) return rc;
    }
    // the full-precision launch: 3 passes (fp32-grade kinds) or the single bf16 pass
    auto launch_full = [&](const FusedParams& fp, bool timed) {
        if (cg == 2)
            return bf16 ? dispatch_cap<1, true, 2>(pl.cap, a0, a1, b0, b1, fp, timed, stream)
                        : dispatch_cap<3, false, 2>(pl.cap, a0, a1, b0, b1, fp, timed, stream);
        return bf16 ? dispatch_cap<1, true, 1>(pl.cap, a0, a1, b0, b1, fp, timed, stream)
                    : dispatch_cap<3, false, 1>(pl.cap, a0, a1, b0, b1, fp, timed, stream);
    };
    FusedParams fp{};
    fp.n_queries = n_queries; fp.num_kb = pl.num_kb; fp.k = k; fp.m_tiles = pl.m_tiles;
    fp.cand = cand; fp.counts = counts; fp.thr_shared = thr; fp.error_flag = err; fp.smax = smax; fp.n_smax = pl.n_smax;
    fp.wait_timeout = wait_timeout_cycles();
    fp.wait_sleep_ns = wait_sleep_ns();
    // query high-plane variant of the last k-block (F32R only): +b at num_kb - 1, -b at num_kb, plain scores at num_kb + 1
    const int akb_upper = pl.num_kb - 1, akb_lower = refine ? pl.num_kb : pl.num_kb - 1, akb_plain = refine ? pl.num_kb + 1 : pl.num_kb - 1;
    fp.a_last_kb = akb_plain;

    if (sampled) {
        // SAMPLE: every G-th item through a strided view of the same planes, ONE tensor pass on the high halves
        // -> 32-column group maxima -> thresholds.  F32X3: plain one-pass scores, the threshold is lowered by the row's
        // statistical error bound; F32R: lower bounds A - b, the threshold needs no margin.
        CUtensorMap s0;
        static const bool strided_only = getenv("ANNCUR_SAMPLE_STRIDED_VIEW") != nullptr;          // testing: never use the contiguous copy
        if (pl.sample_stride == SAMPLE_PLANE_STRIDE && !strided_only)
            rc = make_plane_map(&s0, items + L.off_sample, sample_plane_rows(n_items), pl.num_kb, b_box, bf16);
        else
            rc = make_plane_map(&s0, items, n_items, pl.num_kb, b_box, bf16, pl.sample_stride);
        if (rc != ANNCUR_OK) return rc;
        FusedParams sp = fp;
        sp.mode = MODE_SAMPLE; sp.n_items = pl.s_items; sp.n_tiles = pl.s_tiles; sp.n_chunks = pl.s_chunks;
        sp.smax_wide = pl.smax_wide;
        sp.a_last_kb = akb_lower;
        if (cg == 2) rc = bf16 ? launch_fused<1, true, 8, 2>(a0, a0, s0, s0, sp, false, stream)
                               : launch_fused<1, false, 8, 2>(a0, a0, s0, s0, sp, false, stream);
        else rc = bf16 ? launch_fused<1, true, 8, 1>(a0, a0, s0, s0, sp, false, stream)
                       : launch_fused<1, false, 8, 1>(a0, a0, s0, s0, sp, false, stream);
        if (rc != ANNCUR_OK) return rc;
        if ((rc = launch_sample_threshold(smax, pl.n_smax, n_queries, pl.sample_rank, delta, thr, stream)) != ANNCUR_OK) return rc;
    }
    // MAIN
    fp.mode = MODE_MAIN; fp.n_items = int(n_items); fp.n_tiles = pl.n_tiles; fp.n_chunks = pl.n_chunks;
    fp.close_compact = sampled ? 0 : 1;
    // F32R keeps a row's candidates in shared memory while it re-scores them: 1024 (k up to ~200), 2048 or 4096 per row; past
    // that (cannot happen for k <= 1024) the call is served by the 3-pass launch
    const double expect_row = sampled ? 1.25 * pl.sample_rank * pl.sample_stride : 0.0;
    const int row_cap = 1.8 * expect_row <= 1024.0 ? 1024 : 1.8 * expect_row <= 2048.0 ? 2048 : 4096;
    // Re-scoring costs ~k x 2 KB of gathered reads per row (measured 1.9 us of kernel time per unit of k at B = 4096), the
    // two extra tensor passes ~4.6 ns per item: beyond k ~ N / 400 the 3-pass launch is the faster way to the same answer.
    // (ANNCUR_F32R_K_RATIO overrides the 400 for experiments.  Measured with the 2048 / 4096-candidate refine variants at
    // N = 100k, B = 4096: ratio 250 serves k = 375 by filter + refine in 0.97 instead of 1.02 ms, ratio 190 serves k = 500 in 1.10
    // instead of 1.01 ms -- the crossover stays where it was.)
    static const int64_t k_ratio = [] { const char* e = getenv("ANNCUR_F32R_K_RATIO"); return e ? int64_t(atoi(e)) : int64_t(400); }();
    if (refine && sampled && 1.5 * expect_row <= 4096.0 && int64_t(k) * k_ratio <= n_items) {
        // one f16 pass of upper bounds, then the exact re-scoring of the candidates (refine_topk.cu)
        fp.a_last_kb = akb_upper; fp.filter = 1;
        rc = cg == 2 ? dispatch_cap<1, false, 2>(pl.cap, a0, a0, b0, b0, fp, true, stream)
                     : dispatch_cap<1, false, 1>(pl.cap, a0, a0, b0, b0, fp, true, stream);
        if (rc != ANNCUR_OK) return rc;
        rc = refine_topk_keylists(cand, counts, pl.n_chunks * EPI_HALVES, int(pl.cap), n_queries, k, idx_offset, Q, ldq, k_dim,
                                  reinterpret_cast<const float*>(items + L.off_et), et_ld(k_dim), inv_scale, out_vals, out_idx,
                                  thr, flags, pl.m_tiles, n_items, row_cap, stream);
        if (rc != ANNCUR_OK) return rc;
        fp.a_last_kb = akb_plain; fp.filter = 0;
    } else {
        if ((rc = launch_full(fp, true)) != ANNCUR_OK) return rc;
        rc = select_topk_keylists(cand, counts, pl.n_chunks * EPI_HALVES, int(pl.cap), n_queries, k, idx_offset, inv_scale, out_vals,
                                  out_idx, sampled ? 1 : 0, thr, flags, pl.m_tiles, n_items, big_rows, stream);
        if (rc != ANNCUR_OK || !sampled) return rc;
    }
    // REDO: rows that came up short (F32R: or filled a list, or failed the certificate) restart from -inf in streaming
    // mode with the full-precision launch; unflagged query tiles are skipped
    fp.mtile_flags = flags; fp.close_compact = 1;
    if ((rc = launch_full(fp, false)) != ANNCUR_OK) return rc;
    return select_topk_keylists(cand, counts, pl.n_chunks * EPI_HALVES, int(pl.cap), n_queries, k, idx_offset, inv_scale, out_vals,
                                out_idx, 2, thr, flags, pl.m_tiles, n_items, big_rows, stream);
}

// ---- dense products on the same pipeline -----------------------------------------------------------------------------
// out = Q . E (EPI_DENSE) or per-row sum_j (Q . E - A)^2, sum_j A^2 (EPI_ERR) with the 3-pass fp32-grade arithmetic of
// kind F32X3 (also on the planes of an F32R index).  Replaces the FFMA GEMM for the dense getters
// (eval/matrix_approx_zeshel.py:71-119), the item-embedding build U @ R (:65; pack R, pass U as the queries) and the
// Frobenius errors (eval/run_retrieval_eval_wrt_exact_crossenc.py:146-147).
struct DensePlan {
    int num_kb, m_tiles, n_tiles, n_chunks;
    size_t off_qplanes, off_inv_scale, off_delta, off_thr, off_flags, off_err, total;
};

static DensePlan make_dense_plan(int n_queries, int64_t n_items, int k_dim, int kind) {
    DensePlan pl{};
    pl.num_kb = num_kb_for(k_dim, kind);
    pl.m_tiles = (n_queries + BLOCK_M - 1) / BLOCK_M;
    pl.n_tiles = int((n_items + BLOCK_N - 1) / BLOCK_N);
    const int cg = cta_group_for(pl.m_tiles);
    const int units = sm_count() / cg > 0 ? sm_count() / cg : 1;
    pl.n_chunks = choose_chunks((pl.m_tiles + cg - 1) / cg, pl.n_tiles > 0 ? pl.n_tiles : 1, units, 0.0);
    size_t off = 0;
    pl.off_qplanes = off; off += 2 * plane_bytes(q_plane_rows(n_queries), q_plane_kb(pl.num_kb, kind));
    pl.off_inv_scale = off; off += align_up(sizeof(float) * size_t(n_queries), 256);
    pl.off_delta = off; off += align_up(sizeof(float) * size_t(n_queries), 256);
    pl.off_thr = off; off += align_up(sizeof(uint32_t) * size_t(n_queries), 256);
    pl.off_flags = off; off += align_up(sizeof(uint32_t) * (size_t(pl.m_tiles) + 2), 256);
    pl.off_err = off; off += 256;
    pl.total = off;
    return pl;
}

size_t score_dense_workspace_bytes(int n_queries, int64_t n_items, int k_dim, int kind) {
    if (n_queries <= 0 || n_items <= 0 || k_dim <= 0) return 256;
    return make_dense_plan(n_queries, n_items, k_dim, kind).total;
}

static int run_dense(int epi, int bound_sign, const float* Q, int ldq, int n_queries, const void* packed_items, const float* e_scale,
                     int64_t n_items, int k_dim, int kind, float* out, int64_t ldo, const float* exact, int64_t lda,
                     double* err2, double* norm2, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (bound_sign != 0 && kind != ANNCUR_KIND_F32R) { set_error("score bounds exist for kind F32R only"); return ANNCUR_E_INVALID; }
    if (kind != ANNCUR_KIND_F32X3 && kind != ANNCUR_KIND_F32R) { set_error("score_dense: kind %d has no fp32-grade planes", kind); return ANNCUR_E_INVALID; }
    if (n_queries <= 0 || n_items <= 0) return ANNCUR_OK;
    if (k_dim <= 0) { set_error("score_dense: k_dim must be positive"); return ANNCUR_E_INVALID; }
    if (n_items >= (int64_t(1) << 31) - BLOCK_N) { set_error("score_dense: n_items %lld too large", (long long)n_items); return ANNCUR_E_UNSUPPORTED; }
    const DensePlan pl = make_dense_plan(n_queries, n_items, k_dim, kind);
    if (workspace_bytes < pl.total) { set_error("score_dense workspace too small: %zu < %zu", workspace_bytes, pl.total); return ANNCUR_E_WORKSPACE; }
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) || (reinterpret_cast<uintptr_t>(packed_items) & 255)) {
        set_error("score_dense: workspace and packed_items must be 256-byte aligned");
        return ANNCUR_E_INVALID;
    }
    char* ws = reinterpret_cast<char*>(workspace);
    const int q_kb = q_plane_kb(pl.num_kb, kind);
    const int q_rows = q_plane_rows(n_queries);
    const size_t qpb = plane_bytes(q_rows, q_kb);
    uint16_t* q_h = reinterpret_cast<uint16_t*>(ws + pl.off_qplanes);
    uint16_t* q_l = reinterpret_cast<uint16_t*>(ws + pl.off_qplanes + qpb);
    float* inv_scale = reinterpret_cast<float*>(ws + pl.off_inv_scale);
    float* delta = reinterpret_cast<float*>(ws + pl.off_delta);
    uint32_t* thr = reinterpret_cast<uint32_t*>(ws + pl.off_thr);
    uint32_t* flags = reinterpret_cast<uint32_t*>(ws + pl.off_flags);
    int* err = reinterpret_cast<int*>(ws + pl.off_err);
    const PackedLayout L = packed_layout(n_items, k_dim, kind);
    const char* items = reinterpret_cast<const char*>(packed_items);
    const float* e_rowmax = reinterpret_cast<const float*>(items + L.off_rowmax);
    const int qgrid = (q_rows + 7) / 8;                    // the kernel also zeroes the padding rows of the last query tile
    if (kind == ANNCUR_KIND_F32R) pack_queries_kernel<ANNCUR_KIND_F32R><<<qgrid, 256, 0, stream>>>(Q, ldq, n_queries, q_rows, k_dim, pl.num_kb, e_scale, q_h, q_l, inv_scale, thr, flags, e_rowmax, delta);
    else pack_queries_kernel<ANNCUR_KIND_F32X3><<<qgrid, 256, 0, stream>>>(Q, ldq, n_queries, q_rows, k_dim, pl.num_kb, e_scale, q_h, q_l, inv_scale, thr, flags, e_rowmax, delta);
    ANNCUR_LAUNCH_OK("pack_queries_kernel");
    if (epi == EPI_ERR) {
        ANNCUR_CUDA_OK(cudaMemsetAsync(err2, 0, sizeof(double) * size_t(n_queries), stream));
        ANNCUR_CUDA_OK(cudaMemsetAsync(norm2, 0, sizeof(double) * size_t(n_queries), stream));
    }
    const int cg = cta_group_for(pl.m_tiles);
    const int b_box = BLOCK_N / cg;
    CUtensorMap a0, a1, b0, b1;
    int rc;
    if ((rc = make_plane_map(&a0, q_h, q_rows, q_kb, BLOCK_M, false)) != ANNCUR_OK) return rc;
    if ((rc = make_plane_map(&a1, q_l, q_rows, q_kb, BLOCK_M, false)) != ANNCUR_OK) return rc;
    if ((rc = make_plane_map(&b0, items, n_items, pl.num_kb, b_box, false)) != ANNCUR_OK) return rc;
    if ((rc = make_plane_map(&b1, items + L.plane, n_items, pl.num_kb, b_box, false)) != ANNCUR_OK) return rc;
    FusedParams fp{};
    fp.mode = MODE_MAIN; fp.n_queries = n_queries; fp.n_items = int(n_items); fp.num_kb = pl.num_kb; fp.k = 1;
    fp.m_tiles = pl.m_tiles; fp.n_tiles = pl.n_tiles; fp.n_chunks = pl.n_chunks; fp.error_flag = err;
    fp.wait_timeout = wait_timeout_cycles();
    fp.wait_sleep_ns = wait_sleep_ns();
    fp.a_last_kb = kind == ANNCUR_KIND_F32R ? pl.num_kb + 1 : pl.num_kb - 1;     // plain scores: bound slot = 0
    fp.row_inv_scale = inv_scale; fp.dense_out = out; fp.ldo = ldo; fp.exact = exact; fp.lda = lda; fp.err2 = err2; fp.norm2 = norm2;
    fp.dense_vec_ok = (out != nullptr && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (ldo & 3) == 0) ? 1 : 0;
    fp.smax = reinterpret_cast<float*>(ws); fp.cand = reinterpret_cast<uint64_t*>(ws); fp.counts = thr; fp.thr_shared = thr;   // unused by these epilogues
    if (bound_sign != 0) {
        // what SAMPLE (-) and MAIN (+) of kind F32R see: the one-pass score with the error-bound slot switched on
        fp.a_last_kb = bound_sign > 0 ? pl.num_kb - 1 : pl.num_kb;
        return cg == 2 ? launch_fused<1, false, 8, 2, EPI_DENSE>(a0, a0, b0, b0, fp, true, stream)
                       : launch_fused<1, false, 8, 1, EPI_DENSE>(a0, a0, b0, b0, fp, true, stream);
    }
    if (epi == EPI_DENSE)
        return cg == 2 ? launch_fused<3, false, 8, 2, EPI_DENSE>(a0, a1, b0, b1, fp, true, stream)
                       : launch_fused<3, false, 8, 1, EPI_DENSE>(a0, a1, b0, b1, fp, true, stream);
    return cg == 2 ? launch_fused<3, false, 8, 2, EPI_ERR>(a0, a1, b0, b1, fp, true, stream)
                   : launch_fused<3, false, 8, 1, EPI_ERR>(a0, a1, b0, b1, fp, true, stream);
}

int score_dense(const float* Q, int ldq, int n_queries, const void* packed_items, const float* e_scale, int64_t n_items,
                int k_dim, int kind, float* out, int64_t ldo, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    return run_dense(EPI_DENSE, 0, Q, ldq, n_queries, packed_items, e_scale, n_items, k_dim, kind, out, ldo, nullptr, 0, nullptr,
                     nullptr, workspace, workspace_bytes, stream);
}

int score_bounds_dense(const float* Q, int ldq, int n_queries, const void* packed_items, const float* e_scale, int64_t n_items,
                       int k_dim, int sign, float* out, int64_t ldo, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    return run_dense(EPI_DENSE, sign >= 0 ? 1 : -1, Q, ldq, n_queries, packed_items, e_scale, n_items, k_dim, ANNCUR_KIND_F32R, out,
                     ldo, nullptr, 0, nullptr, nullptr, workspace, workspace_bytes, stream);
}

int recon_error_packed(const float* Q, int ldq, int n_queries, const void* packed_items, const float* e_scale, int64_t n_items,
                       int k_dim, int kind, const float* A, int64_t lda, double* out_err2, double* out_norm2, void* workspace,
                       size_t workspace_bytes, cudaStream_t stream) {
    return run_dense(EPI_ERR, 0, Q, ldq, n_queries, packed_items, e_scale, n_items, k_dim, kind, nullptr, 0, A, lda, out_err2,
                     out_norm2, workspace, workspace_bytes, stream);
}

}  // namespace anncur
