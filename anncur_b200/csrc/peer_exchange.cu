// Candidate exchange of the item-sharded search over NVLink peer memory (not in the reference).
//
// Rank p of P scores the whole query batch against ITS slice of the items and holds a local top-k per row.  Query rows
// are owned in contiguous blocks (owner o merges rows [lo_o, hi_o)).  Instead of packing keys, calling a collective and
// unpacking, ONE kernel turns this rank's (value, index) lists into 64-bit keys and stores each row straight into its
// owner's receive buffer through the peer mapping (NVLink / NVSwitch stores), then publishes an epoch flag in every
// owner; the owner's merge waits for the P flags and selects the best k of the P lists of each of its rows.  No NCCL
// kernel, no staging copy, and only rows_owned x k x P keys arrive per rank instead of B x k x P.
//
// Layout of one channel's buffer (identical on every rank; allocated by anncur_peer_alloc, mapped into the peers with
// CUDA IPC):   keys[2][P][rows_cap][k_cap]  |  flags[P]  |  done counter, error flag, certificate-failure counter
//   keys[e & 1][s][r][:]  = sender s's list for this rank's owned row r at epoch e
//   flags[s]              = last epoch whose lists sender s has completely stored here
// Two key buffers alternate by epoch parity.  Calls on one channel must be stream-ordered on every rank (scatter(e),
// wait(e), merge(e), scatter(e+1), ...): a sender can only reach scatter(e+2) -- which reuses buffer e & 1 -- after its
// own wait(e+1) saw every rank's flag e+1, and a rank publishes flag e+1 only after its merge(e) has finished.
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace anncur {

constexpr int kMaxPeers = 16;
constexpr unsigned long long PEER_WAIT_TIMEOUT_CYCLES = 20000000000ull;   // ~10 s: give up (error flag) instead of hanging

struct PeerPtrs {
    uint64_t* keys[kMaxPeers];      // peer p's key area (base of keys[2][P][rows_cap][k_cap]) as mapped into this process
    uint32_t* flags[kMaxPeers];     // peer p's flags[P]
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// one warp per row: k (value, index) pairs -> keys, stored to the row owner's buffer; the last CTA publishes the flags
__global__ void __launch_bounds__(256)
scatter_keys_signal_kernel(const float* __restrict__ vals, const int64_t* __restrict__ idx, int n_rows, int k, int rank, int world,
                           int rows_cap, int k_cap, uint32_t epoch, PeerPtrs peers, uint32_t* __restrict__ done_counter) {
    const int warps_per_cta = blockDim.x >> 5;
    const int lane = int(lane_id());
    const size_t buf_stride = size_t(world) * rows_cap * k_cap;            // one epoch-parity buffer
    for (int row = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); row < n_rows; row += gridDim.x * warps_per_cta) {
        // owner of `row` under shard_bounds(n_rows, world): the largest o with (o * n_rows) / world <= row
        int owner = int((int64_t(row + 1) * world - 1) / n_rows);
        while (owner + 1 < world && (int64_t(owner + 1) * n_rows) / world <= row) ++owner;
        while (owner > 0 && (int64_t(owner) * n_rows) / world > row) --owner;
        const int lo = int((int64_t(owner) * n_rows) / world);
        uint64_t* dst = peers.keys[owner] + size_t(epoch & 1u) * buf_stride + (size_t(rank) * rows_cap + size_t(row - lo)) * k_cap;
        for (int t = lane; t < k_cap; t += 32) {
            uint64_t key = 0ull;
            if (t < k) {
                const int64_t i = idx[int64_t(row) * k + t];
                if (i >= 0) key = make_key(vals[int64_t(row) * k + t], uint32_t(i));
            }
            dst[t] = key;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const uint32_t prev = atomicAdd(done_counter, 1u);
        if (prev == gridDim.x - 1) {                                        // every CTA's stores are ordered before this point
            *done_counter = 0u;
            __threadfence_system();
            for (int p = 0; p < world; ++p) st_release_sys(peers.flags[p] + rank, epoch);
        }
    }
}

// spins until every sender has published `epoch` here (epochs only grow; compare as a signed difference)
__global__ void wait_flags_kernel(const uint32_t* flags, int world, uint32_t epoch, int* error_flag) {
    const int t = threadIdx.x;
    if (t < world) {
        const unsigned long long t0 = clock64();
        while (int32_t(ld_acquire_sys(flags + t) - epoch) < 0) {
            if (clock64() - t0 > PEER_WAIT_TIMEOUT_CYCLES) { atomicExch(error_flag, 1 + t); break; }
            __nanosleep(200);
        }
    }
}

// Certificate of a merge whose senders shipped only their best k_cap < k_out candidates per row ("rank-budgeted" local
// top-k: on P shards a row's global top-k takes ~k/P items from each, so a sender re-scores and ships k_cap ~ k/P + margin
// instead of k).  The merged top-k_out is exact iff no FULL sender list was consumed entirely: a sender whose last shipped
// key is >= the k_out-th merged key may hold further items above it.  Rows that fail are counted (cumulative counter in the
// channel buffer) and the caller recomputes the batch with k_cap = k_out; rows that pass are exact, whatever the placement
// of the items.  One thread per owned row.
__global__ void merge_certificate_kernel(const uint64_t* __restrict__ keys, int world, int rows, int rows_cap, int k_cap,
                                         const float* __restrict__ out_vals, const int64_t* __restrict__ out_idx, int k_out,
                                         uint32_t* __restrict__ fail_counter) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const int64_t i_k = out_idx[int64_t(row) * k_out + k_out - 1];
    const uint64_t key_k = i_k < 0 ? 0ull : make_key(out_vals[int64_t(row) * k_out + k_out - 1], uint32_t(i_k));
    bool fail = false;
    for (int p = 0; p < world; ++p) {
        const uint64_t last = __ldcg(keys + (size_t(p) * rows_cap + row) * k_cap + (k_cap - 1));
        fail |= last != 0ull && last >= key_k;
    }
    if (fail) atomicAdd(fail_counter, 1u);
}

// ---- owner side in ONE kernel: wait for the senders, merge the P sorted lists of each owned row, certify ------------------
// Every sender's list is sorted best-first (it is a local top-k; padding keys 0 at the end), so a key's place in the merged
// order is known without sorting:  rank(x) = (its position in its own list) + sum over the other lists of #{y > x}, each
// count one binary search.  One warp per row, the row's P x k_cap keys staged in shared memory; keys with rank < k_out go
// straight to out[rank].  The certificate of the rank-budgeted form falls out of the same ranks: the merged row is exact
// unless the LAST key of a full list made it into the top k_out (that sender may hold more above the cut).
constexpr int MG_WARPS = 8;

__device__ __forceinline__ uint32_t count_greater_desc(const uint64_t* __restrict__ list, uint32_t n, uint64_t x) {
    // list[0 .. n) descending; number of entries > x
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (list[mid] > x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(MG_WARPS * 32)
wait_merge_certify_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ flags, int world, uint32_t epoch,
                          int rows, int rows_cap, int k_cap, int k_out, float* __restrict__ out_vals, int64_t* __restrict__ out_idx,
                          uint32_t* __restrict__ fail_counter, int* __restrict__ error_flag) {
    extern __shared__ __align__(16) uint64_t mg_smem[];
    if (threadIdx.x < uint32_t(world)) {
        const unsigned long long t0 = clock64();
        while (int32_t(ld_acquire_sys(flags + threadIdx.x) - epoch) < 0) {
            if (clock64() - t0 > PEER_WAIT_TIMEOUT_CYCLES) { atomicExch(error_flag, 1 + int(threadIdx.x)); break; }
            __nanosleep(100);
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = int(lane_id());
    const int total = world * k_cap;
    uint64_t* mine = mg_smem + size_t(warp) * total;
    for (int row = blockIdx.x * MG_WARPS + warp; row < rows; row += gridDim.x * MG_WARPS) {
        for (int t = lane; t < total; t += 32) {
            const int p = t / k_cap, i = t - p * k_cap;
            mine[t] = __ldcg(keys + (size_t(p) * rows_cap + row) * k_cap + i);
        }
        __syncwarp();
        uint32_t n_real = 0;
        bool fail = false;
        for (int t = lane; t < total; t += 32) {
            const uint64_t x = mine[t];
            if (x == 0ull) continue;
            ++n_real;
            const int p = t / k_cap, i = t - p * k_cap;
            uint32_t rank = uint32_t(i);
            for (int q = 0; q < world; ++q)
                if (q != p) rank += count_greater_desc(mine + q * k_cap, uint32_t(k_cap), x);
            if (rank < uint32_t(k_out)) {
                out_vals[int64_t(row) * k_out + rank] = key_score(x);
                out_idx[int64_t(row) * k_out + rank] = int64_t(key_index(x));
                fail |= (i == k_cap - 1);                 // the last key of a FULL list is part of the answer
            }
        }
        n_real = warp_sum(n_real);
        for (int t = int(n_real) + lane; t < k_out; t += 32) {      // fewer candidates than k_out: pad
            out_vals[int64_t(row) * k_out + t] = ANNCUR_PAD_VAL;
            out_idx[int64_t(row) * k_out + t] = -1;
        }
        if (k_cap < k_out) {
            fail = __any_sync(0xffffffffu, fail);
            // a row that comes out short of k_out although some sender's list was full fails as well
            bool any_full = false;
            for (int q = lane; q < world; q += 32) any_full |= mine[q * k_cap + k_cap - 1] != 0ull;
            any_full = __any_sync(0xffffffffu, any_full);
            if (lane == 0 && (fail || (n_real < uint32_t(k_out) && any_full))) atomicAdd(fail_counter, 1u);
        }
        __syncwarp();
    }
}

// ---- host side --------------------------------------------------------------------------------------------------------
static size_t peer_keys_bytes(int world, int rows_cap, int k_cap) { return align_up(sizeof(uint64_t) * 2 * size_t(world) * rows_cap * k_cap, 256); }

size_t peer_channel_bytes(int world, int rows_cap, int k_cap) {
    if (world < 1 || rows_cap < 1 || k_cap < 1) return 256;
    return peer_keys_bytes(world, rows_cap, k_cap) + 256 /*flags[P]*/ + 256 /*done counter, error flag*/;
}

int peer_alloc(size_t bytes, void** out) {
    void* p = nullptr;
    ANNCUR_CUDA_OK(cudaMalloc(&p, bytes));            // plain cudaMalloc: the only memory CUDA IPC can export
    ANNCUR_CUDA_OK(cudaMemset(p, 0, bytes));
    ANNCUR_CUDA_OK(cudaDeviceSynchronize());
    *out = p;
    return ANNCUR_OK;
}
int peer_free(void* p) {
    if (p) ANNCUR_CUDA_OK(cudaFree(p));
    return ANNCUR_OK;
}
int peer_export(const void* base, void* handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");
    cudaIpcMemHandle_t h;
    ANNCUR_CUDA_OK(cudaIpcGetMemHandle(&h, const_cast<void*>(base)));
    memcpy(handle64, &h, sizeof(h));
    return ANNCUR_OK;
}
int peer_open(const void* handle64, void** out) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* p = nullptr;
    ANNCUR_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *out = p;
    return ANNCUR_OK;
}
int peer_close(void* mapped) {
    if (mapped) ANNCUR_CUDA_OK(cudaIpcCloseMemHandle(mapped));
    return ANNCUR_OK;
}

static int fill_ptrs(PeerPtrs& pp, void* const* peer_bases, int world, int rows_cap, int k_cap) {
    const size_t kb = peer_keys_bytes(world, rows_cap, k_cap);
    for (int p = 0; p < world; ++p) {
        if (!peer_bases[p]) { set_error("peer exchange: null base pointer for rank %d", p); return ANNCUR_E_INVALID; }
        pp.keys[p] = reinterpret_cast<uint64_t*>(peer_bases[p]);
        pp.flags[p] = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(peer_bases[p]) + kb);
    }
    return ANNCUR_OK;
}

int peer_scatter_keys(const float* vals, const int64_t* idx, int n_rows, int k, int rank, int world, int rows_cap, int k_cap,
                      uint32_t epoch, void* const* peer_bases, cudaStream_t stream) {
    if (world < 1 || world > kMaxPeers) { set_error("peer exchange: world %d outside [1, %d]", world, kMaxPeers); return ANNCUR_E_UNSUPPORTED; }
    if (rank < 0 || rank >= world || k < 1 || k > k_cap || n_rows < 0) { set_error("peer exchange: bad rank / k / rows"); return ANNCUR_E_INVALID; }
    if ((int64_t(n_rows) + world - 1) / world > rows_cap) { set_error("peer exchange: %d rows over %d ranks exceed rows_cap %d", n_rows, world, rows_cap); return ANNCUR_E_INVALID; }
    PeerPtrs pp{};
    int rc = fill_ptrs(pp, peer_bases, world, rows_cap, k_cap);
    if (rc != ANNCUR_OK) return rc;
    uint32_t* done = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(peer_bases[rank]) + peer_keys_bytes(world, rows_cap, k_cap) + 256);
    int grid = (n_rows + 7) / 8;
    if (grid < 1) grid = 1;
    if (grid > 4 * sm_count()) grid = 4 * sm_count();
    scatter_keys_signal_kernel<<<grid, 256, 0, stream>>>(vals, idx, n_rows, k, rank, world, rows_cap, k_cap, epoch, pp, done);
    ANNCUR_LAUNCH_OK("scatter_keys_signal_kernel");
    return ANNCUR_OK;
}

int peer_merge_owned(void* local_base, int rank, int world, int rows_owned, int rows_cap, int k_cap, int k_out, uint32_t epoch,
                     float* out_vals, int64_t* out_idx, uint32_t* scratch_rows, cudaStream_t stream) {
    if (world < 1 || world > kMaxPeers) { set_error("peer exchange: world %d outside [1, %d]", world, kMaxPeers); return ANNCUR_E_UNSUPPORTED; }
    if (rows_owned > rows_cap || rows_owned < 0) { set_error("peer exchange: rows_owned %d > rows_cap %d", rows_owned, rows_cap); return ANNCUR_E_INVALID; }
    (void)rank;
    char* base = reinterpret_cast<char*>(local_base);
    const size_t kb = peer_keys_bytes(world, rows_cap, k_cap);
    const uint32_t* flags = reinterpret_cast<const uint32_t*>(base + kb);
    int* err = reinterpret_cast<int*>(base + kb + 256 + 4);
    const uint64_t* keys = reinterpret_cast<const uint64_t*>(base) + size_t(epoch & 1u) * size_t(world) * rows_cap * k_cap;
    const size_t mg_smem = sizeof(uint64_t) * size_t(MG_WARPS) * world * k_cap;
    if (rows_owned > 0 && mg_smem <= 96 * 1024) {
        // the usual case: wait + merge + certificate in one launch
        uint32_t* fails = reinterpret_cast<uint32_t*>(base + kb + 256 + 8);
        ANNCUR_CUDA_OK(cudaFuncSetAttribute(wait_merge_certify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(mg_smem)));
        int grid = (rows_owned + MG_WARPS - 1) / MG_WARPS;
        if (grid > 8 * sm_count()) grid = 8 * sm_count();
        wait_merge_certify_kernel<<<grid, MG_WARPS * 32, mg_smem, stream>>>(keys, flags, world, epoch, rows_owned, rows_cap, k_cap, k_out,
                                                                            out_vals, out_idx, fails, err);
        ANNCUR_LAUNCH_OK("wait_merge_certify_kernel");
        return ANNCUR_OK;
    }
    wait_flags_kernel<<<1, 32 * ((world + 31) / 32), 0, stream>>>(flags, world, epoch, err);
    ANNCUR_LAUNCH_OK("wait_flags_kernel");
    if (rows_owned == 0) return ANNCUR_OK;
    int rc = merge_topk_keys_strided(keys, world, rows_owned, k_cap, int64_t(rows_cap) * k_cap, k_out, out_vals, out_idx, scratch_rows, stream);
    if (rc != ANNCUR_OK || k_cap >= k_out || world == 1) return rc;
    uint32_t* fails = reinterpret_cast<uint32_t*>(base + kb + 256 + 8);
    merge_certificate_kernel<<<(rows_owned + 127) / 128, 128, 0, stream>>>(keys, world, rows_owned, rows_cap, k_cap, out_vals, out_idx, k_out, fails);
    ANNCUR_LAUNCH_OK("merge_certificate_kernel");
    return ANNCUR_OK;
}

// rows (cumulative over the calls on this channel) whose rank-budgeted merge failed its certificate (blocks on the stream)
int peer_cert_failures(void* local_base, int world, int rows_cap, int k_cap, int reset, unsigned* count_host, cudaStream_t stream) {
    char* p = reinterpret_cast<char*>(local_base) + peer_keys_bytes(world, rows_cap, k_cap) + 256 + 8;
    ANNCUR_CUDA_OK(cudaMemcpyAsync(count_host, p, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
    if (reset) ANNCUR_CUDA_OK(cudaMemsetAsync(p, 0, sizeof(unsigned), stream));
    ANNCUR_CUDA_OK(cudaStreamSynchronize(stream));
    return ANNCUR_OK;
}

// 0 = no sender timed out so far; 1 + s = the wait for sender s gave up (blocks on the stream)
int peer_error(void* local_base, int world, int rows_cap, int k_cap, int* err_host, cudaStream_t stream) {
    const char* base = reinterpret_cast<const char*>(local_base);
    ANNCUR_CUDA_OK(cudaMemcpyAsync(err_host, base + peer_keys_bytes(world, rows_cap, k_cap) + 256 + 4, sizeof(int), cudaMemcpyDeviceToHost, stream));
    ANNCUR_CUDA_OK(cudaStreamSynchronize(stream));
    return ANNCUR_OK;
}

}  // namespace anncur
