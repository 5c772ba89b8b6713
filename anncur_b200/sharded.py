"""Item-sharded ANNCUR search over the GPUs of one node (not in the reference).

One process per GPU (torch.distributed, NCCL over NVLink).  Items are independent columns of E, so
rank p of P holds the contiguous slice E[:, lo_p:hi_p]; every rank scores the same query batch
against its slice with the fused kernel and emits a local top-k carrying GLOBAL item indices
(idx_offset = lo_p).  Exactly one exchange step follows, in one of three forms:

``search``          all-gather of the 8-byte candidate keys (B*k*8 bytes per rank), merge of all B rows on
                    EVERY rank: every rank ends with the full answer (the form north_star describes).
``search_rowblock`` query rows are OWNED in contiguous blocks (rank o merges rows [o*B/P, (o+1)*B/P)): each
                    rank's lists travel only to the row's owner and the merge work is split P ways.  The exchange
                    is either ``exchange="p2p"``: one kernel converts (value, index) to keys and stores them
                    straight into the owner's receive buffer through NVLink peer memory, flags instead of a
                    collective (csrc/peer_exchange.cu) -- or ``exchange="nccl"``: topk_to_keys +
                    ``all_to_all_single``.  Returns this rank's block of the answer.
                    With ``local_k`` < k each shard re-scores and ships only its best ``local_k`` candidates per row
                    (a row's global top-k takes ~k/P from each of P shards); the merge certifies every row on the device
                    and ``search_rowblock_verified`` falls back to the full-k exchange when a certificate fails.
``search_owned``    the same for a query batch that arrives distributed (each rank passes ITS block of rows,
                    e.g. straight from its own host buffer): one all-gather of the query blocks, then
                    ``search_rowblock``.  Host traffic per step is then that of a single GPU in total.

Equal to the single-GPU answer up to the (deterministic) tie order.

``local_search`` / ``merge`` are injectable so that the partition / exchange plumbing can be
exercised with the gloo backend on CPU in tests (where the CPU oracle stands in for the kernels);
the defaults are the CUDA kernels and nothing else.
"""
import ctypes as C
import os

import torch
import torch.distributed as dist


def shard_bounds(n_items, world_size):
    """Contiguous balanced item ranges: [(lo_0, hi_0), ...]; sizes differ by at most one."""
    return [((p * n_items) // world_size, ((p + 1) * n_items) // world_size) for p in range(world_size)]


def suggest_local_k(k, world_size):
    """Candidates a shard re-scores and ships per row in the rank-budgeted exchange: the share of a row's global top-k that
    falls into one of P equal shards is Binomial(k, 1/P) for items placed independently of their scores -- mean k/P plus six
    standard deviations plus 8.  (A placement that defeats this only costs the fallback; the certificate catches it.)"""
    P = int(world_size)
    if P <= 1:
        return int(k)
    mean = k / P
    budget = int(mean + 6.0 * (mean * (1.0 - 1.0 / P)) ** 0.5 + 8.0 + 0.999)
    return max(1, min(int(k), max(budget, -(-int(k) // P))))


def pack_candidates(vals, idx):
    """(fp32 [B x k], int64 [B x k]) -> one int64 [B x 2k] buffer (scores bit-cast into the low word)."""
    bits = vals.contiguous().view(torch.int32).to(torch.int64)
    return torch.cat([idx, bits], dim=1).contiguous()


def unpack_candidates(buf, k):
    """Inverse of pack_candidates for a gathered [P x B x 2k] buffer -> ([B x P*k] vals, [B x P*k] idx)."""
    P, B, _ = buf.shape
    idx = buf[:, :, :k].permute(1, 0, 2).reshape(B, P * k)
    vals = buf[:, :, k:].to(torch.int32).view(torch.float32).permute(1, 0, 2).reshape(B, P * k)
    return vals.contiguous(), idx.contiguous()


class PeerMemoryUnavailable(RuntimeError):
    """CUDA IPC / peer access is not available between some pair of ranks (raised on EVERY rank of the group)."""


class PeerChannel:
    """One exchange channel of the peer-memory path: this rank's receive buffer (library-allocated, plain
    cudaMalloc) mapped into every peer with CUDA IPC, and the peers' buffers mapped here.  Collective to
    construct (handles travel through ``all_gather_object``).  One channel serves calls that are stream-ordered;
    use one channel per concurrently used stream."""

    def __init__(self, n_rows, k, device, group=None):
        from . import _lib
        self._lib = lib = _lib.load()
        self._check = _lib.check
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n_rows, self.k = int(n_rows), int(k)
        self.rows_cap = max(1, -(-self.n_rows // self.world))
        self.device = device
        self.epoch = 0
        self.nbytes = int(lib.anncur_peer_channel_bytes(self.world, self.rows_cap, self.k))
        base = C.c_void_p(0)
        with torch.cuda.device(device):
            self._check(lib.anncur_peer_alloc(self.nbytes, C.byref(base)))
            self.base = base.value
            handle = (C.c_ubyte * 64)()
            self._check(lib.anncur_peer_export(C.c_void_p(self.base), handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self._mapped = []
        ptrs = (C.c_void_p * self.world)()
        failure = None
        with torch.cuda.device(device):
            for p, h in enumerate(handles):
                if p == self.rank:
                    ptrs[p] = self.base
                    continue
                m = C.c_void_p(0)
                hb = (C.c_ubyte * 64).from_buffer_copy(h)
                rc = lib.anncur_peer_open(hb, C.byref(m))
                if rc != 0:
                    failure = f"rank {self.rank}: cannot map rank {p}'s buffer: {lib.anncur_last_error().decode('utf-8', 'replace')}"
                    break
                self._mapped.append(m.value)
                ptrs[p] = m.value
        # every rank learns whether EVERY mapping worked (no peer access between some pair of GPUs -> nobody uses the channel)
        failures = [None] * self.world
        dist.all_gather_object(failures, failure, group=group)
        if any(failures):
            with torch.cuda.device(device):
                for mm in self._mapped:
                    lib.anncur_peer_close(C.c_void_p(mm))
                lib.anncur_peer_free(C.c_void_p(self.base))
            self._mapped, self.base = [], None
            raise PeerMemoryUnavailable("; ".join(f for f in failures if f))
        self.ptrs = ptrs
        lo, hi = shard_bounds(self.n_rows, self.world)[self.rank]
        self.row_lo, self.rows_owned = lo, hi - lo
        self._ws = torch.empty(max(int(lib.anncur_merge_topk_keys_workspace_bytes(self.rows_cap)), 256), dtype=torch.uint8, device=device)
        dist.barrier(group=group)                       # every rank has mapped every buffer before the first store

    def exchange(self, vals, idx, k_out=None):
        """Local top-k of all n_rows rows -> merged top-k_out of this rank's owned rows ([rows_owned x k_out])."""
        lib = self._lib
        assert vals.shape == (self.n_rows, self.k) and idx.shape == (self.n_rows, self.k)
        assert vals.is_contiguous() and idx.is_contiguous() and vals.dtype == torch.float32 and idx.dtype == torch.int64
        k_out = self.k if k_out is None else int(k_out)
        self.epoch += 1
        out_v = torch.empty((self.rows_owned, k_out), dtype=torch.float32, device=vals.device)
        out_i = torch.empty((self.rows_owned, k_out), dtype=torch.int64, device=vals.device)
        stream = C.c_void_p(torch.cuda.current_stream(vals.device).cuda_stream)
        with torch.cuda.device(vals.device):
            self._check(lib.anncur_peer_scatter_keys(C.c_void_p(vals.data_ptr()), C.c_void_p(idx.data_ptr()), self.n_rows, self.k,
                                                     self.rank, self.world, self.rows_cap, self.k, self.epoch, self.ptrs, stream))
            self._check(lib.anncur_peer_merge_owned(C.c_void_p(self.base), self.rank, self.world, self.rows_owned, self.rows_cap,
                                                    self.k, k_out, self.epoch, C.c_void_p(out_v.data_ptr()),
                                                    C.c_void_p(out_i.data_ptr()), C.c_void_p(self._ws.data_ptr()),
                                                    self._ws.numel(), stream))
        return out_v, out_i

    def cert_failures(self, reset=False):
        """Rows whose rank-budgeted merge (k_out > this channel's k) failed the exactness certificate since the last reset,
        on THIS rank's owned rows (synchronises the current stream)."""
        n = C.c_uint(0)
        with torch.cuda.device(self.device):
            self._check(self._lib.anncur_peer_cert_failures(C.c_void_p(self.base), self.world, self.rows_cap, self.k, 1 if reset else 0,
                                                            C.byref(n), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return int(n.value)

    def error(self):
        """0, or 1 + s when a wait for sender s timed out (synchronises the current stream)."""
        e = C.c_int(0)
        with torch.cuda.device(self.device):
            self._check(self._lib.anncur_peer_error(C.c_void_p(self.base), self.world, self.rows_cap, self.k, C.byref(e),
                                                    C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return e.value

    def close(self):
        if getattr(self, "base", None):
            with torch.cuda.device(self.device):
                torch.cuda.synchronize(self.device)
                dist.barrier(group=self.group)          # nobody stores into a buffer that is about to go away
                for m in self._mapped:
                    self._lib.anncur_peer_close(C.c_void_p(m))
                self._lib.anncur_peer_free(C.c_void_p(self.base))
            self._mapped, self.base = [], None


class ShardedIndex:
    """This rank's slice of the item-embedding matrix plus the collective search."""

    def __init__(self, E_local, lo, n_items_total, *, precision="f32r", group=None, local_search=None, merge=None,
                 packed=None, exchange=None):
        self.group = group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.lo = int(lo)
        self.n_items_total = int(n_items_total)
        self.precision = precision
        self._E_local = E_local
        self._channels = {}
        self._cuda = local_search is None
        if local_search is None:
            from . import engine
            engine.require_cuda()
            self._packed = packed if packed is not None else engine.PackedItems(E_local, precision)
            local_search = lambda Q, k: engine.score_topk(Q, self._packed, k, idx_offset=self.lo)   # noqa: E731
            merge = merge or engine.merge_topk
            self._key_path = self.n_items_total < 2 ** 32          # 8-byte keys on the wire, merged as gathered
        self._local_search = local_search
        self._merge = merge
        self._key_path = getattr(self, "_key_path", False)
        # exchange form of search_rowblock: "p2p" (NVLink peer stores, default on CUDA) or "nccl" (all_to_all_single)
        want = exchange or os.environ.get("ANNCUR_EXCHANGE") or ("p2p" if (self._cuda and self._key_path) else "nccl")
        if want not in ("p2p", "nccl"):
            raise ValueError(f"exchange must be 'p2p' or 'nccl', not {want!r}")
        if want == "p2p" and not (self._cuda and self._key_path):
            raise ValueError("exchange='p2p' needs the CUDA kernels and global item indices below 2^32")
        self.exchange = want

    @classmethod
    def from_full(cls, E_full, **kw):
        """Convenience: slice a replicated E (k_i x N) for this rank."""
        ws = dist.get_world_size(kw.get("group")) if dist.is_initialized() else 1
        rk = dist.get_rank(kw.get("group")) if dist.is_initialized() else 0
        lo, hi = shard_bounds(E_full.shape[1], ws)[rk]
        return cls(E_full[:, lo:hi], lo, E_full.shape[1], **kw)

    def local_topk(self, Q, k):
        """Local candidates, padded to k with (idx = -1) when the shard holds fewer than k items."""
        return self._local_search(Q, k)

    # ---- all-gather form: the full answer on every rank ---------------------------------------------------------
    def search(self, Q, k):
        """Global top-k for the replicated query batch Q (B x k_i): values fp32, indices int64 (global)."""
        vals, idx = self.local_topk(Q, k)
        if self.world_size == 1:
            return vals, idx
        if self._key_path:
            from . import engine
            mine = engine.topk_to_keys(vals, idx)
            gathered = torch.empty((self.world_size,) + tuple(mine.shape), dtype=mine.dtype, device=mine.device)
            dist.all_gather_into_tensor(gathered, mine, group=self.group)   # [P, B, k]: merged as is
            return engine.merge_topk_keys(gathered, k)
        mine = pack_candidates(vals, idx)
        B = mine.shape[0]
        gathered = torch.empty((self.world_size * B, mine.shape[1]), dtype=mine.dtype, device=mine.device)
        dist.all_gather_into_tensor(gathered, mine, group=self.group)       # rank-major concatenation
        cand_vals, cand_idx = unpack_candidates(gathered.view(self.world_size, B, mine.shape[1]), k)
        return self._merge(cand_vals, cand_idx, k)

    # ---- row-block form: each rank ends with the answer for the rows it owns ------------------------------------
    def row_block(self, n_rows):
        """[lo, hi) of the query rows this rank owns in a batch of n_rows."""
        return shard_bounds(n_rows, self.world_size)[self.rank]

    def _channel(self, n_rows, k, device):
        key = (int(n_rows), int(k), torch.cuda.current_stream(device).cuda_stream)
        ch = self._channels.get(key)
        if ch is None:
            try:
                ch = self._channels[key] = PeerChannel(n_rows, k, device, group=self.group)
            except PeerMemoryUnavailable as exc:          # the same on every rank: switch the whole group to the NCCL exchange
                import sys
                print(f"[anncur_b200.sharded] peer-memory exchange unavailable ({exc}); using NCCL all_to_all_single", file=sys.stderr, flush=True)
                self.exchange = "nccl"
                return None
        return ch

    def prepare(self, n_rows, k, device=None):
        """Collective: set up the peer channel for (n_rows, k) on the current stream ahead of the first search."""
        if self.exchange == "p2p" and self.world_size > 1:
            self._channel(n_rows, k, device if device is not None else self._packed.device)

    def search_rowblock(self, Q, k, local_k=None):
        """Replicated Q (B x k_i) -> (vals, idx) of this rank's owned rows ``row_block(B)``.

        ``local_k`` < k (peer-memory exchange only) is the rank-budgeted form: every shard re-scores and ships only its best
        ``local_k`` candidates per row (see ``suggest_local_k``), the owner's merge evaluates the exactness certificate of
        every row on the device, and ``certificate_failures()`` tells whether any row of any call since the last check
        needs the full-k answer -- ``search_rowblock_verified`` does that bookkeeping."""
        k_loc = int(k) if (local_k is None or self.exchange != "p2p" or self.world_size == 1) else max(1, min(int(local_k), int(k)))
        if k_loc * self.world_size < k:
            k_loc = -(-int(k) // self.world_size)
        vals, idx = self.local_topk(Q, k_loc)
        B = vals.shape[0]
        if self.world_size == 1:
            return vals, idx
        if self.exchange == "p2p":
            ch = self._channel(B, k_loc, vals.device)
            if ch is not None:
                return ch.exchange(vals, idx, k_out=k)
            if k_loc != k:                                # the NCCL form ships full lists: redo the local search with k
                vals, idx = self.local_topk(Q, k)
        bounds = shard_bounds(B, self.world_size)
        rows = [hi - lo for lo, hi in bounds]
        mine = rows[self.rank]
        if self._key_path:
            from . import engine
            keys = engine.topk_to_keys(vals, idx)                                   # [B, k] int64
            recv = torch.empty((self.world_size * mine, k), dtype=keys.dtype, device=keys.device)
            dist.all_to_all_single(recv, keys, output_split_sizes=[mine] * self.world_size, input_split_sizes=rows, group=self.group)
            return engine.merge_topk_keys(recv.view(self.world_size, mine, k), k)
        packed = pack_candidates(vals, idx)                                         # [B, 2k]
        recv = torch.empty((self.world_size * mine, 2 * k), dtype=packed.dtype, device=packed.device)
        dist.all_to_all_single(recv, packed, output_split_sizes=[mine] * self.world_size, input_split_sizes=rows, group=self.group)
        cand_vals, cand_idx = unpack_candidates(recv.view(self.world_size, mine, 2 * k), k)
        return self._merge(cand_vals, cand_idx, k)

    def certificate_failures(self, reset=True):
        """Collective: rows (over all ranks and all rank-budgeted calls since the last reset) whose merge failed its
        certificate.  0 = every answer returned so far is the exact top-k."""
        n = sum(ch.cert_failures(reset) for ch in self._channels.values())
        if self.world_size > 1:
            t = torch.tensor([n], dtype=torch.int64, device=self._packed.device if self._cuda else "cpu")
            dist.all_reduce(t, group=self.group)
            n = int(t.item())
        return n

    def search_rowblock_verified(self, Q, k, local_k=None):
        """``search_rowblock`` with the rank-budgeted shortcut made safe: if any row of the batch fails the certificate (on any
        rank) the batch is recomputed with local_k = k.  Synchronises (one small all-reduce + host read-back per call)."""
        out = self.search_rowblock(Q, k, local_k)
        if local_k is not None and local_k < k and self.exchange == "p2p" and self.world_size > 1:
            if self.certificate_failures(reset=True) > 0:
                out = self.search_rowblock(Q, k, None)
        return out

    def search_owned(self, Q_block, n_rows_total, k, local_k=None):
        """Q_block = this rank's ``row_block(n_rows_total)`` of the batch.  All-gathers the blocks (NVLink), searches,
        and returns the answer for the same rows."""
        if self.world_size == 1:
            return self.local_topk(Q_block, k)
        bounds = shard_bounds(n_rows_total, self.world_size)
        rows = [hi - lo for lo, hi in bounds]
        assert Q_block.shape[0] == rows[self.rank], (Q_block.shape, rows)
        if len(set(rows)) == 1:
            Q = torch.empty((n_rows_total, Q_block.shape[1]), dtype=Q_block.dtype, device=Q_block.device)
            dist.all_gather_into_tensor(Q, Q_block.contiguous(), group=self.group)
        else:                                   # ragged blocks: gather padded blocks, then drop the padding
            cap = max(rows)
            mine = torch.zeros((cap, Q_block.shape[1]), dtype=Q_block.dtype, device=Q_block.device)
            mine[:rows[self.rank]] = Q_block
            allq = torch.empty((self.world_size * cap, Q_block.shape[1]), dtype=Q_block.dtype, device=Q_block.device)
            dist.all_gather_into_tensor(allq, mine, group=self.group)
            Q = torch.cat([allq[p * cap:p * cap + rows[p]] for p in range(self.world_size)], dim=0)
        return self.search_rowblock(Q, k, local_k)

    def search_owned_verified(self, Q_block, n_rows_total, k, local_k=None):
        """``search_owned`` with the rank-budgeted shortcut made safe (see ``search_rowblock_verified``): one small all-reduce
        and a host read-back per call; a failed certificate anywhere recomputes the batch with local_k = k on every rank."""
        out = self.search_owned(Q_block, n_rows_total, k, local_k)
        if local_k is not None and local_k < k and self.exchange == "p2p" and self.world_size > 1:
            if self.certificate_failures(reset=True) > 0:
                out = self.search_owned(Q_block, n_rows_total, k, None)
        return out

    def close(self):
        for ch in self._channels.values():
            ch.close()
        self._channels.clear()
