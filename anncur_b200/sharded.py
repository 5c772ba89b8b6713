"""Item-sharded ANNCUR search over the GPUs of one node (not in the reference).

One process per GPU (torch.distributed, NCCL over NVLink).  Items are independent columns of E, so
rank p of P holds the contiguous slice E[:, lo_p:hi_p]; every rank scores the same query batch
against its slice with the fused kernel and emits a local top-k carrying GLOBAL item indices
(idx_offset = lo_p).  The only exchange is one all-gather of the packed (idx, score) lists --
B*k*16 bytes per rank -- followed by the K9 merge kernel on every rank.  Equal to the single-GPU
answer up to the (deterministic) tie order.

``local_search`` / ``merge`` are injectable so that the partition / gather plumbing can be
exercised with the gloo backend on CPU in tests (where the CPU oracle stands in for the kernels);
the defaults are the CUDA kernels and nothing else.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items, world_size):
    """Contiguous balanced item ranges: [(lo_0, hi_0), ...]; sizes differ by at most one."""
    return [((p * n_items) // world_size, ((p + 1) * n_items) // world_size) for p in range(world_size)]


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


class ShardedIndex:
    """This rank's slice of the item-embedding matrix plus the collective search."""

    def __init__(self, E_local, lo, n_items_total, *, precision="f32r", group=None, local_search=None, merge=None,
                 packed=None):
        self.group = group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.lo = int(lo)
        self.n_items_total = int(n_items_total)
        self.precision = precision
        self._E_local = E_local
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
