"""CPU, world_size 2 (and 3) over gloo: the partition / pack / all-gather / merge plumbing of
anncur_b200.sharded, with the CPU oracle injected in place of the CUDA kernels (tests may do that;
the product default is CUDA-only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cur_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, k, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from anncur_b200.sharded import ShardedIndex, shard_bounds
        E = torch.from_numpy(O.synthetic_scores(12, 1001, rank=6, seed=1))
        Q = torch.from_numpy(O.synthetic_scores(9, 12, rank=6, seed=2))
        lo, hi = shard_bounds(E.shape[1], world)[rank]

        def local_search(Qb, kk, lo=lo, hi=hi):
            kk2 = min(kk, hi - lo)
            t = torch.topk(Qb @ E[:, lo:hi], kk2, dim=1)
            v = torch.full((Qb.shape[0], kk), -np.finfo(np.float32).max)
            i = torch.full((Qb.shape[0], kk), -1, dtype=torch.int64)
            v[:, :kk2], i[:, :kk2] = t.values, t.indices + lo
            return v, i

        def merge(cv, ci, kk):
            v, i = O.merge_topk(cv.numpy()[None], ci.numpy()[None], kk)
            return torch.from_numpy(v), torch.from_numpy(i)

        index = ShardedIndex(E[:, lo:hi], lo, E.shape[1], local_search=local_search, merge=merge)
        v, i = index.search(Q, k)
        ref = O.score_topk(Q, E, k)
        ok = torch.equal(i, ref.indices) and torch.allclose(v, ref.values)
        # row-block form: every rank ends with the rows it owns (all-to-all by row block), from a replicated batch ...
        r0, r1 = index.row_block(Q.shape[0])
        v2, i2 = index.search_rowblock(Q, k)
        ok = ok and torch.equal(i2, ref.indices[r0:r1]) and torch.allclose(v2, ref.values[r0:r1])
        # ... and from a batch that arrives distributed (each rank passes its own rows; ragged blocks: 9 rows over 2 / 3 ranks)
        v3, i3 = index.search_owned(Q[r0:r1].contiguous(), Q.shape[0], k)
        ok = ok and torch.equal(i3, ref.indices[r0:r1]) and torch.allclose(v3, ref.values[r0:r1])
        out = torch.tensor([1 if ok else 0])
        dist.all_reduce(out, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(int(out.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,k", [(2, 10), (3, 400)])
def test_sharded_search_equals_single_rank(world, k):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, k, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1


def test_shard_bounds_and_packing_roundtrip():
    from anncur_b200.sharded import pack_candidates, shard_bounds, unpack_candidates
    b = shard_bounds(10_000_001, 8)
    assert b[0][0] == 0 and b[-1][1] == 10_000_001 and all(b[i][1] == b[i + 1][0] for i in range(7))
    assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    v = torch.randn(5, 7)
    v[0, 0] = float(-np.finfo(np.float32).max)
    i = torch.randint(-1, 2**40, (5, 7))
    buf = torch.stack([pack_candidates(v, i), pack_candidates(v + 1, i + 1)])
    cv, ci = unpack_candidates(buf, 7)
    assert torch.equal(cv[:, :7], v) and torch.equal(ci[:, :7], i) and torch.equal(cv[:, 7:], v + 1)
