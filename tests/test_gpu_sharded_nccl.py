"""GPU, >= 2 devices (skipped otherwise): the item-sharded search over NCCL equals the single-GPU answer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from anncur_b200 import engine
        from anncur_b200.sharded import ShardedIndex, suggest_local_k
        rng = np.random.default_rng(4)
        E = torch.from_numpy(rng.standard_normal((96, 150_001), dtype=np.float32)).cuda()
        Q = torch.from_numpy(rng.standard_normal((300, 96), dtype=np.float32)).cuda()
        ok = 1
        full = engine.PackedItems(E)
        for k in (10, 100):
            rv, ri = engine.score_topk(Q, full, k)
            index = ShardedIndex.from_full(E)                          # all-gather form: the full answer on every rank
            v, i = index.search(Q, k)
            ok &= int(torch.equal(i, ri) and torch.allclose(v, rv, rtol=1e-5, atol=1e-5))
            r0, r1 = index.row_block(Q.shape[0])
            for exchange in ("nccl", "p2p"):                           # row-block forms: all_to_all_single / NVLink peer stores
                ix = ShardedIndex.from_full(E, exchange=exchange)
                for _ in range(3):                                     # both parity buffers of the peer channel
                    v, i = ix.search_rowblock(Q, k)
                    ok &= int(torch.equal(i, ri[r0:r1]) and torch.allclose(v, rv[r0:r1], rtol=1e-5, atol=1e-5))
                v, i = ix.search_owned(Q[r0:r1].contiguous(), Q.shape[0], k)
                ok &= int(torch.equal(i, ri[r0:r1]) and torch.allclose(v, rv[r0:r1], rtol=1e-5, atol=1e-5))
                if exchange == "p2p":
                    ok &= int(all(ch.error() == 0 for ch in ix._channels.values()))
                    # rank-budgeted form: ship local_k < k candidates per row, certified on the device
                    lk = suggest_local_k(k, world)
                    v, i = ix.search_rowblock_verified(Q, k, lk)
                    ok &= int(torch.equal(i, ri[r0:r1]) and torch.allclose(v, rv[r0:r1], rtol=1e-5, atol=1e-5))
                    v, i = ix.search_rowblock(Q, k, max(1, k // world))           # too small a budget: rows must fail, never pass short
                    n_fail = ix.certificate_failures()
                    ok &= int(n_fail > 0 or torch.equal(i, ri[r0:r1]))
                ix.close()
        # adaptive multi-round ANNCUR with the re-score item-sharded: one exchange per round, every rank ends with the same
        # anchors as the single-GPU run ("the per-query solve is replicated; re-score is sharded")
        from anncur_b200 import adaptive_anncur
        from anncur_b200.adaptive import AdaptiveIndex
        from oracle import cur_oracle as O
        A = torch.from_numpy(O.synthetic_scores(60 + 40, 20000, rank=10, noise=0.05, seed=7)).cuda()
        Ra, Xa = A[:60].contiguous(), A[60:].contiguous()
        first = torch.arange(0, 20000, 20000 // 16)[:16]
        a1, i1, v1 = adaptive_anncur(Ra, Xa, first, 4, 16, 10)
        a2, i2, v2 = adaptive_anncur(Ra, Xa, first, 4, 16, 10, index=AdaptiveIndex(Ra, sharded=ShardedIndex.from_full(Ra)))
        ok &= int(torch.equal(a1, a2) and torch.equal(i1, i2) and torch.equal(v1, v2))
        # owned form: every rank solves ITS block of the queries, the re-score is item-sharded, search_owned per round
        ixs = ShardedIndex.from_full(Ra)
        q0, q1 = ixs.row_block(Xa.shape[0])
        a3, i3, v3 = adaptive_anncur(Ra, Xa[q0:q1].contiguous(), first, 4, 16, 10, index=AdaptiveIndex(Ra, sharded=ixs), n_rows_total=Xa.shape[0])
        ok &= int(torch.equal(a1[q0:q1], a3) and torch.equal(i1[q0:q1], i3) and torch.equal(v1[q0:q1], v3))
        out = torch.tensor([ok], device="cuda")
        dist.all_reduce(out, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(int(out.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_item_sharded_search_over_nccl_equals_single_gpu():
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1
