"""GPU: the peer-memory candidate exchange (csrc/peer_exchange.cu) through the C ABI.

On one GPU the P "ranks" are P library-allocated buffers in one process (a peer pointer is then an ordinary local
pointer): scatter-as-rank-r for every r, merge-as-owner-o for every o, against the key merge that the all-gather
form uses (itself checked against the oracle in test_gpu_kernels.py).  The CUDA-IPC mapping and the cross-GPU flags
are covered by test_gpu_sharded_nccl.py on a box with >= 2 GPUs and by `bench.py --gpus N` (which asserts
sharded == single-GPU on a row sample in every multi-GPU run)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _local_topk_lists(rng, P, B, k, n_items, pad_rank=None):
    """P per-shard top-k results with global indices: sorted best-first, shard p draws from its own index range."""
    vals, idxs = [], []
    per = n_items // P
    for p in range(P):
        v = np.sort(rng.standard_normal((B, k)).astype(np.float32), axis=1)[:, ::-1].copy()
        i = np.stack([rng.choice(per, size=k, replace=False) for _ in range(B)]).astype(np.int64) + p * per
        if pad_rank == p:                       # a shard with fewer than k items pads with idx = -1
            i[:, k // 2:] = -1
            v[:, k // 2:] = -np.finfo(np.float32).max
        vals.append(torch.from_numpy(v).cuda())
        idxs.append(torch.from_numpy(i).cuda())
    return vals, idxs


@pytest.mark.parametrize("P,B,k", [(1, 5, 3), (2, 37, 10), (3, 300, 100), (8, 4096, 100), (8, 9, 7)])
def test_peer_exchange_simulated_ranks(P, B, k):
    from anncur_b200 import _lib, engine
    from anncur_b200.sharded import shard_bounds
    lib = _lib.load()
    rng = np.random.default_rng(P * 1000 + B)
    rows_cap = -(-B // P)
    nbytes = lib.anncur_peer_channel_bytes(P, rows_cap, k)
    bases = []
    for _ in range(P):
        b = C.c_void_p(0)
        _lib.check(lib.anncur_peer_alloc(nbytes, C.byref(b)))
        bases.append(b.value)
    ptrs = (C.c_void_p * P)(*bases)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ws = torch.empty(max(int(lib.anncur_merge_topk_keys_workspace_bytes(rows_cap)), 256), dtype=torch.uint8, device="cuda")
    try:
        for epoch in (1, 2, 3):                                     # both parity buffers, and the first one again
            vals, idxs = _local_topk_lists(rng, P, B, k, n_items=1_000_000, pad_rank=(P - 1 if epoch == 2 else None))
            for r in range(P):
                _lib.check(lib.anncur_peer_scatter_keys(C.c_void_p(vals[r].data_ptr()), C.c_void_p(idxs[r].data_ptr()), B, k, r, P,
                                                        rows_cap, k, epoch, ptrs, stream))
            keys = torch.stack([engine.topk_to_keys(vals[r], idxs[r]) for r in range(P)])          # [P, B, k]
            want_v, want_i = engine.merge_topk_keys(keys, k)
            for o, (lo, hi) in enumerate(shard_bounds(B, P)):
                out_v = torch.empty((hi - lo, k), dtype=torch.float32, device="cuda")
                out_i = torch.empty((hi - lo, k), dtype=torch.int64, device="cuda")
                _lib.check(lib.anncur_peer_merge_owned(C.c_void_p(bases[o]), o, P, hi - lo, rows_cap, k, k, epoch,
                                                       C.c_void_p(out_v.data_ptr()), C.c_void_p(out_i.data_ptr()),
                                                       C.c_void_p(ws.data_ptr()), ws.numel(), stream))
                torch.cuda.synchronize()
                assert torch.equal(out_i, want_i[lo:hi]) and torch.equal(out_v, want_v[lo:hi]), (epoch, o)
                err = C.c_int(-1)
                _lib.check(lib.anncur_peer_error(C.c_void_p(bases[o]), P, rows_cap, k, C.byref(err), stream))
                assert err.value == 0
    finally:
        torch.cuda.synchronize()
        for b in bases:
            lib.anncur_peer_free(C.c_void_p(b))


def test_peer_exchange_argument_errors():
    from anncur_b200 import _lib
    lib = _lib.load()
    ptrs = (C.c_void_p * 1)(0)
    v = torch.zeros((4, 2), device="cuda")
    i = torch.zeros((4, 2), dtype=torch.int64, device="cuda")
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.anncur_peer_scatter_keys(C.c_void_p(v.data_ptr()), C.c_void_p(i.data_ptr()), 4, 2, 0, 1, 4, 2, 1, ptrs, s) == _lib.E_INVALID
    assert lib.anncur_peer_scatter_keys(C.c_void_p(v.data_ptr()), C.c_void_p(i.data_ptr()), 4, 2, 0, 17, 4, 2, 1, ptrs, s) == _lib.E_UNSUPPORTED
    assert lib.anncur_peer_scatter_keys(C.c_void_p(v.data_ptr()), C.c_void_p(i.data_ptr()), 4, 3, 0, 1, 4, 2, 1, ptrs, s) == _lib.E_INVALID
