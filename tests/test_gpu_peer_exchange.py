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


def _run_exchange(lib, _lib, vals, idxs, P, B, k_loc, k_out, epoch, bases, ptrs, rows_cap, ws, stream):
    from anncur_b200.sharded import shard_bounds
    for r in range(P):
        _lib.check(lib.anncur_peer_scatter_keys(C.c_void_p(vals[r].data_ptr()), C.c_void_p(idxs[r].data_ptr()), B, k_loc, r, P,
                                                rows_cap, k_loc, epoch, ptrs, stream))
    outs, fails = [], 0
    for o, (lo, hi) in enumerate(shard_bounds(B, P)):
        out_v = torch.empty((hi - lo, k_out), dtype=torch.float32, device="cuda")
        out_i = torch.empty((hi - lo, k_out), dtype=torch.int64, device="cuda")
        _lib.check(lib.anncur_peer_merge_owned(C.c_void_p(bases[o]), o, P, hi - lo, rows_cap, k_loc, k_out, epoch,
                                               C.c_void_p(out_v.data_ptr()), C.c_void_p(out_i.data_ptr()),
                                               C.c_void_p(ws.data_ptr()), ws.numel(), stream))
        n = C.c_uint(0)
        _lib.check(lib.anncur_peer_cert_failures(C.c_void_p(bases[o]), P, rows_cap, k_loc, 1, C.byref(n), stream))
        fails += n.value
        outs.append((out_v, out_i))
    return torch.cat([v for v, _ in outs]), torch.cat([i for _, i in outs]), fails


def test_rank_budgeted_exchange_certificate():
    """Senders ship only their best k_loc < k candidates per row.  Items placed independently of their scores: every row is
    certified and the merged top-k equals the full-k merge.  Adversarial placement (one shard holds a row's whole top-k): the
    certificate must FAIL for exactly those rows -- a short answer is never passed off as the top-k."""
    from anncur_b200 import _lib, engine
    from anncur_b200.sharded import suggest_local_k
    lib = _lib.load()
    P, B, k = 4, 300, 100
    k_loc = suggest_local_k(k, P)
    assert k // P < k_loc < k
    rng = np.random.default_rng(7)
    N = 40000
    S = rng.standard_normal((B, N)).astype(np.float32)
    bad_rows = [3, 77, 299]
    S[bad_rows, : N // P] += 10.0                                        # these rows' winners all live in shard 0
    St = torch.from_numpy(S).cuda()
    per = N // P
    full_v, full_i, loc_v, loc_i = [], [], [], []
    for p in range(P):
        v, i = torch.topk(St[:, p * per:(p + 1) * per], k, dim=1)
        full_v.append(v.contiguous()), full_i.append((i + p * per).contiguous())
        loc_v.append(v[:, :k_loc].contiguous()), loc_i.append((i[:, :k_loc] + p * per).contiguous())
    want_v, want_i = engine.merge_topk_keys(torch.stack([engine.topk_to_keys(full_v[p], full_i[p]) for p in range(P)]), k)
    rows_cap = -(-B // P)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ws = torch.empty(max(int(lib.anncur_merge_topk_keys_workspace_bytes(rows_cap)), 256), dtype=torch.uint8, device="cuda")
    bases = []
    for _ in range(P):
        b = C.c_void_p(0)
        _lib.check(lib.anncur_peer_alloc(lib.anncur_peer_channel_bytes(P, rows_cap, k_loc), C.byref(b)))
        bases.append(b.value)
    ptrs = (C.c_void_p * P)(*bases)
    try:
        got_v, got_i, fails = _run_exchange(lib, _lib, loc_v, loc_i, P, B, k_loc, k, 1, bases, ptrs, rows_cap, ws, stream)
        assert fails == len(bad_rows)
        good = np.setdiff1d(np.arange(B), bad_rows)
        assert torch.equal(got_i[good], want_i[good]) and torch.equal(got_v[good], want_v[good])
        assert not torch.equal(got_i[bad_rows], want_i[bad_rows])          # short lists: what the certificate is for
        # without the planted rows nothing fails, and the counter was reset by the read above
        ok_rows = torch.from_numpy(good).cuda()
        lv = [v[ok_rows][:B - 3].contiguous() for v in loc_v]
        li = [i[ok_rows][:B - 3].contiguous() for i in loc_i]
        got_v, got_i, fails = _run_exchange(lib, _lib, lv, li, P, B - 3, k_loc, k, 2, bases, ptrs, rows_cap, ws, stream)
        assert fails == 0 and torch.equal(got_i, want_i[good][:B - 3])
    finally:
        torch.cuda.synchronize()
        for b in bases:
            lib.anncur_peer_free(C.c_void_p(b))
