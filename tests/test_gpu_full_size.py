"""GPU, BASELINE.json's full sizes: properties that do not need a CPU pass over the whole problem
(the oracle comparison at sizes it finishes in seconds lives in test_gpu_fused.py / test_gpu_kernels.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _enough_memory(gb):
    return torch.cuda.is_available() and torch.cuda.get_device_properties(0).total_memory >= gb * 2**30


@pytest.fixture(scope="module")
def eng():
    from anncur_b200 import engine
    engine.require_cuda()
    return engine


@pytest.mark.skipif(not _enough_memory(100), reason="needs a 100+ GB device")
def test_c4_size_ten_million_items(eng):
    """Config 4 size on one GPU: N = 10 000 000 items, k_i = 500, top-100.
    (1) the fused result equals a plain fp32 torch matmul + topk over the same E (tie-tolerant set equality, 1e-4 scores);
    (2) four item shards merged through the key exchange format equal the single-shard answer exactly;
    (3) results are sorted best-first with unique indices."""
    g = torch.Generator(device="cuda").manual_seed(5)
    N, K, B, k = 10_000_000, 500, 130, 100
    r = 32
    W = torch.randn(K, r, device="cuda", generator=g)
    E = torch.empty((K, N), device="cuda")
    for a in range(0, N, 1_000_000):
        Y = torch.randn(1_000_000, r, device="cuda", generator=g)
        E[:, a:a + 1_000_000] = W @ Y.t() / r ** 0.5 + 0.05 * torch.randn(K, 1_000_000, device="cuda", generator=g)
    Q = torch.randn(B, r, device="cuda", generator=g) @ W.t() / r ** 0.5 + 0.05 * torch.randn(B, K, device="cuda", generator=g)
    packed = eng.PackedItems(E, "f32r")                         # the default kind: one f16 pass of bounds + fp32 re-scoring
    v, i = eng.score_topk(Q, packed, k)
    assert eng.last_redo_rows(B, packed, k) == 0                # served by the fast path
    assert bool((v[:, :-1] >= v[:, 1:]).all())
    assert all(len(set(row.tolist())) == k for row in i.cpu())
    # (1) dense fp32 reference on the device (torch matmul with TF32 off = fp32 FFMA/cuBLAS)
    torch.backends.cuda.matmul.allow_tf32 = False
    dense = Q @ E                                               # 130 x 10M fp32 = 5.2 GB
    ref = torch.topk(dense, k, dim=1)
    scale = dense.abs().amax(dim=1, keepdim=True)
    got_scores = torch.gather(dense, 1, i)
    assert float(((v - got_scores).abs() / scale).max()) <= 1e-5
    kth = ref.values[:, -1:]
    assert bool((got_scores >= kth - 1e-4 * scale).all())       # every returned item is a top-k item up to the tie band
    same = (i.unsqueeze(2) == ref.indices.unsqueeze(1)).any(dim=2).float().mean()
    assert float(same) > 0.999
    del dense, ref, got_scores
    # the 3-pass kind gives the same items (values agree to its own 1e-6-ish accuracy)
    v3, i3 = eng.score_topk(Q, eng.PackedItems(E, "f32x3"), k)
    assert float((i3 == i).float().mean()) > 0.999 and torch.allclose(v3, v, rtol=1e-4, atol=1e-5)
    del packed
    # (2) item shards + key-form merge
    P = 4
    keys = []
    for p in range(P):
        lo, hi = p * N // P, (p + 1) * N // P
        pv, pi = eng.score_topk(Q, eng.PackedItems(E[:, lo:hi], "f32r"), k, idx_offset=lo)
        keys.append(eng.topk_to_keys(pv, pi))
    mv, mi = eng.merge_topk_keys(torch.stack(keys).contiguous(), k)
    assert torch.equal(mi, i) and torch.allclose(mv, v, rtol=1e-5, atol=1e-5)


@pytest.mark.skipif(not _enough_memory(100), reason="needs a 100+ GB device")
def test_c5_size_reconstruction_error_is_additive_over_item_shards(eng):
    """Config 5 size: 10 000 queries x 1 000 000 items, k_i = 200.  sum_j (Q E - A)^2 over all items equals the sum of
    the same quantity over item shards (what a multi-GPU run all-reduces), and a planted exact-rank matrix gives ~0."""
    g = torch.Generator(device="cuda").manual_seed(6)
    n, N, K = 10_000, 1_000_000, 200
    Qm = torch.randn(n, K, device="cuda", generator=g)
    E = torch.randn(K, N, device="cuda", generator=g) / K ** 0.5
    A = torch.empty((n, N), device="cuda")                      # 40 GB
    for a in range(0, N, 100_000):
        A[:, a:a + 100_000] = Qm @ E[:, a:a + 100_000] + 0.01 * torch.randn(n, 100_000, device="cuda", generator=g)
    err2, norm2 = eng.recon_error_rows(Qm, E, A)
    parts_e = torch.zeros_like(err2)
    parts_n = torch.zeros_like(norm2)
    for a in range(0, N, 250_000):
        e2, n2 = eng.recon_error_rows(Qm, E[:, a:a + 250_000], A[:, a:a + 250_000])
        parts_e += e2
        parts_n += n2
    assert torch.allclose(parts_e, err2, rtol=1e-9) and torch.allclose(parts_n, norm2, rtol=1e-9)
    rel = float(torch.sqrt(err2.sum()) / torch.sqrt(norm2.sum()))
    assert abs(rel - 0.01 / (1.0 + 0.01 ** 2) ** 0.5) < 1e-3     # |noise| / |A| with unit-variance Q E
    # the tensor-core form (squared-error epilogue of the fused kernel, fp32-grade 3-pass) agrees with the FFMA form
    e2p, n2p = eng.recon_error_packed(Qm, eng.PackedItems(E, "f32x3"), A)
    assert torch.allclose(e2p, err2, rtol=1e-4) and torch.allclose(n2p, norm2, rtol=1e-6)
    # rows whose A is exactly Q E reconstruct to fp32 rounding
    A[:64] = Qm[:64] @ E
    e2, n2 = eng.recon_error_rows(Qm[:64], E, A[:64])
    assert float(torch.sqrt(e2.sum() / n2.sum())) < 1e-5


def test_c3_size_adaptive_rounds_against_fp64_pinv(eng):
    """BASELINE configs[2] at its own block sizes (k_q = 500 anchor queries, 4 rounds x 125 anchors, top-100) on N = 60 000
    items, for a handful of queries the CPU can follow: every round's picks against the fp64 statement of that round on the
    same anchor set (numpy pinv of the 500 x m matrix), exact outside the tie band tau = 1e-4 max|s|; and the batch answer
    does not depend on which other queries share the batch (the solver's state is per query)."""
    from anncur_b200 import adaptive_anncur
    from anncur_b200.adaptive import AdaptiveIndex
    from oracle import cur_oracle as O
    k_q, N, B, T, kpr, top_k = 500, 60_000, 5, 4, 125, 100
    A = O.synthetic_scores(k_q + 64, N, rank=64, noise=0.05, seed=21)
    R, X = A[:k_q], A[k_q:]
    first = np.sort(np.random.default_rng(3).choice(N, kpr, replace=False))
    index = AdaptiveIndex(torch.from_numpy(R).cuda())
    anc_all, idx_all, val_all = adaptive_anncur(torch.from_numpy(R).cuda(), torch.from_numpy(X).cuda(), first, T, kpr, top_k, index=index)
    anc, idx, val = adaptive_anncur(torch.from_numpy(R).cuda(), torch.from_numpy(X[:B]).cuda(), first, T, kpr, top_k, index=index)
    assert torch.equal(anc, anc_all[:B]) and torch.equal(idx, idx_all[:B]) and torch.equal(val, val_all[:B])
    anc = anc.cpu().numpy()
    R64 = R.astype(np.float64)
    n_same = 0
    for q in range(B):
        for t in range(1, T):
            cur, got = anc[q, :t * kpr], anc[q, t * kpr:(t + 1) * kpr]
            e = X[q, cur].astype(np.float64) @ np.linalg.pinv(R64[:, cur])
            sc = e @ R64
            sc[cur] = -np.inf
            order = np.argsort(-sc, kind="stable")
            tau = 1e-4 * np.abs(sc[np.isfinite(sc)]).max()
            assert len(set(got.tolist())) == kpr and not set(got.tolist()) & set(cur.tolist())
            assert all(sc[j] >= sc[order[kpr - 1]] - tau for j in got), (q, t)
            n_same += set(got.tolist()) == set(order[:kpr].tolist())
    assert n_same >= 0.8 * B * (T - 1), n_same
    exact_top = np.argsort(-X[:B].astype(np.float64), axis=1, kind="stable")[:, :top_k]
    recall = np.mean([len(set(idx[q].cpu().tolist()) & set(exact_top[q].tolist())) / top_k for q in range(B)])
    assert recall > 0.9, recall


@pytest.mark.parametrize("B,K,N,k,kind", [
    (4225, 96, 20_000, 37, "f32r"),        # 34 query tiles: odd tile count for the CTA pairs, last tile holds one row
    (8192, 500, 100_000, 100, "f32r"),     # two bench batches in one call
    (5000, 1100, 30_000, 64, "f32r"),      # K > 1024
    (4500, 300, 50_000, 200, "f32x3"),
    (1, 500, 2_000_000, 100, "f32r"),      # one query, two million items (HBM-bound regime)
    (6000, 64, 3_000, 1000, "f32r"),       # k_r = 1000 of a small item set: streamed from -inf
])
def test_large_batches_against_device_fp32_reference(eng, B, K, N, k, kind):
    """Shapes the CPU oracle cannot follow in seconds, against torch's fp32 matmul (TF32 off) + topk on the device: returned
    scores within the kind's tolerance of the dense scores of the returned items, every returned item a top-k item up to the
    tie band 1e-4 max|s|, indices unique, best first."""
    g = torch.Generator(device="cuda").manual_seed(B + N)
    r = 24
    W = torch.randn(K, r, device="cuda", generator=g)
    E = W @ torch.randn(r, N, device="cuda", generator=g) / r ** 0.5 + 0.05 * torch.randn(K, N, device="cuda", generator=g)
    Q = torch.randn(B, r, device="cuda", generator=g) @ W.t() / r ** 0.5 + 0.05 * torch.randn(B, K, device="cuda", generator=g)
    v, i = eng.score_topk(Q, eng.PackedItems(E, kind), k)
    torch.backends.cuda.matmul.allow_tf32 = False
    for b0 in range(0, B, 2048):
        dense = (Q[b0:b0 + 2048].double() @ E.double()) if K * N <= 40_000_000 else (Q[b0:b0 + 2048] @ E).double()
        vv, ii = v[b0:b0 + 2048].double(), i[b0:b0 + 2048]
        scale = dense.abs().amax(dim=1, keepdim=True)
        got = torch.gather(dense, 1, ii)
        assert float(((vv - got).abs() / scale).max()) <= (2e-5 if kind == "f32r" and K <= 700 else 1e-4)
        kth = torch.topk(dense, k, dim=1).values[:, -1:]
        assert bool((got >= kth - 1e-4 * scale).all())
        assert bool((vv[:, :-1] >= vv[:, 1:]).all())
        srt = torch.sort(ii, dim=1).values
        assert bool((srt[:, 1:] != srt[:, :-1]).all())
