"""GPU parity of the fused tcgen05 score + top-k kernel (anncur_score_topk) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import cur_oracle as O
from tests.parity import assert_scores_close, assert_sorted_desc, assert_topk_sets_match

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from anncur_b200 import engine
    engine.require_cuda()
    return engine


def _rand(shape, seed, scale=1.0):
    return torch.from_numpy(np.random.default_rng(seed).standard_normal(shape, dtype=np.float32) * np.float32(scale))


def _check(eng, Q, E, k, kind="f32x3", rel=1e-4, offset=0):
    packed = eng.PackedItems(E.cuda(), kind)
    v, i = eng.score_topk(Q.cuda(), packed, k, idx_offset=offset)
    torch.cuda.synchronize()
    v, i = v.cpu().numpy(), i.cpu().numpy() - offset
    dense = (Q.double() @ E.double()).numpy()
    ref = O.score_topk(Q, E, min(k, E.shape[1]))
    kk = ref.indices.shape[1]
    assert_topk_sets_match(i[:, :kk], ref.indices.numpy(), full_scores=dense, rel=rel)
    assert_scores_close(v[:, :kk], np.take_along_axis(dense, i[:, :kk], 1), rel=rel, what="scores of returned items")
    assert_sorted_desc(v[:, :kk])
    if k > kk:
        assert (i[:, kk:] == -1 - offset).all()
    return v, i


@pytest.mark.parametrize("B,K,N,k", [
    (128, 32, 256, 10),         # exactly one tile, one k-block
    (1, 50, 10000, 100),        # single query (HBM-bound regime), ragged K
    (7, 500, 3000, 100),        # ragged everything
    (130, 64, 777, 64),         # two query tiles, last one nearly empty; N tail inside a tile
    (300, 200, 20000, 1),       # k = 1
    (64, 500, 50000, 100),
    (256, 96, 4096, 128),       # k at a capacity boundary (cap 256)
    (33, 40, 9000, 129),        # cap 512
    (20, 128, 30000, 500),      # cap 1024
    (12, 64, 20000, 1000),      # cap 2048 (k_r = 1000 is the reference's largest retrieval size)
])
def test_fused_f32x3_matches_oracle(eng, B, K, N, k):
    _check(eng, _rand((B, K), B + K), _rand((K, N), N + k), k)


# ---- kind "f32r": one f16 filter pass of upper bounds + exact fp32 re-scoring of the candidates ----------------------
# Its values are plain fp32 dot products, so the tolerance is 1e-5 (of the row's max |score|) instead of 1e-4.
@pytest.mark.parametrize("B,K,N,k", [
    (128, 32, 256, 10),         # no SAMPLE pass at this size: 3-pass streaming on the F32R planes (extra k-block: K % 32 == 0)
    (1, 50, 10000, 100),
    (7, 500, 3000, 100),
    (130, 64, 60000, 64),       # K % 32 == 0 -> the bound slot opens a new k-block; re-score with 1 float4 per lane
    (300, 200, 50000, 1),       # k = 1; 2 float4 per lane
    (64, 500, 50000, 100),      # the bench shape; 4 float4 per lane
    (257, 512, 70000, 100),     # K = 512 -> 17 k-blocks
    (40, 700, 40000, 50),       # K > 512: query planes not register-cached, re-score through the long path
    (24, 2000, 40000, 100),     # k_i = 2000, the largest anchor count of the reference's grid (..._fixed...:249-251)
    (16, 4096, 48000, 100),     # K a multiple of 64 far above the bench shape
    (33, 40, 90000, 129),
    (20, 128, 60000, 500),
    (12, 64, 120000, 1000),     # k_r = 1000, the reference's largest retrieval size
    (260, 500, 100000, 1000),   # k_r = 1000 at the C2 index: sampled with the 2 j rule (stride 8), served by the 3-pass launch
    (64, 200, 120000, 250),     # k = 250: the refine kernel holds 2048 candidates per row
    (40, 128, 200000, 450),     # k = 450: 4096 candidates per row
    (300, 96, 130000, 300),
])
def test_fused_f32r_matches_oracle(eng, B, K, N, k):
    # plain fp32 dot products: 1e-5 of the row's max |score| up to K ~ 700; beyond, the fp32 summation error of ANY fp32
    # evaluation (the reference's included) approaches 1e-5 and the north_star tolerance 1e-4 applies
    _check(eng, _rand((B, K), B + K), _rand((K, N), N + k), k, kind="f32r", rel=1e-5 if K <= 700 else 1e-4)


def test_fused_f32r_scale_extremes_and_offsets(eng):
    for qs, es in [(15.0, 15.0), (1e-6, 1e-3), (3e4, 2e3)]:
        _check(eng, _rand((40, 100), 3, qs), _rand((100, 40000), 4, es), 50, kind="f32r", rel=1e-5)
    Q = _rand((64, 100), 5)
    Q *= torch.logspace(-6, 6, 64).unsqueeze(1)
    _check(eng, Q, _rand((100, 40000), 6), 50, kind="f32r", rel=1e-5, offset=2**40)
    # items of wildly different norms: the per-item bound b_n differs by orders of magnitude inside one tile
    E = _rand((100, 40000), 7) * torch.logspace(-4, 2, 40000).unsqueeze(0)[:, torch.randperm(40000, generator=torch.Generator().manual_seed(0))]
    _check(eng, _rand((50, 100), 8), E, 100, kind="f32r", rel=1e-5)


def test_fused_f32r_bounds_enclose_the_fp32_score(eng):
    """The guarantee kind f32r rests on: lower <= fp32 score <= upper for EVERY (query, item) pair, where lower / upper are
    what its SAMPLE / MAIN launches compare with the thresholds (anncur_score_bounds_dense).  Checked against fp64 on random
    data, on per-row / per-item magnitudes spread over 8 decades, on a low-rank CUR index (cancelling terms), and on operands
    built so that every fp16 rounding error has the same sign and nearly the worst size (x = 2^e (1 + 2^-11 + 2^-13) rounds
    up by 3/4 of a half-ulp; all products positive) -- the coherent case a statistical error model would miss."""
    rng = np.random.default_rng(17)
    cases = []
    Q, E = _rand((200, 500), 61), _rand((500, 20000), 62)
    cases.append(("random", Q, E))
    cases.append(("magnitudes", Q * torch.logspace(-4, 4, 200).unsqueeze(1), E * torch.logspace(-4, 4, 20000).unsqueeze(0)))
    A = torch.from_numpy(O.synthetic_scores(400, 20000, rank=64, noise=0.05, seed=3))
    rows_i, cols_i = O.sample_anchors(400, 20000, 150, 300, 3)
    f = O.cur_build(A[rows_i, :], A[:, cols_i], rows_i, cols_i, "rows", check=False)
    cases.append(("cur", A[:, cols_i].contiguous(), f.latent_cols.contiguous()))
    worst = np.float32(1.0 + 2.0 ** -11 + 2.0 ** -13)
    Qw = torch.from_numpy((worst * 2.0 ** rng.integers(-3, 4, (130, 512))).astype(np.float32))
    Ew = torch.from_numpy((worst * 2.0 ** rng.integers(-3, 4, (512, 9000))).astype(np.float32))
    cases.append(("coherent worst case", Qw, Ew))
    cases.append(("coherent, mixed signs", Qw * torch.from_numpy(rng.choice([-1.0, 1.0], (130, 512)).astype(np.float32)), -Ew))
    for name, Q, E in cases:
        packed = eng.PackedItems(E.cuda(), "f32r")
        ub = eng.score_bounds_dense(Q.cuda(), packed, +1).cpu().double()
        lb = eng.score_bounds_dense(Q.cuda(), packed, -1).cpu().double()
        S = Q.double() @ E.double()
        S32 = (Q.cuda() @ E.cuda()).cpu().double()            # an fp32 evaluation (cuBLAS), for the "fp32 score" side of the claim
        slack = (ub - lb) / 2                                  # = b_n
        assert (ub >= S).all() and (lb <= S).all(), name
        assert (ub >= S32).all() and (lb <= S32).all(), name
        used = ((S - (ub + lb) / 2).abs() / slack.clamp_min(1e-300)).max().item()
        assert used <= 0.97, (name, used)                      # the bound is never exhausted (the 5 % reserve stays)
        if name == "coherent worst case":
            assert used > 0.2, used                            # ... and this case really does stress it


@pytest.mark.parametrize("K", [768, 2000, 4096, 8192])
def test_fused_f32r_bounds_enclose_at_large_k(eng, K):
    """The same enclosure at the anchor counts the header allows for kind F32R (ANNCUR_MAX_K_DIM_F32R = 8192) -- the
    fp32 accumulation error of the tensor pipe grows with K, and so does the reserve in the bound slot (f32r_slot_factor).
    Coherent worst case: every fp16 rounding error has the same sign and all products are positive, so truncating adders
    would drift one way.  Also random data.  The bound must hold with room to spare at every K."""
    rng = np.random.default_rng(K)
    worst = np.float32(1.0 + 2.0 ** -11 + 2.0 ** -13)
    n_q, n_e = 96, 6000
    cases = [("coherent worst case", torch.from_numpy((worst * 2.0 ** rng.integers(-3, 4, (n_q, K))).astype(np.float32)),
              torch.from_numpy((worst * 2.0 ** rng.integers(-3, 4, (K, n_e))).astype(np.float32))),
             ("coherent, equal magnitudes", torch.full((n_q, K), float(worst)), torch.full((K, n_e), float(worst))),
             ("random", _rand((n_q, K), K + 1), _rand((K, n_e), K + 2))]
    for name, Q, E in cases:
        packed = eng.PackedItems(E.cuda(), "f32r")
        ub = eng.score_bounds_dense(Q.cuda(), packed, +1).cpu().double()
        lb = eng.score_bounds_dense(Q.cuda(), packed, -1).cpu().double()
        S = Q.double() @ E.double()
        S32 = (Q.cuda() @ E.cuda()).cpu().double()
        slack = (ub - lb) / 2
        assert (ub >= S).all() and (lb <= S).all(), (name, K)
        assert (ub >= S32).all() and (lb <= S32).all(), (name, K)
        used = ((S - (ub + lb) / 2).abs() / slack.clamp_min(1e-300)).max().item()
        assert used <= 0.9, (name, K, used)                   # reserve > 0 at every K
        # the one-pass score itself (midpoint of the two bounds) against fp64: what the tensor pipe's accumulation adds on top
        # of the operand rounding stays far inside the K-proportional reserve
        print(f"K={K} {name}: bound used {used:.3f}")


def test_fused_f32r_topk_parity_at_k_dim_8192(eng):
    """Top-k through kind F32R at the largest anchor count the header allows, against the oracle."""
    _check(eng, _rand((9, 8192), 5), _rand((8192, 20000), 6), 50, kind="f32r", rel=1e-4)


def test_fused_f32r_fast_path_serves_generic_inputs(eng):
    """Generic inputs must not lean on the fallback: no row of a random / low-rank batch goes to REDO, while the
    adversarial batch (all-equal scores) is entirely recomputed there."""
    K, N, B, k = 500, 100000, 512, 100
    Q, E = _rand((B, K), 41), _rand((K, N), 42)
    packed = eng.PackedItems(E.cuda(), "f32r")
    eng.score_topk(Q.cuda(), packed, k)
    assert eng.last_redo_rows(B, packed, k) == 0
    A = torch.from_numpy(O.synthetic_scores(600, 60000, rank=64, noise=0.05, seed=1))
    rows_i, cols_i = O.sample_anchors(600, 60000, 100, 200, 1)
    f = O.cur_build(A[rows_i, :], A[:, cols_i], rows_i, cols_i, "rows", check=False)
    packed = eng.PackedItems(f.latent_cols.cuda(), "f32r")
    eng.score_topk(A[:, cols_i].contiguous().cuda(), packed, k)
    assert eng.last_redo_rows(600, packed, k) == 0
    packed = eng.PackedItems(torch.ones(32, 60000).cuda(), "f32r")
    eng.score_topk(torch.ones(130, 32).cuda(), packed, 10)
    assert eng.last_redo_rows(130, packed, 10) == 130


def test_fused_f32r_exact_on_integer_data_with_ties(eng):
    """Small-integer operands: every fp32 dot product is exact, so the answer must equal the exact top-k with ties
    broken towards the lower index -- bit for bit, values and indices."""
    rng = np.random.default_rng(3)
    B, K, N, k = 200, 96, 70000, 100
    Q = torch.from_numpy(rng.integers(-3, 4, (B, K)).astype(np.float32))
    E = torch.from_numpy(rng.integers(-3, 4, (K, N)).astype(np.float32))
    v, i = eng.score_topk(Q.cuda(), eng.PackedItems(E.cuda(), "f32r"), k)
    S = (Q.double() @ E.double()).numpy()
    order = np.lexsort((np.broadcast_to(np.arange(N), S.shape), -S), axis=1)[:, :k]
    assert (i.cpu().numpy() == order).all()
    assert (v.cpu().numpy() == np.take_along_axis(S, order, 1).astype(np.float32)).all()


def test_fused_f32r_adversarial_inputs_fall_back(eng):
    """Inputs built to defeat the filter: ascending scores and all-equal scores (every item passes the threshold ->
    lists fill up -> REDO), a planted sample miss (rows end short -> REDO), and thousands of exact ties at the top."""
    K, N = 32, 60000
    E = torch.zeros(K, N)
    E[0] = torch.arange(N, dtype=torch.float32) / N
    Q = torch.zeros(70, K)
    Q[:, 0] = torch.linspace(0.5, 2.0, 70)
    v, i = _check(eng, Q, E, 100, kind="f32r", rel=1e-5)
    assert (i == np.arange(N - 1, N - 101, -1)[None, :]).all()
    packed = eng.PackedItems(torch.ones(K, N).cuda(), "f32r")
    v, i = eng.score_topk(torch.ones(3, K).cuda(), packed, 10)
    assert (i.cpu().numpy() == np.arange(10)[None, :]).all() and np.allclose(v.cpu().numpy(), K)
    E = _rand((K, N), 7)
    E[0, ::16] += 50.0
    Q = _rand((300, K), 8)
    Q[:, 0] = 0.0
    Q[[5, 131, 299], 0] = 1.0
    _check(eng, Q, E, 100, kind="f32r", rel=1e-5)
    Q, E = _rand((150, 64), 11), _rand((64, 80000), 12)
    E[:, 1000:9000] = E[:, 999:1000]
    _check(eng, Q, E, 100, kind="f32r", rel=1e-5)


def test_fused_f32r_search_host_and_c2_sample(eng):
    """Host-buffer entry point == device entry point for kind f32r; C2-sized call checked on a row sample."""
    K, N, B, k = 500, 100000, 4096, 100
    Q, E = _rand((B, K), 21), _rand((K, N), 22)
    packed = eng.PackedItems(E.cuda(), "f32r")
    v, i = eng.score_topk(Q.cuda(), packed, k)
    rows = np.random.default_rng(0).choice(B, 96, replace=False)
    dense = (Q[rows].double() @ E.double()).numpy()
    ref = torch.topk(torch.from_numpy(dense), k, dim=1)
    assert_topk_sets_match(i[rows].cpu().numpy(), ref.indices.numpy(), full_scores=dense, rel=1e-5)
    assert_scores_close(v[rows].cpu().numpy(), ref.values.numpy(), rel=1e-5)
    vh = torch.empty((B, k), dtype=torch.float32).pin_memory()
    ih = torch.empty((B, k), dtype=torch.int64).pin_memory()
    eng.search_host(Q.pin_memory(), packed, k, vh, ih)
    torch.cuda.synchronize()
    assert torch.equal(ih, i.cpu()) and torch.equal(vh, v.cpu())


def test_score_topk_captured_as_cuda_graph(eng):
    """The whole call (pack, SAMPLE, threshold, MAIN, refine, REDO, select) is capturable: a replayed graph gives the
    eager answer on fresh queries, for the sampled fast path and for the small unsampled one."""
    for (B, K, N, k, kind) in [(64, 500, 60000, 100, "f32r"), (64, 96, 3000, 10, "f32r"), (200, 64, 70000, 20, "bf16")]:
        E = _rand((K, N), 51)
        packed = eng.PackedItems(E.cuda(), kind)
        g = eng.GraphedSearch(packed, B, k, idx_offset=5)
        for seed in (52, 53):
            Q = _rand((B, K), seed).cuda()
            gv, gi = g(Q)
            ev, ei = eng.score_topk(Q, packed, k, idx_offset=5)
            torch.cuda.synchronize()
            assert torch.equal(gi, ei) and torch.equal(gv, ev)


def test_fused_k_larger_than_items_pads(eng):
    _check(eng, _rand((5, 16), 1), _rand((16, 40), 2), 64)


def test_fused_index_offset_is_int64(eng):
    _check(eng, _rand((3, 16), 1), _rand((16, 500), 2), 5, offset=2**40)


def test_fused_scale_extremes(eng):
    # CE-logit scale (+-15), tiny and huge magnitudes, and a mixed-magnitude batch (per-row query scale)
    for qs, es in [(15.0, 15.0), (1e-6, 1e-3), (3e4, 2e3)]:
        _check(eng, _rand((40, 100), 3, qs), _rand((100, 6000), 4, es), 50)
    Q = _rand((64, 100), 5)
    Q *= torch.logspace(-6, 6, 64).unsqueeze(1)
    _check(eng, Q, _rand((100, 6000), 6), 50)


def test_fused_adversarial_ascending_scores(eng):
    # scores increase with the item index: every element beats the running threshold (worst case for the
    # survivor lists, exercises repeated compaction); and all-equal scores: pure index tie-break
    K, N = 32, 20000
    E = torch.zeros(K, N)
    E[0] = torch.arange(N, dtype=torch.float32) / N
    Q = torch.zeros(70, K)
    Q[:, 0] = torch.linspace(0.5, 2.0, 70)
    v, i = _check(eng, Q, E, 100)
    assert (i == np.arange(N - 1, N - 101, -1)[None, :]).all()
    packed = eng.PackedItems(torch.ones(K, N).cuda(), "f32x3")
    v, i = eng.score_topk(torch.ones(3, K).cuda(), packed, 10)
    assert (i.cpu().numpy() == np.arange(10)[None, :]).all() and np.allclose(v.cpu().numpy(), K)


def test_fused_sampled_threshold_miss_is_redone(eng):
    # N large enough for the SAMPLE pass (every 16th item).  Items 0, 16, 32, ... score far above the rest, so
    # the sampled threshold admits only ~j < k items: every row comes up short in MAIN and must be recovered by
    # the REDO pass.  Also a second batch where only a few rows are hit (mixed flagged / unflagged query tiles).
    K, N, B, k = 32, 60000, 300, 100
    E = _rand((K, N), 7)
    Q = _rand((B, K), 8)
    E2 = E.clone()
    E2[0, ::16] += 50.0
    Q2 = Q.clone()
    Q2[:, 0] = 1.0
    _check(eng, Q2, E2, k)
    Q3 = Q.clone()
    Q3[:, 0] = 0.0
    Q3[5, 0] = 1.0
    Q3[131, 0] = 1.0
    Q3[299, 0] = 1.0
    _check(eng, Q3, E2, k)


def test_fused_sampled_heavy_ties_and_small_k(eng):
    # sampled path with k = 1 / 10 and with massive ties at the top (threshold admits thousands: list overflow ->
    # in-kernel compaction), on a size where SAMPLE is active
    K, N, B = 64, 80000, 150
    Q, E = _rand((B, K), 11), _rand((K, N), 12)
    _check(eng, Q, E, 1)
    _check(eng, Q, E, 10)
    E[:, 1000:9000] = E[:, 999:1000]                  # 8001 identical items
    v, i = _check(eng, Q, E, 100)


def test_fused_zero_anchor_items(eng):
    # k_i = 0 is in the reference's grid: all approximate scores are 0
    packed = eng.PackedItems(torch.zeros(0, 50).cuda(), "f32x3")
    v, i = eng.score_topk(torch.zeros(4, 0).cuda(), packed, 5)
    assert (v == 0).all() and i[0].tolist() == [0, 1, 2, 3, 4]


def test_fused_bf16_recall(eng):
    Q, E = _rand((200, 500), 1), _rand((500, 30000), 2)
    packed = eng.PackedItems(E.cuda(), "bf16")
    v, i = eng.score_topk(Q.cuda(), packed, 100)
    ref = O.score_topk(Q, E, 100)
    i = i.cpu().numpy()
    recall = np.mean([len(set(i[r]) & set(ref.indices[r].tolist())) / 100 for r in range(200)])
    assert recall > 0.9, recall                                    # reported, not a parity claim
    dense = (Q @ E).numpy()
    assert_scores_close(v.cpu().numpy(), np.take_along_axis(dense, i, 1), rel=2e-2)


def test_fused_c1_config_against_oracle(eng):
    """BASELINE config 1: 1k queries x 10k items, k_q = k_i = 50, top-100 (the reference's CPU-runnable case)."""
    A = torch.from_numpy(O.synthetic_scores(1000, 10000, rank=64, noise=0.05, seed=0))
    rows_i, cols_i = O.sample_anchors(1000, 10000, 50, 50, 0)
    f = O.cur_build(A[rows_i, :], A[:, cols_i], rows_i, cols_i, "rows", check=False)
    Q = A[:, cols_i].contiguous()
    _check(eng, Q, f.latent_cols, 100)


def test_fused_c2_config_sharded_merge_property(eng):
    """BASELINE config 2 size (N = 100k, k_i = 500, B = 4096): oracle on a row sample, and the
    size-independent property that an item-sharded search merged by anncur_merge_topk equals the
    single-shard answer."""
    K, N, B, k = 500, 100000, 4096, 100
    Q, E = _rand((B, K), 21), _rand((K, N), 22)
    Ec = E.cuda()
    packed = eng.PackedItems(Ec, "f32x3")
    v, i = eng.score_topk(Q.cuda(), packed, k)
    rows = np.random.default_rng(0).choice(B, 96, replace=False)
    dense = (Q[rows].double() @ E.double()).numpy()
    ref = torch.topk(torch.from_numpy(dense), k, dim=1)
    assert_topk_sets_match(i[rows].cpu().numpy(), ref.indices.numpy(), full_scores=dense)
    assert_scores_close(v[rows].cpu().numpy(), ref.values.numpy())
    P = 4
    cv, ci = [], []
    for p in range(P):
        lo, hi = p * N // P, (p + 1) * N // P
        pv, pi = eng.score_topk(Q.cuda(), eng.PackedItems(Ec[:, lo:hi], "f32x3"), k, idx_offset=lo)
        cv.append(pv), ci.append(pi)
    mv, mi = eng.merge_topk(torch.cat(cv, 1), torch.cat(ci, 1), k)
    assert torch.equal(mi, i) and torch.allclose(mv, v, rtol=1e-5, atol=1e-5)


def test_search_host_equals_device_path(eng):
    """anncur_search_host (host buffers: H2D, kernels, D2H on the stream) == anncur_score_topk (device buffers)."""
    Q, E = _rand((300, 77), 31), _rand((77, 70000), 32)
    packed = eng.PackedItems(E.cuda(), "f32x3")
    dv, di = eng.score_topk(Q.cuda(), packed, 50, idx_offset=7)
    for pin in (False, True):
        Qh = torch.zeros(300, 80)                         # row stride 80 > k_dim 77
        Qh[:, :77] = Q
        Qh = Qh.pin_memory() if pin else Qh
        vh = torch.empty((300, 50), dtype=torch.float32)
        ih = torch.empty((300, 50), dtype=torch.int64)
        if pin:
            vh, ih = vh.pin_memory(), ih.pin_memory()
        eng.search_host(Qh[:, :77], packed, 50, vh, ih, idx_offset=7)
        torch.cuda.synchronize()
        assert torch.equal(ih, di.cpu()) and torch.equal(vh, dv.cpu())


@pytest.mark.parametrize("group", ["1", "2"])
def test_fused_both_cta_group_modes_in_subprocess(group):
    """The CTA-pair (cta_group::2) and single-CTA instantiations of the fused kernel give the oracle's answer;
    the mode is chosen per process (ANNCUR_CTA_GROUP), hence the subprocess."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import numpy as np, torch, sys
sys.path.insert(0, %r)
from anncur_b200 import engine as eng
from oracle import cur_oracle as O
from tests.parity import assert_scores_close, assert_topk_sets_match
rng = np.random.default_rng(5)
for (B, K, N, k, kind) in [(700, 96, 90000, 100, "f32x3"), (129, 500, 3000, 10, "f32x3"), (513, 64, 70000, 40, "bf16"),
                           (700, 96, 90000, 100, "f32r"), (100, 500, 60000, 10, "f32r")]:
    Q = torch.from_numpy(rng.standard_normal((B, K), dtype=np.float32)); E = torch.from_numpy(rng.standard_normal((K, N), dtype=np.float32))
    v, i = eng.score_topk(Q.cuda(), eng.PackedItems(E.cuda(), kind), k)
    dense = (Q.double() @ E.double()).numpy(); ref = O.score_topk(Q, E, k)
    if kind != "bf16":
        assert_topk_sets_match(i.cpu().numpy(), ref.indices.numpy(), full_scores=dense)
        assert_scores_close(v.cpu().numpy(), np.take_along_axis(dense, i.cpu().numpy(), 1))
    else:
        rec = np.mean([len(set(i[r].tolist()) & set(ref.indices[r].tolist())) / k for r in range(B)])
        assert rec > 0.9, rec
print("OK")
""" % root
    env = dict(os.environ, ANNCUR_CTA_GROUP=group)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]


# ---- randomised shapes: every kind against the oracle on shapes nobody picked by hand ---------------------------------------
def _sweep_shapes(seed, n):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        B = int(rng.choice([1, 2, 5, 31, 64, 100, 129, 200, 257, 300]))
        K = int(rng.integers(1, 260))
        N = int(np.exp(rng.uniform(np.log(1), np.log(150_000))))
        k = int(min(rng.choice([1, 2, 7, 10, 50, 100, 128, 129, 300, 700]), max(N + 3, 1)))     # k > N happens too (faiss padding)
        out.append((B, K, N, k))
    return out


@pytest.mark.parametrize("kind,rel", [("f32r", 1e-5), ("f32x3", 1e-4)])
def test_fused_random_shapes_match_oracle(eng, kind, rel):
    """36 random (B, K, N, k) per kind, N from 1 to 150 000, on gaussian data and on a CUR-like low-rank index (scores with
    a few dominant directions, as E = U R has); the same parity rules as everywhere else."""
    for n, (B, K, N, k) in enumerate(_sweep_shapes(123 if kind == "f32r" else 321, 36)):
        if n % 2 == 0:
            Q, E = _rand((B, K), 1000 + n), _rand((K, N), 2000 + n)
        else:
            r = max(1, min(K, 8))
            g = np.random.default_rng(3000 + n)
            W = g.standard_normal((K, r), dtype=np.float32)
            E = torch.from_numpy(W @ g.standard_normal((r, N), dtype=np.float32) / np.float32(np.sqrt(r)) + np.float32(0.05) * g.standard_normal((K, N), dtype=np.float32))
            Q = torch.from_numpy(g.standard_normal((B, r), dtype=np.float32) @ W.T / np.float32(np.sqrt(r)) + np.float32(0.05) * g.standard_normal((B, K), dtype=np.float32))
        try:
            # tolerances are relative to the row's largest |score|: with a handful of items that "largest" can itself be a
            # cancelled sum, so tiny item sets are held to north_star's 1e-4 instead of the kind's own 1e-5
            _check(eng, Q, E, k, kind=kind, rel=rel if N >= 64 else 1e-4, offset=int(n % 3) * 1_000_003)
        except AssertionError as exc:
            raise AssertionError(f"shape B={B} K={K} N={N} k={k} (case {n}, {'gaussian' if n % 2 == 0 else 'low-rank'}): {exc}") from exc


def test_fused_bf16_random_shapes_recall(eng):
    """bf16 is reported as recall@k against the exact result (north_star): on random shapes the recall stays high, and the
    returned values are the products of the bf16-ROUNDED operands accumulated in fp32 (error against those: fp32 summation
    only, bounded per row by 1e-5 sum|q| max|e|; against the unrounded operands ~2^-8 of sqrt(K))."""
    for n, (B, K, N, k) in enumerate(_sweep_shapes(77, 16)):
        Q, E = _rand((B, K), 500 + n), _rand((K, N), 700 + n)
        v, i = eng.score_topk(Q.cuda(), eng.PackedItems(E.cuda(), "bf16"), k)
        v, i = v.cpu().numpy(), i.cpu().numpy()
        dense = (Q.double() @ E.double()).numpy()
        Qb, Eb = Q.bfloat16().double(), E.bfloat16().double()
        dense_b = (Qb @ Eb).numpy()
        kk = min(k, N)
        ref = np.argsort(-dense, axis=1, kind="stable")[:, :kk]
        recall = np.mean([len(set(i[r, :kk].tolist()) & set(ref[r].tolist())) / kk for r in range(B)])
        assert recall > 0.9 or kk <= 2, (B, K, N, k, recall)
        assert all(len(set(i[r, :kk].tolist())) == kk for r in range(B))
        bound = 1e-5 * (Qb.abs().sum(1, keepdim=True) * Eb.abs().max()).numpy() + 1e-12
        err = np.abs(v[:, :kk] - np.take_along_axis(dense_b, i[:, :kk], 1))
        assert (err <= bound).all(), (B, K, N, k, float((err / bound).max()))
        assert np.abs(v[:, :kk] - np.take_along_axis(dense, i[:, :kk], 1)).max() <= 2.0 ** -6 * np.sqrt(K) * 4, (B, K, N, k)
        assert_sorted_desc(v[:, :kk])
        if k > kk:
            assert (i[:, kk:] == -1).all()


@pytest.mark.parametrize("B,K,N,k,m,kind", [(70, 64, 9000, 12, 24, "f32r"), (33, 200, 60000, 125, 375, "f32r"), (5, 40, 300, 100, 150, "f32r"),
                                            (130, 96, 20000, 50, 1, "f32x3"), (9, 32, 5000, 10, 0, "f32r")])
def test_masked_search_matches_oracle(eng, B, K, N, k, m, kind):
    """anncur_score_topk_excluding: the k best items of a row that are not in the row's excluded list (the masked re-score of
    the adaptive rounds) against the dense scores with the excluded items set to -inf."""
    Q, E = _rand((B, K), 70 + B), _rand((K, N), 80 + k)
    rng = np.random.default_rng(B + m)
    dense = (Q.double() @ E.double()).numpy()
    order = np.argsort(-dense, axis=1, kind="stable")
    # half of the excluded items come from the row's own top (they would otherwise win), half are random; one padding entry
    excl = np.zeros((B, 0), dtype=np.int64)
    if m > 0:
        excl = np.stack([np.concatenate([order[r, :m // 2], rng.choice(order[r, m // 2:], m - m // 2, replace=False)]) for r in range(B)]).astype(np.int64)
    if m > 2:
        excl[:, -1] = -1
    v, i = eng.score_topk_excluding(Q.cuda(), eng.PackedItems(E.cuda(), kind), k, torch.from_numpy(excl))
    v, i = v.cpu().numpy(), i.cpu().numpy()
    masked = dense.copy()
    for r in range(B):
        masked[r, excl[r][excl[r] >= 0]] = -np.inf
    ref = np.argsort(-masked, axis=1, kind="stable")[:, :k]
    finite = np.where(np.isfinite(masked), masked, 0.0)
    for r in range(B):
        assert not set(i[r].tolist()) & set(excl[r][excl[r] >= 0].tolist())
    assert_topk_sets_match(i, ref, full_scores=np.where(np.isfinite(masked), masked, -1e30), rel=1e-5 if kind == "f32r" else 1e-4)
    assert_scores_close(v, np.take_along_axis(finite, i, 1), rel=1e-5 if kind == "f32r" else 1e-4)
    assert_sorted_desc(v)
