"""GPU parity: every kernel family through the C ABI (anncur_b200.engine) against the CPU oracle /
plain torch fp32 on the same seeded inputs.  Integer/index results are compared with the near-tie
rule of tests/parity.py; floating point with the 1e-4 relative tolerance of BASELINE.json."""
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


def _rand(shape, seed):
    return torch.from_numpy(np.random.default_rng(seed).standard_normal(shape, dtype=np.float32))


# ---------------------------------------------------------------------------------------------- top-k
@pytest.mark.parametrize("n,N,k", [(1, 1, 1), (3, 10, 10), (5, 1000, 7), (17, 4096, 100), (4, 5000, 1000),
                                   (2, 70000, 64), (3, 200000, 2048)])
def test_topk_rows_matches_torch(eng, n, N, k):
    S = _rand((n, N), n * 1000 + k)
    v, i = eng.topk_rows(S.cuda(), k)
    ref = torch.topk(S, k, dim=1)
    assert torch.equal(v.cpu(), ref.values)              # bit-exact values: pure selection
    assert_topk_sets_match(i.cpu().numpy(), ref.indices.numpy(), full_scores=S.numpy())
    assert_sorted_desc(v.cpu().numpy())
    assert torch.equal(torch.gather(S, 1, i.cpu()), v.cpu())


def test_topk_rows_ties_lowest_index_first_and_padding(eng):
    S = torch.zeros(2, 6000)
    S[0, [5, 900, 4500]] = 1.0
    v, i = eng.topk_rows(S.cuda(), 6)
    assert i[0].tolist() == [5, 900, 4500, 0, 1, 2] and i[1].tolist() == [0, 1, 2, 3, 4, 5]
    v, i = eng.topk_rows(torch.tensor([[3.0, 1.0, 2.0]]).cuda(), 5)          # k > N: faiss-style padding
    assert i[0].tolist() == [0, 2, 1, -1, -1]
    assert v[0, 3].item() == -np.finfo(np.float32).max
    v, i = eng.topk_rows(S[:, ::7].cuda(), 3)                                 # strided rows (lds != n_cols)
    assert v.shape == (2, 3)
    v2, i2 = eng.topk_rows(S.cuda(), 4, idx_offset=10**10)                    # int64 offsets survive
    assert i2[1].tolist() == [10**10 + j for j in range(4)]


def test_merge_topk_matches_oracle(eng):
    rng = np.random.default_rng(3)
    P, B, k = 8, 37, 100
    vals = rng.standard_normal((P, B, k)).astype(np.float32)
    idx = np.stack([rng.permutation(10**6)[: B * k].reshape(B, k) + p * 10**6 for p in range(P)]).astype(np.int64)
    idx[3, :, 90:] = -1                                                        # a short shard
    want_v, want_i = O.merge_topk(vals, idx, k)
    cv = torch.from_numpy(vals).permute(1, 0, 2).reshape(B, P * k)
    ci = torch.from_numpy(idx).permute(1, 0, 2).reshape(B, P * k)
    v, i = eng.merge_topk(cv.cuda(), ci.cuda(), k)
    assert np.array_equal(i.cpu().numpy(), want_i) and np.array_equal(v.cpu().numpy(), want_v)


# ---------------------------------------------------------------------------------------------- gemm
@pytest.mark.parametrize("m,k,n", [(1, 1, 1), (50, 50, 10000), (130, 17, 257), (500, 2000, 3000), (64, 0, 33)])
def test_gemm_matches_torch(eng, m, k, n):
    A, B = _rand((m, k), 1), _rand((k, n), 2)
    C = eng.gemm(A.cuda(), B.cuda()).cpu()
    ref = A @ B
    assert C.shape == ref.shape
    if k:
        bound = (A.abs() @ B.abs()).numpy()               # |a|.|b| bounds fp32 summation-order differences
        assert (np.abs(C.numpy() - ref.numpy()) <= 1e-5 * bound + 1e-30).all()
    else:
        assert torch.count_nonzero(C) == 0


def test_gemm_strided_operands(eng):
    A, B = _rand((70, 90), 5), _rand((60, 300), 6)
    C = eng.gemm(A.cuda()[:, :60], B.cuda()[:, 10:250]).cpu()
    assert_scores_close(C.numpy(), (A[:, :60] @ B[:, 10:250]).numpy(), rel=2e-5)


# ---------------------------------------------------------------------------------------------- pinv
@pytest.mark.parametrize("m,n", [(50, 50), (200, 50), (40, 160), (1, 7), (2000, 500), (33, 1)])
def test_pinv_matches_numpy(eng, m, n):
    A = O.synthetic_scores(m, n, rank=min(m, n, 64), noise=0.05, seed=m + n)
    P, cond = eng.pinv(torch.from_numpy(A).cuda(), return_cond=True)
    P = P.cpu().numpy().astype(np.float64)
    ref = np.linalg.pinv(A.astype(np.float64), rcond=1e-15)
    s = np.linalg.svd(A.astype(np.float64), compute_uv=False)
    c = s[0] / s[-1]
    assert np.linalg.norm(P - ref) / np.linalg.norm(ref) <= max(1e-5, 10 * c * np.finfo(np.float32).eps)
    smax, smin = cond.tolist()
    assert abs(smax - s[0]) <= 1e-6 * s[0] and abs(smin - s[-1]) <= 1e-4 * s[-1] + 1e-9 * s[0]
    # Moore-Penrose identities in fp64 on our result
    A64 = A.astype(np.float64)
    assert np.linalg.norm(A64 @ P @ A64 - A64) <= 1e-4 * c * np.linalg.norm(A64) * 1e-2 + 1e-5 * np.linalg.norm(A64)


def test_pinv_drops_null_directions(eng):
    rng = np.random.default_rng(0)
    X = rng.standard_normal((60, 5))
    A = (X @ rng.standard_normal((5, 30))).astype(np.float32)      # rank 5
    P = eng.pinv(torch.from_numpy(A).cuda(), rcond=1e-5).cpu().numpy().astype(np.float64)
    ref = np.linalg.pinv(A.astype(np.float64), rcond=1e-5)
    assert np.linalg.norm(P - ref) / np.linalg.norm(ref) < 1e-4


def test_pinv_rank_deficient_anchor_columns(eng):
    """Duplicated anchor columns (exactly dependent in fp32) and a column that is an fp32-rounded sum of two others.
    With the default rcond = 1e-15 the fp64 Jacobi values of the null directions (~1e-16 s_max) must be DROPPED, not
    inverted to 1e32: the result is the minimum-norm pseudo-inverse (the cutoff is floored at what the sweeps
    resolve).  numpy's fp32 LAPACK path keeps its ~1e-7 noise values in this case, so the comparison is with the fp64
    pinv at a cutoff between the noise and the data."""
    rng = np.random.default_rng(5)
    A = rng.standard_normal((120, 40)).astype(np.float32)
    A[:, 7] = A[:, 3]                                   # exact duplicate
    A[:, 21] = A[:, 3]                                  # and again
    A[:, 30] = 2.0 * A[:, 11]                           # exact multiple (power of two: no rounding)
    P = eng.pinv(torch.from_numpy(A).cuda()).cpu().numpy().astype(np.float64)
    A64 = A.astype(np.float64)
    ref = np.linalg.pinv(A64, rcond=1e-10)
    assert np.abs(P).max() < 1e3 * np.abs(ref).max()    # nothing was inverted as 1 / noise
    assert np.linalg.norm(P - ref) / np.linalg.norm(ref) < 1e-5
    # Moore-Penrose identities
    assert np.linalg.norm(A64 @ P @ A64 - A64) < 1e-5 * np.linalg.norm(A64)
    assert np.linalg.norm(P @ A64 @ P - P) < 1e-5 * np.linalg.norm(P)
    # the same through the wide orientation (k_q < k_i)
    Pw = eng.pinv(torch.from_numpy(np.ascontiguousarray(A.T)).cuda()).cpu().numpy().astype(np.float64)
    assert np.linalg.norm(Pw - ref.T) / np.linalg.norm(ref) < 1e-5
    # all-zero input: every singular value is dropped
    assert float(eng.pinv(torch.zeros(9, 4).cuda()).abs().max()) == 0.0


def _pinv_status(eng, shape):
    import ctypes as C
    from anncur_b200 import _lib
    lib = _lib.load()
    ws = eng.WORKSPACE.get("pinv", lib.anncur_pinv_workspace_bytes(*shape), torch.device("cuda", torch.cuda.current_device()))
    st = torch.zeros(4, dtype=torch.float64, device="cuda")
    _lib.check(lib.anncur_jacobi_status(C.c_void_p(ws.data_ptr()), C.c_void_p(st.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return st.tolist()


@pytest.mark.parametrize("m,n", [(2000, 500), (500, 2000), (64, 64), (333, 97), (40, 700), (1500, 33), (2, 2)])
def test_pinv_cholesky_route_matches_numpy(eng, m, n):
    """Full-rank, well-conditioned intersections (what the CUR build sees) take the fp64 normal-equations route: Gram +
    cooperative Cholesky + triangular solves.  Same answer as np.linalg.pinv to fp32 output rounding, tall and wide; the
    status says which route ran (0 sweeps = no Jacobi)."""
    A = O.synthetic_scores(m, n, rank=min(m, n, 64), noise=0.05, seed=3 * m + n)
    if m == n:
        A = A + 3.0 * np.eye(m, dtype=np.float32)                    # keep the square case well conditioned
    P = eng.pinv(torch.from_numpy(A).cuda()).cpu().numpy().astype(np.float64)
    A64 = A.astype(np.float64)
    s = np.linalg.svd(A64, compute_uv=False)
    c = s[0] / s[-1]
    ref = np.linalg.pinv(A64, rcond=1e-15)
    assert np.linalg.norm(P - ref) / np.linalg.norm(ref) <= max(3e-7, c * c * 1e-14)
    s_max, s_min, conv, sweeps = _pinv_status(eng, (m, n))
    assert conv == 1.0
    if c < 1e3:
        assert sweeps == 0.0, "a well-conditioned input must be served by the Cholesky route"
        assert s_max <= s[0] * (1 + 1e-9) and s_min >= s[-1] * (1 - 1e-9)      # diagonal of L: inside the singular range
    # the Jacobi route on the same input agrees
    PJ = eng.pinv(torch.from_numpy(A).cuda(), return_cond=True)[0].cpu().numpy().astype(np.float64)
    assert np.linalg.norm(P - PJ) / np.linalg.norm(ref) <= max(2e-7, c * c * 1e-14)


def test_pinv_cholesky_route_refuses_hard_inputs(eng):
    """Ill-conditioned (cond 1e7) and rank-deficient inputs must fall through to the Jacobi SVD on the device -- and give
    the SVD's answer -- and a truncating rcond never takes the Cholesky route."""
    rng = np.random.default_rng(8)
    U, _ = np.linalg.qr(rng.standard_normal((400, 60)))
    V, _ = np.linalg.qr(rng.standard_normal((60, 60)))
    A = ((U * np.logspace(0, -7, 60)[None, :]) @ V.T).astype(np.float32)
    P = eng.pinv(torch.from_numpy(A).cuda()).cpu().numpy().astype(np.float64)
    assert _pinv_status(eng, A.shape)[3] >= 1.0                               # Jacobi sweeps ran
    ref = np.linalg.pinv(A.astype(np.float64), rcond=1e-15)           # entries up to ~1e5: compare the inverse itself, the fp32
    assert np.linalg.norm(P - ref) <= 1e-5 * np.linalg.norm(ref)      # rounding of the OUTPUT alone moves A P A by cond * eps_32
    B = rng.standard_normal((200, 30)).astype(np.float32)
    B[:, 11] = B[:, 4]
    PB = eng.pinv(torch.from_numpy(B).cuda()).cpu().numpy().astype(np.float64)
    assert _pinv_status(eng, B.shape)[3] >= 1.0
    assert np.linalg.norm(PB - np.linalg.pinv(B.astype(np.float64), rcond=1e-10)) <= 1e-5 * np.linalg.norm(PB)
    C_ = rng.standard_normal((300, 40)).astype(np.float32)
    eng.pinv(torch.from_numpy(C_).cuda(), rcond=1e-3)
    assert _pinv_status(eng, C_.shape)[3] >= 1.0
    assert float(eng.pinv(torch.zeros(9, 4).cuda()).abs().max()) == 0.0


def test_jacobi_status_reports_convergence(eng):
    import ctypes as C
    from anncur_b200 import _lib
    lib = _lib.load()
    A = torch.from_numpy(O.synthetic_scores(300, 64, rank=32, seed=2)).cuda()
    eng.pinv(A, return_cond=True)                       # singular values wanted -> the Jacobi route; raises if not converged
    ws = eng.WORKSPACE.get("pinv", lib.anncur_pinv_workspace_bytes(300, 64), A.device)
    st = torch.zeros(4, dtype=torch.float64, device="cuda")
    _lib.check(lib.anncur_jacobi_status(C.c_void_p(ws.data_ptr()), C.c_void_p(st.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    s_max, s_min, conv, sweeps = st.tolist()
    s = np.linalg.svd(A.cpu().numpy().astype(np.float64), compute_uv=False)
    assert conv == 1.0 and 1 <= sweeps <= 30 and abs(s_max - s[0]) < 1e-6 * s[0] and abs(s_min - s[-1]) < 1e-4 * s[-1] + 1e-9 * s[0]


# ---------------------------------------------------------------------------------------------- FFMA score + top-k
@pytest.mark.parametrize("B,K,N,k", [(1, 50, 10000, 100), (200, 500, 20000, 100), (37, 7, 300, 300), (9, 64, 5000, 1500)])
def test_score_topk_f32_matches_oracle(eng, B, K, N, k):
    Q, E = _rand((B, K), 11), _rand((K, N), 12)
    v, i = eng.score_topk_f32(Q.cuda(), E.cuda(), k)
    dense = (Q @ E).numpy()
    ref = O.score_topk(Q, E, k)
    assert_topk_sets_match(i.cpu().numpy(), ref.indices.numpy(), full_scores=dense)
    assert_scores_close(v.cpu().numpy(), ref.values.numpy())
    assert_sorted_desc(v.cpu().numpy())


# ---------------------------------------------------------------------------------------------- rerank / overlap
@pytest.mark.parametrize("n,N,k_list,k_r", [(40, 900, [1, 10, 50], 200), (7, 3000, [100], 1000), (5, 64, [64], 64)])
def test_rerank_overlap_matches_oracle(eng, n, N, k_list, k_r):
    exact = O.synthetic_scores(n, N, rank=8, seed=3)
    approx = exact + 0.3 * np.random.default_rng(4).standard_normal((n, N)).astype(np.float32)
    k_max = max(k_list)
    lists = O.retrieve_and_rerank(exact, approx, k_max, k_r)
    ex_i, ap_i = torch.from_numpy(lists["exact"][0]).cuda(), torch.from_numpy(lists["approx"][0]).cuda()
    rr_i, rr_v, common = eng.rerank_overlap(torch.from_numpy(exact).cuda(), ap_i, ex_i, k_list)
    want_i, want_v = lists["reranked"]
    assert np.array_equal(rr_v.cpu().numpy(), want_v)                       # exact scores, bit for bit
    assert_topk_sets_match(rr_i.cpu().numpy(), want_i, full_scores=exact)
    for j, k in enumerate(k_list):
        want = [O.overlap_one(lists["exact"][0][q, :k], want_i[q, :k])["common"] for q in range(n)]
        assert common[:, j].cpu().tolist() == want


# ---------------------------------------------------------------------------------------------- recon error
def test_recon_error_matches_torch(eng):
    n, K, N = 300, 40, 5000
    Q, E, A = _rand((n, K), 1), _rand((K, N), 2), _rand((n, N), 3)
    e2, n2 = eng.recon_error_rows(Q.cuda(), E.cuda(), A.cuda())
    d = (Q.double() @ E.double() - A.double())
    assert torch.allclose(e2.cpu(), (d ** 2).sum(1), rtol=1e-5)
    assert torch.allclose(n2.cpu(), (A.double() ** 2).sum(1), rtol=1e-6)


# ---------------------------------------------------------------------------------------------- adaptive (parity unpinned)
def test_adaptive_round_matches_oracle_restatement(eng):
    rng = np.random.default_rng(0)
    k_q, N, B, m, n_next = 40, 700, 9, 16, 8
    A = O.synthetic_scores(k_q + B, N, rank=12, noise=0.02, seed=2)
    R, X = A[:k_q], A[k_q:]
    anchors = np.stack([np.sort(rng.choice(N, m, replace=False)) for _ in range(B)])
    c = np.take_along_axis(X, anchors, 1)
    nxt, val = eng.adaptive_round(torch.from_numpy(R).cuda(), torch.from_numpy(anchors).cuda(), torch.from_numpy(c).cuda(), n_next)
    for q in range(B):
        e = c[q].astype(np.float64) @ np.linalg.pinv(R[:, anchors[q]].astype(np.float64))
        s = e @ R.astype(np.float64)
        s[anchors[q]] = -np.inf
        want = np.argsort(-s, kind="stable")[:n_next]
        kth = s[want[-1]]
        tau = 1e-4 * np.abs(s[np.isfinite(s)]).max()
        got = nxt[q].cpu().numpy()
        assert all(s[j] >= kth - tau for j in got) and len(set(got.tolist())) == n_next
        assert not set(got.tolist()) & set(anchors[q].tolist())
        assert np.allclose(val[q].cpu().numpy(), s[got], atol=10 * tau)


def test_adaptive_round_split_form_matches_oracle_restatement(eng):
    """The same round through the split form the fused path uses: anncur_adaptive_solve (e_b), the fused tensor-core score +
    top-(n_next + m) on the PACKED R_anc, anncur_filter_excluded.  Picks must equal the fp64 oracle's outside the tie band."""
    rng = np.random.default_rng(1)
    k_q, N, B, m, n_next = 48, 9000, 70, 24, 12
    A = O.synthetic_scores(k_q + B, N, rank=12, noise=0.02, seed=4)
    R, X = A[:k_q], A[k_q:]
    anchors = np.stack([np.sort(rng.choice(N, m, replace=False)) for _ in range(B)])
    c = np.take_along_axis(X, anchors, 1)
    Rc = torch.from_numpy(R).cuda()
    e = eng.adaptive_solve(Rc, torch.from_numpy(anchors).cuda(), torch.from_numpy(c).cuda(), Rt=eng.transpose(Rc))
    e2 = eng.adaptive_solve(Rc, torch.from_numpy(anchors).cuda(), torch.from_numpy(c).cuda())       # Rt rebuilt in the workspace
    assert torch.equal(e, e2)
    cv, ci = eng.score_topk(e, eng.PackedItems(Rc, "f32r"), n_next + m)
    val, nxt = eng.filter_excluded(cv, ci, torch.from_numpy(anchors).cuda(), n_next)
    for q in range(B):
        eq = c[q].astype(np.float64) @ np.linalg.pinv(R[:, anchors[q]].astype(np.float64))
        assert np.allclose(e[q].cpu().numpy(), eq, atol=1e-4 * np.abs(eq).max())
        sc = eq @ R.astype(np.float64)
        sc[anchors[q]] = -np.inf
        want = np.argsort(-sc, kind="stable")[:n_next]
        tau = 1e-4 * np.abs(sc[np.isfinite(sc)]).max()
        got = nxt[q].cpu().numpy()
        assert all(sc[j] >= sc[want[-1]] - tau for j in got) and len(set(got.tolist())) == n_next
        assert not set(got.tolist()) & set(anchors[q].tolist())
        assert np.allclose(val[q].cpu().numpy(), sc[got], atol=10 * tau)


def test_filter_excluded_keeps_order_and_pads(eng):
    cv = torch.tensor([[9.0, 8.0, 7.0, 6.0, 5.0], [4.0, 3.0, 2.0, -3.4e38, -3.4e38]]).cuda()
    ci = torch.tensor([[10, 11, 12, 13, 14], [5, 6, 7, -1, -1]]).cuda()
    ex = torch.tensor([[11, 13, 99], [5, 6, 7]]).cuda()
    v, i = eng.filter_excluded(cv, ci, ex, 3)
    assert i.tolist() == [[10, 12, 14], [-1, -1, -1]] and v[0].tolist() == [9.0, 7.0, 5.0]
    assert (v[1] == -np.finfo(np.float32).max).all()
    v, i = eng.filter_excluded(cv, ci, ex[:, :0], 2)                      # nothing excluded
    assert i.tolist() == [[10, 11], [5, 6]]
    rng = np.random.default_rng(0)                                       # many rows, long lists: against a python loop
    n, k_in, m, n_out = 300, 500, 375, 125
    ci = torch.from_numpy(np.stack([rng.permutation(5000)[:k_in] for _ in range(n)])).cuda()
    cv = torch.sort(torch.randn(n, k_in), dim=1, descending=True).values.cuda()
    ex = torch.from_numpy(np.stack([rng.permutation(5000)[:m] for _ in range(n)])).cuda()
    v, i = eng.filter_excluded(cv, ci, ex, n_out)
    for r in range(0, n, 17):
        keep = [j for j in ci[r].tolist() if j not in set(ex[r].tolist())][:n_out]
        assert i[r].tolist()[:len(keep)] == keep and all(x == -1 for x in i[r].tolist()[len(keep):])


def _np_pinv_e(R, anchors_q, c_q, rcond=1e-15):
    return c_q.astype(np.float64) @ np.linalg.pinv(R[:, anchors_q].astype(np.float64), rcond=rcond)


@pytest.mark.parametrize("k_q,N,B,s,n,rounds", [(60, 700, 9, 7, 12, 3), (64, 900, 5, 0, 16, 3), (500, 5000, 6, 125, 125, 2),
                                                (130, 1500, 3, 128, 1, 2), (96, 800, 4, 20, 33, 2)])
def test_adaptive_incremental_solver_matches_fp64_pinv(eng, k_q, N, B, s, n, rounds):
    """anncur_adaptive_prepare / _begin / _extend (the factor carried across rounds, shared first anchors gathered from
    W_1) against numpy's fp64 pinv of the grown anchor set, and against the from-scratch anncur_adaptive_solve."""
    rng = np.random.default_rng(3)
    A = O.synthetic_scores(k_q + B, N, rank=max(8, k_q // 8), noise=0.05, seed=11)
    R, X = A[:k_q], A[k_q:]
    Rc = torch.from_numpy(R).cuda()
    Rt = eng.transpose(Rc)
    perm = np.stack([rng.permutation(N) for _ in range(B)])
    first = np.sort(perm[0, :s])
    shared = eng.AdaptiveShared(Rt, torch.from_numpy(first), 1e-15)
    state = eng.AdaptiveState(shared, B, n, s + rounds * n)
    e = state.begin(torch.from_numpy(X[:, first]).cuda()).cpu().numpy()
    cur = np.tile(first[None, :], (B, 1))
    for q in range(B):
        if s > 0:
            want = _np_pinv_e(R, cur[q], X[q, cur[q]])
            assert np.allclose(e[q], want, atol=2e-5 * np.abs(want).max()), (q, np.abs(e[q] - want).max(), np.abs(want).max())
        else:
            assert (e[q] == 0).all()
    for t in range(rounds):
        new = np.stack([[j for j in perm[q] if j not in set(first.tolist())][t * n:(t + 1) * n] for q in range(B)]).astype(np.int64)
        c_new = np.take_along_axis(X, new, 1)
        e = state.extend(torch.from_numpy(new).cuda(), torch.from_numpy(c_new).cuda()).cpu().numpy()
        cur = np.concatenate([cur, new], 1)
        full = eng.adaptive_solve(Rc, torch.from_numpy(cur).cuda(), torch.from_numpy(np.take_along_axis(X, cur, 1)).cuda(), Rt=Rt).cpu().numpy()
        for q in range(B):
            want = _np_pinv_e(R, cur[q], X[q, cur[q]])
            scale = np.abs(want).max()
            assert np.allclose(e[q], want, atol=2e-5 * scale), (t, q, np.abs(e[q] - want).max(), scale)
            assert np.allclose(e[q], full[q], atol=2e-5 * scale)


def test_adaptive_incremental_solver_drops_dependent_anchors_like_the_full_solve(eng):
    """Two anchors with identical columns of R_anc: the later one's pivot falls below rcond^2 * max diagonal and its
    coordinate is dropped -- in the block form exactly as in the from-scratch Cholesky."""
    rng = np.random.default_rng(5)
    k_q, N, B, s, n = 48, 600, 4, 8, 8
    A = O.synthetic_scores(k_q + B, N, rank=10, noise=0.05, seed=13)
    A[:, 100] = A[:, 7]                                   # item 100 duplicates item 7 (also in the exact rows: consistent system)
    A[:, 200] = A[:, 300]
    R, X = A[:k_q], A[k_q:]
    Rc = torch.from_numpy(R).cuda()
    Rt = eng.transpose(Rc)
    first = np.array([3, 7, 50, 90, 120, 200, 310, 400])
    new = np.stack([np.array([100, 5, 300, 11, 13, 17, 19, 23]) + 0 * q for q in range(B)]).astype(np.int64)
    new[1, 3] = 555
    shared = eng.AdaptiveShared(Rt, torch.from_numpy(first), 1e-6)
    state = eng.AdaptiveState(shared, B, n, s + n)
    state.begin(torch.from_numpy(X[:, first]).cuda())
    e = state.extend(torch.from_numpy(new).cuda(), torch.from_numpy(np.take_along_axis(X, new, 1)).cuda()).cpu().numpy()
    cur = np.concatenate([np.tile(first[None], (B, 1)), new], 1)
    full = eng.adaptive_solve(Rc, torch.from_numpy(cur).cuda(), torch.from_numpy(np.take_along_axis(X, cur, 1)).cuda(), rcond=1e-6, Rt=Rt).cpu().numpy()
    assert np.isfinite(e).all() and np.allclose(e, full, atol=1e-5 * np.abs(full).max())
    for q in range(B):                                    # consistent duplicates: dropping one of the pair IS the pinv answer
        want = _np_pinv_e(R, cur[q], X[q, cur[q]], rcond=1e-9)
        assert np.allclose(e[q], want, atol=1e-4 * np.abs(want).max())


def test_adaptive_incremental_solver_ignores_invalid_anchor_indices(eng):
    """An index outside [0, N) among the new anchors (the -1 padding of a short candidate list) is a zero column: its pivot is
    dropped and the solve equals the solve without it -- no out-of-range read."""
    rng = np.random.default_rng(8)
    k_q, N, B, s, n = 40, 500, 3, 6, 10
    A = O.synthetic_scores(k_q + B, N, rank=8, noise=0.05, seed=19)
    R, X = A[:k_q], A[k_q:]
    Rt = eng.transpose(torch.from_numpy(R).cuda())
    first = np.sort(rng.choice(N, s, replace=False))
    rest = np.array([j for j in rng.permutation(N) if j not in set(first.tolist())])
    new = np.stack([rest[q * n:(q + 1) * n] for q in range(B)]).astype(np.int64)
    c_new = np.take_along_axis(X, new, 1)
    new_bad = new.copy()
    new_bad[0, 3], new_bad[2, 9], new_bad[2, 0] = -1, N + 7, -1
    state = eng.AdaptiveState(eng.AdaptiveShared(Rt, torch.from_numpy(first), 1e-15), B, n, s + n)
    state.begin(torch.from_numpy(X[:, first]).cuda())
    e = state.extend(torch.from_numpy(new_bad).cuda(), torch.from_numpy(c_new).cuda()).cpu().numpy()
    assert np.isfinite(e).all()
    for q in range(B):
        keep = [j for j in range(n) if 0 <= new_bad[q, j] < N]
        cur = np.concatenate([first, new[q, keep]])
        want = _np_pinv_e(R, cur, X[q, cur])
        assert np.allclose(e[q], want, atol=2e-5 * np.abs(want).max()), q


def test_adaptive_rounds_pick_exactly_outside_the_tie_band(eng):
    """Every round of the whole procedure (incremental solver + fused re-score + anchor filter) against the fp64 statement
    of that round ON THE SAME ANCHOR SET: each pick must score within tau = 1e-4 max|s| of the oracle's n-th best unmasked
    item (so picks differ from the oracle's only among ties inside the band), never repeat an anchor, and come best first."""
    from anncur_b200 import adaptive_anncur
    rng = np.random.default_rng(9)
    k_q, N, B, T, kpr, top_k = 64, 5000, 20, 4, 16, 10
    A = O.synthetic_scores(k_q + B, N, rank=12, noise=0.05, seed=17)
    R, X = A[:k_q], A[k_q:]
    first = np.sort(rng.choice(N, kpr, replace=False))
    anc, idx, val = adaptive_anncur(torch.from_numpy(R), torch.from_numpy(X), first, T, kpr, top_k)
    anc = anc.cpu().numpy()
    R64 = R.astype(np.float64)
    n_exact = 0
    for q in range(B):
        for t in range(1, T):
            cur, got = anc[q, :t * kpr], anc[q, t * kpr:(t + 1) * kpr]
            sc = _np_pinv_e(R, cur, X[q, cur]) @ R64
            sc[cur] = -np.inf
            order = np.argsort(-sc, kind="stable")
            tau = 1e-4 * np.abs(sc[np.isfinite(sc)]).max()
            assert len(set(got.tolist())) == kpr and not set(got.tolist()) & set(cur.tolist())
            assert all(sc[j] >= sc[order[kpr - 1]] - tau for j in got), (q, t)
            assert all(sc[got[i]] >= sc[got[i + 1]] - tau for i in range(kpr - 1))
            n_exact += set(got.tolist()) == set(order[:kpr].tolist())
    assert n_exact >= 0.9 * B * (T - 1)                   # and almost always it IS the oracle's set
    want_anc, want_idx, want_val, _ = O.adaptive_anncur(R, X, first, T, kpr, top_k, rcond=1e-15)
    assert np.mean([len(set(anc[q].tolist()) & set(want_anc[q].tolist())) / (T * kpr) for q in range(B)]) > 0.97


@pytest.mark.parametrize("rescore", ["fused", "fused-full", "ffma"])
def test_adaptive_multi_round_matches_oracle_restatement(eng, rescore):
    """The whole multi-round procedure (anncur_b200.adaptive_anncur: T - 1 K8 calls, exact-score gathers, K9 at the end)
    against oracle.cur_oracle.adaptive_anncur.  Picks of a round may differ from the oracle's only among near-ties of the
    approximate scores, so the comparison is on what the procedure is for: the final exact scores and the overlap of the
    anchor sets."""
    from anncur_b200 import adaptive_anncur
    rng = np.random.default_rng(5)
    k_q, N, B, T, kpr, top_k = 60, 3000, 24, 4, 12, 10
    A = O.synthetic_scores(k_q + B, N, rank=10, noise=0.05, seed=7)
    R, X = A[:k_q], A[k_q:]
    first = np.sort(rng.choice(N, kpr, replace=False))
    want_anc, want_idx, want_val, _ = O.adaptive_anncur(R, X, first, T, kpr, top_k, rcond=1e-15)
    anc, idx, val = adaptive_anncur(torch.from_numpy(R), torch.from_numpy(X), first, T, kpr, top_k, rescore=rescore.split("-")[0],
                                    solver="full" if rescore.endswith("full") else "incremental")
    anc, idx, val = anc.cpu().numpy(), idx.cpu().numpy(), val.cpu().numpy()
    assert anc.shape == (B, T * kpr) and (anc[:, :kpr] == first[None, :]).all()
    assert all(len(set(r.tolist())) == T * kpr for r in anc)                 # anchors are never re-picked
    overlap = np.mean([len(set(anc[q].tolist()) & set(want_anc[q].tolist())) / (T * kpr) for q in range(B)])
    assert overlap > 0.97, overlap
    assert (val == np.take_along_axis(X, idx, 1)).all()                      # returned values are the exact scores of the returned items
    assert np.allclose(val, want_val, atol=2e-3 * np.abs(X).max())


def test_singular_values_and_matrix_rank(eng):
    """eval/compute_m2e_matrix_ranks.py:44-53: np.linalg.matrix_rank of a score matrix."""
    rng = np.random.default_rng(0)
    for (m, n, r) in [(60, 200, 13), (300, 90, 90), (128, 128, 40)]:
        A = (rng.standard_normal((m, r)) @ rng.standard_normal((r, n))).astype(np.float32)
        s = eng.singular_values(torch.from_numpy(A)).cpu().numpy()
        ref = np.linalg.svd(A.astype(np.float64), compute_uv=False)
        assert np.allclose(s[:r], ref[:r], rtol=1e-6, atol=1e-6 * ref[0])
        assert eng.matrix_rank(torch.from_numpy(A)) == np.linalg.matrix_rank(A) == r


def test_orthonormalize_gives_left_singular_vectors(eng):
    rng = np.random.default_rng(1)
    Y = (rng.standard_normal((700, 48)) * np.logspace(0, -5, 48)[None, :]).astype(np.float32)
    Y[:, 5] = Y[:, 2]                                                  # rank-deficient: one zero column must come out
    Q, s = eng.orthonormalize(torch.from_numpy(Y).cuda())
    Q, s = Q.cpu().double().numpy(), s.cpu().numpy()
    keep = np.linalg.norm(Q, axis=0) > 0.5
    assert keep.sum() == 47
    G = Q[:, keep].T @ Q[:, keep]
    assert np.abs(G - np.eye(47)).max() < 1e-5                          # orthonormal (fp32 output)
    ref = np.linalg.svd(Y.astype(np.float64), compute_uv=False)
    assert np.allclose(np.sort(s)[::-1][:47], ref[:47], rtol=1e-6, atol=1e-9 * ref[0])
    # same column space: projecting Y onto span(Q) loses nothing
    P = Q[:, keep] @ (Q[:, keep].T @ Y.astype(np.float64))
    assert np.linalg.norm(P - Y) < 1e-5 * np.linalg.norm(Y)


def test_matrix_rank_large_randomised_matches_numpy(eng):
    """eval/compute_m2e_matrix_ranks.py:50 at a size where both routes run: randomised subspace iteration (GEMMs + Jacobi on
    small matrices) against np.linalg.matrix_rank / np.linalg.svd of the same matrix -- low-rank + noise (the rank is then
    the number of values above numpy's tolerance, not the noise dimension), exact low rank, and a block that must grow."""
    rng = np.random.default_rng(2)
    n, N, r = 900, 20000, 40
    sig = np.linspace(60.0, 5.0, r)
    A = ((rng.standard_normal((n, r)) * sig[None, :]) @ rng.standard_normal((r, N)) / np.sqrt(n) + 0.01 * rng.standard_normal((n, N))).astype(np.float32)
    At = torch.from_numpy(A).cuda()
    ref_s = np.linalg.svd(A.astype(np.float64), compute_uv=False)
    s = eng.leading_singular_values(At, 96, n_iter=4).cpu().numpy()
    assert np.allclose(s[:r], ref_s[:r], rtol=2e-4)
    assert eng.matrix_rank_large(At, n_values=96) == np.linalg.matrix_rank(A)
    assert eng.matrix_rank_large(At.t().contiguous(), n_values=96) == np.linalg.matrix_rank(A)        # tall orientation
    assert eng.matrix_rank_large(At, n_values=16) == np.linalg.matrix_rank(A)                        # block doubles until it holds the rank
    assert eng.matrix_rank_large(At, tol=1.0, n_values=64) == int((ref_s > 1.0).sum())
    B = (rng.standard_normal((300, 7)) @ rng.standard_normal((7, 5000))).astype(np.float32)          # exact rank 7
    assert eng.matrix_rank_large(torch.from_numpy(B).cuda(), n_values=32) == 7 == np.linalg.matrix_rank(B)


def test_topk_rows_large_rows_sampled_path(eng):
    """Rows long enough for the sampled-threshold kernel (>= 32768 columns): random, sorted ascending / descending,
    heavy ties (falls back to the radix path), unaligned row starts, k from 1 to 1000."""
    rng = np.random.default_rng(3)
    S = torch.from_numpy(rng.standard_normal((37, 100003), dtype=np.float32))
    S[1] = torch.arange(100003, dtype=torch.float32)               # ascending
    S[2] = -torch.arange(100003, dtype=torch.float32)              # descending
    S[3] = 0.0                                                     # all ties -> indices 0..k-1
    S[4, :50000] = 7.0                                             # 50000-way tie at the top
    S[5, 99990:] = 100.0                                           # winners in the scalar tail
    for k in (1, 10, 100, 1000):
        v, i = eng.topk_rows(S.cuda(), k)
        ref = torch.topk(S, k, dim=1)
        assert torch.equal(v.cpu(), ref.values)
        assert torch.equal(torch.gather(S, 1, i.cpu()), v.cpu())
        assert (i[3].cpu() == torch.arange(k)).all() and (i[4].cpu() == torch.arange(k)).all()
        assert all(len(set(r.tolist())) == k for r in i.cpu())
    # row stride that breaks 16-byte alignment of odd rows
    big = torch.from_numpy(rng.standard_normal((6, 50001), dtype=np.float32)).cuda()
    v, i = eng.topk_rows(big[:, 1:], 64)
    ref = torch.topk(big[:, 1:].cpu(), 64, dim=1)
    assert torch.equal(v.cpu(), ref.values) and torch.equal(i.cpu(), ref.indices)


def test_merge_topk_keys_equals_pair_merge(eng):
    """Key-form exchange of the item-sharded search: keys[shard][row][k] merged as gathered == the (vals, idx) merge."""
    rng = np.random.default_rng(11)
    P, n, k = 8, 333, 100
    vals = torch.from_numpy(rng.standard_normal((P, n, k), dtype=np.float32)).sort(dim=2, descending=True).values
    vals[:, :, 50:] = vals[:, :, 49:50]                        # ties inside and across shards
    idx = torch.stack([torch.stack([torch.from_numpy(rng.choice(1_000_000, k, replace=False)) + p * 1_000_000 for _ in range(n)])
                       for p in range(P)]).to(torch.int64)
    idx[3, :, 90:] = -1                                         # a shard that holds fewer than k items: padding
    keys = torch.stack([eng.topk_to_keys(vals[p].cuda(), idx[p].cuda()) for p in range(P)]).contiguous()
    for k_out in (1, 100, 500):
        mv, mi = eng.merge_topk_keys(keys, k_out)
        cv = vals.permute(1, 0, 2).reshape(n, P * k)
        ci = idx.permute(1, 0, 2).reshape(n, P * k)
        rv, ri = eng.merge_topk(cv.cuda(), ci.cuda(), k_out)
        assert torch.equal(mv, rv) and torch.equal(mi, ri)
    ov, oi = O.merge_topk(vals.numpy(), idx.numpy(), 100)
    mv, mi = eng.merge_topk_keys(keys, 100)
    assert np.array_equal(mi.cpu().numpy(), oi) and np.array_equal(mv.cpu().numpy(), ov)


def test_overlap_counts_set_semantics(eng):
    """eval/eval_utils.py:139-150: len(set(a) & set(b)) per row, duplicates and -1 padding included."""
    rng = np.random.default_rng(9)
    for (n, k, hi) in [(50, 1, 5), (200, 100, 300), (33, 1000, 5000), (7, 64, 40)]:       # last: many duplicates
        a = rng.integers(-1, hi, size=(n, k))
        b = rng.integers(-1, hi, size=(n, k))
        got = eng.overlap_counts(torch.from_numpy(a), torch.from_numpy(b)).cpu().numpy()
        want = np.array([len(set(a[r].tolist()) & set(b[r].tolist())) for r in range(n)])
        assert np.array_equal(got, want)


# ---- dense products on the tcgen05 pipeline (anncur_score_dense / anncur_recon_error_packed) -------------------------
@pytest.mark.parametrize("B,K,N,kind", [
    (500, 2000, 30000, "f32x3"),     # the item-embedding build U @ R: 63 k-blocks
    (300, 500, 70001, "f32r"),       # get_complete_row on an F32R index (bound slot must not leak), ragged N (unaligned rows)
    (7, 50, 257, "f32x3"),           # tiny / ragged everything, single CTA
    (129, 64, 4096, "f32r"),         # K % 32 == 0 on F32R: extra k-block
])
def test_score_dense_matches_fp64(eng, B, K, N, kind):
    rng = np.random.default_rng(B + K + N)
    Q = torch.from_numpy(rng.standard_normal((B, K), dtype=np.float32) * np.logspace(-2, 2, B, dtype=np.float32)[:, None])
    E = torch.from_numpy(rng.standard_normal((K, N), dtype=np.float32))
    out = eng.score_dense(Q.cuda(), eng.PackedItems(E.cuda(), kind)).cpu().double()
    ref = Q.double() @ E.double()
    scale = (Q.double().abs() @ E.double().abs()).clamp_min(1e-30)          # sum_i |q_i e_in|
    assert ((out - ref).abs() / scale).max().item() < 2e-6


def test_gemm_tc_and_recon_error_packed(eng):
    rng = np.random.default_rng(9)
    U = torch.from_numpy(rng.standard_normal((200, 700), dtype=np.float32))
    R = torch.from_numpy(rng.standard_normal((700, 50000), dtype=np.float32))
    E = eng.gemm_tc(U.cuda(), R.cuda())
    ref = U.double() @ R.double()
    assert ((E.cpu().double() - ref).abs().max() / ref.abs().max()).item() < 1e-5
    Q = torch.from_numpy(rng.standard_normal((333, 200), dtype=np.float32))
    # "exact" matrix = approximation + noise of 5 % of the score scale (the errors being summed are far above fp32 rounding)
    S = Q.double() @ ref
    A = (S + torch.from_numpy(rng.standard_normal((333, 50000))) * 0.05 * S.std()).float()
    for kind in ("f32x3", "f32r"):
        e2, n2 = eng.recon_error_packed(Q.cuda(), eng.PackedItems(E, kind), A.cuda())
        want_e2 = ((Q.double() @ E.cpu().double() - A.double()) ** 2).sum(1)
        want_n2 = (A.double() ** 2).sum(1)
        assert torch.allclose(e2.cpu(), want_e2, rtol=1e-4) and torch.allclose(n2.cpu(), want_n2, rtol=1e-5)


# ---------------------------------------------------------------------------------------------- randomised eval-loop pieces
def test_eval_loop_pieces_random_shapes_with_ties(eng):
    """topk_rows, rerank_overlap and merge_topk on random shapes, half of them on INTEGER-valued scores (thousands of exact
    ties): selection is bit-exact work, so values must be identical and, because the library breaks ties towards the lower
    index, so must the indices (lexicographic order (-score, index) = a stable descending argsort)."""
    rng = np.random.default_rng(2024)
    for case in range(24):
        n = int(rng.choice([1, 3, 17, 40]))
        N = int(np.exp(rng.uniform(np.log(2), np.log(120_000))))
        k = int(min(rng.choice([1, 5, 10, 64, 100, 333, 1000]), N))
        k_r = int(min(max(k, rng.choice([1, 10, 100, 500, 1000, 2000])), N))
        if case % 2:
            exact = rng.integers(-6, 7, size=(n, N)).astype(np.float32)
            approx = exact + rng.integers(-2, 3, size=(n, N)).astype(np.float32)
        else:
            exact = rng.standard_normal((n, N)).astype(np.float32)
            approx = exact + 0.3 * rng.standard_normal((n, N)).astype(np.float32)
        order = np.lexsort((np.broadcast_to(np.arange(N), (n, N)), -exact.astype(np.float64)), axis=1)[:, :k]
        v, i = eng.topk_rows(torch.from_numpy(exact).cuda(), k)
        assert np.array_equal(i.cpu().numpy(), order), (case, n, N, k)
        assert np.array_equal(v.cpu().numpy(), np.take_along_axis(exact, order, 1))
        a_order = np.lexsort((np.broadcast_to(np.arange(N), (n, N)), -approx.astype(np.float64)), axis=1)[:, :k_r]
        av, ai = eng.topk_rows(torch.from_numpy(approx).cuda(), k_r)
        assert np.array_equal(ai.cpu().numpy(), a_order), (case, n, N, k_r)
        # masked re-rank (..._w_fixed_train_test_splits.py:92-99): the k best retrieved items by EXACT score
        got = np.take_along_axis(exact, a_order, 1)
        rr = np.lexsort((a_order, -got.astype(np.float64)), axis=1)[:, :k]
        want_i, want_v = np.take_along_axis(a_order, rr, 1), np.take_along_axis(got, rr, 1)
        ks = sorted({1, min(10, k), k})
        rr_i, rr_v, common = eng.rerank_overlap(torch.from_numpy(exact).cuda(), ai, i, ks)
        assert np.array_equal(rr_v.cpu().numpy(), want_v) and np.array_equal(rr_i.cpu().numpy(), want_i), (case, n, N, k, k_r)
        for j, kk in enumerate(ks):
            want = [len(set(order[q, :kk].tolist()) & set(want_i[q, :kk].tolist())) for q in range(n)]
            assert common[:, j].cpu().tolist() == want, (case, kk)
        # shard merge: split the items into P shards, local top-k of each, merged = global top-k
        P = int(rng.choice([2, 3, 8]))
        bounds = [(p * N) // P for p in range(P + 1)]
        if min(b1 - b0 for b0, b1 in zip(bounds[:-1], bounds[1:])) < 1:
            continue
        cv, ci = [], []
        for p in range(P):
            lv, li = eng.topk_rows(torch.from_numpy(np.ascontiguousarray(exact[:, bounds[p]:bounds[p + 1]])).cuda(), k, idx_offset=bounds[p])
            cv.append(lv)
            ci.append(li)
        mv, mi = eng.merge_topk(torch.cat(cv, 1), torch.cat(ci, 1), k)
        assert np.array_equal(mi.cpu().numpy(), order) and np.array_equal(mv.cpu().numpy(), np.take_along_axis(exact, order, 1)), (case, P)
