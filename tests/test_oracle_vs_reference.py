"""CPU, build container only: the oracle against the reference executed live (skipped where
/root/reference is absent, e.g. on the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import cur_oracle as O
from oracle.ref_shim import load_reference, reference_available
from tests.parity import assert_scores_close, assert_topk_sets_match

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    import warnings
    warnings.filterwarnings("ignore", category=UserWarning)
    return load_reference()


@pytest.mark.parametrize("n,N,k_q,k_i,seed", [(80, 500, 30, 12, 0), (64, 640, 16, 40, 1), (50, 300, 50, 20, 2)])
def test_live_curapprox(ref, n, N, k_q, k_i, seed):
    A = torch.from_numpy(O.synthetic_scores(n, N, rank=8, seed=seed))
    rows_i, cols_i = O.sample_anchors(n, N, k_q, k_i, seed)
    rows, cols = A[rows_i, :], A[:, cols_i]
    want = ref.CURApprox(rows=rows, cols=cols, row_idxs=rows_i, col_idxs=cols_i, approx_preference="rows")
    got = O.cur_build(rows, cols, rows_i, cols_i, "rows")
    assert torch.equal(got.U, want.U) and torch.equal(got.latent_cols, want.latent_cols)
    Q = A[:, cols_i]
    assert torch.equal(O.get_complete_row(got, Q), want.get_complete_row(Q))
    a, b = O.topk_in_row(got, Q, 7), want.topk_in_row(Q, 7)
    assert torch.equal(a.indices, b.indices) and torch.equal(a.values, b.values)
    assert torch.equal(O.get(got, list(range(n)), list(range(N))), want.get(list(range(n)), list(range(N))))


def test_live_sweep_eval(ref):
    A = torch.from_numpy(O.synthetic_scores(70, 600, rank=8, seed=9))
    for method in ("cur", "cur_oracle"):
        want = ref.run_approx_eval_w_seed(method, A, 25, 30, 5, 40, 3, None)
        got = O.run_approx_eval_w_seed(method, A, 25, 30, 5, 40, 3)
        for grp in want:
            assert set(got[grp]) == set(want[grp])
            for k in want[grp]:
                assert float(got[grp][k]) == pytest.approx(float(want[grp][k]), rel=1e-6, abs=1e-6), (grp, k)


def test_live_fixed_split_eval(ref):
    A = torch.from_numpy(O.synthetic_scores(90, 700, rank=8, seed=11))
    train, test = A[:60], A[60:]
    cur = O.fixed_split_cur_scores(train, test, [10, 30], seed=0)
    for k_i, (anc, approx) in cur.items():
        for k_r in (5, 60):
            want = ref.eval_approx_score_mat_for_all_topk(test, approx, [1, 10, 50], k_r)
            got = O.eval_approx_score_mat_for_all_topk(test, approx, [1, 10, 50], k_r)
            assert got == want
        assert O.eval_approx_score_mat(test, approx, 10, 60) == ref.eval_approx_score_mat(test, approx, 10, 60)
