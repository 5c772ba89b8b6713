"""GPU: the drop-in Python surface (CURApprox, build_flat_or_ivff_index, eval functions) against the
reference's own outputs stored in tests/golden and against the oracle."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import cur_oracle as O
from tests.parity import assert_scores_close, assert_sorted_desc, assert_topk_sets_match

pytestmark = pytest.mark.gpu


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


@pytest.mark.parametrize("name", ["curapprox_tall", "curapprox_wide", "curapprox_square"])
@pytest.mark.parametrize("where", ["cpu", "cuda"])
def test_curapprox_dropin_matches_reference_golden(golden_dir, name, where):
    from anncur_b200 import CURApprox
    g = _load(golden_dir, name)
    A = torch.from_numpy(g["A"]).to(where)
    r, c = g["row_idxs"].tolist(), g["col_idxs"].tolist()
    rows, cols = A[r, :], A[:, c]
    ap = CURApprox(rows=rows, cols=cols, row_idxs=r, col_idxs=c, approx_preference="rows", check=True)
    assert ap.latent_cols.device.type == where and ap.n == A.shape[0] and ap.m == A.shape[1]
    tol = max(1e-4, 10 * float(g["cond_intersect"]) * np.finfo(np.float32).eps)
    rel = lambda a, b: np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
    assert rel(ap.U.cpu().numpy(), g["rows_U"]) <= tol
    assert rel(ap.latent_cols.cpu().numpy(), g["rows_latent_cols"]) <= tol
    if name == "curapprox_square":
        return          # square intersections amplify last-bit differences by cond: E-level check only
    Q = A[g["test_rows"].tolist()][:, c]
    k = int(g["k"])
    dense = ap.get_complete_row(Q)
    assert dense.device.type == where
    assert_scores_close(dense.cpu().numpy(), g["rows_get_complete_row"], rel=max(1e-4, tol))
    for precision in (None, "f32r", "f32x3", "f32"):              # None = the default kind (f32r), as a reference script would call it
        tk = ap.topk_in_row(Q, k) if precision is None else ap.topk_in_row(Q, k, precision=precision)
        assert tk.indices.dtype == torch.int64 and tk.values.device.type == where
        assert_sorted_desc(tk.values.cpu().numpy())
        assert_topk_sets_match(tk.indices.cpu().numpy(), g["rows_topk_indices"], full_scores=g["rows_get_complete_row"],
                               rel=max(1e-4, tol))
        assert_scores_close(tk.values.cpu().numpy(), np.take_along_axis(g["rows_get_complete_row"], tk.indices.cpu().numpy(), 1),
                            rel=max(1e-4, tol))
    sel_r, sel_c = g["sel_r"].tolist(), g["sel_c"].tolist()
    assert_scores_close(ap.get(sel_r, sel_c).cpu().numpy(), g["rows_get"], rel=max(1e-4, tol))
    assert_scores_close(ap.get_rows(sel_r).cpu().numpy(), g["rows_get_rows"], rel=max(1e-4, tol))
    assert_scores_close(ap.get_cols(sel_c).cpu().numpy(), g["rows_get_cols"], rel=max(1e-4, tol))
    with pytest.raises(NotImplementedError):
        ap.get_complete_col(Q)
    ac = CURApprox(rows=rows, cols=cols, row_idxs=r, col_idxs=c, approx_preference="cols")
    assert rel(ac.latent_rows.cpu().numpy(), g["cols_latent_rows"]) <= max(1e-4, tol)
    sparse_cols = A[r][:, sel_c[:5]]
    assert_scores_close(ac.get_complete_col(sparse_cols).cpu().numpy(), g["cols_get_complete_col"], rel=max(1e-4, tol))
    tkc = ac.topk_in_col(sparse_cols, 3)
    assert_topk_sets_match(tkc.indices.cpu().numpy(), g["cols_topk_indices"], full_scores=g["cols_get_complete_col"])
    with pytest.raises(NotImplementedError):
        ac.topk_in_row(Q, 3)


def test_curapprox_error_behaviour(golden_dir):
    from anncur_b200 import CURApprox
    g = _load(golden_dir, "curapprox_tall")
    A = torch.from_numpy(g["A"])
    r, c = g["row_idxs"].tolist(), g["col_idxs"].tolist()
    with pytest.raises(NotImplementedError):
        CURApprox(A[r, :], A[:, c], r, c, "diag")
    with pytest.raises(AssertionError):
        CURApprox(A[r, :], A[:, c], r[::-1], c, "rows")
    with pytest.raises(AssertionError):
        CURApprox(A[r, :], A[:, c], r[:-1], c, "rows")
    bad = A[r, :].clone()
    bad[0, c[0]] += 1
    with pytest.raises(AssertionError):
        CURApprox(bad, A[:, c], r, c, "rows", check=True)
    ap = CURApprox(A[r, :], A[:, c], r, c, "rows")
    with pytest.raises(RuntimeError):
        ap.topk_in_row(A[:2][:, c], A.shape[1] + 1)          # torch.topk raises for k > N


def _close_dict(ours, ref, atol):
    assert set(ours) == set(ref)
    for k in ref:
        assert abs(float(ours[k]) - float(ref[k])) <= atol * max(1.0, abs(float(ref[k]))), (k, ours[k], ref[k])


def test_sweep_eval_dropin_matches_reference_golden(golden_dir):
    from anncur_b200 import run_approx_eval_w_seed
    g = _load(golden_dir, "sweep_eval")
    k_q, k_i, top_k, k_r, seed = [int(x) for x in g["params"]]
    for method in ("cur", "cur_oracle"):
        ref = json.loads(str(g[method + "_json"]))
        ours = run_approx_eval_w_seed(method, torch.from_numpy(g["A"]), k_q, k_i, top_k, k_r, seed, None)
        for grp in ("anchor", "non_anchor", "all"):
            _close_dict(ours[grp], ref[grp], atol=3e-3)


def test_sweep_tool_grid_matches_reference_golden_point(golden_dir, tmp_path):
    """run_sweep / the run_sweep_eval CLI (the k_q x k_i grid of eval/run_retrieval_eval_wrt_exact_crossenc.py:227-233 as one
    run): the grid point the golden fixture holds must come out with the reference's numbers, seed-averaged over one seed,
    under the reference's key layout; the other points must be present and sane."""
    from anncur_b200 import data_formats as F
    from anncur_b200.eval_retrieval import run_sweep, sweep_grids
    from anncur_b200.run_sweep_eval import main
    g = _load(golden_dir, "sweep_eval")
    k_q, k_i, top_k, k_r, seed = [int(x) for x in g["params"]]
    assert seed == 0
    A = torch.from_numpy(g["A"])
    gm, ge = sweep_grids(*A.shape)
    assert gm == [v for v in [50, 100, 200, 500, 1000, 2000, 5000] if v <= A.shape[0]] and ge[-1] == A.shape[1]
    res = run_sweep(A, n_seeds=1, top_k_vals=[top_k], top_k_retr_vals=[k_r], n_ment_anchors_vals=[k_q, gm[0]], n_ent_anchors_vals=[k_i, ge[0]])
    for method in ("cur", "cur_oracle"):
        ref = json.loads(str(g[method + "_json"]))
        ours = res[method][f"top_k={top_k}"][f"k_retvr={k_r}"][f"anc_n_m={k_q}~anc_n_e={k_i}"]
        for grp in ("anchor", "non_anchor", "all"):
            _close_dict(ours[grp], ref[grp], atol=3e-3)
        assert len(res[method][f"top_k={top_k}"][f"k_retvr={k_r}"]) == len({k_q, gm[0]}) * len({k_i, ge[0]})
    assert res["other_args"]["n_ment_anchors_vals"] == [k_q, gm[0]]
    # the CLI: pickle in, reference-layout JSON out, matrix rank next to it
    F.save_m2e_pickle(str(tmp_path / "m2e.pkl"), F.make_m2e_dict(A, [{}] * A.shape[0], [[0]] * A.shape[0]))
    f = main(["--m2e_file", str(tmp_path / "m2e.pkl"), "--res_dir", str(tmp_path / "res"), "--methods", "cur", "--k_q", str(gm[0]),
              "--k_i", str(ge[0]), "--rank"])
    d = json.load(open(f))
    assert f.endswith(f"nm={A.shape[0]}_ne={A.shape[1]}_s=1/retrieval_wrt_exact_crossenc.json")
    pt = d["cur"]["top_k=10"]["k_retvr=500"][f"anc_n_m={gm[0]}~anc_n_e={ge[0]}"]
    assert set(pt) == {"anchor", "non_anchor", "all"} and 0.0 < pt["all"]["approx_error_relative"] < 2.0
    assert d["other_args"]["matrix_rank"] == int(np.linalg.matrix_rank(g["A"])) and d["other_args"]["timing"]["grid_points"] == 1


def test_fixed_split_eval_dropin_matches_reference_golden(golden_dir):
    from anncur_b200 import eval_approx_score_mat, eval_approx_score_mat_for_all_topk, fixed_split_cur_eval
    g = _load(golden_dir, "fixed_split_eval")
    ref = json.loads(str(g["results_json"]))
    test = torch.from_numpy(g["test"])
    top_k_vals, k_r_vals, k_i_vals = g["top_k_vals"].tolist(), g["k_r_vals"].tolist(), g["k_i_vals"].tolist()
    for k_i in k_i_vals:
        approx = torch.from_numpy(g[f"approx_{k_i}"])       # the reference's own approximate scores
        for k_r in k_r_vals:
            ours = eval_approx_score_mat_for_all_topk(test, approx, top_k_vals, k_r)
            want = ref[f"all_topk|k_i={k_i}|k_r={k_r}"]
            assert sorted(str(k) for k in ours) == sorted(want)
            for k in ours:
                _close_dict(ours[k], want[str(k)], atol=1e-6)     # same inputs -> identical metrics
            key = f"single|k_i={k_i}|k_r={k_r}|k={min(top_k_vals)}"
            if key in ref:
                _close_dict(eval_approx_score_mat(test, approx, min(top_k_vals), k_r), ref[key], atol=1e-6)
    # whole pipeline (anchor replay + index build + fused retrieval + rerank) against the same golden numbers
    full = fixed_split_cur_eval(torch.from_numpy(g["train"]), test, k_i_vals, top_k_vals, k_r_vals, int(g["seed"]))
    n_train = g["train"].shape[0]
    for k_i in k_i_vals:
        for k_r in k_r_vals:
            want = ref[f"all_topk|k_i={k_i}|k_r={k_r}"]
            for k in want:
                _close_dict(full[f"top_k={k}"][f"k_retvr={k_r}"][f"anc_n_m={n_train}_anc_n_e={k_i}"], want[k], atol=3e-3)


def test_compute_overlap_dropin_strings(golden_dir):
    from anncur_b200 import compute_overlap
    g = _load(golden_dir, "overlap_strings")
    assert {k: list(v) for k, v in compute_overlap(g["a"], g["b"]).items()} == json.loads(str(g["res_json"]))
    assert {k: list(v) for k, v in compute_overlap([], []).items()} == json.loads(str(g["empty_json"]))


def test_flat_index_dropin_matches_faiss_convention():
    from anncur_b200 import build_flat_or_ivff_index
    emb = O.synthetic_scores(12000, 48, rank=8, seed=1)              # > 11000 rows: the reference would pick IVF
    x = O.synthetic_scores(9, 48, rank=8, seed=2)
    index = build_flat_or_ivff_index(emb, force_exact_search=False)
    D, I = index.search(x, 10)
    wantD, wantI = O.flat_ip_search(emb, x, 10)
    assert D.dtype == np.float32 and I.dtype == np.int64 and D.shape == (9, 10)
    assert_topk_sets_match(I, wantI, full_scores=x @ emb.T)
    assert_scores_close(D, wantD)
    D, I = build_flat_or_ivff_index(torch.from_numpy(emb[:7]), True).search(x[:1], 10)      # k > N padding
    assert (I[0, 7:] == -1).all() and (D[0, 7:] == -np.finfo(np.float32).max).all()
    D1, I1 = build_flat_or_ivff_index(list(emb[:100]), True).search(x[:1], 5)               # list input (:29-32)
    assert I1.shape == (1, 5)
