"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the REFERENCE'S OWN functions
(``oracle/ref_shim.py``) on small seeded matrices, in the build container.

    python oracle/make_golden.py            # needs /root/reference

The reference has no tests, golden vectors or fixtures of its own, so these
files are the pin: inputs are stored bit-for-bit next to the reference's outputs, because a
regenerated ``X @ Y.T`` can differ in the last bit between BLAS builds.
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.cur_oracle import synthetic_scores  # noqa: E402
from oracle.ref_shim import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _flat(d):
    return {k: float(v) for k, v in d.items()}


def case_curapprox(ref, name, n, N, k_q, k_i, seed, B, k):
    """CURApprox attributes + every public method (eval/matrix_approx_zeshel.py:21-126)."""
    A = torch.from_numpy(synthetic_scores(n, N, rank=16, noise=0.05, seed=seed))
    rng = np.random.default_rng(seed)
    row_idxs = sorted(rng.choice(n, size=k_q, replace=False).tolist())
    col_idxs = sorted(rng.choice(N, size=k_i, replace=False).tolist())
    rows, cols = A[row_idxs, :], A[:, col_idxs]
    test_rows = [i for i in range(n) if i not in set(row_idxs)][:B]
    Q = A[test_rows][:, col_idxs].contiguous()
    out = {"A": A.numpy(), "row_idxs": np.array(row_idxs), "col_idxs": np.array(col_idxs),
           "test_rows": np.array(test_rows), "k": np.array(k)}
    ap = ref.CURApprox(rows=rows, cols=cols, row_idxs=row_idxs, col_idxs=col_idxs, approx_preference="rows")
    out["rows_U"] = ap.U.numpy()
    out["rows_latent_cols"] = ap.latent_cols.numpy()
    out["rows_get_complete_row"] = ap.get_complete_row(Q).numpy()
    tk = ap.topk_in_row(Q, k)
    out["rows_topk_values"], out["rows_topk_indices"] = tk.values.numpy(), tk.indices.numpy()
    sel_r, sel_c = list(range(0, n, 3)), list(range(0, N, 7))
    out["sel_r"], out["sel_c"] = np.array(sel_r), np.array(sel_c)
    out["rows_get"] = ap.get(sel_r, sel_c).numpy()
    out["rows_get_rows"] = ap.get_rows(sel_r).numpy()
    out["rows_get_cols"] = ap.get_cols(sel_c).numpy()
    ac = ref.CURApprox(rows=rows, cols=cols, row_idxs=row_idxs, col_idxs=col_idxs, approx_preference="cols")
    out["cols_latent_rows"] = ac.latent_rows.numpy()
    sparse_cols = A[row_idxs][:, sel_c[:5]].contiguous()            # (k_q x 5): anchor-row values of 5 columns
    out["cols_get_complete_col"] = ac.get_complete_col(sparse_cols).numpy()
    tkc = ac.topk_in_col(sparse_cols, 3)
    out["cols_topk_values"], out["cols_topk_indices"] = tkc.values.numpy(), tkc.indices.numpy()
    out["cond_intersect"] = np.array(np.linalg.cond(cols[row_idxs, :].numpy().astype(np.float64)))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    return name


def case_sweep(ref, name, n, N, k_q, k_i, top_k, k_r, seed):
    """run_approx_eval_w_seed for cur and cur_oracle (eval/run_retrieval_eval_wrt_exact_crossenc.py:47-158)."""
    A = torch.from_numpy(synthetic_scores(n, N, rank=16, noise=0.05, seed=seed + 100))
    out = {"A": A.numpy(), "params": np.array([k_q, k_i, top_k, k_r, seed])}
    for method in ("cur", "cur_oracle"):
        res = ref.run_approx_eval_w_seed(method, A, k_q, k_i, top_k, k_r, seed, None)
        out[method + "_json"] = np.array(json.dumps({g: _flat(r) for g, r in res.items()}))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    return name


def case_fixed_split(ref, name, n_train, n_test, N, k_i_vals, seed, top_k_vals, k_r_vals):
    """The 'cur' branch of run_eval_method (..._w_fixed_train_test_splits.py:286-303) replayed with the
    reference's CURApprox, then eval_approx_score_mat_for_all_topk (:51-135) / eval_approx_score_mat (:138-206)."""
    A = torch.from_numpy(synthetic_scores(n_train + n_test, N, rank=16, noise=0.05, seed=seed + 200))
    train, test = A[:n_train].contiguous(), A[n_train:].contiguous()
    rng = np.random.default_rng(seed=seed)
    out = {"train": train.numpy(), "test": test.numpy(), "k_i_vals": np.array(k_i_vals),
           "top_k_vals": np.array(top_k_vals), "k_r_vals": np.array(k_r_vals), "seed": np.array(seed)}
    results = {}
    for k_i in k_i_vals:
        anc = sorted(rng.choice(N, size=k_i, replace=False))
        ap = ref.CURApprox(row_idxs=np.arange(n_train), col_idxs=anc, rows=train, cols=train[:, anc],
                           approx_preference="rows")
        approx = ap.get_complete_row(sparse_rows=test[:, anc])
        out[f"anchors_{k_i}"] = np.array(anc)
        out[f"approx_{k_i}"] = approx.numpy()
        for k_r in k_r_vals:
            r = ref.eval_approx_score_mat_for_all_topk(test, approx, top_k_vals, k_r)
            results[f"all_topk|k_i={k_i}|k_r={k_r}"] = {str(k): _flat(v) for k, v in r.items()}
            kk = min(top_k_vals)
            if kk <= k_r:
                results[f"single|k_i={k_i}|k_r={k_r}|k={kk}"] = _flat(ref.eval_approx_score_mat(test, approx, kk, k_r))
    out["results_json"] = np.array(json.dumps(results))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    return name


def case_overlap(ref, name):
    """compute_overlap strings (eval/eval_utils.py:115-150), including the empty-input branch."""
    rng = np.random.default_rng(7)
    a = np.stack([rng.choice(500, size=20, replace=False) for _ in range(37)])
    b = np.stack([rng.choice(500, size=20, replace=False) for _ in range(37)])
    b[:, :8] = a[:, 5:13]
    res = ref.compute_overlap(a, b)
    empty = ref.compute_overlap([], [])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), a=a, b=b,
                        res_json=np.array(json.dumps({k: list(v) for k, v in res.items()})),
                        empty_json=np.array(json.dumps({k: list(v) for k, v in empty.items()})))
    return name


def main():
    warnings.filterwarnings("ignore", category=UserWarning)
    torch.set_num_threads(1)          # single-thread MKL: one summation order for the stored outputs
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    made = [
        case_curapprox(ref, "curapprox_tall", n=160, N=1100, k_q=60, k_i=24, seed=1, B=48, k=25),
        case_curapprox(ref, "curapprox_square", n=120, N=900, k_q=32, k_i=32, seed=2, B=40, k=10),
        case_curapprox(ref, "curapprox_wide", n=100, N=700, k_q=20, k_i=48, seed=3, B=33, k=100),
        case_sweep(ref, "sweep_eval", n=90, N=800, k_q=40, k_i=25, top_k=10, k_r=60, seed=0),
        case_fixed_split(ref, "fixed_split_eval", n_train=70, n_test=40, N=900, k_i_vals=[10, 25, 50],
                         seed=0, top_k_vals=[1, 10, 50], k_r_vals=[5, 50, 200]),
        case_overlap(ref, "overlap_strings"),
    ]
    meta = {"generated_by": "oracle/make_golden.py", "numpy": np.__version__, "torch": torch.__version__,
            "reference": "iesl/anncur @ /root/reference (python, asserts stripped)", "cases": made}
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
