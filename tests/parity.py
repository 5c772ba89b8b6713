"""Shared parity rules (BASELINE.json north_star / SURVEY.md section 8d).

* fp32 scores: |ours - ref| <= REL_TOL * max|ref row|            (REL_TOL = 1e-4)
* top-k index SETS identical, except that items whose reference score lies within
  tau = REL_TOL * max|score row| of the k-th score may be swapped (tie order is unspecified in
  torch.topk / faiss).
"""
import numpy as np

REL_TOL = 1e-4


def assert_scores_close(ours, ref, rel=REL_TOL, what="scores"):
    ours = np.asarray(ours, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert ours.shape == ref.shape, f"{what}: shape {ours.shape} != {ref.shape}"
    if ours.size == 0:
        return
    scale = np.maximum(np.abs(ref).max(axis=-1, keepdims=True), 1e-30) if ref.ndim > 1 else max(np.abs(ref).max(), 1e-30)
    err = np.abs(ours - ref) / scale
    assert err.max() <= rel, f"{what}: max rel err {err.max():.3e} > {rel:g}"


def assert_topk_sets_match(our_idx, ref_idx, full_scores=None, ref_vals=None, rel=REL_TOL, what="topk"):
    """Row-wise set equality with the near-tie exemption.

    full_scores: optional dense [B x N] reference scores used to look up the score of any item
    either side returned; without it ref_vals ([B x k] scores of ref_idx) bounds the k-th score and
    the exemption only applies to reference-side items."""
    our_idx = np.asarray(our_idx)
    ref_idx = np.asarray(ref_idx)
    assert our_idx.shape == ref_idx.shape, f"{what}: shape {our_idx.shape} != {ref_idx.shape}"
    bad_rows = []
    for r in range(our_idx.shape[0]):
        a, b = set(our_idx[r].tolist()), set(ref_idx[r].tolist())
        assert len(a) == our_idx.shape[1], f"{what}: row {r} has duplicate indices"
        if a == b:
            continue
        diff = (a - b) | (b - a)
        if full_scores is None:
            bad_rows.append((r, sorted(diff)[:6]))
            continue
        row = np.asarray(full_scores[r], dtype=np.float64)
        kth = np.sort(row[np.asarray(ref_idx[r])])[0]
        tau = rel * max(np.abs(row).max(), 1e-30)
        if any(abs(row[j] - kth) > tau for j in diff):
            bad_rows.append((r, sorted(diff)[:6]))
    assert not bad_rows, f"{what}: {len(bad_rows)} rows differ beyond the tie tolerance, e.g. {bad_rows[:3]}"


def assert_sorted_desc(vals, what="values"):
    v = np.asarray(vals)
    if v.shape[-1] > 1:
        assert (v[..., :-1] >= v[..., 1:]).all(), f"{what} not sorted best-first"
