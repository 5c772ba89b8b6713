"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's ANNCUR test-time search path.

Plain numpy / CPU-torch fp32, the same third-party primitives the reference leans on
(``np.linalg.pinv`` with its legacy ``rcond=1e-15``, ``torch.matmul``, ``torch.topk``), arranged as
stand-alone functions so that tests can call one step at a time.  Every function names the
reference lines it restates (paths relative to the reference root).

Pinned against: the reference's own functions run in the build container
(tests/test_oracle_vs_reference.py) and tests/golden/*.npz (tests/test_oracle_golden.py).
``adaptive_anncur`` is the exception -- **parity unpinned**, the reference has no such code.

Never imported by ``anncur_b200``.
"""
from itertools import product  # noqa: F401  (kept for grid helpers used by tests)

import numpy as np
import torch

RERANK_FILL = -99999999999999.0  # the literal subtracted from zeros at
#                                  eval/run_retrieval_eval_wrt_exact_crossenc.py:110
OVERLAP_METRICS = ("common", "diff", "total", "common_frac", "diff_frac")


# --------------------------------------------------------------------------------------------
# synthetic inputs -- shared by tests and bench so both sides see one matrix
# --------------------------------------------------------------------------------------------
def synthetic_scores(n_rows, n_items, rank=64, noise=0.05, seed=0, dtype=np.float32):
    """A = X.Y^T/sqrt(r) + noise*G, numpy Generator streams (stable across numpy versions)."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n_rows, rank), dtype=np.float32)
    Y = rng.standard_normal((n_items, rank), dtype=np.float32)
    G = rng.standard_normal((n_rows, n_items), dtype=np.float32)
    A = (X @ Y.T) / np.float32(np.sqrt(rank)) + np.float32(noise) * G
    return A.astype(dtype)


# --------------------------------------------------------------------------------------------
# A1/A2: CUR factorisation  (eval/matrix_approx_zeshel.py:21-69)
# --------------------------------------------------------------------------------------------
def strictly_increasing(idx):
    """eval/matrix_approx_zeshel.py:53-55 (``_is_sorted``: strict ``<`` between neighbours)."""
    idx = list(idx)
    return all(a < b for a, b in zip(idx[:-1], idx[1:]))


def pinv_f32(mat):
    """eval/matrix_approx_zeshel.py:49 -- ``np.linalg.pinv`` on the fp32 array, default rcond 1e-15."""
    arr = mat.numpy() if torch.is_tensor(mat) else np.asarray(mat)
    return torch.from_numpy(np.linalg.pinv(arr))


class CurFactors:
    """Value object for what ``CURApprox.__init__`` leaves on ``self`` (matrix_approx_zeshel.py:25-51)."""

    def __init__(self, n, m, row_idxs, col_idxs, C, R, U, latent_rows, latent_cols, approx_preference):
        self.n, self.m = n, m
        self.row_idxs, self.col_idxs = row_idxs, col_idxs
        self.C, self.R, self.U = C, R, U
        self.latent_rows, self.latent_cols = latent_rows, latent_cols
        self.approx_preference = approx_preference


def cur_build(rows, cols, row_idxs, col_idxs, approx_preference, A=None, check=True):
    """eval/matrix_approx_zeshel.py:21-51 and :57-69.

    ``check=True`` performs the checks the reference *intends* (sorted indices :36-37, lengths
    :39-40, equal intersection :44 with ``torch.equal`` instead of the broken ``torch.eq`` assert).
    """
    rows = torch.as_tensor(rows)
    cols = torch.as_tensor(cols)
    if check:
        if not strictly_increasing(row_idxs):
            raise AssertionError("row_idxs should be sorted")
        if not strictly_increasing(col_idxs):
            raise AssertionError("col_idxs should be sorted")
        if len(row_idxs) != rows.shape[0] or len(col_idxs) != cols.shape[1]:
            raise AssertionError("index list lengths must match rows/cols")
    ridx = torch.as_tensor(np.asarray(row_idxs, dtype=np.int64))
    cidx = torch.as_tensor(np.asarray(col_idxs, dtype=np.int64))
    intersect = cols[ridx, :]                                            # :42  (k_q x k_i)
    if check and not torch.equal(intersect, rows[:, cidx]):
        raise AssertionError("Invalid rows and cols as their intersection does not match")
    if A is not None:                                                    # :46-47 cur_oracle
        U = pinv_f32(cols) @ torch.as_tensor(A) @ pinv_f32(rows)
    else:                                                                # :49
        U = pinv_f32(intersect)
    if approx_preference == "cols":                                      # :60-62
        latent_rows, latent_cols = cols @ U, rows
    elif approx_preference == "rows":                                    # :63-65
        latent_rows, latent_cols = cols, U @ rows
    else:                                                                # :66-67
        raise NotImplementedError(f"approx_preference = {approx_preference} not supported")
    return CurFactors(cols.shape[0], rows.shape[1], row_idxs, col_idxs, cols, rows, U,
                      latent_rows, latent_cols, approx_preference)


def get_rows(f, row_idxs):
    """eval/matrix_approx_zeshel.py:71-75."""
    return f.latent_rows[_as_index(row_idxs), :] @ f.latent_cols


def get_cols(f, col_idxs):
    """eval/matrix_approx_zeshel.py:77-80."""
    return f.latent_rows @ f.latent_cols[:, _as_index(col_idxs)]


def get(f, row_idxs, col_idxs):
    """eval/matrix_approx_zeshel.py:82-86."""
    return f.latent_rows[_as_index(row_idxs), :] @ f.latent_cols[:, _as_index(col_idxs)]


def get_complete_row(f, sparse_rows):
    """A3 -- eval/matrix_approx_zeshel.py:109-119: scores = Q (B x k_i) . E (k_i x N)."""
    if f.approx_preference != "rows":
        raise NotImplementedError("index built with approx_preference != rows")
    return torch.as_tensor(sparse_rows) @ f.latent_cols


def topk_in_row(f, sparse_rows, k):
    """A4 -- eval/matrix_approx_zeshel.py:121-126: torch.topk(scores, k, dim=1)."""
    return torch.topk(get_complete_row(f, sparse_rows), k, dim=1)


def get_complete_col(f, sparse_cols):
    """eval/matrix_approx_zeshel.py:88-98."""
    if f.approx_preference != "cols":
        raise NotImplementedError("index built with approx_preference != cols")
    return f.latent_rows @ torch.as_tensor(sparse_cols)


def topk_in_col(f, sparse_cols, k):
    """eval/matrix_approx_zeshel.py:100-106 (note: top-k along dim=1 of an (n x *) matrix, as shipped)."""
    return torch.topk(get_complete_col(f, sparse_cols), k, dim=1)


def score_topk(Q, E, k):
    """A3+A4 on bare tensors: what ``CURApprox.topk_in_row`` computes, = the bench's CPU arm."""
    return torch.topk(torch.as_tensor(Q) @ torch.as_tensor(E), k, dim=1)


def _as_index(idx):
    if torch.is_tensor(idx):
        return idx.long()
    return torch.as_tensor(np.asarray(idx, dtype=np.int64))


# --------------------------------------------------------------------------------------------
# flat inner-product index (models/nearest_nbr.py:24-38, faiss.IndexFlatIP -- third-party, unpinned)
# --------------------------------------------------------------------------------------------
def flat_ip_search(embeds, x, k):
    """What ``build_flat_or_ivff_index(embeds, force_exact_search=True).search(x, k)`` returns:
    (D float32 [nq x k], I int64 [nq x k]) best-first; faiss pads with (-FLT_MAX, -1) when k > N."""
    embeds = np.ascontiguousarray(np.asarray(embeds, dtype=np.float32))
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    n = embeds.shape[0]
    kk = min(k, n)
    vals, idx = torch.topk(torch.from_numpy(x) @ torch.from_numpy(embeds).T, kk, dim=1)
    D = np.full((x.shape[0], k), -np.finfo(np.float32).max, dtype=np.float32)
    I = np.full((x.shape[0], k), -1, dtype=np.int64)
    D[:, :kk], I[:, :kk] = vals.numpy(), idx.numpy()
    return D, I


# --------------------------------------------------------------------------------------------
# A5: per-query retrieve + rerank loop
#     eval/run_retrieval_eval_wrt_exact_crossenc.py:97-117
#     eval/run_retrieval_eval_wrt_exact_crossenc_w_fixed_train_test_splits.py:80-100 and :160-180
# --------------------------------------------------------------------------------------------
def retrieve_and_rerank(exact, approx, top_k, top_k_retvr):
    """Returns dict of three (indices int64 [n x *], scores fp32 [n x *]) pairs:
    exact top-k, approx top-k_r, and exact-score rerank of the k_r retrieved items cut to top-k."""
    exact = torch.as_tensor(exact)
    approx = torch.as_tensor(approx)
    n = exact.shape[0]
    ex_i, ex_s, ap_i, ap_s, rr_i, rr_s = [], [], [], [], [], []
    for q in range(n):
        row, arow = exact[q], approx[q]
        s, i = row.topk(top_k)
        a_s, a_i = arow.topk(top_k_retvr)
        masked = torch.zeros(row.shape) + RERANK_FILL
        masked[a_i] = row[a_i]
        r_s, r_i = masked.topk(top_k)
        ex_i.append(i), ex_s.append(s), ap_i.append(a_i), ap_s.append(a_s), rr_i.append(r_i), rr_s.append(r_s)

    def pack(ii, ss):
        if n == 0:
            return np.zeros((0, 0), np.int64), np.zeros((0, 0), np.float32)
        return torch.stack(ii).numpy(), torch.stack(ss).numpy()

    return {"exact": pack(ex_i, ex_s), "approx": pack(ap_i, ap_s), "reranked": pack(rr_i, rr_s)}


# --------------------------------------------------------------------------------------------
# A6: overlap metrics  (eval/eval_utils.py:115-150)
# --------------------------------------------------------------------------------------------
def overlap_one(indices1, indices2):
    """eval/eval_utils.py:139-150."""
    if len(indices1) != len(indices2):
        raise AssertionError(f"Len of both indices is not same => {len(indices1)} != {len(indices2)}")
    n = len(indices1)
    common = len(set(np.asarray(indices1).tolist()) & set(np.asarray(indices2).tolist()))
    return {"common": common, "diff": n - common, "total": n,
            "common_frac": common / n, "diff_frac": (n - common) / n}


def compute_overlap(indices_list1, indices_list2):
    """eval/eval_utils.py:115-136 -- returns the reference's formatted strings."""
    per_row = [overlap_one(a, b) for a, b in zip(indices_list1, indices_list2)]
    if not per_row:
        return {m: ("mean 0.0", "std 0.0", "p50 0.0") for m in OVERLAP_METRICS}
    out = {}
    for m in OVERLAP_METRICS:
        vals = [r[m] for r in per_row]
        out[m] = ("mean {:.4f}".format(np.mean(vals)), "std {:.4f}".format(np.std(vals)),
                  "p50 {:.4f}".format(np.percentile(vals, 50)))
    return out


def overlap_floats(indices_list1, indices_list2, prefix="exact_vs_reranked_approx_retvr"):
    """The string -> float re-parse the callers do (..._fixed_train_test_splits.py:116-127,
    run_retrieval_eval_wrt_exact_crossenc.py:125-141): ``{prefix}~{metric}_{stat}`` -> float."""
    strings = compute_overlap(indices_list1, indices_list2)
    flat = {}
    for m, (mean, std, p50) in strings.items():
        flat[f"{prefix}~{m}_mean"] = float(mean[5:])
        flat[f"{prefix}~{m}_std"] = float(std[4:])
        flat[f"{prefix}~{m}_p50"] = float(p50[4:])
    return flat


# --------------------------------------------------------------------------------------------
# eval drivers
# --------------------------------------------------------------------------------------------
def sample_anchors(n_ments, n_ents, n_ment_anchors, n_ent_anchors, seed):
    """eval/run_retrieval_eval_wrt_exact_crossenc.py:65-70: rows first, then columns, one Generator."""
    rng = np.random.default_rng(seed=seed)
    rows = sorted(rng.choice(n_ments, size=n_ment_anchors, replace=False))
    cols = sorted(rng.choice(n_ents, size=n_ent_anchors, replace=False))
    return rows, cols


def run_approx_eval_w_seed(approx_method, all_scores, n_ment_anchors, n_ent_anchors, top_k, top_k_retvr,
                           seed, precomp_approx=None):
    """eval/run_retrieval_eval_wrt_exact_crossenc.py:47-158."""
    all_scores = torch.as_tensor(all_scores)
    n_ments, n_ents = all_scores.shape
    row_idxs, col_idxs = sample_anchors(n_ments, n_ents, n_ment_anchors, n_ent_anchors, seed)
    rows = all_scores[_as_index(row_idxs), :]
    cols = all_scores[:, _as_index(col_idxs)]
    non_anchor = sorted(set(range(n_ments)) - set(row_idxs))
    if approx_method in ("bienc", "fixed_anc_ent") or approx_method.startswith("fixed_anc_ent_cur_"):
        approx = torch.as_tensor(precomp_approx)
    elif approx_method == "cur":
        f = cur_build(rows, cols, row_idxs, col_idxs, "rows", check=False)
        approx = get(f, list(range(n_ments)), list(range(n_ents)))
    elif approx_method == "cur_oracle":
        f = cur_build(rows, cols, row_idxs, col_idxs, "rows", A=all_scores, check=False)
        approx = get(f, list(range(n_ments)), list(range(n_ents)))
    else:
        raise NotImplementedError(f"approx_method = {approx_method} not supported")
    lists = retrieve_and_rerank(all_scores, approx, top_k, top_k_retvr)
    exact_idx, rr_idx = lists["exact"][0], lists["reranked"][0]

    def score(subset):
        sub = np.asarray(subset, dtype=np.int64)
        res = overlap_floats(exact_idx[sub], rr_idx[sub])
        tsub = torch.as_tensor(sub)
        err = torch.norm((approx - all_scores)[tsub, :]).numpy()          # :146
        res["approx_error"] = err
        res["approx_error_relative"] = err / torch.norm(all_scores[tsub, :]).numpy()   # :147
        return res

    return {"anchor": score(row_idxs), "non_anchor": score(non_anchor), "all": score(list(range(n_ments)))}


def eval_approx_score_mat_for_all_topk(all_scores, approx_scores, arg_top_k_vals, top_k_retvr):
    """..._w_fixed_train_test_splits.py:51-135."""
    top_k_vals = [k for k in arg_top_k_vals if k <= top_k_retvr]                       # :72
    if not top_k_vals:
        return {}
    max_k = max(top_k_vals)
    lists = retrieve_and_rerank(all_scores, approx_scores, max_k, top_k_retvr)
    exact_idx, rr_idx = lists["exact"][0], lists["reranked"][0]
    return {k: overlap_floats(exact_idx[:, :k], rr_idx[:, :k]) for k in top_k_vals}


def eval_approx_score_mat(all_scores, approx_scores, top_k, top_k_retvr):
    """..._w_fixed_train_test_splits.py:138-206."""
    lists = retrieve_and_rerank(all_scores, approx_scores, top_k, top_k_retvr)
    return overlap_floats(lists["exact"][0], lists["reranked"][0])


def fixed_split_cur_scores(train_scores, test_scores, n_ent_anchors_vals, seed):
    """..._w_fixed_train_test_splits.py:286-303: ONE generator reused across the k_i grid;
    every training row is an anchor query.  Returns {k_i: (anchor_ent_idxs, approx test scores)}."""
    train_scores = torch.as_tensor(train_scores)
    test_scores = torch.as_tensor(test_scores)
    n_train, n_ents = train_scores.shape
    rng = np.random.default_rng(seed=seed)
    out = {}
    for k_i in n_ent_anchors_vals:
        anc = sorted(rng.choice(n_ents, size=k_i, replace=False))
        cols = train_scores[:, _as_index(anc)]
        f = cur_build(train_scores, cols, np.arange(n_train), anc, "rows", check=False)
        out[k_i] = (anc, get_complete_row(f, test_scores[:, _as_index(anc)]))
    return out


def fixed_anc_ent_scores(test_scores, ent_to_ent_scores, topk_ents_row0, n_fixed_anc_ent):
    """..._w_fixed_train_test_splits.py:321-324 (method ``fixed_anc_ent``; ``bienc`` :283 and ``tfidf`` :383 are the same
    product with model embeddings): entities are embedded by their scores against the first n fixed anchor entities,
    mentions by their exact scores against the same anchors; approx = mention_embeds @ ent_embeds.T."""
    test_scores = torch.as_tensor(test_scores)
    anchor_ent_idxs = _as_index(np.asarray(topk_ents_row0)[:n_fixed_anc_ent])          # :321
    ent_embeds = torch.as_tensor(ent_to_ent_scores)[:, :n_fixed_anc_ent]                # :322
    mention_embeds = test_scores[:, anchor_ent_idxs]                                    # :323
    return mention_embeds @ ent_embeds.T                                                # :324


def fixed_anc_ent_cur_scores(test_scores, ent_to_ent_scores, n_fixed_anc_ent, n_ent_anchors_vals, seed=0):
    """..._w_fixed_train_test_splits.py:340-358 (method ``fixed_anc_ent_cur``): R = the dump transposed (n_fixed x N); ONE
    generator (seed 0) across the anchor-count grid; U = pinv(R[:, anchors]) (numpy, fp32), approx = test[:, anchors] @ (U @ R).
    Returns {n_anc_ent: approx test scores}."""
    test_scores = torch.as_tensor(test_scores)
    R = torch.as_tensor(ent_to_ent_scores)[:, :n_fixed_anc_ent].T                       # :340
    n_ents = R.shape[1]
    rng = np.random.default_rng(seed=seed)                                              # :342
    out = {}
    for n_anc_ent in n_ent_anchors_vals:
        anc = sorted(rng.choice(n_ents, size=n_anc_ent, replace=False))                 # :347
        idx = _as_index(anc)
        U = torch.tensor(np.linalg.pinv(R[:, idx]))                                     # :349-350
        out[n_anc_ent] = test_scores[:, idx] @ (U @ R)                                  # :351-357
    return out


# --------------------------------------------------------------------------------------------
# sharded search -- restated on CPU so the merge can be checked
# --------------------------------------------------------------------------------------------
def merge_topk(cand_vals, cand_idx, k):
    """Best-k of per-shard candidate lists.  cand_vals/idx: [P, B, k_p]; ties -> lower index first;
    entries with idx < 0 are padding."""
    v = np.concatenate(list(cand_vals), axis=1).astype(np.float32)
    i = np.concatenate(list(cand_idx), axis=1).astype(np.int64)
    v = np.where(i < 0, -np.inf, v)
    order = np.lexsort((i, -v.astype(np.float64)), axis=1)[:, :k]
    return np.take_along_axis(v, order, 1), np.take_along_axis(i, order, 1)


def sharded_score_topk(Q, E, k, n_shards):
    Q = torch.as_tensor(Q)
    E = torch.as_tensor(E)
    N = E.shape[1]
    bounds = [(p * N) // n_shards for p in range(n_shards + 1)]
    vals, idxs = [], []
    for p in range(n_shards):
        lo, hi = bounds[p], bounds[p + 1]
        kk = min(k, hi - lo)
        v, i = torch.topk(Q @ E[:, lo:hi], kk, dim=1)
        vp = np.full((Q.shape[0], k), -np.inf, np.float32)
        ip = np.full((Q.shape[0], k), -1, np.int64)
        vp[:, :kk], ip[:, :kk] = v.numpy(), i.numpy() + lo
        vals.append(vp), idxs.append(ip)
    return merge_topk(vals, idxs, k)


# --------------------------------------------------------------------------------------------
# A8: adaptive multi-round ANNCUR -- PARITY UNPINNED (no reference code)
# --------------------------------------------------------------------------------------------
def adaptive_anncur(R_anc, exact_rows, first_anchors, n_rounds, k_per_round, top_k, rcond=1e-15):
    """Per query q, rounds t = 1..T with growing anchor-item set I_t (|I_t| = t * k_per_round):
        c_q  = exact_rows[q, I_t]                      (stands in for CE(q, I_t))
        e_q  = c_q . pinv(R_anc[:, I_t])               (1 x k_q, min-norm least squares)
        s_q  = e_q . R_anc  with I_t masked to -inf    (1 x N)
        I_{t+1} = I_t  U  top-k_per_round(s_q)
    Round 1 uses ``first_anchors`` (shared by all queries).  After the last round the answer is
    the top-``top_k`` of I_T by exact score.  Returns (anchor sets [B x T*k_per_round] int64 in
    selection order, final idx [B x top_k], final exact scores [B x top_k], last-round approx scores)."""
    R = np.asarray(R_anc, dtype=np.float32)
    X = np.asarray(exact_rows, dtype=np.float32)
    B, N = X.shape
    first = np.asarray(first_anchors, dtype=np.int64)
    assert first.shape[0] == k_per_round
    anchors = np.zeros((B, n_rounds * k_per_round), dtype=np.int64)
    last_scores = np.zeros((B, N), dtype=np.float32)
    for q in range(B):
        cur = first.copy()
        for t in range(n_rounds):
            M = R[:, cur]                                   # k_q x m
            c = X[q, cur]                                   # m
            e = c @ np.linalg.pinv(M.astype(np.float64), rcond=rcond)   # k_q  (fp64 solve)
            s = (e @ R.astype(np.float64)).astype(np.float32)
            s[cur] = -np.inf
            last_scores[q] = s
            if t + 1 < n_rounds:
                nxt = np.lexsort((np.arange(N), -s.astype(np.float64)))[:k_per_round]
                cur = np.concatenate([cur, nxt])
        anchors[q] = cur
    ex = np.take_along_axis(X, anchors, 1)
    order = np.lexsort((anchors, -ex.astype(np.float64)), axis=1)[:, :top_k]
    return anchors, np.take_along_axis(anchors, order, 1), np.take_along_axis(ex, order, 1), last_scores
