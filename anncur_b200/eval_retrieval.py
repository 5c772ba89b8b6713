"""Drop-ins for the per-query retrieve / rerank / overlap evaluation of the reference:

* ``compute_overlap``                        eval/eval_utils.py:115-150
* ``run_approx_eval_w_seed``                 eval/run_retrieval_eval_wrt_exact_crossenc.py:47-158
* ``eval_approx_score_mat_for_all_topk``     ..._w_fixed_train_test_splits.py:51-135
* ``eval_approx_score_mat``                  ..._w_fixed_train_test_splits.py:138-206
* ``_get_indices_scores``                    ..._w_fixed_train_test_splits.py:34-47

Same signatures, same result keys (``exact_vs_reranked_approx_retvr~<metric>_<stat>``, values rounded
to 4 decimals exactly as the reference's format-then-parse does).  The Python per-row loop with three
``topk`` calls and an N-long temporary per query becomes three batched kernels: exact row top-k,
approximate row top-k (or the fused score+top-k when an index is given), and rerank+overlap.
"""
import numpy as np
import torch

from . import engine
from .matrix_approx import CURApprox

METRICS = ("common", "diff", "total", "common_frac", "diff_frac")
PREFIX = "exact_vs_reranked_approx_retvr"


def _get_indices_scores(topk_preds):
    indices, scores = zip(*topk_preds)
    if torch.is_tensor(indices[0]):
        indices, scores = torch.cat(indices), torch.cat(scores)
        return {"indices": indices.cpu().numpy(), "scores": scores.cpu().numpy()}
    return {"indices": np.concatenate(indices), "scores": np.concatenate(scores)}


def _stats_strings(common, k):
    """Reference formatting (eval/eval_utils.py:131-136) from per-row intersection counts."""
    common = np.asarray(common, dtype=np.float64)
    if common.size == 0:
        return {m: ("mean 0.0", "std 0.0", "p50 0.0") for m in METRICS}
    k = np.asarray(k, dtype=np.float64)                        # a scalar, or one list length per row
    per_metric = {"common": common, "diff": k - common, "total": np.broadcast_to(k, common.shape).astype(np.float64),
                  "common_frac": common / k, "diff_frac": (k - common) / k}
    return {m: ("mean {:.4f}".format(np.mean(v)), "std {:.4f}".format(np.std(v)),
                "p50 {:.4f}".format(np.percentile(v, 50))) for m, v in per_metric.items()}


def _flatten(strings):
    out = {}
    for m, (mean, std, p50) in strings.items():
        out[f"{PREFIX}~{m}_mean"] = float(mean[5:])
        out[f"{PREFIX}~{m}_std"] = float(std[4:])
        out[f"{PREFIX}~{m}_p50"] = float(p50[4:])
    return out


def _overlap_rows_gpu(a, b):
    """Per-row |set(a_r) & set(b_r)| for two rectangular (n x k) index arrays: the K6 kernel up to k = 4096, sort + binary
    search on the GPU beyond (torch ops on the device; still no host arithmetic)."""
    if a.shape[1] <= 4096:
        return engine.overlap_counts(a, b).cpu().numpy()
    engine.require_cuda()
    dev = a.device if a.is_cuda else torch.device("cuda", torch.cuda.current_device())
    out = np.zeros(a.shape[0], dtype=np.int64)
    for r in range(a.shape[0]):
        ua, ub = torch.unique(a[r].to(dev)), torch.unique(b[r].to(dev))
        out[r] = int(torch.isin(ua, ub).sum().item())
    return out


def compute_overlap(indices_list1, indices_list2):
    """|set(a) & set(b)| per row pair and the reference's mean/std/p50 strings (eval/eval_utils.py:115-150).  Like the
    reference, rows may differ in length from each other (each PAIR must match, :143): rows are grouped by length and every
    group goes through the GPU kernel."""
    n = min(len(indices_list1), len(indices_list2))                  # zip() semantics of the reference (:123)
    if n == 0:
        return {m: ("mean 0.0", "std 0.0", "p50 0.0") for m in METRICS}
    rect = torch.is_tensor(indices_list1) or (isinstance(indices_list1, np.ndarray) and indices_list1.dtype != object)
    if rect:
        a = indices_list1 if torch.is_tensor(indices_list1) else torch.as_tensor(np.asarray(indices_list1))
        b = indices_list2 if torch.is_tensor(indices_list2) else torch.as_tensor(np.asarray(indices_list2))
        assert a.shape == b.shape, f"Len of both indices is not same => {a.shape[-1]} != {b.shape[-1]}"
        return _stats_strings(_overlap_rows_gpu(a, b), a.shape[1])
    lens = np.array([len(x) for x in indices_list1[:n]])
    for r in range(n):
        assert len(indices_list1[r]) == len(indices_list2[r]), f"Len of both indices is not same => {len(indices_list1[r])} != {len(indices_list2[r])}"
    common = np.zeros(n, dtype=np.int64)
    for L in np.unique(lens):
        rows = np.nonzero(lens == L)[0]
        if L == 0:
            raise ZeroDivisionError("division by zero")               # what the reference's n_intersection / n does (:149)
        a = torch.as_tensor(np.asarray([list(indices_list1[r]) for r in rows], dtype=np.int64))
        b = torch.as_tensor(np.asarray([list(indices_list2[r]) for r in rows], dtype=np.int64))
        common[rows] = _overlap_rows_gpu(a, b)
    return _stats_strings(common, lens)


def retrieve_rerank_overlap(all_scores, approx_scores, k_list, top_k_retvr, *, approx_topk=None):
    """Batched body of the reference's per-query loop.  Returns dict with exact / approx / reranked
    (indices, scores) CUDA tensors and ``common`` [n x len(k_list)] intersection counts.
    ``approx_topk``: optional precomputed (vals, idx) of the approximate top-k_retvr (fused path)."""
    exact = engine._f32(all_scores)
    k_max = max(k_list)
    if k_max > top_k_retvr:
        # the reference re-ranks inside an N-long row of -1e14 (..._w_fixed_train_test_splits.py:91-96), so for top_k > k_retvr
        # it pads the re-ranked list with (top_k - k_retvr) arbitrary never-retrieved items (torch.topk's order among the
        # equal -1e14 entries is unspecified): there is no defined answer to reproduce
        raise ValueError(f"top_k = {k_max} > top_k_retvr = {top_k_retvr}: the reference pads the re-ranked list with unspecified "
                         "items in this case; evaluate with top_k <= top_k_retvr")
    if top_k_retvr > engine.MAX_K:
        raise ValueError(f"top_k_retvr = {top_k_retvr} > {engine.MAX_K}: retrieved lists longer than ANNCUR_MAX_K are not supported")
    ex_v, ex_i = engine.topk_rows(exact, k_max)
    if approx_topk is None:
        ap_v, ap_i = engine.topk_rows(engine._f32(approx_scores, device=exact.device), top_k_retvr)
    else:
        ap_v, ap_i = approx_topk
    rr_i, rr_v, common = engine.rerank_overlap(exact, ap_i, ex_i, k_list)
    return {"exact": (ex_i, ex_v), "approx": (ap_i, ap_v), "reranked": (rr_i, rr_v), "common": common}


def eval_approx_score_mat_for_all_topk(all_ment_to_ent_scores, approx_ment_to_ent_scores, arg_top_k_vals, top_k_retvr,
                                       *, approx_topk=None):
    top_k_vals = [int(k) for k in arg_top_k_vals if k <= top_k_retvr]
    if len(top_k_vals) == 0:
        return {}
    res = retrieve_rerank_overlap(all_ment_to_ent_scores, approx_ment_to_ent_scores, top_k_vals, int(top_k_retvr),
                                  approx_topk=approx_topk)
    common = res["common"].cpu().numpy()
    return {k: _flatten(_stats_strings(common[:, j], k)) for j, k in enumerate(top_k_vals)}


def eval_approx_score_mat(all_ment_to_ent_scores, approx_ment_to_ent_scores, top_k, top_k_retvr, *, approx_topk=None):
    res = retrieve_rerank_overlap(all_ment_to_ent_scores, approx_ment_to_ent_scores, [int(top_k)], int(top_k_retvr),
                                  approx_topk=approx_topk)
    return _flatten(_stats_strings(res["common"].cpu().numpy()[:, 0], int(top_k)))


def run_approx_eval_w_seed(approx_method, all_ment_to_ent_scores, n_ment_anchors, n_ent_anchors, top_k, top_k_retvr,
                           seed, precomp_approx_ment_to_ent_scores=None, *, precision="f32r"):
    engine.require_cuda()
    A = engine._f32(all_ment_to_ent_scores)
    n_ments, n_ents = A.shape
    rng = np.random.default_rng(seed=seed)                                       # reference :65-70
    anchor_ment_idxs = sorted(rng.choice(n_ments, size=n_ment_anchors, replace=False))
    anchor_ent_idxs = sorted(rng.choice(n_ents, size=n_ent_anchors, replace=False))
    rows = A[torch.as_tensor(np.asarray(anchor_ment_idxs, dtype=np.int64), device=A.device), :]
    cols = A[:, torch.as_tensor(np.asarray(anchor_ent_idxs, dtype=np.int64), device=A.device)]
    non_anchor_ment_idxs = sorted(set(range(n_ments)) - set(anchor_ment_idxs))

    approx_topk = None
    err2 = norm2 = None
    if approx_method in ["bienc", "fixed_anc_ent"] or approx_method.startswith("fixed_anc_ent_cur_"):
        approx = engine._f32(precomp_approx_ment_to_ent_scores, device=A.device)
    elif approx_method in ("cur", "cur_oracle"):
        cur = CURApprox(row_idxs=anchor_ment_idxs, col_idxs=anchor_ent_idxs, rows=rows, cols=cols,
                        approx_preference="rows", A=(A if approx_method == "cur_oracle" else None), precision=precision)
        approx = None
        # retrieval straight from the index: scores for all rows never leave the tensor core epilogue
        if top_k_retvr <= engine.MAX_K_FUSED and precision != "f32":
            v, i = engine.score_topk(cur._latent_rows_dev, cur.packed_items(), int(top_k_retvr))
        else:
            v, i = engine.score_topk_f32(cur._latent_rows_dev, cur._latent_cols_dev, int(top_k_retvr))
        approx_topk = (v, i)
        if precision in ("f32r", "f32x3") and 2.0 * A.shape[0] * A.shape[1] * cur._latent_cols_dev.shape[0] >= 2e9:
            err2, norm2 = engine.recon_error_packed(cur._latent_rows_dev, cur.packed_items(), A)      # tcgen05, fp32-grade
        else:
            err2, norm2 = engine.recon_error_rows(cur._latent_rows_dev, cur._latent_cols_dev, A)
    else:
        raise NotImplementedError(f"approx_method = {approx_method} not supported")

    res = retrieve_rerank_overlap(A, approx, [int(top_k)], int(top_k_retvr), approx_topk=approx_topk)
    common = res["common"].cpu().numpy()[:, 0]
    if err2 is None:
        diff = approx - A
        err2 = (diff.double() ** 2).sum(dim=1)
        norm2 = (A.double() ** 2).sum(dim=1)
    err2, norm2 = err2.cpu().numpy(), norm2.cpu().numpy()

    def score(idxs):
        idxs = np.asarray(idxs, dtype=np.int64)
        out = _flatten(_stats_strings(common[idxs], int(top_k)))
        e = float(np.sqrt(err2[idxs].sum()))
        out["approx_error"] = e                                                   # reference :146
        nrm = float(np.sqrt(norm2[idxs].sum()))
        out["approx_error_relative"] = e / nrm if nrm > 0 else float("nan")      # reference :147
        return out

    return {"anchor": score(anchor_ment_idxs), "non_anchor": score(non_anchor_ment_idxs),
            "all": score(list(range(n_ments)))}


def fixed_split_cur_eval(train_scores, test_scores, n_ent_anchors_vals, top_k_vals, top_k_retvr_vals, seed,
                         *, precision="f32r", only_k_i=None):
    """The ``cur`` method of run_eval_method (..._w_fixed_train_test_splits.py:286-303 + :403-429) end to end on
    the GPU: ONE numpy Generator replayed across the k_i grid, every training row an anchor query, and for each
    (k_r, k_i) the all-top-k evaluation.  Returns {f"top_k={k}": {f"k_retvr={k_r}": {f"anc_n_m={n_train}_anc_n_e={k_i}": metrics}}}
    (the reference's key layout, :429)
    restricted to the combinations the reference evaluates (k <= k_r, k_r <= N)."""
    engine.require_cuda()
    train = engine._f32(train_scores)
    test = engine._f32(test_scores, device=train.device)
    n_train, n_ents = train.shape
    rng = np.random.default_rng(seed=seed)
    out = {}
    max_kr = max([kr for kr in top_k_retvr_vals if kr <= n_ents] + [0])
    for k_i in n_ent_anchors_vals:
        anc = sorted(rng.choice(n_ents, size=k_i, replace=False))
        if only_k_i is not None and k_i not in only_k_i:
            continue                                   # draw replayed, grid point not evaluated
        if max_kr == 0:
            continue
        n_test = test.shape[0]
        if k_i == 0:
            # empty anchor set (in the reference's grid): E is 0 x N, every approximate score is 0, and the retrieved
            # list is a pure tie -- ours is items 0..k_r-1 (ties -> lower index; torch.topk leaves the order open)
            i = torch.arange(max_kr, device=test.device, dtype=torch.int64).repeat(n_test, 1)
            v = torch.zeros((n_test, max_kr), device=test.device)
        else:
            anc_t = torch.as_tensor(np.asarray(anc, dtype=np.int64), device=train.device)
            cur = CURApprox(row_idxs=np.arange(n_train), col_idxs=anc, rows=train, cols=train[:, anc_t],
                            approx_preference="rows", precision=precision)
            # one retrieval at the largest k_r: every smaller k_r list is a prefix of it
            v, i = cur.topk_in_row(test[:, anc_t], max_kr)
        for k_r in top_k_retvr_vals:
            if k_r > n_ents or k_r == 0:
                continue
            res = eval_approx_score_mat_for_all_topk(test, None, top_k_vals, k_r, approx_topk=(v[:, :k_r].contiguous(), i[:, :k_r].contiguous()))
            for k, metrics in res.items():
                out.setdefault(f"top_k={k}", {}).setdefault(f"k_retvr={k_r}", {})[f"anc_n_m={n_train}_anc_n_e={k_i}"] = metrics
    return out


def _grid_eval(test, v, i, top_k_vals, top_k_retvr_vals, n_ents):
    """{k_r: {k: metrics}} for every retrieved-list length from ONE retrieval (v, i) at the largest k_r: top-k_r lists are
    prefixes of it."""
    out = {}
    for k_r in top_k_retvr_vals:
        if k_r > n_ents or k_r <= 0:
            continue
        out[k_r] = eval_approx_score_mat_for_all_topk(test, None, top_k_vals, k_r,
                                                      approx_topk=(v[:, :k_r].contiguous(), i[:, :k_r].contiguous()))
    return out


def fixed_split_embed_eval(test_scores, query_embeds, item_embeds, n_train, n_ent_anchors_vals, top_k_vals, top_k_retvr_vals,
                           *, precision="f32r"):
    """The ``bienc`` / ``tfidf`` / ``fixed_anc_ent`` methods of run_eval_method
    (..._w_fixed_train_test_splits.py:257-284, :305-325, :360-385 + :403-429): the approximate score matrix is
    ``query_embeds @ item_embeds.T`` (:283, :324, :383) whatever the anchor count, so it is evaluated once per k_r and the
    result is entered under every ``anc_n_e`` key, as the reference does (:411-417).  Same fused tcgen05 score + top-k
    kernel as ANNCUR, with K = the embedding dimension; the (n_test x N) approximate matrix is never formed."""
    engine.require_cuda()
    test = engine._f32(test_scores)
    Q = engine._f32(query_embeds, device=test.device)
    Et = engine._f32(item_embeds, device=test.device)                   # N x d, as the reference holds them
    n_ents = test.shape[1]
    assert Et.shape[0] == n_ents and Q.shape[0] == test.shape[0] and Q.shape[1] == Et.shape[1], (test.shape, Q.shape, Et.shape)
    max_kr = max([kr for kr in top_k_retvr_vals if kr <= n_ents] + [0])
    out = {}
    if max_kr == 0:
        return out
    d = int(Et.shape[1])
    if precision == "f32" or max_kr > engine.MAX_K_FUSED or (precision == "f32r" and d > engine.MAX_K_DIM_F32R):
        v, i = engine.score_topk_f32(Q, Et.t().contiguous(), max_kr)
    else:
        v, i = engine.score_topk(Q, engine.PackedItems(Et.t().contiguous(), precision), max_kr)
    for k_r, res in _grid_eval(test, v, i, top_k_vals, top_k_retvr_vals, n_ents).items():
        for k, metrics in res.items():
            for k_i in n_ent_anchors_vals:
                out.setdefault(f"top_k={k}", {}).setdefault(f"k_retvr={k_r}", {})[f"anc_n_m={n_train}_anc_n_e={k_i}"] = metrics
    return out


def fixed_split_fixed_anc_cur_eval(test_scores, ent_to_fixed_anchor_scores, n_train, n_ent_anchors_vals, top_k_vals,
                                   top_k_retvr_vals, *, seed=0, precision="f32r", only_k_i=None):
    """The ``fixed_anc_ent_cur`` method (..._w_fixed_train_test_splits.py:327-358): CUR with the entity-to-fixed-anchor
    score matrix in the role of the anchor-query rows, R = ent_to_fixed_anchor_scores.T (n_fixed x N); per anchor count one
    generator draw (seed 0, ONE generator across the grid :341-346), U = pinv(R[:, anchors]) (:349), E = U @ R (:350) and the
    test scores against the anchors as queries (:353-356)."""
    engine.require_cuda()
    test = engine._f32(test_scores)
    R = engine._f32(ent_to_fixed_anchor_scores, device=test.device).t().contiguous()       # n_fixed x N
    n_fixed, n_ents = R.shape
    assert test.shape[1] == n_ents
    rng = np.random.default_rng(seed=seed)
    max_kr = max([kr for kr in top_k_retvr_vals if kr <= n_ents] + [0])
    out = {}
    for k_i in n_ent_anchors_vals:
        anc = sorted(rng.choice(n_ents, size=k_i, replace=False))
        if (only_k_i is not None and k_i not in only_k_i) or max_kr == 0:
            continue
        n_test = test.shape[0]
        if k_i == 0:                                    # empty anchor set: all approximate scores are 0 (see fixed_split_cur_eval)
            i = torch.arange(max_kr, device=test.device, dtype=torch.int64).repeat(n_test, 1)
            v = torch.zeros((n_test, max_kr), device=test.device)
        else:
            anc_t = torch.as_tensor(np.asarray(anc, dtype=np.int64), device=test.device)
            cur = CURApprox(row_idxs=np.arange(n_fixed), col_idxs=anc, rows=R, cols=R[:, anc_t], approx_preference="rows",
                            precision=precision)
            v, i = cur.topk_in_row(test[:, anc_t], max_kr)
        for k_r, res in _grid_eval(test, v, i, top_k_vals, top_k_retvr_vals, n_ents).items():
            for k, metrics in res.items():
                out.setdefault(f"top_k={k}", {}).setdefault(f"k_retvr={k_r}", {})[f"anc_n_m={n_train}_anc_n_e={k_i}"] = metrics
    return out


def run_approx_eval(approx_method, all_ment_to_ent_scores, precomp_approx_ment_to_ent_scores, n_ment_anchors, n_ent_anchors,
                    top_k, top_k_retvr, n_seeds, *, precision="f32r"):
    """eval/run_retrieval_eval_wrt_exact_crossenc.py:162-201: seeds 0..n_seeds-1, metric-wise mean."""
    acc = {}
    for seed in range(n_seeds):
        res = run_approx_eval_w_seed(approx_method, all_ment_to_ent_scores, n_ment_anchors, n_ent_anchors, top_k, top_k_retvr,
                                     seed, precomp_approx_ment_to_ent_scores, precision=precision)
        for ment_type, res_dict in res.items():
            for metric, val in res_dict.items():
                acc.setdefault(ment_type, {}).setdefault(metric, []).append(float(val))
    return {mt: {m: float(np.mean(v)) for m, v in d.items()} for mt, d in acc.items()}


def sweep_grids(total_n_ment, total_n_ent):
    """The anchor grids of the random-anchor sweep (eval/run_retrieval_eval_wrt_exact_crossenc.py:227-233)."""
    n_ment_anchors_vals = [v for v in [50, 100, 200, 500, 1000, 2000, 5000] if v <= total_n_ment]
    n_ent_anchors_vals = [v for v in [50, 100, 200, 500, 1000, 2000] if v < total_n_ent] + [total_n_ent]
    return n_ment_anchors_vals, n_ent_anchors_vals


def run_sweep(all_ment_to_ent_scores, n_seeds=1, eval_methods=("cur", "cur_oracle"), top_k_vals=(10,), top_k_retr_vals=(500,),
              n_ment_anchors_vals=None, n_ent_anchors_vals=None, *, precision="f32r", progress=None):
    """The grid loop of the sweep driver's ``run`` (eval/run_retrieval_eval_wrt_exact_crossenc.py:204-371) for the CUR
    methods: eval_res[method]["top_k=k"]["k_retvr=k_r"]["anc_n_m=k_q~anc_n_e=k_i"] = seed-averaged metrics incl.
    ``approx_error`` / ``approx_error_relative`` for anchor / non_anchor / all rows, + ``other_args`` (:373-376 layout)."""
    A = engine._f32(all_ment_to_ent_scores)
    total_n_ment, total_n_ent = A.shape
    g_m, g_e = sweep_grids(total_n_ment, total_n_ent)
    n_ment_anchors_vals = list(n_ment_anchors_vals) if n_ment_anchors_vals is not None else g_m
    n_ent_anchors_vals = list(n_ent_anchors_vals) if n_ent_anchors_vals is not None else g_e
    eval_res = {}
    for method in eval_methods:
        if method not in ("cur", "cur_oracle"):
            raise NotImplementedError(f"Method = {method} not supported")
        for top_k in top_k_vals:
            for k_r in top_k_retr_vals:
                if k_r < top_k or k_r > total_n_ent:                                   # :347-348
                    continue
                for k_q in n_ment_anchors_vals:
                    for k_i in n_ent_anchors_vals:
                        ans = run_approx_eval(method, A, None, k_q, k_i, int(top_k), int(k_r), n_seeds, precision=precision)
                        eval_res.setdefault(method, {}).setdefault(f"top_k={top_k}", {}).setdefault(f"k_retvr={k_r}", {})[
                            f"anc_n_m={k_q}~anc_n_e={k_i}"] = ans
                        if progress is not None:
                            progress(method, top_k, k_r, k_q, k_i, ans)
    eval_res["other_args"] = {"top_k_vals": list(top_k_vals), "top_k_retr_vals": list(top_k_retr_vals),
                              "n_ent_anchors_vals": n_ent_anchors_vals, "n_ment_anchors_vals": n_ment_anchors_vals}
    return eval_res
