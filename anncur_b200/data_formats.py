"""On-disk formats either side of the search path, host-side only:

* the cross-encoder score-matrix pickle            eval/run_cross_encoder_for_ment_ent_matrix_zeshel.py:230-240
* its train / train_train / train_dev / test splits  utils/split_zeshel_ment2ent_for_cur_exps.py:25-129 (adds ``ment_idxs``)
* the entity-to-fixed-anchor dump (``ent_to_ent_scores`` / ``topk_ents``)  ..._w_fixed_train_test_splits.py:313-319
* the retrieval / anchor grids of the fixed-split eval  eval/run_retrieval_eval_wrt_exact_crossenc_w_fixed_train_test_splits.py:238-251
* the result JSON                                   ..._w_fixed_train_test_splits.py:429, :494-500

so that real ZeShEL matrices produced by the reference feed this engine and the reference's compile / plot
scripts read what it writes.  Nothing here computes scores; ``run_cur_method`` hands the matrices to
``eval_retrieval.fixed_split_cur_eval`` (GPU)."""
import json
import os
import pickle
from pathlib import Path

import numpy as np
import torch

M2E_KEYS = ("ment_to_ent_scores", "test_data", "mention_tokens_list", "entity_id_list", "entity_tokens_list", "arg_dict")
SPLIT_NAMES = ("train_dev", "train_train", "train", "test")          # order in which the reference writes them


def make_m2e_dict(ment_to_ent_scores, test_data, mention_tokens_list, entity_id_list=(), entity_tokens_list=(),
                  arg_dict=None, ment_idxs=None):
    """The dict the reference pickles (:230-240); ``ment_idxs`` only in split files (split script :35-43)."""
    scores = torch.as_tensor(ment_to_ent_scores)
    d = {"ment_to_ent_scores": scores, "ment_to_ent_scores.shape": scores.shape, "test_data": list(test_data),
         "mention_tokens_list": list(mention_tokens_list), "entity_id_list": list(entity_id_list),
         "entity_tokens_list": list(entity_tokens_list), "arg_dict": dict(arg_dict or {})}
    if ment_idxs is not None:
        d.pop("ment_to_ent_scores.shape")
        d["ment_idxs"] = list(ment_idxs)
    return d


def save_m2e_pickle(path, m2e_dict):
    Path(os.path.dirname(os.path.abspath(path))).mkdir(exist_ok=True, parents=True)
    with open(path, "wb") as fout:
        pickle.dump(m2e_dict, fout)


def load_m2e_pickle(path, require_ment_idxs=False):
    """Read a score-matrix pickle written by the reference (or by ``save_m2e_pickle``).  The matrix comes back
    as a CPU fp32 torch tensor whatever it was stored as; the other fields are passed through."""
    with open(path, "rb") as fin:
        d = pickle.load(fin)
    missing = [k for k in M2E_KEYS if k not in d]
    if missing:
        raise KeyError(f"{path}: not a ment_to_ent score dump, missing keys {missing}")
    if require_ment_idxs and "ment_idxs" not in d:
        raise KeyError(f"{path}: split files carry 'ment_idxs' (utils/split_zeshel_ment2ent_for_cur_exps.py:39)")
    scores = d["ment_to_ent_scores"]
    scores = scores if torch.is_tensor(scores) else torch.as_tensor(np.asarray(scores))
    d["ment_to_ent_scores"] = scores.detach().to("cpu", torch.float32)
    assert d["ment_to_ent_scores"].dim() == 2
    return d


def split_indices(n_ments, num_train_ment_vals, num_splits, seed, dev_frac):
    """The index sets of utils/split_zeshel_ment2ent_for_cur_exps.py:80-92 with its exact RNG call order (ONE
    generator; per (num_train, split): train draw, then the train-dev draw).  Yields
    (num_train_ments, split_iter, {"train", "test", "train_dev", "train_train"} -> sorted index lists)."""
    assert 0 <= dev_frac < 1
    rng = np.random.default_rng(seed=seed)
    for num_train_ments in num_train_ment_vals:
        for split_iter in range(num_splits):
            if num_train_ments > n_ments:
                continue
            train = sorted(rng.choice(n_ments, size=num_train_ments, replace=False))
            test = sorted(list(set(range(n_ments)) - set(train)))
            train_dev = sorted(rng.choice(a=train, size=int(num_train_ments * dev_frac), replace=False))
            train_train = sorted(list(set(train) - set(train_dev)))
            yield num_train_ments, split_iter, {"train": [int(i) for i in train], "test": [int(i) for i in test],
                                                "train_dev": [int(i) for i in train_dev],
                                                "train_train": [int(i) for i in train_train]}


def write_splits(m2e_dict, num_train_ment_vals, num_splits, seed, dev_frac, base_out_dir):
    """Write the reference's directory layout ``{base}/nm_train={n}/split_idx={i}/{split}.pkl`` (:92-128); empty
    index lists are skipped like the reference does (:28-30).  Returns the files written."""
    scores = torch.as_tensor(m2e_dict["ment_to_ent_scores"])
    data, toks = m2e_dict["test_data"], m2e_dict["mention_tokens_list"]
    assert scores.shape[0] == len(data) == len(toks)
    written = []
    for n_train, split_iter, idx in split_indices(scores.shape[0], num_train_ment_vals, num_splits, seed, dev_frac):
        out_dir = f"{base_out_dir}/nm_train={n_train}/split_idx={split_iter}"
        for name in SPLIT_NAMES:
            ment_idxs = idx[name]
            if len(ment_idxs) == 0:
                continue
            d = make_m2e_dict(scores[ment_idxs, :], [data[i] for i in ment_idxs], [toks[i] for i in ment_idxs],
                              arg_dict=m2e_dict["arg_dict"], ment_idxs=ment_idxs)
            save_m2e_pickle(f"{out_dir}/{name}.pkl", d)
            written.append(f"{out_dir}/{name}.pkl")
    return written


def retrieval_grids(n_ent, method="cur"):
    """top_k / k_retvr / n_ent_anchors grids exactly as ..._w_fixed_train_test_splits.py:238-251 builds them
    (both grids contain 0 = int(1 * 0.1), and n_ent itself)."""
    top_k_vals = [1, 10, 50, 100]
    base = [1, 10, 50, 100, 200, 500, 1000]
    cur = base + [int(k * frac) for k in base for frac in np.arange(0.1, 1.0, 0.1)]
    top_k_retr_vals = cur if ("cur" in method or "fixed_anc_ent" in method) else base
    top_k_retr_vals = sorted(list(set(top_k_retr_vals)))
    anchors_base = [10, 50, 100, 200, 500, 1000, 2000]
    n_ent_anchors_vals = [v for v in anchors_base if v < n_ent] + [n_ent]
    n_ent_anchors_vals = sorted(list(set(n_ent_anchors_vals + cur)))
    return top_k_vals, top_k_retr_vals, n_ent_anchors_vals


E2E_KEYS = ("ent_to_ent_scores", "topk_ents")


def make_e2e_dict(ent_to_ent_scores, topk_ents):
    """The entity-to-fixed-anchor-entity dump the ``fixed_anc_ent`` / ``fixed_anc_ent_cur`` methods read
    (..._w_fixed_train_test_splits.py:313-319, :335-340; eval/run_retrieval_eval_wrt_exact_crossenc.py:299-303): its producer is
    not in the reference repo, the consumers fix the schema -- ``ent_to_ent_scores`` (n_ents x n_fixed_anchors, column j =
    every entity scored against fixed anchor j) and ``topk_ents`` whose ROW 0 lists the n_fixed_anchors anchor entity ids."""
    scores = torch.as_tensor(ent_to_ent_scores)
    topk = np.asarray(topk_ents)
    topk = topk[None, :] if topk.ndim == 1 else topk
    assert scores.dim() == 2 and topk.shape[1] == scores.shape[1], (scores.shape, topk.shape)
    return {"ent_to_ent_scores": scores, "topk_ents": topk}


def save_e2e_pickle(path, e2e_dict):
    Path(os.path.dirname(os.path.abspath(path))).mkdir(exist_ok=True, parents=True)
    with open(path, "wb") as fout:
        pickle.dump(e2e_dict, fout)


def load_e2e_pickle(path, n_fixed_anc_ent=None):
    """-> (ent_embeds [n_ents x n] CPU fp32 tensor, anchor_ent_idxs [n] int64 array): the first ``n_fixed_anc_ent`` fixed
    anchors as the reference slices them (:321-322, :340; all of them when None)."""
    with open(path, "rb") as fin:
        d = pickle.load(fin)
    missing = [k for k in E2E_KEYS if k not in d]
    if missing:
        raise KeyError(f"{path}: not an ent_to_ent score dump, missing keys {missing}")
    scores = d["ent_to_ent_scores"]
    scores = scores if torch.is_tensor(scores) else torch.as_tensor(np.asarray(scores))
    scores = scores.detach().to("cpu", torch.float32)
    anchors = d["topk_ents"][0]
    anchors = anchors.cpu().numpy() if torch.is_tensor(anchors) else np.asarray(anchors)
    anchors = anchors.astype(np.int64)
    assert scores.dim() == 2 and anchors.shape[0] == scores.shape[1], (scores.shape, anchors.shape)
    n = scores.shape[1] if n_fixed_anc_ent is None else int(n_fixed_anc_ent)
    return scores[:, :n].contiguous(), anchors[:n]


def load_embeddings(path):
    """Precomputed embeddings for the ``bienc`` / ``tfidf`` methods (.npy, .npz with one array, or a torch-saved tensor): the
    reference computes them with BERT / TF-IDF models at this point (:259-282, :366-381), which is outside this engine."""
    if path.endswith(".npy"):
        arr = np.load(path)
    elif path.endswith(".npz"):
        z = np.load(path)
        arr = z[z.files[0]]
    else:
        arr = torch.load(path, map_location="cpu")
        arr = arr.numpy() if torch.is_tensor(arr) else np.asarray(arr)
    return torch.as_tensor(np.asarray(arr, dtype=np.float32))


def run_eval_method(curr_method, test_data_file, train_data_file, bienc_args=None, cur_args=None, fixed_anc_ent_args=None,
                    tfidf_args=None, use_wandb=False, *, precision="f32r", n_ent_anchors_vals=None, top_k_retr_vals=None):
    """``run_eval_method`` of the reference (..._w_fixed_train_test_splits.py:209-443), same arguments and return value
    ``(eval_res, retrieval_params)``, every method on the fused score + top-k kernel:

    * ``cur``               :286-303  index E = pinv(C) . R per anchor count, queries = test scores of the anchors
    * ``fixed_anc_ent``     :305-325  items = rows of the e2e dump (N x n_fixed), queries = test scores of the fixed anchors
    * ``fixed_anc_ent_cur`` :327-358  CUR with the e2e dump as anchor rows
    * ``bienc`` / ``tfidf`` :257-284 / :360-385  items / queries = PRECOMPUTED embeddings: ``bienc_args`` / ``tfidf_args`` carry
      ``ent_embed_file`` and ``ment_embed_file`` (the models that produce them are out of scope); tfidf mention embeddings are
      indexed by the test file's ``ment_idxs`` as the reference does (:378)."""
    from . import eval_retrieval as ER
    del use_wandb
    test = load_m2e_pickle(test_data_file, require_ment_idxs=True)
    train = load_m2e_pickle(train_data_file)
    A_test, A_train = test["ment_to_ent_scores"], train["ment_to_ent_scores"]
    assert A_train.shape[1] == A_test.shape[1], "Train and test entities differ! Use entity_id_list from data dump to resolve this"
    n_train, n_ent = A_train.shape[0], A_test.shape[1]
    top_k_vals, kr_all, ki_all = retrieval_grids(n_ent, curr_method)
    kr = [v for v in kr_all if top_k_retr_vals is None or v in set(top_k_retr_vals)]
    keep_ki = None if n_ent_anchors_vals is None else set(n_ent_anchors_vals)
    params = {"top_k_retr_vals": kr_all, "top_k_vals": top_k_vals, "n_ent_anchors_vals": ki_all}
    if curr_method == "cur":
        res = ER.fixed_split_cur_eval(A_train, A_test, ki_all, top_k_vals, kr, (cur_args or {}).get("seed", 0),
                                      precision=precision, only_k_i=keep_ki)
    elif curr_method == "fixed_anc_ent":
        a = fixed_anc_ent_args or {}
        ent_embeds, anchor_ent_idxs = load_e2e_pickle(a["e2e_fname"], a["n_fixed_anc_ent"])
        mention_embeds = A_test[:, torch.as_tensor(anchor_ent_idxs)]                                  # :323
        res = ER.fixed_split_embed_eval(A_test, mention_embeds, ent_embeds, n_train, ki_all, top_k_vals, kr, precision=precision)
    elif curr_method == "fixed_anc_ent_cur":
        a = fixed_anc_ent_args or {}
        ent_embeds, _ = load_e2e_pickle(a["e2e_fname"], a["n_fixed_anc_ent"])
        res = ER.fixed_split_fixed_anc_cur_eval(A_test, ent_embeds, n_train, ki_all, top_k_vals, kr, seed=0,
                                                precision=precision, only_k_i=keep_ki)
    elif curr_method in ("bienc", "tfidf"):
        a = (bienc_args if curr_method == "bienc" else tfidf_args) or {}
        if not a.get("ent_embed_file") or not a.get("ment_embed_file"):
            raise ValueError(f"eval_method={curr_method}: pass precomputed embeddings (ent_embed_file, ment_embed_file); the "
                             "BERT / TF-IDF models that produce them are outside this engine")
        label_embeds = load_embeddings(a["ent_embed_file"])
        mention_embeds = load_embeddings(a["ment_embed_file"])
        if curr_method == "tfidf" or mention_embeds.shape[0] != A_test.shape[0]:
            mention_embeds = mention_embeds[torch.as_tensor(np.asarray(test["ment_idxs"], dtype=np.int64))]   # :378
        res = ER.fixed_split_embed_eval(A_test, mention_embeds, label_embeds, n_train, ki_all, top_k_vals, kr, precision=precision)
    else:
        raise NotImplementedError(f"Method = {curr_method} not supported")
    return res, params


def run_cur_method(test_data_file, train_data_file, seed, *, n_ent_anchors_vals=None, top_k_retr_vals=None,
                   precision="f32r"):
    """``run_eval_method(curr_method="cur", ...)`` (:209-443) on the GPU: load the two split pickles, replay the
    anchor draws, build one index per k_i, retrieve once at the largest k_r and evaluate every (k, k_r).
    Returns (eval_res, retrieval_params) in the reference's layout.  The optional grids restrict the sweep (the
    anchor draws of skipped k_i values are still replayed so that the remaining ones see the reference's anchors)."""
    from .eval_retrieval import fixed_split_cur_eval
    test = load_m2e_pickle(test_data_file, require_ment_idxs=True)
    train = load_m2e_pickle(train_data_file)
    A_test, A_train = test["ment_to_ent_scores"], train["ment_to_ent_scores"]
    assert A_train.shape[1] == A_test.shape[1], "Train and test entities differ! Use entity_id_list from data dump to resolve this"
    n_ent = A_test.shape[1]
    top_k_vals, kr_all, ki_all = retrieval_grids(n_ent, "cur")
    kr = [v for v in kr_all if top_k_retr_vals is None or v in set(top_k_retr_vals)]
    keep_ki = None if n_ent_anchors_vals is None else set(n_ent_anchors_vals)
    res = fixed_split_cur_eval(A_train, A_test, ki_all, top_k_vals, kr, seed, precision=precision, only_k_i=keep_ki)
    params = {"top_k_retr_vals": kr_all, "top_k_vals": top_k_vals, "n_ent_anchors_vals": ki_all}
    return res, params


def write_result_json(res_dir, eval_method, misc, eval_res_by_seed, arg_dict, retriever_params):
    """``{res_dir}/method={eval_method}_{misc}.json`` with the reference's top-level layout (:494-500):
    ``seed=<s>`` -> results, ``other_args`` -> CLI arguments + ``retriever_params``."""
    out = {f"seed={s}": r for s, r in eval_res_by_seed.items()}
    out["other_args"] = dict(arg_dict)
    out["other_args"]["retriever_params"] = retriever_params
    res_file = f"{res_dir}/method={eval_method}_{misc}.json"
    Path(os.path.dirname(res_file)).mkdir(exist_ok=True, parents=True)
    with open(res_file, "w") as fout:
        json.dump(out, fout, indent=4)
    return res_file
