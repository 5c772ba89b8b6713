"""CPU: on-disk formats either side of the search path (anncur_b200/data_formats.py) against fixtures produced by the
reference's own split script and eval driver (oracle/make_golden_formats.py -> tests/golden/formats.json)."""
import json
import os
import pickle

import numpy as np
import pytest
import torch

from anncur_b200 import data_formats as F

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return json.load(open(os.path.join(GOLD, "formats.json")))


@pytest.fixture(scope="module")
def dump():
    A = torch.from_numpy(np.load(os.path.join(GOLD, "formats_inputs.npz"))["A"])
    n = A.shape[0]
    return F.make_m2e_dict(A, [{"mention_id": f"m{i}"} for i in range(n)], [[101, i, 102] for i in range(n)],
                           arg_dict={"data_name": "yugioh"})


def test_m2e_pickle_roundtrip_and_schema(tmp_path, dump):
    p = tmp_path / "sub" / "m2e.pkl"
    F.save_m2e_pickle(str(p), dump)
    raw = pickle.load(open(p, "rb"))
    assert set(F.M2E_KEYS) <= set(raw) and "ment_to_ent_scores.shape" in raw         # reference schema (:230-240)
    back = F.load_m2e_pickle(str(p))
    assert torch.equal(back["ment_to_ent_scores"], dump["ment_to_ent_scores"]) and back["ment_to_ent_scores"].dtype == torch.float32
    with pytest.raises(KeyError):
        F.load_m2e_pickle(str(p), require_ment_idxs=True)
    bad = tmp_path / "bad.pkl"
    pickle.dump({"ment_to_ent_scores": np.zeros((2, 2))}, open(bad, "wb"))
    with pytest.raises(KeyError):
        F.load_m2e_pickle(str(bad))
    # numpy / float64 matrices are accepted and come back as fp32 torch
    alt = dict(dump, ment_to_ent_scores=dump["ment_to_ent_scores"].double().numpy())
    F.save_m2e_pickle(str(tmp_path / "alt.pkl"), alt)
    assert F.load_m2e_pickle(str(tmp_path / "alt.pkl"))["ment_to_ent_scores"].dtype == torch.float32


def test_splits_match_the_reference_script(tmp_path, gold, dump):
    a = gold["splits"]["args"]
    written = F.write_splits(dump, a["num_train_ment_vals"], a["num_splits"], a["seed"], a["dev_frac"], str(tmp_path))
    rel = {os.path.relpath(w, str(tmp_path)): w for w in written}
    assert sorted(rel) == sorted(gold["splits"]["files"])          # nm_train=100 > n_ments is skipped like the reference
    for name, want in gold["splits"]["files"].items():
        d = pickle.load(open(rel[name], "rb"))
        assert [int(i) for i in d["ment_idxs"]] == want["ment_idxs"]
        assert sorted(d.keys()) == want["keys"] and list(d["ment_to_ent_scores"].shape) == want["shape"]
        assert torch.equal(d["ment_to_ent_scores"], dump["ment_to_ent_scores"][want["ment_idxs"], :])
        assert d["mention_tokens_list"] == [dump["mention_tokens_list"][i] for i in want["ment_idxs"]]


def test_retrieval_grids_match_the_reference_driver(gold):
    p = gold["cur_eval"]["retrieval_params"]
    n_ent = 1100
    top_k, k_r, k_i = F.retrieval_grids(n_ent, "cur")
    assert top_k == p["top_k_vals"] and k_r == p["top_k_retr_vals"] and k_i == p["n_ent_anchors_vals"]
    assert 0 in k_r and 0 in k_i and n_ent in k_i and len(k_r) == 41
    assert F.retrieval_grids(n_ent, "bienc")[1] == [1, 10, 50, 100, 200, 500, 1000]


def test_result_json_layout(tmp_path):
    f = F.write_result_json(str(tmp_path / "r"), "cur", "x", {0: {"top_k=1": {}}, 1: {"top_k=1": {}}}, {"data_name": "yugioh"},
                            {"top_k_vals": [1]})
    d = json.load(open(f))
    assert os.path.basename(f) == "method=cur_x.json"
    assert sorted(d) == ["other_args", "seed=0", "seed=1"] and d["other_args"]["retriever_params"] == {"top_k_vals": [1]}


@pytest.mark.gpu
def test_cur_method_on_split_files_matches_reference_results(tmp_path, gold, dump):
    """run_eval_method('cur') of the reference on its own split files vs run_cur_method on ours (GPU)."""
    a = gold["splits"]["args"]
    F.write_splits(dump, a["num_train_ment_vals"], a["num_splits"], a["seed"], a["dev_frac"], str(tmp_path))
    c = gold["cur_eval"]
    k_i_subset = [10, 50, 100, 200, 500, 1000, 1100, 45, 700]
    res, params = F.run_cur_method(str(tmp_path / c["test"]), str(tmp_path / c["train"]), c["seed"], n_ent_anchors_vals=k_i_subset)
    assert params == c["retrieval_params"]
    want = c["eval_res_common_frac_mean_std"]
    n_cmp, n_bad, worst = 0, 0, 0.0
    for tk, v in want.items():
        for kr, v2 in v.items():
            for an, (mean, std) in v2.items():
                k_i = int(an.split("anc_n_e=")[1])
                if k_i not in k_i_subset:
                    continue
                got = res[tk][kr][an]["exact_vs_reranked_approx_retvr~common_frac_mean"]
                n_cmp += 1
                worst = max(worst, abs(got - mean))
                n_bad += abs(got - mean) > 1e-3
    assert n_cmp > 800
    # 34 test queries: one swapped near-tie moves a mean by >= 1/(34 k); allow a few, none large
    assert n_bad <= 0.02 * n_cmp and worst <= 0.05, (n_bad, n_cmp, worst)
    for key, m in c["eval_res_samples"].items():
        tk, kr, an = key.split("|")
        if int(an.split("anc_n_e=")[1]) in k_i_subset:
            assert sorted(res[tk][kr][an]) == sorted(m)             # same metric keys as the reference writes


def test_cli_flags_mirror_the_reference_script():
    from anncur_b200.run_fixed_split_eval import build_parser
    flags = {a.dest for a in build_parser()._actions}
    reference_flags = {"data_name", "eval_method", "res_dir", "test_data_file", "train_data_file", "n_seeds", "bi_model_file",
                       "batch_size", "e2e_fname", "n_fixed_anc_ent", "mention_file", "entity_file", "mode", "misc", "use_wandb"}
    assert reference_flags <= flags                                   # ..._w_fixed_train_test_splits.py:515-542
    with pytest.raises(SystemExit):
        from anncur_b200.run_fixed_split_eval import main
        main(["--res_dir", "x", "--test_data_file", "t.pkl", "--eval_method", "bienc"])


@pytest.mark.gpu
def test_cli_end_to_end_writes_reference_layout(tmp_path, gold, dump):
    from anncur_b200.run_fixed_split_eval import main
    a = gold["splits"]["args"]
    F.write_splits(dump, a["num_train_ment_vals"], a["num_splits"], a["seed"], a["dev_frac"], str(tmp_path / "splits"))
    c = gold["cur_eval"]
    f = main(["--data_name", "yugioh", "--eval_method", "cur", "--res_dir", str(tmp_path / "res"), "--misc", "t",
              "--test_data_file", str(tmp_path / "splits" / c["test"]), "--train_data_file", str(tmp_path / "splits" / c["train"]),
              "--n_seeds", "2", "--k_i", "50", "200", "--k_r", "100", "500"])
    d = json.load(open(f))
    assert sorted(d) == ["other_args", "seed=0", "seed=1"]
    assert d["other_args"]["retriever_params"] == c["retrieval_params"] and d["other_args"]["eval_method"] == "cur"
    want = c["eval_res_common_frac_mean_std"]["top_k=10"]["k_retvr=100"]["anc_n_m=30_anc_n_e=50"][0]
    got = d["seed=0"]["top_k=10"]["k_retvr=100"]["anc_n_m=30_anc_n_e=50"]["exact_vs_reranked_approx_retvr~common_frac_mean"]
    assert abs(got - want) <= 0.03


# ---- the other approximators of the fixed-split eval and the e2e dump they read (8f-2) ----------------------
@pytest.fixture(scope="module")
def methods_gold():
    golden_dir = GOLD
    with open(os.path.join(golden_dir, "methods.json")) as f:
        g = json.load(f)
    z = np.load(os.path.join(golden_dir, "methods_inputs.npz"))
    g["ent_to_ent_scores"], g["topk_ents"] = z["ent_to_ent_scores"], z["topk_ents"]
    return g


def test_e2e_pickle_roundtrip_and_schema(tmp_path, methods_gold):
    g = methods_gold
    path = str(tmp_path / "e2e" / "ent_to_ent.pkl")
    F.save_e2e_pickle(path, F.make_e2e_dict(g["ent_to_ent_scores"], g["topk_ents"]))
    import pickle
    raw = pickle.load(open(path, "rb"))
    assert sorted(raw) == ["ent_to_ent_scores", "topk_ents"]               # what the reference reads (:313-319)
    assert torch.is_tensor(raw["ent_to_ent_scores"]) and np.asarray(raw["topk_ents"]).shape == (1, 40)
    emb, anc = F.load_e2e_pickle(path, 25)
    assert emb.shape == (1100, 25) and emb.dtype == torch.float32 and anc.shape == (25,) and anc.dtype == np.int64
    assert np.array_equal(emb.numpy(), g["ent_to_ent_scores"][:, :25]) and np.array_equal(anc, g["topk_ents"][0][:25])
    assert F.load_e2e_pickle(path)[0].shape == (1100, 40)
    # a 1-D anchor list is promoted to the (1 x n) layout; a score dump is refused
    assert F.make_e2e_dict(g["ent_to_ent_scores"], g["topk_ents"][0])["topk_ents"].shape == (1, 40)
    F.save_m2e_pickle(str(tmp_path / "m2e.pkl"), F.make_m2e_dict(np.zeros((2, 3), np.float32), [{}, {}], [[1], [2]]))
    with pytest.raises(KeyError):
        F.load_e2e_pickle(str(tmp_path / "m2e.pkl"))


def test_oracle_other_methods_match_the_reference_golden(methods_gold, gold, dump):
    """Pins the oracle's restatement of fixed_anc_ent / fixed_anc_ent_cur to the reference's run_eval_method output
    (oracle/make_golden_methods.py) on a sample of grid points -- CPU only."""
    from oracle import cur_oracle as O
    g = methods_gold
    a = g["split_args"]
    idx = [i for n, s, i in F.split_indices(64, a["num_train_ment_vals"], a["num_splits"], a["seed"], a["dev_frac"]) if n == 30 and s == 0][0]
    test = dump["ment_to_ent_scores"][idx["test"], :]
    n_fixed = g["n_fixed_anc_ent"]
    top_k_vals = [1, 10, 50, 100]
    approx = O.fixed_anc_ent_scores(test, g["ent_to_ent_scores"], g["topk_ents"][0], n_fixed)
    want = g["fixed_anc_ent"]["eval_res_common_frac_mean_std"]
    for k_r in (10, 100, 450, 1000):
        res = O.eval_approx_score_mat_for_all_topk(test, approx, top_k_vals, k_r)
        for k, m in res.items():
            for an in ("anc_n_m=30_anc_n_e=0", "anc_n_m=30_anc_n_e=500"):           # one evaluation entered under every key (:411-417)
                mean, std = want[f"top_k={k}"][f"k_retvr={k_r}"][an]
                assert abs(m["exact_vs_reranked_approx_retvr~common_frac_mean"] - mean) < 1e-6
                assert abs(m["exact_vs_reranked_approx_retvr~common_frac_std"] - std) < 1e-6
    ki_all = g["fixed_anc_ent_cur"]["retrieval_params"]["n_ent_anchors_vals"]
    approx_by_ki = O.fixed_anc_ent_cur_scores(test, g["ent_to_ent_scores"], n_fixed, ki_all, seed=0)
    want = g["fixed_anc_ent_cur"]["eval_res_common_frac_mean_std"]
    for k_i in (10, 25, 90, 500, 1100):
        for k_r in (50, 500):
            res = O.eval_approx_score_mat_for_all_topk(test, approx_by_ki[k_i], top_k_vals, k_r)
            for k, m in res.items():
                mean, std = want[f"top_k={k}"][f"k_retvr={k_r}"][f"anc_n_m=30_anc_n_e={k_i}"]
                assert abs(m["exact_vs_reranked_approx_retvr~common_frac_mean"] - mean) < 1e-6, (k_i, k_r, k)


def _compare_grid(res, want, skip_ki=(0,), tol_frac=0.02, only_ki=None):
    n_cmp, n_bad, worst = 0, 0, 0.0
    for tk, v in want.items():
        for kr, v2 in v.items():
            for an, (mean, std) in v2.items():
                k_i = int(an.split("anc_n_e=")[1])
                if k_i in skip_ki or (only_ki is not None and k_i not in only_ki):
                    continue
                got = res[tk][kr][an]["exact_vs_reranked_approx_retvr~common_frac_mean"]
                n_cmp += 1
                worst = max(worst, abs(got - mean))
                n_bad += abs(got - mean) > 1e-3
    return n_cmp, n_bad, worst


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["fixed_anc_ent", "fixed_anc_ent_cur"])
def test_other_methods_on_split_files_match_reference_results(tmp_path, gold, dump, methods_gold, method):
    """run_eval_method(<method>) of the reference on its own split files + e2e dump vs ours through the fused kernel (GPU)."""
    g = methods_gold
    a = gold["splits"]["args"]
    F.write_splits(dump, a["num_train_ment_vals"], a["num_splits"], a["seed"], a["dev_frac"], str(tmp_path))
    e2e = str(tmp_path / "e2e.pkl")
    F.save_e2e_pickle(e2e, F.make_e2e_dict(g["ent_to_ent_scores"], g["topk_ents"]))
    only = None if method == "fixed_anc_ent" else [10, 20, 25, 45, 100, 500, 700, 1000, 1100]
    res, params = F.run_eval_method(method, str(tmp_path / g["test"]), str(tmp_path / g["train"]),
                                    fixed_anc_ent_args={"e2e_fname": e2e, "n_fixed_anc_ent": g["n_fixed_anc_ent"]},
                                    cur_args={"seed": 0}, n_ent_anchors_vals=only)
    assert params == g[method]["retrieval_params"]
    # k_i = 0 under fixed_anc_ent_cur is an all-zero approximation (pure ties; torch.topk's order is unspecified) -> skipped there;
    # fixed_anc_ent enters ONE evaluation under every anchor count, 0 included
    n_cmp, n_bad, worst = _compare_grid(res, g[method]["eval_res_common_frac_mean_std"],
                                        skip_ki=() if method == "fixed_anc_ent" else (0,), only_ki=only)
    assert n_cmp > 900
    assert n_bad <= 0.02 * n_cmp and worst <= 0.05, (n_bad, n_cmp, worst)


@pytest.mark.gpu
def test_bienc_and_tfidf_served_from_precomputed_embeddings(tmp_path, gold, dump, methods_gold):
    """bienc / tfidf: after their (out-of-scope) embedding step the reference executes the fixed_anc_ent lines -- so with the
    fixed-anchor embeddings passed in as 'precomputed embeddings' both must reproduce the fixed_anc_ent golden numbers on
    the base k_r grid (the grid of those methods, :246-247), tfidf indexing the mention embeddings by ment_idxs (:378)."""
    g = methods_gold
    a = gold["splits"]["args"]
    F.write_splits(dump, a["num_train_ment_vals"], a["num_splits"], a["seed"], a["dev_frac"], str(tmp_path))
    n_fixed = g["n_fixed_anc_ent"]
    anchors = g["topk_ents"][0][:n_fixed]
    np.save(str(tmp_path / "ent.npy"), g["ent_to_ent_scores"][:, :n_fixed])
    all_ment = dump["ment_to_ent_scores"].numpy()[:, anchors]                 # embeddings of ALL 64 mentions
    np.save(str(tmp_path / "ment_all.npy"), all_ment)
    test_idx = F.load_m2e_pickle(str(tmp_path / g["test"]))["ment_idxs"]
    np.save(str(tmp_path / "ment_test.npy"), all_ment[test_idx])
    want = g["fixed_anc_ent"]["eval_res_common_frac_mean_std"]
    for method, ment_file in (("tfidf", "ment_all.npy"), ("bienc", "ment_test.npy"), ("bienc", "ment_all.npy")):
        args = {"ent_embed_file": str(tmp_path / "ent.npy"), "ment_embed_file": str(tmp_path / ment_file)}
        res, params = F.run_eval_method(method, str(tmp_path / g["test"]), str(tmp_path / g["train"]), bienc_args=args, tfidf_args=args)
        assert params["top_k_retr_vals"] == [1, 10, 50, 100, 200, 500, 1000]
        n_cmp = 0
        for tk, v in res.items():
            for kr, v2 in v.items():
                for an, m in v2.items():
                    assert abs(m["exact_vs_reranked_approx_retvr~common_frac_mean"] - want[tk][kr][an][0]) <= 1e-3, (method, tk, kr, an)
                    n_cmp += 1
        assert n_cmp > 500
    with pytest.raises(ValueError):
        F.run_eval_method("bienc", str(tmp_path / g["test"]), str(tmp_path / g["train"]), bienc_args={})


@pytest.mark.gpu
def test_cli_serves_fixed_anc_ent(tmp_path, gold, dump, methods_gold):
    from anncur_b200.run_fixed_split_eval import main
    g = methods_gold
    a = gold["splits"]["args"]
    F.write_splits(dump, a["num_train_ment_vals"], a["num_splits"], a["seed"], a["dev_frac"], str(tmp_path / "splits"))
    e2e = str(tmp_path / "e2e.pkl")
    F.save_e2e_pickle(e2e, F.make_e2e_dict(g["ent_to_ent_scores"], g["topk_ents"]))
    f = main(["--eval_method", "fixed_anc_ent", "--res_dir", str(tmp_path / "res"), "--misc", "t", "--e2e_fname", e2e,
              "--n_fixed_anc_ent", str(g["n_fixed_anc_ent"]), "--test_data_file", str(tmp_path / "splits" / g["test"]),
              "--train_data_file", str(tmp_path / "splits" / g["train"]), "--k_r", "100", "500"])
    d = json.load(open(f))
    assert os.path.basename(f) == "method=fixed_anc_ent_t.json" and d["other_args"]["eval_method"] == "fixed_anc_ent"
    want = g["fixed_anc_ent"]["eval_res_common_frac_mean_std"]["top_k=10"]["k_retvr=100"]["anc_n_m=30_anc_n_e=50"][0]
    assert abs(d["seed=0"]["top_k=10"]["k_retvr=100"]["anc_n_m=30_anc_n_e=50"]["exact_vs_reranked_approx_retvr~common_frac_mean"] - want) <= 1e-3
