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
