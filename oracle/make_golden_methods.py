"""TEST INFRASTRUCTURE ONLY.  Golden fixtures for the other approximators of the fixed-split eval, produced
by running the REFERENCE'S OWN ``run_eval_method`` in the build container (needs /root/reference):

    python oracle/make_golden_methods.py

* ``fixed_anc_ent``      ..._w_fixed_train_test_splits.py:305-325
* ``fixed_anc_ent_cur``  ..._w_fixed_train_test_splits.py:327-358
on the tiny train / test split of tests/golden/formats_inputs.npz (the reference's own split script, seed 7) and a
synthetic entity-to-fixed-anchor dump written in the schema the reference reads (:313-319).  ``bienc`` / ``tfidf`` need BERT /
TF-IDF models and are not run; after their embedding step they execute the same lines as ``fixed_anc_ent`` (:283, :383 ==
:324, then :403-429).  Output: tests/golden/methods.json + tests/golden/methods_inputs.npz.
"""
import importlib
import json
import os
import pickle
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.cur_oracle import synthetic_scores  # noqa: E402
from oracle.ref_shim import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
N_FIXED_IN_FILE, N_FIXED_USED = 40, 25


def slim(full):
    out = {}
    for tk, v in full.items():
        for kr, v2 in v.items():
            for an, m in v2.items():
                out.setdefault(tk, {}).setdefault(kr, {})[an] = [m["exact_vs_reranked_approx_retvr~common_frac_mean"],
                                                                m["exact_vs_reranked_approx_retvr~common_frac_std"]]
    return out


def main():
    ref = load_reference()
    split_mod = importlib.import_module("utils.split_zeshel_ment2ent_for_cur_exps")
    A = torch.from_numpy(np.load(os.path.join(OUT, "formats_inputs.npz"))["A"])
    n_ments, n_ents = A.shape
    rng = np.random.default_rng(11)
    e2e = synthetic_scores(n_ents, N_FIXED_IN_FILE, rank=8, noise=0.05, seed=11)
    topk_ents = rng.choice(n_ents, size=(1, N_FIXED_IN_FILE), replace=False).astype(np.int64)
    golden = {"n_fixed_anc_ent": N_FIXED_USED, "train": "nm_train=30/split_idx=0/train.pkl", "test": "nm_train=30/split_idx=0/test.pkl",
              "split_args": {"num_train_ment_vals": [20, 30, 100], "num_splits": 2, "seed": 7, "dev_frac": 0.1}}
    with tempfile.TemporaryDirectory() as tmp:
        dump = {"ment_to_ent_scores": A, "ment_to_ent_scores.shape": A.shape,
                "test_data": [{"mention_id": f"m{i}"} for i in range(n_ments)],
                "mention_tokens_list": [[101, i, 102] for i in range(n_ments)],
                "entity_id_list": [], "entity_tokens_list": [], "arg_dict": {"data_name": "yugioh"}}
        m2e_file = os.path.join(tmp, "m2e.pkl")
        with open(m2e_file, "wb") as f:
            pickle.dump(dump, f)
        split_mod.run(data_name="yugioh", m2e_file=m2e_file, num_train_ment_vals=[20, 30, 100], num_splits=2, seed=7,
                      dev_frac=0.1, base_out_dir=os.path.join(tmp, "m2e_splits"))
        train_f = os.path.join(tmp, "m2e_splits", "nm_train=30", "split_idx=0", "train.pkl")
        test_f = os.path.join(tmp, "m2e_splits", "nm_train=30", "split_idx=0", "test.pkl")
        e2e_f = os.path.join(tmp, "e2e.pkl")
        with open(e2e_f, "wb") as f:                                   # the schema the reference reads (:313-319)
            pickle.dump({"ent_to_ent_scores": torch.from_numpy(e2e), "topk_ents": topk_ents}, f)
        for method in ("fixed_anc_ent", "fixed_anc_ent_cur"):
            res, params = ref.modules.split.run_eval_method(
                curr_method=method, test_data_file=test_f, train_data_file=train_f, bienc_args={}, cur_args={"seed": 0},
                fixed_anc_ent_args={"e2e_fname": e2e_f, "n_fixed_anc_ent": N_FIXED_USED}, tfidf_args={}, use_wandb=False)
            full = json.loads(json.dumps(res))
            golden[method] = {"retrieval_params": json.loads(json.dumps(params)), "eval_res_common_frac_mean_std": slim(full)}
            print(method, "grid points:", sum(len(v2) for v in full.values() for v2 in v.values()))
    np.savez_compressed(os.path.join(OUT, "methods_inputs.npz"), ent_to_ent_scores=e2e, topk_ents=topk_ents)
    with open(os.path.join(OUT, "methods.json"), "w") as f:
        json.dump(golden, f, separators=(",", ":"))
    print("wrote methods.json")


if __name__ == "__main__":
    main()
