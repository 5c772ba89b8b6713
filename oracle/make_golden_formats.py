"""TEST INFRASTRUCTURE ONLY.  Golden fixtures for the on-disk formats either side of the search path, produced by
running the REFERENCE'S OWN scripts in the build container (needs /root/reference):

    python oracle/make_golden_formats.py

* utils/split_zeshel_ment2ent_for_cur_exps.py::run        -> the split index lists it writes (its RNG call order)
* ..._w_fixed_train_test_splits.py::run_eval_method("cur") -> the retrieval / anchor grids (:238-251) and the whole
  result dictionary for a tiny train/test pair, written through the reference's own pickle schema.
Output: tests/golden/formats.json (+ the tiny input matrices in tests/golden/formats_inputs.npz).
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


def main():
    ref = load_reference()
    split_mod = importlib.import_module("utils.split_zeshel_ment2ent_for_cur_exps")
    golden = {}
    # ---- splits ---------------------------------------------------------------------------------------------
    n_ments, n_ents = 64, 1100
    A = torch.from_numpy(synthetic_scores(n_ments, n_ents, rank=8, noise=0.05, seed=3))
    dump = {"ment_to_ent_scores": A, "ment_to_ent_scores.shape": A.shape,
            "test_data": [{"mention_id": f"m{i}"} for i in range(n_ments)],
            "mention_tokens_list": [[101, i, 102] for i in range(n_ments)],
            "entity_id_list": [], "entity_tokens_list": [], "arg_dict": {"data_name": "yugioh"}}
    with tempfile.TemporaryDirectory() as tmp:
        m2e_file = os.path.join(tmp, "m2e.pkl")
        with open(m2e_file, "wb") as f:
            pickle.dump(dump, f)
        split_mod.run(data_name="yugioh", m2e_file=m2e_file, num_train_ment_vals=[20, 30, 100], num_splits=2, seed=7,
                      dev_frac=0.1, base_out_dir=os.path.join(tmp, "m2e_splits"))
        files = {}
        for root, _, names in os.walk(os.path.join(tmp, "m2e_splits")):
            for nm in names:
                with open(os.path.join(root, nm), "rb") as f:
                    d = pickle.load(f)
                rel = os.path.relpath(os.path.join(root, nm), os.path.join(tmp, "m2e_splits"))
                files[rel] = {"ment_idxs": [int(i) for i in d["ment_idxs"]], "keys": sorted(d.keys()),
                              "shape": list(d["ment_to_ent_scores"].shape)}
        golden["splits"] = {"args": {"num_train_ment_vals": [20, 30, 100], "num_splits": 2, "seed": 7, "dev_frac": 0.1},
                            "files": files}
        # ---- the 'cur' method of the fixed-split eval on one of those splits ------------------------------------
        train_f = os.path.join(tmp, "m2e_splits", "nm_train=30", "split_idx=0", "train.pkl")
        test_f = os.path.join(tmp, "m2e_splits", "nm_train=30", "split_idx=0", "test.pkl")
        res, params = ref.modules.split.run_eval_method(
            curr_method="cur", test_data_file=test_f, train_data_file=train_f, fixed_anc_ent_args={}, bienc_args={},
            cur_args={"seed": 0}, tfidf_args={}, use_wandb=False)
        golden["cur_eval"] = {"seed": 0, "train": "nm_train=30/split_idx=0/train.pkl", "test": "nm_train=30/split_idx=0/test.pkl",
                              "retrieval_params": json.loads(json.dumps(params)),
                              "eval_res": json.loads(json.dumps(res))}
        # keep the fixture small: (common_frac mean, std) for every grid point + three full metric dicts
        full = golden["cur_eval"].pop("eval_res")
        slim, sample = {}, {}
        for tk, v in full.items():
            for kr, v2 in v.items():
                for an, m in v2.items():
                    slim.setdefault(tk, {}).setdefault(kr, {})[an] = [m["exact_vs_reranked_approx_retvr~common_frac_mean"],
                                                                     m["exact_vs_reranked_approx_retvr~common_frac_std"]]
                    if len(sample) < 3 and kr in ("k_retvr=100", "k_retvr=500"):
                        sample[f"{tk}|{kr}|{an}"] = m
        golden["cur_eval"]["eval_res_common_frac_mean_std"] = slim
        golden["cur_eval"]["eval_res_samples"] = sample
    np.savez_compressed(os.path.join(OUT, "formats_inputs.npz"), A=A.numpy())
    with open(os.path.join(OUT, "formats.json"), "w") as f:
        json.dump(golden, f, separators=(",", ":"))
    print("wrote formats.json:", len(golden["splits"]["files"]), "split files")


if __name__ == "__main__":
    main()
