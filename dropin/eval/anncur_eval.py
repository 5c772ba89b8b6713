"""Batched GPU versions of the per-query eval functions of the two retrieval-eval scripts, importable as
``eval.anncur_eval`` (the scripts themselves stay the reference's; INTEGRATION.md shows the two-line patch):

    run_approx_eval_w_seed                 eval/run_retrieval_eval_wrt_exact_crossenc.py:47-158
    eval_approx_score_mat_for_all_topk     eval/run_retrieval_eval_wrt_exact_crossenc_w_fixed_train_test_splits.py:51-135
    eval_approx_score_mat                  ..._w_fixed_train_test_splits.py:138-206
    _get_indices_scores                    ..._w_fixed_train_test_splits.py:34-47
"""
from anncur_b200.eval_retrieval import (_get_indices_scores, eval_approx_score_mat,  # noqa: F401
                                        eval_approx_score_mat_for_all_topk, fixed_split_cur_eval,
                                        run_approx_eval_w_seed)
