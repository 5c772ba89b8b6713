"""anncur_b200 -- B200-native engine for the ANNCUR test-time search path of iesl/anncur.

Host-side mirror of the reference's call surface (``CURApprox``, ``build_flat_or_ivff_index``,
the retrieve/rerank/overlap eval functions) over hand-written sm_100a CUDA kernels reached through
the C ABI in ``include/anncur_b200.h`` (``libanncur_b200.so``).  No CPU fallback.
"""
from . import _lib  # noqa: F401
from .adaptive import adaptive_anncur  # noqa: F401
from .matrix_approx import CURApprox  # noqa: F401
from .nearest_nbr import FlatIPIndex, build_flat_or_ivff_index  # noqa: F401
from .eval_retrieval import (compute_overlap, eval_approx_score_mat,  # noqa: F401
                             eval_approx_score_mat_for_all_topk, fixed_split_cur_eval, run_approx_eval_w_seed)

__all__ = ["CURApprox", "adaptive_anncur", "FlatIPIndex", "build_flat_or_ivff_index", "compute_overlap", "eval_approx_score_mat",
           "eval_approx_score_mat_for_all_topk", "fixed_split_cur_eval", "run_approx_eval_w_seed"]
