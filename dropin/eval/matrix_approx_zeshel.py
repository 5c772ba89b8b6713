"""``eval.matrix_approx_zeshel`` served by the B200 engine (reference: eval/matrix_approx_zeshel.py:19-126)."""
from anncur_b200.matrix_approx import CURApprox  # noqa: F401

__all__ = ["CURApprox"]
