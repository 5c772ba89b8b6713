"""``eval.eval_utils`` with ``compute_overlap`` (reference: eval/eval_utils.py:115-150) served by the B200
engine; every other name of the reference module is passed through unchanged when it is importable."""
from _overlay import load_shadowed
from . import __path__ as _pkg_path

try:
    _ref = load_shadowed(_pkg_path, "eval", "eval_utils")
except Exception:
    _ref = None
if _ref is not None:
    globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("__")})

from anncur_b200.eval_retrieval import compute_overlap  # noqa: E402,F401
