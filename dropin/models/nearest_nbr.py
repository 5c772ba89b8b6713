"""``models.nearest_nbr`` served by the B200 engine (reference: models/nearest_nbr.py:24-55).
Only the index factory is replaced; the BERT-embedding helpers of the reference module (:58-80) are
re-exported from the reference file when it is importable."""
from anncur_b200.nearest_nbr import FlatIPIndex, build_flat_or_ivff_index  # noqa: F401
from _overlay import load_shadowed
from . import __path__ as _pkg_path

try:
    _ref = load_shadowed(_pkg_path, "models", "nearest_nbr")
except Exception:                      # the reference module needs faiss at import time
    _ref = None
if _ref is not None:
    for _name in ("embed_tokenized_entities", "index_tokenized_entities"):
        if hasattr(_ref, _name):
            globals()[_name] = getattr(_ref, _name)
