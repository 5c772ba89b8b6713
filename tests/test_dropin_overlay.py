"""CPU: the dropin/ overlay packages shadow exactly the hot-path modules and leave the rest of the reference's
``eval`` / ``models`` packages reachable (runs fully only where the reference checkout exists)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ANNCUR_REFERENCE_ROOT", "/root/reference")

PROBE = r"""
import sys, types
for name in ("IPython", "matplotlib", "matplotlib.pyplot", "faiss"):
    m = types.ModuleType(name); m.embed = lambda *a, **k: None; sys.modules[name] = m
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
import eval.matrix_approx_zeshel as m1, models.nearest_nbr as m2, eval.eval_utils as m3, eval.anncur_eval as m4
import anncur_b200
assert m1.CURApprox is anncur_b200.CURApprox
assert m2.build_flat_or_ivff_index is anncur_b200.build_flat_or_ivff_index
assert m3.compute_overlap is anncur_b200.compute_overlap
assert m4.run_approx_eval_w_seed is anncur_b200.run_approx_eval_w_seed
print("HAVE_REF_PASSTHROUGH", hasattr(m3, "score_topk_preds"))
"""


def _run(paths):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(paths))
    return subprocess.run([sys.executable, "-c", PROBE], capture_output=True, text=True, env=env, timeout=300)


def test_overlay_without_reference():
    r = _run([os.path.join(ROOT, "dropin"), ROOT])
    assert r.returncode == 0, r.stderr
    assert "HAVE_REF_PASSTHROUGH False" in r.stdout


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "eval")), reason="reference checkout not present")
def test_overlay_in_front_of_reference():
    r = _run([os.path.join(ROOT, "dropin"), ROOT, REF])
    assert r.returncode == 0, r.stderr
    assert "HAVE_REF_PASSTHROUGH True" in r.stdout          # eval.eval_utils still offers the reference's other helpers
