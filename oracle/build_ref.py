"""TEST / BENCH INFRASTRUCTURE ONLY.  Makes ``oracle/_ref/``: a verbatim copy of the reference file that holds the timed
path, ``eval/matrix_approx_zeshel.py`` (``CURApprox.topk_in_row`` :121-126 = ``get_complete_row`` :109-119 + ``torch.topk``),
so that ``bench.py --impl reference`` and the ``cpu_baseline`` leg execute THE REFERENCE'S OWN CODE on the GPU box, where
``/root/reference`` does not exist.

    python oracle/build_ref.py              # needs /root/reference; also run by __graft_entry__.build()

The reference is pure Python: there is nothing to compile.  ``oracle/_ref/`` is git-ignored (no reference sources in the
history) and NOT gpurun-ignored (it travels with the snapshot, like a built .so).  The copy is byte-identical
(sha256 in ``oracle/_ref/MANIFEST.json``); it is loaded with asserts stripped (== ``python -O``) because of the broken
assert at :44, see ``oracle/ref_shim.py``.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("ANNCUR_REFERENCE_ROOT", "/root/reference")
FILES = ["eval/matrix_approx_zeshel.py"]


def build(verbose=True):
    out = os.path.join(HERE, "_ref")
    if not os.path.isdir(REFERENCE_ROOT):
        if verbose:
            print(f"[build_ref] {REFERENCE_ROOT} not present: keeping whatever is in {out}")
        return os.path.isdir(out)
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REFERENCE_ROOT, rel), os.path.join(out, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as fh:
            manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(out, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REFERENCE_ROOT, "sha256": manifest}, fh, indent=1)
    if verbose:
        print(f"[build_ref] copied {len(FILES)} file(s) to {out}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
