"""TEST INFRASTRUCTURE ONLY.  Import the *reference's own* hot-path functions in-process.

Only usable where ``/root/reference`` exists (the build container); never on the GPU box.  Used by
``oracle/make_golden.py`` to generate ``tests/golden/*.npz`` and by ``tests/test_oracle_vs_reference.py``
(skipped when the reference tree is absent) to pin ``oracle.cur_oracle``.

Two accommodations, both found by probing:

1. modules that the reference imports but that are not in this image and play no part in the
   arithmetic (IPython, matplotlib, faiss, pytorch_transformers, pytorch_lightning, wandb) are
   replaced by inert stand-ins before import;
2. ``eval/matrix_approx_zeshel.py:44`` asserts on ``torch.eq(...)`` of a matrix, which raises for
   any intersection larger than 1x1.  The reference therefore only runs with asserts stripped
   (``python -O``).  When this interpreter was not started with ``-O`` we compile the reference
   modules ourselves with ``optimize=1`` -- the same byte code ``python -O`` would execute.
"""
import importlib.abc
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ANNCUR_REFERENCE_ROOT", "/root/reference")
# verbatim copy of the ONE reference file the timed path lives in, made by oracle/build_ref.py (git-ignored, travels to
# the GPU box with the snapshot like a built .so) -- what `bench.py --impl reference` and its cpu_baseline leg execute
LOCAL_REF_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "eval", "matrix_approx_zeshel.py"))


class _Inert:
    """Callable, attribute-able nothing."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, name):
        return _Inert()


def _stand_in(name, **attrs):
    if name in sys.modules and not getattr(sys.modules[name], "__anncur_stub__", False):
        return sys.modules[name]
    mod = types.ModuleType(name)
    mod.__anncur_stub__ = True
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def _install_stand_ins():
    import torch

    _stand_in("IPython", embed=lambda *a, **k: None)
    plt = _stand_in("matplotlib.pyplot")
    _stand_in("matplotlib", pyplot=plt)
    _stand_in("faiss")
    _stand_in("wandb", init=_Inert(), log=_Inert(), run=_Inert(), login=_Inert())
    _stand_in("pytorch_transformers")
    _stand_in("pytorch_transformers.modeling_bert", BertModel=_Inert)
    _stand_in("pytorch_transformers.tokenization_bert", BertTokenizer=_Inert)
    _stand_in("pytorch_transformers.optimization", AdamW=_Inert, WarmupLinearSchedule=_Inert)

    class _LightningModule(torch.nn.Module):
        pass

    _stand_in("pytorch_lightning", LightningModule=_LightningModule, Trainer=_Inert,
              LightningDataModule=object, seed_everything=lambda *a, **k: None)
    _stand_in("pytorch_lightning.callbacks", ModelCheckpoint=_Inert, LearningRateMonitor=_Inert)
    _stand_in("pytorch_lightning.loggers", WandbLogger=_Inert)


class _StripAssertsLoader(importlib.abc.SourceLoader):
    """Loads a reference source file compiled with optimize=1 (== what ``python -O`` runs)."""

    def __init__(self, fullname, path):
        self._fullname, self._path = fullname, path

    def get_filename(self, fullname):
        return self._path

    def get_data(self, path):
        with open(path, "rb") as fh:
            return fh.read()

    def source_to_code(self, data, path, *, _optimize=-1):
        return compile(data, path, "exec", dont_inherit=True, optimize=1)

    def set_data(self, path, data):  # never write .pyc next to the read-only reference
        raise NotImplementedError


def _load_reference_module(dotted):
    """Import ``dotted`` (e.g. 'eval.matrix_approx_zeshel') from the reference tree, asserts stripped.

    The reference's top-level packages are called ``eval``/``models``/``utils``; they are registered
    under exactly those names because the reference files import each other that way."""
    if dotted in sys.modules and getattr(sys.modules[dotted], "__anncur_reference__", False):
        return sys.modules[dotted]
    parts = dotted.split(".")
    for i in range(1, len(parts)):
        pkg = ".".join(parts[:i])
        if pkg not in sys.modules or not getattr(sys.modules[pkg], "__anncur_reference__", False):
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(REFERENCE_ROOT, *parts[:i])]
            m.__anncur_reference__ = True
            sys.modules[pkg] = m
    path = os.path.join(REFERENCE_ROOT, *parts) + ".py"
    loader = _StripAssertsLoader(dotted, path)
    spec = importlib.util.spec_from_loader(dotted, loader, origin=path)
    mod = importlib.util.module_from_spec(spec)
    mod.__anncur_reference__ = True
    sys.modules[dotted] = mod
    loader.exec_module(mod)
    return mod


class _Finder(importlib.abc.MetaPathFinder):
    """Routes the reference's own intra-package imports (eval.*, models.*, utils.*) through the
    assert-stripping loader so that nested imports get the same treatment."""

    _roots = ("eval", "models", "utils")

    def find_spec(self, fullname, path=None, target=None):
        head = fullname.split(".")[0]
        if head not in self._roots:
            return None
        cand = os.path.join(REFERENCE_ROOT, *fullname.split("."))
        if os.path.isfile(cand + ".py"):
            return importlib.util.spec_from_loader(fullname, _StripAssertsLoader(fullname, cand + ".py"),
                                                   origin=cand + ".py")
        if os.path.isdir(cand):
            spec = importlib.util.spec_from_loader(fullname, loader=None, is_package=True)
            spec.submodule_search_locations = [cand]
            return spec
        return None


_FINDER = None


def load_reference():
    """Return a namespace with the reference's hot-path callables (executed verbatim, asserts off)."""
    global _FINDER
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stand_ins()
    if _FINDER is None:
        _FINDER = _Finder()
        sys.meta_path.insert(0, _FINDER)
    import importlib

    mat = importlib.import_module("eval.matrix_approx_zeshel")
    utils = importlib.import_module("eval.eval_utils")
    sweep = importlib.import_module("eval.run_retrieval_eval_wrt_exact_crossenc")
    split = importlib.import_module("eval.run_retrieval_eval_wrt_exact_crossenc_w_fixed_train_test_splits")
    return types.SimpleNamespace(
        CURApprox=mat.CURApprox,
        compute_overlap=utils.compute_overlap,
        run_approx_eval_w_seed=sweep.run_approx_eval_w_seed,
        eval_approx_score_mat_for_all_topk=split.eval_approx_score_mat_for_all_topk,
        eval_approx_score_mat=split.eval_approx_score_mat,
        modules=types.SimpleNamespace(mat=mat, utils=utils, sweep=sweep, split=split),
    )


def local_reference_available():
    return os.path.isfile(os.path.join(LOCAL_REF_ROOT, "eval", "matrix_approx_zeshel.py"))


def load_reference_curapprox():
    """The reference's own ``CURApprox`` class (eval/matrix_approx_zeshel.py:20-126) from oracle/_ref, executed verbatim with
    asserts stripped (see the module docstring), or None when oracle/_ref has not been built.  Loaded under a private module
    name so that it never collides with the in-container reference import above."""
    if not local_reference_available():
        return None
    _stand_in("IPython", embed=lambda *a, **k: None)
    plt = _stand_in("matplotlib.pyplot")
    _stand_in("matplotlib", pyplot=plt)
    name = "_anncur_ref_matrix_approx_zeshel"
    if name in sys.modules:
        return sys.modules[name].CURApprox
    path = os.path.join(LOCAL_REF_ROOT, "eval", "matrix_approx_zeshel.py")
    loader = _StripAssertsLoader(name, path)
    spec = importlib.util.spec_from_loader(name, loader, origin=path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    loader.exec_module(mod)
    return mod.CURApprox
