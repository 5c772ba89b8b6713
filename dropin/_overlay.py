"""Helper for the drop-in packages: make ``eval`` / ``models`` overlay packages.

Put ``<repo>/dropin`` and ``<repo>`` *in front of* the reference checkout on ``PYTHONPATH``.  A module that
exists here (``eval/matrix_approx_zeshel.py`` ...) then shadows the reference's module of the same name;
every other ``eval.*`` / ``models.*`` module is still found in the reference tree, because the overlay
package appends the reference's directory of the same name to its ``__path__``."""
import importlib.util
import os
import sys


def overlay_path(pkg_file, pkg_name):
    here = os.path.dirname(os.path.abspath(pkg_file))
    paths = [here]
    for p in sys.path:
        cand = os.path.abspath(os.path.join(p or ".", pkg_name))
        if cand != here and os.path.isdir(cand) and cand not in paths:
            paths.append(cand)
    return paths


def load_shadowed(pkg_path, pkg_name, mod_name):
    """The reference's own ``pkg_name.mod_name`` (the file our module shadows), or None if absent."""
    for d in pkg_path[1:]:
        f = os.path.join(d, mod_name + ".py")
        if os.path.isfile(f):
            spec = importlib.util.spec_from_file_location(f"_shadowed_{pkg_name}_{mod_name}", f)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
    return None
