"""Overlay of the reference's ``eval`` package: modules present here run on the B200 engine, the rest
resolve to the reference checkout that follows on sys.path (see dropin/_overlay.py; ``<repo>/dropin``
must be on PYTHONPATH, which is also what makes this package importable)."""
from _overlay import overlay_path

__path__ = overlay_path(__file__, "eval")
