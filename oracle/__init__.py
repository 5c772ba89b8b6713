"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the ANNCUR test-time search path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as the
checker (or as the CPU arm being timed), never as the thing shipped.  The product package
``anncur_b200`` never imports this package and has no CPU fallback.

Parity status
-------------
* ``oracle.cur_oracle`` (CURApprox algebra, flat-IP search, rerank loop, overlap metrics, the two
  eval functions): **pinned** -- checked against the reference's own functions executed in the build
  container (``oracle/ref_shim.py`` + ``oracle/make_golden.py``) and against the committed outputs
  of those runs in ``tests/golden/*.npz``.
* ``oracle.cur_oracle.adaptive_anncur`` (multi-round ANNCUR): **parity unpinned** -- the reference
  has no implementation; this is our own CPU restatement of that row.
* faiss ``IndexFlatIP.search`` is a third-party dependency that is absent from the reference tree
  and from this image (unpinned version); it is restated as exact inner-product top-k.
"""
