"""CPU oracle for the per-image transform chain — TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and there only as the checker (or as the timed
CPU baseline), never as the thing shipped.  The product package
``image_processor_pipeline_b200`` never imports this package and raises when
its CUDA library is missing.

Tiers:

* ``oracle.ref_loader``  (T1) imports the REAL reference functions from
  ``/root/reference`` behind the shims in ``oracle/ref_shims`` — only possible
  in the build container; used by ``tests/golden/make_golden.py`` to produce the
  committed fixtures.
* ``oracle.libcalls``    (T2) in-memory functions that issue the *identical*
  cv2 / Pillow / numpy calls the reference's transforms issue (same order, same
  arguments) — the pixel parity authority; runs wherever cv2 and Pillow exist.
* ``oracle.intmodels``   (T3) pure-numpy integer restatements of the library
  arithmetic (Pillow 16.16 fixed-point affine, DIV255 paste, 8bpc LANCZOS
  resample, OpenCV BGR2HSV tables, inRange bound rounding, block-raster CCL
  label order).  Each is pinned against T2 by ``tests/test_oracle_*.py``.

Parity pinning: the reference ships no tests or golden vectors.
The oracle is pinned instead against outputs of the reference itself, run in
the build container by ``tests/golden/make_golden.py`` (fixtures committed in
``tests/golden/*.npz``), and against cv2 4.13.0 / Pillow 12.2.0 installed in the
image (the third-party libraries that carry the reference's arithmetic; the
reference pins no versions).
"""
