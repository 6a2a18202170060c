"""CPU oracle for the PyMiniWeather hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the shipped product path (``pyminiweather_b200``) may import this
package.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker or as
the timed CPU baseline.

Contents
--------
``numpy_oracle``  slicing-only NumPy restatement of the reference's per-stage
                  math (bit-identical to the reference's NumPy backend; pinned
                  by ``tests/test_oracle_vs_reference.py`` when ``/root/reference``
                  is mounted and by the fixtures under ``tests/golden/`` always).
``c/``            plain-C (OpenMP) restatement of the same path, used for
                  mid-size parity runs and as the multi-threaded CPU baseline.
``reference_runner``  helpers that import the *real* reference from
                  ``/root/reference`` (only inside the build container) to
                  validate the restatements and generate golden vectors.

Parity status: PINNED -- see ``tests/golden/README.md``.
"""
