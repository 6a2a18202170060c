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
``reference_runner``  helpers that import the *real* reference -- from
                  ``/root/reference`` in the build container, else from the
                  verbatim copy ``oracle/_ref`` -- to validate the restatements,
                  generate golden vectors and time the CPU baseline.
``make_ref``      the recipe for ``oracle/_ref`` (git-ignored copy of the
                  reference package; travels to the GPU box with the snapshot).

Parity status: PINNED -- see ``tests/golden/README.md``.
"""
