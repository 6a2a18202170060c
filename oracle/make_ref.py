"""Recipe for ``oracle/_ref``: a verbatim, git-ignored copy of the reference package so that the
reference itself can be the checker and the timed CPU baseline on the GPU box, where
``/root/reference`` does not exist.  TEST / BASELINE INFRASTRUCTURE ONLY: nothing under
``pyminiweather_b200/`` may import it (``tests/test_cabi.py`` checks that).

    python oracle/make_ref.py          # copies /root/reference/pyminiweather -> oracle/_ref/pyminiweather

``oracle/_ref/`` is listed in ``.gitignore`` (reference sources never enter the history) but not in
``.gpurunignore``: the copy travels to the GPU box with the snapshot, like the built ``.so`` files.
``__graft_entry__.build()`` runs this whenever ``/root/reference`` is mounted.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("PMW_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")


def make(force: bool = False) -> str | None:
    """Copy the reference package (pure Python, ~130 KB) and its unit tests; returns the copy's root or
    None when the reference is not mounted here."""
    src_pkg = os.path.join(SRC, "pyminiweather")
    if not os.path.isdir(src_pkg):
        return DST if os.path.isdir(os.path.join(DST, "pyminiweather")) else None
    if os.path.isdir(os.path.join(DST, "pyminiweather")) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc")
    shutil.copytree(src_pkg, os.path.join(DST, "pyminiweather"), ignore=ignore)
    tests = os.path.join(SRC, "tests", "unit")
    if os.path.isdir(tests):
        shutil.copytree(tests, os.path.join(DST, "tests", "unit"), ignore=ignore)
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Verbatim copy of the reference (shriram-jagan/pyminiweather) made by oracle/make_ref.py.\n"
                "Checker and timed CPU baseline only; git-ignored; never imported by pyminiweather_b200.\n")
    return DST


if __name__ == "__main__":
    out = make(force="--force" in sys.argv)
    print(out or f"reference not found under {SRC} and no earlier copy under {DST}")
    sys.exit(0 if out else 1)
