"""B200-native hot path of PyMiniWeather (per-RK3-stage flux / tendency / update).

Host-side mirror of the reference's operator interface -- same module paths, names,
argument meaning and in-place semantics as ``pyminiweather.solve`` / ``.ics`` / ``.post`` /
``.data`` -- on top of ``libpmw.so`` (hand-written sm_100a CUDA behind a C ABI, see
``include/pmw.h``).  There is no CPU fallback: every operator raises if the library or a
CUDA device is missing.

Reference seam being replaced: ``pyminiweather/__init__.py:4-18`` (array-module alias + IDS).
"""
from enum import IntEnum

__version__ = "0.1.0"


class IDS(IntEnum):
    """Variable ids of the state array (pyminiweather/__init__.py:14-18)."""

    DENS = 0
    UMOM = 1
    WMOM = 2
    RHOT = 3
