"""Extra tendency sources (mirror of pyminiweather/solve/source.py:53-75).  Only ``ic_type ==
"gravity"`` has one in the reference; that configuration is not on the accelerated path yet."""
from .._dispatch import check_ic


def add_source_terms(params, mesh, fields):
    check_ic(params["ic_type"])  # raises NotImplementedError for "gravity"; a no-op otherwise
