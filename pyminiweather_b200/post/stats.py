"""Diagnostics (mirror of pyminiweather/post/stats.py:8-69): domain mass / total energy by a
warp-shuffle block reduction (csrc/pmw_aux.cuh), and the derived output variables."""
from __future__ import annotations

import numpy as np

from .._dispatch import foreign_solver, is_native, writable_f64
from .._lib import PMW_BUF_STATE


def _solver_with_state(params, fields):
    if is_native(fields):
        return fields.device(params)
    solver = foreign_solver(fields, params)
    shape = (4, params["nz"] + 2 * params["hs"], params["nx"] + 2 * params["hs"])
    solver.upload(PMW_BUF_STATE, np.ascontiguousarray(fields.state, dtype=np.float64).reshape(shape))
    return solver


def compute_stats(params, fields):
    """(total_mass, total_energy) of ``fields.state`` over the interior, times dx*dz
    (stats.py:16-33; the kinetic term carries no 1/2, as in the reference)."""
    return _solver_with_state(params, fields).stats(PMW_BUF_STATE)


def compute_solution_variables(params, fields) -> np.ndarray:
    """[4, nz, nx]: rho', u, w, theta' (stats.py:38-69)."""
    return _solver_with_state(params, fields).solution_variables(PMW_BUF_STATE)
