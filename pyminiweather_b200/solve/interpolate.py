"""The reference's individual spatial operators (mirror of pyminiweather/solve/interpolate.py:10-250)
as unfused GPU kernels that fill ``fields.vals_* / d3_vals_* / flux / tend`` on the host.

These are API-parity shims: ``evolve`` / ``discrete_step`` never call them (the fused stage kernels
keep interpolated values, fluxes and tendencies in registers).  Arithmetic follows the reference
expression by expression (no FMA contraction), so interpolation and tendencies are bit-identical to
the NumPy backend and fluxes differ only through ``pow``.
"""
from __future__ import annotations

import numpy as np

from .._dispatch import foreign_solver, is_native, writable_f64
from .._lib import PMW_BUF_STATE, PMW_BUF_TMP, PMW_DIR_X, PMW_DIR_Z


def _solver_and_buf(params, fields, state):
    """Device context plus the logical buffer that holds ``state`` (uploading it if foreign)."""
    shape = (4, params["nz"] + 2 * params["hs"], params["nx"] + 2 * params["hs"])
    if is_native(fields):
        solver = fields.device(params)
        if state is None:
            return solver, PMW_BUF_STATE
        buf = fields.buffer_of(state)
        if buf is not None:
            return solver, buf
        fields.sync_host(PMW_BUF_TMP)
        fields._host_dirty[PMW_BUF_TMP] = True
    else:
        solver = foreign_solver(fields, params)
        state = fields.state if state is None else state
    solver.upload(PMW_BUF_TMP, np.ascontiguousarray(state, dtype=np.float64).reshape(shape))
    return solver, PMW_BUF_TMP


def _dense(a, name):
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.flags.writeable):
        raise ValueError(f"fields.{name} must be a writable C-contiguous float64 array")
    return a


def interpolate_x(params, fields, state=None):
    """fields.vals_x, fields.d3_vals_x <- 4-point interface value / 3rd difference along x
    (interpolate.py:10-43)."""
    solver, buf = _solver_and_buf(params, fields, state)
    solver.interpolate(PMW_DIR_X, buf, _dense(fields.vals_x, "vals_x"), _dense(fields.d3_vals_x, "d3_vals_x"))


def interpolate_z(params, fields, state=None):
    """fields.vals_z, fields.d3_vals_z (interpolate.py:46-79)."""
    solver, buf = _solver_and_buf(params, fields, state)
    solver.interpolate(PMW_DIR_Z, buf, _dense(fields.vals_z, "vals_z"), _dense(fields.d3_vals_z, "d3_vals_z"))


def compute_flux_x(params, fields):
    """fields.flux[:, :nz, :nx+1] from fields.vals_x / d3_vals_x (interpolate.py:82-129)."""
    solver = fields.device(params) if is_native(fields) else foreign_solver(fields, params)
    solver.compute_flux(PMW_DIR_X, _dense(fields.vals_x, "vals_x"), _dense(fields.d3_vals_x, "d3_vals_x"),
                        _dense(fields.flux, "flux"))


def compute_flux_z(params, fields):
    """fields.flux[:, :nz+1, :nx] from fields.vals_z / d3_vals_z (interpolate.py:132-186); like the
    reference it also zeroes the density hyperviscosity rows at the two walls in fields.d3_vals_z."""
    solver = fields.device(params) if is_native(fields) else foreign_solver(fields, params)
    nz = params["nz"]
    fields.d3_vals_z[0, 0, :] = 0.0   # IDS.DENS
    fields.d3_vals_z[0, nz, :] = 0.0
    solver.compute_flux(PMW_DIR_Z, _dense(fields.vals_z, "vals_z"), _dense(fields.d3_vals_z, "d3_vals_z"),
                        _dense(fields.flux, "flux"))


def compute_tend_x(params, fields, state):
    """fields.tend = -(F[i+1] - F[i]) / dx (interpolate.py:189-215)."""
    solver, buf = _solver_and_buf(params, fields, state)
    solver.compute_tend(PMW_DIR_X, _dense(fields.flux, "flux"), buf, _dense(fields.tend, "tend"))


def compute_tend_z(params, fields, state):
    """fields.tend = -(F[k+1] - F[k]) / dz, minus rho' g on the w-momentum (interpolate.py:218-250)."""
    solver, buf = _solver_and_buf(params, fields, state)
    solver.compute_tend(PMW_DIR_Z, _dense(fields.flux, "flux"), buf, _dense(fields.tend, "tend"))
