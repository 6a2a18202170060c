"""Halo fill: periodic in x (or the injection inflow), solid wall in z (mirror of
pyminiweather/ics/bcs.py:8-148).

Both functions mutate ``state_forcing`` in place, like the reference.  The work is done by
``bc_x_kernel`` / ``bc_z_kernel`` (csrc/pmw_aux.cuh).  Inside ``evolve`` these kernels are
not launched at all: the x wrap is written by the producing stage and the wall rows are
rebuilt in shared memory by the z stage (csrc/pmw_tma.cuh).
"""
from __future__ import annotations

from .._dispatch import check_ic, foreign_solver, is_native, sync_inflow, writable_f64
from .._lib import PMW_BUF_STATE, PMW_BUF_TMP


def _apply(which, params, fields, state_forcing, ic_type):
    check_ic(ic_type)
    shape = (4, params["nz"] + 2 * params["hs"], params["nx"] + 2 * params["hs"])
    writable_f64(state_forcing, shape, "state_forcing")
    if is_native(fields):
        buf = fields.buffer_of(state_forcing)
        solver = fields.device(params)
        sync_inflow(solver, params, ic_type)  # the reference branches on the ARGUMENT (bcs.py:37,41)
        fields._source_ic = None              # ... so re-derive the context's setting on the next call
        if buf is not None:
            getattr(solver, which)(buf)
            fields.device_wrote(buf)
            fields.sync_host(buf)  # an explicitly passed array is updated before returning
            return
        # an array that is not one of the two field buffers: borrow TMP's device buffer
        fields.sync_host(PMW_BUF_TMP)
        solver.upload(PMW_BUF_TMP, state_forcing)
        getattr(solver, which)(PMW_BUF_TMP)
        solver.download(PMW_BUF_TMP, out=state_forcing)
        fields._host_dirty[PMW_BUF_TMP] = True  # restore TMP from its host copy on next use
        return
    solver = foreign_solver(fields, params)
    sync_inflow(solver, params, ic_type)
    solver.upload(PMW_BUF_TMP, state_forcing)
    getattr(solver, which)(PMW_BUF_TMP)
    solver.download(PMW_BUF_TMP, out=state_forcing)


def set_bc_x(params, fields, state_forcing, ic_type):
    """Periodic wrap of the two halo columns on every interior row (bcs.py:35-39); with
    ``ic_type == "injection"`` the right halo is left alone and the jet rows of the left halo are
    forced to u = 50 m/s, theta = 298 K (bcs.py:37,41-64)."""
    _apply("bc_x", params, fields, state_forcing, ic_type)


def set_bc_z(params, fields, state_forcing, ic_type):
    """Solid-wall halo rows over all nx+4 columns: w=0, u scaled by the hydrostatic density
    ratio, rho' and (rho*theta)' copied from the nearest interior row (bcs.py:92-148)."""
    _apply("bc_z", params, fields, state_forcing, ic_type)
