"""Time integration: ``evolve`` and ``discrete_step`` (mirror of
pyminiweather/solve/step.py:21-143), running on the fused sm_100a stage kernels.

``evolve`` runs one time step as two fused sweep kernels (or six stage kernels) on rotating device
buffers (no halo-fill kernels, no interpolation/flux/tendency arrays).

Sweep order.  The reference keeps it in the module global ``reverse_direction`` (step.py:18,103,143):
one flag for the whole process, so two simulations stepped alternately swap each other's sweep order.
Here the flag belongs to the device context of a fields object (``pmw_get/set_reverse_direction``): every
simulation alternates Z,X / X,Z on its own, starting Z,X -- for a single simulation per process, the only
way the reference's driver uses it, that is the same sequence.  The flag survives whatever happens to the
context behind a fields object (re-creation after ``params["dt"]`` changed, eviction of a foreign object's
context: ``_dispatch.py`` keeps it by object).  ``step.reverse_direction`` reads the flag of the last call;
assigning it (the reference's ``step.reverse_direction = False`` reset idiom) sets the flag of the NEXT call.
"""
from __future__ import annotations

import weakref

import numpy as np

from .._dispatch import (check_ic, direction_id, foreign_solver, is_native, remember_sweep_order,
                         writable_f64)
from .._lib import PMW_BUF_STATE, PMW_BUF_TMP
from ..ics.directions import Directions  # noqa: F401  (re-exported like the reference)

_last_solver = None  # weakref to the context of the last evolve() call


def __getattr__(name):  # PEP 562: ``step.reverse_direction`` when it has not been assigned
    if name == "reverse_direction":
        s = _last_solver() if _last_solver is not None else None
        return bool(s.reverse_direction) if s is not None else False
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


def _sweep_order(solver):
    """Apply a pending ``step.reverse_direction = X`` assignment to the context of this call."""
    global _last_solver
    g = globals()
    if "reverse_direction" in g:
        solver.reverse_direction = bool(g.pop("reverse_direction"))
    _last_solver = weakref.ref(solver)


# Strict (foreign-fields) mode: also bring state_tmp back after evolve.  It is scratch in
# the reference; leaving it on the device halves the PCIe traffic of the drop-in call.
SYNC_STATE_TMP = False
# Strict mode: row bands of the streamed step (0 = chosen from the grid, 1 = upload, step, download in sequence)
HOST_BANDS = 0


def discrete_step(params, fields, mesh, state_init, state_forcing, state_out, dt, direction) -> None:
    """One RK stage: halo fill on ``state_forcing`` (in place), interpolation, fluxes,
    tendencies, ``state_out[interior] = state_init[interior] + dt * tend`` (step.py:63-82).
    ``state_out`` may alias either input, as in step.py:112-141."""
    check_ic(params["ic_type"])
    d = direction_id(direction)
    shape = (4, params["nz"] + 2 * params["hs"], params["nx"] + 2 * params["hs"])
    for name, a in (("state_init", state_init), ("state_forcing", state_forcing), ("state_out", state_out)):
        writable_f64(a, shape, name)

    if is_native(fields):
        bufs = [fields.buffer_of(a) for a in (state_init, state_forcing, state_out)]
        if None not in bufs:
            solver = fields.device(params)
            solver.discrete_step(d, bufs[0], bufs[1], bufs[2], dt)
            fields.device_wrote(bufs[1], bufs[2])
            fields.sync_host(bufs[1], bufs[2])  # explicitly passed arrays are updated on return
            return
        solver = fields.device(params)
        fields.sync_host()
        fields._host_dirty[PMW_BUF_STATE] = fields._host_dirty[PMW_BUF_TMP] = True
    else:
        solver = foreign_solver(fields, params)

    # Generic path for arbitrary host arrays: forcing -> TMP, init -> STATE (or TMP when it is
    # the same array); BC, then the stage with out aliased to forcing; only the interior of the
    # result is written to state_out, after forcing has received its new halo cells.
    solver.upload(PMW_BUF_TMP, state_forcing)
    init_buf = PMW_BUF_TMP
    if state_init is not state_forcing:
        solver.upload(PMW_BUF_STATE, state_init)
        init_buf = PMW_BUF_STATE
    (solver.bc_x if d == 1 else solver.bc_z)(PMW_BUF_TMP)
    solver.download(PMW_BUF_TMP, out=state_forcing)
    solver.stage(d, init_buf, PMW_BUF_TMP, PMW_BUF_TMP, dt)
    res = solver.download(PMW_BUF_TMP)
    hs = params["hs"]
    state_out[:, hs:-hs, hs:-hs] = res[:, hs:-hs, hs:-hs]


def evolve(params, fields, mesh, dt: float = 1e-4) -> None:
    """Advance by one time step ``dt``: two directional sweeps (Z,X then X,Z, alternating per
    call) of three RK stages each (step.py:85-143)."""
    check_ic(params["ic_type"])
    if is_native(fields):
        solver = fields.device(params)
        _sweep_order(solver)
        solver.evolve(1, dt)
        fields.device_wrote(PMW_BUF_STATE, PMW_BUF_TMP)
        return
    solver = foreign_solver(fields, params)
    _sweep_order(solver)
    shape = (4, params["nz"] + 2 * params["hs"], params["nx"] + 2 * params["hs"])
    state = writable_f64(fields.state, shape, "fields.state")
    if params["ic_type"] != "injection" and not SYNC_STATE_TMP:
        # the whole call as one band-pipelined pass over PCIe (pmw_evolve_host): upload, sweeps and download
        # of successive row bands overlap; same bits as the sequence below
        solver.evolve_host(state, dt, HOST_BANDS)
        remember_sweep_order(fields, solver)
        return
    solver.upload(PMW_BUF_STATE, state)
    if params["ic_type"] == "injection":
        # the right halo columns of state_tmp are never refreshed in this configuration (bcs.py:37):
        # they are caller data that the x stages read
        solver.upload(PMW_BUF_TMP, writable_f64(fields.state_tmp, shape, "fields.state_tmp"))
    solver.evolve(1, dt)
    remember_sweep_order(fields, solver)
    solver.download(PMW_BUF_STATE, out=state)
    if SYNC_STATE_TMP:
        solver.download(PMW_BUF_TMP, out=writable_f64(fields.state_tmp, shape, "fields.state_tmp"))
