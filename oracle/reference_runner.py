"""Drive the REAL reference in-process: ``/root/reference`` (read-only mount of the build
container) or, where that does not exist (the GPU box), the verbatim copy ``oracle/_ref`` that
``oracle/make_ref.py`` makes (git-ignored; travels with the gpurun snapshot).
TEST / BASELINE INFRASTRUCTURE ONLY: the checker in ``tests/`` and the timed CPU baseline of
``bench.py``; the product never imports it.

Used for two things:
  * validating ``oracle.numpy_oracle`` / ``oracle/c`` against the reference's own
    NumPy backend (``tests/test_oracle_vs_reference.py``), and
  * generating the golden fixtures under ``tests/golden/`` (``make_golden.py``).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

_MOUNT = os.environ.get("PMW_REFERENCE_ROOT", "/root/reference")
_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = _MOUNT if os.path.isdir(os.path.join(_MOUNT, "pyminiweather")) else _COPY


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pyminiweather"))


def kind() -> str:
    """Where the reference comes from: 'mount' (/root/reference) or 'copy' (oracle/_ref)."""
    return "mount" if REFERENCE_ROOT == _MOUNT else "copy"


def _import_reference():
    """Import the reference package with the NumPy backend (its __init__ picks
    cupynumeric only when both LEGATE_* variables are set, __init__.py:4)."""
    if not available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    for k in ("LEGATE_MAX_DIM", "LEGATE_MAX_FIELDS"):
        os.environ.pop(k, None)
    sys.dont_write_bytecode = True  # the mount is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with contextlib.redirect_stdout(io.StringIO()):  # it prints the backend name
        import pyminiweather  # noqa: F401
        import pyminiweather.solve.step as step
    return step


def make_params(nx: int, nz: int, ic_type: str = "thermal", xlen: float = 2e4,
                zlen: float = 1e4, dt: float | None = None, cfl: float = 1.0,
                max_speed: float = 500.0) -> dict:
    """The params dict that ``__main__.py:160-195`` builds."""
    p = dict(nx=nx, nz=nz, xlen=xlen, zlen=zlen, dt=dt, nsteps=0, nwarmups=0,
             ic_type=ic_type, hs=2, s=4, max_speed=max_speed, cfl=cfl,
             output_freq=-1, app_filename="PyMiniWeatherData.txt",
             app_log_file=None, verbose=False)
    p["dx"] = p["xlen"] / p["nx"]
    p["dz"] = p["zlen"] / p["nz"]
    if p["dt"] is None:
        p["dt"] = np.minimum(p["dx"], p["dz"]) * p["cfl"] / p["max_speed"]
    return p


class ReferenceRun:
    """One reference simulation: fields + mesh + params, stepped by the
    reference's own ``evolve``."""

    def __init__(self, nx, nz, ic_type="thermal", **kw):
        self.step = _import_reference()
        from pyminiweather.data import initialize_fields
        from pyminiweather.ics import init
        from pyminiweather.mesh import MeshData

        self.params = make_params(nx, nz, ic_type, **kw)
        self.fields = initialize_fields(self.params)
        self.mesh = MeshData(self.params)
        init(self.fields, self.params, self.mesh)
        # module-global direction flag (step.py:18): reset for every run
        self.step.reverse_direction = False
        self._reverse = False

    def evolve(self, nsteps: int = 1):
        self.step.reverse_direction = self._reverse
        for _ in range(nsteps):
            self.step.evolve(self.params, self.fields, self.mesh, dt=self.params["dt"])
        self._reverse = self.step.reverse_direction

    def discrete_step(self, init, forcing, out, dt, direction):
        from pyminiweather.ics import Directions
        d = Directions.X if direction in (1, "x", "X") else Directions.Z
        self.step.discrete_step(self.params, self.fields, self.mesh, init, forcing, out, dt, d)

    def stats(self):
        from pyminiweather.post import compute_stats
        m, e = compute_stats(self.params, self.fields)
        return float(m), float(e)

    def to_oracle_case(self):
        """Snapshot the current reference arrays into an ``OracleCase``."""
        from .numpy_oracle import OracleCase, gravity_source
        f, p = self.fields, self.params
        src = None
        if p["ic_type"] == "gravity":
            src = gravity_source(p["nx"], p["nz"], float(p["dx"]), float(p["dz"]), p["xlen"], p["zlen"],
                                 np.asarray(f.hy_dens_cell))
        return OracleCase(
            nx=p["nx"], nz=p["nz"], dx=float(p["dx"]), dz=float(p["dz"]), dt=float(p["dt"]),
            state=np.array(f.state, copy=True), state_tmp=np.array(f.state_tmp, copy=True),
            hy_dens_cell=np.array(f.hy_dens_cell, copy=True),
            hy_dens_theta_cell=np.array(f.hy_dens_theta_cell, copy=True),
            hy_dens_int=np.array(f.hy_dens_int, copy=True),
            hy_dens_theta_int=np.array(f.hy_dens_theta_int, copy=True),
            hy_pressure_int=np.array(f.hy_pressure_int, copy=True),
            reverse_direction=self._reverse, source_w=src,
            inflow_zlen=float(p["zlen"]) if p["ic_type"] == "injection" else None,
        )
