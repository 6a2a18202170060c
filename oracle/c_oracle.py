"""ctypes wrapper of the C oracle (``oracle/c/pmw_oracle.c``).  TEST INFRASTRUCTURE ONLY.

Operates on ``numpy_oracle.OracleCase`` objects so the two oracles are
interchangeable in tests and in bench.py's CPU-baseline leg.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpmw_oracle.so")
_lib = None

_dp = C.POINTER(C.c_double)


class _Case(C.Structure):
    _fields_ = [("nx", C.c_int), ("nz", C.c_int),
                ("dx", C.c_double), ("dz", C.c_double), ("dt", C.c_double),
                ("hy_dens_cell", _dp), ("hy_dens_theta_cell", _dp),
                ("hy_dens_int", _dp), ("hy_dens_theta_int", _dp), ("hy_pressure_int", _dp),
                ("flux", _dp), ("tend", _dp), ("source_w", _dp),
                ("inflow_rows", C.POINTER(C.c_ubyte))]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "c", "pmw_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "c")])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.pmwo_evolve.argtypes = [C.POINTER(_Case), _dp, _dp, C.c_double, C.c_int, C.POINTER(C.c_int)]
        _lib.pmwo_discrete_step.argtypes = [C.POINTER(_Case), _dp, _dp, _dp, C.c_double, C.c_int]
        _lib.pmwo_set_bc_x.argtypes = [C.POINTER(_Case), _dp]
        _lib.pmwo_set_bc_z.argtypes = [C.POINTER(_Case), _dp]
        _lib.pmwo_stats.argtypes = [C.POINTER(_Case), _dp, _dp]
        _lib.pmwo_set_threads.argtypes = [C.c_int]
        _lib.pmwo_set_threads.restype = C.c_int
        _lib.pmwo_team_size.restype = C.c_int
    return _lib


def set_threads(n: int) -> int:
    """Set the OpenMP team size explicitly (launchers export OMP_NUM_THREADS=1) and return the size of the
    team a parallel region really gets."""
    lib().pmwo_set_threads(int(n))
    return int(lib().pmwo_team_size())


def _p(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


class COracle:
    """Binds an ``OracleCase`` (arrays are used in place, not copied)."""

    def __init__(self, case):
        self.case = case
        self._flux = np.zeros(4 * (case.nz + 1) * (case.nx + 1))
        self._tend = np.zeros(4 * case.nz * case.nx)
        self._inflow = None
        if getattr(case, "inflow_zlen", None) is not None:  # injection: row mask of the jet (bcs.py:43-48)
            from .numpy_oracle import inflow_rows
            self._inflow = np.zeros(case.nz + 4, dtype=np.uint8)
            self._inflow[inflow_rows(case.nz, case.dz, case.inflow_zlen)] = 1
        self._c = _Case(case.nx, case.nz, case.dx, case.dz, case.dt,
                        _p(case.hy_dens_cell), _p(case.hy_dens_theta_cell),
                        _p(case.hy_dens_int), _p(case.hy_dens_theta_int),
                        _p(case.hy_pressure_int), _p(self._flux), _p(self._tend),
                        _p(case.source_w) if getattr(case, "source_w", None) is not None else None,
                        self._inflow.ctypes.data_as(C.POINTER(C.c_ubyte)) if self._inflow is not None else None)

    def evolve(self, nsteps: int = 1, dt: float | None = None):
        rev = C.c_int(1 if self.case.reverse_direction else 0)
        lib().pmwo_evolve(C.byref(self._c), _p(self.case.state), _p(self.case.state_tmp),
                          self.case.dt if dt is None else dt, nsteps, C.byref(rev))
        self.case.reverse_direction = bool(rev.value)

    def discrete_step(self, init, forcing, out, dt, direction):
        lib().pmwo_discrete_step(C.byref(self._c), _p(init), _p(forcing), _p(out), dt, direction)

    def set_bc_x(self, s):
        lib().pmwo_set_bc_x(C.byref(self._c), _p(s))

    def set_bc_z(self, s):
        lib().pmwo_set_bc_z(C.byref(self._c), _p(s))

    def stats(self, s=None):
        out = np.zeros(2)
        lib().pmwo_stats(C.byref(self._c), _p(self.case.state if s is None else s), _p(out))
        return float(out[0]), float(out[1])
