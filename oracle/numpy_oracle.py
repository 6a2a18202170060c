"""NumPy restatement of the PyMiniWeather hot path.  TEST INFRASTRUCTURE ONLY.

This is the *oracle*: a loop-free, slicing-only restatement of what the
reference computes per RK stage, written so that every floating-point operation
happens in the same order as in the reference's NumPy/SciPy backend (the
reference spends 84 % of its time in ``scipy.signal.convolve`` -> ``_correlateND``;
here the two 4-tap correlations are spelled out as shifted-slice sums with the
same left-to-right accumulation).  It is bit-identical to the reference on full
arrays including halos -- that claim is checked by
``tests/test_oracle_vs_reference.py`` (when ``/root/reference`` is mounted) and
by the committed fixtures in ``tests/golden/`` (always).

Reference locations (relative to the reference repo root):

* constants ............ ``pyminiweather/data/constants.py:4-27``
* stencil kernels ...... ``pyminiweather/data/fields.py:94-97``
* ``set_bc_x`` ......... ``pyminiweather/ics/bcs.py:8-64``   (periodic branch :35-39)
* ``set_bc_z`` ......... ``pyminiweather/ics/bcs.py:67-148``
* ``interpolate_x/z`` .. ``pyminiweather/solve/interpolate.py:10-79``
* ``compute_flux_x/z`` . ``pyminiweather/solve/interpolate.py:82-186``
* ``compute_tend_x/z`` . ``pyminiweather/solve/interpolate.py:189-250``
* ``discrete_step`` .... ``pyminiweather/solve/step.py:21-82``
* ``evolve`` ........... ``pyminiweather/solve/step.py:85-143``
* ``compute_stats`` .... ``pyminiweather/post/stats.py:8-35``

State layout everywhere: ``[4, nz+4, nx+4]`` float64, C order, variable ids
DENS=0, UMOM=1, WMOM=2, RHOT=3 (``pyminiweather/__init__.py:14-18``).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# ---- constants.py:4-27 ------------------------------------------------------
HV_BETA = 0.05
P0 = 1.0e5
C0 = 27.5629410929725921310572974482
GAMMA = 1.40027894002789400278940027894
GRAV = 9.8
CP = 1004.0
CV = 717.0
RD = 287.0

DENS, UMOM, WMOM, RHOT = 0, 1, 2, 3
HS = 2
DIR_X, DIR_Z = 1, 2  # ics/directions.py:4-6

# fields.py:94-96 -- symmetric, so the convolution flip is invisible
_C4 = np.array([-1.0 / 12, 7.0 / 12, 7.0 / 12, -1.0 / 12], dtype=np.float64)


@dataclass
class OracleCase:
    """Plain container for one simulation (mirrors the arrays of fields.py:7-55
    that the hot path reads or writes)."""

    nx: int
    nz: int
    dx: float
    dz: float
    dt: float
    state: np.ndarray
    state_tmp: np.ndarray
    hy_dens_cell: np.ndarray
    hy_dens_theta_cell: np.ndarray
    hy_dens_int: np.ndarray
    hy_dens_theta_int: np.ndarray
    hy_pressure_int: np.ndarray
    # step.py:18 is a module global in the reference; here it is per case.
    reverse_direction: bool = False
    # ic_type == "gravity" only: wpert(x,z) * hy_dens_cell[2:nz+2, None], added to the w-momentum
    # tendency in EVERY stage, both directions (source.py:43-50, called at step.py:78)
    source_w: np.ndarray | None = None
    # ic_type == "injection" only: the domain height; switches set_bc_x to the inflow branch
    # (bcs.py:37,41-64).  None = periodic x.
    inflow_zlen: float | None = None
    scratch: dict = field(default_factory=dict)

    def copy(self) -> "OracleCase":
        return OracleCase(
            self.nx, self.nz, self.dx, self.dz, self.dt,
            self.state.copy(), self.state_tmp.copy(),
            self.hy_dens_cell.copy(), self.hy_dens_theta_cell.copy(),
            self.hy_dens_int.copy(), self.hy_dens_theta_int.copy(),
            self.hy_pressure_int.copy(), self.reverse_direction,
            None if self.source_w is None else self.source_w.copy(),
            self.inflow_zlen,
        )


# ---- boundary conditions ------------------------------------------------------
def inflow_rows(nz: int, dz: float, zlen: float) -> np.ndarray:
    """Array rows (halo offset included) of the injection jet, bcs.py:43-48.  The reference samples
    ``z = linspace(0, nz*dz, nz, endpoint=False) + 0.5`` (cell BOTTOMS plus half a metre, not cell
    centres) and keeps rows with |z - 3/4 zlen| <= zlen/16."""
    z = np.linspace(start=0, stop=nz * dz, num=nz, endpoint=False) + 0.5
    return np.nonzero(np.fabs(z - 3.0 * zlen / 4.0) <= zlen / 16.0)[0] + HS


def set_bc_x(case: OracleCase, s: np.ndarray) -> None:
    """Periodic wrap of the two halo columns, interior rows only (bcs.py:35-39); for the injection
    configuration the right halo is left alone and the left halo of the jet rows is forced to
    u = 50 m/s, theta = 298 K (bcs.py:37,41-64)."""
    nx, nz = case.nx, case.nz
    rows = slice(HS, nz + HS)
    s[:, rows, 0] = s[:, rows, nx]
    s[:, rows, 1] = s[:, rows, nx + 1]
    if case.inflow_zlen is None:
        s[:, rows, nx + HS] = s[:, rows, HS]
        s[:, rows, nx + HS + 1] = s[:, rows, HS + 1]
        return
    idx = inflow_rows(nz, case.dz, case.inflow_zlen)
    for col in (0, 1):
        s[UMOM, idx, col] = (s[DENS, idx, col] + case.hy_dens_cell[idx]) * 50.0
    for col in (0, 1):
        s[RHOT, idx, col] = (s[DENS, idx, col] + case.hy_dens_cell[idx]) * 298.0 - case.hy_dens_theta_cell[idx]


def set_bc_z(case: OracleCase, s: np.ndarray) -> None:
    """Solid wall in z over all nx+4 columns (bcs.py:92-148)."""
    nz = case.nz
    hd = case.hy_dens_cell
    top = nz + HS - 1  # last interior row
    # W momentum: zero (bcs.py:92-95)
    for r in (0, 1, nz + HS, nz + HS + 1):
        s[WMOM, r, :] = 0.0
    # U momentum: divide by the interior density, then multiply (bcs.py:98-118)
    s[UMOM, 0, :] = s[UMOM, HS, :] / hd[HS] * hd[0]
    s[UMOM, 1, :] = s[UMOM, HS, :] / hd[HS] * hd[1]
    s[UMOM, nz + HS, :] = s[UMOM, top, :] / hd[top] * hd[nz + HS]
    s[UMOM, nz + HS + 1, :] = s[UMOM, top, :] / hd[top] * hd[nz + HS + 1]
    # density and rho*theta: copy nearest interior row (bcs.py:121-148)
    for v in (DENS, RHOT):
        s[v, 0, :] = s[v, HS, :]
        s[v, 1, :] = s[v, HS, :]
        s[v, nz + HS, :] = s[v, top, :]
        s[v, nz + HS + 1, :] = s[v, top, :]


# ---- interpolation (the two 4-tap correlations) --------------------------------
def interpolate_x(case: OracleCase, s: np.ndarray):
    """``vals_x, d3_vals_x`` of shape [4, nz, nx+1] (interpolate.py:33-43).

    ``convolve(R, k[None,None,:], "same")[:, :, 2:-1]`` with R = s[:, 2:nz+2, :]
    is, at interface i, sum_j w_j * R[..., i+j] for j=0..3 with the *flipped*
    kernel w; accumulation is left to right starting from the first product.
    """
    nx, nz = case.nx, case.nz
    r = s[:, HS:nz + HS, :]
    s0, s1, s2, s3 = (r[:, :, j:j + nx + 1] for j in range(4))
    vals = ((_C4[0] * s0 + _C4[1] * s1) + _C4[2] * s2) + _C4[3] * s3
    # stored kernel [1,-3,3,-1] (fields.py:97) flipped -> weights (-1, 3, -3, 1)
    d3 = ((-1.0 * s0 + 3.0 * s1) + -3.0 * s2) + 1.0 * s3
    return vals, d3


def interpolate_z(case: OracleCase, s: np.ndarray):
    """``vals_z, d3_vals_z`` of shape [4, nz+1, nx] (interpolate.py:69-79)."""
    nx, nz = case.nx, case.nz
    r = s[:, :, HS:nx + HS]
    s0, s1, s2, s3 = (r[:, j:j + nz + 1, :] for j in range(4))
    vals = ((_C4[0] * s0 + _C4[1] * s1) + _C4[2] * s2) + _C4[3] * s3
    d3 = ((-1.0 * s0 + 3.0 * s1) + -3.0 * s2) + 1.0 * s3
    return vals, d3


# ---- fluxes ---------------------------------------------------------------------
def compute_flux_x(case: OracleCase, vals: np.ndarray, d3: np.ndarray) -> np.ndarray:
    """x-fluxes at the nx+1 interfaces of every interior row (interpolate.py:95-129).
    Note ``hv_coeff`` uses the FULL time step ``params["dt"]`` (interpolate.py:99-101)."""
    nz = case.nz
    hv = -HV_BETA * case.dx / (16 * case.dt)
    rho = vals[DENS] + case.hy_dens_cell[HS:nz + HS, np.newaxis]
    u = vals[UMOM] / rho
    w = vals[WMOM] / rho
    t = (vals[RHOT] + case.hy_dens_theta_cell[HS:nz + HS, np.newaxis]) / rho
    p = C0 * np.power(rho * t, GAMMA)
    flux = np.empty_like(vals)
    flux[DENS] = rho * u - hv * d3[DENS]
    flux[UMOM] = rho * u**2 + p - hv * d3[UMOM]
    flux[WMOM] = rho * u * w - hv * d3[WMOM]
    flux[RHOT] = rho * u * t - hv * d3[RHOT]
    return flux


def compute_flux_z(case: OracleCase, vals: np.ndarray, d3: np.ndarray) -> np.ndarray:
    """z-fluxes at the nz+1 interfaces (interpolate.py:144-186).  The wall rows
    k=0 and k=nz get w=0 and a zero density hyperviscosity term (:168-173)."""
    nz = case.nz
    hv = -HV_BETA * case.dz / (16 * case.dt)
    rho = vals[DENS] + case.hy_dens_int[:, np.newaxis]
    u = vals[UMOM] / rho
    w = vals[WMOM] / rho
    t = (vals[RHOT] + case.hy_dens_theta_int[:, np.newaxis]) / rho
    p = C0 * np.power(rho * t, GAMMA) - case.hy_pressure_int[:, np.newaxis]
    w[0, :] = 0.0
    w[nz, :] = 0.0
    d3 = d3.copy()
    d3[DENS, 0, :] = 0.0
    d3[DENS, nz, :] = 0.0
    flux = np.empty_like(vals)
    flux[DENS] = rho * w - hv * d3[DENS]
    flux[UMOM] = rho * w * u - hv * d3[UMOM]
    flux[WMOM] = rho * w**2 + p - hv * d3[WMOM]
    flux[RHOT] = rho * w * t - hv * d3[RHOT]
    return flux


# ---- tendencies -------------------------------------------------------------------
def compute_tend_x(case: OracleCase, flux: np.ndarray) -> np.ndarray:
    """-(F[i+1]-F[i])/dx, true division (interpolate.py:208-215)."""
    nx = case.nx
    return -(flux[:, :, 1:nx + 1] - flux[:, :, 0:nx]) / case.dx


def compute_tend_z(case: OracleCase, flux: np.ndarray, s: np.ndarray) -> np.ndarray:
    """-(F[k+1]-F[k])/dz and the hydrostatic source on WMOM (interpolate.py:238-250)."""
    nx, nz = case.nx, case.nz
    tend = -(flux[:, 1:nz + 1, :] - flux[:, 0:nz, :]) / case.dz
    tend[WMOM] -= s[DENS, HS:nz + HS, HS:nx + HS] * GRAV
    return tend


# ---- one RK stage / one time step ---------------------------------------------------
def tendency(case: OracleCase, forcing: np.ndarray, direction: int) -> np.ndarray:
    """BC fill on ``forcing`` (in place) then the tendency, step.py:67-76."""
    if direction == DIR_X:
        set_bc_x(case, forcing)
        vals, d3 = interpolate_x(case, forcing)
        return compute_tend_x(case, compute_flux_x(case, vals, d3))
    set_bc_z(case, forcing)
    vals, d3 = interpolate_z(case, forcing)
    return compute_tend_z(case, compute_flux_z(case, vals, d3), forcing)


def discrete_step(case: OracleCase, init: np.ndarray, forcing: np.ndarray,
                  out: np.ndarray, dt: float, direction: int) -> None:
    """step.py:21-82; ``out`` may alias ``init`` or ``forcing`` (the tendency is
    fully materialised before the update, exactly as in the reference)."""
    nx, nz = case.nx, case.nz
    tend = tendency(case, forcing, direction)
    if case.source_w is not None:  # add_source_terms, source.py:53-75
        tend[WMOM] += case.source_w
    out[:, HS:nz + HS, HS:nx + HS] = init[:, HS:nz + HS, HS:nx + HS] + dt * tend


def evolve(case: OracleCase, dt: float | None = None) -> None:
    """One full time step = two directional sweeps of three RK stages
    (step.py:85-143).  First call sweeps Z then X; the order alternates."""
    dt = case.dt if dt is None else dt
    dirs = (DIR_X, DIR_Z) if case.reverse_direction else (DIR_Z, DIR_X)
    for d in dirs:
        discrete_step(case, case.state, case.state, case.state_tmp, dt / 3, d)
        discrete_step(case, case.state, case.state_tmp, case.state_tmp, dt / 2, d)
        discrete_step(case, case.state, case.state_tmp, case.state, dt / 1, d)
    case.reverse_direction = not case.reverse_direction


def gravity_source(nx, nz, dx, dz, xlen, zlen, hy_dens_cell) -> np.ndarray:
    """The gravity-wave forcing field of source.py:41-50: a squared-cosine bump of amplitude 0.01
    centred at (xlen/8, 1000 m), radii 500 m, sampled at cell centres (mesh.py:55-77,
    utils/utils.py:40-51), times the hydrostatic density of the row."""
    x = np.linspace(dx / 2.0, xlen + dx / 2.0, nx, endpoint=False)
    z = np.linspace(dz / 2.0, zlen + dz / 2.0, nz, endpoint=False)
    x, z = np.meshgrid(x, z)
    pi = 3.14159265358979323846264338327
    dist = np.sqrt(((x - xlen / 8) / 500.0) ** 2 + ((z - 1000.0) / 500.0) ** 2) * pi / 2.0
    wpert = np.zeros(z.shape)
    np.putmask(wpert, dist <= pi / 2.0, 0.01 * (np.cos(dist) ** 2.0))
    return wpert * hy_dens_cell[HS:nz + HS, np.newaxis]


# ---- diagnostics ----------------------------------------------------------------------
def compute_stats(case: OracleCase, s: np.ndarray | None = None):
    """(total_mass, total_energy) over the interior (stats.py:16-33).  The kinetic
    term has no 1/2 -- that is what the reference computes."""
    nx, nz = case.nx, case.nz
    s = case.state if s is None else s
    inner = (slice(HS, nz + HS), slice(HS, nx + HS))
    rho = s[DENS][inner] + case.hy_dens_cell[HS:nz + HS, np.newaxis]
    u = s[UMOM][inner] / rho
    w = s[WMOM][inner] / rho
    th = (s[RHOT][inner] + case.hy_dens_theta_cell[HS:nz + HS, np.newaxis]) / rho
    p = C0 * np.power(rho * th, GAMMA)
    t = th / np.power(P0 / p, RD / CP)
    ke = rho * (u * u + w * w)
    ie = rho * CV * t
    return rho.sum() * case.dx * case.dz, (ke + ie).sum() * case.dx * case.dz


def compute_solution_variables(case: OracleCase, s: np.ndarray | None = None) -> np.ndarray:
    """Derived output variables rho', u, w, theta' (stats.py:38-69)."""
    nx, nz = case.nx, case.nz
    s = case.state if s is None else s
    inner = (slice(HS, nz + HS), slice(HS, nx + HS))
    hd = case.hy_dens_cell[HS:nz + HS, np.newaxis]
    hdt = case.hy_dens_theta_cell[HS:nz + HS, np.newaxis]
    out = np.zeros((4, nz, nx), dtype=np.float64)
    out[DENS] = s[DENS][inner]
    out[UMOM] = s[UMOM][inner] / (hd + s[DENS][inner])
    out[WMOM] = s[WMOM][inner] / (hd + s[DENS][inner])
    out[RHOT] = (s[RHOT][inner] + hdt) / (hd + s[DENS][inner]) - (hdt / hd)
    return out
