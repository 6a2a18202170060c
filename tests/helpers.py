"""Shared test helpers (tests may import the oracle; the product may not)."""
import numpy as np

from oracle import numpy_oracle as no

HYDRO = ("hy_dens_cell", "hy_dens_theta_cell", "hy_dens_int", "hy_dens_theta_int", "hy_pressure_int")


def make_params(nx, nz, ic_type="thermal", xlen=2e4, zlen=1e4, dt=None, cfl=1.0, max_speed=500.0):
    """The params dict of pyminiweather/__main__.py:160-195."""
    p = dict(nx=nx, nz=nz, xlen=xlen, zlen=zlen, dt=dt, nsteps=0, nwarmups=0, ic_type=ic_type, hs=2, s=4,
             max_speed=max_speed, cfl=cfl, output_freq=-1, app_filename="PyMiniWeatherData.txt",
             app_log_file=None, verbose=False)
    p["dx"] = p["xlen"] / nx
    p["dz"] = p["zlen"] / nz
    if p["dt"] is None:
        p["dt"] = np.minimum(p["dx"], p["dz"]) * cfl / max_speed
    return p


def params_from_golden(g, ic_type="thermal"):
    nx, nz = int(g["nx"]), int(g["nz"])
    p = make_params(nx, nz, ic_type)
    p["dx"], p["dz"], p["dt"] = float(g["dx"]), float(g["dz"]), float(g["dt"])
    return p


def case_from_arrays(p, state, state_tmp, hydro):
    return no.OracleCase(p["nx"], p["nz"], float(p["dx"]), float(p["dz"]), float(p["dt"]),
                         np.array(state, dtype=np.float64, copy=True), np.array(state_tmp, dtype=np.float64, copy=True),
                         *[np.array(hydro[n], dtype=np.float64, copy=True) for n in HYDRO])


def case_from_golden(g, state_key="state0", tmp_key=None, ic_type="thermal"):
    p = params_from_golden(g, ic_type)
    st = g[state_key]
    tmp = g[tmp_key] if tmp_key else st
    case = case_from_arrays(p, st, tmp, g)
    if ic_type == "injection":
        p["zlen"] = float(g["zlen"])
        case.inflow_zlen = p["zlen"]
    return p, case


def new_case(nx, nz, ic_type="thermal", **kw):
    """Initial condition from OUR init (bit-identical to the reference's: tests/test_reference_style.py against the ic_*_32x16 fixtures)."""
    from pyminiweather_b200.data import initialize_fields
    from pyminiweather_b200.ics import init
    from pyminiweather_b200.mesh import MeshData
    p = make_params(nx, nz, ic_type, **kw)
    f = initialize_fields(p)
    init(f, p, MeshData(p))
    hydro = {n: getattr(f, n) for n in HYDRO}
    case = case_from_arrays(p, f._host[0], f._host[1], hydro)
    if ic_type == "gravity":
        case.source_w = no.gravity_source(nx, nz, case.dx, case.dz, p["xlen"], p["zlen"], case.hy_dens_cell)
    if ic_type == "injection":
        case.inflow_zlen = float(p["zlen"])
    return p, case


def synthetic_case(nx, nz, seed=20260101):
    """BASELINE config 5 style input: thermal background + uniform random perturbation
    (SURVEY.md section 8d)."""
    p, case = new_case(8, nz, "thermal")  # 1-D profiles only depend on nz
    p = make_params(nx, nz, "thermal")
    rng = np.random.default_rng(seed)
    amp = np.array([1e-3, 1e-1, 1e-1, 1e-1])[:, None, None]
    st = np.zeros((4, nz + 4, nx + 4))
    st[:, 2:-2, 2:-2] = amp * rng.uniform(-1.0, 1.0, size=(4, nz, nx))
    hydro = {n: getattr(case, n) for n in HYDRO}
    return p, case_from_arrays(p, st, st, hydro)


def interior(a):
    return a[..., 2:-2, 2:-2]


def rel_l2(a, b):
    n = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (n if n > 0 else 1.0)


def per_variable_rel_l2(got, want):
    """Unfloored relative L2 error of every variable on the interior (reported next to the parity
    metric so that the floor in ``worst_rel_l2`` stays visible)."""
    gi, wi = interior(got), interior(want)
    return [float(np.linalg.norm(gi[v] - wi[v]) / max(np.linalg.norm(wi[v]), 1e-300)) for v in range(4)]


def worst_rel_l2(got, want):
    """The parity metric: relative L2 error on the interior, max over the stacked 4-field state
    (the north-star criterion) and over each variable.  A variable whose own norm is below 1e-3 of
    the stacked norm is normalised by that floor instead: its relative error is ill-conditioned --
    e.g. rho' early in a thermal run has norm 1e-4 against 4e2 for (rho*theta)', and two faithful
    CPU evaluations of the reference that differ only in 1-ulp pow() rounding (NumPy SIMD pow vs
    libm) already differ by 1.3e-10 in that variable at 2048x1024 (stacked: 2.7e-14).  See
    DESIGN.md, "Parity metric"."""
    gi, wi = interior(got), interior(want)
    total = np.linalg.norm(wi)
    errs = [np.linalg.norm(gi - wi) / (total if total > 0 else 1.0)]
    for v in range(4):
        errs.append(np.linalg.norm(gi[v] - wi[v]) / max(np.linalg.norm(wi[v]), 1e-3 * total, 1e-300))
    return max(errs)
