"""Measured parity errors of the CUDA path, printed as a table: for every case the north-star metric (relative L2
of the stacked 4-field state, interior), the UNFLOORED relative L2 of each variable (tests/helpers.py floors a
variable's norm at 1e-3 of the stacked norm; this keeps that relaxation visible), and the relative error of the
mass / energy totals.  Run on a GPU:  python tools/parity_report.py > profiles/rXX_parity_report.txt"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import golden
from helpers import HYDRO, case_from_golden, interior, new_case, per_variable_rel_l2, synthetic_case, worst_rel_l2
from oracle import c_oracle, numpy_oracle as no
from pyminiweather_b200.engine import DeviceSolver


def solver_for(case, **tune):
    s = DeviceSolver(case.nx, case.nz, case.dx, case.dz, case.dt)
    s.set_hydrostatic(*[getattr(case, n) for n in HYDRO])
    if tune:
        s.set_tuning(**tune)
    s.upload(0, case.state); s.upload(1, case.state_tmp)
    return s


def line(name, got, want, stats=None, want_stats=None):
    gi, wi = interior(got), interior(want)
    stacked = np.linalg.norm(gi - wi) / np.linalg.norm(wi)
    pv = per_variable_rel_l2(got, want)
    st = ""
    if stats is not None:
        st = f"  mass {abs(stats[0] - want_stats[0]) / abs(want_stats[0]):.1e}  energy {abs(stats[1] - want_stats[1]) / abs(want_stats[1]):.1e}"
    print(f"{name:58s} stacked {stacked:.2e}  floored-max {worst_rel_l2(got, want):.2e}  "
          f"unfloored rho' {pv[0]:.2e} rho*u {pv[1]:.2e} rho*w {pv[2]:.2e} rho*theta' {pv[3]:.2e}{st}", flush=True)


print("# reference fixtures (tests/golden, generated from the reference's NumPy backend)")
for ic, fn, steps in (("thermal", "evolve_thermal_100x50.npz", (1, 2, 10, 100, 1000)),
                      ("collision", "evolve_collision_100x50.npz", (100,)),
                      ("density-current", "evolve_density-current_100x50.npz", (100,))):
    g = golden(fn)
    p, case = case_from_golden(g, "state0", ic_type=ic)
    s = solver_for(case)
    done = 0
    for n in steps:
        s.evolve(n - done); done = n
        line(f"{ic} 100x50, {n} steps vs reference", s.download(0), g[f"state_{n}"], s.stats(0), g[f"stats_{n}"])
    s.close()
g = golden("evolve_thermal_2048x1024_10steps_sub32.npz")
p, case = new_case(2048, 1024, "thermal")
s = solver_for(case)
done = 0
for n in (1, 2, 5, 10):
    s.evolve(n - done); done = n
    got = interior(s.download(0))[:, ::32, ::32]
    want = g[f"sub_{n}"]
    pad = lambda a: np.pad(a, ((0, 0), (2, 2), (2, 2)))  # noqa: E731  (line() looks at the interior)
    line(f"thermal 2048x1024 (config 2), {n} steps vs reference, 1/32 sub-sample", pad(got), pad(want), s.stats(0), g[f"stats_{n}"])
s.close()
print("# C/OpenMP oracle (oracle/c), random-perturbation state of SURVEY.md 8d, 2 steps")
c_oracle.set_threads(len(os.sched_getaffinity(0)))
for nx, nz in ((2048, 1024), (1024, 2048), (2048, 4096), (4096, 8192)):
    p, case = synthetic_case(nx, nz, seed=nx)
    s = solver_for(case)
    c = c_oracle.COracle(case)
    c.evolve(2); s.evolve(2)
    line(f"synthetic {nx}x{nz}, 2 steps vs C oracle", s.download(0), case.state, s.stats(0), c.stats())
    s.close()
