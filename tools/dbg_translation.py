import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import synthetic_case, HYDRO, interior
from pyminiweather_b200.engine import DeviceSolver
def mk(case, variant):
    s = DeviceSolver(case.nx, case.nz, case.dx, case.dz, case.dt, variant=variant, pow_mode="background")
    s.set_hydrostatic(*[getattr(case, n) for n in HYDRO]); s.upload(0, case.state); s.upload(1, case.state_tmp); return s
p, case = synthetic_case(2048, 256, seed=11); m = 333
sh = case.copy(); sh.state[:, :, 2:-2] = np.roll(case.state[:, :, 2:-2], m, axis=2); sh.state_tmp[:] = sh.state
for variant in ("direct", "tma"):
    for what in ("x1", "z1", "x123", "z123", "evolve1", "evolve3"):
        a, b = mk(case, variant), mk(sh, variant)
        for s in (a, b):
            if what == "x1": s.discrete_step(1, 0, 0, 1, case.dt/3); buf = 1
            elif what == "z1": s.discrete_step(2, 0, 0, 1, case.dt/3); buf = 1
            elif what in ("x123", "z123"):
                d = 1 if what[0] == "x" else 2
                s.discrete_step(d, 0, 0, 1, case.dt/3); s.discrete_step(d, 0, 1, 1, case.dt/2); s.discrete_step(d, 0, 1, 0, case.dt); buf = 0
            else: s.evolve(int(what[-1])); buf = 0
        ra, rb = np.roll(interior(a.download(buf)), m, axis=2), interior(b.download(buf))
        bad = ra != rb
        cols = np.unique(np.nonzero(bad)[2])
        print(variant, what, "mismatches", bad.sum(), "of", bad.size, "maxabs %.3e" % np.abs(ra-rb).max(), "cols", cols[:12], "... n=", cols.size)
        a.close(); b.close()
