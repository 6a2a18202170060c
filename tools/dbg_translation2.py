import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import synthetic_case, HYDRO, interior
from pyminiweather_b200.engine import DeviceSolver
def mk(case, variant, **t):
    s = DeviceSolver(case.nx, case.nz, case.dx, case.dz, case.dt, variant=variant, pow_mode="background")
    s.set_hydrostatic(*[getattr(case, n) for n in HYDRO]); s.upload(0, case.state); s.upload(1, case.state_tmp)
    if t: s.set_tuning(**t)
    return s
p, case = synthetic_case(2048, 64, seed=11)
d = mk(case, "direct"); d.discrete_step(1, 0, 0, 1, case.dt/3); rd = interior(d.download(1))
for xp in (2, 4, 5):
    t = mk(case, "tma", x_p=xp); t.discrete_step(1, 0, 0, 1, case.dt/3); rt = interior(t.download(1))
    bad = rd != rt
    cols = np.unique(np.nonzero(bad)[2]); rows = np.unique(np.nonzero(bad)[1]); vs = np.unique(np.nonzero(bad)[0])
    print("x_p", xp, "TC", 32*xp-1, "mismatch", bad.sum(), "vars", vs, "nrows", rows.size, "cols(first 40)", cols[:40], "ncols", cols.size)
    print("  cols mod TC histogram (top):", np.bincount(cols % (32*xp-1), minlength=32*xp-1).nonzero()[0][:60])
    if bad.sum():
        v,k,i = [a[0] for a in np.nonzero(bad)]
        print("  example", v,k,i, repr(rd[v,k,i]), repr(rt[v,k,i]), "rel", abs(rd[v,k,i]-rt[v,k,i])/abs(rd[v,k,i]))
