"""Fused sweeps against the stage-by-stage path, bit for bit, on a list of grids; prints where they differ.
usage: python tools/fused_check.py nx,nz[,lz] ... [key=value ...]   (run on a GPU; PMW_LIB selects the library;
key=value pairs are tuning switches of the fused solver, e.g. sweep_z3=1)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import synthetic_case, HYDRO
from pyminiweather_b200.engine import DeviceSolver


def solver_for(case, variant="tma", pow_mode="background", **tuning):
    s = DeviceSolver(case.nx, case.nz, case.dx, case.dz, case.dt, variant=variant, pow_mode=pow_mode)
    s.set_hydrostatic(*[getattr(case, n) for n in HYDRO])
    if tuning:
        s.set_tuning(**tuning)
    s.upload(0, case.state); s.upload(1, case.state_tmp)
    return s


bad = 0
extra = {k: int(v) for k, v in (a.split("=") for a in sys.argv[1:] if "=" in a)}
for spec in [a for a in sys.argv[1:] if "=" not in a]:
    v = [int(x) for x in spec.split(",")]
    nx, nz = v[:2]
    tune = dict(extra, sweep_lz=v[2]) if len(v) > 2 else dict(extra)
    p, case = synthetic_case(nx, nz, seed=nx + nz)
    a, b = solver_for(case, "tma", "background", fuse=0), solver_for(case, "tma", "background", fuse=1, **tune)
    for n in (1, 4):
        a.evolve(n); b.evolve(n)
        for which in (0, 1):
            ra, rb = a.download(which), b.download(which)
            sl = (slice(None), slice(2, -2), slice(None) if which == 0 else slice(2, -2))
            d = ra[sl] != rb[sl]
            if d.any():
                bad += 1
                vv, kk, ii = np.nonzero(d)
                print(f"{spec}: after {n} calls buf {which}: {d.sum()} cells differ; vars {sorted(set(vv))} rows {sorted(set(kk))[:20]} "
                      f"cols {sorted(set(ii))[:12]}..{max(ii)}  max |d| {np.abs(ra[sl] - rb[sl]).max():.3e}")
            else:
                print(f"{spec}: after {n} calls buf {which}: identical (lz {b.get_tuning('sweep_lz')})")
    a.close(); b.close()
sys.exit(1 if bad else 0)
