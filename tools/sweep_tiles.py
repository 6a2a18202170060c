"""Time the stage kernels per direction / RK stage for every tile shape (development tool).
Usage (GPU box): python tools/sweep_tiles.py [nx nz]"""
import sys, itertools, json
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from helpers import new_case, HYDRO
from pyminiweather_b200.engine import DeviceSolver

nx, nz = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 1024)
p, case = new_case(nx, nz, "thermal")
PEAK = 6532.5

def bench(s, d, reps=60):
    # one sweep = stages 1,2,3 in direction d; time each RK stage kind separately
    out = {}
    for rk in (1, 2, 3):
        for _ in range(5):
            for r in (1, 2, 3): s.evolve_stage(d, r)
        s.synchronize()
    # per-stage timing with events: run full sweeps, timing on; stage kinds interleave -> use 3 separate passes
    res = []
    for rk in (1, 2, 3):
        tot = 0.0; n = 0
        for _ in range(reps):
            for r in (1, 2, 3):
                if r == rk: s.stage_timing(True)
                s.evolve_stage(d, r)
                if r == rk:
                    ms, k = s.stage_timing_read(); s.stage_timing(False); tot += ms * k; n += k
        res.append(tot / n * 1e3)
    return res  # us per launch for rk=1,2,3

def report(tag, us):
    cells = nx * nz
    fr = [cells * b / (t * 1e-6) / 1e9 / PEAK for b, t in zip((64, 96, 96), us)]
    print(f"{tag:28s} us: {us[0]:7.1f} {us[1]:7.1f} {us[2]:7.1f}   frac: {fr[0]:.3f} {fr[1]:.3f} {fr[2]:.3f}   sum {sum(us):7.1f}", flush=True)

def mk(variant="tma"):
    s = DeviceSolver(case.nx, case.nz, case.dx, case.dz, case.dt, variant=variant, pow_mode="background")
    s.set_hydrostatic(*[getattr(case, n) for n in HYDRO]); s.upload(0, case.state); s.upload(1, case.state_tmp)
    return s

s = mk("direct")
report("direct x", bench(s, 1)); report("direct z", bench(s, 2)); s.close()
s = mk()
extra = dict(kv.split("=") for kv in sys.argv[3:])
if extra: s.set_tuning(**{k: int(v) for k, v in extra.items()})
for tr, xp in itertools.product((4, 8), (1, 2, 3)):
    s.set_tuning(x_tr=tr, x_p=xp)
    report(f"x tr={tr} p={xp}", bench(s, 1))
for z in range(1, 9):
    s.set_tuning(z_cfg=z)
    report(f"z cfg={z}", bench(s, 2))
s.close()
