"""Times pmw_evolve with one kernel per RK stage (fuse=0) against one kernel per sweep (fuse=1)
for a few segment heights / tile widths.  usage: python tools/sweep_probe.py [nx nz [steps]]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import new_case, HYDRO
from pyminiweather_b200.engine import DeviceSolver
nx, nz = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 1024)
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 400
p, case = new_case(nx, nz, "thermal")
def run(**tune):
    s = DeviceSolver(nx, nz, case.dx, case.dz, case.dt)
    s.set_hydrostatic(*[getattr(case, n) for n in HYDRO]); s.set_tuning(**tune)
    s.upload(0, case.state); s.upload(1, case.state)
    s.evolve(20); ref = s.download(0); s.synchronize()
    t0 = time.perf_counter(); s.evolve(steps); s.synchronize(); dt = time.perf_counter() - t0
    s.close(); return dt / steps * 1e6, ref
base = None
cfgs = [dict(fuse=0), dict(fuse=1), dict(fuse=1, peer_dbg=8), dict(fuse=1, peer_dbg=16)]  # 8: x sweeps only, 16: z sweeps only
cfgs += [dict(fuse=1, sweep_zt=0), dict(fuse=1, sweep_zt=0, peer_dbg=16), dict(fuse=1, pdl=0)]
for tune in cfgs:
    try:
        us, st = run(**tune)
    except Exception as e:
        print(tune, "FAILED", e, flush=True); continue
    if base is None: base = st
    print(tune, f"{us:7.1f} us/step  {nx*nz/us*1e6:.3e} cells/s  bitwise-equal-to-staged: {np.array_equal(st[:, 2:-2, :], base[:, 2:-2, :])}", flush=True)
