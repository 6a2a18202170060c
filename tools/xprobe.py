import sys, time, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from helpers import new_case, HYDRO
from pyminiweather_b200.engine import DeviceSolver
nx, nz = 2048, 1024
p, case = new_case(nx, nz, "thermal")
for tune in [dict(fuse=1, sweep_zt=0, peer_dbg=8), dict(fuse=1, sweep_zt=0)]:
    s = DeviceSolver(nx, nz, case.dx, case.dz, case.dt)
    s.set_hydrostatic(*[getattr(case, n) for n in HYDRO]); s.set_tuning(**tune)
    s.upload(0, case.state); s.upload(1, case.state)
    s.evolve(20); s.synchronize()
    t0 = time.perf_counter(); s.evolve(400); s.synchronize(); dt = time.perf_counter() - t0
    print(os.environ.get("PMW_LIB", "default")[-16:], tune, f"{dt/400*1e6:7.1f} us/step", flush=True)
    s.close()
