import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import new_case, HYDRO
from pyminiweather_b200.engine import DeviceSolver
p, case = new_case(100, 50, "thermal")
which = sys.argv[1]
s = DeviceSolver(case.nx, case.nz, case.dx, case.dz, case.dt, variant="tma", pow_mode="libdevice")
s.set_hydrostatic(*[getattr(case, n) for n in HYDRO])
s.upload(0, case.state); s.upload(1, case.state_tmp)
if which == "x":
    s.bc_x(0); s.stage(1, 0, 0, 1, case.dt/3)
else:
    s.stage(2, 0, 0, 1, case.dt/3)
s.synchronize()
print("ok", which, np.abs(s.download(1)).max())
