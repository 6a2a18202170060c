"""Effect of L2 eviction hints on the step time at 2048x1024 (development tool)."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import new_case, HYDRO
from pyminiweather_b200.engine import DeviceSolver
p, case = new_case(2048, 1024, "thermal")
def run(h, steps=400):
    s = DeviceSolver(case.nx, case.nz, case.dx, case.dz, case.dt)
    s.set_hydrostatic(*[getattr(case, n) for n in HYDRO]); s.upload(0, case.state); s.upload(1, case.state_tmp)
    s.set_tuning(l2_hints=h)
    s.evolve(40); s.synchronize()
    t0 = time.perf_counter(); s.evolve(steps); s.synchronize(); dt = time.perf_counter() - t0
    s.close(); return dt / steps * 1e6
for h in [0, 1100, 1110, 1101, 1111, 1120, 1121, 1112, 1100, 0, 1100]:
    print(f"l2_hints={h:04d}  {run(h):7.1f} us/step", flush=True)
