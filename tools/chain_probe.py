"""Step time with and without tile-level chaining / PDL (development tool)."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import new_case, HYDRO
from pyminiweather_b200.engine import DeviceSolver
nx, nz = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 1024)
p, case = new_case(nx, nz, "thermal")
def run(steps=400, **tune):
    s = DeviceSolver(case.nx, case.nz, case.dx, case.dz, case.dt)
    s.set_hydrostatic(*[getattr(case, n) for n in HYDRO]); s.upload(0, case.state); s.upload(1, case.state_tmp)
    s.set_tuning(**tune)
    s.evolve(40); s.synchronize()
    t0 = time.perf_counter(); s.evolve(steps); s.synchronize(); dt = time.perf_counter() - t0
    st = s.stats(0); s.close(); return dt / steps * 1e6, st
for tune in [dict(chain=1, pdl=1), dict(chain=0, pdl=1), dict(chain=0, pdl=0), dict(chain=1, pdl=1)]:
    us, st = run(**tune)
    print(tune, f"{us:7.1f} us/step  {nx*nz/us*1e6:.3e} cells/s", st, flush=True)
