"""Chunked sweeps: correctness (bitwise vs unchunked) and step time (development tool)."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from helpers import new_case, HYDRO
from pyminiweather_b200.engine import DeviceSolver
nx, nz = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 1024)
p, case = new_case(nx, nz, "thermal")
def run(steps=400, **tune):
    s = DeviceSolver(case.nx, case.nz, case.dx, case.dz, case.dt)
    s.set_hydrostatic(*[getattr(case, n) for n in HYDRO]); s.upload(0, case.state); s.upload(1, case.state_tmp)
    s.set_tuning(**tune)
    s.evolve(41); s.synchronize()
    ref = s.download(0)
    t0 = time.perf_counter(); s.evolve(steps); s.synchronize(); dt = time.perf_counter() - t0
    s.close(); return dt / steps * 1e6, ref
base = None
for tune in [dict(chunks=1), dict(chunks=2), dict(chunks=2, pdl=0)]:
    us, st = run(**tune)
    if base is None: base = st
    print(tune, f"{us:7.1f} us/step  {nx*nz/us*1e6:.3e} cells/s  bitwise-equal-to-unchunked: {np.array_equal(st[:, 2:-2, 2:-2], base[:, 2:-2, 2:-2])}", flush=True)
