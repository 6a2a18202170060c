"""Single-GPU probe of the peer-ring overheads: ring of one slab mapped onto itself (development tool)."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
from helpers import new_case, HYDRO
from pyminiweather_b200.engine import DeviceSolver
p, case = new_case(2048, 1024, "thermal")
def run(periodic, dbg=0, pdl=1, steps=300):
    s = DeviceSolver(case.nx, case.nz, case.dx, case.dz, case.dt, periodic_x=periodic)
    s.set_hydrostatic(*[getattr(case, n) for n in HYDRO]); s.upload(0, case.state); s.upload(1, case.state_tmp)
    s.set_tuning(pdl=pdl, peer_dbg=dbg)
    if not periodic:
        mine = s.local_ptrs(); s.connect_peers(mine, mine)
    s.evolve(30); s.synchronize()
    t0 = time.perf_counter(); s.evolve(steps); s.synchronize(); dt = time.perf_counter() - t0
    s.close()
    return dt / steps * 1e6
for name, kw in [("periodic", dict(periodic=True)), ("self-peer", dict(periodic=False)),
                 ("self-peer nofence", dict(periodic=False, dbg=1)), ("self-peer nowait", dict(periodic=False, dbg=2)),
                 ("self-peer nofence nowait", dict(periodic=False, dbg=3)), ("periodic nopdl", dict(periodic=True, pdl=0)),
                 ("self-peer nopdl", dict(periodic=False, pdl=0))]:
    print(f"{name:28s} {run(**kw):8.1f} us/step", flush=True)
