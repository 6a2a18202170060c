"""Runs a few fused-sweep steps (target of the ncu captures).  usage: sweep_one.py nx nz steps [key=value ...]
selfring=1: a ring of ONE slab mapped onto itself through the peer path -- the DYNAMIC instantiation of sweep_x with its
halo push, epoch flags and edge-row patch, on a single GPU (ncu cannot attach to a rank of a real ring)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import new_case, HYDRO
from pyminiweather_b200.engine import DeviceSolver
nx, nz, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
tune = {k: int(v) for k, v in (kv.split("=") for kv in sys.argv[4:])}
selfring = bool(tune.pop("selfring", 0))
p, case = new_case(nx, nz, "thermal")
s = DeviceSolver(nx, nz, case.dx, case.dz, case.dt, periodic_x=not selfring)
s.set_hydrostatic(*[getattr(case, n) for n in HYDRO]); s.set_tuning(**tune)
s.upload(0, case.state); s.upload(1, case.state)
if selfring:
    mine = s.local_ptrs(); s.connect_peers(mine, mine)
s.evolve(steps); s.synchronize()
print("lz", s.get_tuning("sweep_lz"), "launches", s.launch_count)
s.close()
