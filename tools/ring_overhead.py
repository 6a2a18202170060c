"""What the slab-ring protocol of the fused x sweep costs by itself, on ONE GPU: the periodic context (static item
walk, no push, no wait) against a ring of one slab mapped onto itself through the peer path (push CTAs, epoch
flags, edge tile columns last) with the static and the dynamic item walk.  NVLink latency is the only thing
missing compared with a real ring.   usage: python tools/ring_overhead.py [nx nz steps]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import new_case, HYDRO
from pyminiweather_b200.engine import DeviceSolver
nx, nz, steps = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (2048, 1024, 500)))
p, case = new_case(nx, nz, "thermal")


def run(periodic, **tune):
    s = DeviceSolver(nx, nz, case.dx, case.dz, case.dt, periodic_x=periodic)
    s.set_hydrostatic(*[getattr(case, k) for k in HYDRO]); s.set_tuning(**tune)
    s.upload(0, case.state); s.upload(1, case.state)
    if not periodic:
        mine = s.local_ptrs(); s.connect_peers(mine, mine)
    s.evolve(20); s.synchronize()
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); s.evolve(steps); s.synchronize(); best = min(best, time.perf_counter() - t0)
    out = s.download(0)[:, 2:-2, 2:-2].copy()
    s.close()
    return best / steps * 1e6, out


t_per, ref = run(True)
print(f"periodic context, static walk            {t_per:7.2f} us/step")
t_dyn2, o = run(True, dyn_items=2)
print(f"periodic context, dynamic walk           {t_dyn2:7.2f} us/step  (+{t_dyn2 - t_per:5.2f})  same bits: {np.array_equal(o, ref)}")
t_s, o = run(False, dyn_items=0)
print(f"self-ring (push + epoch wait), static    {t_s:7.2f} us/step  (+{t_s - t_per:5.2f})  same bits: {np.array_equal(o, ref)}")
t_d, o = run(False, dyn_items=1)
print(f"self-ring (push + epoch wait), dynamic   {t_d:7.2f} us/step  (+{t_d - t_per:5.2f})  same bits: {np.array_equal(o, ref)}")
