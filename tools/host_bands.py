"""The streamed host step (pmw_evolve_host) against the plain sequence upload + evolve(1) + download, on pinned host
memory, for a list of band counts.   usage: python tools/host_bands.py nx nz [bands ...] [key=value tuning ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from helpers import new_case, HYDRO
from pyminiweather_b200.engine import DeviceSolver

nx, nz = int(sys.argv[1]), int(sys.argv[2])
tune = {k: int(v) for k, v in (a.split("=") for a in sys.argv[3:] if "=" in a)}
bands = [int(b) for b in sys.argv[3:] if "=" not in b] or [1, 2, 4, 8, 16, 32, 0]
_, case = new_case(nx, nz, "thermal")
pinned = torch.empty((4, nz + 4, nx + 4), dtype=torch.float64, pin_memory=True)
host = pinned.numpy()
s = DeviceSolver(nx, nz, case.dx, case.dz, case.dt)
s.set_hydrostatic(*[getattr(case, k) for k in HYDRO])
if tune:
    s.set_tuning(**tune)
nbytes = host.nbytes


def timeit(call, n=20):
    for _ in range(3):
        call()
    t0 = time.perf_counter()
    for _ in range(n):
        call()
    return (time.perf_counter() - t0) / n


def plain():
    s.upload(0, host); s.evolve(1); s.download(0, out=host)


host[:] = case.state
ref = host.copy()
t = timeit(plain)
print(f"plain sequence      {t*1e3:7.3f} ms/step   {nx*nz/t:.3e} cells/s   {2*nbytes/t/1e9:6.1f} GB/s both ways", flush=True)
for b in bands:
    host[:] = case.state
    t = timeit(lambda: s.evolve_host(host, None, b))
    print(f"evolve_host bands={b:3d} {t*1e3:7.3f} ms/step   {nx*nz/t:.3e} cells/s   {2*nbytes/t/1e9:6.1f} GB/s both ways", flush=True)
s.close()
