"""Stage time vs grid height: slope = streaming rate, intercept = fixed per-launch cost (development tool)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from helpers import new_case, HYDRO
from pyminiweather_b200.engine import DeviceSolver
sys.path.insert(0, "tools")
from sweep_tiles_lib import bench  # noqa


nx = 2048
rows = []
for nz in (128, 256, 512, 1024, 2048, 4096):
    p, case = new_case(nx, nz, "thermal")
    s = DeviceSolver(case.nx, case.nz, case.dx, case.dz, case.dt, variant="tma", pow_mode="background")
    s.set_hydrostatic(*[getattr(case, n) for n in HYDRO]); s.upload(0, case.state); s.upload(1, case.state_tmp)
    for tune in sys.argv[1:]:
        k, v = tune.split("="); s.set_tuning(**{k: int(v)})
    x = bench(s, 1); z = bench(s, 2)
    print(f"nz={nz:5d} cells={nx*nz:9d}  x us: {x[0]:7.1f} {x[1]:7.1f} {x[2]:7.1f}   z us: {z[0]:7.1f} {z[1]:7.1f} {z[2]:7.1f}", flush=True)
    rows.append((nx * nz, x, z))
    s.close()
cells = np.array([r[0] for r in rows], float)
for name, idx in (("x", 1), ("z", 2)):
    for rk in range(3):
        t = np.array([r[idx][rk] for r in rows])
        A = np.vstack([cells, np.ones_like(cells)]).T
        slope, icpt = np.linalg.lstsq(A, t, rcond=None)[0]
        bytes_per_cell = 64 if rk == 0 else 96
        print(f"{name} S{rk+1}: intercept {icpt:6.2f} us, slope -> {bytes_per_cell / slope / 1e3:7.1f} GB/s streaming")
