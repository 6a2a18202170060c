"""Time of one compute_stats on the device (stats_partial_kernel + stats_final_kernel, no host copy) for library
variants.   usage: python tools/stats_time.py nx nz tag [tag ...]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if sys.argv[1] == "--child":
    import time
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    from helpers import new_case, HYDRO
    from pyminiweather_b200.engine import DeviceSolver
    nx, nz = int(sys.argv[2]), int(sys.argv[3])
    _, case = new_case(nx, nz, "thermal")
    s = DeviceSolver(nx, nz, case.dx, case.dz, case.dt)
    s.set_hydrostatic(*[getattr(case, k) for k in HYDRO]); s.upload(0, case.state)
    s.evolve(50)
    out = torch.zeros(2, dtype=torch.float64, device="cuda")
    vals = s.stats(0)
    best = 1e9
    for _ in range(5):
        s.synchronize(); t0 = time.perf_counter()
        for _ in range(200):
            s.stats_device(0, out.data_ptr())
        s.synchronize(); best = min(best, (time.perf_counter() - t0) / 200)
    print(f"{best*1e6:7.2f} us per compute_stats   {4*8*nx*nz/best/1e9:7.1f} GB/s   mass {vals[0]:.15e} energy {vals[1]:.15e}")
    sys.exit(0)
nx, nz = sys.argv[1:3]
for tag in sys.argv[3:]:
    lib = os.path.join(ROOT, "pyminiweather_b200", "variants", f"libpmw_{tag}.so")
    r = subprocess.run([sys.executable, __file__, "--child", nx, nz], env=dict(os.environ, PMW_LIB=lib), capture_output=True, text=True)
    print(f"{tag:10s} {r.stdout.strip() or r.stderr.strip()[-400:]}", flush=True)
