"""A/B timing of libpmw variants (tools/build_variant.py): for every variant, in a fresh process, the
fused step at nx x nz -- whole, x sweeps only, z sweeps only -- and whether the result still has the
bits of the stage-by-stage path.   usage: python tools/ab_sweeps.py nx nz steps tag [tag ...] [ic=collision] [key=value ...]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import time
    import numpy as np
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import new_case, HYDRO
    from pyminiweather_b200.engine import DeviceSolver
    nx, nz, steps = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    extra = dict(kv.split("=") for kv in sys.argv[5:])
    ic = extra.pop("ic", "thermal")  # ic=collision etc.: the initial condition (default thermal)
    extra = {k: int(v) for k, v in extra.items()}
    p, case = new_case(nx, nz, ic)
    def run(n, **tune):
        s = DeviceSolver(nx, nz, case.dx, case.dz, case.dt)
        s.set_hydrostatic(*[getattr(case, k) for k in HYDRO]); s.set_tuning(**tune)
        s.upload(0, case.state); s.upload(1, case.state)
        s.evolve(7); ref = s.download(0); s.synchronize()
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter(); s.evolve(n); s.synchronize(); best = min(best, time.perf_counter() - t0)
        s.close(); return best / n * 1e6, ref
    _, staged = run(10, fuse=0)
    full, st = run(steps, fuse=1, **extra)
    only_z, _ = run(steps, fuse=1, peer_dbg=16, **extra)  # peer_dbg bit 16 skips the x sweeps, bit 8 the z sweeps
    only_x, _ = run(steps, fuse=1, peer_dbg=8, **extra)
    print(f"step {full:7.2f} us   only-x {only_x:7.2f}   only-z {only_z:7.2f}   {nx*nz/full*1e6:.4e} cells/s   "
          f"bitwise==staged: {np.array_equal(st[:, 2:-2, :], staged[:, 2:-2, :])}", flush=True)
    sys.exit(0)
nx, nz, steps = sys.argv[1:4]
tags = [t for t in sys.argv[4:] if "=" not in t]
extra = [t for t in sys.argv[4:] if "=" in t]
for tag in tags:
    lib = os.path.join(ROOT, "pyminiweather_b200", "variants", f"libpmw_{tag}.so")
    env = dict(os.environ, PMW_LIB=lib)
    r = subprocess.run([sys.executable, __file__, "--child", nx, nz, steps] + extra, env=env, capture_output=True, text=True)
    print(f"{tag:10s} {r.stdout.strip() or r.stderr.strip()[-400:]}", flush=True)
