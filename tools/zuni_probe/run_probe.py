"""CPU check of the universal z iteration (csrc/pmw_zuni.cuh): compile its control flow for the host
(probe.cpp) and compare one fused z sweep -- every segment height, walls, ragged last segment, with and
without the gravity-wave forcing -- with three z stages of the NumPy oracle."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
root = os.path.abspath(os.path.join(here, "../.."))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
from helpers import new_case  # noqa: E402
from oracle import numpy_oracle as no  # noqa: E402

so = os.path.join(here, "probe.so")
subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                       os.path.join(here, "probe.cpp"), "-o", so])
lib = C.CDLL(so)
dp = C.POINTER(C.c_double)
P = lambda a: a.ctypes.data_as(dp)  # noqa: E731
worst = 0.0
for ic, nx, nz, steps in (("collision", 12, 50, 6), ("thermal", 8, 37, 4), ("gravity", 10, 64, 3), ("density-current", 6, 24, 5)):
    p, case = new_case(nx, nz, ic)
    for _ in range(steps):
        no.evolve(case)          # a state with motion everywhere
    for lz in sorted({8, 13, nz // 2, nz - 1, nz}):
        ref = case.copy()
        for dt_s, (i_, f_, o_) in zip((case.dt / 3, case.dt / 2, case.dt),
                                      (("state", "state", "state_tmp"), ("state", "state_tmp", "state_tmp"),
                                       ("state", "state_tmp", "state"))):
            if dt_s == case.dt:
                t2 = ref.state_tmp.copy()   # T2 = the stage-2 array
            no.discrete_step(ref, getattr(ref, i_), getattr(ref, f_), getattr(ref, o_), dt_s, no.DIR_Z)
        S = np.ascontiguousarray(case.state)
        O, T = np.full_like(S, np.nan), np.full_like(S, np.nan)
        src = case.source_w
        lib.zuni_sweep(nx, nz, lz, C.c_double(case.dz), C.c_double(case.dt), C.c_double(case.dt), P(S), P(O), P(T),
                       P(case.hy_dens_cell), P(case.hy_dens_int), P(case.hy_dens_theta_int), P(case.hy_pressure_int),
                       P(np.ascontiguousarray(src)) if src is not None else None)
        gi, wi = O[:, 2:-2, 2:-2], ref.state[:, 2:-2, 2:-2]
        ti, tw = T[:, 2:-2, 2:-2], t2[:, 2:-2, 2:-2]
        assert np.isfinite(gi).all() and np.isfinite(ti).all(), (ic, lz, "cells not written")
        err = max(np.abs(gi - wi).max() / np.abs(wi).max(), np.abs(ti - tw).max() / np.abs(tw).max())
        worst = max(worst, err)
        print(f"{ic:16s} {nx}x{nz} lz={lz:3d}: max relative error {err:.2e}")
        assert err <= 1e-11, (ic, lz)
print("universal z iteration agrees with the oracle; worst", worst)
