"""Measure how GPU-like arithmetic choices move the solution away from the NumPy
oracle (development tool).  Usage: PYTHONPATH=. python tools/arith_probe/run_probe.py"""
import ctypes as C, os, subprocess, sys
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "../..")))
from oracle import numpy_oracle as no, c_oracle as co, reference_runner as rr

here = os.path.dirname(os.path.abspath(__file__))
variants = {
    "fma": ["-ffp-contract=fast", "-mfma"],
    "fma+recip": ["-ffp-contract=fast", "-mfma", "-DVAR_RECIP"],
    "fma+recip+fastpow": ["-ffp-contract=fast", "-mfma", "-DVAR_FASTPOW"],
    "nofma+fastpow": ["-ffp-contract=off", "-mfma", "-DVAR_FASTPOW"],
}
def rel(a, b): return np.linalg.norm(a - b) / np.linalg.norm(a)
def build(name, flags):
    so = f"/tmp/probe_{name.replace('+','_')}.so"
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-fopenmp", "-std=gnu11", "-shared", *flags,
                           os.path.join(here, "probe.c"), "-o", so, "-lm"])
    return so
sos = {k: build(k, v) for k, v in variants.items()}
steps = [1, 10, 100, 1000]
for ic in ["thermal", "collision", "density-current"]:
    ref = rr.ReferenceRun(100, 50, ic)
    base = ref.to_oracle_case()
    a = base.copy(); snaps = {}; done = 0
    for n in steps:
        for _ in range(n - done): no.evolve(a)
        done = n; snaps[n] = (a.state.copy(), no.compute_stats(a))
    for name, so in sos.items():
        co._lib = None; co._SO = so
        b = base.copy(); cb = co.COracle(b); done = 0; out = []
        for n in steps:
            cb.evolve(n - done); done = n
            sa, st = snaps[n]
            pv = max(rel(sa[v][2:-2, 2:-2], b.state[v][2:-2, 2:-2]) for v in range(4))
            sb = cb.stats()
            out.append("n=%d all=%.1e worstvar=%.1e dE=%.0e" % (n, rel(sa[:, 2:-2, 2:-2], b.state[:, 2:-2, 2:-2]), pv, abs(st[1]-sb[1])/st[1]))
        print(f"{ic:16s} {name:20s}", " | ".join(out))
