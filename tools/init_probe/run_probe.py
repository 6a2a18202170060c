"""CPU check of the device init code: compile csrc/pmw_init.cuh's ic_cell for the host and compare
with the NumPy init (pyminiweather_b200.ics.init) for every configuration."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(here, "../..")))
from pyminiweather_b200 import _lib  # noqa: E402
from pyminiweather_b200.data import initialize_fields  # noqa: E402
from pyminiweather_b200.ics import init  # noqa: E402
from pyminiweather_b200.ics.initial_conditions import IC_TYPES, device_spec  # noqa: E402
from pyminiweather_b200.mesh import MeshData  # noqa: E402

so = os.path.join(here, "probe.so")
env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
subprocess.check_call([_lib.nvcc_path(), "-O2", "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared",
                       os.path.join(here, "probe.cu"), "-o", so], env=env)
lib = C.CDLL(so)


class Spec(C.Structure):  # csrc IcSpec = pmw_ic_spec + dx, dz
    _fields_ = _lib.PmwIcSpec._fields_ + [("dx", C.c_double), ("dz", C.c_double)]


for nx, nz in ((100, 50), (37, 19)):
    for ic in IC_TYPES:
        p = dict(nx=nx, nz=nz, xlen=2e4, zlen=1e4, hs=2, s=4, ic_type=ic)
        p["dx"], p["dz"] = p["xlen"] / nx, p["zlen"] / nz
        f = initialize_fields(p)
        m = MeshData(p)
        init(f, p, m)
        want = f._host[0]
        bubbles, wind, bv0 = device_spec(ic, p["xlen"])
        s = Spec()
        s.nbubbles = len(bubbles)
        for n, b in enumerate(bubbles):
            s.amp[n], s.x0[n], s.z0[n], s.xrad[n], s.zrad[n] = b
        s.wind, s.bvfreq, s.bv0, s.dx, s.dz = wind, int(bv0 is not None), bv0 or 0.0, p["dx"], p["dz"]
        xa, za = m.get_axes_int_ext()
        out = np.zeros_like(want)
        dp = C.POINTER(C.c_double)
        lib.probe_init(C.byref(s), nx + 4, nz + 4, xa.ctypes.data_as(dp), za.ctypes.data_as(dp), out.ctypes.data_as(dp))
        errs = [float(np.linalg.norm(out[v] - want[v]) / max(np.linalg.norm(want[v]), 1e-300)) for v in range(4)]
        print(f"{ic:16s} {nx}x{nz}: rel-L2 per variable {['%.1e' % e for e in errs]}  bit-equal: {np.array_equal(out, want)}")
        assert max(errs) <= 1e-13
