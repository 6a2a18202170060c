"""Generate the golden fixtures in this directory by running the REAL reference
(``/root/reference``, NumPy backend) in the build container.

    PYTHONPATH=/root/repo python tests/golden/make_golden.py            # everything
    PYTHONPATH=/root/repo python tests/golden/make_golden.py injection  # only the named sections

Versions used for the committed fixtures are recorded in ``manifest.json``.
The reference is imported read-only; nothing is copied from it.
"""
import json
import os
import sys

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "../..")))
from oracle import reference_runner as rr  # noqa: E402

manifest = {"numpy": np.__version__, "scipy": scipy.__version__,
            "python": sys.version.split()[0], "files": {}}
ONLY = set(sys.argv[1:])
if ONLY and os.path.exists(os.path.join(HERE, "manifest.json")):  # partial run: keep the other entries
    with open(os.path.join(HERE, "manifest.json")) as fh:
        manifest["files"] = json.load(fh)["files"]


def want(section):
    return not ONLY or section in ONLY


def hydro(f):
    return dict(hy_dens_cell=f.hy_dens_cell, hy_dens_theta_cell=f.hy_dens_theta_cell,
                hy_dens_int=f.hy_dens_int, hy_dens_theta_int=f.hy_dens_theta_int,
                hy_pressure_int=f.hy_pressure_int)


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    manifest["files"][name] = sorted(arrs)
    print("wrote", name, os.path.getsize(path) // 1024, "KiB")


def scal(p):
    return dict(nx=p["nx"], nz=p["nz"], dx=p["dx"], dz=p["dz"], dt=p["dt"])


# 1. initial conditions (all five --ic-type choices) on a small grid
if want("ics"):
    for ic in ["thermal", "collision", "density-current", "gravity", "injection"]:
        r = rr.ReferenceRun(32, 16, ic)
        save(f"ic_{ic}_32x16.npz", state=r.fields.state, state_tmp=r.fields.state_tmp,
             stats=np.array(r.stats()), **hydro(r.fields), **scal(r.params))

# 2. single stages through the reference's discrete_step, all three aliasing patterns
#    (step.py:112-141), both directions, full arrays including halos
if want("stages"):
    for ic, (nx, nz) in [("collision", (48, 24)), ("thermal", (37, 19))]:   # second grid is odd-sized
        r = rr.ReferenceRun(nx, nz, ic)
        r.evolve(3)                      # non-trivial momentum everywhere
        f, p = r.fields, r.params
        out = dict(state0=f.state.copy(), tmp0=f.state_tmp.copy(), **hydro(f), **scal(p))
        for dname, d in (("x", 1), ("z", 2)):
            st, tmp = out["state0"].copy(), out["tmp0"].copy()
            r.discrete_step(st, st, tmp, p["dt"] / 3, d)       # S1: init is forcing
            out[f"{dname}_s1_state"], out[f"{dname}_s1_tmp"] = st.copy(), tmp.copy()
            r.discrete_step(st, tmp, tmp, p["dt"] / 2, d)      # S2: out is forcing
            out[f"{dname}_s2_state"], out[f"{dname}_s2_tmp"] = st.copy(), tmp.copy()
            r.discrete_step(st, tmp, st, p["dt"] / 1, d)       # S3: out is init
            out[f"{dname}_s3_state"], out[f"{dname}_s3_tmp"] = st.copy(), tmp.copy()
        save(f"stages_{ic}_{nx}x{nz}.npz", **out)

# 3. boundary conditions alone on a random array (halos start as garbage)
if want("bc"):
    r = rr.ReferenceRun(20, 12, "thermal")
    rng = np.random.default_rng(7)
    s = rng.standard_normal(r.fields.state.shape)
    from pyminiweather.ics import set_bc_x, set_bc_z  # noqa: E402  (reference)
    sx, sz = s.copy(), s.copy()
    set_bc_x(r.params, r.fields, sx, "thermal")
    set_bc_z(r.params, r.fields, sz, "thermal")
    save("bc_random_20x12.npz", s=s, after_bc_x=sx, after_bc_z=sz, **hydro(r.fields), **scal(r.params))

# 4. multi-step evolution, BASELINE config 1 grid
if want("evolve"):
    for ic, snaps in [("thermal", [1, 2, 10, 100, 1000]), ("collision", [100]), ("density-current", [100])]:
        r = rr.ReferenceRun(100, 50, ic)
        out = dict(state0=r.fields.state.copy(), stats0=np.array(r.stats()), **hydro(r.fields), **scal(r.params))
        done = 0
        for n in snaps:
            r.evolve(n - done)
            done = n
            out[f"state_{n}"] = r.fields.state.copy()
            out[f"tmp_{n}"] = r.fields.state_tmp[:, 2:-2, 2:-2].copy() if n <= 2 else np.zeros(0)
            out[f"stats_{n}"] = np.array(r.stats())
        save(f"evolve_{ic}_100x50.npz", **out)

# 5. a mid-size grid, sub-sampled (thermal 512x256, 5 steps)
if want("midsize"):
    r = rr.ReferenceRun(512, 256, "thermal")
    st0 = np.array(r.stats())
    r.evolve(5)
    inner = r.fields.state[:, 2:-2, 2:-2]
    save("evolve_thermal_512x256_5steps_sub8.npz", sub=inner[:, ::8, ::8].copy(),
         l2=np.array([np.linalg.norm(inner[v]) for v in range(4)]),
         stats0=st0, stats5=np.array(r.stats()), **scal(r.params))

# 5b. BASELINE config 2 itself (thermal 2048x1024) after 1, 2, 5 and 10 steps of the reference (~16 s per step):
#     every 32nd cell of every variable, the per-variable L2 norms of the full interior, the totals
if want("config2"):
    r = rr.ReferenceRun(2048, 1024, "thermal")
    out = dict(stats_0=np.array(r.stats()), **scal(r.params))
    done = 0
    for n in (1, 2, 5, 10):
        r.evolve(n - done)
        done = n
        inner = r.fields.state[:, 2:-2, 2:-2]
        out[f"sub_{n}"] = inner[:, ::32, ::32].copy()
        out[f"l2_{n}"] = np.array([np.linalg.norm(inner[v]) for v in range(4)])
        out[f"stats_{n}"] = np.array(r.stats())
    save("evolve_thermal_2048x1024_10steps_sub32.npz", **out)

# 6. gravity-wave configuration (extra w-momentum source in every stage, source.py:20-50)
if want("gravity"):
    r = rr.ReferenceRun(100, 50, "gravity")
    out = dict(state0=r.fields.state.copy(), stats0=np.array(r.stats()), **hydro(r.fields), **scal(r.params))
    done = 0
    for n in (1, 2, 20):
        r.evolve(n - done)
        done = n
        out[f"state_{n}"] = r.fields.state.copy()
        out[f"stats_{n}"] = np.array(r.stats())
    save("evolve_gravity_100x50.npz", **out)

# 7. injection configuration (non-periodic x halo fill with a forced inflow jet, bcs.py:37,41-64):
#    the halo fill alone on random data, and a multi-step evolution (full arrays: the right halo
#    columns must keep their initial values)
if want("injection"):
    r = rr.ReferenceRun(20, 12, "injection")
    from pyminiweather.ics import set_bc_x  # noqa: E402  (reference)
    rng = np.random.default_rng(11)
    s = rng.standard_normal(r.fields.state.shape)
    sx = s.copy()
    set_bc_x(r.params, r.fields, sx, "injection")
    save("bc_injection_random_20x12.npz", s=s, after_bc_x=sx, zlen=r.params["zlen"], **hydro(r.fields),
         **scal(r.params))
    r = rr.ReferenceRun(100, 50, "injection")
    out = dict(state0=r.fields.state.copy(), stats0=np.array(r.stats()), zlen=r.params["zlen"],
               **hydro(r.fields), **scal(r.params))
    done = 0
    for n in (1, 2, 50, 300):
        r.evolve(n - done)
        done = n
        out[f"state_{n}"] = r.fields.state.copy()
        out[f"tmp_{n}"] = r.fields.state_tmp.copy() if n <= 2 else np.zeros(0)
        out[f"stats_{n}"] = np.array(r.stats())
    save("evolve_injection_100x50.npz", **out)

with open(os.path.join(HERE, "manifest.json"), "w") as fh:
    json.dump(manifest, fh, indent=1, sort_keys=True)
