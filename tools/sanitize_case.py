"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck): both kernel
variants, odd and ragged grids, the operator API path and the fused evolve path."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from helpers import new_case, HYDRO, worst_rel_l2
from oracle import numpy_oracle as no
from pyminiweather_b200.engine import DeviceSolver
ONLY_NEW = len(sys.argv) > 1 and sys.argv[1] == "new"   # only the paths added last (bottom of the file)
for variant in (() if ONLY_NEW else ("tma", "direct")):
    for nx, nz in ((130, 37), (64, 12), (254, 23)):
        p, case = new_case(nx, nz, "collision")
        s = DeviceSolver(nx, nz, case.dx, case.dz, case.dt, variant=variant)
        s.set_hydrostatic(*[getattr(case, n) for n in HYDRO]); s.upload(0, case.state); s.upload(1, case.state_tmp)
        s.evolve(2); no.evolve(case); no.evolve(case)
        s.discrete_step(1, 0, 0, 1, case.dt / 3); s.discrete_step(2, 0, 1, 1, case.dt / 2)
        no.discrete_step(case, case.state, case.state, case.state_tmp, case.dt / 3, 1)
        no.discrete_step(case, case.state, case.state_tmp, case.state_tmp, case.dt / 2, 2)
        err = worst_rel_l2(s.download(1), case.state_tmp)
        st = s.stats(0)
        # ring of one through the peer path
        b = DeviceSolver(nx, nz, case.dx, case.dz, case.dt, variant=variant, periodic_x=False)
        b.set_hydrostatic(*[getattr(case, n) for n in HYDRO]); b.upload(0, case.state); b.upload(1, case.state_tmp)
        mine = b.local_ptrs(); b.connect_peers(mine, mine); b.evolve(2)
        assert not b.peer_timed_out()
        print(variant, nx, nz, "err %.2e" % err, st, flush=True)
        assert err < 1e-11
        s.close(); b.close()
        if variant == "tma":  # the other sweep organisations: transposing z sweep, stage-by-stage kernels
            _, c2 = new_case(nx, nz, "collision")
            no.evolve(c2); no.evolve(c2); no.evolve(c2)
            for tune in (dict(fuse=1, sweep_zt=1), dict(fuse=1, sweep_zt=0, sweep_lz=9), dict(fuse=1, sweep_zt=0, sweep_z3=1, sweep_lz=9),
                         dict(fuse=0)):
                t = DeviceSolver(nx, nz, case.dx, case.dz, case.dt, variant=variant)
                t.set_hydrostatic(*[getattr(case, n) for n in HYDRO]); t.set_tuning(**tune)
                _, c0 = new_case(nx, nz, "collision")
                t.upload(0, c0.state); t.upload(1, c0.state_tmp)
                t.evolve(3)
                e2 = worst_rel_l2(t.download(0), c2.state)
                print("  ", tune, "err %.2e" % e2, flush=True)
                assert e2 < 1e-11
                t.close()
if not ONLY_NEW:
    # kernels added later: gravity-wave forcing in the fused sweeps (HAS_SRC), the injection inflow fill, the device-side
    # init, the pow() fallbacks inside the fused sweeps (cold call / bail-out to the generic iteration), the diagnostics
    # kernel on an odd width, the FP64 probe
    from helpers import synthetic_case
    p, case = new_case(130, 33, "gravity")
    for tune in (dict(fuse=1, sweep_zt=0, sweep_lz=9), dict(fuse=1, sweep_zt=0, sweep_z3=1, sweep_lz=9)):
        s = DeviceSolver(130, 33, case.dx, case.dz, case.dt)
        s.set_hydrostatic(*[getattr(case, n) for n in HYDRO]); s.set_tuning(**tune); s.set_source_w(case.source_w)
        s.upload(0, case.state); s.upload(1, case.state_tmp); s.evolve(3); print("gravity", tune, s.stats(0), flush=True); s.close()
    from pyminiweather_b200.data import initialize_fields
    from pyminiweather_b200.ics import init_device
    from pyminiweather_b200.mesh import MeshData
    from pyminiweather_b200.solve import evolve
    from helpers import make_params
    for ic in ("injection", "collision"):
        pp = make_params(64, 32, ic)
        f = initialize_fields(pp); m = MeshData(pp); init_device(f, pp, m)
        for _ in range(3):
            evolve(pp, f, m, dt=pp["dt"])
        print(ic, "device init + evolve", float(np.abs(f.state).max()), flush=True); f.close()
    p, case = synthetic_case(150, 40, seed=5)
    case.state[3, 10:25, 20:90] += 0.2 * case.hy_dens_theta_cell[10:25, None]
    case.state_tmp[:] = case.state
    for tune in (dict(fuse=1, sweep_zt=0, sweep_lz=16), dict(fuse=1, sweep_zt=0, sweep_z3=1, sweep_lz=16), dict(fuse=1, sweep_zt=1)):
        s = DeviceSolver(150, 40, case.dx, case.dz, case.dt)
        s.set_hydrostatic(*[getattr(case, n) for n in HYDRO]); s.set_tuning(**tune)
        s.upload(0, case.state); s.upload(1, case.state_tmp); s.evolve(2); print("pow fallback", tune, s.stats(0), flush=True); s.close()
    p, case = synthetic_case(101, 37, seed=2)
    s = DeviceSolver(101, 37, case.dx, case.dz, case.dt, variant="direct")
    s.set_hydrostatic(*[getattr(case, n) for n in HYDRO]); s.upload(0, case.state); s.upload(1, case.state_tmp)
    print("stats odd nx", s.stats(0), "fp64 probe", s.fp64_peak(), flush=True); s.close()
# round 2, last additions: the streamed host step (row-range launches of the fused sweeps, both sweep orders, ragged
# bands), state_tmp on demand (the re-run of the last sweep), the diagnostics kernel with its prefetch loop
from helpers import new_case as _nc
for nx, nz, bands in ((96, 64, 3), (130, 50, 3)):
    p, case = _nc(nx, nz, "collision")
    a = DeviceSolver(nx, nz, case.dx, case.dz, case.dt); b = DeviceSolver(nx, nz, case.dx, case.dz, case.dt)
    for s_ in (a, b):
        s_.set_hydrostatic(*[getattr(case, n) for n in HYDRO]); s_.set_tuning(sweep_zt=0)
    ha, hb = case.state.copy(), case.state.copy()
    for _ in range(2):
        a.upload(0, ha); a.evolve(1); a.download(0, out=ha)
        b.evolve_host(hb, None, bands)
    assert np.array_equal(ha[:, 2:-2, :], hb[:, 2:-2, :])
    a.evolve(1); t = a.download(1)                      # state_tmp on demand
    b.set_tuning(keep_tmp=2); b.evolve(1)
    assert np.array_equal(t[:, 2:-2, 2:-2], b.download(1)[:, 2:-2, 2:-2])
    print("host step / lazy tmp", nx, nz, a.stats(0), flush=True)
    a.close(); b.close()
print("sanitize case ok")
