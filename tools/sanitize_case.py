"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck): both kernel
variants, odd and ragged grids, the operator API path and the fused evolve path."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from helpers import new_case, HYDRO, worst_rel_l2
from oracle import numpy_oracle as no
from pyminiweather_b200.engine import DeviceSolver
for variant in ("tma", "direct"):
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
            for tune in (dict(fuse=1, sweep_zt=1), dict(fuse=1, sweep_zt=0, sweep_lz=9), dict(fuse=0)):
                t = DeviceSolver(nx, nz, case.dx, case.dz, case.dt, variant=variant)
                t.set_hydrostatic(*[getattr(case, n) for n in HYDRO]); t.set_tuning(**tune)
                _, c0 = new_case(nx, nz, "collision")
                t.upload(0, c0.state); t.upload(1, c0.state_tmp)
                t.evolve(3)
                e2 = worst_rel_l2(t.download(0), c2.state)
                print("  ", tune, "err %.2e" % e2, flush=True)
                assert e2 < 1e-11
                t.close()
print("sanitize case ok")
