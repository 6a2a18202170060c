"""The reference-facing operator API (same names/arguments/in-place semantics as
pyminiweather.solve / .ics / .post) on the GPU: lazy device-resident Fields and the strict
drop-in mode for foreign Fields objects."""
import types

import numpy as np
import pytest

from helpers import HYDRO, interior, make_params, new_case, rel_l2, worst_rel_l2
from oracle import numpy_oracle as no

pytestmark = pytest.mark.gpu


def native_fields(nx, nz, ic):
    from pyminiweather_b200.data import initialize_fields
    from pyminiweather_b200.ics import init
    from pyminiweather_b200.mesh import MeshData
    p = make_params(nx, nz, ic)
    f = initialize_fields(p)
    m = MeshData(p)
    init(f, p, m)
    return p, f, m


def foreign_fields(case):
    """What the reference's dataclass looks like to us: plain attributes holding NumPy arrays."""
    f = types.SimpleNamespace(state=case.state.copy(), state_tmp=case.state_tmp.copy(), nvariables=4)
    for n in HYDRO:
        setattr(f, n, getattr(case, n).copy())
    return f


def test_main_loop_like_the_reference_driver():
    """__main__.py:205-248 with our modules: stats, warm-up, nsteps x evolve, stats."""
    from pyminiweather_b200.post import compute_stats
    from pyminiweather_b200.solve import evolve
    p, f, mesh = native_fields(100, 50, "thermal")
    _, case = new_case(100, 50, "thermal")
    m0, e0 = compute_stats(p, f)
    assert (m0, e0) == pytest.approx(no.compute_stats(case), rel=1e-13)
    for _ in range(25):
        evolve(p, f, mesh, dt=p["dt"])
        no.evolve(case)
    m1, e1 = compute_stats(p, f)
    mo, eo = no.compute_stats(case)
    assert abs(m1 - mo) / mo <= 1e-12 and abs(e1 - eo) / eo <= 1e-12
    assert worst_rel_l2(f.state, case.state) <= 1e-11
    assert rel_l2(interior(f.state_tmp), interior(case.state_tmp)) <= 1e-11
    # the whole loop ran without a single state transfer after the first upload
    assert f._solver.launch_count >= 25 * 6
    f.close()


def test_lazy_fields_host_writes_are_seen():
    from pyminiweather_b200.solve import evolve
    p, f, mesh = native_fields(64, 32, "collision")
    _, case = new_case(64, 32, "collision")
    evolve(p, f, mesh, dt=p["dt"]); no.evolve(case)
    f.state[1, 2:-2, 2:-2] += 0.5          # host-side edit between steps
    case.state[1, 2:-2, 2:-2] += 0.5
    evolve(p, f, mesh, dt=p["dt"]); no.evolve(case)
    assert worst_rel_l2(f.state, case.state) <= 1e-11
    f.state = case.state.copy() * 1.0      # wholesale replacement
    evolve(p, f, mesh, dt=p["dt"]); no.evolve(case)
    assert worst_rel_l2(f.state, case.state) <= 1e-11
    f.close()


def test_strict_dropin_with_foreign_fields_object():
    from pyminiweather_b200.post import compute_stats
    from pyminiweather_b200.solve import evolve
    import pyminiweather_b200.solve.step as step
    p, case = new_case(72, 40, "density-current")
    f = foreign_fields(case)
    step.SYNC_STATE_TMP = True
    try:
        for _ in range(5):
            evolve(p, f, None, dt=p["dt"])   # host arrays are current after every call
            no.evolve(case)
            assert worst_rel_l2(f.state, case.state) <= 1e-11
        assert rel_l2(interior(f.state_tmp), interior(case.state_tmp)) <= 1e-11
    finally:
        step.SYNC_STATE_TMP = False
    m, e = compute_stats(p, f)
    mo, eo = no.compute_stats(case)
    assert abs(m - mo) / mo <= 1e-12 and abs(e - eo) / eo <= 1e-12


@pytest.mark.parametrize("native", [True, False])
def test_discrete_step_and_bcs_operate_in_place_like_the_reference(native):
    from pyminiweather_b200.ics import Directions, set_bc_x, set_bc_z
    from pyminiweather_b200.solve import discrete_step
    if native:
        p, f, mesh = native_fields(48, 24, "collision")
        _, case = new_case(48, 24, "collision")
    else:
        p, case = new_case(48, 24, "collision")
        f, mesh = foreign_fields(case), None
    for _ in range(2):
        no.evolve(case)
    if native:
        f.state = case.state.copy(); f.state_tmp = case.state_tmp.copy()
    else:
        f.state[:] = case.state; f.state_tmp[:] = case.state_tmp
    st, tmp = f.state, f.state_tmp
    for d_ref, d in ((no.DIR_Z, Directions.Z), (no.DIR_X, Directions.X)):
        discrete_step(p, f, mesh, st, st, tmp, p["dt"] / 3, d)
        no.discrete_step(case, case.state, case.state, case.state_tmp, case.dt / 3, d_ref)
        # forcing received its halo cells in place (exact copies of cells that agree to rounding)
        assert np.allclose(st, case.state, rtol=1e-11, atol=1e-12)
        if d is Directions.X:
            assert np.array_equal(st[:, 2:-2, :2], st[:, 2:-2, -4:-2]) and np.array_equal(st[:, 2:-2, -2:], st[:, 2:-2, 2:4])
        else:
            assert not st[2, :2].any() and not st[2, -2:].any()
            assert np.array_equal(st[0, 0], st[0, 2]) and np.array_equal(st[3, -1], st[3, -3])
        assert worst_rel_l2(tmp, case.state_tmp) <= 1e-12
        discrete_step(p, f, mesh, st, tmp, tmp, p["dt"] / 2, d)
        no.discrete_step(case, case.state, case.state_tmp, case.state_tmp, case.dt / 2, d_ref)
        discrete_step(p, f, mesh, st, tmp, st, p["dt"], d)
        no.discrete_step(case, case.state, case.state_tmp, case.state, case.dt, d_ref)
        assert worst_rel_l2(st, case.state) <= 1e-12
        assert worst_rel_l2(tmp, case.state_tmp) <= 1e-12
    # an array that belongs to nobody
    rng = np.random.default_rng(0)
    a = rng.standard_normal(st.shape); b = a.copy()
    set_bc_x(p, f, a, "collision"); no.set_bc_x(case, b)
    assert np.array_equal(a, b)
    set_bc_z(p, f, a, "collision"); no.set_bc_z(case, b)
    assert np.array_equal(a, b)


def test_unsupported_configurations_raise():
    from pyminiweather_b200._lib import PmwError
    from pyminiweather_b200.engine import DeviceSolver
    from pyminiweather_b200.solve import evolve
    p, f, mesh = native_fields(32, 16, "thermal")
    for ic in ("injection", "gravity"):
        with pytest.raises(NotImplementedError):
            evolve(dict(p, ic_type=ic), f, mesh, dt=p["dt"])
    with pytest.raises(PmwError, match="hs must be 2"):
        DeviceSolver(32, 16, 1.0, 1.0, 0.1, hs=3)
    with pytest.raises(PmwError, match=">= 4"):
        DeviceSolver(2, 16, 1.0, 1.0, 0.1)
    s = DeviceSolver(32, 16, 1.0, 1.0, 0.1)
    with pytest.raises(PmwError, match="hydrostatic"):
        s.evolve(1)
    with pytest.raises(ValueError):
        s.upload(0, np.zeros((4, 20, 35)))
    s.close()
