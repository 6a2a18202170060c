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
    assert f._solver.launch_count >= 25 * 2  # two fused sweep kernels per step (six stage kernels with fuse=0)
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
    with pytest.raises(ValueError, match="unknown ic_type"):
        evolve(dict(p, ic_type="squall-line"), f, mesh, dt=p["dt"])
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


def test_gravity_through_the_operator_api():
    """The source field is built and uploaded by the operators themselves from params/fields."""
    from pyminiweather_b200.post import compute_stats
    from pyminiweather_b200.solve import evolve
    p, f, mesh = native_fields(100, 50, "gravity")
    _, case = new_case(100, 50, "gravity")
    for _ in range(8):
        evolve(p, f, mesh, dt=p["dt"])
        no.evolve(case)
    assert worst_rel_l2(f.state, case.state) <= 1e-11
    m, e = compute_stats(p, f)
    mo, eo = no.compute_stats(case)
    assert abs(m - mo) / mo <= 1e-12 and abs(e - eo) / eo <= 1e-12
    f.close()
    ff = foreign_fields(new_case(100, 50, "gravity")[1])
    _, c2 = new_case(100, 50, "gravity")
    for _ in range(2):
        evolve(p, ff, None, dt=p["dt"])
        no.evolve(c2)
    assert worst_rel_l2(ff.state, c2.state) <= 1e-11


# ---- unfused operator shims (reference: pyminiweather/solve/interpolate.py) ------------------------
def test_reference_test_interpolate_on_the_gpu_shims():
    """tests/unit/test_interpolate.py:11-63 of the reference, against our interpolate_x/z: arange
    state at the default 200x100 grid vs scipy.signal.convolve2d, np.allclose."""
    from scipy.signal import convolve2d
    from pyminiweather_b200.data import initialize_fields
    from pyminiweather_b200.solve import interpolate_x, interpolate_z
    p = make_params(200, 100)
    nx, nz = 200, 100
    f = initialize_fields(p)
    state = np.arange(np.prod(f.shape)).astype(np.float64).reshape(f.shape)
    f.state = state
    k4 = np.array([-1.0 / 12, 7.0 / 12, 7.0 / 12, -1.0 / 12])
    interpolate_x(p, f)
    interpolate_z(p, f)
    for v in range(4):
        assert np.allclose(convolve2d(state[v, 2:nz + 2, :], k4[None, :], mode="same")[:, 2:-1], f.vals_x[v])
        assert np.allclose(convolve2d(state[v, :, 2:nx + 2], k4[:, None], mode="same")[2:-1, :], f.vals_z[v])
    f.close()


@pytest.mark.parametrize("native", [True, False])
def test_unfused_operator_chain_matches_oracle(native):
    """interpolate -> flux -> tend, both directions, filling fields.vals_*/d3_vals_*/flux/tend like
    the reference (solve/step.py:67-76 without the fusion)."""
    from pyminiweather_b200.solve import (compute_flux_x, compute_flux_z, compute_tend_x, compute_tend_z,
                                          interpolate_x, interpolate_z)
    if native:
        p, f, _ = native_fields(52, 26, "collision")
        _, case = new_case(52, 26, "collision")
    else:
        p, case = new_case(52, 26, "collision")
        f = foreign_fields(case)
        f.vals_x = np.zeros((4, 26, 53)); f.d3_vals_x = np.zeros((4, 26, 53))
        f.vals_z = np.zeros((4, 27, 52)); f.d3_vals_z = np.zeros((4, 27, 52))
        f.flux = np.zeros((4, 27, 53)); f.tend = np.zeros((4, 26, 52))
    for _ in range(3):
        no.evolve(case)
    no.set_bc_x(case, case.state); no.set_bc_z(case, case.state)
    if native:
        f.state = case.state.copy()
    else:
        f.state[:] = case.state
    st = f.state
    # x
    interpolate_x(p, f, st)
    vals, d3 = no.interpolate_x(case, case.state)
    assert np.array_equal(f.vals_x, vals) and np.array_equal(f.d3_vals_x, d3)        # bit-exact
    compute_flux_x(p, f)
    fx = no.compute_flux_x(case, vals, d3)
    assert rel_l2(f.flux[:, :26, :53], fx) <= 1e-15
    compute_tend_x(p, f, st)
    want = -(f.flux[:, :26, 1:53] - f.flux[:, :26, 0:52]) / case.dx
    assert np.array_equal(f.tend, want)                                               # bit-exact given flux
    # z
    interpolate_z(p, f, st)
    vals, d3 = no.interpolate_z(case, case.state)
    assert np.array_equal(f.vals_z, vals) and np.array_equal(f.d3_vals_z, d3)
    compute_flux_z(p, f)
    fz = no.compute_flux_z(case, vals, d3)
    assert rel_l2(f.flux[:, :27, :52], fz) <= 1e-13   # z flux is a pressure PERTURBATION: pow ulp / cancellation
    assert not f.d3_vals_z[0, 0].any() and not f.d3_vals_z[0, 26].any()
    compute_tend_z(p, f, st)
    want = -(f.flux[:, 1:27, :52] - f.flux[:, 0:26, :52]) / case.dz
    want[2] -= case.state[0, 2:-2, 2:-2] * no.GRAV
    assert np.array_equal(f.tend, want)


# ---- command-line driver (reference: pyminiweather/__main__.py:183-248) ------------------------------
def test_cli_driver_end_to_end(tmp_path, caplog):
    import logging
    from conftest import golden
    from pyminiweather_b200.__main__ import main
    g = golden("evolve_thermal_100x50.npz")
    out = tmp_path / "dump.txt"
    with caplog.at_level(logging.INFO, logger="pyminiweather"):
        rc = main(["--nx", "100", "--nz", "50", "--nsteps", "10", "--output-freq", "5", "--app-filename", str(out)])
    assert rc == 0
    text = caplog.text
    start = [ln for ln in text.splitlines() if "Start: total_mass" in ln][0]
    end = [ln for ln in text.splitlines() if "End: total_mass" in ln][0]
    m0, e0 = [float(x) for x in start.split(":")[-1].split(",")]
    m1, e1 = [float(x) for x in end.split(":")[-1].split(",")]
    assert abs(m0 - g["stats0"][0]) / m0 <= 1e-12 and abs(e0 - g["stats0"][1]) / e0 <= 1e-12
    assert abs(m1 - g["stats_10"][0]) / m1 <= 1e-12 and abs(e1 - g["stats_10"][1]) / e1 <= 1e-12
    # two dumps (before steps 5 and 10), in the reference's text layout: rows of nx+4 comma-separated values
    dumped = np.loadtxt(out, delimiter=",")
    assert dumped.shape == (2 * 4 * 54, 104)
    _, case = new_case(100, 50, "thermal")
    for _ in range(4):
        no.evolve(case)
    first = dumped[:4 * 54].reshape(4, 54, 104)
    assert worst_rel_l2(first, case.state) <= 1e-11
    sv = np.loadtxt(str(out).replace(".txt", "_svars.txt"), delimiter=",")
    assert sv.shape == (2 * 4 * 50, 100)
    assert rel_l2(sv[:200].reshape(4, 50, 100), no.compute_solution_variables(case)) <= 1e-11


def test_injection_through_the_operator_api():
    """ic_type 'injection' (bcs.py:37,41-64) through the reference-shaped functions: native fields,
    a foreign (strict drop-in) container, and set_bc_x's own ic_type argument."""
    from pyminiweather_b200.ics import set_bc_x
    from pyminiweather_b200.post import compute_stats
    from pyminiweather_b200.solve import evolve
    p, f, mesh = native_fields(64, 32, "injection")
    _, case = new_case(64, 32, "injection")
    ff = foreign_fields(case)
    for _ in range(30):
        evolve(p, f, mesh, dt=p["dt"])
        evolve(p, ff, mesh, dt=p["dt"])
        no.evolve(case)
    assert worst_rel_l2(f.state, case.state) <= 1e-11
    assert worst_rel_l2(ff.state, case.state) <= 1e-11
    assert np.linalg.norm(f.state[1, 2:-2, 2:-2]) > 1.0  # the jet is in
    m, e = compute_stats(p, f)
    mo, eo = no.compute_stats(case)
    assert abs(m - mo) / mo <= 1e-12 and abs(e - eo) / eo <= 1e-12
    # set_bc_x branches on its ic_type ARGUMENT, like the reference
    rng = np.random.default_rng(3)
    a = rng.standard_normal(f.state.shape); b = a.copy(); c = a.copy()
    set_bc_x(p, f, a, "injection"); no.set_bc_x(case, b)
    assert np.array_equal(a, b)
    periodic = case.copy(); periodic.inflow_zlen = None
    a = c.copy()
    set_bc_x(p, f, a, "thermal"); no.set_bc_x(periodic, c)
    assert np.array_equal(a, c)
    f.close()


@pytest.mark.parametrize("ic", ["thermal", "collision", "density-current", "gravity", "injection"])
def test_device_init_matches_host_init(ic):
    """ics.init_device (init_state_kernel: the 3x3 quadrature of initial.py:57-80 on the GPU) against
    the host init, which is bit-identical to the reference's: <= 1e-13 relative L2 per variable (CUDA
    vs NumPy pow/cos/exp rounding), exact zeros where the reference has exact zeros, profiles equal."""
    from pyminiweather_b200.data import initialize_fields
    from pyminiweather_b200.ics import init_device
    from pyminiweather_b200.mesh import MeshData
    for nx, nz in ((100, 50), (37, 19)):
        p, fh, mesh = native_fields(nx, nz, ic)
        fd = initialize_fields(p)
        init_device(fd, p, MeshData(p))
        want, got, got_tmp = fh._host[0], fd.state, fd.state_tmp
        for v in range(4):
            n = np.linalg.norm(want[v])
            if n == 0.0:
                assert not got[v].any(), (ic, v)
            else:
                assert np.linalg.norm(got[v] - want[v]) / n <= 1e-13, (ic, v)
        assert np.array_equal(got == 0.0, want == 0.0)          # same support, halo cells included
        assert np.array_equal(got, got_tmp)                      # initial.py:80
        for name in HYDRO:
            assert np.array_equal(getattr(fd, name), getattr(fh, name))
        fd.close(); fh.close()


def test_evolve_from_device_init_vs_oracle():
    """End to end without the host quadrature: device init, 20 steps, against the oracle started from
    the (reference-identical) host init."""
    from pyminiweather_b200.data import initialize_fields
    from pyminiweather_b200.ics import init_device
    from pyminiweather_b200.mesh import MeshData
    from pyminiweather_b200.post import compute_stats
    from pyminiweather_b200.solve import evolve
    for ic in ("collision", "gravity"):
        p, case = new_case(256, 128, ic)
        f = initialize_fields(p)
        mesh = MeshData(p)
        init_device(f, p, mesh)
        m0, e0 = compute_stats(p, f)
        mo, eo = no.compute_stats(case)
        assert abs(m0 - mo) / mo <= 1e-12 and abs(e0 - eo) / eo <= 1e-12
        for _ in range(20):
            evolve(p, f, mesh, dt=p["dt"])
            no.evolve(case)
        assert worst_rel_l2(f.state, case.state) <= 1e-11, ic
        f.close()


def test_device_init_of_slabs_matches_the_whole_domain():
    """pmw_init_state on x-slab contexts (each given its own part of the x axis, SlabMesh) against
    the columns of the whole-domain device init."""
    from pyminiweather_b200.engine import DeviceSolver
    from pyminiweather_b200.ics.initial_conditions import device_spec
    from pyminiweather_b200.mesh import MeshData
    from pyminiweather_b200.slab import SlabMesh
    nx, nz, world = 192, 48, 3
    p = make_params(nx, nz, "collision")
    bubbles, wind, bv0 = device_spec("collision", p["xlen"])
    whole = DeviceSolver(nx, nz, p["dx"], p["dz"], p["dt"])
    whole.init_state(bubbles, wind, bv0, *MeshData(p).get_axes_int_ext())
    want = whole.download(0)
    assert np.linalg.norm(want[3]) > 1.0
    nxl = nx // world
    for r in range(world):
        ps = dict(p, nx=nxl)
        s = DeviceSolver(nxl, nz, p["dx"], p["dz"], p["dt"], periodic_x=False)
        s.init_state(bubbles, wind, bv0, *SlabMesh(ps, r, world).get_axes_int_ext())
        got, ref = s.download(0), want[:, :, r * nxl: (r + 1) * nxl + 4]
        assert np.linalg.norm(got - ref) <= 1e-13 * np.linalg.norm(want), r
        assert np.array_equal(got, s.download(1))
        s.close()
    whole.close()


def test_sweep_order_belongs_to_the_simulation_and_can_be_reset():
    """step.py:18,103,143 of the reference: Z,X then X,Z alternating.  The flag survives the re-creation of a
    fields object's context (dt changed) and the reference's reset idiom ``step.reverse_direction = False``
    restarts the sequence."""
    import pyminiweather_b200.solve.step as step
    from pyminiweather_b200.solve import evolve
    p, f, mesh = native_fields(64, 32, "thermal")
    _, case = new_case(64, 32, "thermal")
    evolve(p, f, mesh, dt=p["dt"]); no.evolve(case)
    assert step.reverse_direction is True
    p2 = dict(p, dt=p["dt"] / 2)            # new context behind the same fields: the order carries over
    evolve(p2, f, mesh, dt=p2["dt"])
    case.dt = p2["dt"]; no.evolve(case)
    assert step.reverse_direction is False
    assert worst_rel_l2(f.state, case.state) <= 1e-11
    evolve(p2, f, mesh, dt=p2["dt"]); no.evolve(case)       # Z,X again -> flag True
    step.reverse_direction = False                          # the reference's reset before a new run
    evolve(p2, f, mesh, dt=p2["dt"])
    case.reverse_direction = False; no.evolve(case)
    assert worst_rel_l2(f.state, case.state) <= 1e-11
    f.close()


@pytest.mark.parametrize("nx,nz,ic,bands", [(1024, 512, "thermal", 0), (896, 480, "collision", 5), (1024, 416, "thermal", 13),
                                             (96, 64, "density-current", 4)])
def test_evolve_host_streams_row_bands_with_the_bits_of_the_plain_sequence(nx, nz, ic, bands):
    """pmw_evolve_host (the strict drop-in's evolve on host arrays): upload, sweeps and download band by band must
    give the bits of upload + evolve(1) + download, in both sweep orders (steps alternate Z,X / X,Z), for band
    counts that do not divide nz, and on grids the banded path does not cover (small ones: plain sequence)."""
    from pyminiweather_b200._lib import PMW_BUF_STATE
    from pyminiweather_b200.engine import DeviceSolver
    _, case = new_case(nx, nz, ic)
    solvers = []
    for _ in range(2):
        s = DeviceSolver(nx, nz, case.dx, case.dz, case.dt)
        s.set_hydrostatic(*[getattr(case, k) for k in HYDRO])
        solvers.append(s)
    plain, banded = solvers
    a, b = case.state.copy(), case.state.copy()
    for step in range(4):
        plain.upload(PMW_BUF_STATE, a); plain.evolve(1); plain.download(PMW_BUF_STATE, out=a)
        banded.evolve_host(b, None, bands)
        assert banded.reverse_direction == plain.reverse_direction == bool((step + 1) & 1)
        assert np.array_equal(interior(a), interior(b)), f"step {step}"
        # the halo columns are the periodic images the last sweep stored, as after a plain download
        assert np.array_equal(a[:, 2:-2, :], b[:, 2:-2, :])
    # the context stays usable for device-resident stepping afterwards
    plain.evolve(2); banded.evolve(2)
    assert np.array_equal(interior(plain.download(PMW_BUF_STATE)), interior(banded.download(PMW_BUF_STATE)))
    plain.close(); banded.close()


def test_strict_dropin_uses_the_streamed_step_and_matches_the_oracle():
    from pyminiweather_b200.solve import evolve
    p, case = new_case(512, 384, "collision")
    f = foreign_fields(case)
    for _ in range(3):
        evolve(p, f, None, dt=p["dt"])
        no.evolve(case)
        assert worst_rel_l2(f.state, case.state) <= 1e-11


def test_streamed_dropin_on_config_2_vs_the_reference_fixture():
    """BASELINE config 2 (thermal 2048x1024) through the strict drop-in -- ``evolve`` on host arrays, i.e. the
    streamed step (pmw_evolve_host, 16 bands) -- against the fixture generated from the reference's NumPy backend:
    every 32nd cell and the per-variable norms of the whole interior after 1, 2, 5 and 10 steps (stacked metric
    <= 1e-11; rho' at its ill-conditioned bound, see test_gpu_parity)."""
    import os
    from pyminiweather_b200.solve import evolve
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "evolve_thermal_2048x1024_10steps_sub32.npz"))
    p, case = new_case(2048, 1024, "thermal")
    f = foreign_fields(case)
    done = 0
    for n in (1, 2, 5, 10):
        for _ in range(n - done):
            evolve(p, f, None, dt=p["dt"])
        done = n
        got = interior(f.state)
        sub, want = got[:, ::32, ::32], g[f"sub_{n}"]
        assert np.linalg.norm(sub - want) / np.linalg.norm(want) <= 1e-11
        for v in range(4):
            assert abs(np.linalg.norm(got[v]) - g[f"l2_{n}"][v]) / g[f"l2_{n}"][v] <= (1e-9 if v == 0 else 1e-12)
