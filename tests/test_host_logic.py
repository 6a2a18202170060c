"""Host logic of the operator layer on the CPU: ``Fields`` (lazy device/host synchronisation),
``_dispatch`` (context cache for foreign containers, source / inflow configuration) and the
``solve`` / ``ics`` / ``post`` functions, with the device context replaced by a stand-in that does the
arithmetic with the NumPy oracle.  What is under test is everything ABOVE the C ABI: which logical buffer
each call touches, the three aliasing patterns of ``discrete_step``, when data crosses the (here
imaginary) PCIe link, how ``ic_type`` reconfigures the context.  Results must equal the oracle bit for
bit, because the stand-in *is* the oracle.  (The real kernels are checked against the same oracle in the
``-m gpu`` suite.)"""
import types

import numpy as np
import pytest

from helpers import HYDRO, make_params, new_case
from oracle import numpy_oracle as no

STATE, TMP = 0, 1


class OracleSolver:
    """Same surface as pyminiweather_b200.engine.DeviceSolver; buffers are NumPy arrays."""
    instances = []

    def __init__(self, nx, nz, dx, dz, dt, *, hs=2, **_kw):
        self.nx, self.nz, self.hs, self.dx, self.dz, self.dt = nx, nz, hs, dx, dz, dt
        self.shape = (4, nz + 4, nx + 4)
        self.buf = [np.zeros(self.shape), np.zeros(self.shape)]
        self.case = None
        self._hydro = self._source = self._inflow_key = None
        self.inflow = None
        self.reverse_direction = False
        self.uploads, self.downloads, self.launch_count, self.closed = [], [], 0, False
        OracleSolver.instances.append(self)

    # data
    def set_hydrostatic(self, *arrs):
        self._hydro = [np.array(a, dtype=np.float64) for a in arrs]

    def hydro_matches(self, arrs):
        return self._hydro is not None and all(np.array_equal(a, b) for a, b in zip(self._hydro, arrs))

    def set_source_w(self, field):
        self._source = None if field is None else np.array(field)

    def set_inflow(self, mask, u_in=50.0, theta_in=298.0):
        self.inflow = None if mask is None else (np.array(mask), u_in, theta_in)

    def upload(self, buf, host, asynchronous=False):
        assert host.shape == self.shape
        self.buf[buf][...] = host
        self.uploads.append(buf)

    def download(self, buf, out=None, asynchronous=False):
        self.downloads.append(buf)
        if out is None:
            return self.buf[buf].copy()
        out[...] = self.buf[buf]
        return out

    def synchronize(self):
        pass

    def close(self):
        self.closed = True

    # operators (oracle arithmetic)
    def _case(self):
        assert self._hydro is not None, "hydrostatic profiles not set"
        c = no.OracleCase(self.nx, self.nz, self.dx, self.dz, self.dt, self.buf[STATE], self.buf[TMP], *self._hydro)
        c.source_w = self._source
        if self.inflow is not None:
            # the mask came from _dispatch.inflow_row_mask(params): recover zlen from it is not possible,
            # so the stand-in applies the mask itself
            c.inflow_zlen = None
        c.reverse_direction = self.reverse_direction
        return c

    def bc_x(self, buf):
        s, nx = self.buf[buf], self.nx
        rows = slice(2, self.nz + 2)
        s[:, rows, 0] = s[:, rows, nx]
        s[:, rows, 1] = s[:, rows, nx + 1]
        if self.inflow is None:
            s[:, rows, nx + 2] = s[:, rows, 2]
            s[:, rows, nx + 3] = s[:, rows, 3]
        else:
            mask, u_in, th_in = self.inflow
            idx = np.nonzero(mask)[0] + 2
            hd, hdt = self._hydro[0], self._hydro[1]
            for col in (0, 1):
                s[1, idx, col] = (s[0, idx, col] + hd[idx]) * u_in
            for col in (0, 1):
                s[3, idx, col] = (s[0, idx, col] + hd[idx]) * th_in - hdt[idx]
        self.launch_count += 1

    def bc_z(self, buf):
        no.set_bc_z(self._case(), self.buf[buf])
        self.launch_count += 1

    def stage(self, d, init_buf, forcing_buf, out_buf, dt_stage):
        c = self._case()
        f = self.buf[forcing_buf]
        if d == no.DIR_X:
            vals, d3 = no.interpolate_x(c, f)
            tend = no.compute_tend_x(c, no.compute_flux_x(c, vals, d3))
        else:
            vals, d3 = no.interpolate_z(c, f)
            tend = no.compute_tend_z(c, no.compute_flux_z(c, vals, d3), f)
        if self._source is not None:
            tend[2] += self._source
        self.buf[out_buf][:, 2:-2, 2:-2] = self.buf[init_buf][:, 2:-2, 2:-2] + dt_stage * tend
        self.launch_count += 1

    def discrete_step(self, d, init_buf, forcing_buf, out_buf, dt_stage):
        (self.bc_x if d == no.DIR_X else self.bc_z)(forcing_buf)
        self.stage(d, init_buf, forcing_buf, out_buf, dt_stage)

    def evolve(self, nsteps=1, dt=None):
        dt = self.dt if dt is None or dt <= 0 else dt
        for _ in range(nsteps):
            for d in ((no.DIR_X, no.DIR_Z) if self.reverse_direction else (no.DIR_Z, no.DIR_X)):
                self.discrete_step(d, STATE, STATE, TMP, dt / 3)
                self.discrete_step(d, STATE, TMP, TMP, dt / 2)
                self.discrete_step(d, STATE, TMP, STATE, dt / 1)
            self.reverse_direction = not self.reverse_direction

    def evolve_host(self, host, dt=None, nbands=0):
        # pmw_evolve_host: one upload, one step and one download of the state (streamed in bands on the device)
        self.upload(STATE, host)
        self.evolve(1, dt)
        self.download(STATE, out=host)
        self.host_steps = getattr(self, "host_steps", 0) + 1

    def stats(self, buf=STATE):
        return no.compute_stats(self._case(), self.buf[buf])

    def solution_variables(self, buf=STATE):
        return no.compute_solution_variables(self._case(), self.buf[buf])


@pytest.fixture()
def fake_device(monkeypatch):
    import pyminiweather_b200._dispatch as dispatch
    import pyminiweather_b200.data.fields as fields_mod
    OracleSolver.instances = []
    monkeypatch.setattr(fields_mod, "DeviceSolver", OracleSolver)
    monkeypatch.setattr(dispatch, "DeviceSolver", OracleSolver)
    dispatch._foreign.clear()
    yield OracleSolver
    dispatch._foreign.clear()


def native(nx, nz, ic):
    from pyminiweather_b200.data import initialize_fields
    from pyminiweather_b200.ics import init
    from pyminiweather_b200.mesh import MeshData
    p = make_params(nx, nz, ic)
    f = initialize_fields(p)
    m = MeshData(p)
    init(f, p, m)
    return p, f, m


def foreign(case):
    f = types.SimpleNamespace(state=case.state.copy(), state_tmp=case.state_tmp.copy(), nvariables=4)
    for n in HYDRO:
        setattr(f, n, getattr(case, n).copy())
    return f


@pytest.mark.parametrize("ic", ["thermal", "gravity", "injection"])
def test_device_resident_loop_moves_data_only_when_asked(fake_device, ic):
    """A loop of evolve() on our Fields: two uploads before the first step, nothing afterwards; reading
    fields.state pulls once; writing through the returned array pushes it back before the next operator."""
    from pyminiweather_b200.post import compute_stats
    from pyminiweather_b200.solve import evolve
    p, f, mesh = native(24, 16, ic)
    _, case = new_case(24, 16, ic)
    for _ in range(5):
        evolve(p, f, mesh, dt=p["dt"])
        no.evolve(case)
    dev = fake_device.instances[-1]
    assert sorted(dev.uploads) == [STATE, TMP] and dev.downloads == []
    assert (dev._source is not None) == (ic == "gravity") and (dev.inflow is not None) == (ic == "injection")
    assert compute_stats(p, f) == no.compute_stats(case)          # still no transfer
    assert dev.downloads == []
    assert np.array_equal(f.state, case.state) and dev.downloads == [STATE]
    assert np.array_equal(f.state, case.state) and dev.downloads == [STATE]   # cached until the device writes again
    f.state[3, 5, 7] += 1.0                                        # the caller edits the host view ...
    case.state[3, 5, 7] += 1.0
    evolve(p, f, mesh, dt=p["dt"])                                 # ... which is pushed before the next operator
    no.evolve(case)
    assert dev.uploads.count(STATE) == 2
    assert np.array_equal(f.state, case.state) and np.array_equal(f.state_tmp, case.state_tmp)
    f.close()
    assert dev.closed


@pytest.mark.parametrize("ic", ["collision", "injection"])
def test_strict_dropin_transfers_every_call(fake_device, ic):
    from pyminiweather_b200.post import compute_solution_variables, compute_stats
    from pyminiweather_b200.solve import evolve
    p, case = new_case(20, 12, ic)
    ff = foreign(case)
    for n in range(3):
        evolve(p, ff, None, dt=p["dt"])
        no.evolve(case)
        assert np.array_equal(ff.state[:, 2:-2, 2:-2], case.state[:, 2:-2, 2:-2])
    dev = fake_device.instances[-1]
    assert len(fake_device.instances) == 1                          # one context, cached per container
    assert dev.uploads.count(STATE) == 3 and dev.downloads.count(STATE) == 3
    assert dev.uploads.count(TMP) == (3 if ic == "injection" else 0)   # state_tmp halos are caller data there
    assert getattr(dev, "host_steps", 0) == (0 if ic == "injection" else 3)  # the streamed step where it applies
    assert compute_stats(p, ff) == no.compute_stats(case)
    assert np.array_equal(compute_solution_variables(p, ff), no.compute_solution_variables(case))
    # another grid for the same container -> a new context, the old one is closed
    p2, case2 = new_case(16, 12, ic)
    ff.state, ff.state_tmp = case2.state.copy(), case2.state_tmp.copy()
    evolve(p2, ff, None, dt=p2["dt"])
    assert len(fake_device.instances) == 2 and fake_device.instances[0].closed


def test_discrete_step_aliasing_patterns_on_native_and_arbitrary_arrays(fake_device):
    """step.py:112-141: (init, forcing, out) = (S,S,T), (S,T,T), (S,T,S) on the fields' own arrays, and the
    generic path for arrays that belong to nobody."""
    from pyminiweather_b200.solve import discrete_step
    p, f, mesh = native(18, 14, "collision")
    _, case = new_case(18, 14, "collision")
    for d in (no.DIR_Z, no.DIR_X):
        st, tmp = f.state, f.state_tmp
        discrete_step(p, f, mesh, st, st, tmp, p["dt"] / 3, d)
        no.discrete_step(case, case.state, case.state, case.state_tmp, case.dt / 3, d)
        discrete_step(p, f, mesh, st, tmp, tmp, p["dt"] / 2, d)
        no.discrete_step(case, case.state, case.state_tmp, case.state_tmp, case.dt / 2, d)
        discrete_step(p, f, mesh, st, tmp, st, p["dt"], d)
        no.discrete_step(case, case.state, case.state_tmp, case.state, case.dt, d)
        assert np.array_equal(f.state, case.state) and np.array_equal(f.state_tmp, case.state_tmp)
    rng = np.random.default_rng(1)
    a = case.state + 1e-3 * rng.standard_normal(case.state.shape)
    b, out = a.copy(), np.zeros_like(a)
    a2, b2, out2 = a.copy(), b.copy(), out.copy()
    discrete_step(p, f, mesh, a, b, out, p["dt"], no.DIR_X)
    no.discrete_step(case, a2, b2, out2, case.dt, no.DIR_X)
    assert np.array_equal(out[:, 2:-2, 2:-2], out2[:, 2:-2, 2:-2]) and np.array_equal(b, b2)   # forcing got its halos
    assert np.array_equal(f.state, case.state)                      # the fields' own buffers were restored
    with pytest.raises(ValueError):
        discrete_step(p, f, mesh, a[:, :, :-1], b, out, p["dt"], no.DIR_X)
    with pytest.raises(ValueError):
        discrete_step(p, f, mesh, a, b, out, p["dt"], 3)


def test_set_bc_x_branches_on_its_ic_type_argument(fake_device):
    from pyminiweather_b200.ics import set_bc_x, set_bc_z
    p, f, mesh = native(20, 12, "injection")
    _, case = new_case(20, 12, "injection")
    rng = np.random.default_rng(2)
    a = rng.standard_normal(f.state.shape)
    want_inj, want_per, want_z = a.copy(), a.copy(), a.copy()
    no.set_bc_x(case, want_inj)
    periodic = case.copy(); periodic.inflow_zlen = None
    no.set_bc_x(periodic, want_per)
    no.set_bc_z(case, want_z)
    b = a.copy(); set_bc_x(p, f, b, "injection"); assert np.array_equal(b, want_inj)
    b = a.copy(); set_bc_x(p, f, b, "thermal"); assert np.array_equal(b, want_per)
    b = a.copy(); set_bc_x(p, f, b, "injection"); assert np.array_equal(b, want_inj)   # and back
    b = a.copy(); set_bc_z(p, f, b, "injection"); assert np.array_equal(b, want_z)
    ff = foreign(case)
    b = a.copy(); set_bc_x(p, ff, b, "injection"); assert np.array_equal(b, want_inj)
    with pytest.raises(ValueError, match="unknown ic_type"):
        set_bc_x(p, f, a, "squall-line")


def test_fields_container_contract(fake_device):
    """Attribute names / shapes of data/fields.py:7-55, lazy scratch arrays, assignment, grid mismatch."""
    from pyminiweather_b200.data import Fields, initialize_fields
    p = make_params(10, 6, "thermal")
    f = initialize_fields(p)
    assert f.state.shape == f.state_tmp.shape == (4, 10, 14) and f.hy_dens_cell.shape == (10,)
    assert f.hy_dens_int.shape == f.hy_pressure_int.shape == (7,)
    assert f.flux.shape == (4, 7, 11) and f.tend.shape == (4, 6, 10) and f.vals_x.shape == (4, 6, 11)
    assert f.vals_z.shape == (4, 7, 10) and f.flux is f.flux
    with pytest.raises(AttributeError):
        f.no_such_array
    with pytest.raises(ValueError):
        f.state = np.zeros((4, 10, 13))
    with pytest.raises(AssertionError):
        Fields(10, 6, hs=2, s=5)                                    # fields.py:65
    with pytest.raises(ValueError, match="do not match"):
        f.device(make_params(12, 6, "thermal"))
