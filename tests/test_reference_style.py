"""The reference's own unit tests, restated against OUR modules (tests/unit/*.py of the
reference): constants, quadrature, mesh, initial conditions, bubble sampler, and the
interpolation-vs-convolve2d check (there against the oracle's interpolation; the GPU shims are
checked in test_gpu_parity.py)."""
import numpy as np
import pytest
from scipy.signal import convolve2d

from conftest import golden
from helpers import HYDRO, make_params
from oracle import numpy_oracle as no
from pyminiweather_b200.data import Constants, Quadrature, initialize_fields
from pyminiweather_b200.ics import init
from pyminiweather_b200.ics.initial_conditions import hydro_const_bvfreq, hydro_const_theta
from pyminiweather_b200.mesh import MeshData
from pyminiweather_b200.utils import sample_ellipse_cosine


def test_constants():  # tests/unit/test_constants.py:7-19
    assert Constants.pi.value == 3.14159265358979323846264338327
    assert Constants.grav.value == 9.8 and Constants.cp.value == 1004.0 and Constants.cv.value == 717.0
    assert Constants.rd.value == 287.0 and Constants.p0.value == 1.0e5
    assert Constants.C0.value == 27.5629410929725921310572974482
    assert Constants.gamma.value == 1.40027894002789400278940027894
    assert Constants.hv_beta.value == 0.05 and Constants.theta0.value == 300.0 and Constants.exner0.value == 1.0


def test_quadrature():  # tests/unit/test_constants.py:22-64
    assert Quadrature.npoints == 3
    np.testing.assert_array_equal(Quadrature.qpoints, [0.112701665379258311482073460022, 0.5,
                                                       0.887298334620741688517926539980])
    np.testing.assert_array_equal(Quadrature.qweights, [0.277777777777777777777777777779,
                                                        0.444444444444444444444444444444,
                                                        0.277777777777777777777777777779])
    with pytest.raises(ValueError):
        Quadrature.qpoints[0] = 1.0
    with pytest.raises(AttributeError):
        Quadrature.qpoints = np.zeros(3)


def test_mesh():  # tests/unit/test_mesh.py:18-82
    p = make_params(64, 32)
    m = MeshData(p)
    x, z = m.get_mesh_int_ext()
    assert x.shape == z.shape == (36, 68)
    np.testing.assert_array_equal(x[0], np.linspace(-2 * p["dx"], (64 + 2) * p["dx"], 68, endpoint=False))
    np.testing.assert_array_equal(z[:, 0], np.linspace(-2 * p["dz"], (32 + 2) * p["dz"], 36, endpoint=False))
    xc, zc = m.get_mesh_cell_centers()
    np.testing.assert_array_equal(xc[0], np.linspace(p["dx"] / 2, p["xlen"] + p["dx"] / 2, 64, endpoint=False))
    np.testing.assert_array_equal(m.get_mesh_vertical_cell_edges(), np.linspace(0.0, 33 * p["dz"], 33, endpoint=False))
    np.testing.assert_array_equal(m.get_mesh_vertical_cell_centers_int_ext(),
                                  np.linspace(-1.5 * p["dz"], 34.5 * p["dz"], 36, endpoint=False))
    assert m.get_mesh_int_ext() is m.get_mesh_int_ext()  # cached


def test_thermal_initial_condition():  # tests/unit/test_initial_conditions.py:13-41
    p = make_params(200, 100, "thermal")
    f = initialize_fields(p)
    init(f, p, MeshData(p))
    s = f.state
    assert not s[0].any() and not s[1].any() and not s[2].any()
    assert np.count_nonzero(s[3]) > 0


def test_hydrostatic_profiles_smooth():  # tests/unit/test_initial_conditions.py:44-82
    z = np.linspace(0, 1e4, 200)
    for hr, ht in (hydro_const_theta(z), hydro_const_bvfreq(z, 0.02)):
        assert np.all(np.diff(hr) < 0)
        assert np.all(np.abs(np.diff(hr, 2)) < 1e-3)
        assert np.all(np.asarray(ht) >= 300.0)


def test_sample_ellipse_cosine():  # tests/unit/test_sample_cosine.py:9-73
    x, z = np.meshgrid(np.linspace(0, 2e4, 41), np.linspace(0, 1e4, 21))
    tiny = sample_ellipse_cosine(x, z, 3.0, 1e4, 5e3, 1.0, 1.0)
    assert np.count_nonzero(tiny) == 1 and tiny.max() == 3.0
    huge = sample_ellipse_cosine(x, z, 3.0, 1e4, 5e3, 1e6, 1e6)
    d = np.sqrt(((x - 1e4) / 1e6) ** 2 + ((z - 5e3) / 1e6) ** 2) * Constants.pi.value / 2
    np.testing.assert_array_equal(huge, 3.0 * np.cos(d) ** 2.0)


@pytest.mark.parametrize("ic", ["thermal", "collision", "density-current", "gravity", "injection"])
def test_init_bit_identical_to_reference(ic):
    g = golden(f"ic_{ic}_32x16.npz")
    p = make_params(32, 16, ic)
    f = initialize_fields(p)
    init(f, p, MeshData(p))
    for n in ("state", "state_tmp") + HYDRO:
        assert np.array_equal(getattr(f, n), g[n]), n


def test_init_row_chunking_is_invisible(monkeypatch):
    import pyminiweather_b200.ics.initial as ini
    p = make_params(40, 24, "collision")
    f1 = initialize_fields(p); init(f1, p, MeshData(p))
    monkeypatch.setattr(ini, "_CHUNK_ELEMS", 9 * 44 * 5)
    f2 = initialize_fields(p); init(f2, p, MeshData(p))
    assert np.array_equal(f1.state, f2.state)


def test_oracle_interpolate_vs_convolve2d():  # tests/unit/test_interpolate.py:11-63 (default 200x100)
    p = make_params(200, 100)
    shape = (4, 104, 204)
    state = np.arange(np.prod(shape)).astype(np.float64).reshape(shape)
    k4 = np.array([-1.0 / 12, 7.0 / 12, 7.0 / 12, -1.0 / 12])
    k3 = np.array([-1.0, 3.0, -3.0, 1.0])  # tests/unit/test_interpolate.py:66-93
    case = no.OracleCase(200, 100, p["dx"], p["dz"], p["dt"], state, state, *[np.zeros(1)] * 5)
    vx, d3x = no.interpolate_x(case, state)
    vz, d3z = no.interpolate_z(case, state)
    for v in range(4):
        assert np.allclose(convolve2d(state[v, 2:102, :], k4[None, :], mode="same")[:, 2:-1], vx[v])
        assert np.allclose(convolve2d(state[v, :, 2:202], k4[:, None], mode="same")[2:-1, :], vz[v])
        assert np.allclose(convolve2d(state[v, 2:102, :], -k3[None, :], mode="same")[:, 2:-1], d3x[v])
        assert np.allclose(convolve2d(state[v, :, 2:202], -k3[:, None], mode="same")[2:-1, :], d3z[v])


# ---- host logic of the injection configuration and of the device-side init (no GPU needed) -----------
@pytest.mark.parametrize("nz,zlen", [(50, 1e4), (64, 1e4), (37, 7.5e3), (1024, 1e4)])
def test_inflow_row_mask_is_the_reference_condition(nz, zlen):
    """_dispatch.inflow_row_mask (what pmw_set_inflow receives) against the oracle's restatement of
    bcs.py:43-48, which is pinned on the reference's fixtures."""
    from pyminiweather_b200._dispatch import inflow_row_mask
    p = make_params(32, nz, "injection", zlen=zlen)
    mask = inflow_row_mask(p)
    assert mask.dtype == np.uint8 and mask.shape == (nz,)
    np.testing.assert_array_equal(np.nonzero(mask)[0] + 2, no.inflow_rows(nz, p["dz"], zlen))
    assert 0 < mask.sum() < nz
    # the band is centred on 3/4 of the domain height and zlen/8 tall
    rows = np.nonzero(mask)[0]
    assert abs((rows.mean() + 0.5) * p["dz"] - 0.75 * zlen) <= p["dz"]
    assert abs(rows.size * p["dz"] - zlen / 8) <= 2 * p["dz"]


def test_check_ic_accepts_the_five_configurations_only():
    from pyminiweather_b200._dispatch import check_ic
    for ic in ("thermal", "collision", "density-current", "gravity", "injection"):
        check_ic(ic)
    with pytest.raises(ValueError, match="unknown ic_type"):
        check_ic("squall-line")


@pytest.mark.parametrize("ic", ["thermal", "collision", "density-current", "gravity", "injection"])
def test_device_spec_describes_the_host_catalogue(ic):
    """ics.device_spec (the pmw_ic_spec pmw_init_state integrates) evaluated with the HOST sampler
    reproduces cell_quantities at arbitrary points: same bubbles, wind and background."""
    from pyminiweather_b200.ics.initial_conditions import background, cell_quantities, device_spec
    xlen = 2e4
    rng = np.random.default_rng(5)
    x, z = rng.uniform(0, xlen, 500), rng.uniform(0, 1e4, 500)
    r, u, w, t, hr, ht = cell_quantities(ic, x, z, xlen)
    bubbles, wind, bv0 = device_spec(ic, xlen)
    t2 = np.zeros_like(x)
    for amp, x0, z0, xrad, zrad in bubbles:
        t2 = t2 + sample_ellipse_cosine(x, z, amp, x0, z0, xrad, zrad)
    assert np.array_equal(t, t2) and np.all(u == wind) and not r.any() and not w.any()
    hr2, ht2 = (hydro_const_bvfreq(z, bv0) if bv0 is not None else hydro_const_theta(z))
    assert np.array_equal(hr, hr2) and np.array_equal(np.broadcast_to(ht, x.shape), np.broadcast_to(ht2, x.shape))
    assert len(bubbles) <= 4


def test_mesh_axes_are_the_meshgrid_rows():
    from pyminiweather_b200.slab import SlabMesh
    p = make_params(48, 20)
    m = MeshData(p)
    xa, za = m.get_axes_int_ext()
    x, z = m.get_mesh_int_ext()
    assert np.array_equal(x[0], xa) and np.array_equal(z[:, 0], za)
    # slabs: each rank's axis is its part of the global one (to rounding of the shifted origin)
    world, nxl = 3, 16
    for r in range(world):
        sa, sz = SlabMesh(dict(p, nx=nxl), r, world).get_axes_int_ext()
        np.testing.assert_allclose(sa, xa[r * nxl: (r + 1) * nxl + 4], rtol=0, atol=1e-9)
        assert np.array_equal(sz, za)


def test_ic_spec_struct_matches_the_header():
    """ctypes mirror of pmw_ic_spec: field order and capacity as declared in include/pmw.h."""
    import os
    import re
    from pyminiweather_b200 import _lib
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "pmw.h")).read()
    body = re.search(r"typedef struct pmw_ic_spec \{(.*?)\} pmw_ic_spec;", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"\b(?:int|double)\s+([^;]+);", body)
    flat = [n.split("[")[0].strip() for decl in names for n in decl.split(",")]
    assert flat == [f[0] for f in _lib.PmwIcSpec._fields_]
    assert int(re.search(r"#define PMW_IC_MAX_BUBBLES (\d+)", hdr).group(1)) == _lib.PMW_IC_MAX_BUBBLES
