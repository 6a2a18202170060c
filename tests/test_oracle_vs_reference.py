"""Live cross-check of the oracles against the real reference -- only where /root/reference is
mounted (the build container).  On the GPU box these tests skip; the golden fixtures cover it."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import numpy_oracle as no, reference_runner as rr

pytestmark = pytest.mark.skipif(not rr.available(), reason="/root/reference not mounted")


@pytest.mark.parametrize("ic", ["thermal", "collision", "density-current", "injection"])
def test_numpy_oracle_bit_identical_to_reference(ic):
    ref = rr.ReferenceRun(96, 40, ic)
    case = ref.to_oracle_case()
    assert no.compute_stats(case) == ref.stats()
    for n in range(12):
        ref.evolve(1)
        no.evolve(case)
    assert np.array_equal(ref.fields.state, case.state)          # halos included
    assert np.array_equal(ref.fields.state_tmp, case.state_tmp)
    assert no.compute_stats(case) == ref.stats()


def test_oracle_interpolation_matches_reference_arrays():
    ref = rr.ReferenceRun(50, 30, "collision")
    ref.evolve(2)
    case = ref.to_oracle_case()
    from pyminiweather.solve import interpolate_x, interpolate_z  # reference
    interpolate_x(ref.params, ref.fields, ref.fields.state)
    interpolate_z(ref.params, ref.fields, ref.fields.state)
    vx, dx3 = no.interpolate_x(case, case.state)
    vz, dz3 = no.interpolate_z(case, case.state)
    assert np.array_equal(vx, ref.fields.vals_x) and np.array_equal(dx3, ref.fields.d3_vals_x)
    assert np.array_equal(vz, ref.fields.vals_z) and np.array_equal(dz3, ref.fields.d3_vals_z)


def test_reference_own_unit_tests_pass_with_numpy_shim(tmp_path):
    """The reference's tests import cupynumeric; with a shim that re-exports NumPy (and SciPy's
    convolve) they pin the 4th-order interpolation against scipy.signal.convolve2d."""
    shim = tmp_path / "cupynumeric.py"
    shim.write_text(textwrap.dedent("""
        from numpy import *          # noqa
        from scipy.signal import convolve  # noqa
    """))
    env = dict(os.environ, PYTHONPATH=f"{tmp_path}:{rr.REFERENCE_ROOT}", PYTHONDONTWRITEBYTECODE="1")
    for k in ("LEGATE_MAX_DIM", "LEGATE_MAX_FIELDS"):
        env.pop(k, None)
    res = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider",
                          os.path.join(rr.REFERENCE_ROOT, "tests/unit/test_interpolate.py"),
                          os.path.join(rr.REFERENCE_ROOT, "tests/unit/test_constants.py")],
                         env=env, capture_output=True, text=True, cwd=str(tmp_path))
    assert res.returncode == 0, res.stdout + res.stderr


def test_gravity_source_restatement_matches_reference():
    ref = rr.ReferenceRun(64, 32, "gravity")
    case = ref.to_oracle_case()
    from pyminiweather.utils import sample_ellipse_cosine  # reference
    x, z = ref.mesh.get_mesh_cell_centers()
    want = sample_ellipse_cosine(x, z, 0.01, ref.params["xlen"] / 8, 1000.0, 500.0, 500.0) * \
        ref.fields.hy_dens_cell[2:34, None]
    assert np.array_equal(case.source_w, want)
    from helpers import make_params
    from pyminiweather_b200.solve.source import gravity_source_field
    assert np.array_equal(gravity_source_field(make_params(64, 32, "gravity"), ref.fields.hy_dens_cell), want)
    for _ in range(6):
        ref.evolve(1)
        no.evolve(case)
    assert np.array_equal(ref.fields.state, case.state)


def test_ref_copy_is_verbatim_and_git_ignored():
    """oracle/_ref (made by oracle/make_ref.py) is what the GPU box uses as THE reference: byte for byte the mounted
    package, and never part of the history."""
    import filecmp
    from oracle import make_ref
    mount = os.path.join(os.environ.get("PMW_REFERENCE_ROOT", "/root/reference"), "pyminiweather")
    if not os.path.isdir(mount):
        pytest.skip("no mounted reference to compare the copy with")
    dst = make_ref.make()
    assert dst and os.path.isdir(os.path.join(dst, "pyminiweather"))

    def walk(cmp):
        assert not cmp.left_only and not cmp.diff_files, (cmp.left_only, cmp.diff_files)
        for sub in cmp.subdirs.values():
            walk(sub)
    walk(filecmp.dircmp(mount, os.path.join(dst, "pyminiweather"), ignore=["__pycache__"]))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run(["git", "-C", root, "check-ignore", "oracle/_ref/pyminiweather/__init__.py"], capture_output=True, text=True)
    assert out.returncode == 0, "oracle/_ref must be listed in .gitignore"
    ignore = os.path.join(root, ".gpurunignore")
    assert not os.path.exists(ignore) or "oracle/_ref" not in open(ignore).read()
