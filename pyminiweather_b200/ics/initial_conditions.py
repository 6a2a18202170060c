"""Initial-condition catalogue (mirror of pyminiweather/ics/initial_conditions.py:26-429).

Each ``--ic-type`` is a hydrostatic background (constant potential temperature, or constant
Brunt-Vaisala frequency for ``gravity``) plus a list of potential-temperature bubbles and a
uniform wind.  Host NumPy, init-time only; the arithmetic follows the reference expression
by expression so that the generated fields are bit-identical (tests/test_reference_style.py against the ic_*_32x16 fixtures).
"""
from __future__ import annotations

import numpy as np

from ..data.constants import Constants
from ..utils import sample_ellipse_cosine

C = Constants


def hydro_const_theta(z):
    """Background (density, potential temperature) for theta = theta0
    (initial_conditions.py:26-52)."""
    ht = C.theta0.value
    exner = C.exner0.value - C.grav.value * z / (C.cp.value * C.theta0.value)
    p = C.p0.value * (exner ** (C.cp.value / C.rd.value))
    hr = ((p / C.C0.value) ** (1.0 / C.gamma.value)) / ht
    return hr, ht


def hydro_const_bvfreq(z, bv_freq0):
    """Background for a constant Brunt-Vaisala frequency (initial_conditions.py:54-81)."""
    ht = C.theta0.value * np.exp(bv_freq0 * bv_freq0 / C.grav.value * z)
    exner = C.exner0.value - C.grav.value * C.grav.value / (C.cp.value * bv_freq0 * bv_freq0) * (
        ht - C.theta0.value) / (ht * C.theta0.value)
    p = C.p0.value * (exner ** (C.cp.value / C.rd.value))
    hr = ((p / C.C0.value) ** (1.0 / C.gamma.value)) / ht
    return hr, ht


# ic_type -> (bubbles [(amplitude, z0, xrad, zrad)], uniform u, background)
#   thermal          initial_conditions.py:166-199
#   collision        initial_conditions.py:96-134
#   density-current  initial_conditions.py:236-268
#   gravity          initial_conditions.py:202-233   (bv_freq0 = 0.02, u = 15)
#   injection        initial_conditions.py:271-302
_CATALOGUE = {
    "thermal": ([(3.0, 2000.0, 2000.0, 2000.0)], 0.0, "theta"),
    "collision": ([(20.0, 2000.0, 2000.0, 2000.0), (-20.0, 8000.0, 2000.0, 2000.0)], 0.0, "theta"),
    "density-current": ([(-20.0, 5000.0, 4000.0, 2000.0)], 0.0, "theta"),
    "gravity": ([], 15.0, "bvfreq"),
    "injection": ([], 0.0, "theta"),
}
IC_TYPES = tuple(_CATALOGUE)
_BV0 = 0.02


def background(ic_type, z):
    """(hr, ht) at heights z -- what VCEQInitFactory(ic)(z) returns (initial_conditions.py:333-367)."""
    if _CATALOGUE[ic_type][2] == "bvfreq":
        return hydro_const_bvfreq(z, _BV0)
    return hydro_const_theta(z)


def cell_quantities(ic_type, x, z, xlen):
    """(r, u, w, t, hr, ht) at points (x, z) -- what CCQInitFactory(ic)(x, z, xlen) returns
    (initial_conditions.py:370-429)."""
    bubbles, wind, _ = _CATALOGUE[ic_type]
    hr, ht = background(ic_type, z)
    r = np.zeros(x.shape)
    w = np.zeros(x.shape)
    u = wind * np.ones(x.shape) if wind else np.zeros(x.shape)
    t = np.zeros(x.shape)
    for amp, z0, xrad, zrad in bubbles:
        t = t + sample_ellipse_cosine(x, z, amp, xlen / 2, z0, xrad, zrad)
    return r, u, w, t, hr, ht


def device_spec(ic_type, xlen):
    """The same catalogue entry in the form ``pmw_init_state`` takes: ([(amp, x0, z0, xrad, zrad)],
    uniform u, bv0 or None).  Every bubble is centred at xlen/2 (initial_conditions.py:131-132,197,266)."""
    bubbles, wind, kind = _CATALOGUE[ic_type]
    return ([(amp, xlen / 2, z0, xrad, zrad) for amp, z0, xrad, zrad in bubbles], wind,
            _BV0 if kind == "bvfreq" else None)
