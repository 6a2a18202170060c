"""Squared-cosine elliptical bubble (mirror of pyminiweather/utils/utils.py:5-51). Init-time only."""
import numpy as np

from ..data.constants import Constants


def sample_ellipse_cosine(x, z, amplitude, x0, z0, xrad, zrad):
    """amplitude * cos^2(pi/2 * r) inside the ellipse r <= 1, r the scaled distance to
    (x0, z0); zero outside."""
    half_pi = Constants.pi.value / 2.0
    dist = np.sqrt(((x - x0) / xrad) ** 2 + ((z - z0) / zrad) ** 2) * Constants.pi.value / 2.0
    out = np.zeros(np.shape(z), dtype=np.asarray(x).dtype)
    np.putmask(out, dist <= half_pi, amplitude * (np.cos(dist) ** 2.0))
    return out
