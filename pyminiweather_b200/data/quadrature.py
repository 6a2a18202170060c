"""3-point Gauss-Legendre rule on [0, 1] used by the initial conditions
(mirror of pyminiweather/data/quadrature.py:4-74).  Init-time only."""
import numpy as np


class GaussianQuadrature:
    __slots__ = ("_pts", "_wts", "_outer", "_gx", "_gz", "_wx", "_wz")

    def __init__(self):
        self._pts = np.array([0.112701665379258311482073460022, 0.5, 0.887298334620741688517926539980])
        self._wts = np.array([0.277777777777777777777777777779, 0.444444444444444444444444444444,
                              0.277777777777777777777777777779])
        for a in (self._pts, self._wts):
            a.setflags(write=False)
        self._outer = np.outer(self._wts, self._wts)
        self._gx, self._gz = np.meshgrid(self._pts, self._pts)
        self._wx, self._wz = np.meshgrid(self._wts, self._wts)

    npoints = property(lambda self: 3)
    qpoints = property(lambda self: self._pts)
    qweights = property(lambda self: self._wts)
    qweights_outer = property(lambda self: self._outer)
    qpoints_grid_x = property(lambda self: self._gx)
    qpoints_grid_z = property(lambda self: self._gz)
    qweights_grid_x = property(lambda self: self._wx)
    qweights_grid_z = property(lambda self: self._wz)


Quadrature = GaussianQuadrature()
