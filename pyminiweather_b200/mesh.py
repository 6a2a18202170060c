"""Coordinate arrays of the uniform grid (mirror of pyminiweather/mesh.py:5-112).
Init/plot-time only; results are cached per instance like the reference's."""
from functools import cached_property

import numpy as np


class MeshData:
    def __init__(self, params):
        self.params = params
        self.xlen, self.zlen = params["xlen"], params["zlen"]
        self.dx, self.dz = params["dx"], params["dz"]
        self.nx, self.nz, self.hs = params["nx"], params["nz"], params["hs"]

    def _axis_int_ext(self, n, d):
        return np.linspace(-self.hs * d, (n + self.hs) * d, n + 2 * self.hs, endpoint=False)

    @cached_property
    def _int_ext(self):
        return np.meshgrid(self._axis_int_ext(self.nx, self.dx), self._axis_int_ext(self.nz, self.dz))

    @cached_property
    def _centers(self):
        x = np.linspace(self.dx / 2.0, self.xlen + self.dx / 2.0, self.nx, endpoint=False)
        z = np.linspace(self.dz / 2.0, self.zlen + self.dz / 2.0, self.nz, endpoint=False)
        return np.meshgrid(x, z)

    @cached_property
    def _vertical_edges(self):
        return np.linspace(0.0, (self.nz + 1) * self.dz, self.nz + 1, endpoint=False)

    @cached_property
    def _vertical_centers_int_ext(self):
        return np.linspace((-self.hs + 0.5) * self.dz, (self.nz + self.hs + 0.5) * self.dz,
                           self.nz + 2 * self.hs, endpoint=False)

    def get_axes_int_ext(self):
        """The two 1-D axes get_mesh_int_ext() is the meshgrid of (x[nx+4], z[nz+4]); what the
        device-side init takes instead of the 2-D arrays."""
        return self._axis_int_ext(self.nx, self.dx), self._axis_int_ext(self.nz, self.dz)

    def get_mesh_int_ext(self):
        """(x, z) 2-D arrays over interior + ghost cells (lower-left corners)."""
        return self._int_ext

    def get_mesh_cell_centers(self):
        """(x, z) 2-D arrays of interior cell centres."""
        return self._centers

    def get_mesh_vertical_cell_edges(self):
        """z of the nz+1 interior cell edges."""
        return self._vertical_edges

    def get_mesh_vertical_cell_centers_int_ext(self):
        """z of the nz+4 cell centres including ghosts."""
        return self._vertical_centers_int_ext
