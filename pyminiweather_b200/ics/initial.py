"""``init``: fill state and the hydrostatic profiles by 3x3 Gauss-Legendre quadrature
(mirror of pyminiweather/ics/initial.py:11-105).

Runs once on the host and produces the hot path's inputs.  Unlike the reference, which
materialises [nz+4, nx+4, 3, 3] temporaries for the whole domain (9x the state), rows are
processed in chunks so that the BASELINE grids (up to 32768 x 8192) fit in host memory.
The per-element arithmetic and summation order are the reference's, so results are
bit-identical.
"""
from __future__ import annotations

import numpy as np

from ..data.constants import Constants
from ..data.quadrature import Quadrature
from .initial_conditions import IC_TYPES, background, cell_quantities, device_spec

_CHUNK_ELEMS = 1 << 21  # quadrature points per chunk (x 8 B x ~12 temporaries)


def init(fields, params, Mesh):
    xlen, dx, dz = params["xlen"], params["dx"], params["dz"]
    nx, nz, hs = params["nx"], params["nz"], params["hs"]
    ic_type = params["ic_type"]
    assert ic_type in IC_TYPES  # initial.py:41-47
    C0, gamma = Constants.C0.value, Constants.gamma.value

    x, z = Mesh.get_mesh_int_ext()  # each [nz+4, nx+4]
    W = Quadrature.qweights_outer
    state = fields.state  # host array (marks it modified)
    rows_per_chunk = max(1, _CHUNK_ELEMS // (9 * (nx + 2 * hs)))
    for k0 in range(0, nz + 2 * hs, rows_per_chunk):
        sl = slice(k0, min(nz + 2 * hs, k0 + rows_per_chunk))
        xq = x[sl, :, np.newaxis, np.newaxis] + Quadrature.qpoints_grid_x * dx
        zq = z[sl, :, np.newaxis, np.newaxis] + Quadrature.qpoints_grid_z * dz
        r, u, w, t, hr, ht = cell_quantities(ic_type, xq, zq, xlen)
        state[0, sl] = np.multiply(r, W).sum(axis=-1).sum(axis=-1)
        state[1, sl] = np.multiply((r + hr) * u, W).sum(axis=-1).sum(axis=-1)
        state[2, sl] = np.multiply((r + hr) * w, W).sum(axis=-1).sum(axis=-1)
        state[3, sl] = np.multiply((r + hr) * (t + ht) - hr * ht, W).sum(axis=-1).sum(axis=-1)
    fields.state_tmp[:] = state[:]  # initial.py:80
    _init_profiles(fields, ic_type, Mesh)


def _init_profiles(fields, ic_type, Mesh):
    C0, gamma = Constants.C0.value, Constants.gamma.value
    # cell-centre profiles over interior + ghosts (initial.py:84-95)
    hr, ht = background(ic_type, Mesh.get_mesh_vertical_cell_centers_int_ext())
    fields.hy_dens_cell[:] = hr * Quadrature.qweights.sum()
    fields.hy_dens_theta_cell[:] = np.multiply((ht * hr)[:, np.newaxis], Quadrature.qweights).sum(axis=-1)
    # interface profiles (initial.py:100-105)
    hr, ht = background(ic_type, Mesh.get_mesh_vertical_cell_edges())
    fields.hy_dens_int[:] = hr[:]
    fields.hy_dens_theta_int[:] = hr * ht
    fields.hy_pressure_int[:] = C0 * ((hr * ht) ** gamma)


def init_device(fields, params, Mesh):
    """``init`` with the 2-D quadrature done on the GPU (``init_state_kernel``, csrc/pmw_init.cuh):
    same arguments and effect as ``init`` for our device-resident ``Fields`` -- state and state_tmp
    end up in HBM (the host arrays are refreshed when read), the 1-D profiles are computed on the
    host exactly as above.  Nothing of size [nz+4, nx+4, 3, 3] is ever built (the reference needs
    9x the state for it, README.md:156), so the largest BASELINE grids initialise in milliseconds.
    Agreement with ``init``: <= 1e-13 relative L2 (CUDA vs NumPy pow/cos rounding)."""
    from ..data.fields import Fields
    if not isinstance(fields, Fields):
        raise TypeError("init_device needs pyminiweather_b200.data.Fields (device-resident); use init() otherwise")
    ic_type = params["ic_type"]
    assert ic_type in IC_TYPES  # initial.py:41-47
    _init_profiles(fields, ic_type, Mesh)
    bubbles, wind, bv0 = device_spec(ic_type, params["xlen"])
    x_axis, z_axis = Mesh.get_axes_int_ext()
    fields.init_on_device(params, bubbles, wind, bv0, x_axis, z_axis)
