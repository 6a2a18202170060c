"""Extra tendency sources (mirror of pyminiweather/solve/source.py:20-75).

Only ``ic_type == "gravity"`` has one: a constant squared-cosine bump ``wpert`` (amplitude 0.01,
centre (xlen/8, 1000 m), radii 500 m) times the hydrostatic density of the row, added to the rho*w
tendency in EVERY stage of both sweeps (source.py:43-50, called at step.py:78).  The reference
re-samples the bump on every call; here the field is built once on the host and lives in HBM, and
the fused stage kernels add it in registers (``StageArgs::src_w``).
"""
from __future__ import annotations

import numpy as np

from ..mesh import MeshData
from ..utils import sample_ellipse_cosine


def gravity_source_field(params, hy_dens_cell, mesh=None) -> np.ndarray:
    """[nz, nx]: wpert(x, z) * hy_dens_cell[2:nz+2, None] at the interior cell centres."""
    nz = params["nz"]
    x, z = (mesh or MeshData(params)).get_mesh_cell_centers()
    wpert = sample_ellipse_cosine(x, z, 0.01, params["xlen"] / 8, 1000.0, 500.0, 500.0)
    return wpert * np.asarray(hy_dens_cell)[2:nz + 2, np.newaxis]


def source_field_for(params, hy_dens_cell):
    """The forcing field the device context needs for this configuration, or None."""
    if params["ic_type"] == "gravity":
        return gravity_source_field(params, hy_dens_cell)
    return None


def add_source_terms(params, mesh, fields):
    """fields.tend[WMOM] += source (source.py:53-75) on the host scratch array -- for code that
    drives the unfused operators; ``evolve``/``discrete_step`` apply the source inside the kernels."""
    if params["ic_type"] == "gravity":
        fields.tend[2, :, :] += gravity_source_field(params, fields.hy_dens_cell, mesh)
