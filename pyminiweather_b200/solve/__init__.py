from .interpolate import (compute_flux_x, compute_flux_z, compute_tend_x, compute_tend_z, interpolate_x,
                          interpolate_z)
from .source import add_source_terms
from .step import discrete_step, evolve

__all__ = ["compute_flux_x", "compute_flux_z", "compute_tend_x", "compute_tend_z", "interpolate_x",
           "interpolate_z", "add_source_terms", "discrete_step", "evolve"]
