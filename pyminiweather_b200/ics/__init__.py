from .bcs import set_bc_x, set_bc_z
from .directions import Directions
from .initial import init, init_device

__all__ = ["set_bc_x", "set_bc_z", "Directions", "init", "init_device"]
