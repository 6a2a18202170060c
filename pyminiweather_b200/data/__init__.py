from .constants import Constants
from .fields import Fields, initialize_fields
from .quadrature import Quadrature

__all__ = ["Constants", "Fields", "initialize_fields", "Quadrature"]
