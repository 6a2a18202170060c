"""Sweep directions (mirror of pyminiweather/ics/directions.py:4-6; the integer values are
also the C ABI's PMW_DIR_X / PMW_DIR_Z)."""
from enum import Enum


class Directions(Enum):
    X = 1
    Z = 2
