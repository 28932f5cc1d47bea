"""Importing this module enables CuPy arrays as operator inputs
(API mirror of ``tomosipo/cupy.py``)."""
from .links import cupy as _cupy_link  # noqa: F401
