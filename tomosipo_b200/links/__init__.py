"""Array back-ends of the projector (numpy always; torch / cupy when installed)."""
import importlib.util

from . import base
from . import numpy
from .base import are_compatible, geometry_shape

if importlib.util.find_spec("torch") is not None:
    from . import torch

if importlib.util.find_spec("cupy") is not None:
    from . import cupy
