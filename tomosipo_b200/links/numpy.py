"""NumPy link (host arrays).  API mirror of ``tomosipo/links/numpy.py``."""
import warnings
from contextlib import contextmanager

import numpy as np

from .. import _backend
from .base import Link, RawBuffer, backends


def _pinned_like(shape, dtype, fill=None):
    """Array for data this link creates itself: page-locked when large (float32 only), see _backend.pinned_empty."""
    if np.dtype(dtype) != np.float32:
        return np.empty(shape, dtype=dtype) if fill is None else np.full(shape, fill, dtype=dtype)
    out = _backend.pinned_empty(shape, np.float32)
    if fill is not None:
        out[...] = fill
    return out


class NumpyLink(Link):
    """Wraps a C-contiguous float32 ``ndarray``; other inputs are converted with a warning."""

    def __init__(self, shape, initial_value):
        super().__init__(shape, initial_value)
        if initial_value is None:
            self._data = np.zeros(shape, dtype=np.float32)
            return
        if np.isscalar(initial_value):
            self._data = np.full(shape, initial_value, dtype=np.float32)
            return
        arr = np.asarray(initial_value)  # NumPy 2: the reference's np.array(copy=False) raises here
        if arr.shape != tuple(shape):
            raise ValueError(f"Cannot link array. Expected array of shape {shape}. Got {arr.shape}")
        if arr.dtype != np.float32:
            warnings.warn(
                f"The parameter initial_value is of type {arr.dtype}; expected `np.float32`. "
                f"The type has been Automatically converted. "
                f"Use `ts.link(x.astype(np.float32))' to inhibit this warning. "
            )
            conv = _pinned_like(arr.shape, np.float32)
            conv[...] = arr
            arr = conv
        if not (arr.flags["C_CONTIGUOUS"] and arr.flags["ALIGNED"]):
            warnings.warn(
                f"The parameter initial_value should be C_CONTIGUOUS and ALIGNED. "
                f"It has been automatically made contiguous and aligned. "
                f"Use `ts.link(np.ascontiguousarray(x))' to inhibit this warning. "
            )
            arr = np.ascontiguousarray(arr)
        self._data = arr

    @staticmethod
    def __accepts__(initial_value):
        # the default backend: also takes None and scalars
        return initial_value is None or isinstance(initial_value, np.ndarray) or np.isscalar(initial_value)

    def __compatible_with__(self, other):
        return True if isinstance(other, NumpyLink) else NotImplemented

    @property
    def linked_data(self):
        return RawBuffer(self._data.ctypes.data, tuple(self._data.shape), "host", 0, 0, self._data)

    @property
    def data(self):
        """The shared ndarray; projection data is ordered (v, angle, u)."""
        return self._data

    @data.setter
    def data(self, val):
        raise AttributeError(
            "You cannot change which array backs a dataset.\n"
            "To change the underlying data instead, use: \n"
            " >>> x.data[:] = new_data\n"
        )

    @contextmanager
    def context(self):
        yield

    def new_zeros(self, shape):
        return NumpyLink(shape, _pinned_like(shape, self._data.dtype, 0.0))

    def new_full(self, shape, value):
        return NumpyLink(shape, _pinned_like(shape, self._data.dtype, value))

    def new_empty(self, shape):
        return NumpyLink(shape, _pinned_like(shape, self._data.dtype))

    def clone(self):
        return NumpyLink(self._data.shape, np.copy(self._data))


backends.append(NumpyLink)
