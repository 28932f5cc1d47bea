"""CuPy link.  API mirror of ``tomosipo/links/cupy.py`` (import it explicitly, or
via ``tomosipo_b200.cupy``; CuPy is optional and absent from the build image)."""
import warnings
from contextlib import contextmanager

import cupy

from .base import Link, RawBuffer, backends


def _torch_link_class():
    """TorchLink if PyTorch has been imported by the application, else None (never imports torch itself)."""
    import sys

    mod = sys.modules.get("tomosipo_b200.links.torch")
    return getattr(mod, "TorchLink", None)


class CupyLink(Link):
    """Wraps a C-contiguous float32 ``cupy.ndarray``."""

    def __init__(self, shape, initial_value):
        super().__init__(shape, initial_value)
        if not isinstance(initial_value, cupy.ndarray):
            raise ValueError(f"Expected initial_value to be a `cupy.ndarray'. Got {initial_value.__class__}")
        a = initial_value
        if a.shape == ():
            self._data = cupy.zeros(shape, dtype=cupy.float32)
            self._data[:] = a
            return
        if tuple(shape) != tuple(a.shape):
            raise ValueError(f"Expected initial_value with shape {shape}. Got {a.shape}")
        if a.dtype != cupy.float32:
            warnings.warn(
                f"The parameter initial_value is of type {a.dtype}; expected `cupy.float32`. "
                f"The type has been automatically converted. "
                f"Use `ts.link(x.astype(cupy.float32))' to inhibit this warning. "
            )
            a = a.astype(cupy.float32)
        # the reference only checks contiguity inside the dtype branch (links/cupy.py:41-55)
        if not a.flags["C_CONTIGUOUS"]:
            warnings.warn(
                f"The parameter initial_value should be contiguous. "
                f"It has been automatically made contiguous. "
                f"Use `ts.link(cupy.ascontiguousarray(x))' to inhibit this warning. "
            )
            a = cupy.ascontiguousarray(a)
        self._data = a

    @staticmethod
    def __accepts__(initial_value):
        return isinstance(initial_value, cupy.ndarray)

    def __compatible_with__(self, other):
        if isinstance(other, CupyLink):
            return self._data.device == other._data.device
        # CUDA tensors of PyTorch on the same GPU (a TODO in the reference, links/cupy.py:66-72): both are plain
        # device pointers to the C ABI; direct_project orders the two libraries' current streams
        torch_link = _torch_link_class()
        if torch_link is not None and isinstance(other, torch_link):
            t = other.data
            return bool(t.is_cuda) and t.device.index == self._data.device.id
        return NotImplemented

    @property
    def linked_data(self):
        a = self._data
        stream = cupy.cuda.get_current_stream().ptr
        return RawBuffer(a.data.ptr, tuple(a.shape), "device", a.device.id, stream, a)

    @property
    def data(self):
        return self._data

    @data.setter
    def data(self, val):
        raise AttributeError(
            "You cannot change which cupy array backs a dataset.\n"
            "To change the underlying data instead, use: \n"
            " >>> vd.data[:] = new_data\n"
        )

    @contextmanager
    def context(self):
        with self._data.device:
            yield

    def new_zeros(self, shape):
        with self._data.device:
            return CupyLink(shape, cupy.zeros(shape, dtype=self._data.dtype))

    def new_full(self, shape, value):
        with self._data.device:
            return CupyLink(shape, cupy.full(shape, value, dtype=self._data.dtype))

    def new_empty(self, shape):
        with self._data.device:
            return CupyLink(shape, cupy.empty(shape, dtype=self._data.dtype))

    def clone(self):
        return CupyLink(self._data.shape, self._data.copy())


backends.append(CupyLink)
