"""PyTorch link (CPU and CUDA tensors).  API mirror of ``tomosipo/links/torch.py``.

CUDA tensors are projected in place on their own device *and on the current
torch stream* (the reference only sets the device, ``links/torch.py:139-155``);
CPU tensors take the host path of the C ABI.
"""
import warnings
from contextlib import contextmanager

import torch

from .base import Link, RawBuffer, backends
from .numpy import NumpyLink


class TorchLink(Link):
    """Wraps a contiguous float32 tensor; other inputs are converted with a warning."""

    def __init__(self, shape, initial_value):
        super().__init__(shape, initial_value)
        if not isinstance(initial_value, torch.Tensor):
            raise ValueError(f"Expected initial_value to be a `torch.Tensor'. Got {initial_value.__class__}")
        t = initial_value
        if t.shape == torch.Size([]):
            self._data = torch.zeros(shape, dtype=torch.float32, device=t.device)
            self._data[:] = t
            return
        if tuple(shape) != tuple(t.shape):
            raise ValueError(f"Expected initial_value with shape {shape}. Got {t.shape}")
        if t.dtype != torch.float32:
            warnings.warn(
                f"The parameter initial_value is of type {t.dtype}; expected `torch.float32`. "
                f"The type has been automatically converted. "
                f"Use `ts.link(x.to(dtype=torch.float32))' to inhibit this warning. "
            )
            t = t.to(dtype=torch.float32)
        if not t.is_contiguous():
            warnings.warn(
                f"The parameter initial_value should be contiguous. "
                f"It has been automatically made contiguous. "
                f"Use `ts.link(x.contiguous())' to inhibit this warning. "
            )
            t = t.contiguous()
        self._data = t

    @staticmethod
    def __accepts__(initial_value):
        return isinstance(initial_value, torch.Tensor)

    def __compatible_with__(self, other):
        if isinstance(other, NumpyLink):
            theirs = torch.device("cpu")
        elif isinstance(other, TorchLink):
            theirs = other._data.device
        else:
            return NotImplemented
        return self._data.device == theirs

    @property
    def linked_data(self):
        t = self._data
        if t.is_cuda:
            stream = torch.cuda.current_stream(t.device).cuda_stream
            return RawBuffer(t.data_ptr(), tuple(t.shape), "device", t.device.index, stream, t)
        # may be part of an autograd graph: detach to reach the storage
        return RawBuffer(t.detach().data_ptr(), tuple(t.shape), "host", 0, 0, t)

    @property
    def data(self):
        """The shared tensor; projection data is ordered (v, angle, u)."""
        return self._data

    @data.setter
    def data(self, val):
        raise AttributeError(
            "You cannot change which torch tensor backs a dataset.\n"
            "To change the underlying data instead, use: \n"
            " >>> vd.data[:] = new_data\n"
        )

    @contextmanager
    def context(self):
        if self._data.is_cuda:
            with torch.cuda.device_of(self._data):
                yield
        else:
            yield

    def new_zeros(self, shape):
        return TorchLink(shape, self._data.new_zeros(shape))

    def new_full(self, shape, value):
        return TorchLink(shape, self._data.new_full(shape, value))

    def new_empty(self, shape):
        return TorchLink(shape, self._data.new_empty(shape))

    def clone(self):
        return TorchLink(self._data.shape, self._data.clone())


if not hasattr(torch, "__sphinx_mock__"):
    backends.append(TorchLink)
