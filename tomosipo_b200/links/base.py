"""Array links: how the projector sees numpy / torch / cupy arrays.

API mirror of the reference's ``tomosipo/links/base.py``: ``link``,
``geometry_shape``, ``are_compatible``, the ``Link`` protocol and the
``backends`` registry.  Where the reference's links hand ASTRA either an
``ndarray`` or an ``astra.data3d.GPULink(ptr, x, y, z, pitch)``, these links
hand the C ABI a :class:`RawBuffer` -- the same information, no ASTRA types.
"""
import warnings
from collections import namedtuple
from contextlib import contextmanager

import tomosipo_b200 as ts

#: registered link classes, tried in order (numpy first, then torch, then cupy)
backends = []

#: What ``Link.linked_data`` returns: a dense float32 [z][y][x] buffer.
#: ``kind`` is ``"host"`` or ``"device"``; ``device``/``stream`` only matter for
#: device buffers; ``keepalive`` pins the owning array for the call's duration.
RawBuffer = namedtuple("RawBuffer", "ptr shape kind device stream keepalive")


def link(geometry, arr):
    """Link ``arr`` (array, scalar or ``None``) to the data layout of ``geometry``."""
    shape = geometry_shape(geometry)
    for backend in backends:
        if backend.__accepts__(arr):
            return backend(shape, arr)
    raise ValueError(f"An initial_value of class {type(arr)} is not supported. ")


def geometry_shape(geometry):
    """Array shape of data living on ``geometry``.

    Volumes are ``(z, y, x)``; projection stacks are ``(v, angle, u)``.
    """
    if ts.geometry.is_volume(geometry):
        return geometry.shape
    if ts.geometry.is_projection(geometry):
        rows, cols = geometry.det_shape
        return (rows, geometry.num_angles, cols)
    raise ValueError(
        f"Geometry '{type(geometry)}' is not supported. Cannot determine if volume or projection geometry."
    )


def are_compatible(link_a, link_b):
    """Can the projector run from one link to the other (same device)?"""
    for first, second in ((link_a, link_b), (link_b, link_a)):
        verdict = first.__compatible_with__(second)
        if verdict is True:
            return True
        if verdict is not NotImplemented:
            return False
    warnings.warn(
        f"Cannot determine if link of type {type(link_a)} is compatible with {type(link_b)}. "
        "Continuing anyway."
    )
    # The reference returns None here (links/base.py:42-46) although it says
    # "continuing"; we do continue.
    return True


class Link(object):
    """Base class of array links."""

    def __init__(self, shape, initial_value):
        self._shape = shape
        super().__init__()

    # protocol ---------------------------------------------------------------
    @staticmethod
    def __accepts__(initial_value):
        """Can this link class wrap ``initial_value``?"""
        raise NotImplementedError()

    def __compatible_with__(self, other):
        """True / False / NotImplemented: can we project between self and other?"""
        raise NotImplementedError()

    # properties -------------------------------------------------------------
    @property
    def linked_data(self):
        """:class:`RawBuffer` describing the memory the projector reads / writes."""
        raise NotImplementedError()

    @property
    def data(self):
        """The wrapped array (shared, not copied)."""
        raise NotImplementedError()

    @data.setter
    def data(self, val):
        raise AttributeError(
            "You cannot change which array backs a dataset.\n"
            "To change the underlying data instead, use: \n"
            " >>> x.data[:] = new_data\n"
        )

    @property
    def shape(self):
        return self._shape

    @contextmanager
    def context(self):
        """Make the array's device current for the duration of a projection."""
        raise NotImplementedError()

    # allocation -------------------------------------------------------------
    def new_zeros(self, shape):
        raise NotImplementedError()

    def new_full(self, shape, value):
        raise NotImplementedError()

    def new_empty(self, shape):
        raise NotImplementedError()

    def clone(self):
        raise NotImplementedError()
