"""Argument coercion helpers (shapes, sizes, positions, scalars, vectors).

API mirror of the reference's ``tomosipo/types.py`` (same function names,
same accepted inputs, same exception types and messages) re-stated for
NumPy >= 2: the reference's ``np.array(..., copy=False)`` calls
(``tomosipo/types.py:279,333,337,355,359``) raise on NumPy 2, so everything
here goes through ``np.asarray``.
"""
from collections import abc
from numbers import Integral
from typing import Any, Collection, Iterable, Tuple, TypeVar, Union

import numpy as np

T = TypeVar("T")

Shape2D = Tuple[int, int]
Shape3D = Tuple[int, int, int]
Size2D = Tuple[float, float]
Size3D = Tuple[float, float, float]
ToShape2D = Union[int, Tuple[int, int], Iterable[int]]
ToShape3D = Union[int, Tuple[int, int, int], Iterable[int]]
ToSize2D = Union[float, Tuple[float, float], Iterable[float]]
ToSize3D = Union[float, Tuple[float, float, float], Iterable[float]]
Pos = Tuple[float, float, float]
ToPos = Union[float, Collection[float]]
Scalars = np.ndarray
ToScalars = Union[float, Collection[float], np.ndarray]
Vec = np.ndarray
HomogeneousVec = np.ndarray
ToVec = Union[Tuple[float, float, float], Iterable[Tuple[float, float, float]], np.ndarray]
ToHomogeneousVec = Union[ToVec, Tuple[float, float, float, float], Iterable[Tuple[float, float, float, float]]]

_EPS = 1e-8  # == tomosipo_b200.epsilon (kept local to avoid an import cycle)


def to_tuple(val, n):
    """Broadcast a scalar to an ``n``-tuple, or check the length of an iterable.

    >>> to_tuple(1, 2)
    (1, 1)
    >>> to_tuple((1, 2), n=2)
    (1, 2)
    >>> to_tuple((1, 2, 3), n=2)
    Traceback (most recent call last):
    ...
    TypeError: Expected tuple with 2 elements. Got (1, 2, 3).
    """
    n = int(n)
    if not isinstance(val, abc.Iterable):
        return (val,) * n
    if len(tuple(val)) != n:
        raise TypeError(f"Expected tuple with {n} elements. Got {repr(val)}.")
    return val


def to_float_tuple(val, n, var_name="value"):
    """``n`` floats from a scalar or an iterable.

    >>> to_float_tuple(1, 2)
    (1.0, 1.0)
    >>> to_float_tuple(('a', 0), n=2)
    Traceback (most recent call last):
    ...
    TypeError: value must contain only floats. Got ('a', 0).
    """
    val = to_tuple(val, n)
    try:
        return tuple(float(v) for v in val)
    except ValueError:
        raise TypeError(f"{var_name} must contain only floats. Got {repr(val)}.")


def to_shape_nd(shape, n):
    shape = to_tuple(shape, n)
    for s in shape:
        if not isinstance(s, Integral):
            raise TypeError(f"Shape must contain only integers. Got {shape} with type {type(s)}.")
    shape = tuple(int(s) for s in shape)
    if min(shape) < 1:
        raise TypeError(f"Shape must be positive. Got {shape}.")
    return shape


def to_shape2d(shape: ToShape2D) -> Shape2D:
    """
    >>> to_shape2d(1)
    (1, 1)
    >>> to_shape2d((5.0, 3))
    Traceback (most recent call last):
    ...
    TypeError: Shape must contain only integers. Got (5.0, 3) with type <class 'float'>.
    """
    return to_shape_nd(shape, 2)


def to_shape3d(shape: ToShape3D) -> Shape3D:
    """
    >>> to_shape3d((5, 3, 2))
    (5, 3, 2)
    >>> to_shape3d((5.0, 3))
    Traceback (most recent call last):
    ...
    TypeError: Expected tuple with 3 elements. Got (5.0, 3).
    """
    return to_shape_nd(shape, 3)


def to_size_nd(size, n):
    size = to_float_tuple(size, n, var_name="Size")
    if min(size) < -_EPS:
        raise TypeError(f"Size must be non-negative. Got {size}.")
    return size


def to_size2d(size: ToSize2D) -> Size2D:
    """
    >>> to_size2d((5, 3))
    (5.0, 3.0)
    """
    return to_size_nd(size, 2)


def to_size3d(size: ToSize3D) -> Size3D:
    """
    >>> to_size3d(1)
    (1.0, 1.0, 1.0)
    """
    return to_size_nd(size, 3)


def to_pos(pos: ToPos) -> Pos:
    """
    >>> to_pos(0)
    (0.0, 0.0, 0.0)
    >>> to_pos((3, 2, 1))
    (3.0, 2.0, 1.0)
    """
    if np.isscalar(pos) and pos == 0.0:
        return (0.0, 0.0, 0.0)
    if isinstance(pos, abc.Iterable):
        return to_float_tuple(pos, 3, "Position")
    raise TypeError("Cannot convert value to position. Expected (float, float, float). Got {pos}. ")


def to_scalars(s: ToScalars, var_name="scalars", accept_empty=False) -> Scalars:
    """1-D float64 array from a float or a collection of floats.

    >>> to_scalars((1, 1, 1))
    array([1., 1., 1.])
    >>> to_scalars(1).shape
    (1,)
    >>> to_scalars("string")
    Traceback (most recent call last):
    ...
    TypeError: Could not convert scalars to np.array. Got: 'string'.
    """
    try:
        arr = np.atleast_1d(np.asarray(s, dtype=np.float64))
    except (ValueError, TypeError):
        raise TypeError(f"Could not convert {var_name} to np.array. Got: {repr(s)}.")
    if np.isnan(arr).any():
        raise TypeError("Could not convert to array of scalars: array contains NaN.")
    if arr.ndim == 1 and (accept_empty or arr.size > 0):
        return arr
    raise TypeError(f"Value cannot be converted to {var_name}. Expected shape: (N,). Got shape: {arr.shape}.")


def to_vec(vec: ToVec, var_name="vector") -> Vec:
    """(N, 3) float64 array from one vector or a collection of vectors.

    >>> to_vec((1, 1, 1))
    array([[1., 1., 1.]])
    >>> to_vec([(1, 1, 1), (2, 2, 2)]).shape
    (2, 3)
    >>> to_vec("string")
    Traceback (most recent call last):
    ...
    TypeError: Could not convert vector to np.array. Got: 'string'.
    """
    try:
        arr = np.asarray(vec, dtype=np.float64)
    except (ValueError, TypeError):
        raise TypeError(f"Could not convert {var_name} to np.array. Got: {repr(vec)}.")
    original = arr.shape
    arr = np.atleast_2d(arr)
    if arr.ndim == 2 and arr.shape[1] == 3:
        return arr
    raise TypeError(
        f"Value cannot be converted to {var_name}. Expected shape: (3,) or (N, 3). Got shape: {original}."
    )


def to_homogeneous(vec, s) -> HomogeneousVec:
    """(N, 4) array: appends the homogeneous coordinate ``s`` to 3-vectors."""
    s = float(s)
    try:
        arr = np.asarray(vec, dtype=np.float64)
    except (ValueError, TypeError):
        raise TypeError(f"Could not convert value to np.array. Got: {repr(vec)}.")
    original = arr.shape
    arr = np.atleast_2d(arr)
    if arr.ndim == 2 and arr.shape[1] == 4:
        return arr
    if arr.ndim == 2 and arr.shape[1] == 3:
        return np.concatenate([arr, np.full((arr.shape[0], 1), s)], axis=1)
    raise TypeError(
        f"Value cannot be converted to homogeneous coordinates. "
        f"Expected shape: (3,), (4,), (N, 3), or (N, 4). Got shape: {original}. "
    )


def to_homogeneous_vec(vec: ToVec) -> HomogeneousVec:
    """Direction vectors: homogeneous coordinate 0."""
    return to_homogeneous(vec, 0.0)


def to_homogeneous_pos(vec: ToVec) -> HomogeneousVec:
    """Positions: homogeneous coordinate 1."""
    return to_homogeneous(vec, 1.0)
