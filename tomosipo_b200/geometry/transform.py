"""4x4 homogeneous transforms over time steps.

API mirror of the reference's ``tomosipo/geometry/transform.py``
(``Transform``, ``identity``, ``translate``, ``scale``, ``rotate``,
``reflect``, ``to_perspective``, ``from_perspective``, ``random_transform``).
Coordinates are (z, y, x); matrices are stored as ``(num_steps, 4, 4)``.
``from_perspective`` is on the hot path's set-up side: ``Operator`` uses it
to un-rotate vector volumes (reference ``Operator.py:41-49``).
"""
import warnings
from typing import Any

import numpy as np

import tomosipo_b200 as ts
from .. import vector_calc as vc
from ..types import ToHomogeneousVec, ToScalars


class Transform(object):
    """A sequence of projective 4x4 matrices, one per time step."""

    def __init__(self, matrix):
        super().__init__()
        self.matrix, _ = vc._broadcastmm(matrix, matrix)

    def __mul__(self, other):
        if not isinstance(other, Transform):
            return NotImplemented
        n, m = self.num_steps, other.num_steps
        if not (n == 1 or m == 1 or n == m):
            raise ValueError(
                f"Cannot multiply transforms with different number of time steps. Got steps: {n} and {m}"
            )
        return Transform(vc.matrix_matrix_transform(self.matrix, other.matrix))

    def __repr__(self):
        return f"Transform(\n    {self.matrix}\n)"

    def __eq__(self, other):
        if not isinstance(other, Transform):
            return False
        A, B = vc._broadcastmm(self.matrix, other.matrix)
        return bool(np.all(np.abs(A - B) < ts.epsilon))

    def __getitem__(self, i):
        if not isinstance(i, (slice, int)):
            raise TypeError(f"Transform only support one-dimensional indexing. Got: {i}")
        return Transform(self.matrix[i])

    @property
    def num_steps(self):
        return self.matrix.shape[0]

    @property
    def inv(self):
        return Transform(vc.invert_transformation_matrix(self.matrix))

    def transform_vec(self, vec):
        """Apply to direction vectors (translation has no effect)."""
        return vc.to_vec(vc.matrix_transform(self.matrix, vc.to_homogeneous_vec(vec)))

    def transform_point(self, points):
        """Apply to positions."""
        return vc.to_vec(vc.matrix_transform(self.matrix, vc.to_homogeneous_point(points)))


def _from_columns(c0, c1, c2, c3):
    """Stack four (N, 4) column arrays into (N, 4, 4) matrices."""
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    return Transform(np.stack((c0, c1, c2, c3), axis=2))


def identity():
    """The identity transform (one step)."""
    return Transform(np.eye(4))


def translate(axis: ToHomogeneousVec, *, alpha: ToScalars = 1):
    """Translation by ``alpha * axis`` (both may vary per step).

    >>> ts.translate((1, 0, 0)).transform_point((0, 0, 0))
    array([[1., 0., 0.]])
    """
    axis = vc.to_homogeneous_point(axis)
    alpha = ts.types.to_scalars(alpha, var_name="alpha")
    if not (len(axis) == 1 or len(alpha) == 1 or len(axis) == len(alpha)):
        raise ValueError(
            f"Expected parameters `axis` and `alpha` to be the same length. Got: {len(axis)}, {len(alpha)}"
        )
    shift = alpha[:, None] * axis
    shift[:, 3] = 1.0
    eye = np.eye(4)
    return _from_columns(eye[None, :, 0], eye[None, :, 1], eye[None, :, 2], shift)


def scale(scale, *, pos=0, alpha=1.0):
    """Anisotropic scaling by ``alpha * scale`` around ``pos``."""
    if np.isscalar(scale):
        scale = ts.types.to_size3d(scale)
    if np.isscalar(pos):
        pos = ts.types.to_pos(pos)
    scale = ts.types.to_homogeneous_vec(scale)
    pos = ts.types.to_homogeneous_pos(pos)
    alpha = ts.types.to_scalars(alpha)
    l1, l2, l3 = len(scale), len(pos), len(alpha)
    try:
        n = vc.broadcast_lengths(vc.broadcast_lengths(l1, l2), l3)
    except ValueError:
        raise ValueError(
            f"Expected `scale`, `pos`, and `alpha` to be broadcastable. Got lengths: {l1}, {l2}, and {l3}."
        )
    s = alpha[:, None] * scale  # (n', 4), last column 0
    S = np.zeros((s.shape[0], 4, 4))
    S[:, 0, 0], S[:, 1, 1], S[:, 2, 2], S[:, 3, 3] = s[:, 0], s[:, 1], s[:, 2], 1.0
    T = translate(pos)
    return T * Transform(S) * T.inv


def rotate(*, pos, axis, angles=None, rad=None, deg=None, right_handed=True):
    """Rotation by ``angles`` (radians) around the line through ``pos`` along ``axis``.

    The (z, y, x) frame is left-handed; ``right_handed=True`` (default) turns
    counter-clockwise when looking against the axis in the usual right-handed
    drawing of (x, y, z).
    """
    if np.isscalar(pos):
        pos = ts.types.to_pos(pos)
    pos = vc.to_homogeneous_point(pos)
    axis = vc.to_homogeneous_vec(axis)
    pos, axis = vc._broadcastv(pos, axis)
    axis = axis / vc.norm(axis)[:, None]

    legacy = rad is not None or deg is not None
    if angles is None and not legacy:
        raise ValueError("The `angles=` parameter is required.")
    if angles is not None and legacy:
        raise TypeError(
            "The `angles` parameter is not compatible with the `rad` or `deg` parameter. "
            "The `rad` and `deg` parameters are deprecated. "
        )
    if angles is None:
        warnings.warn(
            "The `rad` and `deg` parameters of `ts.rotate` are deprecated. Please use `angles` instead.",
            category=DeprecationWarning,
            stacklevel=2,
        )
        angles = np.deg2rad(deg) if deg is not None else rad
    theta = vc.to_scalar(angles)[:, 0]
    if not right_handed:
        theta = -theta

    # Rodrigues in the left-handed (z, y, x) frame: R = cos I + (1 - cos) a a^T - sin [a]_x
    # (the transpose of the textbook right-handed matrix, as in the reference,
    # transform.py:413-431, whose rows are stacked as columns)
    n = vc.broadcast_lengths(len(theta), len(axis))
    a = np.broadcast_to(axis[:, :3], (n, 3))
    c = np.broadcast_to(np.cos(theta), (n,))
    s = np.broadcast_to(np.sin(theta), (n,))
    K = np.zeros((n, 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -a[:, 2], a[:, 1]
    K[:, 1, 0], K[:, 1, 2] = a[:, 2], -a[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -a[:, 1], a[:, 0]
    R = np.zeros((n, 4, 4))
    R[:, :3, :3] = (
        c[:, None, None] * np.eye(3)[None]
        + (1 - c)[:, None, None] * a[:, :, None] * a[:, None, :]
        - s[:, None, None] * K
    )
    R[:, 3, 3] = 1.0
    T = translate(-pos)
    return T.inv * Transform(R) * T


def reflect(*, pos, axis):
    """Reflection in the plane through ``pos`` with normal ``axis``."""
    if np.isscalar(pos):
        pos = ts.types.to_pos(pos)
    pos = vc.to_homogeneous_point(pos)
    axis = vc.to_homogeneous_vec(axis)
    pos, axis = vc._broadcastv(pos, axis)
    axis = axis / vc.norm(axis)[:, None]
    H = np.eye(4)[None] - 2.0 * axis[:, :, None] * axis[:, None, :]
    T = translate(pos)
    return T * Transform(H) * T.inv


def to_perspective(*, pos: ToHomogeneousVec = None, w: ToHomogeneousVec = None, v: ToHomogeneousVec = None,
                   u: ToHomogeneousVec = None, vol: Any = None, ignore_scale: bool = True):
    """Transform mapping the standard frame onto the frame ``(pos; w, v, u)``.

    With ``ignore_scale`` the basis vectors are normalised first, so volumes
    keep their size.
    """
    if vol is not None:
        pos, w, v, u = vol.pos, vol.w, vol.v, vol.u
    if any(x is None for x in (pos, w, v, u)):
        raise ValueError("Not enough arguments provided: one of pos, w, v, u is missing.")
    pos = vc.to_homogeneous_point(pos)
    w, v, u = (vc.to_homogeneous_vec(x) for x in (w, v, u))
    pos, w, v, u = np.broadcast_arrays(pos, w, v, u)
    vc.check_same_shapes(pos, w, v, u)
    if ignore_scale:
        w, v, u = (x / vc.norm(x)[:, None] for x in (w, v, u))
    assert pos.ndim == 2
    return Transform(np.stack((w, v, u, pos), axis=2))


def from_perspective(*, pos: ToHomogeneousVec = None, w: ToHomogeneousVec = None, v: ToHomogeneousVec = None,
                     u: ToHomogeneousVec = None, vol: Any = None, ignore_scale: bool = True):
    """Inverse of :func:`to_perspective`: re-express geometry in the frame ``(pos; w, v, u)``."""
    return to_perspective(pos=pos, w=w, v=v, u=u, vol=vol, ignore_scale=ignore_scale).inv


def random_transform():
    """A random rotation * scaling * translation (unseeded, as in the reference)."""
    t, pos, axis, s = np.random.normal(size=(4, 3))
    angle = np.random.normal()
    return rotate(pos=pos, axis=axis, angles=angle) * scale(abs(s)) * translate(t)
