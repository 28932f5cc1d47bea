"""Arbitrarily oriented, possibly moving, volume geometries.

API mirror of the reference's ``tomosipo/geometry/volume_vec.py``.  A vector
volume is a box of ``shape`` voxels centred on ``pos`` whose voxel edges are
the vectors ``w`` (z-like), ``v`` (y-like) and ``u`` (x-like), one set per time
step.  ``Operator`` turns such a volume into an axis-aligned one by moving the
projection geometry instead (reference ``Operator.py:11-60``).
"""
from numbers import Integral
from typing import Union

import numpy as np

import tomosipo_b200 as ts
from .. import vector_calc as vc
from ..types import ToPos, ToShape3D, ToSize3D, ToVec
from ..utils import slice_interval
from .transform import Transform


def volume_vec(*, shape: ToShape3D, pos: Union[float, ToVec] = 0, w: ToVec = (1, 0, 0), v: ToVec = (0, 1, 0),
               u: ToVec = (0, 0, 1)):
    """Create an arbitrarily oriented volume geometry.

    >>> ts.volume_vec(shape=1, pos=(0, 0, 0), w=(1, 0, 0), v=(0, 1, 0), u=(0, 0, 1)).num_steps
    1
    """
    return VolumeVectorGeometry(shape, pos, w, v, u)


def random_volume_vec():
    """A random vector volume (unseeded)."""
    vg = volume_vec(shape=np.random.uniform(4, 10, size=3).astype(int), pos=np.random.normal(size=3))
    return ts.geometry.random_transform() * vg


class VolumeVectorGeometry(object):
    """Volume given by a centre and three voxel-edge vectors per time step."""

    def __init__(self, shape, pos, w=(1, 0, 0), v=(0, 1, 0), u=(0, 0, 1)):
        super().__init__()
        self._shape = ts.types.to_shape3d(shape)
        if np.isscalar(pos) and pos == 0.0:
            pos = (0.0, 0.0, 0.0)
        parts = [
            ts.types.to_vec(pos, "position"),
            ts.types.to_vec(w, "w axis"),
            ts.types.to_vec(v, "v axis"),
            ts.types.to_vec(u, "u axis"),
        ]
        try:
            parts = np.broadcast_arrays(*parts)
        except ValueError:
            shapes = [x.shape for x in parts]
            raise ValueError(f"Not all arguments pos, w, v, u are the same shape. Got: {shapes}")
        self._pos, self._w, self._v, self._u = parts

    def __repr__(self):
        with ts.utils.print_options():
            return (
                f"ts.volume_vec(\n"
                f"    shape={self._shape},\n"
                f"    pos={repr(self.pos)},\n"
                f"    w={repr(self.w)},\n"
                f"    v={repr(self.v)},\n"
                f"    u={repr(self.u)},\n"
                f")"
            )

    def __eq__(self, other):
        if not isinstance(other, VolumeVectorGeometry):
            return False
        if self.shape != other.shape:
            return False
        pairs = ((self._pos, other._pos), (self._w, other._w), (self._v, other._v), (self._u, other._u))
        try:
            return bool(all(np.all(np.abs(a - b) < ts.epsilon) for a, b in pairs))
        except ValueError:  # different number of steps
            return False

    def __getitem__(self, key):
        """Index as ``vg[step, w, v, u]``; spatial slices may bin voxels (step > 1).

        >>> ts.volume_vec(shape=4, pos=0)[:, :2].shape
        (2, 4, 4)
        """
        everything = slice(None, None, None)
        if isinstance(key, tuple) and len(key) > 4:
            raise ValueError(f"VolumeVectorGeometry supports indexing in 4 dimensions. Got {key}.")
        if isinstance(key, (Integral, slice)):
            key = (key,)
        if isinstance(key, tuple):
            key = key + (everything,) * (4 - len(key))
            steps = key[0]
            lo, hi, counts, scales = [], [], [], []
            for n, k in zip(self.shape, key[1:]):
                a, b, count, scale = slice_interval(0, n, n, k)
                lo.append(a); hi.append(b); counts.append(count); scales.append(scale)
            axes = (self.w, self.v, self.u)
            corner = self.lower_left_corner
            first = corner + sum(a * e for a, e in zip(lo, axes))
            last = corner + sum(b * e for b, e in zip(hi, axes))
            centre = ((first + last) / 2)[steps]
            return VolumeVectorGeometry(
                tuple(counts), centre, scales[0] * self.w[steps], scales[1] * self.v[steps], scales[2] * self.u[steps]
            )
        return VolumeVectorGeometry(self.shape, self.pos[key], self.w[key], self.v[key], self.u[key])

    def __len__(self):
        return self.num_steps

    def to_vec(self):
        return self

    @property
    def num_steps(self):
        """Number of positions / orientations."""
        return len(self._pos)

    @property
    def pos(self):
        return np.copy(self._pos)

    @property
    def w(self):
        return np.copy(self._w)

    @property
    def v(self):
        return np.copy(self._v)

    @property
    def u(self):
        return np.copy(self._u)

    @property
    def shape(self):
        return self._shape

    @property
    def sizes(self):
        """(num_steps, 3) physical edge lengths of the whole box."""
        edges = [n * vc.norm(e) for n, e in zip(self._shape, (self._w, self._v, self._u))]
        return np.stack(edges, axis=1)

    @property
    def size(self):
        """Box size when constant over time; raises ``ValueError`` otherwise."""
        sizes = self.sizes
        if np.all(np.ptp(sizes, axis=0) < ts.epsilon):
            return tuple(float(x) for x in sizes[0])
        raise ValueError("The size of the volume is not constant. To prevent this error, use `vg.sizes'. ")

    @property
    def voxel_sizes(self):
        return self.sizes / np.array([self.shape])

    @property
    def voxel_size(self):
        return tuple(sz / n for sz, n in zip(self.size, self.shape))

    @property
    def corners(self):
        """(num_steps, 8, 3): corners ordered by (w, v, u) bits."""
        bits = np.array([(a, b, c) for a in (0, 1) for b in (0, 1) for c in (0, 1)], dtype=np.float64) - 0.5
        n = self._shape
        out = (
            self._pos[None]
            + bits[:, 0, None, None] * self._w[None] * n[0]
            + bits[:, 1, None, None] * self._v[None] * n[1]
            + bits[:, 2, None, None] * self._u[None] * n[2]
        )
        return out.swapaxes(0, 1)

    @property
    def lower_left_corner(self):
        n = self._shape
        return self._pos - (self._w * n[0] + self._v * n[1] + self._u * n[2]) / 2

    def reshape(self, new_shape):
        """Change the voxel count, keeping the physical box."""
        new_shape = ts.types.to_shape3d(new_shape)
        factors = [old / max(new, 1) for old, new in zip(self._shape, new_shape)]
        return VolumeVectorGeometry(
            new_shape, self.pos, self._w * factors[0], self._v * factors[1], self._u * factors[2]
        )

    def __rmul__(self, other):
        if isinstance(other, Transform):
            return VolumeVectorGeometry(
                self.shape,
                other.transform_point(self._pos),
                other.transform_vec(self._w),
                other.transform_vec(self._v),
                other.transform_vec(self._u),
            )
        return NotImplemented
