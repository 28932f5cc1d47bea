"""Axis-aligned volume geometries.

API mirror of the reference's ``tomosipo/geometry/volume.py``.  The ASTRA
volume dict this class emits (``to_astra``, reference ``volume.py:286-301``) is
one half of what ``create_astra_projector`` hands to the backend; the other
half is the projection geometry's vectors.
"""
import warnings
from typing import Tuple, Union

import numpy as np

import tomosipo_b200 as ts
from .. import astra_compat
from ..types import ToPos, ToShape3D, ToSize3D
from .transform import Transform
from .volume_vec import VolumeVectorGeometry

Extent = Union[Tuple[float, float], Tuple[Tuple[float, float], Tuple[float, float], Tuple[float, float]]]


def is_volume(g):
    """True for axis-aligned and vector volume geometries."""
    return isinstance(g, (VolumeGeometry, VolumeVectorGeometry))


def volume(*, shape: ToShape3D = (1, 1, 1), pos: ToPos = None, size: ToSize3D = None, extent: Extent = None):
    """Create an axis-aligned volume geometry.

    Give ``pos`` and/or ``size`` (defaults: origin; one unit per voxel), or
    ``extent`` = ((min_z, max_z), (min_y, max_y), (min_x, max_x)).

    >>> ts.volume(shape=2, size=1).voxel_size
    (0.5, 0.5, 0.5)
    """
    shape = ts.types.to_shape3d(shape)
    if extent is not None:
        if pos is not None:
            raise ValueError("ts.volume does not accept both `extent` and `pos` arguments. ")
        if size is not None:
            raise ValueError("ts.volume does not accept both `extent` and `size` arguments. ")
        pos, size = _extent_to_pos_size(extent)
        return VolumeGeometry(shape, pos, size)
    return VolumeGeometry(shape, pos=0 if pos is None else pos, size=shape if size is None else size)


def random_volume():
    """A random axis-aligned volume (unseeded)."""
    return volume(
        shape=np.random.uniform(2, 100, 3).astype(int),
        pos=np.random.normal(size=3),
        size=np.random.uniform(1, 10, size=3),
    )


def _pos_size_to_extent(pos, size):
    pos = np.array(ts.types.to_pos(pos))
    half = 0.5 * np.array(ts.types.to_size3d(size))
    return tuple((lo, hi) for lo, hi in zip(pos - half, pos + half))


def _extent_to_pos_size(extent):
    size = ts.types.to_size3d(tuple(hi - lo for lo, hi in extent))
    pos = ts.types.to_pos(tuple((hi + lo) / 2 for lo, hi in extent))
    return pos, size


class VolumeGeometry:
    """Axis-aligned box of ``shape`` voxels centred on ``pos`` with physical ``size``."""

    def __init__(self, shape=(1, 1, 1), pos=0, size=None):
        shape = ts.types.to_shape3d(shape)
        pos = ts.types.to_pos(pos)
        if size is None:
            self._inner = ts.volume_vec(shape=shape, pos=pos)
        else:
            size = ts.types.to_size3d(size)
            vz, vy, vx = (sz / n for sz, n in zip(size, shape))
            self._inner = ts.volume_vec(shape=shape, pos=pos, w=(vz, 0, 0), v=(0, vy, 0), u=(0, 0, vx))

    def __repr__(self):
        return (
            f"ts.volume(\n"
            f"    shape={self._inner.shape},\n"
            f"    pos={tuple(float(p) for p in self.pos[0])},\n"
            f"    size={self.size},\n"
            f")"
        )

    def __eq__(self, other):
        return isinstance(other, VolumeGeometry) and self._inner == other._inner

    def __getitem__(self, key):
        """Slice in (z, y, x); a step > 1 bins voxels.

        >>> ts.volume(shape=4)[:2].shape
        (2, 4, 4)
        """
        key = key if isinstance(key, tuple) else (key,)
        sub = self._inner[(0,) + key]
        return VolumeGeometry(sub.shape, sub.pos[0], sub.size)

    def __contains__(self, other):
        return all(s[0] <= o[0] and o[1] <= s[1] for s, o in zip(self.extent, other.extent))

    def __len__(self):
        return 1

    def to_astra(self):
        """ASTRA volume-geometry dict (GridRowCount = Y, GridColCount = X, GridSliceCount = Z)."""
        nz, ny, nx = self.shape
        ez, ey, ex = self.extent
        return astra_compat.create_vol_geom(ny, nx, nz, *ex, *ey, *ez)

    def to_vec(self):
        return self._inner

    @property
    def num_steps(self):
        return 1

    @property
    def pos(self):
        return self._inner.pos

    @property
    def w(self):
        return self._inner.w

    @property
    def v(self):
        return self._inner.v

    @property
    def u(self):
        return self._inner.u

    @property
    def shape(self):
        return self._inner.shape

    @property
    def sizes(self):
        return self._inner.sizes

    @property
    def size(self):
        return self._inner.size

    @property
    def voxel_sizes(self):
        return self._inner.voxel_sizes

    @property
    def voxel_size(self):
        return self._inner.voxel_size

    @property
    def extent(self):
        """((min_z, max_z), (min_y, max_y), (min_x, max_x))."""
        return _pos_size_to_extent(self._inner.pos[0], self._inner.size)

    @property
    def corners(self):
        return self._inner.corners

    @property
    def lower_left_corner(self):
        return self._inner.lower_left_corner

    def with_voxel_size(self, voxel_size):
        """Same centre, as many whole voxels of ``voxel_size`` as fit."""
        voxel_size = ts.types.to_size3d(voxel_size)
        new_shape = (np.array(self.size) / voxel_size).astype(int)
        return VolumeGeometry(new_shape, pos=self.pos[0], size=new_shape * voxel_size)

    def reshape(self, new_shape):
        return VolumeGeometry(new_shape, pos=self.pos[0], size=self.size)

    def translate(self, t):
        t = ts.types.to_pos(t)
        return VolumeGeometry(self.shape, pos=tuple(p + d for p, d in zip(self.pos[0], t)), size=self.size)

    def untranslate(self, t):
        return self.translate(-np.array(t))

    def scale(self, scale):
        """Scale around the volume centre."""
        scale = ts.types.to_size3d(scale)
        return VolumeGeometry(self.shape, pos=self.pos[0], size=tuple(a * b for a, b in zip(scale, self.size)))

    def multiply(self, scale):
        """Scale around the origin (moves the centre as well)."""
        scale = ts.types.to_size3d(scale)
        return VolumeGeometry(
            self.shape,
            pos=tuple(a * b for a, b in zip(scale, self.pos[0])),
            size=tuple(a * b for a, b in zip(scale, self.size)),
        )

    def __rmul__(self, other):
        if not isinstance(other, Transform):
            return NotImplemented
        if other.num_steps == 1:
            # translation * positive scaling keeps the box axis-aligned
            shift = other.matrix[0, :3, 3]
            factors = abs(other.matrix[0].diagonal()[:3])
            if ts.translate(shift) * ts.scale(factors) == other:
                return self.multiply(factors).translate(shift)
        warnings.warn(
            "Converting VolumeGeometry to VolumeVectorGeometry. "
            "Use `T * vg.to_vec()` to inhibit this warning. ",
            stacklevel=2,
        )
        return other * self.to_vec()


def from_astra(astra_vol_geom):
    """``VolumeGeometry`` from an ASTRA 3D volume-geometry dict."""
    opt = astra_vol_geom["option"]
    shape = (astra_vol_geom["GridSliceCount"], astra_vol_geom["GridRowCount"], astra_vol_geom["GridColCount"])
    extent = tuple((opt[f"WindowMin{a}"], opt[f"WindowMax{a}"]) for a in "ZYX")
    pos, size = _extent_to_pos_size(extent)
    return VolumeGeometry(shape=shape, pos=pos, size=size)
