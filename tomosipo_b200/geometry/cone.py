"""Circular cone-beam geometry (single rotation axis, flat detector).

API mirror of the reference's ``tomosipo/geometry/cone.py``.  ``to_astra`` /
``to_vec`` (reference ``cone.py:251-265``) produce the ASTRA ``'cone'`` dict and
its vector form; the vector form is what the projector consumes.
"""
import warnings
from typing import Union

import numpy as np

import tomosipo_b200 as ts
from .. import astra_compat
from ..types import ToScalars, ToShape2D, ToSize2D
from .base_projection import _CircularGeometry
from .cone_vec import ConeVectorGeometry
from .transform import Transform


def cone(*, angles: Union[int, ToScalars] = 1, shape: ToShape2D = (1, 1), size: ToSize2D = None,
         cone_angle: float = None, src_orig_dist: float = None, src_det_dist: float = None):
    """Create a circular cone-beam geometry.

    An integer ``angles`` means that many equi-spaced angles on [0, 2 pi).
    Either ``cone_angle`` (= detector height / source-detector distance, with
    the detector through the origin) or the two distances must be given.

    >>> ts.cone(angles=3, cone_angle=1/2).num_angles
    3
    """
    shape = ts.types.to_shape2d(shape)
    size = ts.types.to_size2d(shape if size is None else size)
    if cone_angle is None and src_orig_dist is None and src_det_dist is None:
        raise ValueError(
            "ts.cone requires at least one of `cone_angle`, `src_orig_dist`, or `src_det_dist` parameters. "
        )
    if cone_angle is not None and src_orig_dist is not None:
        raise ValueError("ts.cone does not accept both `cone_angle` and src_orig_dist` arguments at the same time. ")
    if cone_angle is not None and src_det_dist is not None:
        raise ValueError("ts.cone does not accept both `cone_angle` and src_det_dist` arguments at the same time. ")
    if cone_angle is not None:
        src_det_dist = src_orig_dist = size[0] / cone_angle
    elif src_orig_dist is None:
        src_orig_dist = src_det_dist
    elif src_det_dist is None:
        src_det_dist = src_orig_dist
    return ConeGeometry(angles=angles, shape=shape, size=size, src_orig_dist=src_orig_dist, src_det_dist=src_det_dist)


def random_cone():
    """A random circular cone geometry (unseeded)."""
    return cone(
        angles=np.random.normal(size=20),
        shape=np.random.uniform(10, 20, size=2).astype(int),
        size=np.random.uniform(10, 20, size=2),
        src_orig_dist=np.random.uniform(0, 10),
        src_det_dist=np.random.uniform(0, 20),
    )


class ConeGeometry(_CircularGeometry):
    """Angles, detector shape/size and the two source distances."""

    _is_cone = True

    def __init__(self, angles=1, shape=1, size=None, src_orig_dist=None, src_det_dist=None):
        super().__init__(shape=shape)
        self.angles_original = angles
        if np.isscalar(angles) and isinstance(angles, int):
            angles = np.linspace(0, 2 * np.pi, angles, endpoint=False)
        else:
            angles = ts.types.to_scalars(angles, var_name="angles")
        if len(angles) == 0:
            raise ValueError(f"ConeGeometry expects non-empty array of angles; got {self.angles_original}")
        if src_orig_dist is None:
            raise ValueError("Expected `src_orig_dist` parameter. Got `None`. ")
        if src_det_dist is None:
            raise ValueError("Expected `src_det_dist` parameter. Got `None`. ")
        self._angles = angles
        self._size = tuple(ts.types.to_size2d(shape if size is None else size))
        self._src_orig_dist = float(src_orig_dist)
        self._src_det_dist = float(src_det_dist)

    def __repr__(self):
        with ts.utils.print_options():
            return (
                f"ts.cone(\n"
                f"    angles={repr(self.angles_original)},\n"
                f"    shape={self.det_shape},\n"
                f"    size={self.det_size},\n"
                f"    src_orig_dist={self._src_orig_dist},\n"
                f"    src_det_dist={self._src_det_dist},\n"
                f")"
            )

    def __eq__(self, other):
        if not isinstance(other, ConeGeometry):
            return False
        if len(self._angles) != len(other._angles) or self.det_shape != other.det_shape:
            return False
        return bool(
            np.all(np.abs(self._angles - other._angles) < ts.epsilon)
            and np.all(np.abs(np.array(self._size) - np.array(other._size)) < ts.epsilon)
            and abs(self._src_orig_dist - other._src_orig_dist) < ts.epsilon
            and abs(self._src_det_dist - other._src_det_dist) < ts.epsilon
        )

    def __getitem__(self, key):
        """Select angles.  Detector indexing needs the vector form."""
        if isinstance(key, tuple):
            raise ValueError(
                f"Expected 1 index to ConeGeometry, got {len(key)}. "
                f"Indexing on the detector plane is not supported, "
                f"since it might move the detector center. "
            )
        picked = np.atleast_1d(self._angles[key])
        return ConeGeometry(picked, self.det_shape, self.det_size, self._src_orig_dist, self._src_det_dist)

    def to_astra(self):
        spacing_v, spacing_u = np.array(self._size) / np.array(self.det_shape)
        return astra_compat.create_proj_geom_cone(
            spacing_u, spacing_v, *self.det_shape, self.angles, self._src_orig_dist,
            self._src_det_dist - self._src_orig_dist,
        )

    @staticmethod
    def from_astra(astra_pg):
        if astra_pg["type"] != "cone":
            raise ValueError("ConeGeometry.from_astra only supports 'cone' type astra geometries.")
        shape = (astra_pg["DetectorRowCount"], astra_pg["DetectorColCount"])
        spacing = (astra_pg["DetectorSpacingY"], astra_pg["DetectorSpacingX"])
        sod = astra_pg["DistanceOriginSource"]
        return ConeGeometry(
            angles=astra_pg["ProjectionAngles"],
            shape=shape,
            size=np.array(spacing) * np.array(shape),
            src_orig_dist=sod,
            src_det_dist=sod + astra_pg["DistanceOriginDetector"],
        )

    def to_vec(self):
        return ConeVectorGeometry.from_astra(astra_compat.geom_2vec(self.to_astra()))

    @property
    def src_orig_dist(self):
        return self._src_orig_dist

    @property
    def src_det_dist(self):
        return self._src_det_dist

    @property
    def src_pos(self):
        return self.to_vec().src_pos

    @property
    def ray_dir(self):
        raise NotImplementedError()

    def rescale_det(self, scale):
        sv, su = (int(s) for s in ts.types.to_size2d(scale))
        shape = (self.det_shape[0] // sv, self.det_shape[1] // su)
        return ConeGeometry(self.angles_original, shape, self.det_size, self._src_orig_dist, self._src_det_dist)

    def reshape(self, new_shape):
        new_shape = ts.types.to_shape2d(new_shape)
        return ConeGeometry(self.angles_original, new_shape, self.det_size, self._src_orig_dist, self._src_det_dist)

    def __rmul__(self, other):
        if not isinstance(other, Transform):
            return NotImplemented
        warnings.warn(
            "Converting cone geometry to vector geometry. Use `T * pg.to_vec()` to inhibit this warning. ",
            stacklevel=2,
        )
        return other * self.to_vec()
