"""Circular parallel-beam geometry.

API mirror of the reference's ``tomosipo/geometry/parallel.py``.  ``to_astra``
(reference ``parallel.py:162-172``) emits the plain ``'parallel3d'`` dict and
``to_vec`` (``parallel.py:188-195``) its vector form.
"""
import warnings
from typing import Union

import numpy as np

import tomosipo_b200 as ts
from .. import astra_compat
from ..types import ToScalars, ToShape2D, ToSize2D
from .base_projection import _CircularGeometry
from .parallel_vec import ParallelVectorGeometry
from .transform import Transform


def parallel(*, angles: Union[int, ToScalars] = 1, shape: ToShape2D = 1, size: ToSize2D = None):
    """Create a circular parallel-beam geometry.

    An integer ``angles`` means that many equi-spaced angles on [0, pi).

    >>> ts.parallel(angles=3, shape=10).det_shape
    (10, 10)
    """
    return ParallelGeometry(angles, shape, size)


def random_parallel():
    """A random parallel geometry (unseeded)."""
    return parallel(
        angles=np.random.normal(size=20),
        size=np.random.uniform(10, 20, size=2),
        shape=np.random.uniform(10, 20, size=2).astype(int),
    )


class ParallelGeometry(_CircularGeometry):
    """Angles and detector shape/size; the detector is centred on the origin."""

    _is_parallel = True

    def __init__(self, angles=1, shape=1, size=None):
        super().__init__(shape=shape)
        self._angles_original = angles
        if np.isscalar(angles) and isinstance(angles, int):
            angles = np.linspace(0, np.pi, angles, endpoint=False)
        else:
            angles = ts.types.to_scalars(angles, var_name="angles")
        if len(angles) == 0:
            raise TypeError(f"ParallelGeometry expects non-empty array of angles; got {self._angles_original}")
        self._angles = angles
        self._size = ts.types.to_size2d(shape if size is None else size)

    def __repr__(self):
        with ts.utils.print_options():
            return (
                f"ts.parallel(\n"
                f"    angles={repr(self._angles_original)},\n"
                f"    shape={repr(self.det_shape)},\n"
                f"    size={repr(self._size)},\n"
                f")"
            )

    def __eq__(self, other):
        if not isinstance(other, ParallelGeometry):
            return False
        if self.det_shape != other.det_shape or len(self._angles) != len(other._angles):
            return False
        return bool(
            np.all(np.abs(self._angles - other._angles) < ts.epsilon)
            and np.all(np.abs(np.array(self._size) - np.array(other._size)) < ts.epsilon)
        )

    def __getitem__(self, key):
        """Select angles.  Detector indexing needs the vector form."""
        if isinstance(key, tuple):
            raise ValueError(
                f"Expected 1 index to ParallelGeometry, got {len(key)}. "
                f"Indexing on the detector plane is not supported, "
                f"since it might move the detector center. "
                f"To prevent this error, use `pg.to_vec()[a:b, c:d, e:f]'. "
            )
        return parallel(angles=np.atleast_1d(self._angles[key]), shape=self.det_shape, size=self._size)

    def to_astra(self):
        rows, cols = self.det_shape
        spacing_v, spacing_u = np.array(self._size) / np.array(self.det_shape)
        return {
            "type": "parallel3d",
            "DetectorSpacingX": spacing_u,
            "DetectorSpacingY": spacing_v,
            "DetectorRowCount": rows,
            "DetectorColCount": cols,
            "ProjectionAngles": np.copy(self._angles),
        }

    @staticmethod
    def from_astra(astra_pg):
        if astra_pg["type"] != "parallel3d":
            raise ValueError("ParallelGeometry.from_astra only supports 'parallel3d' type astra geometries.")
        shape = (astra_pg["DetectorRowCount"], astra_pg["DetectorColCount"])
        spacing = (astra_pg["DetectorSpacingY"], astra_pg["DetectorSpacingX"])
        return parallel(
            angles=np.copy(astra_pg["ProjectionAngles"]), shape=shape, size=np.array(spacing) * np.array(shape)
        )

    def to_vec(self):
        return ParallelVectorGeometry.from_astra(astra_compat.geom_2vec(self.to_astra()))

    @property
    def src_pos(self):
        raise NotImplementedError()

    @property
    def ray_dir(self):
        return self.to_vec().ray_dir

    def rescale_det(self, scale):
        sv, su = (int(s) for s in ts.types.to_size2d(scale))
        shape = (self.det_shape[0] // sv, self.det_shape[1] // su)
        return parallel(angles=np.copy(self._angles), shape=shape, size=self._size)

    def reshape(self, new_shape):
        return parallel(angles=np.copy(self._angles), shape=new_shape, size=self._size)

    def __rmul__(self, other):
        if not isinstance(other, Transform):
            return NotImplemented
        warnings.warn(
            "Converting parallel geometry to vector geometry. Use `T * pg.to_vec()` to inhibit this warning. ",
            stacklevel=2,
        )
        return other * self.to_vec()
