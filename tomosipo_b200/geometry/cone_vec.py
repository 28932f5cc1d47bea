"""Arbitrarily oriented cone-beam geometry.

API mirror of the reference's ``tomosipo/geometry/cone_vec.py``.  ``to_astra``
(reference ``cone_vec.py:174-191``) is the geometry -> 12-column vector
conversion that feeds the projector.
"""
import numpy as np

import tomosipo_b200 as ts
from ..types import ToShape2D, ToVec
from .base_projection import _BeamVectorGeometry


def cone_vec(*, shape: ToShape2D, src_pos: ToVec, det_pos: ToVec, det_v: ToVec, det_u: ToVec):
    """Create an arbitrarily oriented cone-beam geometry.

    >>> ts.cone_vec(shape=10, src_pos=(0, -2, 0), det_pos=(0, 1, 0), det_v=(1, 0, 0), det_u=(0, 0, 1)).num_angles
    1
    """
    return ConeVectorGeometry(shape=shape, src_pos=src_pos, det_pos=det_pos, det_v=det_v, det_u=det_u)


def random_cone_vec():
    """A randomly transformed random circular cone geometry (unseeded)."""
    return ts.geometry.random_transform() * ts.geometry.random_cone().to_vec()


class ConeVectorGeometry(_BeamVectorGeometry):
    """Source position and detector (centre, v, u) per projection angle."""

    _is_cone = True
    _beam_name = "src_pos"
    _beam_label = "source position"
    _beam_is_point = True
    _astra_type = "cone_vec"
    _ctor_name = "cone_vec"

    def __init__(self, *, shape, src_pos, det_pos, det_v, det_u):
        self._init_vectors(shape, src_pos, det_pos, det_v, det_u)

    @property
    def _src_pos(self):
        return self._beam

    @property
    def src_pos(self):
        return np.copy(self._beam)

    @property
    def ray_dir(self):
        raise NotImplementedError()

    def _ray_through(self, points):
        return points - self._beam
