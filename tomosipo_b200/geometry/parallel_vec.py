"""Arbitrarily oriented parallel-beam geometry.

API mirror of the reference's ``tomosipo/geometry/parallel_vec.py``; ``to_astra``
(reference ``parallel_vec.py:175-192``) emits the ``parallel3d_vec`` rows
``[ray_dir | det centre | u | v]`` consumed by the projector.
"""
import numpy as np

import tomosipo_b200 as ts
from ..types import ToShape2D, ToVec
from .base_projection import _BeamVectorGeometry


def parallel_vec(*, shape: ToShape2D, ray_dir: ToVec, det_pos: ToVec, det_v: ToVec, det_u: ToVec):
    """Create an arbitrarily oriented parallel-beam geometry.

    >>> ts.parallel_vec(shape=10, ray_dir=(0, 1, 0), det_pos=(0, 0, 0), det_v=(1, 0, 0), det_u=(0, 0, 1)).num_angles
    1
    """
    return ParallelVectorGeometry(shape, ray_dir, det_pos, det_v, det_u)


def random_parallel_vec():
    """A random parallel vector geometry (unseeded)."""
    n = int(np.random.uniform(1, 20))
    return parallel_vec(
        shape=np.random.uniform(10, 20, size=2).astype(int),
        ray_dir=np.random.normal(size=(n, 3)),
        det_pos=np.random.normal(size=(n, 3)),
        det_v=np.random.normal(size=(n, 3)),
        det_u=np.random.normal(size=(n, 3)),
    )


class ParallelVectorGeometry(_BeamVectorGeometry):
    """Ray direction and detector (centre, v, u) per projection angle."""

    _is_parallel = True
    _beam_name = "ray_dir"
    _beam_label = "ray direction"
    _beam_is_point = False
    _astra_type = "parallel3d_vec"
    _ctor_name = "parallel_vec"

    def __init__(self, shape, ray_dir, det_pos, det_v, det_u):
        self._init_vectors(shape, ray_dir, det_pos, det_v, det_u)

    @property
    def _ray_dir(self):
        return self._beam

    @property
    def src_pos(self):
        raise NotImplementedError()

    @property
    def ray_dir(self):
        return np.copy(self._beam)

    def _ray_through(self, points):
        return self._beam
