"""Concatenation of geometries / transforms along the time axis.

API mirror of the reference's ``tomosipo/geometry/concatenate.py``.
"""
from typing import Collection, Union

import numpy as np

import tomosipo_b200 as ts
from . import ProjectionGeometry, Transform, VolumeGeometry, VolumeVectorGeometry


def _stack(items, names):
    return {n: np.concatenate([getattr(i, n) for i in items]) for n in names}


def concatenate(items: Union[Collection[ProjectionGeometry], Collection[VolumeGeometry],
                             Collection[VolumeVectorGeometry], Collection[Transform]]):
    """Concatenate same-kind items; the result is always in vector form.

    >>> ts.concatenate([ts.translate((0, 0, 1)), ts.translate((0, 0, 2))]).num_steps
    2
    """
    if len(items) == 0:
        raise ValueError("ts.concatenate expected at least one argument. ")
    if all(isinstance(i, Transform) for i in items):
        return Transform(np.concatenate([i.matrix for i in items]))
    for test, ctor, beam in ((ts.geometry.is_parallel, ts.parallel_vec, "ray_dir"),
                             (ts.geometry.is_cone, ts.cone_vec, "src_pos")):
        if all(test(i) for i in items):
            if any(i.det_shape != items[0].det_shape for i in items):
                raise ValueError("Cannot concatenate geometries. Not all detector shapes are equal.")
            return ctor(shape=items[0].det_shape, **_stack(items, (beam, "det_pos", "det_v", "det_u")))
    if isinstance(items, VolumeGeometry):
        raise TypeError("items must be iterable. ")
    if all(ts.geometry.is_volume(i) for i in items):
        if any(i.shape != items[0].shape for i in items):
            raise ValueError("Cannot concatenate volumes. Not all shapes are equal.")
        return ts.volume_vec(shape=items[0].shape, **_stack(items, ("pos", "w", "v", "u")))
    raise TypeError(f"Concatenating objects of types {set(type(i) for i in items)} is not supported. ")
