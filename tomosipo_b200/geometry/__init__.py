"""Geometry model (volumes, projection geometries, transforms); host-side fp64.

Same public names as the reference's ``tomosipo.geometry`` package: the
geometry classes, their ``random_*`` constructors (test helpers) and the
``is_*`` predicates.
"""
from . import (base_projection, cone, cone_vec, conversion, det_vec, parallel, parallel_vec, transform, volume,  # noqa: F401
               volume_vec)
from .base_projection import ProjectionGeometry, is_cone, is_parallel, is_projection
from .cone import ConeGeometry, random_cone
from .cone_vec import ConeVectorGeometry, random_cone_vec
from .parallel import ParallelGeometry, random_parallel
from .parallel_vec import ParallelVectorGeometry, random_parallel_vec
from .transform import Transform, random_transform
from .volume import VolumeGeometry, is_volume, random_volume
from .volume_vec import VolumeVectorGeometry, random_volume_vec

__all__ = [
    "ProjectionGeometry", "ConeGeometry", "ConeVectorGeometry", "ParallelGeometry", "ParallelVectorGeometry",
    "VolumeGeometry", "VolumeVectorGeometry", "Transform",
    "random_cone", "random_cone_vec", "random_parallel", "random_parallel_vec", "random_volume", "random_volume_vec",
    "random_transform", "is_projection", "is_cone", "is_parallel", "is_volume",
]
