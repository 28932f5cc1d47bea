"""Dispatch of ASTRA projection-geometry dicts to the matching class
(API mirror of the reference's ``tomosipo/geometry/conversion.py``)."""
from .cone import ConeGeometry
from .cone_vec import ConeVectorGeometry
from .det_vec import DetectorVectorGeometry
from .parallel import ParallelGeometry
from .parallel_vec import ParallelVectorGeometry

_BY_TYPE = {
    "cone": ConeGeometry,
    "cone_vec": ConeVectorGeometry,
    "det_vec": DetectorVectorGeometry,
    "parallel3d_vec": ParallelVectorGeometry,
    "parallel3d": ParallelGeometry,
}


def from_astra_projection_geometry(astra_pg):
    try:
        cls = _BY_TYPE[astra_pg["type"]]
    except KeyError:
        raise ValueError("ProjectionGeometry.from_astra only supports 3d astra geometries")
    return cls.from_astra(astra_pg)
