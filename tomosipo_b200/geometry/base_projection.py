"""Common interface of projection geometries.

API mirror of the reference's ``tomosipo/geometry/base_projection.py``
(``ProjectionGeometry``, ``is_projection``, ``is_cone``, ``is_parallel``).

Two helper bases factor what the reference spells out per class:
``_BeamVectorGeometry`` (shared by ``cone_vec`` / ``parallel_vec``: a
``DetectorVectorGeometry`` plus one beam vector per angle) and
``_CircularGeometry`` (shared by ``cone`` / ``parallel``: everything is
answered by converting to the vector form).
"""
import numpy as np

import tomosipo_b200 as ts
from .. import vector_calc as vc
from .transform import Transform


def is_projection(g):
    """True for any projection geometry."""
    return isinstance(g, ProjectionGeometry)


def is_cone(g):
    return is_projection(g) and g.is_cone


def is_parallel(g):
    return is_projection(g) and g.is_parallel


class ProjectionGeometry(object):
    """Abstract projection geometry: a detector of ``det_shape`` pixels per angle."""

    _is_cone = False
    _is_parallel = False
    _is_vector = False

    def __init__(self, shape=1):
        height, width = ts.types.to_shape2d(shape)
        self._shape = (height, width)

    def __repr__(self):
        raise NotImplementedError()

    def __eq__(self, other):
        raise NotImplementedError()

    def __len__(self):
        return self.num_steps

    def to_astra(self):
        raise NotImplementedError()

    def to_vec(self):
        raise NotImplementedError()

    @property
    def is_cone(self):
        return self._is_cone

    @property
    def is_parallel(self):
        return self._is_parallel

    @property
    def is_vec(self):
        return self._is_vector

    @property
    def det_shape(self):
        """(rows, columns) = (V, U)."""
        return self._shape

    @property
    def num_angles(self):
        raise NotImplementedError()

    @property
    def num_steps(self):
        return self.num_angles

    @property
    def angles(self):
        raise NotImplementedError()

    @property
    def src_pos(self):
        raise NotImplementedError()

    @property
    def det_pos(self):
        raise NotImplementedError()

    @property
    def det_v(self):
        raise NotImplementedError()

    @property
    def det_u(self):
        raise NotImplementedError()

    @property
    def det_normal(self):
        raise NotImplementedError()

    @property
    def ray_dir(self):
        raise NotImplementedError()

    @property
    def det_size(self):
        raise NotImplementedError()

    @property
    def det_sizes(self):
        raise NotImplementedError()

    @property
    def corners(self):
        raise NotImplementedError()

    @property
    def lower_left_corner(self):
        raise NotImplementedError()

    def rescale_det(self, scale):
        raise NotImplementedError()

    def reshape(self, new_shape):
        raise NotImplementedError()

    def project_point(self, point):
        raise NotImplementedError()

    def __rmul__(self, other):
        if isinstance(other, Transform):
            raise NotImplementedError()
        return NotImplemented


class _BeamVectorGeometry(ProjectionGeometry):
    """A detector vector geometry plus one beam vector (source position or ray
    direction) per angle.  Subclasses set ``_beam_name`` / ``_beam_label`` /
    ``_astra_type`` / ``_ctor_name`` and say whether the beam vector is a point."""

    _is_vector = True
    _beam_name = None       # keyword of the constructor: "src_pos" | "ray_dir"
    _beam_label = None      # used in error messages
    _beam_is_point = None   # transforms as a point (source) or as a vector (ray)
    _astra_type = None
    _ctor_name = None

    def _init_vectors(self, shape, beam, det_pos, det_v, det_u):
        from . import det_vec as dv

        ProjectionGeometry.__init__(self, shape=shape)
        parts = [
            ts.types.to_vec(beam, self._beam_label),
            ts.types.to_vec(det_pos, "detector position"),
            ts.types.to_vec(det_v, "v axis"),
            ts.types.to_vec(det_u, "u axis"),
        ]
        try:
            parts = np.broadcast_arrays(*parts)
        except ValueError:
            shapes = [x.shape for x in parts]
            raise ValueError(
                f"Not all arguments {self._beam_name}, det_pos, det_v, det_u are the same shape. Got: {shapes}"
            )
        self._beam = parts[0]
        self._det_vec = dv.det_vec(shape, parts[1], parts[2], parts[3])

    def _rebuild(self, shape, beam, det):
        return type(self)(**{
            "shape": shape, self._beam_name: beam, "det_pos": det.det_pos, "det_v": det.det_v, "det_u": det.det_u
        })

    def __repr__(self):
        with ts.utils.print_options():
            return (
                f"ts.{self._ctor_name}(\n"
                f"    shape={self.det_shape},\n"
                f"    {self._beam_name}={repr(self._beam)},\n"
                f"    det_pos={repr(self._det_vec.det_pos)},\n"
                f"    det_v={repr(self._det_vec.det_v)},\n"
                f"    det_u={repr(self._det_vec.det_u)},\n"
                f")"
            )

    def __eq__(self, other):
        if type(other) is not type(self):
            return False
        if self._det_vec != other._det_vec:
            return False
        return bool(np.all(np.abs(self._beam - other._beam) < ts.epsilon))

    def __getitem__(self, key):
        """Index as ``pg[angle, v, u]``; detector slices move ``det_pos`` accordingly."""
        det = self._det_vec[key]
        first = key[0] if isinstance(key, tuple) else key
        return self._rebuild(det.det_shape, self._beam[first], det)

    def to_astra(self):
        """ASTRA ``*_vec`` dict: rows [beam | det centre | u | v], each reversed to (x, y, z)."""
        rows, cols = self.det_shape
        d = self._det_vec
        vectors = np.concatenate(
            [self._beam[:, ::-1], d._det_pos[:, ::-1], d._det_u[:, ::-1], d._det_v[:, ::-1]], axis=1
        )
        return {"type": self._astra_type, "DetectorRowCount": rows, "DetectorColCount": cols, "Vectors": vectors}

    @classmethod
    def from_astra(cls, astra_pg):
        if astra_pg["type"] != cls._astra_type:
            raise ValueError(
                f"{cls.__name__}.from_astra only supports '{cls._astra_type}' type astra geometries."
            )
        vec = np.asarray(astra_pg["Vectors"], dtype=np.float64)
        shape = (astra_pg["DetectorRowCount"], astra_pg["DetectorColCount"])
        return cls(**{
            "shape": shape,
            cls._beam_name: vec[:, 0:3][:, ::-1],
            "det_pos": vec[:, 3:6][:, ::-1],
            "det_u": vec[:, 6:9][:, ::-1],
            "det_v": vec[:, 9:12][:, ::-1],
        })

    def to_vec(self):
        return self

    def to_vol(self):
        """Thin vector volume that coincides with the detector."""
        return self._det_vec.to_vol()

    @property
    def num_angles(self):
        return self._det_vec.num_angles

    @property
    def angles(self):
        raise NotImplementedError()

    @property
    def det_pos(self):
        return self._det_vec.det_pos

    @property
    def det_v(self):
        return self._det_vec.det_v

    @property
    def det_u(self):
        return self._det_vec.det_u

    @property
    def det_normal(self):
        return self._det_vec.det_normal

    @property
    def det_size(self):
        return self._det_vec.det_size

    @property
    def det_sizes(self):
        return self._det_vec.det_sizes

    @property
    def corners(self):
        return self._det_vec.corners

    @property
    def lower_left_corner(self):
        return self._det_vec.lower_left_corner

    def rescale_det(self, scale):
        det = self._det_vec.rescale_det(scale)
        return self._rebuild(det.det_shape, np.copy(self._beam), det)

    def reshape(self, new_shape):
        det = self._det_vec.reshape(new_shape)
        return self._rebuild(new_shape, np.copy(self._beam), det)

    def _ray_through(self, points):
        """Direction of the ray through ``points`` (per angle)."""
        raise NotImplementedError()

    def project_point(self, point):
        """(num_angles, 2) detector coordinates (v, u), in pixels from the detector centre."""
        if np.isscalar(point):
            point = ts.types.to_pos(point)
        origin = ts.types.to_vec(point)
        det_pos, det_v, det_u = self._det_vec.det_pos, self.det_v, self.det_u
        hit = vc.intersect(origin, self._ray_through(origin), det_pos, self.det_normal)
        iu = vc.dot(hit - det_pos, det_u) / vc.squared_norm(det_u)
        iv = vc.dot(hit - det_pos, det_v) / vc.squared_norm(det_v)
        return np.stack((iv, iu), axis=-1)

    def __rmul__(self, other):
        if not isinstance(other, Transform):
            return NotImplemented
        beam = other.transform_point(self._beam) if self._beam_is_point else other.transform_vec(self._beam)
        return self._rebuild(self.det_shape, beam, other * self._det_vec)


class _CircularGeometry(ProjectionGeometry):
    """Parametrised single-axis geometries: derived vectors come from ``to_vec()``."""

    def to_vol(self):
        return self.to_vec().to_vol()

    @property
    def num_angles(self):
        return len(self._angles)

    @property
    def angles(self):
        return np.copy(self._angles)

    @property
    def det_pos(self):
        return self.to_vec().det_pos

    @property
    def det_v(self):
        return self.to_vec().det_v

    @property
    def det_u(self):
        return self.to_vec().det_u

    @property
    def det_normal(self):
        return self.to_vec().det_normal

    @property
    def det_size(self):
        return self._size

    @property
    def det_sizes(self):
        return np.repeat([self._size], self.num_angles, axis=0)

    @property
    def corners(self):
        return self.to_vec().corners

    @property
    def lower_left_corner(self):
        return self.to_vec().lower_left_corner

    def project_point(self, point):
        return self.to_vec().project_point(point)
