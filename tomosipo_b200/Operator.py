"""The tomographic operator ``A = ts.operator(vg, pg)``.

API mirror of the reference's ``tomosipo/Operator.py``: ``operator``,
``Operator`` (``__call__``/``_fp``/``_bp``/``T``/``domain``/``range``/
``domain_shape``/``range_shape``), ``BackprojectionOperator`` and
``to_astra_compatible_operator_geometry``.  The only difference is below the
``direct_fp`` / ``direct_bp`` calls, which reach libtsproj instead of ASTRA.
"""
import numpy as np

import tomosipo_b200 as ts
from .Data import Data
from .astra import create_astra_projector, direct_bp, direct_fp


def to_astra_compatible_operator_geometry(vg, pg):
    """Axis-align a vector volume by moving the projection geometry instead.

    The projector only handles axis-aligned volumes centred on the origin.  For
    a ``VolumeVectorGeometry`` both geometries are re-expressed in the volume's
    own (normalised) frame -- a rigid change of perspective that keeps voxel
    sizes (reference ``Operator.py:11-60``).
    """
    if isinstance(vg, ts.geometry.VolumeGeometry):
        return (vg, pg)
    if not isinstance(vg, ts.geometry.VolumeVectorGeometry):
        raise TypeError(f"Expected volume geometry. Got {type(vg)}. ")

    vg = vg.to_vec()
    unit = lambda x: x / ts.vector_calc.norm(x)[:, None]  # noqa: E731
    P = ts.from_perspective(pos=vg.pos, w=unit(vg.w), v=unit(vg.v), u=unit(vg.u))
    vg = P * vg
    pg = P * pg.to_vec()

    sz, sy, sx = vg.voxel_size
    assert np.allclose(vg.pos, 0.0)
    assert np.allclose(vg.w, (sz, 0, 0))
    assert np.allclose(vg.v, (0, sy, 0))
    assert np.allclose(vg.u, (0, 0, sx))
    return ts.volume(shape=vg.shape, pos=0, size=vg.size), pg


def operator(volume_geometry, projection_geometry, voxel_supersampling=1, detector_supersampling=1,
             additive=False):
    """Create a tomographic projection operator.

    Parameters
    ----------
    volume_geometry:
        domain of the operator (``VolumeGeometry`` or ``VolumeVectorGeometry``)
    projection_geometry:
        range of the operator (any projection geometry)
    voxel_supersampling: int
        sub-voxels per voxel edge used by the backprojection
    detector_supersampling: int
        rays per detector pixel edge used by the forward projection
    additive: bool
        accumulate into the output instead of overwriting it
    """
    return Operator(
        volume_geometry,
        projection_geometry,
        voxel_supersampling=voxel_supersampling,
        detector_supersampling=detector_supersampling,
        additive=additive,
    )


def _to_link(geometry, x):
    return x.link if isinstance(x, Data) else ts.link(geometry, x)


class Operator:
    """Linear operator from volumes ``(z, y, x)`` to projection stacks ``(v, angle, u)``."""

    def __init__(self, volume_geometry, projection_geometry, voxel_supersampling=1, detector_supersampling=1,
                 additive=False):
        super().__init__()
        self.volume_geometry = volume_geometry
        self.projection_geometry = projection_geometry
        self.astra_compat_vg, self.astra_compat_pg = to_astra_compatible_operator_geometry(
            volume_geometry, projection_geometry
        )
        self.astra_projector = create_astra_projector(
            self.astra_compat_vg,
            self.astra_compat_pg,
            voxel_supersampling=voxel_supersampling,
            detector_supersampling=detector_supersampling,
        )
        self.additive = additive
        self._transpose = BackprojectionOperator(self)

    def _apply(self, forward, x, out):
        """Shared body of ``_fp`` / ``_bp``: link, allocate, project, unwrap."""
        in_geom, out_geom = (
            (self.astra_compat_vg, self.astra_compat_pg) if forward else (self.astra_compat_pg, self.astra_compat_vg)
        )
        src = _to_link(in_geom, x)
        if out is not None:
            dst = _to_link(out_geom, out)
        else:
            out_shape = self.range_shape if forward else self.domain_shape
            dst = src.new_zeros(out_shape) if self.additive else src.new_empty(out_shape)
        if forward:
            direct_fp(self.astra_projector, src, dst, additive=self.additive)
        else:
            direct_bp(self.astra_projector, dst, src, additive=self.additive)
        if isinstance(x, Data):
            return ts.data(self.projection_geometry if forward else self.volume_geometry, dst.data)
        return dst.data

    def _fp(self, volume, out=None):
        return self._apply(True, volume, out)

    def _bp(self, projection, out=None):
        return self._apply(False, projection, out)

    def __call__(self, volume, out=None):
        """Forward-project ``volume`` (array or ``Data``), optionally into ``out``."""
        return self._fp(volume, out)

    def transpose(self):
        return self._transpose

    @property
    def T(self):
        """The backprojection operator (always the same object)."""
        return self.transpose()

    @property
    def domain(self):
        return self.volume_geometry

    @property
    def range(self):
        return self.projection_geometry

    @property
    def domain_shape(self):
        return ts.links.geometry_shape(self.astra_compat_vg)

    @property
    def range_shape(self):
        return ts.links.geometry_shape(self.astra_compat_pg)


class BackprojectionOperator:
    """Transpose view of an operator; holds only a reference to its parent.

    >>> A = ts.operator(ts.volume(shape=10), ts.parallel(angles=10, shape=10))
    >>> A.T is A.T.T.T
    True
    """

    def __init__(self, parent):
        super().__init__()
        self.parent = parent

    def __call__(self, projection, out=None):
        """Back-project ``projection`` (array or ``Data``), optionally into ``out``."""
        return self.parent._bp(projection, out)

    def transpose(self):
        return self.parent

    @property
    def T(self):
        return self.transpose()

    @property
    def domain(self):
        return self.parent.range

    @property
    def range(self):
        return self.parent.domain

    @property
    def domain_shape(self):
        return self.parent.range_shape

    @property
    def range_shape(self):
        return self.parent.domain_shape
