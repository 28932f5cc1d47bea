"""Legacy ``Data`` wrapper: a geometry plus a linked array.

API mirror of the reference's ``tomosipo/Data.py``.  The reference registers
the array with ``astra.data3d.link`` and keeps the id; there is no id registry
here -- the link itself is what the projector consumes.
"""
import tomosipo_b200 as ts


def data(geometry, initial_value=None):
    """Create a dataset on ``geometry``; an existing matching ``Data`` is returned as is."""
    if isinstance(initial_value, Data):
        if geometry == initial_value.geometry:
            return initial_value
        raise ValueError(f"Got initial_value={initial_value}, but its geometry does not match {geometry}.")
    return Data(geometry, initial_value)


class Data(object):
    """A geometry together with the array that lives on it."""

    def __init__(self, geometry, initial_value=None):
        super().__init__()
        self.geometry = geometry
        if not hasattr(geometry, "to_astra"):
            raise TypeError(
                f"Cannot create data object with geometry because it is not convertible to ASTRA: {geometry}"
            )
        self.astra_geom = geometry.to_astra()
        if not (self.is_volume() or self.is_projection()):
            raise ValueError(
                f"Geometry '{type(geometry)}' is not supported. Cannot determine if volume or projection geometry."
            )
        self._link = ts.link(geometry, initial_value)
        self.astra_id = id(self)  # opaque, for code that logs it

    def clone(self):
        """New ``Data`` on the same geometry with a copy of the array."""
        return Data(self.geometry, self._link.clone().data)

    def __enter__(self):
        return self

    def __exit__(self, type, value, traceback):
        return None

    @property
    def data(self):
        """The underlying array (shared); projections are ordered (v, angle, u)."""
        return self._link.data

    @data.setter
    def data(self, val):
        self._link.data = val

    @property
    def link(self):
        return self._link

    def is_volume(self):
        return ts.geometry.is_volume(self.geometry)

    def is_projection(self):
        return ts.geometry.is_projection(self.geometry)

    def to_astra(self):
        return self.astra_id
