"""Array links and the Data wrapper (host logic; mirrors the reference's
tests/links/test_numpy.py, tests/test_Data.py, tests/test_torch_support.py:31-59)."""
import warnings

import numpy as np
import pytest
import torch

import tomosipo_b200 as ts
import tomosipo_b200.torch_support  # noqa: F401  (registers nothing extra; import parity with the reference)
from tomosipo_b200.links.numpy import NumpyLink
from tomosipo_b200.links.torch import TorchLink


def test_geometry_shape():
    assert ts.links.geometry_shape(ts.volume(shape=(2, 3, 4))) == (2, 3, 4)
    assert ts.links.geometry_shape(ts.parallel(angles=5, shape=(6, 7))) == (6, 5, 7)
    assert ts.links.geometry_shape(ts.cone(angles=5, shape=(6, 7), cone_angle=1).to_vec()) == (6, 5, 7)
    with pytest.raises(ValueError):
        ts.links.geometry_shape(object())


def test_numpy_link_shape_dtype_contiguity():
    vg = ts.volume(shape=(3, 4, 5))
    with pytest.raises(ValueError):
        ts.link(vg, np.zeros((3, 4, 6), dtype=np.float32))
    with pytest.warns(UserWarning, match="float32"):
        lk = ts.link(vg, np.zeros((3, 4, 5), dtype=np.float64))
    assert lk.data.dtype == np.float32
    with pytest.warns(UserWarning, match="C_CONTIGUOUS"):
        ts.link(vg, np.zeros((5, 4, 3), dtype=np.float32).T)
    x = np.ones((3, 4, 5), dtype=np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        lk = ts.link(vg, x)
    assert lk.data is x                       # shared, not copied
    assert ts.link(vg, None).data.sum() == 0  # default: zeros
    assert ts.link(vg, 3.0).data.mean() == 3.0
    assert lk.linked_data.ptr == x.ctypes.data and lk.linked_data.kind == "host"
    with pytest.raises(AttributeError):
        lk.data = x
    with pytest.raises(ValueError):
        ts.link(vg, "nope")


def test_new_arrays_and_clone():
    vg = ts.volume(shape=(2, 2, 2))
    for lk in (ts.link(vg, np.ones((2, 2, 2), np.float32)), ts.link(vg, torch.ones(2, 2, 2))):
        assert float(lk.new_zeros((1, 2, 3)).data.sum()) == 0.0
        assert float(lk.new_full((1, 2, 3), 2.0).data.sum()) == 12.0
        assert tuple(lk.new_empty((1, 2, 3)).data.shape) == (1, 2, 3)
        c = lk.clone()
        c.data[:] = 5
        assert float(lk.data.sum()) == 8.0


def test_torch_link_conversions():
    vg = ts.volume(shape=(2, 3, 4))
    with pytest.warns(UserWarning, match="float32"):
        lk = ts.link(vg, torch.zeros(2, 3, 4, dtype=torch.float64))
    assert lk.data.dtype == torch.float32
    with pytest.warns(UserWarning, match="contiguous"):
        ts.link(vg, torch.zeros(4, 3, 2).permute(2, 1, 0))
    with pytest.raises(ValueError):
        ts.link(vg, torch.zeros(2, 3, 5))
    assert isinstance(ts.link(vg, torch.tensor(2.0)), TorchLink)
    assert float(ts.link(vg, torch.tensor(2.0)).data.mean()) == 2.0


def test_compatibility_rules():
    vg = ts.volume(shape=(2, 2, 2))
    a = ts.link(vg, np.zeros((2, 2, 2), np.float32))
    b = ts.link(vg, torch.zeros(2, 2, 2))
    assert ts.links.are_compatible(a, a) and ts.links.are_compatible(a, b) and ts.links.are_compatible(b, a)
    assert isinstance(a, NumpyLink)


def test_data_wrapper():
    vg = ts.volume(shape=(2, 3, 4))
    pg = ts.parallel(angles=5, shape=(6, 7))
    vd, pd = ts.data(vg), ts.data(pg)
    assert vd.data.shape == (2, 3, 4) and pd.data.shape == (6, 5, 7)
    assert vd.is_volume() and pd.is_projection() and not vd.is_projection()
    assert ts.data(vg, vd) is vd
    with pytest.raises(ValueError):
        ts.data(ts.volume(shape=3), vd)
    with pytest.warns(UserWarning):
        ts.data(vg, np.zeros((2, 3, 4), dtype=np.float64))
    with pytest.raises(TypeError):
        ts.data(object())
    c = vd.clone()
    c.data[:] = 1
    assert vd.data.sum() == 0
    with ts.data(vg, 1.0) as d:
        assert d.data.mean() == 1.0


def test_direct_project_checks():
    A = ts.operator(ts.volume(shape=4), ts.parallel(angles=3, shape=4))
    v = ts.link(A.astra_compat_vg, np.zeros((4, 4, 4), np.float32))
    p = ts.link(A.astra_compat_pg, np.zeros((4, 3, 4), np.float32))
    with pytest.raises(ValueError, match="forward"):
        ts.astra.direct_project(A.astra_projector, v, p)
    with pytest.raises(ValueError, match="expects"):
        ts.astra.direct_project(A.astra_projector, p, v, forward=True)


def test_operator_surface():
    vg, pg = ts.volume(shape=10), ts.parallel(angles=10, shape=10)
    A = ts.operator(vg, pg)
    assert A.T is A.T.T.T and A.T.T is A
    assert A.domain is vg and A.range is pg and A.T.domain is pg and A.T.range is vg
    assert A.domain_shape == (10, 10, 10) and A.range_shape == (10, 10, 10)
    assert A.T.domain_shape == A.range_shape and A.T.range_shape == A.domain_shape
    assert not A.additive
    with pytest.raises(TypeError):
        ts.operator(pg, pg)


def test_vector_volume_is_unrotated():
    # reference Operator.py:11-60: rotate the detector instead of the volume
    R = ts.rotate(pos=0, axis=(1, 0, 0), angles=0.4)
    T = ts.translate((0.2, -0.1, 0.3))
    vg = ts.volume(shape=(4, 5, 6), size=(2, 2.5, 3))
    pg = ts.cone(angles=7, shape=(8, 9), size=(4, 5), src_orig_dist=5, src_det_dist=9)
    A = ts.operator(T * R * vg.to_vec(), pg)
    assert isinstance(A.astra_compat_vg, ts.geometry.VolumeGeometry)
    assert A.astra_compat_vg == ts.volume(shape=(4, 5, 6), pos=0, size=(2, 2.5, 3))
    assert A.astra_compat_pg == (T * R).inv * pg.to_vec()
    assert A.domain_shape == (4, 5, 6) and A.range_shape == (8, 7, 9)


def test_cupy_link_against_a_stand_in_module(monkeypatch):
    """CuPy is absent from the image: exercise CupyLink's own logic (reference links/cupy.py:25-159: coercion
    warnings, shape checks, pointer export, same-device allocation) against a numpy-backed stand-in for the
    handful of cupy names it touches."""
    import contextlib
    import importlib
    import sys
    import types

    class _Dev:
        def __init__(self, id_=0):
            self.id = id_

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def __eq__(self, other):
            return isinstance(other, _Dev) and other.id == self.id

    class FakeArray:
        """The slice of cupy.ndarray the link uses, over a numpy array (not a subclass: numpy's link must not take it)."""
        device = _Dev(0)

        def __init__(self, a):
            self.a = a

        shape = property(lambda self: self.a.shape)
        dtype = property(lambda self: self.a.dtype)
        flags = property(lambda self: {"C_CONTIGUOUS": self.a.flags["C_CONTIGUOUS"]})
        data = property(lambda self: types.SimpleNamespace(ptr=self.a.ctypes.data))  # cupy: MemoryPointer

        def astype(self, dt):
            return FakeArray(self.a.astype(dt))

        def copy(self):
            return FakeArray(self.a.copy())

        def __setitem__(self, k, v):
            self.a[k] = v.a if isinstance(v, FakeArray) else v

    def wrap(a):
        return FakeArray(np.asarray(a))

    fake = types.ModuleType("cupy")
    fake.ndarray = FakeArray
    fake.float32 = np.float32
    fake.zeros = lambda shape, dtype=np.float32: wrap(np.zeros(shape, dtype))
    fake.empty = lambda shape, dtype=np.float32: wrap(np.empty(shape, dtype))
    fake.full = lambda shape, value, dtype=np.float32: wrap(np.full(shape, value, dtype))
    fake.ascontiguousarray = lambda a: wrap(np.ascontiguousarray(a.a))
    fake.cuda = types.SimpleNamespace(get_current_stream=lambda: types.SimpleNamespace(ptr=1234))
    monkeypatch.setitem(sys.modules, "cupy", fake)
    sys.modules.pop("tomosipo_b200.links.cupy", None)
    n_backends = len(ts.links.base.backends)
    try:
        mod = importlib.import_module("tomosipo_b200.links.cupy")
        CupyLink = mod.CupyLink
        vg = ts.volume(shape=(2, 3, 4))
        a = wrap(np.arange(24, dtype=np.float32).reshape(2, 3, 4))
        lk = ts.link(vg, a)
        assert isinstance(lk, CupyLink) and lk.data is a
        rb = lk.linked_data
        assert (rb.ptr, rb.shape, rb.kind, rb.device, rb.stream) == (a.a.ctypes.data, (2, 3, 4), "device", 0, 1234)
        with pytest.raises(ValueError):
            CupyLink((2, 3, 5), a)
        with pytest.raises(ValueError):
            CupyLink((2, 3, 4), np.zeros((2, 3, 4), np.float32))          # a plain ndarray is not a cupy array
        with pytest.warns(UserWarning, match="float32"):
            assert CupyLink((2, 3, 4), wrap(np.zeros((2, 3, 4)))).data.dtype == np.float32
        with pytest.warns(UserWarning, match="contiguous"):
            t = CupyLink((2, 3, 4), wrap(np.zeros((4, 3, 2), np.float32).T))
        assert t.data.flags["C_CONTIGUOUS"]
        s0 = CupyLink((2, 3, 4), wrap(np.float32(3.0)))                       # scalars fill a new array
        assert s0.data.shape == (2, 3, 4) and float(s0.data.a.min()) == 3.0
        z, f, e, c = lk.new_zeros((1, 2, 3)), lk.new_full((1, 2, 3), 2.5), lk.new_empty((3, 2, 1)), lk.clone()
        assert z.data.shape == (1, 2, 3) and float(z.data.a.sum()) == 0 and float(f.data.a.mean()) == 2.5
        assert e.data.shape == (3, 2, 1) and c.data is not a and np.array_equal(c.data.a, a.a)
        with pytest.raises(AttributeError):
            lk.data = a
        assert lk.__compatible_with__(c) is True and lk.__compatible_with__(ts.link(vg, np.zeros((2, 3, 4), np.float32))) is NotImplemented
        with lk.context():
            pass
    finally:
        del ts.links.base.backends[n_backends:]
        sys.modules.pop("tomosipo_b200.links.cupy", None)
