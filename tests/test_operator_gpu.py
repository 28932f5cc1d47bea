"""The public operator API on the GPU: the reference's own operator / autograd
tests (tests/test_Operator.py, tests/test_torch_support.py) restated against
this package, value parity against the oracle, the ASTRA known answer, and
size-independent properties at the full benchmark size."""
import numpy as np
import pytest
import torch

import tomosipo_b200 as ts
from tomosipo_b200.torch_support import AutogradOperator, autograd_operator, to_autograd
from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def oracle_of(A):
    """Oracle projector for the ASTRA-compatible geometries of an operator."""
    avg, apg = A.astra_compat_vg.to_astra(), A.astra_compat_pg.to_vec().to_astra()
    o = avg["option"]
    kind = O.CONE_VEC if apg["type"] == "cone_vec" else O.PARALLEL_VEC
    return O.OracleProjector(
        kind, A.domain_shape, [o["WindowMinX"], o["WindowMinY"], o["WindowMinZ"]],
        [o["WindowMaxX"], o["WindowMaxY"], o["WindowMaxZ"]], (A.range_shape[0], A.range_shape[2]), apg["Vectors"])


# ---------------------------------------------------------------- test_Operator.py
def test_operator_data_equals_array():
    pg = ts.cone(angles=150, shape=(100, 100), size=(4, 4), src_orig_dist=4, src_det_dist=8)   # as tests/test_Operator.py:15-32
    vg = ts.volume(shape=100)
    A = ts.operator(vg, pg)
    vd = ts.phantom.hollow_box(ts.data(vg))
    pd = A(vd)
    assert isinstance(pd, ts.Data.Data) and pd.geometry is pg
    np.testing.assert_array_equal(pd.data, A(vd.data))
    np.testing.assert_array_equal(A.T(pd).data, A.T(pd.data))
    assert A.T(pd).geometry is vg


def test_operator_additive():
    pg = ts.cone(angles=33, shape=(24, 32), size=(3, 4), src_orig_dist=5, src_det_dist=9)
    vg = ts.volume(shape=(20, 24, 28), size=(1, 1.2, 1.4))
    A, B = ts.operator(vg, pg, additive=False), ts.operator(vg, pg, additive=True)
    x = ts.phantom.hollow_box(ts.data(vg)).data
    y = np.zeros(A.range_shape, np.float32)
    B(x, out=y); B(x, out=y)
    assert np.allclose(y, 2 * A(x), rtol=1e-5, atol=1e-6)
    z = np.zeros(A.domain_shape, np.float32)
    B.T(y, out=z); B.T(y, out=z)
    assert np.allclose(z, 2 * A.T(y), rtol=1e-5, atol=1e-5)
    assert np.allclose(B(x), A(x))    # fresh output of an additive operator starts from zeros


@pytest.mark.parametrize("pg", [
    ts.cone(angles=12, shape=(20, 24), size=(3, 3.6), src_orig_dist=6, src_det_dist=10),
    ts.parallel(angles=12, shape=(20, 24), size=(3, 3.6)),
])
@pytest.mark.parametrize("scale", [(1, 1, 1), (1.5, 1.0, 0.75)])
def test_operator_volume_vector(pg, scale):
    # rotated / translated / scaled vector volume == inverse-transformed detector (tests/test_Operator.py:52-83)
    T = ts.translate((0.1, -0.2, 0.15))
    R = ts.rotate(pos=0, axis=(0.3, 1.0, -0.4), angles=0.6)
    S = ts.scale(scale)
    vg = ts.volume(shape=(16, 18, 20), size=(1.6, 1.8, 2.0))
    M = T * R * S
    A1 = ts.operator(M * vg.to_vec(), pg)
    A2 = ts.operator(S * vg, (T * R).inv * pg.to_vec())
    x = np.random.default_rng(0).random(vg.shape).astype(np.float32)
    assert rel_l2(A1(x), A2(x)) < 1e-5
    assert rel_l2(A1.T(A1(x)), A2.T(A2(x))) < 1e-5


# ------------------------------------------------------------------ value parity
@pytest.mark.parametrize("make", [
    lambda: (ts.volume(shape=(24, 32, 40), pos=(0.1, 0.2, -0.1), size=(1.2, 1.6, 2.0)),
             ts.cone(angles=21, shape=(28, 36), size=(3, 4), src_orig_dist=5, src_det_dist=8)),
    lambda: (ts.volume(shape=(24, 32, 40), size=(1.2, 1.6, 2.0)),
             ts.parallel(angles=np.linspace(0, 2 * np.pi, 17, endpoint=False), shape=(28, 36), size=(3, 4))),
    lambda: (ts.volume(shape=64, size=1)[:1],
             ts.parallel(angles=48, shape=(64, 96), size=(1, 1.5)).to_vec()[:, :1, :]),
])
def test_operator_matches_oracle(make):
    vg, pg = make()
    A = ts.operator(vg, pg)
    Q = oracle_of(A)
    rng = np.random.default_rng(0)
    x = rng.random(A.domain_shape).astype(np.float32)
    y = rng.random(A.range_shape).astype(np.float32)
    assert rel_l2(A(torch.from_numpy(x).cuda()).cpu().numpy(), Q.fp(x)) < TOL
    assert rel_l2(A.T(torch.from_numpy(y).cuda()).cpu().numpy(), Q.bp(y)) < TOL
    assert rel_l2(A(x), Q.fp(x)) < TOL          # numpy (host) path
    assert rel_l2(A.T(y), Q.bp(y)) < TOL


def test_astra_known_answer_on_gpu():
    # notebooks/cupy.ipynb cell 4 of the reference (real ASTRA output): mean 0.109146185, max 0.8471311
    vg = ts.volume(shape=128, size=1)
    pg = ts.parallel(angles=np.linspace(0, 2 * np.pi, 180, endpoint=False), shape=128, size=np.sqrt(2))
    A = ts.operator(vg, pg)
    x = torch.from_numpy(ts.phantom.hollow_box(ts.data(vg)).data).cuda()
    y = A(x)
    assert abs(float(y.double().mean()) - 0.109146185) / 0.109146185 < 5e-6
    assert abs(float(y.max()) - 0.8471311) / 0.8471311 < 5e-6
    assert float(y.min()) == 0.0


def test_readme_sirt_matches_oracle():
    # README.md:139-164 of the reference at reduced size: same iteration, float64 numpy arrays in
    n = 32
    pg = ts.cone(size=np.sqrt(2), cone_angle=1 / 2, angles=25, shape=(n, 48))
    vg = ts.volume(shape=n)
    A = ts.operator(vg, pg)
    Q = oracle_of(A)
    phantom = np.zeros(A.domain_shape); phantom[5:12, 5:12, 5:12] = 1.0
    with pytest.warns(UserWarning):
        R = 1 / A(np.ones(A.domain_shape))
    R = np.minimum(R, 1 / ts.epsilon)
    with pytest.warns(UserWarning):
        C = 1 / A.T(np.ones(A.range_shape))
    C = np.minimum(C, 1 / ts.epsilon)
    Rq = np.minimum(1 / Q.fp(np.ones(A.domain_shape)), 1 / ts.epsilon)
    Cq = np.minimum(1 / Q.bp(np.ones(A.range_shape)), 1 / ts.epsilon)
    assert rel_l2(R, Rq) < 1e-4 and rel_l2(C, Cq) < 1e-4
    with pytest.warns(UserWarning):
        y = A(phantom)
        x = np.zeros(A.domain_shape)
        xq = np.zeros(A.domain_shape)
        yq = Q.fp(phantom)
        for _ in range(5):
            x += C * A.T(R * (y - A(x)))
            xq += Cq * Q.bp(Rq * (yq - Q.fp(xq)))
    assert rel_l2(x, xq) < 1e-4


# ------------------------------------------------------------ test_torch_support.py
@pytest.mark.parametrize("device", ["cpu", "cuda"])
def test_fp_bp_torch(device):
    A = ts.operator(ts.volume(shape=10), ts.parallel(angles=10, shape=10))
    x = torch.ones(A.domain_shape, device=device)
    y = A(x)
    assert y.device.type == device and y.dtype == torch.float32 and float(y.sum()) > 1
    bp = A.T(y)
    assert bp.device.type == device and float(bp.sum()) > 1
    out = torch.empty(A.range_shape, device=device)
    assert A(x, out=out) is out and torch.equal(out, y)


def test_mixed_devices_raise():
    A = ts.operator(ts.volume(shape=10), ts.parallel(angles=10, shape=10))
    with pytest.raises(ValueError, match="not compatible"):
        A(torch.ones(A.domain_shape).cuda(), out=torch.ones(A.range_shape))


def test_float64_input():
    A = ts.operator(ts.volume(shape=10), ts.parallel(angles=10, shape=10))
    x = torch.ones(A.domain_shape, dtype=torch.float64, device="cuda")
    with pytest.warns(UserWarning):
        y = A(x)
    with pytest.warns(UserWarning):
        y2 = to_autograd(A)(x)
    assert y.dtype == torch.float32 and torch.equal(y, y2)


def test_autograd_gradient_is_transpose():
    A = ts.operator(ts.volume(shape=10), ts.parallel(angles=10, shape=10))
    f = to_autograd(A)
    x = torch.rand(A.domain_shape, device="cuda", requires_grad=True)
    y = f(x)
    y.backward(y)
    assert torch.allclose(x.grad, A.T(A(x.detach())))
    g = to_autograd(A.T)
    p = torch.rand(A.range_shape, device="cuda", requires_grad=True)
    q = g(p)
    q.backward(q)
    assert torch.allclose(p.grad, A(A.T(p.detach())))


@pytest.mark.parametrize("extra", [(1, 1), (2, 3)])
def test_autograd_extra_dims_match_loop(extra):
    A = ts.operator(ts.volume(shape=(1, 12, 12)), ts.parallel(angles=9, shape=(1, 16)))
    f = to_autograd(A, num_extra_dims=2, is_2d=True)
    x = torch.rand(*extra, 12, 12, device="cuda", requires_grad=True)
    y = f(x)
    assert y.shape == (*extra, 9, 16)
    for i in range(extra[0]):
        for j in range(extra[1]):
            assert torch.equal(y[i, j], A(x.detach()[i, j][None])[0])      # batched call == per-item calls
    y.sum().backward()
    ref = A.T(torch.ones(A.range_shape, device="cuda"))[0]
    assert torch.allclose(x.grad[0, 0], ref)
    with pytest.raises(AssertionError):
        f(torch.rand(12, 12, device="cuda"))
    # non-contiguous input falls back to the per-item loop with the same values
    xt = torch.rand(*extra, 12, 12, device="cuda").transpose(-1, -2)
    with pytest.warns(UserWarning):
        yt = f(xt)
    assert torch.allclose(yt, f(xt.contiguous()))
    # the expanded (stride-0) gradient of a sum takes the batched path after one copy
    xg = torch.rand(*extra, 12, 12, device="cuda", requires_grad=True)
    with pytest.warns(UserWarning):
        f(xg).sum().backward()
    g_ref = to_autograd(A.T, num_extra_dims=len(extra), is_2d=True)(torch.ones_like(f(xg.detach())))
    assert torch.allclose(xg.grad, g_ref)


def test_autograd_operator():
    vg, pg = ts.volume(shape=10), ts.parallel(angles=10, shape=10)
    A = ts.operator(vg, pg)
    B = autograd_operator(vg, pg)
    assert isinstance(B, AutogradOperator) and B.domain_shape == A.domain_shape and B.T.T is B
    x = torch.rand(A.domain_shape, device="cuda")
    assert torch.equal(A(x), B(x)) and torch.equal(A.T(A(x)), B.T(B(x)))
    out = torch.zeros(A.range_shape, device="cuda")
    assert B(x, out=out) is out and torch.equal(out, A(x))
    with pytest.raises(ValueError):
        AutogradOperator(ts.operator(vg, pg, additive=True))
    p = torch.rand(A.range_shape, device="cuda", requires_grad=True)
    q = B.T(p)
    q.backward(q)
    assert torch.allclose(p.grad, A(A.T(p.detach())))


def test_legacy_data_projection():
    # tests/test_astra.py:48-84 of the reference (runs, accumulates all-to-all)
    vg, pg = ts.volume(shape=16), ts.parallel(angles=8, shape=16)
    vd, pd = ts.data(vg, 1.0), ts.data(pg)
    ts.astra.forward(vd, pd)
    assert pd.data.sum() > 1
    vd2 = ts.data(vg)
    ts.astra.backward(vd2, pd, voxel_supersampling=2)
    assert vd2.data.sum() > 1


def test_stream_semantics():
    # device arrays are projected on the current torch stream, not on the legacy default stream
    A = ts.operator(ts.volume(shape=32), ts.parallel(angles=16, shape=32))
    x = torch.rand(A.domain_shape, device="cuda")
    ref = A(x)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        y = A(x)
    s.synchronize()
    assert torch.equal(y, ref)


def test_concurrent_host_threads_share_a_projector():
    """One projector, two host threads, two streams: device calls are independent (per-call scratch from the
    stream-ordered pool, descriptors by value, immutable geometry tables)."""
    import threading

    A = ts.operator(ts.volume(shape=(64, 64, 64), size=1),
                    ts.cone(angles=48, shape=(64, 96), size=(1.875, 2.8125), src_orig_dist=4, src_det_dist=6))
    g = torch.Generator(device="cuda").manual_seed(0)
    xs = [torch.rand(A.domain_shape, device="cuda", generator=g) for _ in range(2)]
    refs = [(A(x), A.T(A(x))) for x in xs]
    torch.cuda.synchronize()
    out, errors = [None, None], []

    def work(i):
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(20):
                    y = A(xs[i])
                    xb = A.T(y)
            s.synchronize()
            out[i] = (y, xb)
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors
    for i in range(2):
        assert torch.equal(out[i][0], refs[i][0]) and torch.equal(out[i][1], refs[i][1])


def test_device_calls_capture_into_a_cuda_graph():
    """FP + BP on device tensors are plain stream work (kernel launches with by-value TMA descriptors,
    stream-ordered scratch): they capture into a CUDA graph and replay on new data."""
    vg = ts.volume(shape=(48, 64, 64), size=(0.75, 1, 1))
    for pg in (ts.cone(angles=40, shape=(48, 96), size=(1.5, 3.0), src_orig_dist=4, src_det_dist=6),      # TMA kernels
               ts.parallel(angles=33, shape=(2, 96), size=(2 / 48, 1.5)).to_vec()):                         # thin FP
        A = ts.operator(vg, pg)
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.rand(A.domain_shape, device="cuda", generator=g)
        y = torch.empty(A.range_shape, device="cuda")
        xb = torch.empty(A.domain_shape, device="cuda")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                          # warm-up outside the capture (device tables, attributes)
            A(x, out=y); A.T(y, out=xb)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            A(x, out=y); A.T(y, out=xb)
        x.copy_(torch.rand(A.domain_shape, device="cuda", generator=g))
        y.zero_(); xb.zero_()
        graph.replay()
        torch.cuda.synchronize()
        y_ref = A(x)
        assert torch.equal(y, y_ref) and torch.equal(xb, A.T(y_ref))


# ----------------------------------------------- properties at the benchmark size
@pytest.mark.slow
def test_full_size_properties():
    n = 512
    vg = ts.volume(shape=n, size=1)
    pg = ts.cone(angles=720, shape=(n, 768), size=(1.875, 2.8125), src_orig_dist=4, src_det_dist=6).to_vec()
    A = ts.operator(vg, pg)
    g = torch.Generator(device="cuda").manual_seed(0)
    x1 = torch.rand(A.domain_shape, device="cuda", generator=g)
    x2 = torch.rand(A.domain_shape, device="cuda", generator=g)
    y1, y2 = A(x1), A(x2)
    lin = A(2.0 * x1 - 0.5 * x2)
    err = torch.linalg.vector_norm(lin - (2.0 * y1 - 0.5 * y2)) / torch.linalg.vector_norm(lin)
    assert float(err) < 1e-5                                   # linearity
    w = torch.rand(A.range_shape, device="cuda", generator=g)
    lhs = torch.sum(y1.double() * w.double())
    rhs = torch.sum(x1.double() * A.T(w).double())
    assert abs(float(lhs / rhs) - 1) < 0.01                    # scaled adjoint
    ones = A(torch.ones(A.domain_shape, device="cuda"))
    assert abs(float(ones[n // 2, 0, 384]) - 1.0) < 1e-3       # central chord of the unit cube
    assert abs(float(ones[n // 2 - 1, 180, 383]) - 1.0) < 1e-3
    out = torch.zeros_like(ones)
    assert torch.equal(A(torch.ones(A.domain_shape, device="cuda"), out=out), ones)  # deterministic


def test_full_size_properties_1024():
    """BASELINE configs[3] (cone 1024^3, 1440 angles, 1024 x 1536) on one GPU: size-independent properties."""
    n = 1024
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * 2 ** 30:
        pytest.skip("needs ~50 GB of device memory")
    vg = ts.volume(shape=n, size=1)
    pg = ts.cone(angles=1440, shape=(n, 1536), size=(1.875, 2.8125), src_orig_dist=4, src_det_dist=6)
    A = ts.operator(vg, pg)
    assert A.astra_projector.info().n_march_x + A.astra_projector.info().n_march_y == 1440
    ones = A(torch.ones(A.domain_shape, device="cuda"))
    assert abs(float(ones[n // 2, 0, 768]) - 1.0) < 1e-3       # central chord of the unit cube
    assert abs(float(ones[n // 2 - 1, 360, 767]) - 1.0) < 1e-3
    assert float(ones.min()) >= 0.0 and float(ones.max()) < 1.8  # longest chord of a unit cube is sqrt(3)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand(A.domain_shape, device="cuda", generator=g)
    y = A(x)
    y -= ones
    assert float(y.max()) <= 1e-4                              # monotone: 0 <= x <= 1  =>  A x <= A 1
    y += ones
    lhs = (y * ones).sum(dtype=torch.float64)
    bp = A.T(ones)
    rhs = (x * bp).sum(dtype=torch.float64)
    assert abs(float(lhs / rhs) - 1) < 0.01                    # scaled adjoint
    del y, x
    out = torch.zeros_like(bp)
    assert torch.equal(A.T(ones, out=out), bp)                 # deterministic


def test_numpy_outputs_live_in_cached_pinned_buffers():
    """Arrays the host path creates itself (operator outputs, float32 copies of float64 inputs) are page-locked
    (tsp_host_alloc) so that their transfers overlap the kernels; a freed buffer is handed out again."""
    import ctypes
    import gc

    from tomosipo_b200 import _backend as B

    a = B.pinned_empty((256, 1024, 2))            # 2 MB: above the pinning threshold
    assert a.dtype == np.float32 and a.flags.c_contiguous and a.flags.writeable
    attr = ctypes.c_uint(0)
    err = torch.cuda.cudart().cudaHostGetFlags(a.ctypes.data) if hasattr(torch.cuda.cudart(), "cudaHostGetFlags") else None
    ptr = a.ctypes.data
    a[...] = 3.0
    del a
    gc.collect()
    b = B.pinned_empty((256, 1024, 2))
    assert b.ctypes.data == ptr                   # served from the cache
    assert B.pinned_empty((4, 4)).base is None    # small arrays stay ordinary numpy arrays
    vg = ts.volume(shape=96, size=1)
    pg = ts.cone(angles=48, shape=(96, 144), size=(1.875, 2.8125), src_orig_dist=4, src_det_dist=6)
    A = ts.operator(vg, pg)
    x64 = np.random.default_rng(0).random(A.domain_shape)              # float64, like the README loop
    with pytest.warns(UserWarning):
        y = A(x64)
    assert y.dtype == np.float32 and y.base is not None                # pinned-backed output
    assert rel_l2(y, oracle_of(A).fp(x64)) < TOL
    del attr, err


def test_cupy_link_on_real_device_memory(monkeypatch):
    """CuPy itself is absent from the image (a11).  The link only touches a handful of cupy names, so a stand-in
    module whose arrays live in REAL device memory (torch CUDA storage behind the cupy.ndarray surface) drives the
    whole CupyLink -> RawBuffer -> C ABI -> kernel path: device pointer, device id, current stream, same-device
    allocation of the output.  Also the cupy <-> torch compatibility the reference leaves as a TODO
    (tomosipo/links/cupy.py:66-72): projecting a CuPy volume into a torch CUDA tensor."""
    import importlib
    import sys
    import types

    class _Dev:
        def __init__(self, id_):
            self.id = id_
            self._ctx = None

        def __enter__(self):
            self._ctx = torch.cuda.device(self.id)
            self._ctx.__enter__()
            return self

        def __exit__(self, *a):
            return self._ctx.__exit__(*a)

        def __eq__(self, other):
            return isinstance(other, _Dev) and other.id == self.id

    class FakeArray:
        def __init__(self, t):
            self.t = t

        shape = property(lambda self: tuple(self.t.shape))
        dtype = property(lambda self: {torch.float32: np.dtype("float32"), torch.float64: np.dtype("float64")}[self.t.dtype])
        flags = property(lambda self: {"C_CONTIGUOUS": self.t.is_contiguous()})
        data = property(lambda self: types.SimpleNamespace(ptr=self.t.data_ptr()))
        device = property(lambda self: _Dev(self.t.device.index))

        def astype(self, dt):
            return FakeArray(self.t.to(torch.float32))

        def copy(self):
            return FakeArray(self.t.clone())

    fake = types.ModuleType("cupy")
    fake.ndarray = FakeArray
    fake.float32 = np.dtype("float32")
    mk = lambda f: (lambda shape, *a, dtype=None: FakeArray(f(tuple(shape), *a, dtype=torch.float32, device="cuda")))  # noqa: E731
    fake.zeros, fake.empty = mk(torch.zeros), mk(torch.empty)
    fake.full = lambda shape, value, dtype=None: FakeArray(torch.full(tuple(shape), float(value), device="cuda"))
    fake.ascontiguousarray = lambda a: FakeArray(a.t.contiguous())
    # CuPy keeps its own current stream (here: the default stream), whatever torch's current stream is
    fake.cuda = types.SimpleNamespace(
        get_current_stream=lambda: types.SimpleNamespace(ptr=torch.cuda.default_stream().cuda_stream))
    monkeypatch.setitem(sys.modules, "cupy", fake)
    sys.modules.pop("tomosipo_b200.links.cupy", None)
    n_backends = len(ts.links.base.backends)
    try:
        importlib.import_module("tomosipo_b200.links.cupy")
        vg = ts.volume(shape=(40, 48, 56), size=(0.8, 1, 1.1))
        pg = ts.cone(angles=30, shape=(40, 72), size=(1.6, 2.9), src_orig_dist=4, src_det_dist=6)
        A = ts.operator(vg, pg)
        g = torch.Generator(device="cuda").manual_seed(3)
        x = torch.rand(A.domain_shape, device="cuda", generator=g)
        y_t = A(x)
        y_c = A(FakeArray(x))                                   # CuPy in -> CuPy out, allocated by the link
        assert isinstance(y_c, FakeArray) and torch.equal(y_c.t, y_t)
        xb_c = A.T(y_c)
        assert isinstance(xb_c, FakeArray) and torch.equal(xb_c.t, A.T(y_t))
        out = torch.zeros_like(y_t)                             # CuPy volume -> torch projections (mixed links)
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):                           # ... with the output bound to another stream
            A(FakeArray(x), out=out)
        side.synchronize()
        assert torch.equal(out, y_t)
        assert rel_l2(y_c.t.cpu().numpy(), oracle_of(A).fp(x.cpu().numpy().astype(np.float64))) < TOL
    finally:
        del ts.links.base.backends[n_backends:]
        sys.modules.pop("tomosipo_b200.links.cupy", None)
