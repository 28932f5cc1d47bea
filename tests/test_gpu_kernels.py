"""Parity of the CUDA kernels (through the C ABI) against the fp64 oracle.

Tolerance (BASELINE.json north_star): relative L2 <= 1e-5 against the fp64
Joseph / voxel-driven reference.
"""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-5


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-300))


def _run(P, direction, vol, proj, additive=False):
    import torch
    from tomosipo_b200 import _backend as B

    dv = torch.from_numpy(np.ascontiguousarray(vol, dtype=np.float32)).cuda()
    dp = torch.from_numpy(np.ascontiguousarray(proj, dtype=np.float32)).cuda()
    s = torch.cuda.current_stream().cuda_stream
    P.project(direction, additive, dv.data_ptr(), dp.data_ptr(), B.MEM_DEVICE, 0, s)
    torch.cuda.synchronize()
    return dv.cpu().numpy(), dp.cpu().numpy()


def make(kind, vol_shape, window, det_shape, vectors, vss=1, dss=1):
    from tomosipo_b200 import _backend as B

    P = B.Projector(kind, vol_shape, window, det_shape, vectors, vss, dss)
    Q = O.OracleProjector(kind, vol_shape, [w[0] for w in window], [w[1] for w in window], det_shape, vectors, vss, dss)
    return P, Q


def cases():
    rng = np.random.default_rng(1)
    out = []
    ang = np.linspace(0, 2 * np.pi, 40, endpoint=False)
    # circular cone, cubic voxels, scaled-down cfg 3
    out.append(("cone", O.CONE_VEC, (64, 64, 64), [(-.5, .5)] * 3, (64, 96),
                O.cone_vectors(ang, 2.8125 / 96, 1.875 / 64, 4.0, 2.0)))
    # parallel, ragged sizes, anisotropic voxels, off-centre volume
    out.append(("par_aniso", O.PARALLEL_VEC, (37, 50, 45), [(-1.0, 1.4), (-2.0, 1.0), (-.7, .9)], (41, 67),
                O.parallel_vectors(np.linspace(0, np.pi, 31, endpoint=False), 0.05, 0.06)))
    # cone, ragged + anisotropic, source fairly close
    out.append(("cone_aniso", O.CONE_VEC, (33, 47, 29), [(-1.0, 1.2), (-1.5, 1.0), (-.8, .9)], (45, 53),
                O.cone_vectors(np.linspace(0, 2 * np.pi, 23, endpoint=False), 0.09, 0.08, 6.0, 3.0)))
    # random cone_vec (tilted detectors, arbitrary directions -> all three marching axes)
    v = np.zeros((24, 12))
    for i in range(24):
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        u = np.cross(d, rng.normal(size=3)); u /= np.linalg.norm(u)
        w = np.cross(d, u)
        v[i, 0:3] = -d * 9.0 + rng.normal(size=3) * 0.1
        v[i, 3:6] = d * 5.0 + rng.normal(size=3) * 0.1
        v[i, 6:9] = u * 0.11
        v[i, 9:12] = w * 0.13
    out.append(("cone_vec_random", O.CONE_VEC, (40, 36, 44), [(-1.1, 1.1), (-.9, .9), (-1, 1)], (48, 56), v))
    # random parallel_vec
    v = np.zeros((24, 12))
    for i in range(24):
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        u = np.cross(d, rng.normal(size=3)); u /= np.linalg.norm(u)
        w = np.cross(d, u)
        v[i, 0:3] = d
        v[i, 3:6] = rng.normal(size=3) * 0.2
        v[i, 6:9] = u * 0.06
        v[i, 9:12] = w * 0.05
    out.append(("par_vec_random", O.PARALLEL_VEC, (40, 36, 44), [(-1.1, 1.1), (-.9, .9), (-1, 1)], (64, 60), v))
    # slab (cfg 5 shape): 1 x N x N volume, single detector row
    out.append(("slab", O.PARALLEL_VEC, (1, 64, 64), [(-.5, .5), (-.5, .5), (-.5, -.5 + 1 / 64)], (1, 96),
                O.parallel_vectors(np.linspace(0, np.pi, 64, endpoint=False), 1.5 / 96, 1 / 64)
                + np.array([0, 0, 0, 0, 0, -.5 + .5 / 64, 0, 0, 0, 0, 0, 0])))
    return out


@pytest.mark.parametrize("case", cases(), ids=lambda c: c[0])
def test_fp_matches_oracle(case):
    name, kind, vs, win, ds, vec = case
    P, Q = make(kind, vs, win, ds, vec)
    rng = np.random.default_rng(0)
    x = rng.random(vs).astype(np.float32)
    _, y = _run(P, 0, x, np.zeros(Q.proj_shape, np.float32))
    ref = Q.fp(x.astype(np.float64))
    assert np.array_equal(P.marching_axes(), Q.marching_axes())
    assert rel_l2(y, ref) <= TOL, rel_l2(y, ref)


@pytest.mark.parametrize("case", cases(), ids=lambda c: c[0])
def test_bp_matches_oracle(case):
    name, kind, vs, win, ds, vec = case
    P, Q = make(kind, vs, win, ds, vec)
    rng = np.random.default_rng(0)
    y = rng.random(Q.proj_shape).astype(np.float32)
    x, _ = _run(P, 1, np.zeros(vs, np.float32), y)
    ref = Q.bp(y.astype(np.float64))
    assert rel_l2(x, ref) <= TOL, rel_l2(x, ref)


def test_additive_mode():
    name, kind, vs, win, ds, vec = cases()[0]
    P, Q = make(kind, vs, win, ds, vec)
    rng = np.random.default_rng(0)
    x = rng.random(vs).astype(np.float32)
    y0 = rng.random(Q.proj_shape).astype(np.float32)
    _, y = _run(P, 0, x, y0, additive=True)
    ref = Q.fp(x.astype(np.float64)) + y0
    assert rel_l2(y, ref) <= TOL
    x2, _ = _run(P, 1, x, y0, additive=True)
    ref = Q.bp(y0.astype(np.float64)) + x
    assert rel_l2(x2, ref) <= TOL


@pytest.mark.parametrize("direct", [False, True], ids=["staged", "direct"])
@pytest.mark.parametrize("kind", ["cone", "par_aniso"])
def test_supersampling(kind, direct):
    """VoxelSuperSampling / DetectorSuperSampling (reference tomosipo/astra.py:84-97, doc/topics/operator.rst:69-86):
    through the staged kernels on the refined geometry + pooling (default), and the direct kernels (TSP_SS_DIRECT=1);
    SET, ADD, and factors 2 and 3."""
    import os

    case = [c for c in cases() if c[0] == kind][0]
    name, k, vs, win, ds, vec = case
    vs = tuple(max(1, s // 2) for s in vs)
    old = os.environ.pop("TSP_SS_DIRECT", None)
    if direct:
        os.environ["TSP_SS_DIRECT"] = "1"
    try:
        for vss, dss in ((2, 2), (3, 1), (1, 3)):
            P, Q = make(k, vs, win, ds, vec, vss=vss, dss=dss)
            rng = np.random.default_rng(0)
            x = rng.random(vs).astype(np.float32)
            y = rng.random(Q.proj_shape).astype(np.float32)
            _, yy = _run(P, 0, x, np.zeros(Q.proj_shape, np.float32))
            xx, _ = _run(P, 1, np.zeros(vs, np.float32), y)
            assert rel_l2(yy, Q.fp(x.astype(np.float64))) <= TOL
            assert rel_l2(xx, Q.bp(y.astype(np.float64))) <= TOL
            if (vss, dss) == (2, 2):
                _, y2 = _run(P, 0, x, y.copy(), additive=True)
                x2, _ = _run(P, 1, x.copy(), y, additive=True)
                assert rel_l2(y2, Q.fp(x.astype(np.float64)) + y) <= TOL
                assert rel_l2(x2, Q.bp(y.astype(np.float64)) + x) <= TOL
                if not direct and kind == "cone":
                    assert P.info().fp_uses_tma == 1 and P.info().bp_uses_tma == 1
    finally:
        os.environ.pop("TSP_SS_DIRECT", None)
        if old is not None:
            os.environ["TSP_SS_DIRECT"] = old


def test_host_memory_path():
    from tomosipo_b200 import _backend as B

    name, kind, vs, win, ds, vec = cases()[1]
    P, Q = make(kind, vs, win, ds, vec)
    rng = np.random.default_rng(0)
    x = rng.random(vs).astype(np.float32)
    y = np.zeros(Q.proj_shape, np.float32)
    P.project(0, False, x.ctypes.data, y.ctypes.data, B.MEM_HOST, 0, 0)
    assert rel_l2(y, Q.fp(x.astype(np.float64))) <= TOL
    x2 = np.zeros(vs, np.float32)
    P.project(1, False, x2.ctypes.data, y.ctypes.data, B.MEM_HOST, 0, 0)
    assert rel_l2(x2, Q.bp(y.astype(np.float64))) <= TOL
